#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q --deselect tests/test_runs_unchanged.py > gpurun_out/r2h_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2h_pytest.log
tail -4 gpurun_out/r2h_pytest.log
echo "== specialised instances (multi-CTA per SM)"; python tools/bench_zc_sweep.py 1:128 1:72 1:40 2:128 2:72 2:40 1:144 1:120 1:64 2>&1 | tee gpurun_out/r2h_zc_spec.log
echo "== table-driven kernel (NRLDPC_NO_SPEC=1)"; NRLDPC_NO_SPEC=1 python tools/bench_zc_sweep.py 1:128 1:72 1:40 2:128 2:72 2:40 2>&1 | tee gpurun_out/r2h_zc_tab.log
python tools/mc_stage_times.py 2>&1 | tee gpurun_out/r2h_mc_stages.log
