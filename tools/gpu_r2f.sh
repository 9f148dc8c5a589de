#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/r2f_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2f_pytest.log
tail -4 gpurun_out/r2f_pytest.log
for rep in 1 2; do
  echo "== r1 build"; NRLDPC_SO=build/variants/libnrldpc_r1.so python tools/profile_decode.py 65536 3 | tail -2
  echo "== current build"; python tools/profile_decode.py 65536 3 | tail -2
done 2>&1 | tee gpurun_out/r2f_ab.log
timeout 600 python tools/bench_bp.py 2>&1 | tee gpurun_out/r2f_bp.log
