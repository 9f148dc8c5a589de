#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2d_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2d_pytest.log
tail -6 gpurun_out/r2d_pytest.log
timeout 300 python tools/pdsch_slot_bench.py > gpurun_out/r2d_pdsch.log 2>&1; cat gpurun_out/r2d_pdsch.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2d_bench.json 2> gpurun_out/r2d_bench.err; echo "bench rc=$?"
tail -3 gpurun_out/r2d_bench.err
timeout 600 python bench.py --workload mc --steps 2 --warmup 1 > gpurun_out/r2d_mc.json 2> gpurun_out/r2d_mc.err; echo "mc rc=$?"; cat gpurun_out/r2d_mc.json; tail -3 gpurun_out/r2d_mc.err
timeout 400 python bench.py --impl reference --steps 10 --warmup 3 > gpurun_out/r2d_ref.json 2> gpurun_out/r2d_ref.err; echo "ref rc=$?"; cat gpurun_out/r2d_ref.json
