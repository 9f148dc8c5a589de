#!/bin/bash
# round-2 GPU check A: parity suite, transport-block latency, headline bench
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version --format=csv,noheader > gpurun_out/r2a_gpu.txt 2>&1
nproc >> gpurun_out/r2a_gpu.txt; free -g | head -2 >> gpurun_out/r2a_gpu.txt
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2a_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2a_pytest.log
tail -15 gpurun_out/r2a_pytest.log
timeout 300 python tools/pdsch_slot_bench.py > gpurun_out/r2a_pdsch.log 2>&1; echo "rc=$?" >> gpurun_out/r2a_pdsch.log
NRLDPC_SOFT_D2H=1 timeout 300 python tools/pdsch_slot_bench.py > gpurun_out/r2a_pdsch_d2h.log 2>&1
cat gpurun_out/r2a_pdsch.log gpurun_out/r2a_pdsch_d2h.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err; echo "bench rc=$?"
cat gpurun_out/r2a_bench.json | cut -c1-1500
tail -5 gpurun_out/r2a_bench.err
