#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_sch_chain.py -m gpu -x -q > gpurun_out/r2c_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c_pytest.log
tail -4 gpurun_out/r2c_pytest.log
timeout 300 python tools/tb_latency_probe.py > gpurun_out/r2c_tb.log 2>&1
echo "---- NRLDPC_TRACE=1" >> gpurun_out/r2c_tb.log
NRLDPC_TRACE=1 timeout 300 python tools/tb_latency_probe.py 2>&1 | awk '/trace/ {n++; if (n % 8 == 0) print; next} {print}' >> gpurun_out/r2c_tb.log
echo "---- NRLDPC_TRACE=2" >> gpurun_out/r2c_tb.log
NRLDPC_TRACE=2 timeout 300 python tools/tb_latency_probe.py 2>&1 | awk '/trace/ {n++; if (n % 8 == 0) print; next} {print}' >> gpurun_out/r2c_tb.log
cat gpurun_out/r2c_tb.log
