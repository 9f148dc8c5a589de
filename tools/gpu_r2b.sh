#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_sch_chain.py -m gpu -x -q > gpurun_out/r2b_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2b_pytest.log
tail -5 gpurun_out/r2b_pytest.log
timeout 300 python tools/tb_latency_probe.py > gpurun_out/r2b_tb.log 2>&1
NRLDPC_TRACE=1 timeout 300 python tools/tb_latency_probe.py 2>&1 | grep -m 12 "trace" >> gpurun_out/r2b_tb.log
cat gpurun_out/r2b_tb.log
for cfg in "3 1 8" "7 1 8" "7 0 8" "11 1 8" "15 1 8" "7 1 32" "15 1 32" "15 0 32"; do
  set -- $cfg
  NRLDPC_COPY_THREADS=$1 NRLDPC_COPY_NT=$2 NRLDPC_STAGE_MB=$3 timeout 300 python tools/pageable_probe.py 4096 2>&1 | tail -1
done | tee gpurun_out/r2b_pageable.log
