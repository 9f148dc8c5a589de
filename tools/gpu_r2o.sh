#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
echo "== new (packed in-loop syndrome)"; python tools/et_ab.py 2>&1 | tee gpurun_out/r2o_et_new.log
echo "== previous build (in-row syndrome)"; NRLDPC_SO=build/variants/libnrldpc_prev.so python tools/et_ab.py 2>&1 | tee gpurun_out/r2o_et_prev.log
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/r2o_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2o_pytest.log
