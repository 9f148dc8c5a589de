#!/bin/bash
# A/B of the sum-product kernel against saved builds (build/variants/libnrldpc_*.so) + its tests + an ncu capture
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
for V in build/variants/libnrldpc_*.so; do
  [ -f "$V" ] && { echo "== $V"; NRLDPC_SO=$PWD/$V python tools/bench_bp.py 2>&1 | tail -2; }
done
echo "== shipped"; python tools/bench_bp.py 2>&1 | tee gpurun_out/r2_bp_bench.log | tail -2
timeout 900 python -m pytest tests -m gpu -x -q -k "bp" 2>&1 | tail -3
python tools/profile_bp.py 2>&1 | tail -1
ncu --clock-control none --set full --import-source on --kernel-name regex:bp_qc_kernel --launch-skip 1 --launch-count 1 -o gpurun_out/prof_r2_bp -f python tools/profile_bp.py > gpurun_out/r2_bp_ncu.log 2>&1
ls -la gpurun_out/prof_r2_bp.ncu-rep; tail -2 gpurun_out/r2_bp_ncu.log
