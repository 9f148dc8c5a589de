#!/bin/bash
# same-box A/B of specialised-decoder variants (build/variants/libnrldpc_*.so, tools/build_variant_zc.sh) at the given sizes
# MODES: "fixed et" (default) or "et" (the sizes below 144 have early-termination instances only)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
SIZES="${SIZES:-1:176 1:144}"
MODES="${MODES:-fixed et}"
for V in shipped build/variants/libnrldpc_*.so shipped; do
  echo "== $V"
  if [ "$V" = shipped ]; then unset NRLDPC_SO; else export NRLDPC_SO=$PWD/$V; fi
  for m in $MODES; do
    if [ $m = et ]; then python tools/bench_zc_sweep.py --et $SIZES 2>&1 | grep "^BG" | cut -c1-175
    else python tools/bench_zc_sweep.py $SIZES 2>&1 | grep "^BG" | cut -c1-150; fi
  done
done 2>&1 | tee gpurun_out/zc_ab.log
