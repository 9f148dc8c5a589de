#!/bin/bash
# time every build/variants/libnrldpc_*.so on the bench workload (BG1 Zc=384, 16384 codeblocks, 10 iterations)
for so in build/variants/libnrldpc_*.so; do
  echo -n "$(basename $so): "
  NRLDPC_SO=$so python tools/profile_decode.py ${1:-16384} 4 | tail -2 | awk '{printf "%s ms %s Gbit/s | ", $3, $5} END {print ""}'
done
