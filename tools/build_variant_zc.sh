#!/bin/bash
# Variant of libnrldpc_b200.so in which the specialised decoder instances of the listed lifting sizes ("176" = BG1, "2:208" = BG2) are rebuilt with
# extra -D flags and every other specialised size is stubbed out (those fall back to the table-driven kernel, the library
# stays small): kernel experiments at other sizes than 384 (tools/build_variant.sh is the BG1 Zc=384 one).
# usage: tools/build_variant_zc.sh NAME "ZC ZC ..." [-DFLAG=...]...  ->  build/variants/libnrldpc_NAME.so
set -e
cd "$(dirname "$0")/../python_5gtoolbox_b200/csrc"
name=$1; zcs=$2; shift 2
V=../../build/variants
mkdir -p $V
ARCH="-gencode arch=compute_100a,code=sm_100a"
repl=""; keep=""
for item in $zcs; do   # "176" = BG1 Zc=176, "2:208" = BG2 Zc=208
  case $item in *:*) g=${item%%:*}; z=${item##*:};; *) g=1; z=$item;; esac
  nvcc -O3 -std=c++17 -lineinfo $ARCH -Xcompiler -fPIC -Xptxas -v -fmad=false --expt-relaxed-constexpr "$@" \
    -c nrldpc_decode_spec_bg${g}_$z.cu -o $V/${name}_bg${g}_$z.o 2> $V/${name}_bg${g}_$z.log &
  repl="$repl $V/${name}_bg${g}_$z.o"; keep="$keep|X\\($g, $z\\)"
done
grep -o 'X([12], [0-9]*)' nrldpc_decode_spec.cu | sort -u | grep -vE "^(${keep#|})$" | sed 's/X(\(.*\))/STUB(\1)/' | tr '\n' ' ' > $V/stub_list_$name.inc
cat > $V/stubs_$name.cu <<EOS
#include "nrldpc_decode.cuh"
namespace nrldpc {
#define STUB(BGN, ZC) \\
    int launch_decode_spec_##BGN##_##ZC(const DecArgs &, int, cudaStream_t) { return 1; } \\
    void decode_spec_geometry_##BGN##_##ZC(int *t, int *s) { if (t) *t = 0; if (s) *s = 0; }
#include "stub_list_$name.inc"
}
EOS
nvcc -O3 -std=c++17 $ARCH -Xcompiler -fPIC -I. -I$V -c $V/stubs_$name.cu -o $V/stubs_$name.o
wait
OBJS=$(ls *.o | grep -v '^nrldpc_decode_spec_bg')
nvcc -shared $ARCH -o $V/libnrldpc_$name.so $OBJS $repl $V/stubs_$name.o -lcudart
echo "$name: $(grep -h -E 'Used' $V/${name}_bg*.log | sed -E 's/.*Used ([0-9]+) registers.*/\1/' | sort -u | tr '\n' ' ')registers"
