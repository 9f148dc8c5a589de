#!/bin/bash
# Build a variant of libnrldpc_b200.so with extra -D flags for the specialised BG1 Zc=384 decoder (kernel experiments).
# usage: tools/build_variant.sh NAME [-DFLAG=...]...   ->  build/variants/libnrldpc_NAME.so
# The variant library carries only the BG1 Zc=384 instance (every other specialised size is stubbed out, so those sizes
# fall back to the table-driven kernel and the library stays small); NRLDPC_NO_MAKE=1 skips the rebuild of the main
# library's objects.  Load it with NRLDPC_SO=build/variants/libnrldpc_NAME.so (tools/profile_decode.py, tools/et_ab.py).
set -e
cd "$(dirname "$0")/../python_5gtoolbox_b200/csrc"
name=$1; shift
V=../../build/variants
mkdir -p $V
[ -n "$NRLDPC_NO_MAKE" ] || make -s -j8 >/dev/null
ARCH="-gencode arch=compute_100a,code=sm_100a"
# one stub per entry of NRLDPC_SPEC_LIST except (1, 384): "no such variant" (kNoVariant) -> table-driven kernel
grep -o 'X([12], [0-9]*)' nrldpc_decode_spec.cu | sort -u | grep -v 'X(1, 384)' | sed 's/X(\(.*\))/STUB(\1)/' | tr '\n' ' ' > $V/stub_list.inc
cat > $V/stubs.cu <<'EOS'
#include "nrldpc_decode.cuh"
namespace nrldpc {
#define STUB(BGN, ZC) \
    int launch_decode_spec_##BGN##_##ZC(const DecArgs &, int, cudaStream_t) { return 1; } \
    void decode_spec_geometry_##BGN##_##ZC(int *t, int *s) { if (t) *t = 0; if (s) *s = 0; }
#include "stub_list.inc"
}
EOS
nvcc -O3 -std=c++17 $ARCH -Xcompiler -fPIC -I. -I$V -c $V/stubs.cu -o $V/stubs.o
nvcc -O3 -std=c++17 -lineinfo $ARCH -Xcompiler -fPIC -Xptxas -v -fmad=false \
  --expt-relaxed-constexpr "$@" -c nrldpc_decode_spec_bg1_384.cu -o $V/spec_$name.o 2> $V/spec_$name.log
echo "$name: $(grep -E 'Used' $V/spec_$name.log | sed -E 's/.*Used ([0-9]+) registers.*/\1/' | tr '\n' ' ')registers"
OBJS=$(ls *.o | grep -v '^nrldpc_decode_spec_bg')
nvcc -shared $ARCH -o $V/libnrldpc_$name.so $OBJS $V/spec_$name.o $V/stubs.o -lcudart
