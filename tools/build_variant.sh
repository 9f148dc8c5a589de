#!/bin/bash
# Build a variant of libnrldpc_b200.so with extra -D flags for the specialised BG1 Zc=384 decoder (kernel experiments).
# usage: tools/build_variant.sh NAME [-DFLAG=...]...   ->  build/variants/libnrldpc_NAME.so
# The variant library carries only the BG1 Zc=384 instance (the other specialised sizes are stubbed out so that the
# library stays ~10 MB); NRLDPC_NO_MAKE=1 skips the rebuild of the main library's objects.
set -e
cd "$(dirname "$0")/../python_5gtoolbox_b200/csrc"
name=$1; shift
V=../../build/variants
mkdir -p $V
[ -n "$NRLDPC_NO_MAKE" ] || make -s -j8 >/dev/null
ARCH="-gencode arch=compute_100a,code=sm_100a"
if [ ! -f $V/stubs.o ]; then
  cat > $V/stubs.cu <<'EOS'
#include "nrldpc_decode.cuh"
namespace nrldpc {
#define STUB(BGN, ZC) \
    int launch_decode_spec_##BGN##_##ZC(const DecArgs &, int, cudaStream_t) { return NRLDPC_EINVAL; } \
    void decode_spec_geometry_##BGN##_##ZC(int *t, int *s) { if (t) *t = 0; if (s) *s = 0; }
STUB(2, 384) STUB(1, 352) STUB(2, 352) STUB(1, 320) STUB(2, 320) STUB(1, 288) STUB(2, 288) STUB(1, 208) STUB(2, 208) STUB(1, 176) STUB(2, 176)
}
EOS
  nvcc -O3 -std=c++17 $ARCH -Xcompiler -fPIC -I. -c $V/stubs.cu -o $V/stubs.o
fi
nvcc -O3 -std=c++17 -lineinfo $ARCH -Xcompiler -fPIC -Xptxas -v -fmad=false \
  --expt-relaxed-constexpr "$@" -c nrldpc_decode_spec_bg1_384.cu -o $V/spec_$name.o 2> $V/spec_$name.log
echo "$name: $(grep -E 'Used' $V/spec_$name.log | sed -E 's/.*Used ([0-9]+) registers.*/\1/' | tr '\n' ' ')registers (ET0/B00 ET1/B00 ET0/B01 ET1/B01)"
nvcc -shared $ARCH -o $V/libnrldpc_$name.so nrldpc_api.o nrldpc_tables.o nrldpc_encode.o nrldpc_decode_qc.o \
  nrldpc_decode_spec.o $V/spec_$name.o $V/stubs.o nrldpc_generic.o nrldpc_bf_qc.o nrldpc_util.o nrldpc_ratematch.o -lcudart
