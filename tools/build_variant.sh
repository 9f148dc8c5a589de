#!/bin/bash
# Build a variant of libnrldpc_b200.so with extra -D flags for the specialised decoder (kernel experiments).
# usage: tools/build_variant.sh NAME [-DFLAG=...]...   ->  build/variants/libnrldpc_NAME.so
set -e
cd "$(dirname "$0")/../python_5gtoolbox_b200/csrc"
name=$1; shift
mkdir -p ../../build/variants
make -s -j8 >/dev/null
nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC -Xptxas -v -fmad=false \
  --expt-relaxed-constexpr "$@" -c nrldpc_decode_spec_bg1_384.cu -o ../../build/variants/spec_$name.o 2> ../../build/variants/spec_$name.log
grep -E "registers|spill" ../../build/variants/spec_$name.log | sed -n 3,4p | tr '\n' ' '; echo
nvcc -shared -gencode arch=compute_100a,code=sm_100a -o ../../build/variants/libnrldpc_$name.so nrldpc_api.o nrldpc_tables.o \
  nrldpc_encode.o nrldpc_decode_qc.o nrldpc_decode_spec.o ../../build/variants/spec_$name.o nrldpc_generic.o nrldpc_util.o nrldpc_ratematch.o nrldpc_decode_spec_bg2_384.o nrldpc_decode_spec_bg1_352.o nrldpc_decode_spec_bg2_352.o nrldpc_decode_spec_bg1_320.o nrldpc_decode_spec_bg2_320.o nrldpc_decode_spec_bg1_288.o nrldpc_decode_spec_bg2_288.o nrldpc_decode_spec_bg1_208.o nrldpc_decode_spec_bg2_208.o nrldpc_decode_spec_bg1_176.o nrldpc_decode_spec_bg2_176.o -lcudart
