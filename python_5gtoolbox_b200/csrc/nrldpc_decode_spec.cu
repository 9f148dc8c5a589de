// nrldpc_decode_spec.cu -- dispatch to the compile-time specialised decoder kernels (nrldpc_decode_spec.cuh).
#include "nrldpc_decode.cuh"

namespace nrldpc {

#define NRLDPC_SPEC_LIST(X) X(1, 384) X(2, 384) X(1, 352) X(2, 352) X(1, 320) X(2, 320) X(1, 288) X(2, 288) X(1, 256) X(2, 256) X(1, 240) X(2, 240) X(1, 224) X(2, 224) X(1, 208) X(2, 208) X(1, 192) X(2, 192) X(1, 176) X(2, 176) X(1, 160) X(2, 160) X(1, 144) X(2, 144)

#define NRLDPC_DECLARE(BGN, ZC)                                                                 \
    int launch_decode_spec_##BGN##_##ZC(const DecArgs &a, int early_term, cudaStream_t s);      \
    void decode_spec_geometry_##BGN##_##ZC(int *threads, int *smem);
NRLDPC_SPEC_LIST(NRLDPC_DECLARE)

int launch_decode_spec(int bgn, int Zc, const DecArgs &a, int early_term, cudaStream_t s, bool *handled)
{
    *handled = true;
#define NRLDPC_CASE(BGN, ZC) if (bgn == BGN && Zc == ZC) return launch_decode_spec_##BGN##_##ZC(a, early_term, s);
    NRLDPC_SPEC_LIST(NRLDPC_CASE)
#undef NRLDPC_CASE
    *handled = false;
    return NRLDPC_OK;
}

bool decode_spec_geometry(int bgn, int Zc, int *threads, int *smem)
{
#define NRLDPC_CASE(BGN, ZC) if (bgn == BGN && Zc == ZC) { decode_spec_geometry_##BGN##_##ZC(threads, smem); return true; }
    NRLDPC_SPEC_LIST(NRLDPC_CASE)
#undef NRLDPC_CASE
    return false;
}

}  // namespace nrldpc
