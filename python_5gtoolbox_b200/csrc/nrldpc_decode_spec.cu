// nrldpc_decode_spec.cu -- dispatch to the compile-time specialised decoder kernels (nrldpc_decode_spec.cuh).
#include <mutex>

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

// Ring of {ticket, done} slots per device for the dynamic codeblock queue of the early-termination kernels.  A slot is
// zero whenever no kernel is using it (the last CTA of a launch zeroes it), so a launch only draws the next index;
// two launches share a slot only if kSlots launches are in flight at once.
int decode_queue_slot(int **slot)
{
    constexpr int kSlots = 4096, kMaxDev = 64;
    static std::mutex mu;
    static int *ring[kMaxDev] = {};
    static unsigned next[kMaxDev] = {};
    int dev = 0;
    NRLDPC_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= kMaxDev) { *slot = nullptr; return NRLDPC_OK; }
    std::lock_guard<std::mutex> lock(mu);
    if (!ring[dev]) {
        NRLDPC_CUDA(cudaMalloc(reinterpret_cast<void **>(&ring[dev]), kSlots * 2 * sizeof(int)));
        NRLDPC_CUDA(cudaMemset(ring[dev], 0, kSlots * 2 * sizeof(int)));
    }
    *slot = ring[dev] + 2 * (next[dev]++ % kSlots);
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
