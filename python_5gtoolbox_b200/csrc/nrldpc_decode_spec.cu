// nrldpc_decode_spec.cu -- dispatch to the compile-time specialised decoder kernels (nrldpc_decode_spec.cuh).
#include <mutex>

#include "nrldpc_decode.cuh"

namespace nrldpc {

#define NRLDPC_SPEC_LIST(X) X(1, 384) X(2, 384) X(1, 352) X(2, 352) X(1, 320) X(2, 320) X(1, 288) X(2, 288) X(1, 256) X(2, 256) X(1, 240) X(2, 240) X(1, 224) X(2, 224) X(1, 208) X(2, 208) X(1, 192) X(2, 192) X(1, 176) X(2, 176) X(1, 160) X(2, 160) X(1, 144) X(2, 144) X(1, 128) X(2, 128) X(1, 72) X(2, 72) X(1, 40) X(2, 40) X(1, 120) X(2, 120) X(1, 112) X(2, 112) X(1, 104) X(2, 104) X(1, 96) X(2, 96) X(1, 88) X(2, 88) X(1, 80) X(2, 80) X(1, 64) X(2, 64) X(1, 60) X(2, 60) X(1, 56) X(2, 56) X(1, 52) X(2, 52) X(1, 48) X(2, 48) X(1, 44) X(2, 44) X(1, 36) X(2, 36) X(1, 32) X(2, 32) X(1, 28) X(2, 28) X(1, 24) X(1, 30) X(2, 30) X(1, 26) X(1, 22) X(1, 20)

#define NRLDPC_DECLARE(BGN, ZC)                                                                 \
    int launch_decode_spec_##BGN##_##ZC(const DecArgs &a, int early_term, cudaStream_t s);      \
    void decode_spec_geometry_##BGN##_##ZC(int *threads, int *smem);
NRLDPC_SPEC_LIST(NRLDPC_DECLARE)

int launch_decode_spec(int bgn, int Zc, const DecArgs &a, int early_term, cudaStream_t s, bool *handled)
{
    *handled = true;
    int rc = 1;  // 1 = no specialised kernel for this request (kNoVariant of nrldpc_decode_spec.cuh)
#define NRLDPC_CASE(BGN, ZC) if (bgn == BGN && Zc == ZC) rc = launch_decode_spec_##BGN##_##ZC(a, early_term, s);
    NRLDPC_SPEC_LIST(NRLDPC_CASE)
#undef NRLDPC_CASE
    if (rc != 1) return rc;
    *handled = false;
    return NRLDPC_OK;
}

// Ring of ticket counters per device for the dynamic codeblock queue of the early-termination kernels.  A counter is
// never reset: a launch draws exactly `tickets` tickets, so the host keeps the value every counter will have when the
// next launch that draws it starts (launches that share a counter are serialised only if kSlots launches are in flight
// at once).  Nothing here may run inside a stream capture (allocation, and a graph replay would reuse work_base):
// capturing streams get no counter and the kernel uses its static stride.
int decode_queue_slot(cudaStream_t s, unsigned tickets, unsigned **work, unsigned *work_base)
{
    constexpr int kSlots = 4096, kMaxDev = 64;
    static std::mutex mu;
    static unsigned *ring[kMaxDev] = {};
    static unsigned *base[kMaxDev] = {};
    static unsigned next[kMaxDev] = {};
    *work = nullptr;
    *work_base = 0;
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(s, &cap) != cudaSuccess) { cudaGetLastError(); return NRLDPC_OK; }
    if (cap != cudaStreamCaptureStatusNone) return NRLDPC_OK;
    int dev = 0;
    NRLDPC_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= kMaxDev) return NRLDPC_OK;
    std::lock_guard<std::mutex> lock(mu);
    if (!ring[dev]) {
        unsigned *p = nullptr;
        NRLDPC_CUDA(cudaMalloc(reinterpret_cast<void **>(&p), kSlots * sizeof(unsigned)));
        NRLDPC_CUDA(cudaMemset(p, 0, kSlots * sizeof(unsigned)));
        ring[dev] = p;
        base[dev] = new unsigned[kSlots]();
    }
    const unsigned i = next[dev]++ % kSlots;
    *work = ring[dev] + i;
    *work_base = base[dev][i];
    base[dev][i] += tickets;  // wraps like the device counter
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
