// nrldpc_decode.cuh -- pieces shared by the table-driven (nrldpc_decode_qc.cu) and the compile-time
// specialised (nrldpc_decode_spec.cu) flooding min-sum decoder kernels.
#pragma once
#include "nrldpc_common.cuh"
#include "nrldpc_raterecover.cuh"

namespace nrldpc {

struct DecArgs {
    const float *llr;
    int B, max_iter;
    float alpha, beta;
    int8_t *ck;
    uint32_t *info;
    uint8_t *status;
    int32_t *iters;
    // early-termination kernels: ticket counter of the dynamic codeblock queue (or null: static stride).  The counter is
    // never reset: a launch draws exactly B tickets (one per decoded codeblock), so the host knows the value it has when
    // the launch starts (work_base) and the kernel's ticket is atomicAdd(work, 1) - work_base.
    unsigned *work = nullptr;
    unsigned work_base = 0;
    RrArgs rr;  // rr.src != nullptr: the LLR load is the rate recovery of a transport block (llr is then unused)
};

// Ticket counter for a launch that will draw `tickets` tickets on stream s, from the current device's ring of counters
// (nrldpc_decode_spec.cu).  *work = nullptr when the stream is being captured into a CUDA graph (a replay would reuse
// work_base): the caller then falls back to the static stride.
int decode_queue_slot(cudaStream_t s, unsigned tickets, unsigned **work, unsigned *work_base);

constexpr uint32_t kInfBits = 0x7f800000u;

// d = min(|a|,|b|) with sign(a) xor sign(b): running "sign product * first minimum" of a check row.
__device__ __forceinline__ float min_xorsign_abs(float a, float b)
{
    float d;
    asm("min.xorsign.abs.f32 %0, %1, %2;" : "=f"(d) : "f"(a), "f"(b));
    return d;
}

// Check-to-variable message decoded from a row record: magnitude m.y on the argmin edge (isidx),
// m.x elsewhere; `signword` carries the edge's sign in bit 31.
__device__ __forceinline__ float record_lr(float2 m, bool isidx, uint32_t signword)
{
    const float mag = isidx ? m.y : m.x;
    return __uint_as_float(__float_as_uint(mag) ^ (signword & 0x80000000u));
}

// Specialised kernels (nrldpc_decode_spec.cu).  `*handled` = false when (bgn, Zc) has no specialisation.
int launch_decode_spec(int bgn, int Zc, const DecArgs &a, int early_term, cudaStream_t s, bool *handled);
bool decode_spec_geometry(int bgn, int Zc, int *threads, int *smem);

}  // namespace nrldpc
