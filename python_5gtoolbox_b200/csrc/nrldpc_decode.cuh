// nrldpc_decode.cuh -- pieces shared by the table-driven (nrldpc_decode_qc.cu) and the compile-time
// specialised (nrldpc_decode_spec.cu) flooding min-sum decoder kernels.
#pragma once
#include "nrldpc_common.cuh"

namespace nrldpc {

struct DecArgs {
    const float *llr;
    int B, max_iter;
    float alpha, beta;
    int8_t *ck;
    uint32_t *info;
    uint8_t *status;
    int32_t *iters;
    int *work = nullptr;  // early-termination kernels: {ticket counter, CTAs done} slot of the dynamic codeblock queue, both zero
                          // at launch and zeroed again by the last CTA (or null: static stride)
};

// next slot of the current device's ring of queue slots (nrldpc_decode_spec.cu)
int decode_queue_slot(int **slot);

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
