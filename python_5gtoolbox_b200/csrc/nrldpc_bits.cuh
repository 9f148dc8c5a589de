// nrldpc_bits.cuh -- Zc-bit circular vectors packed in 32-bit words (encoder and bit-flipping decoder).
#pragma once
#include <stdint.h>

namespace nrldpc {

// 32 bits of the Zc-bit circular vector v starting at bit position `pos` (0 <= pos < Zc), i.e.
// result bit t = v[(pos + t) mod Zc].  v has W = ceil(Zc/32) words plus one zero pad word, and the
// bits at positions >= Zc of the last word are zero.
__device__ __forceinline__ uint32_t circ_window(const uint32_t *v, int pos, int Zc)
{
    if (Zc >= 32) {
        const int q = pos >> 5, sh = pos & 31;
        uint32_t out = __funnelshift_r(v[q], v[q + 1], sh);
        const int n1 = Zc - pos;  // bits available before the wrap
        if (n1 < 32) out = (out & ((1u << n1) - 1u)) | (v[0] << n1);
        return out;
    }
    // Zc < 32: replicate the vector until it covers pos + 32 bits
    unsigned long long rep = v[0];
    for (int len = Zc; len < 64; len <<= 1) rep |= rep << len;
    return (uint32_t)(rep >> pos);
}

// word w of (circulant block with shift P) @ v :  out[r] = v[(r + P) mod Zc], r = 32w .. 32w+31
__device__ __forceinline__ uint32_t rot_word(const uint32_t *v, int P, int w, int Zc)
{
    int pos = 32 * w + P;
    if (pos >= Zc) pos -= Zc;  // 32w < Zc and P < Zc
    uint32_t out = circ_window(v, pos, Zc);
    const int nv = Zc - 32 * w;  // valid bits of this word
    if (nv < 32) out &= (1u << nv) - 1u;
    return out;
}

}  // namespace nrldpc
