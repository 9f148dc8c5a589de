// nrldpc_util.cu -- device-side Monte-Carlo helpers: Philox bits, BPSK/AWGN LLRs, error counters.
// Replaces the per-codeblock host loop of for_test_5g_ldpc_encoder (py5gphy/ldpc/nr_ldpc_decode.py:247-257)
// and the np.array_equal bookkeeping of scripts/internal/sim_ldpc_internal.py:61-62.
#include <algorithm>

#include "nrldpc_common.cuh"

namespace nrldpc {

namespace {

// Philox4x32-10 (Salmon et al., SC'11): counter-based, so any (seed, offset) slice is reproducible
// on any number of GPUs.
__device__ __forceinline__ uint4 philox4x32_10(uint4 ctr, uint2 key)
{
    constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int round = 0; round < 10; ++round) {
        const uint32_t hi0 = __umulhi(M0, ctr.x), lo0 = M0 * ctr.x;
        const uint32_t hi1 = __umulhi(M1, ctr.z), lo1 = M1 * ctr.z;
        ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
        key.x += W0;
        key.y += W1;
    }
    return ctr;
}

__device__ __forceinline__ float sqrt_approx(float x)
{
    float y;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

__device__ __forceinline__ uint4 philox_at(unsigned long long seed, unsigned long long block)
{
    return philox4x32_10(make_uint4((uint32_t)block, (uint32_t)(block >> 32), 0x4c445043u, 0u),
                         make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
}

// rows x cols outputs; row j uses the Philox counters (first_id + j*id_stride) * bpr + blk (+ ctr0), so a
// codeblock's random stream depends only on its global id, not on how the batch is sharded.
__global__ void random_bits_kernel(int8_t *bits, long long rows, long long cols, unsigned long long seed,
                                   unsigned long long ctr0, long long first_id, long long id_stride)
{
    const long long bpr = (cols + 127) / 128, nblk = rows * bpr;  // one Philox block = 128 bits
    const bool vec8 = (cols % 8 == 0) && (reinterpret_cast<uintptr_t>(bits) % 8 == 0);
    // (row, blk) of block t advance incrementally: one 64-bit division per thread instead of one per block
    const long long stride = (long long)gridDim.x * blockDim.x, t0 = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const long long srow = stride / bpr, sblk = stride - srow * bpr;
    long long row = t0 / bpr, blk = t0 - row * bpr;
    for (long long t = t0; t < nblk; t += stride, row += srow, blk += sblk) {
        if (blk >= bpr) { blk -= bpr; ++row; }
        const uint4 x = philox_at(seed, ctr0 + (unsigned long long)((first_id + row * id_stride) * bpr + blk));
        const uint32_t w[4] = {x.x, x.y, x.z, x.w};
        const long long base = blk * 128;
        int8_t *out = bits + row * cols;
        if (vec8 && base + 128 <= cols) {
            // 8 bits -> 8 bytes per store: nibble * 0x00204081 & 0x01010101 spreads 4 bits over 4 bytes
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                const uint32_t byte = (w[i >> 2] >> (8 * (i & 3))) & 0xffu;
                const uint32_t lo = ((byte & 0xfu) * 0x00204081u) & 0x01010101u, hi = ((byte >> 4) * 0x00204081u) & 0x01010101u;
                *reinterpret_cast<uint2 *>(out + base + 8 * i) = make_uint2(lo, hi);
            }
            continue;
        }
#pragma unroll
        for (int i = 0; i < 128; ++i)
            if (base + i < cols) out[base + i] = (int8_t)((w[i >> 5] >> (i & 31)) & 1u);
    }
}

// Bit-packed twin: a Philox block IS four packed words (bit i of the block = bit i % 32 of word i / 32, exactly the bit
// random_bits_kernel stores as a byte).  One thread per group of four words of a row; the words beyond `cols` are zero.
__global__ void random_bits_packed_kernel(uint32_t *words, long long rows, long long cols, long long row_words,
                                          unsigned long long seed, long long first_id, long long id_stride)
{
    const long long bpr = (cols + 127) / 128, gpr = (row_words + 3) / 4, ngrp = rows * gpr;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < ngrp; t += stride) {
        const long long row = t / gpr, blk = t - row * gpr;
        uint32_t w[4] = {0u, 0u, 0u, 0u};
        if (blk < bpr) {
            const uint4 x = philox_at(seed, (unsigned long long)((first_id + row * id_stride) * bpr + blk));
            w[0] = x.x; w[1] = x.y; w[2] = x.z; w[3] = x.w;
        }
        uint32_t *out = words + row * row_words + blk * 4;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const long long first_bit = blk * 128 + 32 * q, left = cols - first_bit;  // valid bits of this word
            if (blk * 4 + q < row_words) out[q] = left >= 32 ? w[q] : (left > 0 ? w[q] & ((1u << left) - 1u) : 0u);
        }
    }
}

// A CTA walks whole rows (one long row: stretches of kAwgnStretch Philox blocks), so that the 64-bit counter base and the row
// pointers are formed once per stretch and everything per Philox block is 32-bit arithmetic.  The counter of block blk of row j
// is ctr0 + (first_id + j * id_stride) * bpr + blk, whatever the launch geometry.
constexpr unsigned kAwgnStretch = 8192;

// PACKED: dn is bit-packed (`dnw`, row_words words per row, no fillers) instead of one int8 per bit.
template <bool PACKED>
__global__ void __launch_bounds__(256, 8)
awgn_llr_kernel(const int8_t *__restrict__ dn, const uint32_t *__restrict__ dnw, long long row_words, long long rows,
                long long cols, float sigma, float scale,
                unsigned long long seed, unsigned long long ctr0, long long first_id, long long id_stride,
                float *__restrict__ llr)
{
    const unsigned long long bpr = (unsigned long long)(cols + 3) / 4;  // one Philox block = 4 normals
    const bool vec4 = (cols % 4 == 0) && (PACKED || reinterpret_cast<uintptr_t>(dn) % 4 == 0) && (reinterpret_cast<uintptr_t>(llr) % 16 == 0);
    const unsigned long long spr = (bpr + kAwgnStretch - 1) / kAwgnStretch, nstretch = (unsigned long long)rows * spr;
    const uint2 key = make_uint2((uint32_t)(seed ^ 0x9E3779B97F4A7C15ull), (uint32_t)((seed ^ 0x9E3779B97F4A7C15ull) >> 32));
    const float sig_scale = sigma * scale;  // LLR = scale * (en + sigma n) = scale * en + (sigma scale) * n
    for (unsigned long long st = blockIdx.x; st < nstretch; st += gridDim.x) {
        const unsigned long long row = spr == 1 ? st : st / spr, s0 = (st - row * spr) * kAwgnStretch;
        const unsigned nb = (unsigned)min((unsigned long long)kAwgnStretch, bpr - s0);
        const unsigned long long cbase = ctr0 + (unsigned long long)(first_id + (long long)row * id_stride) * bpr + s0;
        const int8_t *d = PACKED ? nullptr : dn + row * cols + s0 * 4;
        const uint32_t *dw = PACKED ? dnw + row * row_words : nullptr;
        const unsigned p0 = (unsigned)(s0 * 4);  // bit position of the stretch inside the packed row
        float *o = llr + row * cols + s0 * 4;
        const long long left = cols - (long long)s0 * 4;  // values of this row from the stretch on
        for (unsigned b = threadIdx.x; b < nb; b += blockDim.x) {
            const unsigned long long c = cbase + b;
            const uint4 x = philox4x32_10(make_uint4((uint32_t)c, (uint32_t)(c >> 32), 0x4c445043u, 0u), key);
            // Box-Muller on two uniform pairs, on the special-function unit: lg2 / sqrt / sin / cos are one MUFU instruction
            // each (relative error ~2^-21 on the radius, absolute ~2^-21 on the angle part: far below the Monte-Carlo
            // resolution of a BLER point)
            const float u0 = ((x.x >> 8) + 0.5f) * (1.0f / 16777216.0f), u1 = ((x.y >> 8) + 0.5f) * (1.0f / 16777216.0f);
            const float u2 = ((x.z >> 8) + 0.5f) * (1.0f / 16777216.0f), u3 = ((x.w >> 8) + 0.5f) * (1.0f / 16777216.0f);
            const float r0 = sqrt_approx(-2.0f * __logf(u0)), r1 = sqrt_approx(-2.0f * __logf(u2));
            float s0f, c0f, s1f, c1f;
            __sincosf(6.28318530717958647692f * u1, &s0f, &c0f);
            __sincosf(6.28318530717958647692f * u3, &s1f, &c1f);
            const float n[4] = {r0 * c0f, r0 * s0f, r1 * c1f, r1 * s1f};
            if (vec4) {
                // :252-257  en = 1 - 2 dn ; fn = en + N(0, sigma) ; LLR = 2 fn / sigma^2 ; a -1 filler is sent as LLR 0
                if constexpr (PACKED) {
                    const unsigned p = p0 + 4u * b;  // a multiple of 4: the nibble never straddles two words
                    const uint32_t w = dw[p >> 5] >> (p & 31u);
                    float4 v;
                    v.x = __fmaf_rn(sig_scale, n[0], __uint_as_float(__float_as_uint(scale) ^ ((w << 31) & 0x80000000u)));
                    v.y = __fmaf_rn(sig_scale, n[1], __uint_as_float(__float_as_uint(scale) ^ ((w << 30) & 0x80000000u)));
                    v.z = __fmaf_rn(sig_scale, n[2], __uint_as_float(__float_as_uint(scale) ^ ((w << 29) & 0x80000000u)));
                    v.w = __fmaf_rn(sig_scale, n[3], __uint_as_float(__float_as_uint(scale) ^ ((w << 28) & 0x80000000u)));
                    *reinterpret_cast<float4 *>(o + 4 * b) = v;
                    continue;
                }
                const uint32_t d4 = *reinterpret_cast<const uint32_t *>(d + 4 * b);
                float4 v;
                float *vp = &v.x;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const uint32_t byte = (d4 >> (8 * i)) & 0xffu;
                    // scale * (+-1) by flipping the sign bit with the data bit; fused multiply-add for the noise term
                    const float en = __uint_as_float(__float_as_uint(scale) ^ (byte << 31));
                    vp[i] = (byte & 0x80u) ? 0.0f : __fmaf_rn(sig_scale, n[i], en);
                }
                *reinterpret_cast<float4 *>(o + 4 * b) = v;
            } else {
#pragma unroll
                for (int i = 0; i < 4; ++i)
                    if ((long long)4 * b + i < left) {
                        int dv;
                        if constexpr (PACKED) { const unsigned p = p0 + 4u * b + i; dv = (int)((dw[p >> 5] >> (p & 31u)) & 1u); }
                        else dv = d[4 * b + i];
                        const float en = dv ? -scale : scale;
                        o[4 * b + i] = dv < 0 ? 0.0f : __fmaf_rn(sig_scale, n[i], en);
                    }
            }
        }
    }
}

__global__ void count_errors_kernel(const int8_t *__restrict__ ref, long long ref_stride, const int8_t *__restrict__ got,
                                    long long got_stride, int B, int K, const int32_t *__restrict__ iters,
                                    unsigned long long *__restrict__ counters)
{
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    const int nwarps = (gridDim.x * blockDim.x) >> 5;
    unsigned long long blk = 0, bit = 0, its = 0, n = 0;
    for (int b = warp; b < B; b += nwarps) {
        int diff = 0;
        for (int k = lane; k < K; k += 32) diff += ref[b * ref_stride + k] != got[b * got_stride + k];
        diff = __reduce_add_sync(0xffffffffu, diff);
        if (lane == 0) {
            ++n; bit += diff; blk += diff != 0;
            if (iters) its += iters[b];
        }
    }
    if (lane == 0 && n) {
        atomicAdd(counters + 0, n);
        atomicAdd(counters + 1, blk);
        atomicAdd(counters + 2, bit);
        atomicAdd(counters + 3, its);
    }
}


// nr_crc_encode / nr_crc_decode long division (py5gphy/crc/crc.py:28-33, :72-77), one thread per block
// of bits.  rem holds the L-bit remainder register, MSB = remainder[0].  mode 0: append the CRC to
// `in` [B,A] -> out [B,A+L].  mode 1: `in` is [B,A+L]; err[b] = remainder != 0.
__global__ void crc_kernel(const int8_t *__restrict__ in, int B, int A, int L, uint32_t poly, int mode,
                           int8_t *__restrict__ out, uint8_t *__restrict__ err)
{
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const int len_in = mode ? A + L : A;
    const int8_t *x = in + (size_t)b * len_in;
    const uint32_t mask = (L == 32) ? 0xffffffffu : ((1u << L) - 1u);
    uint32_t rem = 0;
    for (int k = 0; k < L; ++k) rem = (rem << 1) | (uint32_t)((k < len_in) ? (x[k] & 1) : 0);
    for (int idx = 0; idx < A; ++idx) {
        const uint32_t first = (rem >> (L - 1)) & 1u;
        const int k = idx + L;
        rem = ((rem << 1) & mask) | (uint32_t)((k < len_in) ? (x[k] & 1) : 0);
        if (first) rem ^= poly;
    }
    if (mode == 0) {
        int8_t *y = out + (size_t)b * (A + L);
        for (int k = 0; k < A; ++k) y[k] = x[k];
        for (int k = 0; k < L; ++k) y[A + k] = (int8_t)((rem >> (L - 1 - k)) & 1u);
    } else {
        err[b] = rem != 0;
    }
}

// ---- the same CRC for long blocks (transport blocks of up to ~10^6 bits): one CTA per block of bits.
// The CRC is linear over GF(2): with the message cut into chunks, M(x) = sum_t chunk_t(x) x^{s_t}
// (s_t = message bits after chunk t), so  M(x) x^L mod P = sum_t (r_t * x^{s_t}) mod P  with r_t the CRC of
// chunk t alone.  Each thread runs the bit-serial register over its own chunk, multiplies by x^{s_t} mod P
// (square-and-multiply in GF(2)[x]/P, ~40 L-step products) and the partial remainders are XOR-reduced.
constexpr int kCrcBlockThreads = 256;  // long blocks; medium blocks (a codeblock's payload) use 64: fewer x^s products

__host__ __device__ __forceinline__ uint32_t gf2_mulmod(uint32_t a, uint32_t b, int L, uint32_t poly, uint32_t mask)
{
    uint32_t res = 0;
    for (int i = L - 1; i >= 0; --i) {
        const uint32_t top = (res >> (L - 1)) & 1u;
        res = (res << 1) & mask;
        if (top) res ^= poly;
        if ((b >> i) & 1u) res ^= a;
    }
    return res;
}

// x^(A - k1(t)) mod P of every thread's chunk, computed on the host per launch (one product per thread in the kernel
// instead of a square-and-multiply chain of ~40)
struct CrcPow { uint32_t p[kCrcBlockThreads]; };

// STAGE: the block of bits is first copied into shared memory with coalesced loads (a thread walking its own chunk of a
// global row reads one byte per 32-byte sector and instruction: 0.21 ms per 16 384 codeblock payloads, 5x the staged version)
template <int T, bool STAGE>
__global__ void __launch_bounds__(T)
crc_block_kernel(const __grid_constant__ CrcPow pw, const int8_t *__restrict__ in, int A, int L, uint32_t poly, int mode,
                 int8_t *__restrict__ out, uint8_t *__restrict__ err)
{
    constexpr int kCrcBlockThreads = T;
    extern __shared__ int8_t s_x[];
    __shared__ uint32_t s_part[kCrcBlockThreads / 32];
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31;
    const int len_in = mode ? A + L : A;
    const int8_t *x = in + (size_t)b * len_in;
    if (STAGE) {
        for (int k = tid; k < len_in; k += kCrcBlockThreads) s_x[k] = x[k];
        __syncthreads();
        x = s_x;
    }
    const uint32_t mask = (L == 32) ? 0xffffffffu : ((1u << L) - 1u);
    const int chunk = (A + kCrcBlockThreads - 1) / kCrcBlockThreads;
    const int k0 = min(tid * chunk, A), k1 = min(k0 + chunk, A);
    uint32_t rem = 0;
    for (int k = k0; k < k1; ++k) {  // chunk(x) * x^L mod P
        const uint32_t fb = ((rem >> (L - 1)) & 1u) ^ (uint32_t)(x[k] & 1);
        rem = (rem << 1) & mask;
        if (fb) rem ^= poly;
    }
    if (rem && k1 < A) rem = gf2_mulmod(rem, pw.p[tid], L, poly, mask);  // times x^(A - k1) mod P
#pragma unroll
    for (int o = 16; o; o >>= 1) rem ^= __shfl_xor_sync(0xffffffffu, rem, o);
    if (lane == 0) s_part[tid >> 5] = rem;
    __syncthreads();
    rem = 0;
#pragma unroll
    for (int w = 0; w < kCrcBlockThreads / 32; ++w) rem ^= s_part[w];
    if (mode == 0) {
        int8_t *y = out + (size_t)b * (A + L);
        for (int k = tid; k < A; k += kCrcBlockThreads) y[k] = x[k];
        if (tid < L) y[A + tid] = (int8_t)((rem >> (L - 1 - tid)) & 1u);
    } else if (tid == 0) {
        uint32_t got = 0;
        for (int k = 0; k < L; ++k) got = (got << 1) | (uint32_t)(x[A + k] & 1);
        err[b] = rem != got;
    }
}

// Bit-packed twin of count_errors_kernel: K bits per codeblock, ref / got rows of packed words.
__global__ void count_errors_packed_kernel(const uint32_t *__restrict__ ref, long long ref_stride,
                                           const uint32_t *__restrict__ got, long long got_stride, int B, int K,
                                           const int32_t *__restrict__ iters, unsigned long long *__restrict__ counters)
{
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    const int nwarps = (gridDim.x * blockDim.x) >> 5, nw = (K + 31) / 32;
    const uint32_t last = (K & 31) ? ((1u << (K & 31)) - 1u) : 0xffffffffu;
    unsigned long long blk = 0, bit = 0, its = 0, n = 0;
    for (int b = warp; b < B; b += nwarps) {
        int diff = 0;
        for (int w = lane; w < nw; w += 32) {
            uint32_t x = ref[b * ref_stride + w] ^ got[b * got_stride + w];
            if (w == nw - 1) x &= last;
            diff += __popc(x);
        }
        diff = __reduce_add_sync(0xffffffffu, diff);
        if (lane == 0) {
            ++n; bit += diff; blk += diff != 0;
            if (iters) its += iters[b];
        }
    }
    if (lane == 0 && n) {
        atomicAdd(counters + 0, n);
        atomicAdd(counters + 1, blk);
        atomicAdd(counters + 2, bit);
        atomicAdd(counters + 3, its);
    }
}

// Bit-packed twin of crc_block_kernel, mode 0, in place: one CTA per row; the row's A payload bits are staged in shared
// memory as words, every thread runs the remainder register over its chunk of bits, multiplies by x^(bits after the chunk)
// mod P and the partial remainders are XOR-reduced; thread 0 stores the L CRC bits at bit positions A .. A+L-1 of the row
// (remainder MSB first, py5gphy/crc/crc.py:34-38).
template <int T>
__global__ void __launch_bounds__(T)
crc_attach_packed_kernel(const __grid_constant__ CrcPow pw, uint32_t *__restrict__ words, long long row_words, int A, int L,
                         uint32_t poly)
{
    extern __shared__ uint32_t s_w[];
    __shared__ uint32_t s_part[T / 32];
    const int tid = threadIdx.x, lane = tid & 31, nw = (A + 31) / 32;
    uint32_t *x = words + (size_t)blockIdx.x * row_words;
    for (int k = tid; k < nw; k += T) s_w[k] = x[k];
    __syncthreads();
    const uint32_t mask = (L == 32) ? 0xffffffffu : ((1u << L) - 1u);
    const int chunk = (A + T - 1) / T;
    const int k0 = min(tid * chunk, A), k1 = min(k0 + chunk, A);
    uint32_t rem = 0;
    for (int k = k0; k < k1; ++k) {  // chunk(x) * x^L mod P
        const uint32_t fb = ((rem >> (L - 1)) ^ (s_w[k >> 5] >> (k & 31))) & 1u;
        rem = (rem << 1) & mask;
        if (fb) rem ^= poly;
    }
    if (rem && k1 < A) rem = gf2_mulmod(rem, pw.p[tid], L, poly, mask);  // times x^(A - k1) mod P
#pragma unroll
    for (int o = 16; o; o >>= 1) rem ^= __shfl_xor_sync(0xffffffffu, rem, o);
    if (lane == 0) s_part[tid >> 5] = rem;
    __syncthreads();
    if (tid == 0) {
        rem = 0;
#pragma unroll
        for (int w = 0; w < T / 32; ++w) rem ^= s_part[w];
        // CRC bit k (0 = first) = remainder bit L-1-k -> row bit A + k
        const unsigned long long crc = (unsigned long long)(__brev(rem) >> (32 - L));
        const int w0 = A >> 5, sh = A & 31;
        const unsigned long long val = crc << sh, m = (unsigned long long)mask << sh;
        x[w0] = (x[w0] & ~(uint32_t)m) | (uint32_t)val;
        if ((uint32_t)(m >> 32)) x[w0 + 1] = (x[w0 + 1] & ~(uint32_t)(m >> 32)) | (uint32_t)(val >> 32);
    }
}

}  // namespace

}  // namespace nrldpc

using namespace nrldpc;

static int grid_for(long long nblk, int per_sm)
{
    const long long g = (nblk + 255) / 256;
    return (int)(g > 148LL * per_sm ? 148LL * per_sm : (g < 1 ? 1 : g));
}

static int awgn_grid(long long rows, long long cols)
{
    const long long bpr = (cols + 3) / 4, spr = (bpr + 8191) / 8192, n = rows * spr;
    return (int)(n > 148LL * 8 ? 148LL * 8 : (n < 1 ? 1 : n));
}

extern "C" int nrldpc_random_bits_rows(int8_t *d_bits, long long rows, long long cols, unsigned long long seed,
                                       long long first_id, long long id_stride, void *stream)
{
    if (rows <= 0 || cols <= 0) return NRLDPC_OK;
    random_bits_kernel<<<grid_for(rows * ((cols + 127) / 128), 16), 256, 0, (cudaStream_t)stream>>>(d_bits, rows, cols, seed, 0, first_id, id_stride);
    NRLDPC_CUDA(cudaGetLastError());
    return NRLDPC_OK;
}

extern "C" int nrldpc_random_bits(int8_t *d_bits, long long count, unsigned long long seed, unsigned long long offset,
                                  void *stream)
{
    if (count <= 0) return NRLDPC_OK;
    random_bits_kernel<<<grid_for((count + 127) / 128, 16), 256, 0, (cudaStream_t)stream>>>(d_bits, 1, count, seed, offset, 0, 0);
    NRLDPC_CUDA(cudaGetLastError());
    return NRLDPC_OK;
}

extern "C" int nrldpc_awgn_llr_rows(const int8_t *d_dn, long long rows, long long cols, float snr_db, unsigned long long seed,
                                    long long first_id, long long id_stride, float *d_llr, void *stream)
{
    if (rows <= 0 || cols <= 0) return NRLDPC_OK;
    const double sigma = pow(10.0, -(double)snr_db / 20.0), np = pow(10.0, -(double)snr_db / 10.0);
    awgn_llr_kernel<false><<<awgn_grid(rows, cols), 256, 0, (cudaStream_t)stream>>>(
        d_dn, nullptr, 0, rows, cols, (float)sigma, (float)(2.0 / np), seed, 0, first_id, id_stride, d_llr);
    NRLDPC_CUDA(cudaGetLastError());
    return NRLDPC_OK;
}

extern "C" int nrldpc_awgn_llr(const int8_t *d_dn, long long count, float snr_db, unsigned long long seed,
                               unsigned long long offset, float *d_llr, void *stream)
{
    if (count <= 0) return NRLDPC_OK;
    const double sigma = pow(10.0, -(double)snr_db / 20.0), np = pow(10.0, -(double)snr_db / 10.0);
    awgn_llr_kernel<false><<<awgn_grid(1, count), 256, 0, (cudaStream_t)stream>>>(
        d_dn, nullptr, 0, 1, count, (float)sigma, (float)(2.0 / np), seed, offset, 0, 0, d_llr);
    NRLDPC_CUDA(cudaGetLastError());
    return NRLDPC_OK;
}

extern "C" int nrldpc_count_errors(const int8_t *d_ref, long long ref_stride, const int8_t *d_got, long long got_stride,
                                   int B, int K, const int32_t *d_iters, long long *d_counters, void *stream)
{
    if (B <= 0) return NRLDPC_OK;
    const int grid = (B + 7) / 8 > 148 * 8 ? 148 * 8 : (B + 7) / 8;
    count_errors_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(d_ref, ref_stride, d_got, got_stride, B, K, d_iters,
                                                               reinterpret_cast<unsigned long long *>(d_counters));
    NRLDPC_CUDA(cudaGetLastError());
    return NRLDPC_OK;
}

extern "C" int nrldpc_random_bits_packed_rows(uint32_t *d_words, long long rows, long long cols, long long row_words,
                                              unsigned long long seed, long long first_id, long long id_stride, void *stream)
{
    if (rows < 0 || cols < 0 || row_words < (cols + 31) / 32 || (rows > 0 && !d_words)) {
        set_error("random_bits_packed_rows: bad argument (row_words >= ceil(cols / 32))");
        return NRLDPC_EINVAL;
    }
    if (rows == 0 || row_words == 0) return NRLDPC_OK;
    random_bits_packed_kernel<<<grid_for(rows * ((row_words + 3) / 4), 16), 256, 0, (cudaStream_t)stream>>>(
        d_words, rows, cols, row_words, seed, first_id, id_stride);
    NRLDPC_CUDA(cudaGetLastError());
    return NRLDPC_OK;
}

extern "C" int nrldpc_awgn_llr_packed_rows(const uint32_t *d_dn_words, long long rows, long long cols, long long row_words,
                                           float snr_db, unsigned long long seed, long long first_id, long long id_stride,
                                           float *d_llr, void *stream)
{
    if (rows < 0 || cols < 0 || row_words < (cols + 31) / 32 || (rows > 0 && cols > 0 && (!d_dn_words || !d_llr))) {
        set_error("awgn_llr_packed_rows: bad argument (row_words >= ceil(cols / 32))");
        return NRLDPC_EINVAL;
    }
    if (rows == 0 || cols == 0) return NRLDPC_OK;
    const double sigma = pow(10.0, -(double)snr_db / 20.0), np = pow(10.0, -(double)snr_db / 10.0);
    awgn_llr_kernel<true><<<awgn_grid(rows, cols), 256, 0, (cudaStream_t)stream>>>(
        nullptr, d_dn_words, row_words, rows, cols, (float)sigma, (float)(2.0 / np), seed, 0, first_id, id_stride, d_llr);
    NRLDPC_CUDA(cudaGetLastError());
    return NRLDPC_OK;
}

extern "C" int nrldpc_count_errors_packed(const uint32_t *d_ref_words, long long ref_row_words, const uint32_t *d_got_words,
                                          long long got_row_words, int B, int K, const int32_t *d_iters, long long *d_counters,
                                          void *stream)
{
    const long long nw = ((long long)K + 31) / 32;
    if (B < 0 || K < 0 || ref_row_words < nw || got_row_words < nw || !d_counters || (B > 0 && (!d_ref_words || !d_got_words))) {
        set_error("count_errors_packed: bad argument");
        return NRLDPC_EINVAL;
    }
    if (B == 0) return NRLDPC_OK;
    const int grid = (B + 7) / 8 > 148 * 8 ? 148 * 8 : (B + 7) / 8;
    count_errors_packed_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(d_ref_words, ref_row_words, d_got_words, got_row_words, B, K,
                                                                      d_iters, reinterpret_cast<unsigned long long *>(d_counters));
    NRLDPC_CUDA(cudaGetLastError());
    return NRLDPC_OK;
}

// polynomial bit patterns of py5gphy/crc/crc.py:96-106 (x^L term dropped), MSB = first array element
static int crc_poly(int poly_id, int *L, uint32_t *poly)
{
    static const int len[6] = {6, 11, 16, 24, 24, 24};
    static const uint32_t pat[6] = {0x21u, 0x621u, 0x1021u, 0x864CFBu, 0x800063u, 0xB2B117u};
    if (poly_id < 0 || poly_id > 5) { set_error("crc: poly_id must be 0..5 ('6','11','16','24A','24B','24C')"); return NRLDPC_EINVAL; }
    *L = len[poly_id];
    *poly = pat[poly_id];
    return NRLDPC_OK;
}

// long blocks: one CTA per block of bits
// (one thread per block walks its bits serially with uncoalesced byte loads: only good for short blocks)
static bool crc_use_block_kernel(int B, int A) { (void)B; return A >= 1024; }

static void crc_pow_table(int A, int L, uint32_t poly, int T, CrcPow *pw)
{
    const uint32_t mask = (L == 32) ? 0xffffffffu : ((1u << L) - 1u);
    auto powx = [&](uint32_t e) {
        uint32_t base = 2u, acc = 1u;
        while (e) {
            if (e & 1u) acc = gf2_mulmod(acc, base, L, poly, mask);
            base = gf2_mulmod(base, base, L, poly, mask);
            e >>= 1;
        }
        return acc;
    };
    const int chunk = (A + T - 1) / T;
    const uint32_t xchunk = powx((uint32_t)chunk);
    int e_next = -1;
    for (int t = T - 1; t >= 0; --t) {
        const int k0 = std::min(t * chunk, A), k1 = std::min(k0 + chunk, A), e = A - k1;
        pw->p[t] = (e_next >= 0 && e == e_next + chunk) ? gf2_mulmod(pw->p[t + 1], xchunk, L, poly, mask) : powx((uint32_t)e);
        e_next = e;
    }
}

template <int T>
static void crc_block_launch(const int8_t *d_in, int B, int A, int L, uint32_t poly, int mode, int8_t *d_out, uint8_t *d_err,
                             cudaStream_t s)
{
    CrcPow pw;
    crc_pow_table(A, L, poly, T, &pw);
    const int len_in = mode ? A + L : A;
    if (len_in <= 40 * 1024) crc_block_kernel<T, true><<<B, T, len_in, s>>>(pw, d_in, A, L, poly, mode, d_out, d_err);
    else crc_block_kernel<T, false><<<B, T, 0, s>>>(pw, d_in, A, L, poly, mode, d_out, d_err);
}

extern "C" int nrldpc_crc_encode(const int8_t *d_in, int B, int A, int poly_id, int8_t *d_out, void *stream)
{
    int L; uint32_t poly;
    if (int rc = crc_poly(poly_id, &L, &poly)) return rc;
    if (B < 0 || A < 0 || !d_in || !d_out) { set_error("crc_encode: bad argument"); return NRLDPC_EINVAL; }
    if (B == 0) return NRLDPC_OK;
    if (crc_use_block_kernel(B, A)) {
        if (A >= 32768) crc_block_launch<kCrcBlockThreads>(d_in, B, A, L, poly, 0, d_out, nullptr, (cudaStream_t)stream);
        else crc_block_launch<64>(d_in, B, A, L, poly, 0, d_out, nullptr, (cudaStream_t)stream);
    }
    else crc_kernel<<<(B + 127) / 128, 128, 0, (cudaStream_t)stream>>>(d_in, B, A, L, poly, 0, d_out, nullptr);
    NRLDPC_CUDA(cudaGetLastError());
    return L;
}

extern "C" int nrldpc_crc_check(const int8_t *d_in, int B, int A, int poly_id, uint8_t *d_err, void *stream)
{
    int L; uint32_t poly;
    if (int rc = crc_poly(poly_id, &L, &poly)) return rc;
    if (B < 0 || A < 0 || !d_in || !d_err) { set_error("crc_check: bad argument"); return NRLDPC_EINVAL; }
    if (B == 0) return NRLDPC_OK;
    if (crc_use_block_kernel(B, A)) {
        if (A >= 32768) crc_block_launch<kCrcBlockThreads>(d_in, B, A, L, poly, 1, nullptr, d_err, (cudaStream_t)stream);
        else crc_block_launch<64>(d_in, B, A, L, poly, 1, nullptr, d_err, (cudaStream_t)stream);
    }
    else crc_kernel<<<(B + 127) / 128, 128, 0, (cudaStream_t)stream>>>(d_in, B, A, L, poly, 1, nullptr, d_err);
    NRLDPC_CUDA(cudaGetLastError());
    return L;
}

extern "C" int nrldpc_crc_attach_packed(uint32_t *d_words, int B, int A, int poly_id, long long row_words, void *stream)
{
    int L; uint32_t poly;
    if (int rc = crc_poly(poly_id, &L, &poly)) return rc;
    if (B < 0 || A < 0 || row_words < ((long long)A + L + 31) / 32 || (B > 0 && !d_words)) {
        set_error("crc_attach_packed: bad argument (row_words >= ceil((A + L) / 32))");
        return NRLDPC_EINVAL;
    }
    if (B == 0) return L;
    const size_t smem = (size_t)((A + 31) / 32) * 4;
    if (smem > 200 * 1024) { set_error("crc_attach_packed: rows of more than 1.6 Mbit are not supported"); return NRLDPC_EINVAL; }
    CrcPow pw;
    if (A >= 32768) {
        crc_pow_table(A, L, poly, kCrcBlockThreads, &pw);
        if (smem > 48 * 1024)
            NRLDPC_CUDA(cudaFuncSetAttribute(crc_attach_packed_kernel<kCrcBlockThreads>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        crc_attach_packed_kernel<kCrcBlockThreads><<<B, kCrcBlockThreads, smem, (cudaStream_t)stream>>>(pw, d_words, row_words, A, L, poly);
    } else {
        crc_pow_table(A, L, poly, 64, &pw);
        crc_attach_packed_kernel<64><<<B, 64, smem, (cudaStream_t)stream>>>(pw, d_words, row_words, A, L, poly);
    }
    NRLDPC_CUDA(cudaGetLastError());
    return L;
}

namespace {
using TmpBuf = ScratchBuf;
}

extern "C" int nrldpc_crc_encode_host(const int8_t *in, int B, int A, int poly_id, int8_t *out)
{
    int L; uint32_t poly;
    if (int rc = crc_poly(poly_id, &L, &poly)) return rc;
    if (B <= 0 || A < 0) return B == 0 ? L : NRLDPC_EINVAL;
    TmpBuf di, dout;
    NRLDPC_CUDA(di.alloc((size_t)B * A + 1));
    NRLDPC_CUDA(dout.alloc((size_t)B * (A + L)));
    NRLDPC_CUDA(cudaMemcpy(di.p, in, (size_t)B * A, cudaMemcpyHostToDevice));
    int rc = nrldpc_crc_encode((const int8_t *)di.p, B, A, poly_id, (int8_t *)dout.p, nullptr);
    if (rc < 0) return rc;
    NRLDPC_CUDA(cudaMemcpy(out, dout.p, (size_t)B * (A + L), cudaMemcpyDeviceToHost));
    return L;
}

extern "C" int nrldpc_crc_check_host(const int8_t *in, int B, int A, int poly_id, uint8_t *err)
{
    int L; uint32_t poly;
    if (int rc = crc_poly(poly_id, &L, &poly)) return rc;
    if (B <= 0 || A < 0) return B == 0 ? L : NRLDPC_EINVAL;
    TmpBuf di, de;
    NRLDPC_CUDA(di.alloc((size_t)B * (A + L)));
    NRLDPC_CUDA(de.alloc((size_t)B));
    NRLDPC_CUDA(cudaMemcpy(di.p, in, (size_t)B * (A + L), cudaMemcpyHostToDevice));
    int rc = nrldpc_crc_check((const int8_t *)di.p, B, A, poly_id, (uint8_t *)de.p, nullptr);
    if (rc < 0) return rc;
    NRLDPC_CUDA(cudaMemcpy(err, de.p, (size_t)B, cudaMemcpyDeviceToHost));
    return L;
}
