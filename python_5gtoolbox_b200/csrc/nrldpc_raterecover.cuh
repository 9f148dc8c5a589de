// nrldpc_raterecover.cuh -- rate recovery (+ HARQ soft combining) of ONE codeblock by ONE CTA, as a device function.
//
// Used by the stand-alone raterecover_kernel (nrldpc_ratematch.cu) and, fused, by the prologue of the min-sum decoder
// kernels: the decoder's LLR load IS the rate recovery, so a transport block goes from the
// concatenated received LLRs to decoded bits in one launch and the recovered fp32 row never leaves L2.
//
// Replaces  nr_ldpc_raterecover.raterecover_ldpc               (py5gphy/ldpc/nr_ldpc_raterecover.py:6-65)
//           the HARQ combining loop of DLSCHDecode / ULSCH_decoding
//                                                              (py5gphy/nr_pdsch/nr_dlsch_decode.py:74-87,
//                                                               py5gphy/nr_pusch/nr_ulsch_decode.py:75-88)
// Arithmetic is float64 in the reference's operation order (sum of the repetitions in arrival order / their number;
// (a + c) / 2 where both transmissions are non-zero), rounded ONCE to fp32 for the decoder.
#pragma once
#include <stdint.h>

namespace nrldpc {

// What the decoder kernels need to recover their own LLRs.  src == nullptr: plain llr[B,N] input (no recovery).
struct RrArgs {
    const void *src = nullptr;     // concatenated received LLRs of the transport block, float32 or float64 (in_f64)
    const int32_t *E = nullptr;    // [B] received length of every codeblock (a multiple of Qm)
    const long long *goff = nullptr;  // [B] offset of codeblock b inside src
    const double *cur = nullptr;   // [B,N] soft buffer of the earlier transmissions to combine with, or null
    double *soft = nullptr;        // [B,N] combined soft buffer out (device memory or mapped pinned host memory), or null
    float *scratch = nullptr;      // fp32 rows the decoder iterates on: [gridDim.x, N] (persistent kernels) or [B, N]
    int in_f64 = 0;
    int Ncb = 0, k0 = 0, Qm = 1, F0 = 0, F1 = 0;  // circular buffer, start, modulation order, fillers [F0, F1) of dn
};

// Number of filler positions met strictly before step t of the walk that starts at k0: the fillers are
// the buffer positions [f0, f1) (py5gphy/ldpc/nr_ldpc_raterecover.py:34).
__device__ __forceinline__ int rr_fillers_before(int t, int k0, int f0, int f1, int Ncb)
{
    if (f1 <= f0) return 0;
    auto overlap = [t](int a, int b) { return max(0, min(t, b) - a); };  // |[0,t) & [a,b)|, a >= 0
    if (k0 <= f0) return overlap(f0 - k0, f1 - k0);
    if (k0 >= f1) return overlap(f0 - k0 + Ncb, f1 - k0 + Ncb);
    return overlap(0, f1 - k0) + overlap(f0 - k0 + Ncb, Ncb);
}

// 10 * max|LLr_fe| over the CTA (:30), all threads of the CTA must call; s_red = 33 doubles of shared memory.
template <typename TIn>
__device__ __forceinline__ double rr_block_max10(const TIn *__restrict__ fe, int E, double *s_red)
{
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = (blockDim.x + 31) >> 5;
    double m = 0.0;
    for (int e = tid; e < E; e += blockDim.x) m = fmax(m, fabs((double)fe[e]));
#pragma unroll
    for (int o = 16; o; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
    __syncthreads();  // s_red may still be read from an earlier call
    if (lane == 0) s_red[warp] = m;
    __syncthreads();
    if (tid == 0) {
        double mm = 0.0;
        for (int w = 0; w < nwarps; ++w) mm = fmax(mm, s_red[w]);
        s_red[32] = mm * 10.0;
    }
    __syncthreads();
    return s_red[32];
}

// De-interleaving + de-selection with averaging of repeated bits (:27-63): position pos of the circular
// buffer, met at step t of the walk with `rank` non-filler positions before it, receives the received
// values k = rank, rank + S, ... < E; the output is their float64 sum in that order divided by their
// number (0 when there is none), the fillers get 10 * max|LLr_fe| (:30,:64), the rest of [0,N) is 0.
// Then, with `cur`, the HARQ combination with the stored soft value (a zero on either side = not received there).
// out64 / out32 may be null.  All threads of the CTA must call (block reduction inside when there are fillers).
template <typename TIn>
__device__ __forceinline__ void rr_codeblock(const TIn *__restrict__ fe, int E, int N, int Ncb, int k0, int Qm, int F0, int F1,
                                             const double *__restrict__ cur, double *__restrict__ out64,
                                             float *__restrict__ out32, double *s_red)
{
    const int tid = threadIdx.x;
    const double max_llr = (F1 > F0) ? rr_block_max10(fe, E, s_red) : 0.0;
    const int f0 = min(max(F0, 0), Ncb), f1 = min(max(F1, f0), Ncb);
    const int S = Ncb - (f1 - f0);
    const int cols = Qm > 0 ? E / Qm : 0;
    const float inv_cols = cols > 0 ? 1.0f / (float)cols : 0.f;
    for (int pos = tid; pos < N; pos += blockDim.x) {
        double v = 0.0;
        if (pos >= F0 && pos < F1) {
            v = max_llr;
        } else if (pos < Ncb && S > 0 && E > 0) {
            int t = pos - k0;
            if (t < 0) t += Ncb;
            const int rank = t - rr_fillers_before(t, k0, f0, f1, Ncb);
            double sum = 0.0;
            int cnt = 0;
            for (int k = rank; k < E; k += S) {
                // q = k / cols, e = k % cols without an integer division (k < 2^24: exact in fp32, fixed up by one)
                int q = __float2int_rz(__int2float_rn(k) * inv_cols);
                int e = k - q * cols;
                if (e < 0) { --q; e += cols; }
                else if (e >= cols) { ++q; e -= cols; }
                sum += (double)fe[(size_t)e * Qm + q];
                ++cnt;
            }
            v = cnt ? sum / (double)cnt : 0.0;
        }
        if (cur) {  // py5gphy/nr_pdsch/nr_dlsch_decode.py:80-87
            const double c = cur[pos];
            v = (v == 0.0 || c == 0.0) ? v + c : (v + c) / 2.0;
        }
        if (out64) out64[pos] = v;
        if (out32) out32[pos] = (float)v;
    }
}

// Codeblock cb of a transport block described by rr -> its fp32 row `row` (and the float64 soft buffer rr.soft).
__device__ __forceinline__ void rr_row(const RrArgs &rr, int cb, int N, float *row, double *s_red)
{
    const int E = rr.E[cb];
    const long long off = rr.goff[cb];
    const double *cur = rr.cur ? rr.cur + (size_t)cb * N : nullptr;
    double *soft = rr.soft ? rr.soft + (size_t)cb * N : nullptr;
    if (rr.in_f64)
        rr_codeblock(reinterpret_cast<const double *>(rr.src) + off, E, N, rr.Ncb, rr.k0, rr.Qm, rr.F0, rr.F1, cur, soft, row, s_red);
    else
        rr_codeblock(reinterpret_cast<const float *>(rr.src) + off, E, N, rr.Ncb, rr.k0, rr.Qm, rr.F0, rr.F1, cur, soft, row, s_red);
}

}  // namespace nrldpc
