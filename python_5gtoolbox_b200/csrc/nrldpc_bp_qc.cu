// nrldpc_bp_qc.cu -- sum-product ('BP') decoder on the quasi-cyclic structure of the 5G matrices, device-resident.
//
// Replaces nr_decode_ldpc(..., algo='BP') = decode_ldpc + _BP_process (py5gphy/ldpc/nr_ldpc_decode.py:51-143, :145-176)
// for the 5G matrices, which before went through the generic CSR kernel with a host round trip per batch.  Same flooding
// schedule and operation order as the reference, float64 arithmetic like the reference (tanh / atanh are this library's
// own, nrldpc_bp_math.cuh, within 3 ulp: intermediate LLRs are not bit-identical to NumPy's, as with any other libm; hard
// decisions, status and iteration counts are what the goldens pin):
//   per row:  t_e = tanh(Lq_e / 2), P = prod t_e, Lr_e = 2 atanh(P / t_e) clipped to +-38.14 when |x| >= 1 (:158-163);
//             exactly one zero input: that edge gets prod tanh(others) WITHOUT the 2 atanh (:164-170); >= 2 zeros: 0.
// Per-edge messages are needed (no min1/min2 compression for BP): Lr[nnz][Zc] doubles = 970 KB per codeblock at BG1
// Zc=384 lives in a per-CTA global workspace (persistent CTAs; 287 MB in all, so half of it comes from HBM), the posteriors of the core columns in
// shared memory, the degree-1 extension variables are recomputed from their single message.  The circulant shift is
// index arithmetic on the lifted index, the tables sit in the constant bank (QcCfg by value).
#include <algorithm>

#include "nrldpc_bp_math.cuh"
#include "nrldpc_common.cuh"

namespace nrldpc {
namespace {

#ifndef NRLDPC_BP_THREADS
#define NRLDPC_BP_THREADS 512
#endif
constexpr int kBpThreads = NRLDPC_BP_THREADS;  // 1024 / kBpThreads persistent CTAs per SM
constexpr int kBpBatch = 4;  // edges of a row whose workspace words are in flight together

template <typename TIn>
__global__ void __launch_bounds__(kBpThreads)
bp_qc_kernel(const __grid_constant__ QcCfg c, const TIn *__restrict__ llr, int B, int max_iter, int early_term,
             double *__restrict__ work, int8_t *__restrict__ ck, uint8_t *__restrict__ status, int32_t *__restrict__ iters)
{
    extern __shared__ double LQ[];  // [ncore][Zc], then (Zc a multiple of 32) the packed hard decisions: uint32 [Nfull / 32]
    __shared__ int s_any;
    const int Zc = c.Zc, tid = threadIdx.x, nnz = c.rowptr[c.nrows];
    double *Lr = work + (size_t)blockIdx.x * (size_t)nnz * Zc;  // [edge][r]
    const int nchk = c.nrows * Zc, ncorev = c.ncore * Zc;
    // In-loop syndrome (:107-114) on bit-packed decisions when Zc is a multiple of 32: whoever computes a posterior (the
    // variable pass; sweep 2 of a check row for its degree-1 extension variable) ballots its sign into hard[n / 32], a warp
    // always holds 32 consecutive lifted indices of one block; the syndrome word of (row-block, word) is then the XOR of
    // funnel-shifted words, 32 checks per lane, instead of one walk over the edges per check.
    const bool packed = (Zc & 31) == 0;
    const int W = Zc >> 5;
    uint32_t *hard = reinterpret_cast<uint32_t *>(LQ + ncorev);

    for (int cb = blockIdx.x; cb < B; cb += gridDim.x) {
        const TIn *L0 = llr + (size_t)cb * c.N;
        // channel LLR of variable n of the full codeword (the 2Zc punctured columns start at 0, :43); -0.0 -> +0.0
        auto chan = [&](int n) -> double { return n < 2 * Zc ? 0.0 : (double)L0[n - 2 * Zc] + 0.0; };
        // :94-101  LQ = LLRin, Lr = 0
        for (int n = tid; n < ncorev; n += kBpThreads) LQ[n] = chan(n);
        if (packed)
            for (int n = tid; n < c.Nfull; n += kBpThreads) {  // Lr = 0: every posterior is its channel LLR
                const uint32_t b = __ballot_sync(0xffffffffu, chan(n) < 0.0);
                if ((tid & 31) == 0) hard[n >> 5] = b;
            }
        for (int e = tid; e < nnz * Zc; e += kBpThreads) Lr[e] = 0.0;
        if (tid == 0) s_any = 0;
        __syncthreads();

        // posterior of the variable on edge e (column-block j, shift P) seen from check r
        auto post = [&](int j, int P, int e, int r) -> double {
            if (j < c.ncore) {
                int v = r + P;
                if (v >= Zc) v -= Zc;
                return LQ[j * Zc + v];
            }
            return chan(j * Zc + r) + Lr[e * Zc + r];  // degree-1 extension variable: LQ = LLR + its one message (:126)
        };
        auto syndrome = [&](bool final_rule) -> int {
            int any = 0;
            for (int m = tid; m < nchk; m += kBpThreads) {
                const int i = m / Zc, r = m - i * Zc;
                int p = 0;
                for (int e = c.rowptr[i]; e < c.rowptr[i + 1]; ++e) {
                    const uint32_t w = c.edge[e];
                    const double x = post((int)(w & 0xff), (int)(w >> 8), e, r);
                    p ^= final_rule ? (x <= 0.0) : (x < 0.0);
                }
                any |= p;
            }
            return any;
        };

        auto syndrome_packed = [&]() -> int {
            uint32_t any = 0;
            for (int t = tid; t < c.nrows * W; t += kBpThreads) {
                const int i = t / W, w = t - i * W;
                uint32_t acc = i >= 4 ? hard[(c.ncore + i - 4) * W + w] : 0u;
                for (int e = c.rowptr[i]; e < c.rowptr[i + 1] - (i >= 4); ++e) {
                    const uint32_t ew = c.edge[e];
                    const int P = (int)(ew >> 8);
                    int w0 = w + (P >> 5);  // check 32 w + b reads variable 32 w + b + P (mod Zc)
                    if (w0 >= W) w0 -= W;
                    const int w1 = w0 + 1 < W ? w0 + 1 : 0;
                    const uint32_t *col = hard + (int)(ew & 0xff) * W;
                    acc ^= __funnelshift_r(col[w0], col[w1], P & 31);
                }
                any |= acc;
            }
            return any != 0;
        };

        bool done = false;
        int it = 0;
        for (; it < max_iter; ++it) {
            // :107-114 syndrome of the decisions LQ < 0 (its only use is the early exit)
            if (early_term && (packed ? syndrome_packed() : syndrome(false))) s_any = 1;
            __syncthreads();
            const int bad = s_any;
            __syncthreads();
            if (tid == 0) s_any = 0;
            if (!bad && early_term) { done = true; break; }

            // :117-123 every check row from the old Lq = LQ - Lr; sweep 1 leaves tanh(Lq/2) in the message slot.
            // Both sweeps request the workspace words of kBpBatch edges before they start on the arithmetic (the loads are
            // L2 / HBM latency); the degree-1 extension edge of rows >= 4 (the row's last edge) is handled on its own.
            for (int m = tid; m < nchk; m += kBpThreads) {
                const int i = m / Zc, r = m - i * Zc;
                const int e0 = c.rowptr[i], ne = c.rowptr[i + 1] - e0, nce = ne - (i >= 4);
                double *slot0 = Lr + (e0 * Zc + r);  // edge k of the row: slot0[k * Zc] (32-bit offsets: nnz Zc < 2^17)
                int nz = 0, zi = -1;
                double prod = 1.0;
                auto take = [&](int k, double q) {
                    double t = 0.0;
                    if (q == 0.0) { ++nz; if (zi < 0) zi = k; }
                    else { t = bpmath::tanh_half(q); prod *= t; }
                    slot0[k * Zc] = t;
                };
                double qx = 0.0, chx = 0.0;
                if (nce < ne) {  // LQ = LLR + its one message (:126), Lq = LQ - Lr
                    const double lr = slot0[nce * Zc];
                    chx = chan((c.ncore + i - 4) * Zc + r);
                    qx = __dsub_rn(chx + lr, lr);
                }
                for (int kb = 0; kb < nce; kb += kBpBatch) {
                    double q[kBpBatch];
#pragma unroll
                    for (int u = 0; u < kBpBatch; ++u) {
                        const int k = kb + u < nce ? kb + u : nce - 1;  // (a repeated load instead of a branch)
                        const uint32_t w = c.edge[e0 + k];
                        int v = r + (int)(w >> 8);
                        if (v >= Zc) v -= Zc;
                        q[u] = __dsub_rn(LQ[(int)(w & 0xff) * Zc + v], slot0[k * Zc]);
                    }
#pragma unroll
                    for (int u = 0; u < kBpBatch; ++u)
                        if (kb + u < nce) take(kb + u, q[u]);
                }
                if (nce < ne) take(nce, qx);
                double last = 0.0;  // the new message on the row's last edge
                for (int kb = 0; kb < ne; kb += kBpBatch) {
                    double t[kBpBatch];
#pragma unroll
                    for (int u = 0; u < kBpBatch; ++u) t[u] = slot0[(kb + u < ne ? kb + u : ne - 1) * Zc];
#pragma unroll
                    for (int u = 0; u < kBpBatch; ++u) {
                        const int k = kb + u;
                        if (k < ne) {
                            double out = 0.0;
                            if (nz == 0) {
                                const double x = prod / t[u];
                                out = x >= 1.0 ? 2 * 19.07 : (x <= -1.0 ? -2 * 19.07 : bpmath::atanh_twice(x));
                            } else if (nz == 1 && k == zi) {
                                out = prod;  // :170 the reference omits 2*atanh here
                            }
                            slot0[k * Zc] = out;
                            last = out;
                        }
                    }
                }
                if (packed && early_term && nce < ne) {  // (warp-uniform: 32 consecutive checks of one row-block)
                    const uint32_t b = __ballot_sync(0xffffffffu, chx + last < 0.0);
                    if ((tid & 31) == 0) hard[((c.ncore + i - 4) * Zc + r) >> 5] = b;
                }
            }
            __syncthreads();
            // :126 LQ = LLRin + Lr.sum(axis=0) in ascending check index, core columns (the others are recomputed on use)
            for (int n = tid; n < ncorev; n += kBpThreads) {
                const int j = n / Zc, v = n - j * Zc;
                double s = 0.0;
#pragma unroll 4
                for (int q = c.colptr[j]; q < c.colptr[j + 1]; ++q) {
                    const uint32_t en = c.centry[q];
                    const int i = en & 63, k = (en >> 6) & 31, back = en >> 16;
                    int r = v + back;
                    if (r >= Zc) r -= Zc;
                    s = __dadd_rn(s, Lr[(c.rowptr[i] + k) * Zc + r]);
                }
                const double x = __dadd_rn(chan(n), s);
                LQ[n] = x;
                if (packed) {
                    const uint32_t b = __ballot_sync(0xffffffffu, x < 0.0);
                    if ((tid & 31) == 0) hard[n >> 5] = b;
                }
            }
            __syncthreads();
        }
        int ok = 1;
        if (!done) {  // :134-143 final decision with the other tie rule
            if (syndrome(true)) s_any = 1;
            __syncthreads();
            ok = !s_any;
        }
        // hard decisions: core columns from LQ, extension column kb+4+i4 from its single message (row 4+i4, last edge)
        int8_t *out = ck + (size_t)cb * c.Nfull;
        for (int n = tid; n < c.Nfull; n += kBpThreads) {
            double x;
            if (n < ncorev) x = LQ[n];
            else {
                const int i = 4 + (n - ncorev) / Zc, r = (n - ncorev) % Zc;
                x = chan(n) + Lr[(c.rowptr[i + 1] - 1) * Zc + r];
            }
            out[n] = (int8_t)(done ? (x < 0.0) : (x <= 0.0));
        }
        if (tid == 0) {
            if (status) status[cb] = (uint8_t)ok;
            if (iters) iters[cb] = it;
        }
        __syncthreads();  // s_any and the state are re-initialised for the next codeblock
    }
}

}  // namespace

int launch_bp_qc(const QcCfg &c, const void *d_llr, int is_f64, int B, int max_iter, int early_term, int8_t *d_ck,
                 uint8_t *d_status, int32_t *d_iters, cudaStream_t s)
{
    if (B <= 0) return NRLDPC_OK;
    const int nnz = c.rowptr[c.nrows];
    const int smem = c.ncore * c.Zc * (int)sizeof(double) + (c.Zc % 32 == 0 ? c.Nfull / 32 * 4 : 0);  // posteriors + packed decisions
    int dev = 0, sms = 148;
    NRLDPC_CUDA(cudaGetDevice(&dev));
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int grid = std::min(B, (1024 / kBpThreads) * sms);  // two CTAs per SM (83 KB of shared memory each at Zc = 384)
    ScratchBuf work;
    NRLDPC_CUDA(work.alloc((size_t)grid * nnz * c.Zc * sizeof(double), s));
    if (is_f64) {
        NRLDPC_CUDA(cudaFuncSetAttribute(bp_qc_kernel<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        bp_qc_kernel<double><<<grid, kBpThreads, smem, s>>>(c, (const double *)d_llr, B, max_iter, early_term, work.as<double>(), d_ck, d_status, d_iters);
    } else {
        NRLDPC_CUDA(cudaFuncSetAttribute(bp_qc_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        bp_qc_kernel<float><<<grid, kBpThreads, smem, s>>>(c, (const float *)d_llr, B, max_iter, early_term, work.as<double>(), d_ck, d_status, d_iters);
    }
    NRLDPC_CUDA(cudaGetLastError());
    return NRLDPC_OK;
}

}  // namespace nrldpc

using namespace nrldpc;

extern "C" int nrldpc_decode_bp(const void *d_llr, int is_f64, int B, int bgn, int Zc, int max_iter, int early_term,
                                int8_t *d_ck, uint8_t *d_status, int32_t *d_iters, void *stream)
{
    const QcCfg *c = get_cfg(bgn, Zc);
    if (!c) return NRLDPC_EINVAL;
    if (B < 0 || max_iter < 0 || !d_llr || !d_ck) { set_error("decode_bp: bad argument"); return NRLDPC_EINVAL; }
    return launch_bp_qc(*c, d_llr, is_f64, B, max_iter, early_term, d_ck, d_status, d_iters, (cudaStream_t)stream);
}

extern "C" int nrldpc_decode_bp_host(const void *llr, int is_f64, int B, int bgn, int Zc, int max_iter, int early_term,
                                     int8_t *ck, uint8_t *status, int32_t *iters)
{
    const QcCfg *c = get_cfg(bgn, Zc);
    if (!c) return NRLDPC_EINVAL;
    if (B < 0 || max_iter < 0 || !llr || !ck) { set_error("decode_bp: bad argument"); return NRLDPC_EINVAL; }
    if (B == 0) return NRLDPC_OK;
    cudaStream_t s;
    if (int rc = host_stream(&s)) return rc;
    const size_t esz = is_f64 ? 8 : 4;
    ScratchBuf d_llr, d_ck, d_st, d_it;
    NRLDPC_CUDA(d_llr.alloc((size_t)B * c->N * esz, s));
    NRLDPC_CUDA(d_ck.alloc((size_t)B * c->Nfull, s));
    NRLDPC_CUDA(d_st.alloc((size_t)B, s));
    NRLDPC_CUDA(d_it.alloc((size_t)B * 4, s));
    int rc = h2d_async(d_llr.p, llr, (size_t)B * c->N * esz, s);
    if (rc == NRLDPC_OK) rc = launch_bp_qc(*c, d_llr.p, is_f64, B, max_iter, early_term, d_ck.as<int8_t>(), d_st.as<uint8_t>(), d_it.as<int32_t>(), s);
    if (rc == NRLDPC_OK && status && cudaMemcpyAsync(status, d_st.p, (size_t)B, cudaMemcpyDeviceToHost, s) != cudaSuccess) rc = cuda_fail(cudaGetLastError(), "cudaMemcpyAsync");
    if (rc == NRLDPC_OK && iters && cudaMemcpyAsync(iters, d_it.p, (size_t)B * 4, cudaMemcpyDeviceToHost, s) != cudaSuccess) rc = cuda_fail(cudaGetLastError(), "cudaMemcpyAsync");
    if (rc == NRLDPC_OK) rc = d2h_sync(ck, d_ck.p, (size_t)B * c->Nfull, s);
    else cudaStreamSynchronize(s);
    return rc;
}
