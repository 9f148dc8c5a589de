// nrldpc_ratematch.cu -- LDPC rate matching / rate recovery of TS 38.212 5.4.2, batched over the
// codeblocks of a transport block (the steps either side of the encoder / decoder in PDSCH and PUSCH).
//
// Replaces   nr_ldpc_ratematch.ratematch_ldpc      (py5gphy/ldpc/nr_ldpc_ratematch.py:64-97)
//            nr_ldpc_raterecover.raterecover_ldpc  (py5gphy/ldpc/nr_ldpc_raterecover.py:6-65)
//            the HARQ soft-combining loop of DLSCHDecode / ULSCH_decoding
//                                                  (py5gphy/nr_pdsch/nr_dlsch_decode.py:74-87)
// One CTA per codeblock; both kernels are pure streaming (HBM-bound): every input element is read once,
// every output element written once.
#include <cstdint>
#include <vector>

#include "nrldpc_common.cuh"
#include "nrldpc_raterecover.cuh"

namespace nrldpc {
namespace {

constexpr int kRmThreads = 256;

// Exclusive block-wide prefix sum of one int per thread (kRmThreads threads); *total = sum of all.
__device__ __forceinline__ int block_exclusive_scan(int v, int *total)
{
    __shared__ int warp_sums[kRmThreads / 32];
    __shared__ int s_total;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int n = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += n;
    }
    if (lane == 31) warp_sums[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        int w = lane < kRmThreads / 32 ? warp_sums[lane] : 0;
#pragma unroll
        for (int o = 1; o < kRmThreads / 32; o <<= 1) {
            const int n = __shfl_up_sync(0xffffffffu, w, o);
            if (lane >= o) w += n;
        }
        if (lane < kRmThreads / 32) warp_sums[lane] = w;
        if (lane == kRmThreads / 32 - 1) s_total = w;
    }
    __syncthreads();
    *total = s_total;
    const int base = warp ? warp_sums[warp - 1] : 0;
    __syncthreads();
    return base + incl - v;
}

// Bit selection + bit interleaving.  The circular buffer is walked from k0; a position holding -1
// (filler) is skipped, whatever the pattern of fillers (:80-87); output bit k of the selection goes to
// fe[(k mod E/Qm) * Qm + k div (E/Qm)] (:90-93).
__global__ void __launch_bounds__(kRmThreads)
ratematch_kernel(const int8_t *__restrict__ dn, int N, int Ncb, int k0, int Qm, const int32_t *__restrict__ E_of,
                 const long long *__restrict__ goff, int8_t *__restrict__ g)
{
    const int b = blockIdx.x, tid = threadIdx.x;
    const int8_t *d = dn + (size_t)b * N;
    const int E = E_of[b];
    if (E <= 0) return;
    int8_t *out = g + goff[b];
    const int per = (Ncb + kRmThreads - 1) / kRmThreads;
    const int t0 = min(tid * per, Ncb), t1 = min(t0 + per, Ncb);  // this thread's stretch of the walk
    int cnt = 0;
    for (int t = t0; t < t1; ++t) {
        int pos = k0 + t;
        if (pos >= Ncb) pos -= Ncb;
        cnt += (d[pos] != -1);
    }
    int S;
    int rank = block_exclusive_scan(cnt, &S);
    if (S <= 0) return;  // nothing but fillers: the reference would not terminate
    const int cols = E / Qm;
    for (int t = t0; t < t1; ++t) {
        int pos = k0 + t;
        if (pos >= Ncb) pos -= Ncb;
        const int8_t v = d[pos];
        if (v == -1) continue;
        for (int k = rank; k < E; k += S) {  // repetition when E exceeds the number of transmittable bits
            const int q = k / cols, e = k - q * cols;
            out[(size_t)e * Qm + q] = v;
        }
        ++rank;
    }
}

// Rate recovery of codeblock blockIdx.x: the device function shared with the decoder kernels' fused LLR load
// (nrldpc_raterecover.cuh).
template <typename TIn, typename TOut>
__global__ void __launch_bounds__(kRmThreads)
raterecover_kernel(const TIn *__restrict__ llr, int N, int Ncb, int k0, int Qm, int F0, int F1,
                   const int32_t *__restrict__ E_of, const long long *__restrict__ goff, TOut *__restrict__ outp)
{
    __shared__ double s_red[33];
    const int b = blockIdx.x;
    TOut *out = outp + (size_t)b * N;
    if constexpr (sizeof(TOut) == 8)
        rr_codeblock<TIn>(llr + goff[b], E_of[b], N, Ncb, k0, Qm, F0, F1, nullptr, reinterpret_cast<double *>(out), nullptr, s_red);
    else
        rr_codeblock<TIn>(llr + goff[b], E_of[b], N, Ncb, k0, Qm, F0, F1, nullptr, nullptr, reinterpret_cast<float *>(out), s_red);
}

// HARQ soft combining (py5gphy/nr_pdsch/nr_dlsch_decode.py:80-87): a zero on either side means "not
// received there" -> plain sum, otherwise the average of the two transmissions.
__global__ void harq_combine_kernel(const double *__restrict__ nw, const double *__restrict__ cur, long long n,
                                    double *__restrict__ out)
{
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const double a = nw[i], c = cur[i];
        out[i] = (a == 0.0 || c == 0.0) ? a + c : (a + c) / 2.0;
    }
}

int check_rm_args(const char *what, int B, int N, int Ncb, int k0, int Qm)
{
    if (B < 0 || N <= 0 || Ncb <= 0 || Ncb > N || k0 < 0 || k0 >= Ncb || Qm <= 0) {
        set_error("%s: bad argument (B=%d N=%d Ncb=%d k0=%d Qm=%d)", what, B, N, Ncb, k0, Qm);
        return NRLDPC_EINVAL;
    }
    return NRLDPC_OK;
}

using DevMem = ScratchBuf;

// offsets of the concatenated per-codeblock sequences (code block concatenation, 38.212 5.5)
int host_offsets(const char *what, const int32_t *E, int B, int Qm, std::vector<long long> *off, long long *total)
{
    off->resize(B > 0 ? B : 1);
    long long t = 0;
    for (int b = 0; b < B; ++b) {
        if (E[b] < 0 || E[b] % Qm) { set_error("%s: E[%d]=%d must be a non-negative multiple of Qm=%d", what, b, E[b], Qm); return NRLDPC_EINVAL; }
        (*off)[b] = t;
        t += E[b];
    }
    *total = t;
    return NRLDPC_OK;
}

}  // namespace
}  // namespace nrldpc

using namespace nrldpc;

extern "C" int nrldpc_ratematch(const int8_t *d_dn, int B, int N, int Ncb, int k0, int Qm, const int32_t *d_E,
                                const long long *d_goff, int8_t *d_g, void *stream)
{
    if (int rc = check_rm_args("ratematch", B, N, Ncb, k0, Qm)) return rc;
    if (!d_dn || !d_E || !d_goff || !d_g) { set_error("ratematch: null pointer"); return NRLDPC_EINVAL; }
    if (B == 0) return NRLDPC_OK;
    ratematch_kernel<<<B, kRmThreads, 0, (cudaStream_t)stream>>>(d_dn, N, Ncb, k0, Qm, d_E, d_goff, d_g);
    NRLDPC_CUDA(cudaGetLastError());
    return NRLDPC_OK;
}

extern "C" int nrldpc_ratematch_host(const int8_t *dn, int B, int N, int Ncb, int k0, int Qm, const int32_t *E, int8_t *g)
{
    if (int rc = check_rm_args("ratematch", B, N, Ncb, k0, Qm)) return rc;
    if (B == 0) return NRLDPC_OK;
    if (!dn || !E || !g) { set_error("ratematch: null pointer"); return NRLDPC_EINVAL; }
    std::vector<long long> off;
    long long total = 0;
    if (int rc = host_offsets("ratematch", E, B, Qm, &off, &total)) return rc;
    DevMem d_dn, d_E, d_off, d_g;
    NRLDPC_CUDA(d_dn.alloc((size_t)B * N));
    NRLDPC_CUDA(d_E.alloc((size_t)B * 4));
    NRLDPC_CUDA(d_off.alloc((size_t)B * 8));
    NRLDPC_CUDA(d_g.alloc((size_t)total));
    NRLDPC_CUDA(cudaMemcpy(d_dn.p, dn, (size_t)B * N, cudaMemcpyHostToDevice));
    NRLDPC_CUDA(cudaMemcpy(d_E.p, E, (size_t)B * 4, cudaMemcpyHostToDevice));
    NRLDPC_CUDA(cudaMemcpy(d_off.p, off.data(), (size_t)B * 8, cudaMemcpyHostToDevice));
    if (int rc = nrldpc_ratematch((const int8_t *)d_dn.p, B, N, Ncb, k0, Qm, (const int32_t *)d_E.p,
                                  (const long long *)d_off.p, (int8_t *)d_g.p, nullptr)) return rc;
    NRLDPC_CUDA(cudaMemcpy(g, d_g.p, (size_t)total, cudaMemcpyDeviceToHost));
    return NRLDPC_OK;
}

extern "C" int nrldpc_raterecover(const void *d_llr_g, int in_f64, int B, int N, int Ncb, int k0, int Qm, int F0, int F1,
                                  const int32_t *d_E, const long long *d_goff, void *d_out, int out_f64, void *stream)
{
    if (int rc = check_rm_args("raterecover", B, N, Ncb, k0, Qm)) return rc;
    if (!d_llr_g || !d_E || !d_goff || !d_out) { set_error("raterecover: null pointer"); return NRLDPC_EINVAL; }
    if (B == 0) return NRLDPC_OK;
    cudaStream_t s = (cudaStream_t)stream;
    if (in_f64 && out_f64)
        raterecover_kernel<double, double><<<B, kRmThreads, 0, s>>>((const double *)d_llr_g, N, Ncb, k0, Qm, F0, F1, d_E, d_goff, (double *)d_out);
    else if (in_f64)
        raterecover_kernel<double, float><<<B, kRmThreads, 0, s>>>((const double *)d_llr_g, N, Ncb, k0, Qm, F0, F1, d_E, d_goff, (float *)d_out);
    else if (out_f64)
        raterecover_kernel<float, double><<<B, kRmThreads, 0, s>>>((const float *)d_llr_g, N, Ncb, k0, Qm, F0, F1, d_E, d_goff, (double *)d_out);
    else
        raterecover_kernel<float, float><<<B, kRmThreads, 0, s>>>((const float *)d_llr_g, N, Ncb, k0, Qm, F0, F1, d_E, d_goff, (float *)d_out);
    NRLDPC_CUDA(cudaGetLastError());
    return NRLDPC_OK;
}

extern "C" int nrldpc_raterecover_host(const void *llr_g, int in_f64, int B, int N, int Ncb, int k0, int Qm, int Zc, int K_apo,
                                       int K, const int32_t *E, void *out, int out_f64)
{
    if (int rc = check_rm_args("raterecover", B, N, Ncb, k0, Qm)) return rc;
    if (B == 0) return NRLDPC_OK;
    if (!llr_g || !E || !out || K_apo > K) { set_error("raterecover: bad argument"); return NRLDPC_EINVAL; }
    std::vector<long long> off;
    long long total = 0;
    if (int rc = host_offsets("raterecover", E, B, Qm, &off, &total)) return rc;
    const size_t isz = in_f64 ? 8 : 4, osz = out_f64 ? 8 : 4;
    DevMem d_in, d_E, d_off, d_out;
    NRLDPC_CUDA(d_in.alloc((size_t)total * isz));
    NRLDPC_CUDA(d_E.alloc((size_t)B * 4));
    NRLDPC_CUDA(d_off.alloc((size_t)B * 8));
    NRLDPC_CUDA(d_out.alloc((size_t)B * N * osz));
    NRLDPC_CUDA(cudaMemcpy(d_in.p, llr_g, (size_t)total * isz, cudaMemcpyHostToDevice));
    NRLDPC_CUDA(cudaMemcpy(d_E.p, E, (size_t)B * 4, cudaMemcpyHostToDevice));
    NRLDPC_CUDA(cudaMemcpy(d_off.p, off.data(), (size_t)B * 8, cudaMemcpyHostToDevice));
    if (int rc = nrldpc_raterecover(d_in.p, in_f64, B, N, Ncb, k0, Qm, K_apo - 2 * Zc, K - 2 * Zc, (const int32_t *)d_E.p,
                                    (const long long *)d_off.p, d_out.p, out_f64, nullptr)) return rc;
    NRLDPC_CUDA(cudaMemcpy(out, d_out.p, (size_t)B * N * osz, cudaMemcpyDeviceToHost));
    return NRLDPC_OK;
}

extern "C" int nrldpc_harq_combine(const double *d_new, const double *d_cur, long long count, double *d_out, void *stream)
{
    if (count < 0 || !d_new || !d_cur || !d_out) { set_error("harq_combine: bad argument"); return NRLDPC_EINVAL; }
    if (count == 0) return NRLDPC_OK;
    const long long blocks = (count + 255) / 256;
    harq_combine_kernel<<<(int)(blocks < 148 * 16 ? blocks : 148 * 16), 256, 0, (cudaStream_t)stream>>>(d_new, d_cur, count, d_out);
    NRLDPC_CUDA(cudaGetLastError());
    return NRLDPC_OK;
}

extern "C" int nrldpc_harq_combine_host(const double *nw, const double *cur, long long count, double *out)
{
    if (count < 0 || (count && (!nw || !cur || !out))) { set_error("harq_combine: bad argument"); return NRLDPC_EINVAL; }
    if (count == 0) return NRLDPC_OK;
    DevMem a, c, o;
    NRLDPC_CUDA(a.alloc((size_t)count * 8));
    NRLDPC_CUDA(c.alloc((size_t)count * 8));
    NRLDPC_CUDA(o.alloc((size_t)count * 8));
    NRLDPC_CUDA(cudaMemcpy(a.p, nw, (size_t)count * 8, cudaMemcpyHostToDevice));
    NRLDPC_CUDA(cudaMemcpy(c.p, cur, (size_t)count * 8, cudaMemcpyHostToDevice));
    if (int rc = nrldpc_harq_combine((const double *)a.p, (const double *)c.p, count, (double *)o.p, nullptr)) return rc;
    NRLDPC_CUDA(cudaMemcpy(out, o.p, (size_t)count * 8, cudaMemcpyDeviceToHost));
    return NRLDPC_OK;
}
