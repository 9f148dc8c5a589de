// nrldpc_generic.cu -- decoders for an ARBITRARY parity-check matrix given as CSR/CSC, sm_100a.
//
// These serve the reference entry points that take a dense H (decode_ldpc,
// py5gphy/ldpc/nr_ldpc_decode.py:51-143; ldpc_decoder_BF, ldpc_decoder_bit_flipping.py:5-73) and,
// instantiated in float64, give a device path that reproduces the reference's float64 arithmetic
// exactly (used to validate the fp32 hot kernel at scale).  One CTA per codeblock, message state in
// a global-memory workspace; no attempt at speed -- the quasi-cyclic kernel is the hot path.
#include "nrldpc_common.cuh"

namespace nrldpc {

namespace {

constexpr int kGenThreads = 512;

template <typename T> struct Inf;
template <> struct Inf<float> { static __device__ float v() { return __uint_as_float(0x7f800000u); } };
template <> struct Inf<double> { static __device__ double v() { return __longlong_as_double(0x7ff0000000000000LL); } };

__device__ __forceinline__ float add_rn(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ double add_rn(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ float sub_rn(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ double sub_rn(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ float mul_rn(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ double mul_rn(double a, double b) { return __dmul_rn(a, b); }

// Workspace per codeblock: Lr[E] then LQ[Nv] then llr_full[Nv]  (T)
template <typename T>
__global__ void __launch_bounds__(kGenThreads)
soft_csr_kernel(const T *__restrict__ llr, int B, int M, int Nv, int E, const int32_t *__restrict__ rowptr,
                const int32_t *__restrict__ colidx, const int32_t *__restrict__ cptr,
                const int32_t *__restrict__ cedge, int prepend, int max_iter, int algo, T alpha, T beta,
                int early_term, T *__restrict__ work, int8_t *__restrict__ ck, uint8_t *__restrict__ status,
                int32_t *__restrict__ iters)
{
    const int cb = blockIdx.x, tid = threadIdx.x;
    T *Lr = work + (size_t)cb * ((size_t)E + 2 * (size_t)Nv);
    T *LQ = Lr + E;
    T *L0 = LQ + Nv;
    const int Nin = Nv - prepend;
    __shared__ int s_any;

    // :94-101  LQ = LLRin, Lr = 0 (and the 5G wrapper's prepended zeros, :43)
    for (int n = tid; n < Nv; n += kGenThreads) {
        T v = (n < prepend) ? (T)0 : add_rn(llr[(size_t)cb * Nin + (n - prepend)], (T)0);
        L0[n] = v;
        LQ[n] = v;
    }
    for (int e = tid; e < E; e += kGenThreads) Lr[e] = (T)0;
    if (tid == 0) s_any = 0;
    __syncthreads();

    bool done = false;
    int it = 0;
    for (; it < max_iter; ++it) {
        // :107-114 syndrome of LQ<0
        int any = 0;
        for (int m = tid; m < M; m += kGenThreads) {
            int p = 0;
            for (int e = rowptr[m]; e < rowptr[m + 1]; ++e) p ^= (LQ[colidx[e]] < (T)0);
            any |= p;
        }
        if (any) s_any = 1;
        __syncthreads();
        const int bad = s_any;
        __syncthreads();
        if (tid == 0) s_any = 0;
        if (!bad && early_term) { done = true; break; }

        // :117-123 check rows from the old Lq = LQ - Lr
        for (int m = tid; m < M; m += kGenThreads) {
            const int e0 = rowptr[m], e1 = rowptr[m + 1];
            if (algo == NRLDPC_ALGO_MINSUM) {
                // _min_sum_process :178-227 via (min1, min2, first argmin, sign product), sign(0) := +
                T m1 = Inf<T>::v(), m2 = Inf<T>::v();
                int idx = -1, neg = 0;
                for (int e = e0; e < e1; ++e) {
                    const T q = sub_rn(LQ[colidx[e]], Lr[e]);
                    const T aq = q < (T)0 ? -q : q;
                    neg ^= (q < (T)0);
                    if (aq < m1) { m2 = m1; m1 = aq; idx = e; }
                    else if (aq < m2) m2 = aq;
                }
                T s1 = sub_rn(m1, beta); s1 = s1 > (T)0 ? s1 : (T)0;
                T s2 = sub_rn(m2, beta); s2 = s2 > (T)0 ? s2 : (T)0;
                const T mag1 = mul_rn(alpha, s1), mag2 = mul_rn(alpha, s2);
                for (int e = e0; e < e1; ++e) {
                    const T q = sub_rn(LQ[colidx[e]], Lr[e]);
                    const T mag = (e == idx) ? mag2 : mag1;
                    Lr[e] = (neg ^ (q < (T)0)) ? -mag : mag;
                }
            } else {
                // _BP_process :145-176 (tanh rule; device libm, so not bit-identical to NumPy)
                int nz = 0, zi = -1;
                double prod = 1.0;
                for (int e = e0; e < e1; ++e) {
                    const double q = (double)sub_rn(LQ[colidx[e]], Lr[e]);
                    if (q == 0.0) { ++nz; if (zi < 0) zi = e; }
                    else prod *= tanh(q / 2);
                }
                for (int e = e0; e < e1; ++e) {
                    const double q = (double)sub_rn(LQ[colidx[e]], Lr[e]);
                    double out = 0.0;
                    if (nz == 0) {
                        const double x = prod / tanh(q / 2);
                        out = x >= 1.0 ? 2 * 19.07 : (x <= -1.0 ? -2 * 19.07 : 2 * atanh(x));
                    } else if (nz == 1 && e == zi) {
                        out = prod;  // :170 the reference omits 2*atanh here
                    }
                    Lr[e] = (T)out;
                }
            }
        }
        __syncthreads();
        // :126 LQ = LLRin + Lr.sum(axis=0), ascending check index
        for (int n = tid; n < Nv; n += kGenThreads) {
            T s = (T)0;
            for (int q = cptr[n]; q < cptr[n + 1]; ++q) s = add_rn(s, Lr[cedge[q]]);
            LQ[n] = add_rn(L0[n], s);
        }
        __syncthreads();
    }
    // outputs
    int ok = 1;
    if (done) {
        for (int n = tid; n < Nv; n += kGenThreads) ck[(size_t)cb * Nv + n] = LQ[n] < (T)0 ? 1 : 0;
    } else {
        // :134-143 final tie rule LQ<=0 -> 1
        int any = 0;
        for (int m = tid; m < M; m += kGenThreads) {
            int p = 0;
            for (int e = rowptr[m]; e < rowptr[m + 1]; ++e) p ^= (LQ[colidx[e]] <= (T)0);
            any |= p;
        }
        if (any) s_any = 1;
        __syncthreads();
        ok = !s_any;
        for (int n = tid; n < Nv; n += kGenThreads) ck[(size_t)cb * Nv + n] = LQ[n] <= (T)0 ? 1 : 0;
    }
    if (tid == 0) {
        if (status) status[cb] = (uint8_t)ok;
        if (iters) iters[cb] = it;
    }
}

// ldpc_decoder_BF: py5gphy/ldpc/ldpc_decoder_bit_flipping.py:41-73.  Workspace per codeblock: int32 S[M].
__global__ void __launch_bounds__(kGenThreads)
bf_csr_kernel(const double *__restrict__ llr, int B, int M, int Nv, const int32_t *__restrict__ rowptr,
              const int32_t *__restrict__ colidx, const int32_t *__restrict__ cptr, const int32_t *__restrict__ cedge_row,
              int prepend, int max_iter, int32_t *__restrict__ work, int8_t *__restrict__ ck_out,
              uint8_t *__restrict__ status, int32_t *__restrict__ iters)
{
    const int cb = blockIdx.x, tid = threadIdx.x;
    int32_t *S = work + (size_t)cb * M;
    int8_t *ck = ck_out + (size_t)cb * Nv;
    const int Nin = Nv - prepend;
    __shared__ int s_any, s_max;
    // :41-43  LLR>0 -> 0, LLR<0 -> 1, LLR==0 stays 0
    for (int n = tid; n < Nv; n += kGenThreads)
        ck[n] = (n >= prepend && llr[(size_t)cb * Nin + (n - prepend)] < 0.0) ? 1 : 0;
    if (tid == 0) { s_any = 0; s_max = -0x7fffffff; }
    __syncthreads();
    int ok = 0, it = 0;
    for (; it < max_iter; ++it) {
        int any = 0;
        for (int m = tid; m < M; m += kGenThreads) {  // :47 S = H ck mod 2
            int p = 0;
            for (int e = rowptr[m]; e < rowptr[m + 1]; ++e) p ^= ck[colidx[e]];
            S[m] = p;
            any |= p;
        }
        if (any) s_any = 1;
        __syncthreads();
        const int bad = s_any;
        if (!bad) { ok = 1; break; }
        // :61 En = (2S-1) @ H ; :62 max
        int mymax = -0x7fffffff;
        for (int n = tid; n < Nv; n += kGenThreads) {
            int en = 0;
            for (int q = cptr[n]; q < cptr[n + 1]; ++q) en += 2 * S[cedge_row[q]] - 1;
            mymax = max(mymax, en);
        }
        atomicMax(&s_max, mymax);
        __syncthreads();
        const int mx = s_max;
        // :67-70 flip every bit with En == max
        for (int n = tid; n < Nv; n += kGenThreads) {
            int en = 0;
            for (int q = cptr[n]; q < cptr[n + 1]; ++q) en += 2 * S[cedge_row[q]] - 1;
            if (en == mx) ck[n] = (int8_t)(1 - ck[n]);
        }
        __syncthreads();
        if (tid == 0) { s_any = 0; s_max = -0x7fffffff; }
        __syncthreads();
    }
    if (tid == 0) {
        if (status) status[cb] = (uint8_t)ok;
        if (iters) iters[cb] = it;
    }
}

}  // namespace

template <typename T>
int launch_soft_csr(const T *d_llr, int B, int M, int Nv, int E, const int32_t *d_rowptr, const int32_t *d_colidx,
                    const int32_t *d_cptr, const int32_t *d_cedge, int prepend, int max_iter, int algo, T alpha,
                    T beta, int early_term, T *d_work, int8_t *d_ck, uint8_t *d_status, int32_t *d_iters,
                    cudaStream_t s)
{
    if (B <= 0) return NRLDPC_OK;
    soft_csr_kernel<T><<<B, kGenThreads, 0, s>>>(d_llr, B, M, Nv, E, d_rowptr, d_colidx, d_cptr, d_cedge, prepend,
                                                 max_iter, algo, alpha, beta, early_term, d_work, d_ck, d_status,
                                                 d_iters);
    NRLDPC_CUDA(cudaGetLastError());
    return NRLDPC_OK;
}
template int launch_soft_csr<float>(const float *, int, int, int, int, const int32_t *, const int32_t *, const int32_t *,
                                    const int32_t *, int, int, int, float, float, int, float *, int8_t *, uint8_t *,
                                    int32_t *, cudaStream_t);
template int launch_soft_csr<double>(const double *, int, int, int, int, const int32_t *, const int32_t *,
                                     const int32_t *, const int32_t *, int, int, int, double, double, int, double *,
                                     int8_t *, uint8_t *, int32_t *, cudaStream_t);

int launch_bf_csr(const double *d_llr, int B, int M, int Nv, int E, const int32_t *d_rowptr, const int32_t *d_colidx,
                  const int32_t *d_cptr, const int32_t *d_cedge_row, int prepend, int max_iter, int32_t *d_work,
                  int8_t *d_ck, uint8_t *d_status, int32_t *d_iters, cudaStream_t s)
{
    (void)E;
    if (B <= 0) return NRLDPC_OK;
    bf_csr_kernel<<<B, kGenThreads, 0, s>>>(d_llr, B, M, Nv, d_rowptr, d_colidx, d_cptr, d_cedge_row, prepend,
                                            max_iter, d_work, d_ck, d_status, d_iters);
    NRLDPC_CUDA(cudaGetLastError());
    return NRLDPC_OK;
}

}  // namespace nrldpc
