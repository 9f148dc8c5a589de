// nrldpc_bf_qc.cu -- bit-flipping decoder on the quasi-cyclic 5G matrices, sm_100a.
//
// ldpc_decoder_BF (py5gphy/ldpc/ldpc_decoder_bit_flipping.py:5-73) as reached through
// nr_decode_ldpc(..., algo='BF') (py5gphy/ldpc/nr_ldpc_decode.py:43,65-67), i.e. on getH(Zc, bgn, iLS) with 2*Zc
// zero LLRs prepended.  One CTA per codeblock; the whole state (hard bits, syndrome, flip metric) is resident in
// shared memory as bytes and the circulant structure replaces the CSR/CSC index arrays of the generic kernel:
// check r of row-block i reads variable (r + P) mod Zc of column-block j (ldpc_info.py:126-137).  Only the sign of
// an LLR is used, so HBM traffic is the LLR read and the int8 ck write.
//
// Thread mapping: t = ty * RT + tr; tr walks the lifted index (stride RT = Zc rounded up to a warp, or to a power
// of two below 32), ty walks the row-blocks / column-blocks (RY of them in flight).
#include "nrldpc_common.cuh"

namespace nrldpc {

namespace {

template <typename T>
__global__ void __launch_bounds__(384)
bf_qc_kernel(const __grid_constant__ QcCfg cfg, const T *__restrict__ llr, int B, int max_iter, int RT,
             int8_t *__restrict__ ck_out, uint8_t *__restrict__ status, int32_t *__restrict__ iters)
{
    extern __shared__ __align__(16) unsigned char smem[];
    const int Zc = cfg.Zc, Nfull = cfg.Nfull, M = cfg.M, N = cfg.N;
    unsigned char *ck = smem;                  // [ncols][Zc] hard decisions
    signed char *En = (signed char *)(ck + Nfull);  // [ncols][Zc] #unsatisfied - #satisfied checks (:61)
    unsigned char *S = (unsigned char *)(En + Nfull);  // [nrows][Zc] syndrome (:47)
    __shared__ int s_max;

    const int cb = blockIdx.x, tid = threadIdx.x;
    const int ty = tid / RT, tr = tid - ty * RT, RY = blockDim.x / RT;
    if (tid == 0) s_max = -128;

    // :41-43  LLR>0 -> 0, LLR<0 -> 1, LLR==0 stays 0; the 2*Zc prepended zeros (nr_ldpc_decode.py:43) decide 0
    const T *x = llr + (size_t)cb * N;
    for (int n = tid; n < Nfull; n += blockDim.x) ck[n] = (n >= 2 * Zc && x[n - 2 * Zc] < (T)0) ? 1 : 0;
    __syncthreads();

    int ok = 0, it = 0;
    for (; it < max_iter; ++it) {
        // :47 S = H ck mod 2
        int any = 0;
        for (int i = ty; i < cfg.nrows; i += RY) {
            const int e0 = cfg.rowptr[i], e1 = cfg.rowptr[i + 1];
            for (int r = tr; r < Zc; r += RT) {
                unsigned p = 0;
                for (int e = e0; e < e1; ++e) {
                    const unsigned w = cfg.edge[e];
                    int c = r + (int)(w >> 8);
                    c -= (c >= Zc) ? Zc : 0;
                    p ^= ck[(w & 0xffu) * Zc + c];
                }
                S[i * Zc + r] = (unsigned char)p;
                any |= (int)p;
            }
        }
        if (!__syncthreads_or(any)) { ok = 1; break; }  // :50-56

        // :61-62 En = (2S-1) @ H and its maximum over all N' variables
        int mymax = -128;
        for (int j = ty; j < cfg.ncols; j += RY) {
            if (j < cfg.ncore) {
                const int q0 = cfg.colptr[j], q1 = cfg.colptr[j + 1];
                for (int c = tr; c < Zc; c += RT) {
                    int u = 0;
                    for (int q = q0; q < q1; ++q) {
                        const unsigned w = cfg.centry[q];
                        int r = c + (int)(w >> 16);
                        r -= (r >= Zc) ? Zc : 0;
                        u += S[(w & 0x3fu) * Zc + r];
                    }
                    const int en = 2 * u - (q1 - q0);
                    En[j * Zc + c] = (signed char)en;
                    mymax = max(mymax, en);
                }
            } else {
                const int i = j - cfg.ncore + 4;  // degree-1 extension column of row-block i, shift 0
                for (int c = tr; c < Zc; c += RT) {
                    const int en = 2 * (int)S[i * Zc + c] - 1;
                    En[j * Zc + c] = (signed char)en;
                    mymax = max(mymax, en);
                }
            }
        }
        mymax = __reduce_max_sync(0xffffffffu, mymax);
        if ((tid & 31) == 0) atomicMax(&s_max, mymax);
        __syncthreads();
        const int mx = s_max;
        // :67-70 flip every bit whose metric equals the maximum
        for (int n = tid; n < Nfull; n += blockDim.x)
            if ((int)En[n] == mx) ck[n] ^= 1;
        __syncthreads();
        if (tid == 0) s_max = -128;
    }

    int8_t *out = ck_out + (size_t)cb * Nfull;
    if ((Nfull & 3) == 0) {
        for (int n = tid; n < Nfull / 4; n += blockDim.x) ((uint32_t *)out)[n] = ((const uint32_t *)ck)[n];
    } else {
        for (int n = tid; n < Nfull; n += blockDim.x) out[n] = (int8_t)ck[n];
    }
    if (tid == 0) {
        if (status) status[cb] = (uint8_t)ok;
        if (iters) iters[cb] = it;
    }
}

}  // namespace

int launch_bf_qc(const QcCfg &cfg, const void *d_llr, int is_f64, int B, int max_iter, int8_t *d_ck,
                 uint8_t *d_status, int32_t *d_iters, cudaStream_t s)
{
    if (B <= 0) return NRLDPC_OK;
    int RT;
    if (cfg.Zc >= 32) {
        RT = (cfg.Zc + 31) / 32 * 32;
    } else {
        RT = 1;
        while (RT < cfg.Zc) RT <<= 1;
    }
    const int RY = RT >= 256 ? 1 : 256 / RT;
    const int threads = RT * RY;  // 256 for Zc <= 128, else Zc rounded up to a warp (<= 384)
    const size_t smem = (size_t)2 * cfg.Nfull + cfg.M;
    if (is_f64) {
        NRLDPC_CUDA(cudaFuncSetAttribute(bf_qc_kernel<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, 72 * 1024));
        bf_qc_kernel<double><<<B, threads, smem, s>>>(cfg, (const double *)d_llr, B, max_iter, RT, d_ck, d_status, d_iters);
    } else {
        NRLDPC_CUDA(cudaFuncSetAttribute(bf_qc_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, 72 * 1024));
        bf_qc_kernel<float><<<B, threads, smem, s>>>(cfg, (const float *)d_llr, B, max_iter, RT, d_ck, d_status, d_iters);
    }
    NRLDPC_CUDA(cudaGetLastError());
    return NRLDPC_OK;
}

}  // namespace nrldpc
