// nrldpc_bf_qc.cu -- bit-flipping decoder on the quasi-cyclic 5G matrices, bit-packed, sm_100a.
//
// ldpc_decoder_BF (py5gphy/ldpc/ldpc_decoder_bit_flipping.py:5-73) as reached through
// nr_decode_ldpc(..., algo='BF') (py5gphy/ldpc/nr_ldpc_decode.py:43,65-67), i.e. on getH(Zc, bgn, iLS) with 2*Zc
// zero LLRs prepended.  Only the sign of an LLR is used, so a codeblock is N' bits: every Zc-bit column-block of
// the hard decisions and every Zc-bit row-block of the syndrome is a row of ceil(Zc/32) words in shared memory,
// and a circulant block with shift P is a bit rotation (check r of row-block i reads variable (r + P) mod Zc of
// column-block j, ldpc_info.py:126-137):
//   S = H ck mod 2 (:47)            one thread per syndrome word: XOR of rotated hard-decision words;
//   En = (2S - 1) @ H (:61)         one thread per word of 32 variables: the number u of unsatisfied checks of each
//                                   variable is counted in five bit planes (ripple-carry add of the rotated syndrome
//                                   words, column degree <= 30), En = 2u - degree; the plane scan from the top gives
//                                   the word's maximum and the mask of the variables that reach it;
//   flip En == max(En) (:62-70)     block maximum by redux.sync + one shared atomicMax per warp, then every word
//                                   whose maximum equals it flips its mask.
// One CTA per codeblock, three barriers per iteration; HBM traffic is the LLR read and the int8 ck write.
#include <algorithm>

#include "nrldpc_common.cuh"
#include "nrldpc_bits.cuh"

namespace nrldpc {

namespace {

constexpr int kBfMaxThreads = 256;

template <typename T>
__global__ void __launch_bounds__(kBfMaxThreads)
bf_qc_kernel(const __grid_constant__ QcCfg cfg, const T *__restrict__ llr, int B, int max_iter, int vec,
             int8_t *__restrict__ ck_out, uint8_t *__restrict__ status, int32_t *__restrict__ iters)
{
    extern __shared__ uint32_t smem[];
    const int Zc = cfg.Zc, W = cfg.tiles, Wp = W + 1, ncols = cfg.ncols, nrows = cfg.nrows;
    uint32_t *CK = smem;                  // [ncols][Wp] hard decisions (pad word stays zero)
    uint32_t *SW = CK + ncols * Wp;       // [nrows][Wp] syndrome
    uint32_t *MK = SW + nrows * Wp;       // [ncols][W]  variables of the word that reach the word's maximum
    int *EN = (int *)(MK + ncols * W);    // [ncols][W]  the word's maximum of En
    __shared__ int s_max;

    const int cb = blockIdx.x, tid = threadIdx.x, nthr = blockDim.x;
    const int lane = tid & 31, warp = tid >> 5, nwarps = nthr >> 5;
    for (int t = tid; t < ncols * Wp + nrows * Wp; t += nthr) smem[t] = 0;
    if (tid == 0) s_max = -128;
    __syncthreads();

    // :41-43  LLR>0 -> 0, LLR<0 -> 1, LLR==0 stays 0; the 2*Zc prepended zeros (nr_ldpc_decode.py:43) decide 0
    const T *x = llr + (size_t)cb * cfg.N;
    for (int t = warp; t < (ncols - 2) * W; t += nwarps) {
        const int jo = t / W, w = t - jo * W, r = 32 * w + lane;
        const bool neg = (r < Zc) && (x[jo * Zc + r] < (T)0);
        const uint32_t bits = __ballot_sync(0xffffffffu, neg);
        if (lane == 0) CK[(jo + 2) * Wp + w] = bits;
    }
    __syncthreads();

    int ok = 0, it = 0;
    for (; it < max_iter; ++it) {
        // :47 S = H ck mod 2
        uint32_t any = 0;
        for (int t = tid; t < nrows * W; t += nthr) {
            const int i = t / W, w = t - i * W;
            uint32_t acc = 0;
            for (int e = cfg.rowptr[i]; e < cfg.rowptr[i + 1]; ++e) {
                const uint32_t ed = cfg.edge[e];
                acc ^= rot_word(CK + (ed & 0xffu) * Wp, (int)(ed >> 8), w, Zc);
            }
            SW[i * Wp + w] = acc;
            any |= acc;
        }
        if (!__syncthreads_or(any != 0)) { ok = 1; break; }  // :50-56

        // :61-62 En and its maximum over all N' variables
        int mymax = -128;
        for (int t = tid; t < ncols * W; t += nthr) {
            const int j = t / W, w = t - j * W;
            const int nv = Zc - 32 * w;
            const uint32_t valid = nv < 32 ? (1u << nv) - 1u : 0xffffffffu;
            uint32_t mask = valid;
            int u = 0, deg = 1;
            if (j < cfg.ncore) {
                uint32_t c0 = 0, c1 = 0, c2 = 0, c3 = 0, c4 = 0;
                const int q0 = cfg.colptr[j], q1 = cfg.colptr[j + 1];
                deg = q1 - q0;
                for (int q = q0; q < q1; ++q) {
                    const uint32_t ce = cfg.centry[q];
                    const uint32_t b = rot_word(SW + (ce & 0x3fu) * Wp, (int)(ce >> 16), w, Zc);  // S[(c + back) mod Zc]
                    const uint32_t t0 = c0 & b;  c0 ^= b;
                    const uint32_t t1 = c1 & t0; c1 ^= t0;
                    const uint32_t t2 = c2 & t1; c2 ^= t1;
                    const uint32_t t3 = c3 & t2; c3 ^= t2;
                    c4 ^= t3;
                }
                uint32_t m;
                m = mask & c4; if (m) { mask = m; u |= 16; }
                m = mask & c3; if (m) { mask = m; u |= 8; }
                m = mask & c2; if (m) { mask = m; u |= 4; }
                m = mask & c1; if (m) { mask = m; u |= 2; }
                m = mask & c0; if (m) { mask = m; u |= 1; }
            } else {
                const uint32_t b = SW[(j - cfg.ncore + 4) * Wp + w];  // degree-1 extension column of that row-block, shift 0
                if (b) { mask = b; u = 1; }
            }
            const int en = 2 * u - deg;
            MK[t] = mask;
            EN[t] = en;
            mymax = max(mymax, en);
        }
        mymax = __reduce_max_sync(0xffffffffu, mymax);
        if (lane == 0) atomicMax(&s_max, mymax);
        __syncthreads();
        const int mx = s_max;
        // :67-70 flip every bit whose metric equals the maximum
        for (int t = tid; t < ncols * W; t += nthr) {
            const int j = t / W, w = t - j * W;
            if (EN[t] == mx) CK[j * Wp + w] ^= MK[t];
        }
        __syncthreads();
        if (tid == 0) s_max = -128;
    }

    int8_t *out = ck_out + (size_t)cb * cfg.Nfull;
    if (vec) {  // Zc % 16 == 0, 16-byte aligned rows: 16 decisions -> one 128-bit store
        const int H = Zc >> 4;
        for (int t = tid; t < ncols * H; t += nthr) {
            const int j = t / H, h = t - j * H;
            const uint32_t bits = reinterpret_cast<const uint16_t *>(CK + j * Wp)[h];
            uint32_t o[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) o[q] = (((bits >> (4 * q)) & 0xfu) * 0x00204081u) & 0x01010101u;
            *reinterpret_cast<uint4 *>(out + j * Zc + 16 * h) = make_uint4(o[0], o[1], o[2], o[3]);
        }
    } else {
        for (int t = warp; t < ncols * W; t += nwarps) {
            const int j = t / W, w = t - j * W, r = 32 * w + lane;
            if (r < Zc) out[j * Zc + r] = (int8_t)((CK[j * Wp + w] >> lane) & 1u);
        }
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
    const int W = cfg.tiles, Wp = W + 1;
    const int threads = std::min(kBfMaxThreads, (cfg.ncols * W + 31) / 32 * 32);
    const size_t smem = (size_t)(cfg.ncols * Wp + cfg.nrows * Wp + 2 * cfg.ncols * W) * 4;  // 12.5 KB at BG1 Zc=384
    const int vec = (cfg.Zc % 16 == 0) && (reinterpret_cast<uintptr_t>(d_ck) % 16 == 0);
    if (is_f64)
        bf_qc_kernel<double><<<B, threads, smem, s>>>(cfg, (const double *)d_llr, B, max_iter, vec, d_ck, d_status, d_iters);
    else
        bf_qc_kernel<float><<<B, threads, smem, s>>>(cfg, (const float *)d_llr, B, max_iter, vec, d_ck, d_status, d_iters);
    NRLDPC_CUDA(cudaGetLastError());
    return NRLDPC_OK;
}

}  // namespace nrldpc
