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
#include <cstdlib>

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


// ---------------------------------------------------------------------------------------------------------------
// Word-parallel bit-flipping kernel for Zc % 32 == 0 (W = Zc/32 words per block row/column, no partial word).
// Same algorithm and bit-plane counters as bf_qc_kernel above; the mapping follows encode_words_kernel:
// thread <-> (codeblock g, word w, part p); the 8 parts of a CTA split the row-blocks (syndrome) and the
// column-blocks (metric) and all threads of a part walk ONE flat edge list from the constant bank
// (descriptor = byte offset << 16 | last-of-row/column << 5 | shift & 31: address = one LEA.HI, the descriptor
// is the funnel-shift amount).  Hard decisions of the core column-blocks and every syndrome row-block are stored
// twice back to back, so a rotation is 2 LDS + 1 SHF without wrap arithmetic.  Two edges enter the bit-plane
// counters through one full adder (2 LOP3) before the ripple.  A CTA holds G = 96/W codeblocks, each with its
// own maximum, zero-syndrome test and iteration count; a thread flips the words it evaluated itself.
constexpr int kBwParts = 8, kBwPart = 96, kBwThreads = kBwParts * kBwPart, kBwList = 64, kBwIds = 12;

struct BfWordArgs {
    int W, Zc, ncols, nrows, ncore, N, Nfull;
    int G, slot;  // codeblocks per CTA, words of shared memory per codeblock
    uint32_t mNW, mW;  // floor(2^32 / d) + 1 for d = (ncols - 2) W and d = W
    uint16_t nrow[kBwParts], ncol[kBwParts];
    uint8_t nrid[kBwParts], ncid[kBwParts];            // row-blocks / core column-blocks of a part (balanced by edge count)
    uint8_t rowid[kBwParts][kBwIds], colid[kBwParts][kBwIds];
    uint32_t rowl[kBwParts][kBwList];  // core-column edges of the part's row-blocks, in rowid order (offsets into CK2)
    uint32_t coll[kBwParts][kBwList];  // edges of the part's core column-blocks, in colid order  (offsets into SW2)
};

__device__ __forceinline__ uint32_t bw_rot(const uint32_t *v2w, uint32_t d)
{
    const uint32_t *p = reinterpret_cast<const uint32_t *>(reinterpret_cast<const char *>(v2w) + (d >> 16));
    return __funnelshift_r(p[0], p[1], d);
}

template <typename T>
__global__ void __launch_bounds__(kBwThreads, 2)
bf_words_kernel(const __grid_constant__ BfWordArgs a, const T *__restrict__ llr, int B, int max_iter,
                int8_t *__restrict__ ck_out, uint8_t *__restrict__ status, int32_t *__restrict__ iters)
{
    extern __shared__ uint32_t smem[];
    __shared__ int s_max[kBwPart], s_any[kBwPart];
    const int W = a.W, W2 = 2 * W, Zc = a.Zc, ncore = a.ncore, nrows = a.nrows, ncols = a.ncols, next = nrows - 4;
    // per codeblock: CK2[ncore][2W], CKX[next][W], SW2[nrows][2W], MK[ncols][W], EN[ncols][W] as int8
    const int oCKX = ncore * W2, oSW = oCKX + next * W, oMK = oSW + nrows * W2, oEN = oMK + ncols * W;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int cb0 = blockIdx.x * a.G, g_cnt = min(a.G, B - cb0);
    if (tid < kBwPart) { s_max[tid] = -128; s_any[tid] = 0; }

    // hard decisions (:41-43): LLR < 0 -> 1, else 0; the two punctured column-blocks decide 0
    for (int t = tid; t < g_cnt * 2 * W2; t += kBwThreads) {
        const int g = t / (2 * W2);
        smem[g * a.slot + (t - g * 2 * W2)] = 0;
    }
    {   // the CTA's LLRs are contiguous: word t <-> floats [32 t, 32 t + 32); a warp takes 8 words at a time with the
        // eight 128-byte loads issued back to back (the read is the only HBM traffic of the kernel besides ck)
        const int NW = (ncols - 2) * W, total = g_cnt * NW;
        const T *base = llr + (size_t)cb0 * a.N + lane;
        for (int t0 = 8 * warp; t0 < total; t0 += 8 * (kBwThreads / 32)) {
            T v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) v[u] = (t0 + u < total) ? base[(size_t)(t0 + u) * 32] : (T)0;
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const uint32_t bits = __ballot_sync(0xffffffffu, v[u] < (T)0);
                const int t = t0 + u;
                if (lane == 0 && t < total) {
                    const int g = __umulhi((uint32_t)t, a.mNW), rem = t - g * NW;
                    const int jo = W == 1 ? rem : (int)__umulhi((uint32_t)rem, a.mW), w = rem - jo * W, j = jo + 2;
                    uint32_t *cb = smem + g * a.slot;
                    if (j < ncore) cb[j * W2 + w] = cb[j * W2 + W + w] = bits;
                    else cb[oCKX + (j - ncore) * W + w] = bits;
                }
            }
        }
    }
    __syncthreads();

    const int part = tid / kBwPart, x = tid - part * kBwPart;
    const int g = x / W, w = x - g * W;
    bool act = g < g_cnt;
    uint32_t *cb = smem + g * a.slot;
    const uint32_t *CKw = cb + w, *SWw = cb + oSW + w;
    int8_t *EN = reinterpret_cast<int8_t *>(cb + oEN);
    const int nr = a.nrow[part], nc = a.ncol[part];
    int it_done = max_iter, ok = 0;

    for (int it = 0; it < max_iter; ++it) {
        // :47 S = H ck mod 2 for the row-blocks of this part
        if (act) {
            uint32_t acc = 0, any = 0;
            int k = 0;
            for (int e = 0; e < nr; ++e) {
                const uint32_t d = a.rowl[part][e];
                acc ^= bw_rot(CKw, d);
                if (d & 32u) {
                    const int i = a.rowid[part][k++];
                    if (i >= 4) acc ^= cb[oCKX + (i - 4) * W + w];  // the row-block's own degree-1 column, shift 0
                    cb[oSW + i * W2 + w] = acc;
                    cb[oSW + i * W2 + W + w] = acc;
                    any |= acc;
                    acc = 0;
                }
            }
            if (any) s_any[g] = 1;
        }
        __syncthreads();
        if (act && s_any[g] == 0) { act = false; ok = 1; it_done = it; }  // :50-56

        // :61-62 En of the words of this part's column-blocks, maximum per codeblock
        int mymax = -128;
        if (act) {
            int k = 0;
            uint32_t c0 = 0, c1 = 0, c2 = 0, c3 = 0, c4 = 0;
            int deg = 0;
            for (int e = 0; e < nc; ++e) {
                const uint32_t d = a.coll[part][e];
                uint32_t b = bw_rot(SWw, d), carry;
                ++deg;
                if (!(d & 32u)) {
                    // two edges through a full adder: sum into plane 0, carry into plane 1
                    const uint32_t d2 = a.coll[part][++e];
                    const uint32_t b2 = bw_rot(SWw, d2);
                    ++deg;
                    carry = (c0 & b) | (c0 & b2) | (b & b2);
                    c0 ^= b ^ b2;
                    const uint32_t t1 = c1 & carry; c1 ^= carry;
                    const uint32_t t2 = c2 & t1; c2 ^= t1;
                    const uint32_t t3 = c3 & t2; c3 ^= t2;
                    c4 ^= t3;
                    if (!(d2 & 32u)) continue;
                } else {
                    const uint32_t t0 = c0 & b; c0 ^= b;
                    const uint32_t t1 = c1 & t0; c1 ^= t0;
                    const uint32_t t2 = c2 & t1; c2 ^= t1;
                    const uint32_t t3 = c3 & t2; c3 ^= t2;
                    c4 ^= t3;
                }
                // column-block closed: plane scan from the top = word maximum + the variables that reach it
                uint32_t mask = 0xffffffffu, m;
                int u = 0;
                m = mask & c4; if (m) { mask = m; u |= 16; }
                m = mask & c3; if (m) { mask = m; u |= 8; }
                m = mask & c2; if (m) { mask = m; u |= 4; }
                m = mask & c1; if (m) { mask = m; u |= 2; }
                m = mask & c0; if (m) { mask = m; u |= 1; }
                const int en = 2 * u - deg;
                const int j = a.colid[part][k++];
                cb[oMK + j * W + w] = mask;
                EN[j * W + w] = (int8_t)en;
                mymax = max(mymax, en);
                c0 = c1 = c2 = c3 = c4 = 0;
                deg = 0;
            }
            // degree-1 extension column-blocks ncore + k, k = part, part + 8, ...: En = +1 where the row-block's check fails
            for (int k = part; k < next; k += kBwParts) {
                const uint32_t b = SWw[(4 + k) * W2];
                const int en = b ? 1 : -1;
                cb[oMK + (ncore + k) * W + w] = b ? b : 0xffffffffu;
                EN[(ncore + k) * W + w] = (int8_t)en;
                mymax = max(mymax, en);
            }
        }
        {   // one shared atomicMax per (warp, codeblock)
            const uint32_t peers = __match_any_sync(0xffffffffu, act ? g : -1);
            const int wmax = __reduce_max_sync(peers, mymax);
            if (act && lane == (__ffs(peers) - 1)) atomicMax(&s_max[g], wmax);
        }
        if (!__syncthreads_or(act)) break;
        if (tid < kBwPart) s_any[tid] = 0;
        // :67-70 flip every bit whose metric equals the codeblock's maximum (each thread: the words it evaluated)
        if (act) {
            const int mx = s_max[g];
            for (int k = 0; k < a.ncid[part]; ++k) {
                const int j = a.colid[part][k];
                if (EN[j * W + w] == mx) {
                    const uint32_t v = cb[j * W2 + w] ^ cb[oMK + j * W + w];
                    cb[j * W2 + w] = v;
                    cb[j * W2 + W + w] = v;
                }
            }
            for (int k = part; k < next; k += kBwParts)
                if (EN[(ncore + k) * W + w] == mx) cb[oCKX + k * W + w] ^= cb[oMK + (ncore + k) * W + w];
        }
        __syncthreads();
        if (tid < kBwPart) s_max[tid] = -128;
    }

    __syncthreads();
    // ck: 16 decisions -> one 128-bit store
    const int H = Zc >> 4;
    for (int t = tid; t < g_cnt * ncols * H; t += kBwThreads) {
        const int gg = t / (ncols * H), rem = t - gg * ncols * H, j = rem / H, h = rem - j * H;
        const uint32_t *c = smem + gg * a.slot;
        const uint32_t bits = reinterpret_cast<const uint16_t *>(j < ncore ? c + j * W2 : c + oCKX + (j - ncore) * W)[h];
        uint32_t o[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) o[q] = (((bits >> (4 * q)) & 0xfu) * 0x00204081u) & 0x01010101u;
        *reinterpret_cast<uint4 *>(ck_out + (size_t)(cb0 + gg) * a.Nfull + j * Zc + 16 * h) = make_uint4(o[0], o[1], o[2], o[3]);
    }
    if (part == 0 && w == 0 && g < g_cnt) {
        if (status) status[cb0 + g] = (uint8_t)ok;
        if (iters) iters[cb0 + g] = it_done;
    }
}

int launch_bf_words(const QcCfg &c, const void *d_llr, int is_f64, int B, int max_iter, int8_t *d_ck, uint8_t *d_status,
                    int32_t *d_iters, cudaStream_t s, bool *handled)
{
    *handled = false;
    const int W = c.tiles;
    BfWordArgs a = {};
    a.W = W; a.Zc = c.Zc; a.ncols = c.ncols; a.nrows = c.nrows; a.ncore = c.ncore; a.N = c.N; a.Nfull = c.Nfull;
    a.G = kBwPart / W;
    const int next = c.nrows - 4;
    int slot = c.ncore * 2 * W + next * W + c.nrows * 2 * W + c.ncols * W + (c.ncols * W + 3) / 4;
    slot += ((W - slot) % 32 + 32) % 32;  // slot = W (mod 32): thread (g, w) falls in bank (g W + w) mod 32
    a.slot = slot;
    a.mNW = (uint32_t)((1ull << 32) / (uint32_t)((c.ncols - 2) * W)) + 1u;
    a.mW = (uint32_t)((1ull << 32) / (uint32_t)W) + 1u;
    {   // row-blocks and core column-blocks go to the least loaded part, heaviest first (round robin would give the
        // part that owns BG1 column-blocks 0 / row-blocks 0-3 1.6x / 1.3x the average number of edges)
        int load[kBwParts] = {}, order[kMaxRows];
        int deg[kMaxRows];
        for (int i = 0; i < c.nrows; ++i) { deg[i] = c.rowptr[i + 1] - c.rowptr[i] - (i >= 4 ? 1 : 0); order[i] = i; }
        std::stable_sort(order, order + c.nrows, [&](int x, int y) { return deg[x] > deg[y]; });
        for (int t = 0; t < c.nrows; ++t) {
            const int i = order[t], p = (int)(std::min_element(load, load + kBwParts) - load);
            if (a.nrid[p] >= kBwIds || a.nrow[p] + deg[i] > kBwList) return NRLDPC_OK;  // not handled: bf_qc_kernel
            load[p] += deg[i] + 2;
            a.rowid[p][a.nrid[p]++] = (uint8_t)i;
            const int e1 = c.rowptr[i + 1] - (i >= 4 ? 1 : 0);  // rows >= 4: the last edge is the degree-1 column
            for (int e = c.rowptr[i]; e < e1; ++e) {
                const int j = c.edge[e] & 0xff, P = c.edge[e] >> 8;
                if (j >= c.ncore) return NRLDPC_OK;
                a.rowl[p][a.nrow[p]++] = (uint32_t)((j * 2 * W + (P >> 5)) * 4) << 16 | (uint32_t)(e == e1 - 1) << 5 | (uint32_t)(P & 31);
            }
        }
        std::fill(load, load + kBwParts, 0);
        for (int j = 0; j < c.ncore; ++j) { deg[j] = c.colptr[j + 1] - c.colptr[j]; order[j] = j; }
        std::stable_sort(order, order + c.ncore, [&](int x, int y) { return deg[x] > deg[y]; });
        for (int t = 0; t < c.ncore; ++t) {
            const int j = order[t], p = (int)(std::min_element(load, load + kBwParts) - load);
            if (a.ncid[p] >= kBwIds || a.ncol[p] + deg[j] > kBwList) return NRLDPC_OK;
            load[p] += deg[j] + 3;
            a.colid[p][a.ncid[p]++] = (uint8_t)j;
            const int q1 = c.colptr[j + 1];
            for (int q = c.colptr[j]; q < q1; ++q) {
                const int i = c.centry[q] & 0x3f, back = c.centry[q] >> 16;
                a.coll[p][a.ncol[p]++] = (uint32_t)((i * 2 * W + (back >> 5)) * 4) << 16 | (uint32_t)(q == q1 - 1) << 5 | (uint32_t)(back & 31);
            }
        }
    }
    const int smem_bytes = a.G * slot * 4;
    if (smem_bytes > 200 * 1024) return NRLDPC_OK;
    *handled = true;
    if (is_f64) {
        NRLDPC_CUDA(cudaFuncSetAttribute(bf_words_kernel<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
        NRLDPC_CUDA(cudaFuncSetAttribute(bf_words_kernel<double>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        bf_words_kernel<double><<<(B + a.G - 1) / a.G, kBwThreads, smem_bytes, s>>>(a, (const double *)d_llr, B, max_iter, d_ck, d_status, d_iters);
    } else {
        NRLDPC_CUDA(cudaFuncSetAttribute(bf_words_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
        NRLDPC_CUDA(cudaFuncSetAttribute(bf_words_kernel<float>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        bf_words_kernel<float><<<(B + a.G - 1) / a.G, kBwThreads, smem_bytes, s>>>(a, (const float *)d_llr, B, max_iter, d_ck, d_status, d_iters);
    }
    NRLDPC_CUDA(cudaGetLastError());
    return NRLDPC_OK;
}

}  // namespace

int launch_bf_qc(const QcCfg &cfg, const void *d_llr, int is_f64, int B, int max_iter, int8_t *d_ck,
                 uint8_t *d_status, int32_t *d_iters, cudaStream_t s)
{
    if (B <= 0) return NRLDPC_OK;
    static const bool no_words = getenv("NRLDPC_BF_NO_WORDS") != nullptr;  // A/B timing against bf_qc_kernel
    if (cfg.Zc % 32 == 0 && reinterpret_cast<uintptr_t>(d_ck) % 16 == 0 && !no_words) {
        bool handled = false;
        const int rc = launch_bf_words(cfg, d_llr, is_f64, B, max_iter, d_ck, d_status, d_iters, s, &handled);
        if (rc != NRLDPC_OK || handled) return rc;
    }
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
