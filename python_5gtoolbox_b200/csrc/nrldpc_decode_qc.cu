// nrldpc_decode_qc.cu -- batched flooding min-sum decoder for 5G NR LDPC on sm_100a (fp32).
//
// Replaces nr_ldpc_decode.nr_decode_ldpc / decode_ldpc / _min_sum_process
// (py5gphy/ldpc/nr_ldpc_decode.py:11-49, :51-143, :178-227) for algo = 'min-sum' with any alpha/beta.
//
// Schedule = the reference's FLOODING schedule, reproduced operation for operation in fp32:
//   per iteration  (1) hard decision LQ<0 and syndrome of every check (early exit, :107-114)
//                  (2) every check row from the OLD variable-to-check messages Lq (:117-123)
//                  (3) LQ = LLRin + sum of the new check-to-variable messages Lr in ascending check
//                      index (:126), then Lq = LQ - Lr per edge (:129-131)
//   after L iterations the final decision uses the other tie rule LQ<=0 -> 1 (:134-143).
//
// Data layout (all of a codeblock's state is resident in shared memory; nothing per-edge is stored):
//   LQ[ncore][Zc] fp32        posteriors of the kb+4 column-blocks of degree > 1
//   rec[nrows][Zc]            one compressed record per check row: float2 {mag1, mag2} = the two
//                             possible |Lr| = alpha*max(min{1,2}-beta,0), plus a sign/index word
//                             (bit deg-1-k = sign of Lr on edge k, upper bits = argmin edge)
//   Lq of an edge is recomputed as LQ - Lr(record); the degree-1 extension columns never store a
//   posterior: LQ_ext = fl(LLR + Lr), Lq_ext = fl(LQ_ext - Lr), with the LLR re-read from L2.
// The three zero-input cases of _min_sum_process (:188-225) fall out of the record with sign(0) := +.
//
// Mapping: a CTA owns G codeblocks; a warp processes "tiles" = (row-block or column-block, 32
// consecutive lifted indices r); the circulant shift is the shared-memory index rotation (r+P) mod Zc,
// conflict-free because a warp reads 32 consecutive words.  For Zc < 32 several codeblocks share a warp.
#include "nrldpc_common.cuh"

namespace nrldpc {

namespace {

struct DecArgs {
    const float *llr;
    int B, G, max_iter, early_term;
    float alpha, beta;
    int8_t *ck;
    uint32_t *info;
    uint8_t *status;
    int32_t *iters;
    int slot_bytes, off_mags, off_bits, off_ext;
};

constexpr int kMaxG = 32;
constexpr uint32_t kInfBits = 0x7f800000u;

// d = min(|a|,|b|) with sign(a) xor sign(b): running "sign product * first minimum" of a check row.
__device__ __forceinline__ float min_xorsign_abs(float a, float b)
{
    float d;
    asm("min.xorsign.abs.f32 %0, %1, %2;" : "=f"(d) : "f"(a), "f"(b));
    return d;
}

// Check-to-variable message of edge k decoded from a row record.
// WIDE: idx at bit 24, else at bit 12; sign of edge k at bit `bitpos`.
__device__ __forceinline__ float record_lr(float2 m, uint32_t bits, int idx, int k, int bitpos)
{
    const float mag = (idx == k) ? m.y : m.x;
    return __uint_as_float(__float_as_uint(mag) ^ (((bits >> bitpos) & 1u) << 31));
}

struct Lane {
    int g;        // codeblock slot inside the CTA
    int r;        // lifted index (clamped to 0 when inactive)
    int cb;       // global codeblock (clamped to B-1 when inactive)
    bool active;  // this lane owns a real (codeblock, r) that is still being decoded
};

// Decompose tile-column `col` (0 .. G*tiles/per) into the lane's codeblock slot and lifted index.
__device__ __forceinline__ Lane lane_of(const QcCfg &c, const DecArgs &a, int col, int lane, uint32_t donemask)
{
    Lane L;
    const int g = (col / c.tiles) * c.per + (lane >> c.lanes_log2);
    const int r = (col % c.tiles) * 32 + (lane & (c.lanes - 1));
    const int cb = blockIdx.x * a.G + g;
    L.active = (r < c.Zc) && (g < a.G) && (cb < a.B) && !((donemask >> (g & 31)) & 1u);
    L.g = (g < a.G) ? g : 0;
    L.r = (r < c.Zc) ? r : 0;
    L.cb = (cb < a.B) ? cb : a.B - 1;
    return L;
}

// One check row (row-block i, lifted index r): syndrome bit of the current hard decisions, then the
// min-sum update of its record from Lq = LQ - Lr_old.  py5gphy/ldpc/nr_ldpc_decode.py:107-123,178-227.
template <int DEG, bool EXT, bool WIDE>
__device__ __forceinline__ void cn_row(const QcCfg &c, const DecArgs &a, char *slot, int i, const Lane &L,
                                       int lane, int *flag)
{
    const int Zc = c.Zc, r = L.r;
    const int e0 = c.rowptr[i];
    float2 *mp = reinterpret_cast<float2 *>(slot + a.off_mags) + i * Zc + r;
    char *bp = slot + a.off_bits + c.bits_off[i] * Zc;
    const float2 m = *mp;
    const uint32_t bits = WIDE ? reinterpret_cast<uint32_t *>(bp)[r] : reinterpret_cast<uint16_t *>(bp)[r];
    const int idx = bits >> (WIDE ? 24 : 12);
    float llr_e = 0.f;
    if (EXT) llr_e = __fadd_rn(__ldg(a.llr + (size_t)L.cb * c.N + (size_t)(c.kb + i - 2) * Zc + r), 0.0f);
    const float *LQ = reinterpret_cast<const float *>(slot);

    float vmin = __uint_as_float(kInfBits);  // sign = running sign product, |vmin| = first minimum
    float min2 = vmin;
    int nidx = 0;
    uint32_t sacc = 0, synd = 0;
#pragma unroll
    for (int k = 0; k < DEG; ++k) {
        const float lr = record_lr(m, bits, idx, k, DEG - 1 - k);
        float x;
        if (EXT && k == DEG - 1) {
            x = __fadd_rn(llr_e, lr);  // posterior of the degree-1 extension variable (:126)
            const uint32_t hb = __ballot_sync(0xffffffffu, x < 0.f);
            if (L.active && (r & (c.lanes - 1)) == 0)
                reinterpret_cast<uint32_t *>(slot + a.off_ext)[(i - 4) * c.tiles + (r >> 5)] = hb;
        } else {
            const uint32_t ew = c.edge[e0 + k];
            int cc = r + (int)(ew >> 8);
            if (cc >= Zc) cc -= Zc;
            x = LQ[(ew & 0xff) * Zc + cc];
        }
        synd ^= __float_as_uint(x);           // sign bit = hard decision LQ<0 (:107-108)
        const float q = __fsub_rn(x, lr);     // Lq = LQ - Lr (:131)
        const float aq = fabsf(q), a1 = fabsf(vmin);
        min2 = fminf(min2, fmaxf(a1, aq));
        nidx = (aq < a1) ? k : nidx;
        vmin = min_xorsign_abs(vmin, q);
        sacc = __funnelshift_l(__float_as_uint(q), sacc, 1);  // sign of Lq on edge k -> bit DEG-1-k
    }
    // :199-202  Lr = alpha * sign_prod * sign(Lq) * max(minv - beta, 0), minv = min2 on the argmin edge
    const float mag1 = __fmul_rn(a.alpha, fmaxf(__fsub_rn(fabsf(vmin), a.beta), 0.f));
    const float mag2 = __fmul_rn(a.alpha, fmaxf(__fsub_rn(min2, a.beta), 0.f));
    const uint32_t sp = (uint32_t)((int)__float_as_uint(vmin) >> 31);
    const uint32_t nb = ((sacc ^ sp) & ((1u << DEG) - 1u)) | ((uint32_t)nidx << (WIDE ? 24 : 12));
    if (L.active) {
        *mp = make_float2(mag1, mag2);
        if (WIDE) reinterpret_cast<uint32_t *>(bp)[r] = nb;
        else reinterpret_cast<uint16_t *>(bp)[r] = (uint16_t)nb;
        if (synd >> 31) flag[L.g] = 1;
    }
}

__device__ __forceinline__ void cn_dispatch(const QcCfg &c, const DecArgs &a, char *slot, int i, const Lane &L,
                                            int lane, int *flag)
{
    const int deg = c.rowptr[i + 1] - c.rowptr[i];
    if (i < 4) {
        switch (deg) {
        case 19: cn_row<19, false, true>(c, a, slot, i, L, lane, flag); break;
        case 10: cn_row<10, false, false>(c, a, slot, i, L, lane, flag); break;
        default: cn_row<8, false, false>(c, a, slot, i, L, lane, flag); break;
        }
    } else {
        switch (deg) {
        case 3: cn_row<3, true, false>(c, a, slot, i, L, lane, flag); break;
        case 4: cn_row<4, true, false>(c, a, slot, i, L, lane, flag); break;
        case 5: cn_row<5, true, false>(c, a, slot, i, L, lane, flag); break;
        case 6: cn_row<6, true, false>(c, a, slot, i, L, lane, flag); break;
        case 7: cn_row<7, true, false>(c, a, slot, i, L, lane, flag); break;
        case 8: cn_row<8, true, false>(c, a, slot, i, L, lane, flag); break;
        case 9: cn_row<9, true, false>(c, a, slot, i, L, lane, flag); break;
        default: cn_row<10, true, false>(c, a, slot, i, L, lane, flag); break;
        }
    }
}

// One core variable (column-block j, lifted index cc): LQ = LLRin + sum_i Lr(i) in ascending
// row-block order = ascending check index (py5gphy/ldpc/nr_ldpc_decode.py:126).
__device__ __forceinline__ void vn_col(const QcCfg &c, const DecArgs &a, char *slot, int j, const Lane &L)
{
    const int Zc = c.Zc, cc = L.r;
    float lv = 0.f;  // the 2Zc punctured systematic bits start at LLR 0 (:43)
    if (j >= 2) lv = __fadd_rn(__ldg(a.llr + (size_t)L.cb * c.N + (size_t)(j - 2) * Zc + cc), 0.0f);
    const float2 *mags = reinterpret_cast<const float2 *>(slot + a.off_mags);
    const char *bbase = slot + a.off_bits;
    float acc = 0.f;
    const int q1 = c.colptr[j + 1];
    for (int q = c.colptr[j]; q < q1; ++q) {
        const uint32_t en = c.centry[q];
        const int i = en & 63, k = (en >> 6) & 31, bitpos = (en >> 11) & 31;
        int r = cc + (int)(en >> 16);
        if (r >= Zc) r -= Zc;
        const float2 m = mags[i * Zc + r];
        const char *bp = bbase + c.bits_off[i] * Zc;
        uint32_t bits;
        int idx;
        if (c.wide[i]) { bits = reinterpret_cast<const uint32_t *>(bp)[r]; idx = bits >> 24; }
        else { bits = reinterpret_cast<const uint16_t *>(bp)[r]; idx = bits >> 12; }
        acc = __fadd_rn(acc, record_lr(m, bits, idx, k, bitpos));
    }
    if (L.active) reinterpret_cast<float *>(slot)[j * Zc + cc] = __fadd_rn(lv, acc);
}

// Final syndrome with the post-loop tie rule LQ<=0 -> 1 (py5gphy/ldpc/nr_ldpc_decode.py:134-143).
__device__ __forceinline__ void final_row(const QcCfg &c, const DecArgs &a, char *slot, int i, const Lane &L,
                                          int *flag)
{
    const int Zc = c.Zc, r = L.r;
    const int e0 = c.rowptr[i], deg = c.rowptr[i + 1] - e0;
    const float *LQ = reinterpret_cast<const float *>(slot);
    const int ncoredeg = (i >= 4) ? deg - 1 : deg;
    uint32_t synd = 0;
    for (int k = 0; k < ncoredeg; ++k) {
        const uint32_t ew = c.edge[e0 + k];
        int cc = r + (int)(ew >> 8);
        if (cc >= Zc) cc -= Zc;
        synd ^= (LQ[(ew & 0xff) * Zc + cc] <= 0.f) ? 1u : 0u;
    }
    if (i >= 4) {
        const float2 m = reinterpret_cast<const float2 *>(slot + a.off_mags)[i * Zc + r];
        const uint32_t bits = reinterpret_cast<const uint16_t *>(slot + a.off_bits + c.bits_off[i] * Zc)[r];
        const float lr = record_lr(m, bits, bits >> 12, deg - 1, 0);
        const float llr_e = __fadd_rn(__ldg(a.llr + (size_t)L.cb * c.N + (size_t)(c.kb + i - 2) * Zc + r), 0.0f);
        const bool hb1 = __fadd_rn(llr_e, lr) <= 0.f;
        const uint32_t hb = __ballot_sync(0xffffffffu, hb1);
        if (L.active && (r & (c.lanes - 1)) == 0)
            reinterpret_cast<uint32_t *>(slot + a.off_ext)[(i - 4) * c.tiles + (r >> 5)] = hb;
        synd ^= hb1 ? 1u : 0u;
    }
    if (L.active && synd) flag[L.g] = 1;
}

template <int NT>
__global__ void __launch_bounds__(NT, 1)
decode_minsum_kernel(const __grid_constant__ QcCfg c, const __grid_constant__ DecArgs a)
{
    extern __shared__ __align__(16) char smem[];
    __shared__ int s_flag[2][kMaxG];
    const int Zc = c.Zc, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int NW = NT / 32;
    const int G = a.G, cb0 = blockIdx.x * G;
    const int ncolumns = (G / c.per) * c.tiles;  // warp-tile columns across the CTA's codeblocks

    // ---- init: records = 0 (Lr = 0, :101), LQ = LLRin (:94) with the punctured columns at 0 (:43)
    {
        uint32_t *w = reinterpret_cast<uint32_t *>(smem);
        const int nw = G * a.slot_bytes / 4;
        for (int t = tid; t < nw; t += NT) w[t] = 0;
        if (tid < kMaxG) { s_flag[0][tid] = 0; s_flag[1][tid] = 0; }
    }
    __syncthreads();
    uint32_t donemask = 0;
    for (int g = 0; g < G; ++g)
        if (cb0 + g >= a.B) donemask |= 1u << g;
    const uint32_t fullmask = (G >= 32) ? 0xffffffffu : ((1u << G) - 1u);
    for (int t = warp; t < (c.ncore - 2) * ncolumns; t += NW) {
        const int j = 2 + t / ncolumns;
        const Lane L = lane_of(c, a, t % ncolumns, lane, donemask);
        if (L.active)
            reinterpret_cast<float *>(smem + L.g * a.slot_bytes)[j * Zc + L.r] =
                __fadd_rn(__ldg(a.llr + (size_t)L.cb * c.N + (size_t)(j - 2) * Zc + L.r), 0.0f);
    }
    __syncthreads();

    uint32_t et_mask = 0;  // codeblocks that left through the in-loop syndrome check (tie rule LQ<0)
    int my_iters = a.max_iter;
    int it = 0;
    for (; it < a.max_iter; ++it) {
        int *flag = s_flag[it & 1];
        // ---- check-node pass (+ syndrome of the current hard decisions)
        for (int t = warp; t < c.nrows * ncolumns; t += NW) {
            const int i = c.cn_order[t / ncolumns];
            const Lane L = lane_of(c, a, t % ncolumns, lane, donemask);
            cn_dispatch(c, a, smem + L.g * a.slot_bytes, i, L, lane, flag);
        }
        __syncthreads();
        if (a.early_term) {
            for (int g = 0; g < G; ++g)
                if (!((donemask >> g) & 1u) && flag[g] == 0) {
                    donemask |= 1u << g;
                    et_mask |= 1u << g;
                    if (tid == g) my_iters = it;
                }
        }
        if (tid < kMaxG) s_flag[(it + 1) & 1][tid] = 0;
        if (donemask == fullmask) break;
        // ---- variable-node pass
        for (int t = warp; t < c.ncore * ncolumns; t += NW) {
            const int j = c.vn_order[t / ncolumns];
            const Lane L = lane_of(c, a, t % ncolumns, lane, donemask);
            vn_col(c, a, smem + L.g * a.slot_bytes, j, L);
        }
        __syncthreads();
    }

    // ---- final decision + syndrome for the codeblocks that did not leave early
    uint32_t okmask = et_mask;
    if (donemask != fullmask) {
        int *flag = s_flag[it & 1];  // cleared above, not yet written
        for (int t = warp; t < c.nrows * ncolumns; t += NW) {
            const int i = t / ncolumns;
            const Lane L = lane_of(c, a, t % ncolumns, lane, donemask);
            final_row(c, a, smem + L.g * a.slot_bytes, i, L, flag);
        }
        __syncthreads();
        for (int g = 0; g < G; ++g)
            if (!((donemask >> g) & 1u) && flag[g] == 0) okmask |= 1u << g;
    }

    // ---- outputs
    if (tid < G && cb0 + tid < a.B) {
        if (a.status) a.status[cb0 + tid] = (okmask >> tid) & 1u;
        if (a.iters) a.iters[cb0 + tid] = my_iters;
    }
    for (int g = 0; g < G; ++g) {
        const int cb = cb0 + g;
        if (cb >= a.B) break;
        const char *slot = smem + g * a.slot_bytes;
        const float *LQ = reinterpret_cast<const float *>(slot);
        const uint32_t *ext = reinterpret_cast<const uint32_t *>(slot + a.off_ext);
        const bool et = (et_mask >> g) & 1u;
        const int sub = (g % c.per) << c.lanes_log2;
        if (a.ck) {
            int8_t *out = a.ck + (size_t)cb * c.Nfull;
            const int ncoreN = c.ncore * Zc;
            for (int n = tid; n < c.Nfull; n += NT) {
                int bit;
                if (n < ncoreN) {
                    const float x = LQ[n];
                    bit = et ? (x < 0.f) : (x <= 0.f);
                } else {
                    const int m = n - ncoreN, i4 = m / Zc, r = m - i4 * Zc;
                    bit = (ext[i4 * c.tiles + (r >> 5)] >> (sub + (r & 31))) & 1u;
                }
                out[n] = (int8_t)bit;
            }
        }
        if (a.info) {
            const int nwords = (c.K + 31) / 32;
            for (int w = warp; w < nwords; w += NW) {
                const int n = 32 * w + lane;
                bool bit = false;
                if (n < c.K) { const float x = LQ[n]; bit = et ? (x < 0.f) : (x <= 0.f); }
                const uint32_t word = __ballot_sync(0xffffffffu, bit);
                if (lane == 0) a.info[(size_t)cb * nwords + w] = word;
            }
        }
    }
}

void fill_layout(const QcCfg &cfg, DecArgs *a)
{
    const int Zc = cfg.Zc;
    a->off_mags = cfg.ncore * Zc * 4;
    a->off_mags = (a->off_mags + 7) & ~7;
    a->off_bits = a->off_mags + cfg.nrows * Zc * 8;
    a->off_ext = (a->off_bits + cfg.bits_bytes_per_zc * Zc + 3) & ~3;
    a->slot_bytes = (a->off_ext + (cfg.nrows - 4) * cfg.tiles * 4 + 15) & ~15;
}

constexpr int kSmemMax = 227 * 1024;

}  // namespace

int decode_minsum_geometry(const QcCfg &cfg, int *G_out, int *threads, int *smem)
{
    DecArgs a;
    fill_layout(cfg, &a);
    int gmax = kSmemMax / a.slot_bytes;
    if (gmax < 1) return NRLDPC_EINVAL;
    if (gmax > kMaxG) gmax = kMaxG;
    // aim for ~12 warp-tile columns per CTA (what one Zc=384 codeblock provides)
    int want = ((12 + cfg.tiles - 1) / cfg.tiles) * cfg.per;
    int G = want < gmax ? want : gmax;
    if (cfg.per > 1) G = (G / cfg.per) * cfg.per;
    if (G < 1) G = 1;
    const int columns = (G / cfg.per) * cfg.tiles;
    int nt = columns >= 8 ? 1024 : (columns >= 4 ? 512 : 256);
    if (G_out) *G_out = G;
    if (threads) *threads = nt;
    if (smem) *smem = G * a.slot_bytes;
    return NRLDPC_OK;
}

int launch_decode_minsum(const QcCfg &cfg, const float *d_llr, int B, int max_iter, float alpha, float beta,
                         int early_term, int8_t *d_ck, uint32_t *d_info, uint8_t *d_status, int32_t *d_iters,
                         cudaStream_t s)
{
    if (B <= 0) return NRLDPC_OK;
    DecArgs a;
    fill_layout(cfg, &a);
    int G, nt, smem;
    if (decode_minsum_geometry(cfg, &G, &nt, &smem)) { set_error("decode_minsum: codeblock does not fit in shared memory"); return NRLDPC_EINVAL; }
    a.llr = d_llr; a.B = B; a.G = G; a.max_iter = max_iter; a.early_term = early_term;
    a.alpha = alpha; a.beta = beta; a.ck = d_ck; a.info = d_info; a.status = d_status; a.iters = d_iters;
    const int grid = (B + G - 1) / G;
    auto launch = [&](auto kern) -> int {
        NRLDPC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemMax));
        kern<<<grid, nt, smem, s>>>(cfg, a);
        NRLDPC_CUDA(cudaGetLastError());
        return NRLDPC_OK;
    };
    if (nt == 1024) return launch(decode_minsum_kernel<1024>);
    if (nt == 512) return launch(decode_minsum_kernel<512>);
    return launch(decode_minsum_kernel<256>);
}

}  // namespace nrldpc
