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
// Mapping: a CTA owns G codeblocks.  Every thread owns ONE (codeblock, lifted index r) for the whole
// kernel; its warp walks a fixed subset of the row-blocks (check pass) and core column-blocks
// (variable pass), so per-tile index arithmetic disappears and the per-edge tables (byte offsets,
// built on the host, passed by value in the constant bank) are warp-uniform.  The circulant shift is
// the shared-memory index rotation (r+P) mod Zc: conflict-free, 32 consecutive words per warp.
// For Zc < 32 several codeblocks share a warp.
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <vector>

#include <utility>

#include "nrldpc_decode.cuh"

namespace nrldpc {

namespace {

constexpr int kMaxS = 16;   // warp groups per warp column
constexpr int kNumCls = 11; // check-row classes: <19,noext,wide> <10,noext> <8,noext> <10..3,ext>

struct RowInfo {
    uint32_t mrel;     // i * Zc * 8: byte offset of the row-block's float2 magnitudes inside the mags region;
                       // its sign words sit at (mrel + r*8)/2 (32-bit) or /4 (16-bit) inside their regions
    uint32_t ext_llr;  // byte offset of the extension column's LLRs inside a codeblock's LLR row
    uint32_t e0;       // first edge in the edge table
    uint32_t pad;
};
struct ColInfo {
    uint32_t lq_off;   // byte offset of the column-block's posteriors inside a slot
    uint32_t llr_off;  // byte offset of its channel LLRs inside a codeblock's LLR row, or kNoLlr (punctured)
    uint16_t q0;       // first entry in the variable-pass table
    uint8_t nwide, deg;
};
constexpr uint32_t kNoLlr = 0xffffffffu;

// Everything the kernel needs, by value (constant bank).  The two per-edge tables are staged into
// shared memory at kernel start (one LDS.64 / LDS.128 per edge instead of constant-bank reads).
struct DecTab {
    int Zc, Z4, Z8, nrows, ncore, kb, K, N, Nfull, tiles, lanes, lanes_log2, per;
    int G, ncolumns, S, nwarps;
    int slot_bytes, off_mags, off_bw, off_bn, off_ext;  // off_bn already rebased by -(#wide rows)*Zc*2
    int tab_cn, tab_vn, smem_bytes;                     // byte offsets of the staged tables, total dynamic smem
    int nedges, nventries;
    uint2 cn_edge[kMaxEdges];       // {byte offset of the column's LQ array, shift * 4}
    uint4 vn_entry[kMaxCoreEdges];  // {row mrel, ((Zc - shift) % Zc) * 8, k << idxshift | (31 - bitpos), 0}
    RowInfo row[kMaxRows];
    ColInfo col[kMaxCore];
    uint16_t cls_begin[kNumCls * kMaxS], cls_end[kNumCls * kMaxS];  // ranges of cn_list per (class, warp group)
    uint8_t cn_list[kMaxRows];
    uint16_t vn_begin[kMaxS], vn_end[kMaxS];                        // ranges of vn_list per warp group
    uint8_t vn_list[kMaxCore];
};

constexpr int kMaxG = 32;
struct Me {               // what a thread owns for the whole kernel
    char *slot;           // its codeblock's shared-memory slot
    char *mags, *bw, *bn; // slot + off_mags / off_bw / off_bn
    const char *llr;      // CTA-uniform base of the CTA's first codeblock's LLR row
    uint32_t llr_off;     // this thread's byte offset from it: (g*N + r) * 4
    const uint2 *tcn;     // staged edge tables
    const uint4 *tvn;
    int g, r;             // slot index, lifted index (clamped to a valid value when !valid)
    bool valid;
};

// RR: the rows were written by this CTA (fused rate recovery): L2-coherent load instead of the read-only path
template <bool RR> __device__ __forceinline__ float load_llr(const Me &me, uint32_t off)
{
    const float *p = reinterpret_cast<const float *>(me.llr + (me.llr_off + off));
    return __fadd_rn(RR ? __ldcg(p) : __ldg(p), 0.0f);  // -0.0 -> +0.0
}

// One check row (row-block i, lifted index r): syndrome bit of the current hard decisions, then the
// min-sum update of its record from Lq = LQ - Lr_old.  py5gphy/ldpc/nr_ldpc_decode.py:107-123,178-227.
template <int DEG, bool EXT, bool WIDE, bool ET, bool RR>
__device__ __forceinline__ void cn_row(const DecTab &T, const DecArgs &a, const Me &me, const int i, const bool active,
                                       int *flag)
{
    constexpr int SH = WIDE ? 24 : 12;
    constexpr uint32_t IDXMASK = WIDE ? 0x1f000000u : 0xf000u;
    const RowInfo ri = T.row[i];
    const uint32_t r4 = (uint32_t)me.r * 4u;
    const uint32_t rel = ri.mrel + 2 * r4;
    char *rec = me.mags + rel;
    char *bp = WIDE ? me.bw + (rel >> 1) : me.bn + (rel >> 2);
    const float2 m = *reinterpret_cast<const float2 *>(rec);
    const uint32_t bits = WIDE ? *reinterpret_cast<const uint32_t *>(bp) : *reinterpret_cast<const uint16_t *>(bp);
    float llr_e = 0.f;
    if (EXT) llr_e = load_llr<RR>(me, ri.ext_llr);
    const uint2 *et = me.tcn + ri.e0;

    float vmin = __uint_as_float(kInfBits);  // sign = running sign product, |vmin| = first minimum
    float min2 = vmin;
    uint32_t nidx = 0, sacc = 0, synd = 0;
#pragma unroll
    for (int k = 0; k < DEG; ++k) {
        const bool isidx = ((bits ^ ((uint32_t)k << SH)) & IDXMASK) == 0;
        const float lr = record_lr(m, isidx, bits << (31 - (DEG - 1 - k)));
        float x;
        if (EXT && k == DEG - 1) {
            x = __fadd_rn(llr_e, lr);  // posterior of the degree-1 extension variable (:126)
            if (ET) {
                const uint32_t hb = __ballot_sync(0xffffffffu, x < 0.f);
                if (active && (me.r & (T.lanes - 1)) == 0)
                    reinterpret_cast<uint32_t *>(me.slot + T.off_ext)[(i - 4) * T.tiles + (me.r >> 5)] = hb;
            }
        } else {
            const uint2 ew = et[k];
            const uint32_t t = r4 + ew.y;
            x = *reinterpret_cast<const float *>(me.slot + ew.x + min(t, t - (uint32_t)T.Z4));
        }
        if (ET) synd ^= __float_as_uint(x);   // sign bit = hard decision LQ<0 (:107-108)
        const float q = __fsub_rn(x, lr);     // Lq = LQ - Lr (:131)
        const float aq = fabsf(q), a1 = fabsf(vmin);
        min2 = fminf(min2, fmaxf(a1, aq));
        nidx = (aq < a1) ? (uint32_t)k << SH : nidx;
        vmin = min_xorsign_abs(vmin, q);
        sacc = __funnelshift_l(__float_as_uint(q), sacc, 1);  // sign of Lq on edge k -> bit DEG-1-k
    }
    // :199-202  Lr = alpha * sign_prod * sign(Lq) * max(minv - beta, 0), minv = min2 on the argmin edge
    const float mag1 = __fmul_rn(a.alpha, fmaxf(__fsub_rn(fabsf(vmin), a.beta), 0.f));
    const float mag2 = __fmul_rn(a.alpha, fmaxf(__fsub_rn(min2, a.beta), 0.f));
    const uint32_t sp = (uint32_t)((int)__float_as_uint(vmin) >> 31);
    const uint32_t nb = ((sacc ^ sp) & ((1u << DEG) - 1u)) | nidx;
    if (active) {
        *reinterpret_cast<float2 *>(rec) = make_float2(mag1, mag2);
        if (WIDE) *reinterpret_cast<uint32_t *>(bp) = nb;
        else *reinterpret_cast<uint16_t *>(bp) = (uint16_t)nb;
        if (ET && (synd >> 31)) flag[me.g] = 1;
    }
}

// All rows of one class that belong to this warp's group.
template <int CLS, int DEG, bool EXT, bool WIDE, bool ET, bool RR>
__device__ __forceinline__ void cn_class(const DecTab &T, const DecArgs &a, const Me &me, int sub, bool active, int *flag)
{
    const int e = T.cls_end[CLS * kMaxS + sub];
    for (int o = T.cls_begin[CLS * kMaxS + sub]; o < e; ++o) cn_row<DEG, EXT, WIDE, ET, RR>(T, a, me, T.cn_list[o], active, flag);
}

template <bool ET, bool RR>
__device__ __forceinline__ void cn_pass(const DecTab &T, const DecArgs &a, const Me &me, int sub, bool active, int *flag)
{
    cn_class<0, 19, false, true, ET, RR>(T, a, me, sub, active, flag);
    cn_class<1, 10, false, false, ET, RR>(T, a, me, sub, active, flag);
    cn_class<2, 8, false, false, ET, RR>(T, a, me, sub, active, flag);
    cn_class<3, 10, true, false, ET, RR>(T, a, me, sub, active, flag);
    cn_class<4, 9, true, false, ET, RR>(T, a, me, sub, active, flag);
    cn_class<5, 8, true, false, ET, RR>(T, a, me, sub, active, flag);
    cn_class<6, 7, true, false, ET, RR>(T, a, me, sub, active, flag);
    cn_class<7, 6, true, false, ET, RR>(T, a, me, sub, active, flag);
    cn_class<8, 5, true, false, ET, RR>(T, a, me, sub, active, flag);
    cn_class<9, 4, true, false, ET, RR>(T, a, me, sub, active, flag);
    cn_class<10, 3, true, false, ET, RR>(T, a, me, sub, active, flag);
}

// One core variable (column-block j, lifted index c = me.r): LQ = LLRin + sum_i Lr(i) in ascending
// row-block order = ascending check index (py5gphy/ldpc/nr_ldpc_decode.py:126).
template <bool RR>
__device__ __forceinline__ void vn_col(const DecTab &T, const Me &me, int j, bool active)
{
    const ColInfo ci = T.col[j];
    float lv = 0.f;  // the 2Zc punctured systematic bits start at LLR 0 (:43)
    if (ci.llr_off != kNoLlr) lv = load_llr<RR>(me, ci.llr_off);
    const uint32_t c8 = (uint32_t)me.r * 8u, Z8 = (uint32_t)T.Z8;
    float acc = 0.f;
    const uint4 *q = me.tvn + ci.q0;
    const uint4 *qw = q + ci.nwide, *qe = q + ci.deg;
    for (; q < qw; ++q) {  // rows with 32-bit sign words (BG1 rows 0-3) come first: ascending row order
        const uint4 en = *q;
        const uint32_t t = c8 + en.y, rel = en.x + min(t, t - Z8);
        const float2 m = *reinterpret_cast<const float2 *>(me.mags + rel);
        const uint32_t bits = *reinterpret_cast<const uint32_t *>(me.bw + (rel >> 1));
        acc = __fadd_rn(acc, record_lr(m, ((bits ^ en.z) & 0x1f000000u) == 0, bits << (en.z & 31u)));
    }
#pragma unroll 4
    for (; q < qe; ++q) {
        const uint4 en = *q;
        const uint32_t t = c8 + en.y, rel = en.x + min(t, t - Z8);
        const float2 m = *reinterpret_cast<const float2 *>(me.mags + rel);
        const uint32_t bits = *reinterpret_cast<const uint16_t *>(me.bn + (rel >> 2));
        acc = __fadd_rn(acc, record_lr(m, ((bits ^ en.z) & 0xf000u) == 0, bits << (en.z & 31u)));
    }
    if (active) *reinterpret_cast<float *>(me.slot + ci.lq_off + (c8 >> 1)) = __fadd_rn(lv, acc);
}

// Syndrome pass without a record update, with the post-loop tie rule LQ<=0 -> 1
// (py5gphy/ldpc/nr_ldpc_decode.py:134-143).  Also stores the hard decisions of the extension
// variables, which have no resident posterior.
template <bool RR>
__device__ __forceinline__ void final_row(const DecTab &T, const Me &me, int i, bool active, int *flag)
{
    const RowInfo ri = T.row[i];
    const uint32_t r4 = (uint32_t)me.r * 4u;
    const int deg = (i + 1 < T.nrows ? T.row[i + 1].e0 : T.nedges) - ri.e0;
    const int ncoredeg = (i >= 4) ? deg - 1 : deg;
    uint32_t synd = 0;
    for (int k = 0; k < ncoredeg; ++k) {
        const uint2 ew = me.tcn[ri.e0 + k];
        const uint32_t t = r4 + ew.y;
        const float x = *reinterpret_cast<const float *>(me.slot + ew.x + min(t, t - (uint32_t)T.Z4));
        synd ^= (x <= 0.f) ? 1u : 0u;
    }
    if (i >= 4) {
        const uint32_t rel = ri.mrel + 2 * r4;
        const float2 m = *reinterpret_cast<const float2 *>(me.mags + rel);
        const uint32_t bits = *reinterpret_cast<const uint16_t *>(me.bn + (rel >> 2));
        const float lr = record_lr(m, ((bits ^ ((uint32_t)(deg - 1) << 12)) & 0xf000u) == 0, bits << 31);
        const float x = __fadd_rn(load_llr<RR>(me, ri.ext_llr), lr);
        const bool hb1 = x <= 0.f;
        const uint32_t hb = __ballot_sync(0xffffffffu, hb1);
        if (active && (me.r & (T.lanes - 1)) == 0)
            reinterpret_cast<uint32_t *>(me.slot + T.off_ext)[(i - 4) * T.tiles + (me.r >> 5)] = hb;
        synd ^= hb1 ? 1u : 0u;
    }
    if (active && synd) flag[me.g] = 1;
}

template <int MAXNT, bool ET, bool RR = false>
__global__ void __launch_bounds__(MAXNT, 1)
decode_minsum_kernel(const __grid_constant__ DecTab T, const __grid_constant__ DecArgs a)
{
    extern __shared__ __align__(16) char smem[];
    __shared__ int s_flag[2][kMaxG];
    const int Zc = T.Zc, tid = threadIdx.x, lane = tid & 31, NT = blockDim.x;
    const int warp = __reduce_min_sync(0xffffffffu, tid >> 5);  // warp-uniform (REDUX writes a uniform register)
    const int G = T.G, cb0 = blockIdx.x * G;

    // ---- the (codeblock, r) this thread owns, and the row-/column-block subset of its warp
    const int colm = warp % T.ncolumns, sub = warp / T.ncolumns;
    Me me;
    {
        const int g = (colm / T.tiles) * T.per + (lane >> T.lanes_log2);
        const int r = (colm % T.tiles) * 32 + (lane & (T.lanes - 1));
        const int cb = cb0 + g;
        me.valid = (r < Zc) && (cb < a.B);
        me.g = g;
        me.r = (r < Zc) ? r : 0;
        me.slot = smem + g * T.slot_bytes;
        me.mags = me.slot + T.off_mags;
        me.bw = me.slot + T.off_bw;
        me.bn = me.slot + T.off_bn;
        me.llr = reinterpret_cast<const char *>((RR ? a.rr.scratch : a.llr) + (size_t)cb0 * T.N);
        me.llr_off = (uint32_t)(((cb < a.B) ? g : a.B - 1 - cb0) * T.N + me.r) * 4u;
        me.tcn = reinterpret_cast<const uint2 *>(smem + T.tab_cn);
        me.tvn = reinterpret_cast<const uint4 *>(smem + T.tab_vn);
    }

    if constexpr (RR) {
        // the LLR load is the rate recovery (+ HARQ combining) of the CTA's codeblocks: received sequence -> their fp32 rows
        // (read back below through L2) and the float64 soft buffer the caller keeps.  Shared memory is not in use yet.
        for (int g = 0; g < G && cb0 + g < a.B; ++g)
            rr_row(a.rr, cb0 + g, T.N, a.rr.scratch + (size_t)(cb0 + g) * T.N, reinterpret_cast<double *>(smem));
        __syncthreads();
    }
    // ---- init: records = 0 (Lr = 0, :101), tables staged, LQ = LLRin (:94) with the punctured columns at 0 (:43)
    {
        uint32_t *w = reinterpret_cast<uint32_t *>(smem);
        const int nw = G * T.slot_bytes / 4;
        for (int t = tid; t < nw; t += NT) w[t] = 0;
        uint2 *tc = reinterpret_cast<uint2 *>(smem + T.tab_cn);
        for (int t = tid; t < T.nedges; t += NT) tc[t] = T.cn_edge[t];
        uint4 *tv = reinterpret_cast<uint4 *>(smem + T.tab_vn);
        for (int t = tid; t < T.nventries; t += NT) tv[t] = T.vn_entry[t];
        if (tid < kMaxG) { s_flag[0][tid] = 0; s_flag[1][tid] = 0; }
    }
    __syncthreads();
    uint32_t donemask = 0;
    for (int g = 0; g < G; ++g)
        if (cb0 + g >= a.B) donemask |= 1u << g;
    const uint32_t fullmask = (G >= 32) ? 0xffffffffu : ((1u << G) - 1u);
    if (me.valid)
        for (int o = sub; o < T.ncore; o += T.S) {
            const ColInfo ci = T.col[o];
            if (ci.llr_off != kNoLlr) *reinterpret_cast<float *>(me.slot + ci.lq_off + me.r * 4) = load_llr<RR>(me, ci.llr_off);
        }
    __syncthreads();

    uint32_t et_mask = 0;  // codeblocks that left through the in-loop syndrome check (tie rule LQ<0)
    int my_iters = a.max_iter;
    int it = 0;
    for (; it < a.max_iter; ++it) {
        int *flag = s_flag[it & 1];
        const bool active = me.valid && !((donemask >> me.g) & 1u);
        // ---- check-node pass (+ syndrome of the current hard decisions when ET)
        if (__any_sync(0xffffffffu, active)) cn_pass<ET, RR>(T, a, me, sub, active, flag);
        __syncthreads();
        if (ET) {
            for (int g = 0; g < G; ++g)
                if (!((donemask >> g) & 1u) && flag[g] == 0) {
                    donemask |= 1u << g;
                    et_mask |= 1u << g;
                    if (tid == g) my_iters = it;
                }
            if (tid < kMaxG) s_flag[(it + 1) & 1][tid] = 0;
            if (donemask == fullmask) break;
        }
        // ---- variable-node pass
        const bool active2 = me.valid && !((donemask >> me.g) & 1u);
        if (__any_sync(0xffffffffu, active2)) {
            const int e = T.vn_end[sub];
            for (int o = T.vn_begin[sub]; o < e; ++o) vn_col<RR>(T, me, T.vn_list[o], active2);
        }
        __syncthreads();
    }

    // ---- final decision + syndrome for the codeblocks that did not leave early
    uint32_t okmask = et_mask;
    if (donemask != fullmask) {
        int *flag = s_flag[it & 1];  // cleared, not yet written
        const bool active = me.valid && !((donemask >> me.g) & 1u);
        for (int o = sub; o < T.nrows; o += T.S) final_row<RR>(T, me, o, active, flag);
        __syncthreads();
        for (int g = 0; g < G; ++g)
            if (!((donemask >> g) & 1u) && flag[g] == 0) okmask |= 1u << g;
    }

    // ---- outputs
    if (tid < G && cb0 + tid < a.B) {
        if (a.status) a.status[cb0 + tid] = (okmask >> tid) & 1u;
        if (a.iters) a.iters[cb0 + tid] = my_iters;
    }
    const int NW = NT >> 5;
    for (int g = 0; g < G; ++g) {
        const int cb = cb0 + g;
        if (cb >= a.B) break;
        const char *slot = smem + g * T.slot_bytes;
        const float *LQ = reinterpret_cast<const float *>(slot);
        const uint32_t *ext = reinterpret_cast<const uint32_t *>(slot + T.off_ext);
        const bool et = (et_mask >> g) & 1u;
        const int subpos = (g % T.per) << T.lanes_log2;
        if (a.ck) {
            int8_t *out = a.ck + (size_t)cb * T.Nfull;
            const int ncoreN = T.ncore * Zc;
            for (int n = tid; n < T.Nfull; n += NT) {
                int bit;
                if (n < ncoreN) {
                    const float x = LQ[n];
                    bit = et ? (x < 0.f) : (x <= 0.f);
                } else {
                    const int m = n - ncoreN, i4 = m / Zc, r = m - i4 * Zc;
                    bit = (ext[i4 * T.tiles + (r >> 5)] >> (subpos + (r & 31))) & 1u;
                }
                out[n] = (int8_t)bit;
            }
        }
        if (a.info) {
            const int nwords = (T.K + 31) / 32;
            for (int w = (tid >> 5); w < nwords; w += NW) {
                const int n = 32 * w + lane;
                bool bit = false;
                if (n < T.K) { const float x = LQ[n]; bit = et ? (x < 0.f) : (x <= 0.f); }
                const uint32_t word = __ballot_sync(0xffffffffu, bit);
                if (lane == 0) a.info[(size_t)cb * nwords + w] = word;
            }
        }
    }
}

constexpr int kSmemMax = 227 * 1024 - 512;  // 227 KB per CTA minus the static flags

int class_of(int i, int deg)
{
    if (i < 4) return deg == 19 ? 0 : (deg == 10 ? 1 : (deg == 8 ? 2 : -1));
    return (deg >= 3 && deg <= 10) ? 13 - deg : -1;  // 10 -> 3 ... 3 -> 10
}

// Host: derive the kernel tables and launch geometry for one (bgn, Zc).
int build_dec_tab(const QcCfg &c, DecTab *T)
{
    std::memset(T, 0, sizeof(*T));
    const int Zc = c.Zc;
    T->Zc = Zc; T->Z4 = 4 * Zc; T->Z8 = 8 * Zc;
    T->nrows = c.nrows; T->ncore = c.ncore; T->kb = c.kb; T->K = c.K; T->N = c.N; T->Nfull = c.Nfull;
    T->tiles = c.tiles; T->lanes = c.lanes; T->lanes_log2 = c.lanes_log2; T->per = c.per;
    T->nedges = c.rowptr[c.nrows]; T->nventries = c.colptr[c.ncore];
    // rows with 32-bit sign words must be the leading row-blocks (BG1: 0-3, BG2: none)
    int nwide_rows = 0;
    for (int i = 0; i < c.nrows; ++i) {
        if (c.wide[i]) { if (i != nwide_rows) return NRLDPC_EINVAL; ++nwide_rows; }
    }
    // slot layout
    T->off_mags = (c.ncore * Zc * 4 + 7) & ~7;
    T->off_bw = T->off_mags + c.nrows * Zc * 8;
    const int bn_start = T->off_bw + nwide_rows * Zc * 4;
    T->off_bn = bn_start - nwide_rows * Zc * 2;  // so that row i >= nwide_rows lands at off_bn + (i*Zc + r)*2
    T->off_ext = (bn_start + (c.nrows - nwide_rows) * Zc * 2 + 3) & ~3;
    T->slot_bytes = (T->off_ext + (c.nrows - 4) * c.tiles * 4 + 15) & ~15;
    const int tab_bytes = T->nedges * 8 + T->nventries * 16;
    // geometry: as many codeblocks per CTA as fit, at most 16 warp columns and 32 codeblocks
    int gmax = (kSmemMax - tab_bytes) / T->slot_bytes;
    if (gmax < 1) return NRLDPC_EINVAL;
    gmax = std::min(gmax, kMaxG);
    int G = std::min(gmax, std::max(1, 16 / c.tiles) * c.per);
    if (c.per > 1) G = std::max(c.per, (G / c.per) * c.per);
    if (G > gmax) return NRLDPC_EINVAL;
    T->G = G;
    T->ncolumns = (G / c.per) * c.tiles;
    T->S = std::min(kMaxS, std::max(1, 32 / T->ncolumns));
    T->nwarps = T->ncolumns * T->S;
    if (T->nwarps > 32) return NRLDPC_EINVAL;
    T->tab_vn = G * T->slot_bytes;  // 16-byte aligned
    T->tab_cn = T->tab_vn + T->nventries * 16;
    T->smem_bytes = T->tab_cn + T->nedges * 8;
    const int S = T->S;

    std::vector<int> deg(c.nrows);
    for (int i = 0; i < c.nrows; ++i) {
        RowInfo &r = T->row[i];
        r.mrel = (uint32_t)i * Zc * 8;
        r.ext_llr = (i >= 4) ? (uint32_t)(c.kb + i - 2) * Zc * 4 : 0;
        r.e0 = c.rowptr[i];
        deg[i] = c.rowptr[i + 1] - c.rowptr[i];
        if (class_of(i, deg[i]) < 0 || (c.wide[i] != 0) != (deg[i] == 19)) return NRLDPC_EINVAL;
    }
    for (int e = 0; e < T->nedges; ++e) {
        const int j = c.edge[e] & 0xff, P = c.edge[e] >> 8;
        T->cn_edge[e] = make_uint2((uint32_t)(j < c.ncore ? j * Zc * 4 : 0), (uint32_t)P * 4);
    }
    for (int j = 0; j < c.ncore; ++j) {
        ColInfo &ci = T->col[j];
        ci.lq_off = j * Zc * 4;
        ci.llr_off = j >= 2 ? (uint32_t)(j - 2) * Zc * 4 : kNoLlr;
        ci.q0 = c.colptr[j];
        ci.deg = (uint8_t)(c.colptr[j + 1] - c.colptr[j]);
        int nwide = 0;
        for (int q = c.colptr[j]; q < c.colptr[j + 1]; ++q) {
            const uint32_t en = c.centry[q];
            const int i = en & 63, k = (en >> 6) & 31, bitpos = (en >> 11) & 31, back = en >> 16;
            const bool wide = c.wide[i];
            if (wide) { if (nwide != q - c.colptr[j]) return NRLDPC_EINVAL; ++nwide; }  // wide rows lead (ascending i)
            T->vn_entry[q] = make_uint4(T->row[i].mrel, (uint32_t)back * 8,
                                        ((uint32_t)k << (wide ? 24 : 12)) | (uint32_t)(31 - bitpos), 0u);
        }
        ci.nwide = (uint8_t)nwide;
    }
    // check rows: per class, deal the rows round-robin to the S warp groups, continuing the rotation
    // across classes so that no group collects all the remainders
    {
        std::vector<std::vector<std::vector<int>>> lists(kNumCls, std::vector<std::vector<int>>(S));
        int rot = 0;
        for (int cls = 0; cls < kNumCls; ++cls)
            for (int i = 0; i < c.nrows; ++i)
                if (class_of(i, deg[i]) == cls) lists[cls][rot++ % S].push_back(i);
        int pos = 0;
        for (int cls = 0; cls < kNumCls; ++cls)
            for (int s = 0; s < S; ++s) {
                T->cls_begin[cls * kMaxS + s] = (uint16_t)pos;
                for (int i : lists[cls][s]) T->cn_list[pos++] = (uint8_t)i;
                T->cls_end[cls * kMaxS + s] = (uint16_t)pos;
            }
        if (pos != c.nrows) return NRLDPC_EINVAL;
    }
    // core columns: longest-processing-time-first onto the S groups
    {
        std::vector<int> idx(c.ncore);
        for (int j = 0; j < c.ncore; ++j) idx[j] = j;
        std::stable_sort(idx.begin(), idx.end(), [&](int x, int y) { return T->col[x].deg > T->col[y].deg; });
        std::vector<std::vector<int>> grp(S);
        std::vector<int> load(S, 0);
        for (int j : idx) {
            int best = 0;
            for (int s = 1; s < S; ++s) if (load[s] < load[best]) best = s;
            grp[best].push_back(j);
            load[best] += T->col[j].deg + 3;
        }
        int pos = 0;
        for (int s = 0; s < S; ++s) {
            T->vn_begin[s] = (uint16_t)pos;
            for (int j : grp[s]) T->vn_list[pos++] = (uint8_t)j;
            T->vn_end[s] = (uint16_t)pos;
        }
    }
    return NRLDPC_OK;
}

const DecTab *get_dec_tab(const QcCfg &cfg)
{
    static std::mutex mu;
    static std::map<int, DecTab *> cache;
    std::lock_guard<std::mutex> lk(mu);
    const int key = cfg.bgn * 1024 + cfg.Zc;
    auto it = cache.find(key);
    if (it != cache.end()) return it->second;
    DecTab *T = new DecTab;
    if (build_dec_tab(cfg, T) != NRLDPC_OK) { delete T; return nullptr; }
    cache[key] = T;
    return T;
}


}  // namespace

int decode_minsum_geometry(const QcCfg &cfg, int *G_out, int *threads, int *smem)
{
    if (!std::getenv("NRLDPC_NO_SPEC") && decode_spec_geometry(cfg.bgn, cfg.Zc, threads, smem)) {
        if (G_out) *G_out = 1;
        return NRLDPC_OK;
    }
    const DecTab *T = get_dec_tab(cfg);
    if (!T) return NRLDPC_EINVAL;
    if (G_out) *G_out = T->G;
    if (threads) *threads = T->nwarps * 32;
    if (smem) *smem = T->smem_bytes;
    return NRLDPC_OK;
}

int launch_decode_minsum(const QcCfg &cfg, const float *d_llr, int B, int max_iter, float alpha, float beta,
                         int early_term, int8_t *d_ck, uint32_t *d_info, uint8_t *d_status, int32_t *d_iters,
                         cudaStream_t s, const RrArgs *rr)
{
    if (B <= 0) return NRLDPC_OK;
    const DecTab *T = get_dec_tab(cfg);
    if (!T) { set_error("decode_minsum: no kernel geometry for bgn=%d Zc=%d", cfg.bgn, cfg.Zc); return NRLDPC_EINVAL; }
    DecArgs a;
    a.llr = d_llr; a.B = B; a.max_iter = max_iter; a.alpha = alpha; a.beta = beta;
    a.ck = d_ck; a.info = d_info; a.status = d_status; a.iters = d_iters;
    if (rr) {
        if (!early_term || !rr->src || !rr->E || !rr->goff) { set_error("decode: the fused rate recovery needs early_term = 1 and src, E, goff"); return NRLDPC_EINVAL; }
        a.rr = *rr;
    }
    static const bool no_spec = std::getenv("NRLDPC_NO_SPEC") != nullptr;  // tests: force the generic kernel
    if (!no_spec) {
        bool handled = false;
        const int rc = launch_decode_spec(cfg.bgn, cfg.Zc, a, early_term, s, &handled);
        if (handled) return rc;
    }
    const int grid = (B + T->G - 1) / T->G, nt = T->nwarps * 32, smem = T->smem_bytes;
    ScratchBuf rows;  // fused rate recovery: the recovered fp32 rows [B, N], stream-ordered
    if (rr && !a.rr.scratch) {
        NRLDPC_CUDA(rows.alloc((size_t)B * T->N * sizeof(float), s));
        a.rr.scratch = rows.as<float>();
    }
    auto launch = [&](auto kern) -> int {
        NRLDPC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        kern<<<grid, nt, smem, s>>>(*T, a);
        NRLDPC_CUDA(cudaGetLastError());
        return NRLDPC_OK;
    };
    if (rr) return nt <= 768 ? launch(decode_minsum_kernel<768, true, true>) : launch(decode_minsum_kernel<1024, true, true>);
    if (nt <= 768) return early_term ? launch(decode_minsum_kernel<768, true>) : launch(decode_minsum_kernel<768, false>);
    return early_term ? launch(decode_minsum_kernel<1024, true>) : launch(decode_minsum_kernel<1024, false>);
}

}  // namespace nrldpc
