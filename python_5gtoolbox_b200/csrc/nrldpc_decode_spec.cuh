// nrldpc_decode_spec.cuh -- flooding min-sum decoder kernels with (bgn, Zc) known at compile time.
// (template code; one translation unit per instantiated (bgn, Zc): nrldpc_decode_spec_bgX_ZZZ.cu)
//
// Same algorithm, schedule and fp32 operation order as nrldpc_decode_qc.cu (the table-driven kernel):
// py5gphy/ldpc/nr_ldpc_decode.py:51-143 (decode_ldpc) and :178-227 (_min_sum_process).  The whole base
// graph is unrolled at compile time from the constexpr TS 38.212 tables, so every shift, column offset,
// record offset and sign-bit position is an immediate and each warp group runs straight-line code.
//
// One persistent CTA per SM decodes one codeblock at a time with its whole state in shared memory (see
// DESIGN.md 3-4).  HBM sees 4N + K/8 bytes per codeblock and is ~1 % busy; the kernel is bound by warp
// instruction issue and the half-rate ALU pipe (LOP3/SHF/FMNMX/ISETP/SEL), so the design is about
// instructions per edge:
//   * circulant shifts cost no instruction: every shared-memory array that is read through a rotation
//     (posteriors LQ in the check pass, row records in the variable pass) carries 32 extra elements that
//     mirror its first 32, and the (r + P) mod Zc wrap becomes a WARP-UNIFORM choice between offset 0
//     and -Zc (tile >= ceil((Zc-P)/32)): LDS [R + UR + imm] or a per-lane base pre-added once;
//   * the argmin of a check row is found after the two minima are known, on the FMA pipe:
//     sat((min1 - |Lq|) * -inf) is exactly 1 on every edge but the minimum one;
//   * sign/argmin words are 8-bit for rows of degree <= 5, 16-bit up to degree 12 and 32-bit above,
//     interleaved so that all of them have a 4-byte stride (one set of wrap offsets, no bank conflicts);
//   * the two magnitudes of a row live in separate arrays a multiple of 128 B apart: the variable pass
//     selects the address and reads ONE 128-byte wavefront per edge;
//   * rows / columns are software-pipelined by hand across the record stores (the compiler cannot prove
//     that the shared-memory stores of one row do not alias the loads of the next).
// Compile-time switches (NRLDPC_*) keep the measured alternatives buildable: tools/build_variant.sh.
#pragma once
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <utility>

#include "nrldpc_decode.cuh"
#define NRLDPC_TABLE static constexpr
#include "../../include/nrldpc_bg_tables.inc"

namespace nrldpc {

namespace {

#ifndef NRLDPC_MAX_S
#define NRLDPC_MAX_S 16
#endif
constexpr int kMaxS = 16;  // warp groups per r-tile (array bound)
#ifndef NRLDPC_MULTI_CTA_BELOW
#define NRLDPC_MULTI_CTA_BELOW 144  // (176 measured: BG1 Zc=160 785 vs 826, Zc=144 712 vs 745 G edge-iterations/s -- one CTA of 30 warps wins there)
#endif
constexpr int kMultiCtaBelow = NRLDPC_MULTI_CTA_BELOW;  // lifting sizes below this run several persistent CTAs per SM
constexpr int kNoVariant = 1;  // launch_spec: this (early_term) combination is not instantiated -> table-driven kernel
#ifndef NRLDPC_COLD_FMA
#define NRLDPC_COLD_FMA 1
#endif
#ifndef NRLDPC_VN_PSEL
#define NRLDPC_VN_PSEL 1
#endif
#ifndef NRLDPC_SUB_INTERLEAVE
#define NRLDPC_SUB_INTERLEAVE 0
#endif
#ifndef NRLDPC_SIGN_FMA
#define NRLDPC_SIGN_FMA 0
#endif
#ifndef NRLDPC_MASKREG
#define NRLDPC_MASKREG 1     // argmin-field test as ONE LOP3 with predicate output: ((bits ^ k) & mask register) != 0
#endif
#ifndef NRLDPC_FIRST_IT
#define NRLDPC_FIRST_IT 1    // first check pass without record decode (all records are zero: Lq = LQ): 1 = kernels without
                             // early termination only (with it, the extra code costs more than it saves: measured), 2 = all
#endif
#ifndef NRLDPC_MIRROR_LOOP
#define NRLDPC_MIRROR_LOOP 0   // 1: the mirror stores of a check row sit in a one-trip loop (trip count opaque to ptxas), which
#endif                         // keeps them behind a real warp-uniform branch instead of predicated off in every warp
#ifndef NRLDPC_TILE_ROT
#define NRLDPC_TILE_ROT 0      // 1: group g's tiles rotated by g, so the tile-0 warps of the groups sit on different schedulers
#endif
#ifndef NRLDPC_COLD_FMA_MAXDEG
#define NRLDPC_COLD_FMA_MAXDEG 99  // rows of a larger degree find the argmin with FADD + SHF + FLO (ALU pipe) instead of 3 FMA-pipe ops
#endif
#ifndef NRLDPC_GROUP_SYNC_CN
#define NRLDPC_GROUP_SYNC_CN 0  // > 0: named barrier among the warps of a group every this many check rows (bounds their drift
#endif                          // in the straight-line code, which is streamed from L2: the instruction caches hold 6 KB / 32 KB)
#ifndef NRLDPC_GROUP_SYNC_VN
#define NRLDPC_GROUP_SYNC_VN 0  // same, every this many columns of the variable pass
#endif
#ifndef NRLDPC_TWO_CTAS
#define NRLDPC_TWO_CTAS 0  // two persistent CTAs per SM where two codeblock states fit in its shared memory (small Zc): the
#endif                     // check pass of one could overlap the variable pass of the other -- measured: +1 % (BG1 Zc=176), -11 % (BG2 Zc=208)
#ifndef NRLDPC_PF_LLR2
#define NRLDPC_PF_LLR2 1  // channel LLR of a check row's extension variable fetched two rows ahead (L2 latency) instead of one
#endif
#ifndef NRLDPC_PF_VLLR
#define NRLDPC_PF_VLLR 2  // channel LLR of a core column fetched this many columns ahead in the variable pass (1 or 2)
#endif
#ifndef NRLDPC_FINAL_PACKED
#define NRLDPC_FINAL_PACKED 1  // final syndrome on bit-packed hard decisions (lifting sizes that are a multiple of 32)
#endif
#ifndef NRLDPC_XSIGN
#define NRLDPC_XSIGN 1   // 1: rows of odd degree apply the old message's sign to the posterior instead of the magnitude
#endif                   // (Lq' = s Lq = s x - |Lr|: two predicated FADDs choose the magnitude, no FSEL; see cn_edge_s)
#ifndef NRLDPC_XSIGN_MIN_ZC
#define NRLDPC_XSIGN_MIN_ZC 288  // measured: +1.0-1.6 % at Zc = 288 ... 384 (24-30 warps, ALU-pipe bound), -0.5 ... -3 % at 144 ... 256
#endif
#ifndef NRLDPC_LLR_OPAQUE
#define NRLDPC_LLR_OPAQUE 0
#endif
#ifndef NRLDPC_PF_SUM
#define NRLDPC_PF_SUM 20  // prefetch the next check row's inputs when deg(current) + deg(next) <= this (registers)
#endif

template <int BGN> struct Graph;
template <> struct Graph<1> {
    static constexpr int rows = NRLDPC_BG1_ROWS, cols = NRLDPC_BG1_COLS, nnz = NRLDPC_BG1_NNZ, kb = 22;
    static constexpr int rowptr(int i) { return nrldpc_bg1_rowptr[i]; }
    static constexpr int col(int e) { return nrldpc_bg1_col[e]; }
    static constexpr int shift(int s, int e) { return nrldpc_bg1_shift[s][e]; }
};
template <> struct Graph<2> {
    static constexpr int rows = NRLDPC_BG2_ROWS, cols = NRLDPC_BG2_COLS, nnz = NRLDPC_BG2_NNZ, kb = 10;
    static constexpr int rowptr(int i) { return nrldpc_bg2_rowptr[i]; }
    static constexpr int col(int e) { return nrldpc_bg2_col[e]; }
    static constexpr int shift(int s, int e) { return nrldpc_bg2_shift[s][e]; }
};

constexpr int ils_of(int Zc)
{
    constexpr int a[8] = {2, 3, 5, 7, 9, 11, 13, 15};
    for (int s = 0; s < 8; ++s)
        for (int z = a[s]; z <= 384; z *= 2)
            if (z == Zc) return s;
    return -1;
}

// RR_: the kernel recovers its LLRs itself (rate recovery + HARQ combining fused into the LLR load, DecArgs::rr) and reads
// them back from a row it wrote (coherent loads instead of the read-only path).
template <int BGN, int ZC_, bool RR_ = false> struct Code {
    using G = Graph<BGN>;
    static constexpr int bgn = BGN, ZC = ZC_, iLS = ils_of(ZC_);
    static constexpr bool RR = RR_;
    // Any lifting size >= 64: r-tile 0 must be full (it writes the 32 mirrored elements).  In a partial last
    // tile the lanes beyond Zc duplicate lane Zc-1 (same loads, same stores of the same values).
    // (below 32 the single r-tile is partly filled: its lanes beyond Zc duplicate lane Zc-1, and the 32 mirrored elements
    // still cover every rotated read, r + P <= 2 Zc - 2 < Zc + 32.  Measured against the table-driven kernel, which packs
    // several codeblocks into a warp there: BG1 Zc = 20 ... 30 1.24-1.43x, Zc = 18 1.07x, Zc = 16 0.59x; BG2 1.0-1.09x at
    // Zc = 20 ... 30 -- so BG1 is instantiated from 20 up, BG2 from 28 up)
    static_assert(iLS >= 0 && ZC_ >= 20, "specialised kernels are instantiated for lifting sizes >= 20");
    static constexpr int nrows = G::rows, kb = G::kb, ncore = G::kb + 4, nnz = G::nnz;
    static constexpr int K = kb * ZC, N = (G::cols - 2) * ZC, Nfull = G::cols * ZC;
    static constexpr int tiles = (ZC + 31) / 32;
    static constexpr int LQS = ZC + 32;  // elements per rotated array: Zc + the 32 mirrored ones
    static constexpr int deg(int i) { return G::rowptr(i + 1) - G::rowptr(i); }
    static constexpr int P(int e) { return G::shift(iLS, e) % ZC; }
    static constexpr int row_of(int e) { int i = 0; while (G::rowptr(i + 1) <= e) ++i; return i; }
    // column view, ascending row-block = the reference's summation order (:126)
    static constexpr int cdeg(int j) { int n = 0; for (int e = 0; e < nnz; ++e) n += (G::col(e) == j); return n; }
    static constexpr int cedge(int j, int n)
    {
        for (int e = 0; e < nnz; ++e)
            if (G::col(e) == j && n-- == 0) return e;
        return -1;
    }
    // Tile index from which the rotation by `shift` wraps for the whole warp: lanes of tile T read
    // index r + shift - (T >= wrap_tile ? Zc : 0), which stays inside [0, Zc + 32).
    static constexpr int wrap_tile(int shift) { return (ZC - shift + 31) / 32; }

    // sign/argmin word of a row: kind 0 = 8 bit (deg <= 5: 5 signs + 3 index bits), 1 = 16 bit
    // (deg <= 12: 12 + 4), 2 = 32 bit (deg <= 24: 24 + 5).  Bit deg-1-k = sign of Lr on edge k.
    static constexpr int kind(int i) { return deg(i) > 12 ? 2 : (deg(i) > 5 ? 1 : 0); }
    static constexpr int rank(int i) { int n = 0; for (int x = 0; x < i; ++x) n += (kind(x) == kind(i)); return n; }
    static constexpr int count(int kd) { int n = 0; for (int x = 0; x < nrows; ++x) n += (kind(x) == kd); return n; }
    static constexpr int idx_shift(int i) { return kind(i) == 2 ? 24 : (kind(i) == 1 ? 12 : 5); }
    static constexpr uint32_t idx_mask(int i) { return kind(i) == 2 ? 0x1f000000u : (kind(i) == 1 ? 0xf000u : 0xe0u); }
    // shared-memory layout (bytes)
    static constexpr int off_lq = 0;                                          // float LQ[ncore][LQS]
    static constexpr int off_mags = off_lq + ncore * LQS * 4;                 // float mag1[nrows][LQS], then mag2[nrows][LQS]
    static constexpr int mag2_dist = (nrows * LQS * 4 + 127) & ~127;          // bytes from a row's mag1 to its mag2 (multiple of 128: same banks)
    static constexpr int off_b32 = off_mags + 2 * mag2_dist;                  // uint32 [count(2)][LQS]
    static constexpr int off_b16 = off_b32 + count(2) * LQS * 4;              // uint16 pairs of rows interleaved
    static constexpr int off_b8 = off_b16 + ((count(1) + 1) / 2) * LQS * 4;   // uint8 quads of rows interleaved
    static constexpr int off_ext = off_b8 + ((count(0) + 3) / 4) * LQS * 4;   // packed hard bits of the extension columns
    static constexpr int smem_bytes = (off_ext + (nrows - 4) * tiles * 4 + 15) & ~15;
    // CTAs per SM: two when two codeblock states (+ 1 KB of system shared memory each) fit in the SM's 228 KB; each then
    // runs 16 / tiles warp groups (at most 512 threads, so that two CTAs keep >= 64 registers per thread)
    // Small lifting sizes (ZC < 144): one codeblock no longer fills an SM, so as many persistent CTAs as fit in its shared
    // memory run side by side (each its own codeblock, its own barriers: the check pass of one overlaps the variable pass of
    // another), with the SM's 32 warps shared between them.
    static constexpr int fit = (228 * 1024) / (smem_bytes + 1024);
    static constexpr int ctas_small = fit > 8 ? 8 : (fit < 1 ? 1 : fit);
    static constexpr int ctas = ZC < kMultiCtaBelow ? ctas_small
                                         : ((NRLDPC_TWO_CTAS && 2 * (smem_bytes + 1024) <= 228 * 1024 && 16 / tiles >= 1) ? 2 : 1);
    static constexpr int Smax0 = (32 / ctas) / tiles;
    static constexpr int Smax = Smax0 < 1 ? 1 : Smax0;
    // At most 3 warp groups from Zc = 144 up (4 with 7 r-tiles; 2 at Zc = 320, whose 3 x 10 warps would cap the registers at 64:
    // +6.5-7 % with fixed iterations, +9-14 % with early termination at an operating point, both base graphs).  Measured on Zc = 144 ... 256 (profiles/r2_zc_groups_ab.md):
    // an iteration takes ~9 us whatever the lifting size and whether 12 or 30 warps run it -- ncu shows the warps waiting
    // for instructions (no_instruction 8-11 per issue at Zc = 176 / 144 against 0.4 at Zc = 384), every group streams
    // its own straight-line code -- so more groups buy nothing, and 3 x tiles warps keep > 64 registers per thread.
    // Below 144 (several CTAs per SM) the best group count is irregular and depends on the operating point: one group per
    // CTA is 20-27 % faster at BG1 Zc = 56 / 60 / 64 when nothing converges (-3 dB, 10 iterations), 40 % slower at
    // 44 / 52 / 112, and 4-9 % SLOWER at the same 56 / 60 / 64 at +1 dB, where codeblocks leave after ~7 iterations and
    // the fewer warps pay at the codeblock boundaries.  Only the BG2 sizes that win at both points carry a count here
    // (profiles/r2_zc_groups_ab.md, same box).  -DNRLDPC_MAX_S=n (n <= 3) overrides, -DNRLDPC_GROUP_TABLE=0 disables.
    static constexpr int small_groups()
    {
        if (BGN == 2 && (ZC == 44 || ZC == 72 || ZC == 88 || ZC == 96 || ZC == 120 || ZC == 128)) return 1;
        return NRLDPC_MAX_S;
    }
#ifndef NRLDPC_GROUP_TABLE
#define NRLDPC_GROUP_TABLE 1  // 0: as many groups as 32 warps per SM allow (the layout before the measurements above)
#endif
    static constexpr int Scap = (NRLDPC_MAX_S <= 3 || !NRLDPC_GROUP_TABLE) ? NRLDPC_MAX_S
                              : (ZC >= kMultiCtaBelow ? (tiles == 7 ? 4 : (tiles == 10 ? 2 : 3)) : small_groups());
    static constexpr int S = Smax < Scap ? Smax : Scap, nwarps = tiles * S;
    static constexpr int lq_base(int j) { return off_lq + j * LQS * 4; }
    static constexpr int mags_base(int i) { return off_mags + i * LQS * 4; }  // mag1 of check 0; mag2 sits mag2_dist further
    static constexpr int bits_base(int i)  // the word of check r sits at bits_base(i) + 4 r
    {
        return kind(i) == 2 ? off_b32 + rank(i) * LQS * 4
             : kind(i) == 1 ? off_b16 + (rank(i) / 2) * LQS * 4 + (rank(i) % 2) * 2
                            : off_b8 + (rank(i) / 4) * LQS * 4 + (rank(i) % 4);
    }
};

struct Deal { int n[kMaxS]; int item[kMaxS][64]; };

// rows by degree (descending, stable), longest-processing-time-first onto the S warp groups with the
// measured cost model of a row: 13 issue slots per edge + 25 per row
template <class C> constexpr Deal deal_rows()
{
    Deal d{};
    int order[64] = {}, load[kMaxS] = {};
    for (int i = 0; i < C::nrows; ++i) order[i] = i;
    for (int i = 1; i < C::nrows; ++i)  // insertion sort, stable
        for (int p = i; p > 0 && C::deg(order[p - 1]) < C::deg(order[p]); --p) { int t = order[p]; order[p] = order[p - 1]; order[p - 1] = t; }
    for (int p = 0; p < C::nrows; ++p) {
        int best = 0;
        for (int s = 1; s < C::S; ++s) if (load[s] < load[best]) best = s;
        d.item[best][d.n[best]++] = order[p];
        load[best] += 13 * C::deg(order[p]) + 25;
    }
    return d;
}
// core columns: longest-processing-time-first onto the S groups
template <class C> constexpr Deal deal_cols()
{
    Deal d{};
    int order[64] = {}, load[kMaxS] = {};
    for (int j = 0; j < C::ncore; ++j) order[j] = j;
    for (int i = 1; i < C::ncore; ++i)
        for (int p = i; p > 0 && C::cdeg(order[p - 1]) < C::cdeg(order[p]); --p) { int t = order[p]; order[p] = order[p - 1]; order[p - 1] = t; }
    for (int p = 0; p < C::ncore; ++p) {
        int best = 0;
        for (int s = 1; s < C::S; ++s) if (load[s] < load[best]) best = s;
        d.item[best][d.n[best]++] = order[p];
        load[best] += 7 * C::cdeg(order[p]) + 6;
    }
    return d;
}
template <class C> inline constexpr Deal kRows = deal_rows<C>();
template <class C> inline constexpr Deal kCols = deal_cols<C>();

// Final syndrome on bit-packed hard decisions (below): lifting sizes that are a multiple of 32 whose packed words fit in the
// dead records of row-blocks 0-3.
template <class C> constexpr bool kFinalPacked = NRLDPC_FINAL_PACKED && C::ZC % 32 == 0 && 2 * C::tiles * C::ncore * 4 <= 4 * C::LQS * 4;
// Early termination: do the check passes store the hard bits of the extension variables every iteration?  Only where they cannot
// be derived at the exit: a codeblock leaves with every parity check satisfied, so the bit of a row-block's degree-1 extension
// variable IS the XOR of its core bits -- one pass over the packed core decisions at the exit (ext_from_core), and only when the
// caller wants ck at all, instead of FSETP + VOTE + predicated STS per extension row and iteration on the ALU-bound check pass.
#ifndef NRLDPC_ET_EXT_AT_EXIT
#define NRLDPC_ET_EXT_AT_EXIT 1
#endif
template <class C> constexpr bool kEtStoresExt = !(NRLDPC_ET_EXT_AT_EXIT && kFinalPacked<C>);

template <class C> struct Th {  // per-thread constants
    char *smem;                 // the CTA's codeblock state
    const float *llr;           // this thread's codeblock LLR row, already offset by r
    uint32_t r4;                // r * 4
    uint32_t r4m2;              // r * 4 + mag2_dist
    char *p4, *p4m2;            // smem + r4, smem + r4m2
    int r, tile;
    int w4[C::tiles + 1];       // warp-uniform wrap offsets in bytes of a 4-byte-stride array: w4[t] = tile >= t ? -4 Zc : 0
    uint32_t mk[3];             // argmin-field masks of the three word kinds, held in (opaque) registers
    int one;                    // 1, opaque to ptxas
};

// Is edge K the argmin edge of the row whose sign/argmin word is `bits`?  With the mask in a register the
// test is one LOP3 with a predicate output, (bits ^ K<<s) & mask != 0, instead of LOP3 + ISETP.
template <class C, int I, int K> __device__ __forceinline__ bool is_argmin(const Th<C> &th, uint32_t bits)
{
#if NRLDPC_MASKREG
    return ((bits ^ ((uint32_t)K << C::idx_shift(I))) & th.mk[C::kind(I)]) == 0;
#else
    return ((bits ^ ((uint32_t)K << C::idx_shift(I))) & C::idx_mask(I)) == 0;
#endif
}

// y - (edge K is the argmin edge of the row ? m.y : m.x): the magnitude is chosen by predicating two FADDs on the argmin test
// (FMA pipe) instead of an FSEL on the ALU pipe, which bounds the check pass.
template <class C, int I, int K> __device__ __forceinline__ float sub_record_mag(const Th<C> &th, float y, float2 m, uint32_t bits)
{
    float q = __fsub_rn(y, m.x);
    asm("{\n\t.reg .pred p;\n\t.reg .b32 t;\n\t"
        "lop3.b32 t, %3, %4, %5, 0x28;\n\t"
        "setp.eq.b32 p, t, 0;\n\t"
        "@p sub.rn.f32 %0, %1, %2;\n\t}"
        : "+f"(q) : "f"(y), "f"(m.y), "r"(bits), "r"((uint32_t)K << C::idx_shift(I)), "r"(th.mk[C::kind(I)]));
    return q;
}

// sign/argmin word access
template <class C, int I> __device__ __forceinline__ uint32_t load_bits(const char *p)
{
    if constexpr (C::kind(I) == 2) return *reinterpret_cast<const uint32_t *>(p);
    else if constexpr (C::kind(I) == 1) return *reinterpret_cast<const uint16_t *>(p);
    else return *reinterpret_cast<const uint8_t *>(p);
}
template <class C, int I> __device__ __forceinline__ void store_bits(char *p, uint32_t v)
{
    if constexpr (C::kind(I) == 2) *reinterpret_cast<uint32_t *>(p) = v;
    else if constexpr (C::kind(I) == 1) *reinterpret_cast<uint16_t *>(p) = (uint16_t)v;
    else *reinterpret_cast<uint8_t *>(p) = (uint8_t)v;
}

// acc = acc << 1 | (top bit of v)
__device__ __forceinline__ uint32_t push_top_bit(uint32_t acc, uint32_t v) { return __funnelshift_l(v, acc, 1); }

// Posterior LQ of the variable on edge e of a check row with lifted index r: LQ[col][(r + P) mod Zc].
template <class C, int E> __device__ __forceinline__ float load_lq_rot(const Th<C> &th)
{
    constexpr int P = C::P(E), t = C::wrap_tile(P);
    return *reinterpret_cast<const float *>(th.smem + (C::lq_base(C::G::col(E)) + 4 * P) + th.r4 + th.w4[t]);
}

// What a check row reads from shared memory: its record and the posteriors of its core variables.
template <class C, int I> struct RowIn {
    static constexpr int NC = (I >= 4) ? C::deg(I) - 1 : C::deg(I);
    float2 m;
    uint32_t bits;
    float x[NC];
};
template <class C, int I, int... K>
__device__ __forceinline__ void load_row_x(RowIn<C, I> &in, const Th<C> &th, std::integer_sequence<int, K...>)
{
    ((in.x[K] = load_lq_rot<C, C::G::rowptr(I) + K>(th)), ...);
}
// r4 (+ the uniform wrap offset) must not be folded into per-lane registers hoisted out of the iteration
// loop: the sum is made opaque here so that the wrap offsets stay in uniform registers (LDS [R + UR + imm]).
template <class C> __device__ __forceinline__ Th<C> opaque(Th<C> th)
{
    asm volatile("" : "+r"(th.r4), "+r"(th.r4m2));
    return th;
}

// FIRST: the first check pass of a codeblock -- every record is zero (Lr = 0, :101), so nothing is read of it.
template <class C, int I, bool FIRST = false> __device__ __forceinline__ RowIn<C, I> load_row(const Th<C> &th_)
{
    const Th<C> th = opaque(th_);
    RowIn<C, I> in;
    if constexpr (FIRST) {
        in.m = make_float2(0.f, 0.f);
        in.bits = 0;
    } else {
        in.m.x = *reinterpret_cast<const float *>(th.smem + C::mags_base(I) + th.r4);
        in.m.y = *reinterpret_cast<const float *>(th.smem + (C::mags_base(I) + C::mag2_dist) + th.r4);
        in.bits = load_bits<C, I>(th.smem + C::bits_base(I) + th.r4);
    }
    load_row_x<C, I>(in, th, std::make_integer_sequence<int, RowIn<C, I>::NC>{});
    return in;
}

// Rows handled with the sign on the posterior (NRLDPC_XSIGN): odd degree only -- the parity of a stored sign word is
// XOR_k (sp ^ sign Lq_k) = (DEG + 1) sp, i.e. 0 for an odd degree, so XOR_k s_k needs no instruction there.
template <class C, int I, bool FIRST> __device__ __forceinline__ constexpr bool xsign_row()
{
    return NRLDPC_XSIGN && C::ZC >= NRLDPC_XSIGN_MIN_ZC && !FIRST && (C::deg(I) % 2 == 1);
}

template <class C, int I, int K, bool ET, bool FIRST>
__device__ __forceinline__ void cn_edge_s(float (&q)[C::deg(I)], float &m1, float &m2, uint32_t &synd, const Th<C> &th,
                                          const RowIn<C, I> &in, uint32_t idxf, float llr_e)
{
    constexpr int DEG = C::deg(I);
    constexpr bool EXT = I >= 4;
    float lr = 0.f;  // FIRST: Lr = +0 on every edge
    if constexpr (xsign_row<C, I, FIRST>()) {
        // Lq' = s Lq with s = the sign of the old message: s x - |Lr| = s (x - s |Lr|) exactly, so |Lq'| = |Lq| and
        // sign(Lq) = s ^ sign(Lq') (a zero Lq' gets the sign s: the sign given to a zero cancels in every message, see
        // cn_row_s).  The row's sign product and sign word are corrected once per row.
        const uint32_t sbit = (in.bits << (31 - (DEG - 1 - K))) & 0x80000000u;
        float x;
        if constexpr (EXT && K == DEG - 1 && !ET) {
            // s x = s LLR + |Lr| and Lq' = s x - |Lr|: both magnitudes by predicated FADDs (x itself is not needed)
            const float t = __uint_as_float(__float_as_uint(llr_e) ^ sbit);
            x = 0.f;
            const float y = sub_record_mag<C, I, K>(th, t, make_float2(-in.m.x, -in.m.y), in.bits);
            q[K] = sub_record_mag<C, I, K>(th, y, in.m, in.bits);
        } else if constexpr (EXT && K == DEG - 1) {
            const bool isidx = is_argmin<C, I, K>(th, in.bits);
            const float mag = isidx ? in.m.y : in.m.x;
            x = __fadd_rn(llr_e, __uint_as_float(__float_as_uint(mag) ^ sbit));  // posterior of the extension variable (:126)
            if constexpr (ET && kEtStoresExt<C>) {
                const uint32_t hb = __ballot_sync(0xffffffffu, x < 0.f);
                if ((th.r & 31) == 0) reinterpret_cast<uint32_t *>(th.smem + C::off_ext)[(I - 4) * C::tiles + th.tile] = hb;
            }
            q[K] = __fsub_rn(__uint_as_float(__float_as_uint(x) ^ sbit), mag);
        } else {
            x = in.x[K];
            q[K] = sub_record_mag<C, I, K>(th, __uint_as_float(__float_as_uint(x) ^ sbit), in.m, in.bits);
        }
        if constexpr (ET) synd ^= __float_as_uint(x);
        if constexpr (K == 0) {
            m1 = q[0];
        } else if constexpr (K == 1) {
            m2 = fmaxf(fabsf(m1), fabsf(q[1]));
            m1 = min_xorsign_abs(m1, q[1]);
        } else {
            m2 = fminf(m2, fmaxf(fabsf(m1), fabsf(q[K])));
            m1 = min_xorsign_abs(m1, q[K]);
        }
        return;
    }
    if constexpr (!FIRST) {
#if NRLDPC_MASKREG
        const bool isidx = is_argmin<C, I, K>(th, in.bits);
#else
        const bool isidx = idxf == ((uint32_t)K << C::idx_shift(I));
#endif
        lr = record_lr(in.m, isidx, in.bits << (31 - (DEG - 1 - K)));
    }
    float x;
    if constexpr (EXT && K == DEG - 1) {
        x = __fadd_rn(llr_e, lr);  // posterior of the degree-1 extension variable (:126); -0.0 + 0 = +0.0
        if constexpr (ET && kEtStoresExt<C>) {
            const uint32_t hb = __ballot_sync(0xffffffffu, x < 0.f);
            if ((th.r & 31) == 0) reinterpret_cast<uint32_t *>(th.smem + C::off_ext)[(I - 4) * C::tiles + th.tile] = hb;
        }
    } else {
        x = in.x[K];
    }
    if constexpr (ET) synd ^= __float_as_uint(x);  // sign bit = hard decision LQ<0 (:107-108)
    q[K] = FIRST ? x : __fsub_rn(x, lr);            // Lq = LQ - Lr (:131); LQ is never -0.0, so LQ - 0 == LQ bit for bit
    // |m1| = first minimum, sign of m1 = running sign product, m2 = second minimum (the first two edges
    // spelled out: no arithmetic against the +inf start values)
    if constexpr (K == 0) {
        m1 = q[0];
    } else if constexpr (K == 1) {
        m2 = fmaxf(fabsf(m1), fabsf(q[1]));
        m1 = min_xorsign_abs(m1, q[1]);
    } else {
        m2 = fminf(m2, fmaxf(fabsf(m1), fabsf(q[K])));
        m1 = min_xorsign_abs(m1, q[K]);
    }
}

template <class C, int I, bool ET, bool FIRST, int... K>
__device__ __forceinline__ void cn_edges_s(float (&q)[C::deg(I)], float &m1, float &m2, uint32_t &synd, const Th<C> &th,
                                           const RowIn<C, I> &in, uint32_t idxf, float llr_e,
                                           std::integer_sequence<int, K...>)
{
    (cn_edge_s<C, I, K, ET, FIRST>(q, m1, m2, synd, th, in, idxf, llr_e), ...);
}

// One check row of row-block I (compile time), lifted index r: syndrome bit of the current hard
// decisions, then the min-sum update of its record from Lq = LQ - Lr_old
// (py5gphy/ldpc/nr_ldpc_decode.py:107-123, :178-227).
template <class C, int I, bool ET, bool B0, bool FIRST = false>
__device__ __forceinline__ void cn_row_s(const DecArgs &a, const Th<C> &th, int *flag, const RowIn<C, I> &in,
                                         const float llr_e)
{
    constexpr int DEG = C::deg(I);
    char *rec = th.smem + C::mags_base(I) + th.r4;
    char *bp = th.smem + C::bits_base(I) + th.r4;
    const uint32_t idxf = in.bits & C::idx_mask(I);
    float q[DEG];
    float m1 = __uint_as_float(kInfBits), m2 = m1;
    uint32_t synd = 0;
    cn_edges_s<C, I, ET, FIRST>(q, m1, m2, synd, th, in, idxf, llr_e, std::make_integer_sequence<int, DEG>{});
    // signs of Lq and the argmin: an edge is "hot" when |Lq| == min1, i.e. min1 - |Lq| is +0 (never NaN
    // for finite inputs).  Edges that tie at min1 make min2 == min1, so any hot edge may carry the index.
    constexpr uint32_t ALL = (1u << DEG) - 1u;
#if NRLDPC_SIGN_FMA
    // Sign word on the FMA pipe (the check pass is bound by the ALU pipe): [Lq < 0] = sat(Lq * -inf)
    // exactly -- -inf/+inf saturate to 0/1 and 0 * inf = NaN saturates to +0 -- accumulated as an exact
    // small integer on top of 2^23, whose mantissa then holds the word.
    float saccf = 8388608.0f;
#pragma unroll
    for (int k = 0; k < DEG; ++k) {
        const float neg = __saturatef(__fmul_rn(q[k], __uint_as_float(0xff800000u)));
        saccf = __fmaf_rn(neg, (float)(1u << (DEG - 1 - k)), saccf);  // sign of Lq on edge k -> bit DEG-1-k
    }
    const uint32_t sacc = __float_as_uint(saccf);
#else
    uint32_t sacc = 0;
#pragma unroll
    for (int k = 0; k < DEG; ++k) sacc = push_top_bit(sacc, __float_as_uint(q[k]));  // sign of Lq on edge k -> bit DEG-1-k
#endif
    // The argmin: an edge is "hot" when |Lq| == min1, i.e. min1 - |Lq| is +0 (never NaN for finite inputs),
    // "cold" when it is negative.  Edges that tie at min1 make min2 == min1, so then ANY index gives the
    // same messages.
#if NRLDPC_COLD_FMA
    uint32_t kmin;
    if constexpr (DEG <= NRLDPC_COLD_FMA_MAXDEG) {
    // position of the hot edge = sum(all k) - sum(cold k), with cold = sat(d * -inf) in {0, 1} exactly and the
    // sum kept as an exact small integer on top of 2^23, whose low mantissa bits are then the position
    float hotpos = 8388608.0f + (float)(DEG * (DEG - 1) / 2);
#pragma unroll
    for (int k = 1; k < DEG; ++k) {
        const float c = __saturatef(__fmul_rn(__fsub_rn(fabsf(m1), fabsf(q[k])), __uint_as_float(0xff800000u)));
        hotpos = __fmaf_rn(c, -(float)k, hotpos);
    }
    kmin = __float_as_uint(hotpos);  // low bits; masked to the index field below
    } else {
        uint32_t cold = 0;
#pragma unroll
        for (int k = 0; k < DEG; ++k) cold = push_top_bit(cold, __float_as_uint(__fsub_rn(fabsf(m1), fabsf(q[k]))));
        kmin = (uint32_t)(DEG - 32) + (uint32_t)__clz(~cold & ALL);  // DEG-1 - (31 - clz): smallest hot k
    }
#else
    uint32_t cold = 0;
#pragma unroll
    for (int k = 0; k < DEG; ++k) cold = push_top_bit(cold, __float_as_uint(__fsub_rn(fabsf(m1), fabsf(q[k]))));
    const uint32_t hot = ~cold & ALL;
    const uint32_t kmin = (uint32_t)(DEG - 32) + (uint32_t)__clz(hot);  // DEG-1 - (31 - clz): smallest hot k
#endif
    // :199-202  Lr = alpha * sign_prod * sign(Lq) * max(minv - beta, 0), minv = min2 on the argmin edge
    // (beta == 0: max(m - 0, 0) == m for m >= +0)
    const float mag1 = B0 ? __fmul_rn(a.alpha, fabsf(m1)) : __fmul_rn(a.alpha, fmaxf(__fsub_rn(fabsf(m1), a.beta), 0.f));
    const float mag2 = B0 ? __fmul_rn(a.alpha, m2) : __fmul_rn(a.alpha, fmaxf(__fsub_rn(m2, a.beta), 0.f));
    const uint32_t sp = (uint32_t)((int)__float_as_uint(m1) >> 31);  // all ones when the sign product is -
    // (sign on the posterior: sacc holds sign(Lq') = s ^ sign(Lq), and XOR_k s_k = 0, so the row's sign product is sp as it is)
    const uint32_t nb = (xsign_row<C, I, FIRST>() ? ((sacc ^ in.bits ^ sp) & ALL) : ((sacc ^ sp) & ALL))
                      | ((kmin << C::idx_shift(I)) & C::idx_mask(I));
    *reinterpret_cast<float *>(rec) = mag1;
    *reinterpret_cast<float *>(rec + C::mag2_dist) = mag2;
    store_bits<C, I>(bp, nb);
    if (th.tile == 0) {  // mirror of the first 32 records behind the array (rotated reads of the variable pass)
#if NRLDPC_MIRROR_LOOP
        int n = th.one;
#pragma unroll 1
        do {  // (volatile: loop-invariant stores would be sunk out of the loop and the loop deleted)
            *reinterpret_cast<volatile float *>(rec + C::ZC * 4) = mag1;
            *reinterpret_cast<volatile float *>(rec + C::mag2_dist + C::ZC * 4) = mag2;
            if constexpr (C::kind(I) == 2) *reinterpret_cast<volatile uint32_t *>(bp + C::ZC * 4) = nb;
            else if constexpr (C::kind(I) == 1) *reinterpret_cast<volatile uint16_t *>(bp + C::ZC * 4) = (uint16_t)nb;
            else *reinterpret_cast<volatile uint8_t *>(bp + C::ZC * 4) = (uint8_t)nb;
        } while (--n > 0);
    }
    if (false) {
#endif
        *reinterpret_cast<float *>(rec + C::ZC * 4) = mag1;
        *reinterpret_cast<float *>(rec + C::mag2_dist + C::ZC * 4) = mag2;
        store_bits<C, I>(bp + C::ZC * 4, nb);
    }
    if (ET && (synd >> 31)) flag[0] = 1;
}

// A channel LLR: read-only path for the caller's llr[B,N]; L2-coherent load for a row this CTA recovered itself.
template <class C> __device__ __forceinline__ float ld_llr(const float *p)
{
    if constexpr (C::RR) return __ldcg(p);
    else return __ldg(p);
}

// Channel LLR of the degree-1 extension variable of row-block I (0 for the core rows), -0.0 -> +0.0.
// (-0.0 -> +0.0 only when the syndrome reads the sign bit of LLR + Lr (ET); the messages do not depend on
// the sign of a zero LLR: x = LLR + Lr and Lq = x - Lr give the same values either way.)
template <class C, int I, bool ET> __device__ __forceinline__ float load_ext_llr(const Th<C> &th)
{
    if constexpr (I >= 4) {
        const float v = ld_llr<C>(th.llr + (C::kb + I - 2) * C::ZC);
        return ET ? __fadd_rn(v, 0.0f) : v;
    } else return 0.f;
}

// The rows of warp group SUB, software-pipelined: the compiler cannot move a row's shared-memory loads
// above the previous row's record stores (it cannot prove they do not alias), so the next row's inputs
// are loaded, in program order, BEFORE the current row is computed and stored; its channel LLR (L2) too.
// Rows whose combined degree exceeds NRLDPC_PF_SUM load their inputs after the store instead (registers).
// named barrier (id 1 + SUB) among the warps of group SUB
template <class C, int SUB> __device__ __forceinline__ void group_sync()
{
    asm volatile("bar.sync %0, %1;" ::"n"(1 + SUB), "n"(C::tiles * 32) : "memory");
}

template <class C, bool ET, bool B0, bool FIRST, int SUB, int O = 0>
__device__ __forceinline__ void cn_pass_s(const DecArgs &a, const Th<C> &th, int *flag,
                                          const RowIn<C, kRows<C>.item[SUB][O]> &in_cur, const float llr_cur,
                                          const float llr_nxt_in = 0.f)
{
    constexpr int I = kRows<C>.item[SUB][O];
#if NRLDPC_GROUP_SYNC_CN > 0
    if constexpr (O > 0 && O % NRLDPC_GROUP_SYNC_CN == 0) group_sync<C, SUB>();
#endif
    if constexpr (O + 1 < kRows<C>.n[SUB]) {
        constexpr int In = kRows<C>.item[SUB][O + 1];
#if NRLDPC_PF_LLR2
        const float llr_nxt = llr_nxt_in;
        float llr_nn = 0.f;
        if constexpr (O + 2 < kRows<C>.n[SUB]) llr_nn = load_ext_llr<C, kRows<C>.item[SUB][O + 2], ET>(th);
#else
        const float llr_nxt = load_ext_llr<C, In, ET>(th), llr_nn = 0.f;
#endif
        if constexpr (C::deg(I) + C::deg(In) <= NRLDPC_PF_SUM) {
            const RowIn<C, In> in_nxt = load_row<C, In, FIRST>(th);
            cn_row_s<C, I, ET, B0, FIRST>(a, th, flag, in_cur, llr_cur);
            cn_pass_s<C, ET, B0, FIRST, SUB, O + 1>(a, th, flag, in_nxt, llr_nxt, llr_nn);
        } else {
            cn_row_s<C, I, ET, B0, FIRST>(a, th, flag, in_cur, llr_cur);
            const RowIn<C, In> in_nxt = load_row<C, In, FIRST>(th);
            cn_pass_s<C, ET, B0, FIRST, SUB, O + 1>(a, th, flag, in_nxt, llr_nxt, llr_nn);
        }
    } else {
        cn_row_s<C, I, ET, B0, FIRST>(a, th, flag, in_cur, llr_cur);
    }
}

template <class C, int J, int Nn>
__device__ __forceinline__ void vn_entry_s(float &acc, const Th<C> &th)
{
    constexpr int e = C::cedge(J, Nn), i = C::row_of(e), k = e - C::G::rowptr(i), DEGI = C::deg(i);
    constexpr int back = (C::ZC - C::P(e)) % C::ZC, t = C::wrap_tile(back);  // check r = (c - P) mod Zc
    const uint32_t bits = load_bits<C, i>(th.smem + (C::bits_base(i) + 4 * back) + th.r4 + th.w4[t]);
    const bool isidx = is_argmin<C, i, k>(th, bits);
    // only the magnitude this edge uses is read (one 128-byte wavefront): mag2 on the argmin edge, mag1 elsewhere
#if NRLDPC_VN_PSEL
    const float mag = *reinterpret_cast<const float *>((isidx ? th.p4m2 : th.p4) + ((C::mags_base(i) + 4 * back) + th.w4[t]));
#else
    const float mag = *reinterpret_cast<const float *>(th.smem + (C::mags_base(i) + 4 * back) + (isidx ? th.r4m2 : th.r4) + th.w4[t]);
#endif
    acc = __fadd_rn(acc, __uint_as_float(__float_as_uint(mag) ^ ((bits << (31 - (DEGI - 1 - k))) & 0x80000000u)));
}

template <class C, int J, int... Nn>
__device__ __forceinline__ void vn_entries_s(float &acc, const Th<C> &th, std::integer_sequence<int, Nn...>)
{
    (vn_entry_s<C, J, Nn>(acc, th), ...);
}

// One core variable of column-block J (compile time), lifted index c = r: LQ = LLRin + sum_i Lr(i) in
// ascending row-block order = ascending check index (py5gphy/ldpc/nr_ldpc_decode.py:126).
template <class C, int J>
__device__ __forceinline__ float vn_col_s(const Th<C> &th_, const float lv)
{
    const Th<C> th = opaque(th_);
    float acc = 0.f;
    vn_entries_s<C, J>(acc, th, std::make_integer_sequence<int, C::cdeg(J)>{});
    return __fadd_rn(lv, acc);
}
template <class C, int J>
__device__ __forceinline__ void vn_store_s(const Th<C> &th, const float lq)
{
    float *dst = reinterpret_cast<float *>(th.smem + C::lq_base(J) + th.r4);
    *dst = lq;
    if (th.tile == 0) dst[C::ZC] = lq;  // mirror for the rotated reads of the check pass
}

// Channel LLR of core column-block J; the 2Zc punctured systematic bits start at LLR 0 (:43).
// (no -0.0 -> +0.0 here: the sum of the messages is never -0.0, so LLR + sum has the same value and sign.)
template <class C, int J> __device__ __forceinline__ float load_col_llr(const Th<C> &th)
{
    if constexpr (J >= 2) return ld_llr<C>(th.llr + (J - 2) * C::ZC);
    else return 0.f;
}

// The columns of warp group SUB.  A column's posterior is stored after the NEXT column's records have
// been read (same aliasing argument as in cn_pass_s), its channel LLR is fetched one column ahead.
template <class C, int SUB, int O = 0>
__device__ __forceinline__ void vn_pass_s(const Th<C> &th, const float lv_cur, const float lq_prev, const float lv_nxt_in = 0.f)
{
    if constexpr (O < kCols<C>.n[SUB]) {
#if NRLDPC_GROUP_SYNC_VN > 0
        if constexpr (O > 0 && O % NRLDPC_GROUP_SYNC_VN == 0) group_sync<C, SUB>();
#endif
#if NRLDPC_PF_VLLR == 2
        const float lv_nxt = lv_nxt_in;
        float lv_nn = 0.f;
        if constexpr (O + 2 < kCols<C>.n[SUB]) lv_nn = load_col_llr<C, kCols<C>.item[SUB][O + 2]>(th);
#else
        float lv_nxt = 0.f, lv_nn = 0.f;
        if constexpr (O + 1 < kCols<C>.n[SUB]) lv_nxt = load_col_llr<C, kCols<C>.item[SUB][O + 1]>(th);
#endif
        const float lq = vn_col_s<C, kCols<C>.item[SUB][O]>(th, lv_cur);
        if constexpr (O > 0) vn_store_s<C, kCols<C>.item[SUB][O - 1]>(th, lq_prev);
        vn_pass_s<C, SUB, O + 1>(th, lv_nxt, lq, lv_nn);
    } else if constexpr (O > 0) {
        vn_store_s<C, kCols<C>.item[SUB][O - 1]>(th, lq_prev);
    }
}

// Final syndrome of row-block I with the post-loop tie rule LQ<=0 -> 1 (:134-143) + extension hard bits.
template <class C, int I, int... K>
__device__ __forceinline__ void final_edges_s(uint32_t &synd, const Th<C> &th, std::integer_sequence<int, K...>)
{
    ((synd ^= (load_lq_rot<C, C::G::rowptr(I) + K>(th) <= 0.f) ? 1u : 0u), ...);
}
// CORE = false: only the hard bit of the extension variable (the packed final syndrome does the rest).
template <class C, int I, bool CORE = true>
__device__ __forceinline__ void final_row_s(const Th<C> &th, int *flag, const float llr_e)
{
    constexpr int DEG = C::deg(I), NCORE = (I >= 4) ? DEG - 1 : DEG;
    uint32_t synd = 0;
    if constexpr (CORE) final_edges_s<C, I>(synd, th, std::make_integer_sequence<int, NCORE>{});
    if constexpr (I >= 4) {
        float2 m;
        m.x = *reinterpret_cast<const float *>(th.smem + C::mags_base(I) + th.r4);
        m.y = *reinterpret_cast<const float *>(th.smem + (C::mags_base(I) + C::mag2_dist) + th.r4);
        const uint32_t bits = load_bits<C, I>(th.smem + C::bits_base(I) + th.r4);
        const bool isidx = ((bits ^ ((uint32_t)(DEG - 1) << C::idx_shift(I))) & C::idx_mask(I)) == 0;
        const float lr = record_lr(m, isidx, bits << 31);
        const float x = __fadd_rn(llr_e, lr);
        const bool hb1 = x <= 0.f;
        const uint32_t hb = __ballot_sync(0xffffffffu, hb1);
        if ((th.r & 31) == 0) reinterpret_cast<uint32_t *>(th.smem + C::off_ext)[(I - 4) * C::tiles + th.tile] = hb;
        synd ^= hb1 ? 1u : 0u;
    }
    if (CORE && synd) flag[0] = 1;
    asm volatile("" ::: "memory");  // rows stay in program order (register pressure)
}
// The channel LLRs of all extension variables of the group are requested up front: one exposed L2 latency
// per final pass instead of one per row.
template <class C, int SUB, bool CORE, int... O>
__device__ __forceinline__ void final_pass_s(const Th<C> &th, int *flag, std::integer_sequence<int, O...>)
{
    const float llr_e[] = {load_ext_llr<C, kRows<C>.item[SUB][O], true>(th)...};  // -0.0 -> +0.0
    (final_row_s<C, kRows<C>.item[SUB][O], CORE>(th, flag, llr_e[O]), ...);
}

// ---- Final syndrome on bit-packed hard decisions (:134-143), for Zc a multiple of 32.
// The hard decisions LQ <= 0 of the core column-blocks are ballot-packed into tiles words per column-block
// (stored twice in a row, so that word index w + q needs no modulo); the syndrome word w of row-block i is then
// the XOR over its core edges of bits [32 w + P, 32 w + P + 32) of the column's Zc-bit vector, i.e. a funnel shift
// of two neighbouring words, XOR the packed hard bits of its extension column: Zc/32 lanes per row-block
// instead of Zc threads, 2 shared-memory words per edge and syndrome word instead of 32.
template <class C> struct FinalTab {
    uint16_t start[C::nrows + 1];
    uint16_t e[C::nnz];  // core edges only: col | (P / 32) << 5 | (P % 32) << 9
};
template <class C> constexpr FinalTab<C> make_final_tab()
{
    FinalTab<C> t{};
    int n = 0;
    for (int i = 0; i < C::nrows; ++i) {
        t.start[i] = (uint16_t)n;
        const int ncore = (i >= 4) ? C::deg(i) - 1 : C::deg(i);
        for (int k = 0; k < ncore; ++k) {
            const int e = C::G::rowptr(i) + k;
            t.e[n++] = (uint16_t)(C::G::col(e) | ((C::P(e) / 32) << 5) | ((C::P(e) % 32) << 9));
        }
    }
    t.start[C::nrows] = (uint16_t)n;
    return t;
}
template <class C> __constant__ FinalTab<C> kFinalTab = make_final_tab<C>();


// step 1 (all warps, after the last variable pass): pack the hard decisions of this warp group's columns
// (LT: the in-loop tie rule LQ < 0 -> 1 of a codeblock that leaves early, :107-108, instead of the final LQ <= 0 -> 1)
template <class C, bool LT = false> __device__ __forceinline__ void final_pack(int sub, const Th<C> &th, uint32_t *hbw)
{
    const int lane = th.r & 31;
    for (int j = sub; j < C::ncore; j += C::S) {
        const float x = *reinterpret_cast<const float *>(th.smem + C::lq_base(j) + th.r4);
        const uint32_t word = __ballot_sync(0xffffffffu, LT ? (x < 0.f) : (x <= 0.f));
        if (lane == 0) hbw[j * 2 * C::tiles + th.tile] = hbw[j * 2 * C::tiles + C::tiles + th.tile] = word;
    }
}
// step 2 (after a barrier): warp `warp` checks row-blocks warp, warp + nwarps, ...; lane w < tiles owns syndrome word w
template <class C> __device__ __forceinline__ void final_syndrome(int warp, int lane, const uint32_t *hbw, const uint32_t *ext, int *flag)
{
    for (int i = warp; i < C::nrows; i += C::nwarps) {
        const int e0 = kFinalTab<C>.start[i], e1 = kFinalTab<C>.start[i + 1];
        if (lane < C::tiles) {
            uint32_t acc = (i >= 4) ? ext[(i - 4) * C::tiles + lane] : 0u;
#pragma unroll 4
            for (int e = e0; e < e1; ++e) {
                const uint32_t t = kFinalTab<C>.e[e];
                const uint32_t *v = hbw + (t & 31u) * (2 * C::tiles) + ((t >> 5) & 15u) + lane;
                acc ^= __funnelshift_r(v[0], v[1], t >> 9);
            }
            if (acc) flag[0] = 1;
        }
    }
}

// Early exit with all parity checks satisfied: hard-bit word w of the extension variable of row-block i = XOR over the row-block's
// core edges of the rotated core words (same walk as final_syndrome); warp `warp` takes row-blocks 4 + warp, 4 + warp + nwarps, ...
template <class C> __device__ __forceinline__ void ext_from_core(int warp, int lane, const uint32_t *hbw, uint32_t *ext)
{
    for (int i = 4 + warp; i < C::nrows; i += C::nwarps) {
        const int e0 = kFinalTab<C>.start[i], e1 = kFinalTab<C>.start[i + 1];
        if (lane < C::tiles) {
            uint32_t acc = 0u;
#pragma unroll 4
            for (int e = e0; e < e1; ++e) {
                const uint32_t t = kFinalTab<C>.e[e];
                const uint32_t *v = hbw + (t & 31u) * (2 * C::tiles) + ((t >> 5) & 15u) + lane;
                acc ^= __funnelshift_r(v[0], v[1], t >> 9);
            }
            ext[(i - 4) * C::tiles + lane] = acc;
        }
    }
}

// run PASS<SUB> for the warp's group (sub is warp-uniform)
template <class C, bool ET, bool B0, bool FIRST, int SUB = 0>
__device__ __forceinline__ void run_cn(int sub, const DecArgs &a, const Th<C> &th, int *flag)
{
    if constexpr (SUB < C::S) {
        if (sub == SUB) cn_pass_s<C, ET, B0, FIRST, SUB>(a, th, flag, load_row<C, kRows<C>.item[SUB][0], FIRST>(th), load_ext_llr<C, kRows<C>.item[SUB][0], ET>(th),
                                                        load_ext_llr<C, kRows<C>.item[SUB][kRows<C>.n[SUB] > 1 ? 1 : 0], ET>(th));
        else run_cn<C, ET, B0, FIRST, SUB + 1>(sub, a, th, flag);
    }
}
template <class C, int SUB = 0>
__device__ __forceinline__ void run_vn(int sub, const Th<C> &th)
{
    if constexpr (SUB < C::S) {
        if (sub == SUB) vn_pass_s<C, SUB>(th, load_col_llr<C, kCols<C>.item[SUB][0]>(th), 0.f,
                                          load_col_llr<C, kCols<C>.item[SUB][kCols<C>.n[SUB] > 1 ? 1 : 0]>(th));
        else run_vn<C, SUB + 1>(sub, th);
    }
}
template <class C, bool CORE = true, int SUB = 0>
__device__ __forceinline__ void run_final(int sub, const Th<C> &th, int *flag)
{
    if constexpr (SUB < C::S) {
        if (sub == SUB) final_pass_s<C, SUB, CORE>(th, flag, std::make_integer_sequence<int, kRows<C>.n[SUB]>{});
        else run_final<C, CORE, SUB + 1>(sub, th, flag);
    }
}

// Pull `bytes` (a multiple of 16) at a 16-byte aligned global address into L2, asynchronously.
__device__ __forceinline__ void prefetch_l2(const void *p, uint32_t bytes)
{
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}

// Pull what codeblock cb will read into L2: its row of llr[B,N], or (fused rate recovery) its stretch of the received
// sequence.  Thread 0 only.
// (the plain-LLR path keeps the exact statements of the round-1 kernel: its SASS -- and its 0.7 % -- depend on them)
template <class C> __device__ __forceinline__ void prefetch_cb(const DecArgs &a, int cb)
{
    if constexpr (C::RR) {
        const size_t esz = a.rr.in_f64 ? 8 : 4;
        const uintptr_t p0 = reinterpret_cast<uintptr_t>(a.rr.src) + (size_t)a.rr.goff[cb] * esz;
        const uintptr_t lo = p0 & ~(uintptr_t)15, hi = (p0 + (size_t)a.rr.E[cb] * esz) & ~(uintptr_t)15;
        if (hi > lo) prefetch_l2(reinterpret_cast<const void *>(lo), (uint32_t)(hi - lo));
    } else {
        prefetch_l2(a.llr + (size_t)cb * C::N, C::N * 4);
    }
}

int num_sms()
{
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    }
    return n;
}

template <class C, bool ET, bool B0>
__global__ void __launch_bounds__(C::nwarps * 32, C::ctas)
decode_spec_kernel(const __grid_constant__ DecArgs a)
{
    extern __shared__ __align__(16) char smem[];
    __shared__ int s_flag[2];
    __shared__ int s_next;
    constexpr int NT = C::nwarps * 32, ZC = C::ZC;
    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __reduce_min_sync(0xffffffffu, tid >> 5);  // warp-uniform (REDUX writes a uniform register)
#if NRLDPC_SUB_INTERLEAVE
    // warp -> (tile, group) with the group in the low bits: the warps of one scheduler (warp % 4) then all
    // run the same instruction stream (S = 2: even schedulers group 0, odd schedulers group 1)
    const int sub = warp % C::S;
    Th<C> th;
    th.smem = smem;
    th.tile = warp / C::S;
#else
    const int sub = warp / C::tiles;
    Th<C> th;
    th.smem = smem;
    th.tile = warp % C::tiles;
#if NRLDPC_TILE_ROT
    th.tile = (th.tile + sub) % C::tiles;
#endif
#endif
    th.r = min(th.tile * 32 + lane, ZC - 1);
    th.r4 = (uint32_t)th.r * 4u;
    th.r4m2 = th.r4 + (uint32_t)C::mag2_dist;
    th.p4 = smem + th.r4;
    th.p4m2 = smem + th.r4m2;
#pragma unroll
    for (int t = 0; t <= C::tiles; ++t) th.w4[t] = (th.tile >= t) ? -4 * ZC : 0;
    {   // masks ptxas cannot fold back into immediates: B > 0 in every launch, so `zero` is 0 -- at run time only
        const uint32_t zero = (uint32_t)a.B >> 31;
        th.mk[0] = 0xe0u + zero, th.mk[1] = 0xf000u + zero, th.mk[2] = 0x1f000000u + zero;
        asm volatile("add.s32 %0, %1, 1;" : "=r"(th.one) : "r"(zero));  // 1, to neither nvvm nor ptxas
    }

    // Persistent CTA (one per SM: the codeblock state fills shared memory): codeblocks blockIdx.x,
    // blockIdx.x + gridDim.x, ...  The channel LLRs are read straight from global memory every iteration
    // (L2 hits); the NEXT codeblock's row is pulled into L2 while the current one is decoded, so that no
    // HBM latency is exposed at a codeblock boundary.
    const bool pf = tid == 0 && (C::RR || (reinterpret_cast<uintptr_t>(a.llr) & 15) == 0);  // bulk prefetch needs 16-byte alignment
    if (pf) prefetch_cb<C>(a, blockIdx.x);
    // With early termination the codeblocks take different numbers of iterations, so the CTAs draw them from a
    // ticket counter (a.work) instead of a fixed stride: the first one is blockIdx.x, every further one
    // gridDim.x + ticket.  Thread 0 takes the ticket of the NEXT codeblock at the top (its latency hides behind the
    // state initialisation), prefetches that codeblock's LLRs and publishes the index before the first barrier.
    const bool dyn = ET && a.work != nullptr;
    for (int cb = blockIdx.x; cb < a.B; cb += gridDim.x) {
        int nx = 0;
        if (ET) { if (dyn && tid == 0) nx = (int)gridDim.x + (int)(atomicAdd(a.work, 1u) - a.work_base); }
        if constexpr (C::RR) {
            // the LLR load is the rate recovery (+ HARQ combining) of this codeblock: received sequence -> this CTA's fp32
            // row (L2-resident, rewritten for every codeblock) and the float64 soft buffer the caller keeps.  The codeblock
            // state in shared memory is dead here, its first bytes serve the block reduction.
            float *row = a.rr.scratch + (size_t)blockIdx.x * C::N;
            rr_row(a.rr, cb, C::N, row, reinterpret_cast<double *>(smem));
            __syncthreads();  // the row is read back by other threads, the reduction scratch is re-initialised below
            th.llr = row + th.r;
        } else {
            th.llr = a.llr + (size_t)cb * C::N + th.r;
#if NRLDPC_LLR_OPAQUE
            asm volatile("" : "+l"(th.llr));  // keep the row pointer in a register pair (otherwise re-derived from cb and r at every LDG)
#endif
        }
#ifndef NRLDPC_EXP_NO_PF
        if (!dyn) {
            if (pf && cb + (int)gridDim.x < a.B) prefetch_cb<C>(a, cb + (int)gridDim.x);
        }

#endif

        // ---- init: records = 0 (Lr = 0, :101), LQ = LLRin (:94) with the punctured columns at 0 (:43)
        {
            constexpr int z0 = C::off_mags / 16, z1 = C::smem_bytes / 16;
            static_assert(C::off_mags % 16 == 0 && C::smem_bytes % 16 == 0, "16-byte zero fill");
            // (with the specialised first check pass every record is written before it is read, unless no pass runs)
            if (!(NRLDPC_FIRST_IT > 1 || (NRLDPC_FIRST_IT && !ET)) || a.max_iter <= 0)
                for (int t = z0 + tid; t < z1; t += NT) reinterpret_cast<uint4 *>(smem)[t] = make_uint4(0, 0, 0, 0);
            for (int t = tid; t < 2 * C::LQS; t += NT) reinterpret_cast<float *>(smem + C::lq_base(0))[t] = 0.f;
            if (tid < 2) s_flag[tid] = 0;
            for (int j = 2 + sub; j < C::ncore; j += C::S) {
#ifdef NRLDPC_EXP_NO_INIT_LLR
                const float v = 1.0f;
#else
                const float v = __fadd_rn(ld_llr<C>(th.llr + (j - 2) * ZC), 0.0f);  // -0.0 -> +0.0
#endif
                float *dst = reinterpret_cast<float *>(smem + C::lq_base(j) + th.r4);
                *dst = v;
                if (th.tile == 0) dst[ZC] = v;
            }
        }
        if (ET) {
            if (dyn && tid == 0) {
                s_next = nx;
                if (pf && nx < a.B) prefetch_cb<C>(a, nx);
            }
        }
        __syncthreads();
        if (ET) { if (dyn) nx = s_next; }

        bool et_done = false;
        int it = 0;
        for (; it < a.max_iter; ++it) {
            int *flag = &s_flag[it & 1];
#ifndef NRLDPC_EXP_NO_CN
#if NRLDPC_FIRST_IT
            if ((NRLDPC_FIRST_IT > 1 || !ET) && it == 0) run_cn<C, ET, B0, true>(sub, a, th, flag);
            else run_cn<C, ET, B0, false>(sub, a, th, flag);
#else
            run_cn<C, ET, B0, false>(sub, a, th, flag);
#endif
#endif
            __syncthreads();
            if (ET) {
                if (*flag == 0) { et_done = true; break; }  // all parity checks hold for the current decisions (:111-114)
                if (tid == 0) s_flag[(it + 1) & 1] = 0;
            }
#ifndef NRLDPC_EXP_NO_VN
            run_vn<C>(sub, th);
#endif
            __syncthreads();
        }
        bool ok = et_done;
#ifndef NRLDPC_EXP_NO_FINAL
        if (!et_done) {
            int *flag = &s_flag[it & 1];
            if constexpr (kFinalPacked<C>) {
                // the records of the core row-blocks 0-3 are dead after the last variable pass: their mag1 rows hold the packed words
                uint32_t *hbw = reinterpret_cast<uint32_t *>(smem + C::mags_base(0));
                final_pack<C>(sub, th, hbw);
                run_final<C, false>(sub, th, flag);  // hard bits of the extension columns
                __syncthreads();
                final_syndrome<C>(warp, lane, hbw, reinterpret_cast<const uint32_t *>(smem + C::off_ext), flag);
            } else {
                run_final<C>(sub, th, flag);
            }
            __syncthreads();
            ok = (*flag == 0);
        }
#endif

        // ---- outputs
        if (tid == 0) {
            if (a.status) a.status[cb] = ok ? 1 : 0;
            if (a.iters) a.iters[cb] = it;
        }
        const uint32_t *ext = reinterpret_cast<const uint32_t *>(smem + C::off_ext);
        if constexpr (ET && !kEtStoresExt<C>) {
            if (et_done && a.ck) {  // (block-uniform) the extension bits of a converged codeblock, from its core decisions
                uint32_t *hbw = reinterpret_cast<uint32_t *>(smem + C::mags_base(0));  // the records are dead
                final_pack<C, true>(sub, th, hbw);
                __syncthreads();
                ext_from_core<C>(warp, lane, hbw, reinterpret_cast<uint32_t *>(smem + C::off_ext));
                __syncthreads();
            }
        }
        if (a.ck) {
            int8_t *out = a.ck + (size_t)cb * C::Nfull;
            for (int j = 0; j < C::ncore; ++j) {
                const float *LQ = reinterpret_cast<const float *>(smem + C::lq_base(j));
                for (int c = tid; c < ZC; c += NT) out[j * ZC + c] = (int8_t)(et_done ? (LQ[c] < 0.f) : (LQ[c] <= 0.f));
            }
            for (int n = tid; n < (C::nrows - 4) * ZC; n += NT) {
                const int i4 = n / ZC, r = n - i4 * ZC;
                out[C::ncore * ZC + n] = (int8_t)((ext[i4 * C::tiles + (r >> 5)] >> (r & 31)) & 1u);
            }
        }
#ifndef NRLDPC_EXP_NO_OUT
        if (kFinalPacked<C> && a.info && !et_done) {
            // the packed hard decisions of the final syndrome are the info words (K = kb Zc, Zc a multiple of 32)
            constexpr int nwords = C::kb * C::tiles;
            const uint32_t *hbw = reinterpret_cast<const uint32_t *>(smem + C::mags_base(0));
            for (int n = tid; n < nwords; n += NT) a.info[(size_t)cb * nwords + n] = hbw[(n / C::tiles) * 2 * C::tiles + n % C::tiles];
        } else if (a.info) {
            constexpr int nwords = (C::K + 31) / 32;
            for (int w = (tid >> 5); w < nwords; w += C::nwarps) {
                const int n = min(32 * w + lane, C::K - 1), j = n / ZC, c = n - j * ZC;
                const float x = *reinterpret_cast<const float *>(smem + C::lq_base(j) + 4 * c);
                uint32_t word = __ballot_sync(0xffffffffu, et_done ? (x < 0.f) : (x <= 0.f));
                if (32 * w + 32 > C::K) word &= (1u << (C::K - 32 * w)) - 1u;   // last, partial word
                if (lane == 0) a.info[(size_t)cb * nwords + w] = word;
            }
        }
#endif
        __syncthreads();  // the state is re-initialised for the next codeblock
        if (ET) { if (dyn) cb = nx - (int)gridDim.x; }  // the loop adds the stride back
    }
}

template <class C>
int launch_spec(const DecArgs &a, int early_term, cudaStream_t s)
{
    static_assert(C::smem_bytes <= 227 * 1024 - 64, "codeblock state does not fit in shared memory");
    auto launch = [&](auto kern) -> int {
        NRLDPC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::smem_bytes));
        NRLDPC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        if (getenv("NRLDPC_DEBUG")) {
            int nb = 0;
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, C::nwarps * 32, C::smem_bytes);
            fprintf(stderr, "nrldpc: decode_spec<%d,%d> threads=%d smem=%d ctas/SM wanted=%d resident=%d\n", C::bgn, C::ZC,
                    C::nwarps * 32, C::smem_bytes, C::ctas, nb);
        }
        const int grid = std::min(a.B, C::ctas * num_sms());
        DecArgs b = a;
        // dynamic codeblock queue of the early-termination kernels, once every CTA decodes several codeblocks: a ticket
        // counter from a per-device ring (no allocation or memset on the launch path; NRLDPC_STATIC_QUEUE=1 or a
        // capturing stream: fixed stride)
        static const bool static_queue = getenv("NRLDPC_STATIC_QUEUE") != nullptr;
        if (early_term && a.B >= 2 * grid && !static_queue) {
            if (int rc = decode_queue_slot(s, (unsigned)a.B, &b.work, &b.work_base)) return rc;
        }
        ScratchBuf rows;  // fused rate recovery: one fp32 row per persistent CTA, stream-ordered
        if (b.rr.src && !b.rr.scratch) {
            NRLDPC_CUDA(rows.alloc((size_t)grid * C::N * sizeof(float), s));
            b.rr.scratch = rows.as<float>();
        }
        kern<<<grid, C::nwarps * 32, C::smem_bytes, s>>>(b);
        NRLDPC_CUDA(cudaGetLastError());
        return NRLDPC_OK;
    };
    // the lifting sizes below 144 are instantiated for the reference's semantics only (early termination; beta = 0 and
    // beta as a run-time value, plus the fused rate recovery): fixed-iteration runs go to the table-driven kernel there.
    // This keeps the library's build time and size in check (6 kernels of ~170 KB of SASS per (bgn, Zc) otherwise).
    constexpr bool lean = C::ZC < 144;
    if (a.rr.src) {
        using CR = Code<C::bgn, C::ZC, true>;
        if (!early_term) { set_error("decode: the fused rate recovery runs with early termination only"); return NRLDPC_EINVAL; }
        if constexpr (!lean) if (a.beta == 0.f) return launch(decode_spec_kernel<CR, true, true>);
        return launch(decode_spec_kernel<CR, true, false>);
    }
    if constexpr (lean) {
        if (!early_term) return kNoVariant;
        return a.beta == 0.f ? launch(decode_spec_kernel<C, true, true>) : launch(decode_spec_kernel<C, true, false>);
    } else {
        if (a.beta == 0.f) return early_term ? launch(decode_spec_kernel<C, true, true>) : launch(decode_spec_kernel<C, false, true>);
        return early_term ? launch(decode_spec_kernel<C, true, false>) : launch(decode_spec_kernel<C, false, false>);
    }
}

}  // namespace


// One instantiation per translation unit.
#define NRLDPC_SPEC_INSTANCE(BGN, ZC)                                                                              \
    int launch_decode_spec_##BGN##_##ZC(const DecArgs &a, int early_term, cudaStream_t s)                          \
    {                                                                                                              \
        return launch_spec<Code<BGN, ZC>>(a, early_term, s);                                                       \
    }                                                                                                              \
    void decode_spec_geometry_##BGN##_##ZC(int *threads, int *smem)                                                \
    {                                                                                                              \
        if (threads) *threads = Code<BGN, ZC>::nwarps * 32;                                                        \
        if (smem) *smem = Code<BGN, ZC>::smem_bytes;                                                               \
    }

}  // namespace nrldpc
