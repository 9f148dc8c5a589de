// nrldpc_encode.cu -- batched 5G NR LDPC encoder for sm_100a: bit-packed GF(2) rotate/XOR.
//
// Replaces nr_ldpc_encode.encode_ldpc / _gen_ldpc_parity_bit (py5gphy/ldpc/nr_ldpc_encode.py:8-115),
// which multiplies by dense int8 circulant blocks.  Here every Zc-bit column-block of the codeword is
// a row of ceil(Zc/32) 32-bit words in shared memory (packed with warp ballots); multiplying by a
// circulant block with shift P is a bit rotation (one funnel shift per output word) and the GF(2)
// sums are XORs.  The core parity follows the same A/B/C block split as the reference (:88-113):
//   L1 = A ck (4 row-blocks), L2 = sum L1, pc1 = roll(L2, s), pc2/pc4 from pc1, pc3 from pc4 (BG1) or
//   pc2 (BG2), extension parity pe = C [ck; pc].
#include "nrldpc_common.cuh"
#include "nrldpc_bits.cuh"

namespace nrldpc {

namespace {

constexpr int kEncThreads = 256;

// Rate matching fused into the encoder's store ("bit-select + Qm interleave into the encoder's store"):
// nr_ldpc_ratematch.ratematch_ldpc (py5gphy/ldpc/nr_ldpc_ratematch.py:64-97) + code block concatenation
// (py5gphy/nr_pdsch/nr_dlsch.py:66-68) in gather form, straight from the packed codeword in shared memory -- dn never
// exists.  Output byte o of codeblock cb is bit k = (o mod Qm) * E/Qm + o div Qm of the selection (:90-93), i.e. the
// (k mod S)-th non-filler position of the walk that starts at k0 on the circular buffer of length Ncb (:80-87; S = the
// number of non-filler positions, the fillers are the buffer positions [F0, F1)).
template <class BitAt>
__device__ __forceinline__ void rm_store(const EncRmArgs &rm, int cb, BitAt bit_at, int tid, int nthreads)
{
    const int E = rm.E[cb], Qm = rm.Qm, Ncb = rm.Ncb, k0 = rm.k0;
    int8_t *out = rm.g + rm.goff[cb];
    const int f0 = min(max(rm.F0, 0), Ncb), f1 = min(max(rm.F1, f0), Ncb), nf = f1 - f0, S = Ncb - nf;
    if (S <= 0 || E <= 0) return;
    // rank r of a non-filler position -> its step t of the walk: the fillers form one stretch [a, a + nf) of the walk, or
    // (k0 inside the fillers) its first `lead` steps and its tail
    int a, skip, lead = 0;
    if (k0 <= f0) { a = f0 - k0; skip = nf; }
    else if (k0 >= f1) { a = f0 - k0 + Ncb; skip = nf; }
    else { a = Ncb; skip = 0; lead = f1 - k0; }
    const int cols = E / Qm;
    for (int o = tid; o < E; o += nthreads) {
        const int e = o / Qm, q = o - e * Qm;
        int r = q * cols + e;
        if (r >= S) r -= (r / S) * S;  // repetition when E exceeds the buffer
        int t = r + lead;
        if (t >= a) t += skip;
        int pos = k0 + t;
        if (pos >= Ncb) pos -= Ncb;
        out[o] = (int8_t)bit_at(pos);
    }
}

__device__ __forceinline__ int find_edge(const QcCfg &c, int i, int j)
{
    for (int e = c.rowptr[i]; e < c.rowptr[i + 1]; ++e)
        if ((int)(c.edge[e] & 0xff) == j) return (int)(c.edge[e] >> 8);
    return 0;
}

__global__ void __launch_bounds__(kEncThreads)
encode_kernel(const __grid_constant__ QcCfg c, int8_t *__restrict__ ck, int B, int G, int fix_fillers,
              int8_t *__restrict__ dn, int vec, const __grid_constant__ EncRmArgs rm)
{
    extern __shared__ uint32_t smem[];
    const int Zc = c.Zc, W = c.tiles, Wp = W + 1;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = kEncThreads / 32;
    const int cb0 = blockIdx.x * G;
    const int g_cnt = min(G, B - cb0);
    // per codeblock: v[ncols][Wp] codeword column-blocks, fm[kb][W] filler masks, l1[4][Wp], l2[Wp]
    const int slot = c.ncols * Wp + c.kb * W + 5 * Wp;
    auto V = [&](int g, int j) { return smem + g * slot + j * Wp; };
    auto FM = [&](int g, int j) { return smem + g * slot + c.ncols * Wp + j * W; };
    auto L1 = [&](int g, int i) { return smem + g * slot + c.ncols * Wp + c.kb * W + i * Wp; };

    for (int t = threadIdx.x; t < g_cnt * slot; t += kEncThreads) smem[t] = 0;
    __syncthreads();

    // A. pack the K input bits.  Zc % 16 == 0 and 16-byte aligned rows (vec): one thread per 16 input bytes, one
    // 128-bit load -> 16 codeword bits + 16 filler-mask bits written as a half word; otherwise one warp per
    // (codeblock, column-block, word) with a ballot.
    const int H = Zc >> 4;  // 16-bit chunks per column-block (vec)
    // vec thread mapping: thread = (row tj, chunk h) of the [rows][H] array of 16-byte chunks, RYv rows per sweep;
    // rows of consecutive codeblocks are contiguous in memory, so a sweep reads one contiguous range and the only
    // divisions are by the compile-time kb / (ncols - 2) of the base graph
    const int tj = vec ? (int)threadIdx.x / H : 0, h = (int)threadIdx.x - tj * H, RYv = vec ? kEncThreads / H : 1;
    if (vec) {
        for (int R = tj; R < g_cnt * c.kb && tj < RYv; R += RYv) {
            const int g = c.bgn == 1 ? R / 22 : R / 10, j = R - g * c.kb;
            uint4 *src = reinterpret_cast<uint4 *>(ck + (long long)(cb0 + g) * c.K + j * Zc + 16 * h);
            uint4 x = *src;
            uint32_t xs[4] = {x.x, x.y, x.z, x.w};
            uint32_t bits = 0, fmask = 0;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                // bytes equal to 0xFF (-1): fillers when j >= 2 (nr_ldpc_encode.py:32-35)
                uint32_t m = (j >= 2) ? (((xs[q] & 0x7f7f7f7fu) + 0x01010101u) & xs[q] & 0x80808080u) >> 7 : 0u;
                const uint32_t b = xs[q] & 0x01010101u & ~m;
                bits |= ((b * 0x01020408u) >> 24) << (4 * q);
                fmask |= ((m * 0x01020408u) >> 24) << (4 * q);
                xs[q] &= ~(m * 0xffu);
            }
            reinterpret_cast<uint16_t *>(V(g, j))[h] = (uint16_t)bits;
            reinterpret_cast<uint16_t *>(FM(g, j))[h] = (uint16_t)fmask;
            if (fmask && fix_fillers) *src = make_uint4(xs[0], xs[1], xs[2], xs[3]);
        }
    }
    for (int it = vec ? 0x7fffffff : warp; it < g_cnt * c.kb * W; it += nwarps) {
        const int g = it / (c.kb * W), rem = it % (c.kb * W), j = rem / W, w = rem % W;
        const int r = 32 * w + lane;
        int val = 0;
        const long long k = (long long)(cb0 + g) * c.K + j * Zc + r;
        if (r < Zc) val = ck[k];
        // nr_ldpc_encode.py:32-35: a -1 at k >= 2Zc is a filler: encoded as 0, reported as -1 in dn
        const bool filler = (val == -1) && (j >= 2);
        const uint32_t bits = __ballot_sync(0xffffffffu, !filler && (val & 1));
        const uint32_t fmask = __ballot_sync(0xffffffffu, filler);
        if (lane == 0) { V(g, j)[w] = bits; FM(g, j)[w] = fmask; }
        if (filler && fix_fillers) ck[k] = 0;
    }
    __syncthreads();

    // B. core parity.  L1[i] = sum over systematic edges of row-block i (i < 4)
    for (int it = threadIdx.x; it < g_cnt * 4 * W; it += kEncThreads) {
        const int g = it / (4 * W), rem = it % (4 * W), i = rem / W, w = rem % W;
        uint32_t acc = 0;
        for (int e = c.rowptr[i]; e < c.rowptr[i + 1]; ++e) {
            const int j = c.edge[e] & 0xff;
            if (j < c.kb) acc ^= rot_word(V(g, j), c.edge[e] >> 8, w, Zc);
        }
        L1(g, i)[w] = acc;
    }
    __syncthreads();
    for (int it = threadIdx.x; it < g_cnt * W; it += kEncThreads) {
        const int g = it / W, w = it % W;
        L1(g, 4)[w] = L1(g, 0)[w] ^ L1(g, 1)[w] ^ L1(g, 2)[w] ^ L1(g, 3)[w];  // L2 (:94)
    }
    __syncthreads();
    const int kb = c.kb;
    // pc1 = np.roll(L2, s), s = shift of block (1,kb) for BG1 / (2,kb) for BG2 (:95-96,:101-102)
    const int s = find_edge(c, c.bgn == 1 ? 1 : 2, kb);
    for (int it = threadIdx.x; it < g_cnt * W; it += kEncThreads) {
        const int g = it / W, w = it % W;
        V(g, kb)[w] = rot_word(L1(g, 4), (Zc - s) % Zc, w, Zc);
    }
    __syncthreads();
    // pc2 = L1[0] + B(0,0) pc1 ; pc4 = L1[3] + B(3,0) pc1 (:97-98,:103-104)
    const int p00 = find_edge(c, 0, kb), p30 = find_edge(c, 3, kb);
    for (int it = threadIdx.x; it < g_cnt * 2 * W; it += kEncThreads) {
        const int g = it / (2 * W), rem = it % (2 * W), which = rem / W, w = rem % W;
        if (which == 0) V(g, kb + 1)[w] = L1(g, 0)[w] ^ rot_word(V(g, kb), p00, w, Zc);
        else V(g, kb + 3)[w] = L1(g, 3)[w] ^ rot_word(V(g, kb), p30, w, Zc);
    }
    __syncthreads();
    // BG1: pc3 = L1[2] + B(2,3) pc4 (:99) ; BG2: pc3 = L1[1] + B(1,1) pc2 (:105)
    const int p3 = c.bgn == 1 ? find_edge(c, 2, kb + 3) : find_edge(c, 1, kb + 1);
    for (int it = threadIdx.x; it < g_cnt * W; it += kEncThreads) {
        const int g = it / W, w = it % W;
        if (c.bgn == 1) V(g, kb + 2)[w] = L1(g, 2)[w] ^ rot_word(V(g, kb + 3), p3, w, Zc);
        else V(g, kb + 2)[w] = L1(g, 1)[w] ^ rot_word(V(g, kb + 1), p3, w, Zc);
    }
    __syncthreads();

    // C. extension parity pe = C [ck; pc]  (:90,:110-112): every edge of rows >= 4 except the last
    const int next = c.nrows - 4;
    for (int it = threadIdx.x; it < g_cnt * next * W; it += kEncThreads) {
        const int g = it / (next * W), rem = it % (next * W), i = 4 + rem / W, w = rem % W;
        uint32_t acc = 0;
        for (int e = c.rowptr[i]; e < c.rowptr[i + 1] - 1; ++e)
            acc ^= rot_word(V(g, c.edge[e] & 0xff), c.edge[e] >> 8, w, Zc);
        V(g, kb + i)[w] = acc;
    }
    __syncthreads();

    // D'. rate matching fused into the store: g straight from the packed codeword
    if (rm.g) {
        const uint32_t mZ = 0xffffffffu / (uint32_t)Zc + 1u;  // pos / Zc by multiplication: exact for pos * Zc < 2^32
        for (int g = 0; g < g_cnt; ++g)
            rm_store(rm, cb0 + g, [&](int pos) {
                const int jo = (int)__umulhi((uint32_t)pos, mZ), r = pos - jo * Zc;
                return (V(g, jo + 2)[r >> 5] >> (r & 31)) & 1u;
            }, threadIdx.x, kEncThreads);
        return;
    }
    // D. unpack dn = codeword without the first 2Zc bits; -1 at filler positions (:31-37,:47-48)
    const int nout = c.ncols - 2;
    if (vec) {
        for (int R = tj; R < g_cnt * nout && tj < RYv; R += RYv) {
            const int g = c.bgn == 1 ? R / 66 : R / 50, jo = R - g * nout, j = jo + 2;
            const uint32_t bits = reinterpret_cast<const uint16_t *>(V(g, j))[h];
            const uint32_t fm = (j < kb) ? reinterpret_cast<const uint16_t *>(FM(g, j))[h] : 0u;
            uint32_t o[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const uint32_t b = (((bits >> (4 * q)) & 0xfu) * 0x00204081u) & 0x01010101u;
                const uint32_t m = (((fm >> (4 * q)) & 0xfu) * 0x00204081u) & 0x01010101u;
                o[q] = b | (m * 0xffu);
            }
            *reinterpret_cast<uint4 *>(dn + (long long)(cb0 + g) * c.N + jo * Zc + 16 * h) = make_uint4(o[0], o[1], o[2], o[3]);
        }
        return;
    }
    for (int it = warp; it < g_cnt * nout * W; it += nwarps) {
        const int g = it / (nout * W), rem = it % (nout * W), j = 2 + rem / W, w = rem % W;
        const int r = 32 * w + lane;
        if (r < Zc) {
            int8_t v = (int8_t)((V(g, j)[w] >> lane) & 1u);
            if (j < kb && ((FM(g, j)[w] >> lane) & 1u)) v = -1;
            dn[(long long)(cb0 + g) * c.N + (j - 2) * Zc + r] = v;
        }
    }
}


// ---------------------------------------------------------------------------------------------------------------
// Word-parallel encoder for Zc % 32 == 0 (W = Zc/32 words per column-block, no partial word).
// Thread <-> (codeblock g, word w); the block is 4 parts of kWPart threads and all threads of a part walk the SAME
// row-block, so the per-edge table reads are warp-uniform and there is no per-edge index arithmetic:
// every column-block that gets rotated is stored twice back to back (word k and word k + W), hence
//   word w of (circulant P) @ v  =  funnelshift_r(v2[w + (P >> 5)], v2[w + (P >> 5) + 1], P & 31)
// with a uniform offset.  The core parity is written in closed form over the four L1 rows
// (rotations compose: rot(rot(v, a), b) = rot(v, a + b)), so pc1..pc4 are computed by the four parts in parallel
// from L1 alone (nr_ldpc_encode.py:92-106) and the kernel has 4 barriers.  dn's words are collected in one
// contiguous array so that the byte <-> bit conversion needs no division on the output side.
constexpr int kWParts = 4, kWPart = 96, kWThreads = kWParts * kWPart;

__device__ __forceinline__ uint32_t rot2(const uint32_t *v2w, int P)
{
    const uint32_t *p = v2w + (P >> 5);
    return __funnelshift_r(p[0], p[1], P & 31);
}

// 16 int8 bits (values & 1) -> 16 packed bits
__device__ __forceinline__ uint32_t pack16(const uint4 &x)
{
    const uint32_t a = ((x.x & 0x0f0f0f0fu) | ((x.y << 4) & 0xf0f0f0f0u)) & 0x11111111u;
    const uint32_t b = ((x.z & 0x0f0f0f0fu) | ((x.w << 4) & 0xf0f0f0f0u)) & 0x11111111u;
    // byte k of a: bit 0 = byte k of x.x, bit 4 = byte k of x.y; the multiply gathers them in bits 24..31
    return ((a * 0x01020408u) >> 24) | (((b * 0x01020408u) >> 16) & 0xff00u);
}

// mask of the bytes equal to 0xFF, one bit per byte
__device__ __forceinline__ uint32_t ffmask16(const uint4 &x)
{
    uint32_t xs[4] = {x.x, x.y, x.z, x.w}, out = 0;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const uint32_t m = (((xs[q] & 0x7f7f7f7fu) + 0x01010101u) & xs[q] & 0x80808080u) >> 7;
        out |= ((m * 0x01020408u) >> 24) << (4 * q);
    }
    return out;
}

__device__ __forceinline__ uint32_t spread4(uint32_t nib) { return (nib * 0x00204081u) & 0x01010101u; }

constexpr int kWSysMax = 20, kWExtMax = 72;

// Everything the kernel needs, passed by value (constant bank).  An edge descriptor is
//   (byte offset of word j*2W + (P >> 5) inside the codeblock's doubled column array) << 16 | last-of-row << 5 | (P & 31)
// so that the address is one LEA.HI and the descriptor itself is the funnel-shift amount (the shift wraps mod 32).
struct EncWordArgs {
    int bgn, W, H, kb, nout, K, N;   // W = ceil(Zc/32) words (threads) per column-block, H = Zc/16 half words
    int G, slot;                 // codeblocks per CTA, words of shared memory per codeblock (without dn's words)
    uint32_t mKH, mNH, mH;       // floor(2^32 / d) + 1 for d = kb*H, nout*H, H   (H = Zc/16)
    int s1, s2, s3, s4;          // composed rotations of the closed-form core parity
    int p3;                      // B(2,3) (BG1) / B(1,1) (BG2)
    uint16_t nsys[kWParts], next[kWParts];
    uint32_t sys[kWParts][kWSysMax];   // systematic edges of row-block `part`
    uint32_t ext[kWParts][kWExtMax];   // edges (all but the last) of row-blocks 4 + part, 8 + part, ...
    EncRmArgs rm;                      // rm.g != nullptr: rate matching fused into the store, dn is not written
};

__device__ __forceinline__ uint32_t rot2d(const uint32_t *v2w, uint32_t d)
{
    const uint32_t *p = reinterpret_cast<const uint32_t *>(reinterpret_cast<const char *>(v2w) + (d >> 16));
    return __funnelshift_r(p[0], p[1], d);
}

// PACKED: ck / dn are bit-packed (nrldpc_encode_packed: K/32 and N/32 little-endian words per codeblock, bit k of a codeblock
// at word k / 32, bit k % 32 -- SURVEY 8(d)'s K/8 + N/8 algorithmic bytes); the byte <-> bit stages A and D are plain word
// copies, there are no fillers (a packed bit cannot be -1) and Zc is a multiple of 32.
template <bool ODD, bool PACKED = false>   // ODD: Zc = 16 * odd (half-word stores into the doubled arrays)
__global__ void __launch_bounds__(kWThreads)
encode_words_kernel(const __grid_constant__ EncWordArgs a, int8_t *__restrict__ ck, int B, int fix_fillers,
                    int8_t *__restrict__ dn)
{
    extern __shared__ uint32_t smem[];
    __shared__ int any_filler;
    // A column-block is H = Zc/16 half words.  Stored twice back to back it is a bit-contiguous array of 2 Zc bits = H
    // words (W2), whatever the parity of H: for Zc = 16 * odd the second copy simply starts on a half-word boundary, which
    // is why every store into a doubled array below is made of half words.  W = ceil(H/2) threads cover a column-block;
    // in the last word of an odd H only the low half is valid.
    const int W = a.W, H = a.H, W2 = H, kb = a.kb, nout = a.nout;
    constexpr bool odd = ODD;
    const int cb0 = blockIdx.x * a.G, g_cnt = min(a.G, B - cb0);
    // per codeblock slot: D2[kb+4][H words] (systematic + core parity, doubled), L2x[4][H words] (L1 rows, doubled),
    // FM[kb-2][H half words] (filler masks, indexed like dn's half words); then OUT[G][nout][H half words]: dn packed,
    // contiguous over the CTA's codeblocks like dn itself
    const int oL1 = (kb + 4) * W2, oFM = oL1 + 4 * W2;
    uint32_t *OUT = smem + a.G * a.slot;
    const int NH = nout * H;
    // word `v` of thread (g, w) into a doubled array (both copies) / into dn's packed half words
    auto store2 = [&](uint32_t *arr, int w, uint32_t v) {
        if (!odd) { arr[w] = v; arr[W + w] = v; return; }
        uint16_t *h = reinterpret_cast<uint16_t *>(arr);
        h[2 * w] = (uint16_t)v; h[H + 2 * w] = (uint16_t)v;
        if (2 * w + 1 < H) { h[2 * w + 1] = (uint16_t)(v >> 16); h[H + 2 * w + 1] = (uint16_t)(v >> 16); }
    };
    auto store_out = [&](uint16_t *o16, int w, uint32_t v) {  // o16 = the column-block's first half word inside OUT
        if (!odd) { reinterpret_cast<uint32_t *>(o16)[w] = v; return; }
        o16[2 * w] = (uint16_t)v;
        if (2 * w + 1 < H) o16[2 * w + 1] = (uint16_t)(v >> 16);
    };
    if (threadIdx.x == 0) any_filler = 0;
    __syncthreads();

    // A. pack: one thread per 16 input bytes; the CTA's codeblocks are contiguous in ck
    if constexpr (PACKED) {
        const int KW = kb * W, total = g_cnt * KW;
        const uint32_t *base = reinterpret_cast<const uint32_t *>(ck) + (long long)cb0 * KW;
        if ((W & 3) == 0) {
            // four words at a time: they share their column-block (W % 4 == 0), every offset below is a multiple of 4 words
            for (int idx = 4 * threadIdx.x; idx < total; idx += 4 * kWThreads) {
                const uint4 v = *reinterpret_cast<const uint4 *>(base + idx);
                const int g = __umulhi(2u * (uint32_t)idx, a.mKH), u = idx - g * KW;  // kb H = 2 kb W
                const int j = __umulhi(2u * (uint32_t)u, a.mH);
                uint32_t *cbw = smem + g * a.slot + u + j * W;  // word u of [kb][W] -> word u + j W of the doubled [kb][2][W]
                *reinterpret_cast<uint4 *>(cbw) = v;
                *reinterpret_cast<uint4 *>(cbw + W) = v;
                if (j >= 2) *reinterpret_cast<uint4 *>(OUT + g * (NH / 2) + u - 2 * W) = v;
            }
        } else {
            for (int idx = threadIdx.x; idx < total; idx += kWThreads) {
                const uint32_t v = base[idx];
                const int g = __umulhi(2u * (uint32_t)idx, a.mKH), u = idx - g * KW;
                const int j = __umulhi(2u * (uint32_t)u, a.mH);
                uint32_t *cbw = smem + g * a.slot + u + j * W;
                cbw[0] = v;
                cbw[W] = v;
                if (j >= 2) OUT[g * (NH / 2) + u - 2 * W] = v;
            }
        }
    } else {
        const int KH = kb * H, total = g_cnt * KH;
        uint4 *base = reinterpret_cast<uint4 *>(ck + (long long)cb0 * a.K);
        uint16_t *OUT16 = reinterpret_cast<uint16_t *>(OUT);
        for (int idx = threadIdx.x; idx < total; idx += kWThreads) {
            uint4 x = base[idx];
            const int g = __umulhi((uint32_t)idx, a.mKH), u = idx - g * KH;
            const int j = __umulhi((uint32_t)u, a.mH);
            uint32_t fm = 0;
            if (j >= 2 && ((x.x | x.y | x.z | x.w) & 0x80808080u)) {
                // bytes equal to -1 at k >= 2Zc are fillers: encoded as 0, reported as -1 (nr_ldpc_encode.py:32-37)
                fm = ffmask16(x);
                if (fm) {
                    uint32_t xs[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
                    for (int q = 0; q < 4; ++q) xs[q] &= ~(spread4((fm >> (4 * q)) & 0xfu) * 0xffu);
                    x = make_uint4(xs[0], xs[1], xs[2], xs[3]);
                    if (fix_fillers) base[idx] = x;
                    any_filler = 1;
                }
            }
            const uint16_t bits = (uint16_t)pack16(x);
            // halfword u of the [kb][H] input lands at halfword u + j*H of the doubled array [kb][2][H]
            uint16_t *cb16 = reinterpret_cast<uint16_t *>(smem + g * a.slot) + u + j * H;
            cb16[0] = bits;
            cb16[H] = bits;
            if (j >= 2) {
                OUT16[g * NH + u - 2 * H] = bits;
                reinterpret_cast<uint16_t *>(smem + g * a.slot + oFM)[u - 2 * H] = (uint16_t)fm;
            }
        }
    }
    __syncthreads();

    const int part = threadIdx.x / kWPart, x = threadIdx.x - part * kWPart;
    const int g = x / W, w = x - g * W;
    const bool act = g < g_cnt;
    uint32_t *cb = smem + g * a.slot;
    uint16_t *out16 = reinterpret_cast<uint16_t *>(OUT) + g * NH;  // dn's half words of codeblock g
    const uint32_t *Dw = cb + w;

    // B. L1[part] = sum over the systematic edges of row-block `part` (:92)
    if (act) {
        uint32_t acc = 0;
        const int n = a.nsys[part];
#pragma unroll 4
        for (int e = 0; e < n; ++e) acc ^= rot2d(Dw, a.sys[part][e]);
        store2(cb + oL1 + part * W2, w, acc);
    }
    __syncthreads();

    // core parity in closed form; part p computes pc_{p+1}.  s1 = roll(L2) as a rotation, s2 = s1 + B(0,0),
    // s4 = s1 + B(3,0); BG1: pc3 = L1[2] + B(2,3) pc4 (s3 = s4 + p3), BG2: pc3 = L1[1] + B(1,1) pc2 (s3 = s2 + p3)
    if (act) {
        const uint32_t *Lw = cb + oL1 + w;
        const int sh = part == 0 ? a.s1 : part == 1 ? a.s2 : part == 2 ? a.s3 : a.s4;
        uint32_t acc = rot2(Lw, sh) ^ rot2(Lw + W2, sh) ^ rot2(Lw + 2 * W2, sh) ^ rot2(Lw + 3 * W2, sh);
        if (part == 1) acc ^= Lw[0];
        else if (part == 3) acc ^= Lw[3 * W2];
        else if (part == 2) {
            if (a.bgn == 1) acc ^= Lw[2 * W2] ^ rot2(Lw + 3 * W2, a.p3);
            else acc ^= Lw[W2] ^ rot2(Lw, a.p3);
        }
        store2(cb + (kb + part) * W2, w, acc);
        store_out(out16 + (kb - 2 + part) * H, w, acc);
    }
    __syncthreads();

    // C. extension parity pe = C [ck; pc] (:110-113): every edge of a row-block >= 4 except its last; one flat
    // edge list per part, the descriptor's bit 5 closes a row-block
    if (act) {
        uint32_t acc = 0;
        uint16_t *o = out16 + (kb + 2 + part) * H;  // row-block 4 + part -> dn column-block kb - 2 + 4 + part
        const int n = a.next[part];
#pragma unroll 4
        for (int e = 0; e < n; ++e) {
            const uint32_t d = a.ext[part][e];
            acc ^= rot2d(Dw, d);
            if (d & 32u) {
                store_out(o, w, acc);
                acc = 0;
                o += kWParts * H;
            }
        }
    }
    __syncthreads();

    // D'. rate matching fused into the store: OUT holds dn's N bits of every codeblock contiguously (Zc = 32 W), so
    // dn[pos] is bit pos of the codeblock's OUT words
    if (a.rm.g) {
        for (int gg = 0; gg < g_cnt; ++gg) {
            const uint32_t *words = OUT + gg * (NH / 2);  // NH = nout * H is even (nout = 66 | 50)
            rm_store(a.rm, cb0 + gg, [&](int pos) { return (words[pos >> 5] >> (pos & 31)) & 1u; }, threadIdx.x, kWThreads);
        }
        return;
    }
    // D. unpack dn: one thread per 16 output bytes, contiguous in dn and in OUT for the CTA's codeblocks
    if constexpr (PACKED) {
        const int total = g_cnt * (NH / 2);
        uint32_t *dst = reinterpret_cast<uint32_t *>(dn) + (long long)cb0 * (NH / 2);
        if (((NH / 2) & 3) == 0 && ((a.G * a.slot) & 3) == 0) {
            for (int idx = 4 * threadIdx.x; idx < total; idx += 4 * kWThreads)
                *reinterpret_cast<uint4 *>(dst + idx) = *reinterpret_cast<const uint4 *>(OUT + idx);
        } else {
            for (int idx = threadIdx.x; idx < total; idx += kWThreads) dst[idx] = OUT[idx];
        }
    } else {
        const int total = g_cnt * NH, FH = (kb - 2) * H;
        uint4 *dst = reinterpret_cast<uint4 *>(dn + (long long)cb0 * a.N);
        const uint16_t *OUT16 = reinterpret_cast<const uint16_t *>(OUT);
        const bool fill = any_filler != 0;
        for (int idx = threadIdx.x; idx < total; idx += kWThreads) {
            const uint32_t bits = OUT16[idx];
            uint4 o = make_uint4(spread4(bits & 0xfu), spread4((bits >> 4) & 0xfu), spread4((bits >> 8) & 0xfu),
                                 spread4(bits >> 12));
            if (fill) {
                const int gg = __umulhi((uint32_t)idx, a.mNH), u = idx - gg * NH;
                if (u < FH) {
                    const uint32_t fm = reinterpret_cast<const uint16_t *>(smem + gg * a.slot + oFM)[u];
                    o.x |= spread4(fm & 0xfu) * 0xffu;
                    o.y |= spread4((fm >> 4) & 0xfu) * 0xffu;
                    o.z |= spread4((fm >> 8) & 0xfu) * 0xffu;
                    o.w |= spread4(fm >> 12) * 0xffu;
                }
            }
            dst[idx] = o;
        }
    }
}

int find_edge_host(const QcCfg &c, int i, int j)
{
    for (int e = c.rowptr[i]; e < c.rowptr[i + 1]; ++e)
        if ((int)(c.edge[e] & 0xff) == j) return (int)(c.edge[e] >> 8);
    return 0;
}

int launch_encode_words(const QcCfg &c, int8_t *d_ck, int B, int fix_fillers, int8_t *d_dn, cudaStream_t s, const EncRmArgs &rm,
                        bool packed = false)
{
    const int Zc = c.Zc, H = Zc / 16, W = (H + 1) / 2, kb = c.kb, nout = c.ncols - 2;
    EncWordArgs a = {};
    a.bgn = c.bgn; a.W = W; a.H = H; a.kb = kb; a.nout = nout; a.K = c.K; a.N = c.N;
    a.rm = rm;
    a.G = kWPart / W;
    int slot = (kb + 8) * H + ((kb - 2) * H + 1) / 2 + 1;  // doubled arrays of H words, filler masks, one pad word (the
                                                           // last word of an odd H reads one word past its doubled array)
    slot += ((W - slot) % 32 + 32) % 32;  // slot = W (mod 32): thread (g, w) falls in bank (g W + w) mod 32
    a.slot = slot;
    auto magic = [](uint32_t d) { return (uint32_t)((1ull << 32) / d) + 1u; };
    a.mKH = magic(kb * H);
    a.mNH = magic(nout * H);
    a.mH = magic(H);
    const int sroll = find_edge_host(c, c.bgn == 1 ? 1 : 2, kb);
    const int p00 = find_edge_host(c, 0, kb), p30 = find_edge_host(c, 3, kb);
    a.p3 = c.bgn == 1 ? find_edge_host(c, 2, kb + 3) : find_edge_host(c, 1, kb + 1);
    a.s1 = (Zc - sroll) % Zc;
    a.s2 = (a.s1 + p00) % Zc;
    a.s4 = (a.s1 + p30) % Zc;
    a.s3 = ((c.bgn == 1 ? a.s4 : a.s2) + a.p3) % Zc;
    auto desc = [&](uint32_t ed, int last) {
        const int j = ed & 0xff, P = ed >> 8;
        return (uint32_t)((j * H + (P >> 5)) * 4) << 16 | (uint32_t)last << 5 | (uint32_t)(P & 31);
    };
    for (int p = 0; p < kWParts; ++p) {
        int n = 0;
        for (int e = c.rowptr[p]; e < c.rowptr[p + 1]; ++e)
            if ((int)(c.edge[e] & 0xff) < kb) {
                if (n >= kWSysMax) { set_error("encode: systematic edge list overflow"); return NRLDPC_EINVAL; }
                a.sys[p][n++] = desc(c.edge[e], 0);
            }
        a.nsys[p] = (uint16_t)n;
        n = 0;
        for (int i = 4 + p; i < c.nrows; i += kWParts) {
            const int e1 = c.rowptr[i + 1] - 1;
            for (int e = c.rowptr[i]; e < e1; ++e) {
                if (n >= kWExtMax) { set_error("encode: extension edge list overflow"); return NRLDPC_EINVAL; }
                a.ext[p][n++] = desc(c.edge[e], e == e1 - 1);
            }
        }
        a.next[p] = (uint16_t)n;
    }
    const int smem_bytes = a.G * (slot + nout * H / 2) * 4 + 16;
    static bool attr_done[64] = {};
    int dev = 0;
    NRLDPC_CUDA(cudaGetDevice(&dev));
    if (dev < 64 && !attr_done[dev]) {
        NRLDPC_CUDA(cudaFuncSetAttribute(encode_words_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 72 * 1024));
        NRLDPC_CUDA(cudaFuncSetAttribute(encode_words_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 72 * 1024));
        NRLDPC_CUDA(cudaFuncSetAttribute(encode_words_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 72 * 1024));
        attr_done[dev] = true;
    }
    const int grid = (B + a.G - 1) / a.G;
    if (packed) encode_words_kernel<false, true><<<grid, kWThreads, smem_bytes, s>>>(a, d_ck, B, 0, d_dn);
    else if (H & 1) encode_words_kernel<true><<<grid, kWThreads, smem_bytes, s>>>(a, d_ck, B, fix_fillers, d_dn);
    else encode_words_kernel<false><<<grid, kWThreads, smem_bytes, s>>>(a, d_ck, B, fix_fillers, d_dn);
    NRLDPC_CUDA(cudaGetLastError());
    return NRLDPC_OK;
}

}  // namespace

int launch_encode_packed(const QcCfg &cfg, const uint32_t *d_ck_words, int B, uint32_t *d_dn_words, cudaStream_t s)
{
    if (B <= 0) return NRLDPC_OK;
    if (cfg.Zc % 32 != 0) { set_error("encode_packed: the lifting size must be a multiple of 32"); return NRLDPC_EINVAL; }
    if ((reinterpret_cast<uintptr_t>(d_ck_words) | reinterpret_cast<uintptr_t>(d_dn_words)) % 16 != 0) {
        set_error("encode_packed: ck_words and dn_words must be 16-byte aligned (128-bit word copies)");
        return NRLDPC_EINVAL;
    }
    return launch_encode_words(cfg, reinterpret_cast<int8_t *>(const_cast<uint32_t *>(d_ck_words)), B, 0,
                               reinterpret_cast<int8_t *>(d_dn_words), s, EncRmArgs{}, true);
}

int launch_encode(const QcCfg &cfg, int8_t *d_ck, int B, int fix_fillers, int8_t *d_dn, cudaStream_t s, const EncRmArgs *rm_in)
{
    if (B <= 0) return NRLDPC_OK;
    EncRmArgs rm;
    if (rm_in) rm = *rm_in;
    if (cfg.Zc % 16 == 0 && cfg.Zc >= 32 && (reinterpret_cast<uintptr_t>(d_ck) | reinterpret_cast<uintptr_t>(d_dn)) % 16 == 0)
        return launch_encode_words(cfg, d_ck, B, fix_fillers, d_dn, s, rm);
    const int Wp = cfg.tiles + 1;
    const int slot_bytes = (cfg.ncols * Wp + cfg.kb * cfg.tiles + 5 * Wp) * 4;
    // enough codeblocks per CTA to give 256 threads work, bounded by 48 KB of static-limit shared memory
    int G = 1;
    while (G < 16 && (G * 2) * slot_bytes <= 48 * 1024 && G * cfg.nrows * cfg.tiles < 2 * kEncThreads) G *= 2;
    const int grid = (B + G - 1) / G;
    const int vec = (cfg.Zc % 16 == 0) && ((reinterpret_cast<uintptr_t>(d_ck) | reinterpret_cast<uintptr_t>(d_dn)) % 16 == 0);
    encode_kernel<<<grid, kEncThreads, G * slot_bytes, s>>>(cfg, d_ck, B, G, fix_fillers, d_dn, vec, rm);
    NRLDPC_CUDA(cudaGetLastError());
    return NRLDPC_OK;
}

}  // namespace nrldpc
