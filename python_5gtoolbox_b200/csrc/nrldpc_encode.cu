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

__device__ __forceinline__ int find_edge(const QcCfg &c, int i, int j)
{
    for (int e = c.rowptr[i]; e < c.rowptr[i + 1]; ++e)
        if ((int)(c.edge[e] & 0xff) == j) return (int)(c.edge[e] >> 8);
    return 0;
}

__global__ void __launch_bounds__(kEncThreads)
encode_kernel(const __grid_constant__ QcCfg c, int8_t *__restrict__ ck, int B, int G, int fix_fillers,
              int8_t *__restrict__ dn, int vec)
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

}  // namespace

int launch_encode(const QcCfg &cfg, int8_t *d_ck, int B, int fix_fillers, int8_t *d_dn, cudaStream_t s)
{
    if (B <= 0) return NRLDPC_OK;
    const int Wp = cfg.tiles + 1;
    const int slot_bytes = (cfg.ncols * Wp + cfg.kb * cfg.tiles + 5 * Wp) * 4;
    // enough codeblocks per CTA to give 256 threads work, bounded by 48 KB of static-limit shared memory
    int G = 1;
    while (G < 16 && (G * 2) * slot_bytes <= 48 * 1024 && G * cfg.nrows * cfg.tiles < 2 * kEncThreads) G *= 2;
    const int grid = (B + G - 1) / G;
    const int vec = (cfg.Zc % 16 == 0) && ((reinterpret_cast<uintptr_t>(d_ck) | reinterpret_cast<uintptr_t>(d_dn)) % 16 == 0);
    encode_kernel<<<grid, kEncThreads, G * slot_bytes, s>>>(cfg, d_ck, B, G, fix_fillers, d_dn, vec);
    NRLDPC_CUDA(cudaGetLastError());
    return NRLDPC_OK;
}

}  // namespace nrldpc
