// nrldpc_sch.cu -- whole-transport-block entry points around the LDPC hot path.
//
// Replaces the per-codeblock Python loops of
//   DLSCHDecode / ULSCH_decoding   (py5gphy/nr_pdsch/nr_dlsch_decode.py:13-109, py5gphy/nr_pusch/nr_ulsch_decode.py:13-110)
//   DLSCHEncode / ULSCH_*          (py5gphy/nr_pdsch/nr_dlsch.py:12-74, py5gphy/nr_pusch/nr_ulsch.py:13-68)
// Receive side: ONE decoder launch whose LLR load is the rate recovery + HARQ combining of every codeblock
// (nrldpc_raterecover.cuh), then one small kernel for CB CRC, code block de-segmentation and the TB CRC.  The
// float64 soft buffer the caller keeps for HARQ is written by the decoder's prologue; when the caller's buffer is
// pinned host memory the kernel stores straight into it over PCIe while the iterations run (no D2H copy).
// Transmit side: TB CRC + segmentation + CB CRC in two small kernels, then the encoder and the rate matcher.
#include <immintrin.h>
#include <sched.h>

#include <algorithm>
#include <array>
#include <chrono>
#include <cstdio>
#include <condition_variable>
#include <cstring>
#include <deque>
#include <map>
#include <mutex>
#include <thread>
#include <vector>

#include "nrldpc_common.cuh"
#include "nrldpc_raterecover.cuh"

// inside a host-buffer entry point, once work has been queued on stream `s`: fail only after the stream has drained, so that
// no copy from / into the caller's buffers is still in flight when the function returns
#define NRLDPC_TRY_SYNC(expr)                                         \
    do {                                                              \
        if (int rc__ = (expr)) { cudaStreamSynchronize(s); return rc__; } \
    } while (0)

namespace nrldpc {

// ------------------------------------------------------------------ pinned host memory: pool + staging

namespace {

bool is_pinned(const void *p)
{
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return at.type == cudaMemoryTypeHost || at.type == cudaMemoryTypeManaged;
}

// Pool of pinned (page-locked, device-mapped) host blocks: cudaHostAlloc is millisecond-class, so blocks are
// recycled by size class; at most kKeep bytes stay cached.
struct PinnedPool {
    std::mutex mu;
    std::multimap<size_t, void *> free_blocks;
    std::map<void *, size_t> live;
    size_t cached = 0;
    static constexpr size_t kKeep = (size_t)2 << 30;
    static size_t round_up(size_t n)
    {
        const size_t g = n <= ((size_t)1 << 20) ? 4096 : ((size_t)1 << 20);
        return (std::max<size_t>(n, 1) + g - 1) / g * g;
    }
    cudaError_t get(size_t n, void **p)
    {
        const size_t sz = round_up(n);
        {
            std::lock_guard<std::mutex> lk(mu);
            auto it = free_blocks.find(sz);
            if (it != free_blocks.end()) {
                *p = it->second;
                free_blocks.erase(it);
                cached -= sz;
                live[*p] = sz;
                return cudaSuccess;
            }
        }
        cudaError_t e = cudaHostAlloc(p, sz, cudaHostAllocPortable | cudaHostAllocMapped);
        if (e != cudaSuccess) return e;
        std::lock_guard<std::mutex> lk(mu);
        live[*p] = sz;
        return cudaSuccess;
    }
    int put(void *p)
    {
        std::unique_lock<std::mutex> lk(mu);
        auto it = live.find(p);
        if (it == live.end()) return NRLDPC_EINVAL;
        const size_t sz = it->second;
        live.erase(it);
        if (cached + sz <= kKeep) {
            free_blocks.emplace(sz, p);
            cached += sz;
            return NRLDPC_OK;
        }
        lk.unlock();
        cudaFreeHost(p);
        return NRLDPC_OK;
    }
};
PinnedPool &pinned_pool()
{
    static PinnedPool *p = new PinnedPool;  // never destroyed: blocks may outlive static destruction
    return *p;
}

// memcpy with non-temporal stores (AVX2) for the large pageable -> pinned copies: the destination is read next by the DMA
// engine, not by this core, so it should neither be fetched for ownership nor evict the caller's working set.
__attribute__((target("avx2"))) void copy_nt_avx2(char *dst, const char *src, size_t n)
{
    while (n && (reinterpret_cast<uintptr_t>(dst) & 31)) { *dst++ = *src++; --n; }
    size_t i = 0;
    for (; i + 128 <= n; i += 128) {
        const __m256i a = _mm256_loadu_si256(reinterpret_cast<const __m256i *>(src + i));
        const __m256i b = _mm256_loadu_si256(reinterpret_cast<const __m256i *>(src + i + 32));
        const __m256i c = _mm256_loadu_si256(reinterpret_cast<const __m256i *>(src + i + 64));
        const __m256i d = _mm256_loadu_si256(reinterpret_cast<const __m256i *>(src + i + 96));
        _mm256_stream_si256(reinterpret_cast<__m256i *>(dst + i), a);
        _mm256_stream_si256(reinterpret_cast<__m256i *>(dst + i + 32), b);
        _mm256_stream_si256(reinterpret_cast<__m256i *>(dst + i + 64), c);
        _mm256_stream_si256(reinterpret_cast<__m256i *>(dst + i + 96), d);
    }
    _mm_sfence();
    if (i < n) memcpy(dst + i, src + i, n - i);
}
void copy_piece(void *dst, const void *src, size_t n)
{
    static const bool nt = __builtin_cpu_supports("avx2") && !(getenv("NRLDPC_COPY_NT") && atoi(getenv("NRLDPC_COPY_NT")) == 0);
    if (nt && n >= ((size_t)64 << 10)) copy_nt_avx2((char *)dst, (const char *)src, n);
    else memcpy(dst, src, n);
}

// Host threads that copy between pageable and pinned memory (a single core moves ~10 GB/s, a PCIe 5 x16 link 55).
struct CopyThreads {
    struct Job { char *dst; const char *src; size_t n; };
    std::mutex mu;
    std::condition_variable cv, cv_done;
    std::deque<Job> q;
    int pending = 0;
    std::vector<std::thread> th;
    int nthreads = 0;
    CopyThreads()
    {
        // helpers = this process's share of the cores it may run on, minus the calling thread (torchrun exports the
        // number of ranks on the node), at most 11 (measured on a 16-core B200 host: 3 helpers 34 GB/s, 7 44 GB/s, 11 46 GB/s, 15 44 GB/s)
        int n = 3;
        if (const char *e = getenv("NRLDPC_COPY_THREADS")) n = atoi(e);
        else {
            int cores = (int)std::thread::hardware_concurrency();
            cpu_set_t set;
            if (sched_getaffinity(0, sizeof(set), &set) == 0) cores = CPU_COUNT(&set);
            int ranks = 1;
            if (const char *e = getenv("LOCAL_WORLD_SIZE")) ranks = std::max(1, atoi(e));
            n = std::max(0, std::min(cores / ranks - 1, 11));
        }
        nthreads = std::max(0, std::min(n, 15));
        for (int i = 0; i < nthreads; ++i) th.emplace_back([this] { loop(); });
        for (auto &t : th) t.detach();
    }
    void loop()
    {
        for (;;) {
            Job j;
            {
                std::unique_lock<std::mutex> lk(mu);
                cv.wait(lk, [this] { return !q.empty(); });
                j = q.front();
                q.pop_front();
            }
            copy_piece(j.dst, j.src, j.n);
            {
                std::lock_guard<std::mutex> lk(mu);
                if (--pending == 0) cv_done.notify_all();
            }
        }
    }
    // memcpy split over the helper threads and the caller
    void copy(void *dst, const void *src, size_t n)
    {
        constexpr size_t kMin = (size_t)256 << 10;
        const int parts = (int)std::min<size_t>((size_t)nthreads + 1, std::max<size_t>(1, n / kMin));
        if (parts <= 1) { copy_piece(dst, src, n); return; }
        const size_t per = ((n + parts - 1) / parts + 63) & ~(size_t)63;
        {
            std::lock_guard<std::mutex> lk(mu);
            for (int i = 1; i < parts; ++i) {
                const size_t o = (size_t)i * per;
                if (o >= n) break;
                q.push_back({(char *)dst + o, (const char *)src + o, std::min(per, n - o)});
                ++pending;
            }
        }
        cv.notify_all();
        copy_piece(dst, src, std::min(per, n));
        std::unique_lock<std::mutex> lk(mu);
        cv_done.wait(lk, [this] { return pending == 0; });
    }
};
std::mutex g_copy_mu;  // one multi-threaded copy at a time (the job queue has a single completion count)
CopyThreads &copy_threads()
{
    static CopyThreads *c = new CopyThreads;
    return *c;
}

}  // namespace

void host_copy(void *dst, const void *src, size_t n)
{
    if (n < ((size_t)2 << 20)) { memcpy(dst, src, n); return; }   // waking the helpers costs more than a 2 MiB memcpy saves
    std::lock_guard<std::mutex> lk(g_copy_mu);
    copy_threads().copy(dst, src, n);
}

// Per-device ring of pinned staging slots for pageable caller memory.
struct StageRing {
    static constexpr int kSlots = 4;
    size_t kSlotBytes = (size_t)8 << 20;
    void *slot[kSlots] = {};
    cudaEvent_t ev[kSlots] = {};
    int next = 0;
    std::mutex mu;
    int init()
    {
        if (slot[0]) return NRLDPC_OK;
        if (const char *e = getenv("NRLDPC_STAGE_MB")) kSlotBytes = (size_t)std::max(1, atoi(e)) << 20;
        for (int i = 0; i < kSlots; ++i) {
            NRLDPC_CUDA(cudaHostAlloc(&slot[i], kSlotBytes, cudaHostAllocPortable));
            NRLDPC_CUDA(cudaEventCreateWithFlags(&ev[i], cudaEventDisableTiming));
        }
        return NRLDPC_OK;
    }
};
static StageRing &stage_ring()
{
    static std::mutex mu;
    static std::map<int, StageRing *> rings;
    int dev = 0;
    cudaGetDevice(&dev);
    std::lock_guard<std::mutex> lk(mu);
    auto it = rings.find(dev);
    if (it == rings.end()) it = rings.emplace(dev, new StageRing).first;
    return *it->second;
}

// Host -> device on stream s.  Pinned source: one asynchronous copy.  Pageable source: chunks staged through the
// pinned ring by the copy threads, so the DMA of chunk i overlaps the host copy of chunk i+1; the source may be
// reused as soon as the call returns.
int h2d_async(void *dst, const void *src, size_t n, cudaStream_t s)
{
    if (n == 0) return NRLDPC_OK;
    if (n <= ((size_t)64 << 10) || is_pinned(src)) {
        NRLDPC_CUDA(cudaMemcpyAsync(dst, src, n, cudaMemcpyHostToDevice, s));
        return NRLDPC_OK;
    }
    StageRing &R = stage_ring();
    std::lock_guard<std::mutex> lk(R.mu);
    if (int rc = R.init()) return rc;
    for (size_t o = 0; o < n; o += R.kSlotBytes) {
        const size_t m = std::min(R.kSlotBytes, n - o);
        const int i = R.next++ % StageRing::kSlots;
        NRLDPC_CUDA(cudaEventSynchronize(R.ev[i]));  // the slot's previous DMA is done
        host_copy(R.slot[i], (const char *)src + o, m);
        NRLDPC_CUDA(cudaMemcpyAsync((char *)dst + o, R.slot[i], m, cudaMemcpyHostToDevice, s));
        NRLDPC_CUDA(cudaEventRecord(R.ev[i], s));
    }
    return NRLDPC_OK;
}

// Device -> host, complete when the call returns (the stream is synchronised).  Pageable destination: through the ring.
int d2h_sync(void *dst, const void *src, size_t n, cudaStream_t s)
{
    if (n == 0) { NRLDPC_CUDA(cudaStreamSynchronize(s)); return NRLDPC_OK; }
    if (n <= ((size_t)64 << 10) || is_pinned(dst)) {
        NRLDPC_CUDA(cudaMemcpyAsync(dst, src, n, cudaMemcpyDeviceToHost, s));
        NRLDPC_CUDA(cudaStreamSynchronize(s));
        return NRLDPC_OK;
    }
    StageRing &R = stage_ring();
    std::lock_guard<std::mutex> lk(R.mu);
    if (int rc = R.init()) return rc;
    // DMA of chunk i+1 overlaps the host copy of chunk i
    size_t o_prev = 0, m_prev = 0;
    int i_prev = -1;
    for (size_t o = 0; o < n; o += R.kSlotBytes) {
        const size_t m = std::min(R.kSlotBytes, n - o);
        const int i = R.next++ % StageRing::kSlots;
        NRLDPC_CUDA(cudaEventSynchronize(R.ev[i]));
        NRLDPC_CUDA(cudaMemcpyAsync(R.slot[i], (const char *)src + o, m, cudaMemcpyDeviceToHost, s));
        NRLDPC_CUDA(cudaEventRecord(R.ev[i], s));
        if (i_prev >= 0) {
            NRLDPC_CUDA(cudaEventSynchronize(R.ev[i_prev]));
            host_copy((char *)dst + o_prev, R.slot[i_prev], m_prev);
        }
        o_prev = o, m_prev = m, i_prev = i;
    }
    NRLDPC_CUDA(cudaEventSynchronize(R.ev[i_prev]));
    host_copy((char *)dst + o_prev, R.slot[i_prev], m_prev);
    NRLDPC_CUDA(cudaStreamSynchronize(s));
    return NRLDPC_OK;
}

// One non-blocking stream per (host thread, device) for the synchronous host-buffer entry points of this file: calls
// from different host threads (one transport block each) overlap their copies and kernels on the device.
int host_stream(cudaStream_t *s)
{
    thread_local std::map<int, cudaStream_t> streams;
    int dev = 0;
    NRLDPC_CUDA(cudaGetDevice(&dev));
    auto it = streams.find(dev);
    if (it == streams.end()) {
        cudaStream_t st;
        NRLDPC_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
        it = streams.emplace(dev, st).first;
    }
    *s = it->second;
    return NRLDPC_OK;
}

namespace {

// ------------------------------------------------------------------ CRC over GF(2)[x] / P

constexpr int kTbThreads = 256, kMaxCbs = 256;  // TS 38.214 transport blocks have at most 152 codeblocks

__host__ __device__ __forceinline__ uint32_t gf2_mulmod(uint32_t a, uint32_t b, int L, uint32_t poly, uint32_t mask)
{
    uint32_t res = 0;
    for (int i = L - 1; i >= 0; --i) {
        const uint32_t top = (res >> (L - 1)) & 1u;
        res = (res << 1) & mask;
        if (top) res ^= poly;
        if ((b >> i) & 1u) res ^= a;
    }
    return res;
}

uint32_t gf2_powx(uint32_t e, int L, uint32_t poly)
{
    const uint32_t mask = (L == 32) ? 0xffffffffu : ((1u << L) - 1u);
    uint32_t base = 2u, acc = 1u;
    while (e) {
        if (e & 1u) acc = gf2_mulmod(acc, base, L, poly, mask);
        base = gf2_mulmod(base, base, L, poly, mask);
        e >>= 1;
    }
    return acc;
}

// polynomial bit patterns of py5gphy/crc/crc.py:96-106 (x^L term dropped)
constexpr uint32_t kPoly24A = 0x864CFBu, kPoly24B = 0x800063u, kPoly16 = 0x1021u;

// Powers of x the transport-block kernels multiply their partial remainders with.  A codeblock row holds cbz payload
// bits (+ Lcb CRC bits); thread t of a CTA owns the bits [t*chunk, (t+1)*chunk) of the row.
struct TbPow {
    uint32_t in_tb[kTbThreads];  // x^(cbz - min(k1(t), cbz)) mod P_tb : bits of the row's payload after the thread's chunk
    uint32_t in_cb[kTbThreads];  // x^(Kp  - min(k1(t), Kp))  mod P_24B, Kp = cbz + Lcb : CRC check over payload + CRC
    uint32_t in_cbp[kTbThreads]; // x^(cbz - min(k1(t), cbz)) mod P_24B                  : CRC generation over the payload
    uint32_t after[kMaxCbs];     // x^(bits of the block after codeblock c's payload (+ tail)) mod P_tb
    int chunk, cbz, Kp, Ltb, Lcb;
    uint32_t ptb;
};
using TbGeom = TbPow;

// x^-n mod P (P has a constant term, so x is invertible): n steps of r -> r / x
uint32_t gf2_pow_neg(int n, int L, uint32_t poly)
{
    uint32_t r = 1u;
    for (int i = 0; i < n; ++i) r = (r & 1u) ? (((r ^ poly) >> 1) | (1u << (L - 1))) : (r >> 1);
    return r;
}

// The block has B = C * cbz bits (transport block + its CRC).  tail = 0: CRC check, the remainder of the whole block.
// tail = -Ltb: CRC generation -- the kernels run their shift registers over the whole block with zeros in the CRC's
// positions, which computes M(x) x^Ltb x^Ltb, and the factor x^-Ltb folded into `after` takes one x^Ltb out again.
void tb_geometry(int C, int cbz, int Lcb, int Ltb, int tail, TbGeom *g)
{
    TbPow &p = *g;
    p.cbz = cbz; p.Lcb = Lcb; p.Kp = cbz + Lcb; p.Ltb = Ltb;
    p.ptb = Ltb == 24 ? kPoly24A : kPoly16;
    p.chunk = (p.Kp + kTbThreads - 1) / kTbThreads;
    const uint32_t mtb = (1u << Ltb) - 1u, m24 = (1u << 24) - 1u;
    // descending exponents step by `chunk`: one product per entry
    const uint32_t xc_tb = gf2_powx((uint32_t)p.chunk, Ltb, p.ptb), xc_cb = gf2_powx((uint32_t)p.chunk, 24, kPoly24B);
    int e_tb_next = -1, e_cb_next = -1;
    for (int t = kTbThreads - 1; t >= 0; --t) {
        const int k1 = std::min((t + 1) * p.chunk, p.Kp);
        const int e_tb = cbz - std::min(k1, cbz), e_cb = p.Kp - k1;
        p.in_tb[t] = (e_tb_next >= 0 && e_tb == e_tb_next + p.chunk) ? gf2_mulmod(p.in_tb[t + 1], xc_tb, Ltb, p.ptb, mtb) : gf2_powx((uint32_t)e_tb, Ltb, p.ptb);
        p.in_cb[t] = (e_cb_next >= 0 && e_cb == e_cb_next + p.chunk) ? gf2_mulmod(p.in_cb[t + 1], xc_cb, 24, kPoly24B, m24) : gf2_powx((uint32_t)e_cb, 24, kPoly24B);
        p.in_cbp[t] = (e_tb_next >= 0 && e_tb == e_tb_next + p.chunk) ? gf2_mulmod(p.in_cbp[t + 1], xc_cb, 24, kPoly24B, m24) : gf2_powx((uint32_t)e_tb, 24, kPoly24B);
        e_tb_next = e_tb, e_cb_next = e_cb;
    }
    const uint32_t xcb = gf2_powx((uint32_t)cbz, Ltb, p.ptb);
    uint32_t f = tail >= 0 ? gf2_powx((uint32_t)tail, Ltb, p.ptb) : gf2_pow_neg(-tail, Ltb, p.ptb);
    for (int c = C - 1; c >= 0; --c) {
        p.after[c] = f;
        f = gf2_mulmod(f, xcb, Ltb, p.ptb, mtb);
    }
}

const TbGeom *get_tb_geometry(int C, int cbz, int Lcb, int Ltb, int tail)
{
    static std::mutex mu;
    static std::map<std::array<int, 5>, TbGeom *> cache;
    std::lock_guard<std::mutex> lk(mu);
    const std::array<int, 5> key{C, cbz, Lcb, Ltb, tail};
    auto it = cache.find(key);
    if (it != cache.end()) return it->second;
    if (cache.size() > 1024) {  // bounded: a link adaptation loop visits a few hundred (C, cbz) pairs at most
        for (auto &kv : cache) delete kv.second;
        cache.clear();
    }
    TbGeom *g = new TbGeom;
    tb_geometry(C, cbz, Lcb, Ltb, tail, g);
    cache[key] = g;
    return g;
}

// XOR-reduce one word per thread over the CTA (kTbThreads threads); result valid in every thread.
__device__ __forceinline__ uint32_t block_xor(uint32_t v, uint32_t *s_part)
{
#pragma unroll
    for (int o = 16; o; o >>= 1) v ^= __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = v;
    __syncthreads();
    uint32_t r = 0;
#pragma unroll
    for (int w = 0; w < kTbThreads / 32; ++w) r ^= s_part[w];
    return r;
}

// Receive side, CTA c = codeblock c: CB CRC24B of the decoded row (C > 1; nr_dlsch_decode.py:93-98 computes and ignores
// it), its cbz payload bits into the transport block (:101-102), and the codeblock's share of the TB CRC remainder
// (:105; crc.nr_crc_decode, py5gphy/crc/crc.py:43-88 = remainder of the whole block is zero).  The last CTA to finish
// publishes the TB verdict.  acc = {remainder, CTAs done}, zero at launch.
__global__ void __launch_bounds__(kTbThreads)
tb_finish_kernel(const __grid_constant__ TbPow pw, const int8_t *__restrict__ ck,
                 long long ck_stride, int C, int A, int8_t *__restrict__ tbblk, uint8_t *__restrict__ tb_err,
                 uint8_t *__restrict__ cb_err, uint32_t *__restrict__ acc)
{
    __shared__ uint32_t s_part[kTbThreads / 32];
    const int c = blockIdx.x, tid = threadIdx.x;
    const int8_t *x = ck + (size_t)c * ck_stride;
    const uint32_t mtb = (1u << pw.Ltb) - 1u, m24 = (1u << 24) - 1u;
    const int k0 = min(tid * pw.chunk, pw.Kp), k1 = min(k0 + pw.chunk, pw.Kp);
    uint32_t rt = 0, rc = 0;
    for (int k = k0; k < k1; ++k) {
        const uint32_t bit = (uint32_t)(x[k] & 1);
        if (k < pw.cbz) {
            const uint32_t fb = ((rt >> (pw.Ltb - 1)) & 1u) ^ bit;
            rt = (rt << 1) & mtb;
            if (fb) rt ^= pw.ptb;
            const long long n = (long long)c * pw.cbz + k;
            if (n < A) tbblk[n] = (int8_t)bit;
        }
        const uint32_t fc = ((rc >> 23) & 1u) ^ bit;
        rc = (rc << 1) & m24;
        if (fc) rc ^= kPoly24B;
    }
    if (rt) rt = gf2_mulmod(rt, pw.in_tb[tid], pw.Ltb, pw.ptb, mtb);
    if (pw.Lcb && rc) rc = gf2_mulmod(rc, pw.in_cb[tid], 24, kPoly24B, m24);
    rt = block_xor(rt, s_part);
    if (pw.Lcb) {
        rc = block_xor(rc, s_part);
        if (tid == 0 && cb_err) cb_err[c] = rc != 0;
    } else if (tid == 0 && cb_err) cb_err[c] = 0;
    if (tid == 0) {
        if (rt) atomicXor(acc, gf2_mulmod(rt, pw.after[c], pw.Ltb, pw.ptb, mtb));
        __threadfence();
        if (atomicAdd(acc + 1, 1u) == (unsigned)C - 1) {
            __threadfence();
            *tb_err = atomicXor(acc, 0u) != 0;
        }
    }
}

// Transmit side, pass 1, CTA c: the share of the TB CRC of payload bits [c*cbz, (c+1)*cbz) & [0, A) of trblk.
// acc[1] |= 1 when a value other than 0 / 1 is met (the reference asserts, py5gphy/crc/crc.py:18-20).
__global__ void __launch_bounds__(kTbThreads)
tb_crc_partial_kernel(const __grid_constant__ TbPow pw, const int8_t *__restrict__ trblk,
                      int A, uint32_t *__restrict__ acc)
{
    __shared__ uint32_t s_part[kTbThreads / 32];
    const int c = blockIdx.x, tid = threadIdx.x;
    const uint32_t mtb = (1u << pw.Ltb) - 1u;
    const int k0 = min(tid * pw.chunk, pw.cbz), k1 = min(k0 + pw.chunk, pw.cbz);
    uint32_t rt = 0, bad = 0;
    for (int k = k0; k < k1; ++k) {
        const long long n = (long long)c * pw.cbz + k;
        const uint32_t raw = n < A ? (uint32_t)(uint8_t)trblk[n] : 0u;  // the CRC's own positions count as zeros: M(x) x^L
        bad |= raw > 1u;
        const uint32_t bit = raw & 1u;
        const uint32_t fb = ((rt >> (pw.Ltb - 1)) & 1u) ^ bit;
        rt = (rt << 1) & mtb;
        if (fb) rt ^= pw.ptb;
    }
    if (rt) rt = gf2_mulmod(rt, pw.in_tb[tid], pw.Ltb, pw.ptb, mtb);
    rt = block_xor(rt, s_part);
    if (tid == 0 && rt) atomicXor(acc, gf2_mulmod(rt, pw.after[c], pw.Ltb, pw.ptb, mtb));
    if (bad) atomicOr(acc + 1, 1u);
}

// Transmit side, pass 2, CTA c: row c of cbs[C,K] = cbz bits of [trblk; TB CRC], CRC24B over them when C > 1, -1 fillers
// (py5gphy/ldpc/nr_ldpc_cbsegment.py:7-33).  acc[0] = the TB CRC remainder of pass 1.
__global__ void __launch_bounds__(kTbThreads)
tb_segment_kernel(const __grid_constant__ TbPow pw, const int8_t *__restrict__ trblk, int A, int K,
                  const uint32_t *__restrict__ acc, int8_t *__restrict__ cbs)
{
    __shared__ uint32_t s_part[kTbThreads / 32];
    const int c = blockIdx.x, tid = threadIdx.x;
    const uint32_t m24 = (1u << 24) - 1u, crc_tb = acc[0];
    int8_t *row = cbs + (size_t)c * K;
    const int k0 = min(tid * pw.chunk, pw.cbz), k1 = min(k0 + pw.chunk, pw.cbz);
    uint32_t rc = 0;
    for (int k = k0; k < k1; ++k) {
        const long long n = (long long)c * pw.cbz + k;
        const uint32_t bit = n < A ? (uint32_t)(trblk[n] & 1) : (crc_tb >> (pw.Ltb - 1 - (int)(n - A))) & 1u;
        row[k] = (int8_t)bit;
        const uint32_t fc = ((rc >> 23) & 1u) ^ bit;
        rc = (rc << 1) & m24;
        if (fc) rc ^= kPoly24B;
    }
    if (pw.Lcb) {
        if (rc) rc = gf2_mulmod(rc, pw.in_cbp[tid], 24, kPoly24B, m24);  // times x^(payload bits after the chunk)
        rc = block_xor(rc, s_part);
        if (tid < 24) row[pw.cbz + tid] = (int8_t)((rc >> (23 - tid)) & 1u);
    }
    for (int k = pw.Kp + tid; k < K; k += kTbThreads) row[k] = -1;
}


// Rate recovery + HARQ combining of codeblock blockIdx.x without a decoder behind it (BF / BP / float64 service paths).
template <typename TIn>
__global__ void __launch_bounds__(256)
sch_recover_kernel(const TIn *__restrict__ src, int N, int Ncb, int k0, int Qm, int F0, int F1, const int32_t *__restrict__ E,
                   const long long *__restrict__ goff, const double *__restrict__ cur, double *__restrict__ soft,
                   float *__restrict__ llr32)
{
    __shared__ double s_red[33];
    const int b = blockIdx.x;
    rr_codeblock<TIn>(src + goff[b], E[b], N, Ncb, k0, Qm, F0, F1, cur ? cur + (size_t)b * N : nullptr,
                      soft ? soft + (size_t)b * N : nullptr, llr32 ? llr32 + (size_t)b * N : nullptr, s_red);
}

struct SchDims { int K, N, Nfull, F0, F1, Ltb, Lcb, cbz; };

int sch_dims(const char *what, const QcCfg &c, int C, int A, int K_apo, int Ncb, int k0, int Qm, SchDims *d)
{
    d->K = c.K; d->N = c.N; d->Nfull = c.Nfull;
    d->Ltb = A > 3824 ? 24 : 16;  // nr_dlsch.py:31-35
    d->Lcb = C > 1 ? 24 : 0;      // ldpc_info.get_cbs_info :27-34
    const long long B = (long long)A + d->Ltb;
    if (C < 1 || C > kMaxCbs || A < 1 || B % C || K_apo != (int)(B / C) + d->Lcb || K_apo > c.K || Ncb <= 0 || Ncb > c.N || k0 < 0 ||
        k0 >= Ncb || Qm <= 0) {
        set_error("%s: inconsistent transport block (C=%d A=%d K_apo=%d K=%d Ncb=%d k0=%d Qm=%d)", what, C, A, K_apo, c.K, Ncb, k0, Qm);
        return NRLDPC_EINVAL;
    }
    d->cbz = (int)(B / C);
    d->F0 = K_apo - 2 * c.Zc;
    d->F1 = c.K - 2 * c.Zc;
    return NRLDPC_OK;
}

int host_offsets(const char *what, const int32_t *E, int C, int Qm, std::vector<long long> *off, long long *total)
{
    off->resize(C);
    long long t = 0;
    for (int b = 0; b < C; ++b) {
        if (E[b] <= 0 || E[b] % Qm || E[b] >= (1 << 24)) { set_error("%s: E[%d]=%d must be a positive multiple of Qm=%d below 2^24", what, b, E[b], Qm); return NRLDPC_EINVAL; }
        (*off)[b] = t;
        t += E[b];
    }
    *total = t;
    return NRLDPC_OK;
}

}  // namespace
}  // namespace nrldpc

using namespace nrldpc;

extern "C" int nrldpc_host_alloc(size_t bytes, void **p)
{
    if (!p) { set_error("host_alloc: null pointer"); return NRLDPC_EINVAL; }
    NRLDPC_CUDA(pinned_pool().get(bytes, p));
    return NRLDPC_OK;
}

extern "C" int nrldpc_host_free(void *p)
{
    if (!p) return NRLDPC_OK;
    if (int rc = pinned_pool().put(p)) { set_error("host_free: not a block of nrldpc_host_alloc"); return rc; }
    return NRLDPC_OK;
}

extern "C" int nrldpc_sch_recover(const void *d_llr_g, int in_f64, int C, int N, int Ncb, int k0, int Qm, int F0, int F1,
                                  const int32_t *d_E, const long long *d_goff, const double *d_cur, double *d_soft,
                                  float *d_llr32, void *stream)
{
    if (C < 0 || N <= 0 || Ncb <= 0 || Ncb > N || k0 < 0 || k0 >= Ncb || Qm <= 0 || !d_llr_g || !d_E || !d_goff) {
        set_error("sch_recover: bad argument");
        return NRLDPC_EINVAL;
    }
    if (C == 0) return NRLDPC_OK;
    cudaStream_t s = (cudaStream_t)stream;
    if (in_f64) sch_recover_kernel<double><<<C, 256, 0, s>>>((const double *)d_llr_g, N, Ncb, k0, Qm, F0, F1, d_E, d_goff, d_cur, d_soft, d_llr32);
    else sch_recover_kernel<float><<<C, 256, 0, s>>>((const float *)d_llr_g, N, Ncb, k0, Qm, F0, F1, d_E, d_goff, d_cur, d_soft, d_llr32);
    NRLDPC_CUDA(cudaGetLastError());
    return NRLDPC_OK;
}

extern "C" int nrldpc_sch_recover_host(const void *llr_g, int in_f64, int C, int N, int Ncb, int k0, int Qm, int Zc, int K_apo,
                                       int K, const int32_t *E, const double *cur, double *soft)
{
    if (C <= 0 || !llr_g || !E || !soft || K_apo > K) { set_error("sch_recover: bad argument"); return C == 0 ? NRLDPC_OK : NRLDPC_EINVAL; }
    std::vector<long long> off;
    long long total = 0;
    if (int rc = host_offsets("sch_recover", E, C, Qm, &off, &total)) return rc;
    cudaStream_t s;
    if (int rc = host_stream(&s)) return rc;
    const size_t esz = in_f64 ? 8 : 4, nsoft = (size_t)C * N * 8;
    ScratchBuf d_in, d_E, d_off, d_cur, d_soft;
    NRLDPC_CUDA(d_in.alloc((size_t)total * esz, s));
    NRLDPC_CUDA(d_E.alloc((size_t)C * 4, s));
    NRLDPC_CUDA(d_off.alloc((size_t)C * 8, s));
    NRLDPC_CUDA(d_soft.alloc(nsoft, s));
    if (cur) NRLDPC_CUDA(d_cur.alloc(nsoft, s));
    NRLDPC_TRY_SYNC(h2d_async(d_in.p, llr_g, (size_t)total * esz, s));
    NRLDPC_TRY_SYNC(h2d_async(d_E.p, E, (size_t)C * 4, s));
    NRLDPC_TRY_SYNC(h2d_async(d_off.p, off.data(), (size_t)C * 8, s));
    if (cur) NRLDPC_TRY_SYNC(h2d_async(d_cur.p, cur, nsoft, s));
    int rc = nrldpc_sch_recover(d_in.p, in_f64, C, N, Ncb, k0, Qm, K_apo - 2 * Zc, K - 2 * Zc, d_E.as<int32_t>(), d_off.as<long long>(),
                                cur ? d_cur.as<double>() : nullptr, d_soft.as<double>(), nullptr, s);
    if (rc == NRLDPC_OK) rc = d2h_sync(soft, d_soft.p, nsoft, s);
    else cudaStreamSynchronize(s);
    return rc;
}

extern "C" int nrldpc_sch_decode(const void *d_llr_g, int in_f64, int C, int bgn, int Zc, int Ncb, int k0, int Qm, int K_apo,
                                 const int32_t *d_E, const long long *d_goff, const double *d_cur, double *d_soft, int max_iter,
                                 float alpha, float beta, int A, int8_t *d_ck, int8_t *d_tbblk, uint8_t *d_tb_err,
                                 uint8_t *d_cb_err, uint8_t *d_status, int32_t *d_iters, void *stream)
{
    const QcCfg *c = get_cfg(bgn, Zc);
    if (!c) return NRLDPC_EINVAL;
    SchDims d;
    if (int rc = sch_dims("sch_decode", *c, C, A, K_apo, Ncb, k0, Qm, &d)) return rc;
    if (max_iter < 0 || !d_llr_g || !d_E || !d_goff || !d_tbblk || !d_tb_err) { set_error("sch_decode: bad argument"); return NRLDPC_EINVAL; }
    cudaStream_t s = (cudaStream_t)stream;
    RrArgs rr;
    rr.src = d_llr_g; rr.E = d_E; rr.goff = d_goff; rr.cur = d_cur; rr.soft = d_soft; rr.in_f64 = in_f64;
    rr.Ncb = Ncb; rr.k0 = k0; rr.Qm = Qm; rr.F0 = d.F0; rr.F1 = d.F1;
    ScratchBuf ck, acc;
    if (!d_ck) {
        NRLDPC_CUDA(ck.alloc((size_t)C * d.Nfull, s));
        d_ck = ck.as<int8_t>();
    }
    NRLDPC_CUDA(acc.alloc(8, s));
    NRLDPC_CUDA(cudaMemsetAsync(acc.p, 0, 8, s));
    if (int rc = launch_decode_minsum(*c, nullptr, C, max_iter, alpha, beta, 1, d_ck, nullptr, d_status, d_iters, s, &rr)) return rc;
    const TbGeom *g = get_tb_geometry(C, d.cbz, d.Lcb, d.Ltb, 0);
    tb_finish_kernel<<<C, kTbThreads, 0, s>>>(*g, d_ck, d.Nfull, C, A, d_tbblk, d_tb_err, d_cb_err, acc.as<uint32_t>());
    NRLDPC_CUDA(cudaGetLastError());
    return NRLDPC_OK;
}

extern "C" int nrldpc_sch_decode_host(const void *llr_g, int in_f64, int C, int bgn, int Zc, int Ncb, int k0, int Qm, int K_apo,
                                      const int32_t *E, const double *cur, double *soft, int max_iter, float alpha, float beta,
                                      int A, int8_t *tbblk, uint8_t *tb_err, uint8_t *cb_err, uint8_t *status, int32_t *iters)
{
    const QcCfg *c = get_cfg(bgn, Zc);
    if (!c) return NRLDPC_EINVAL;
    SchDims d;
    if (int rc = sch_dims("sch_decode", *c, C, A, K_apo, Ncb, k0, Qm, &d)) return rc;
    if (max_iter < 0 || !llr_g || !E || !tbblk || !tb_err) { set_error("sch_decode: bad argument"); return NRLDPC_EINVAL; }
    std::vector<long long> off;
    long long total = 0;
    if (int rc = host_offsets("sch_decode", E, C, Qm, &off, &total)) return rc;
    cudaStream_t s;
    if (int rc = host_stream(&s)) return rc;
    // NRLDPC_TRACE=1: host clock at every stage boundary (=2: with a stream synchronisation in front, so that the
    // device time of a stage is attributed to it); timing aid only
    static const int trace = getenv("NRLDPC_TRACE") ? atoi(getenv("NRLDPC_TRACE")) : 0;
    double tt[10] = {};
    int nt = 0;
    auto mark = [&]() {
        if (!trace || nt >= 10) return;
        if (trace > 1) cudaStreamSynchronize(s);
        tt[nt++] = std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now().time_since_epoch()).count();
    };
    mark();
    const size_t esz = in_f64 ? 8 : 4, nsoft = (size_t)C * d.N * 8;
    // small results in one block: tb_err | cb_err[C] | status[C] | iters[C], each part 16-byte aligned
    const size_t o_cb = 16, o_st = o_cb + (((size_t)C + 15) & ~(size_t)15), o_it = o_st + (((size_t)C + 15) & ~(size_t)15);
    const size_t nsmall = o_it + (size_t)C * 4;
    ScratchBuf d_in, d_E, d_off, d_cur, d_soft, d_tb, d_small;
    NRLDPC_CUDA(d_in.alloc((size_t)total * esz, s));
    NRLDPC_CUDA(d_E.alloc((size_t)C * 4, s));
    NRLDPC_CUDA(d_off.alloc((size_t)C * 8, s));
    NRLDPC_CUDA(d_tb.alloc((size_t)A, s));
    NRLDPC_CUDA(d_small.alloc(nsmall, s));
    // the float64 soft buffer: stored straight into the caller's buffer by the decoder when that is pinned host memory
    // (zero-copy over PCIe, overlapping the iterations; NRLDPC_SOFT_D2H=1: device buffer + copy, for A/B timing)
    static const bool soft_d2h = getenv("NRLDPC_SOFT_D2H") != nullptr;
    double *soft_dev = nullptr;
    bool soft_copy = false;
    if (soft) {
        void *dp = nullptr;
        if (!soft_d2h && is_pinned(soft) && cudaHostGetDevicePointer(&dp, soft, 0) == cudaSuccess) soft_dev = (double *)dp;
        else {
            cudaGetLastError();
            NRLDPC_CUDA(d_soft.alloc(nsoft, s));
            soft_dev = d_soft.as<double>();
            soft_copy = true;
        }
    }
    // the soft buffer of the earlier transmissions: read by the decoder's prologue straight from the caller's pinned host
    // memory (zero-copy: the reads go host -> device while the new soft buffer is stored device -> host, the two directions
    // of the link; NRLDPC_CUR_H2D=1: copied to the device first, for A/B timing); pageable memory is staged and copied
    static const bool cur_h2d = getenv("NRLDPC_CUR_H2D") != nullptr;
    const double *cur_dev = nullptr;
    if (cur) {
        void *dp = nullptr;
        if (!cur_h2d && is_pinned(cur) && cudaHostGetDevicePointer(&dp, const_cast<double *>(cur), 0) == cudaSuccess) cur_dev = (const double *)dp;
        else {
            cudaGetLastError();
            NRLDPC_CUDA(d_cur.alloc(nsoft, s));
            cur_dev = d_cur.as<double>();
        }
    }
    mark();
    // from here on copies from / kernels storing into the caller's buffers may be in flight: no return without a
    // synchronisation of the stream
    int rc = h2d_async(d_E.p, E, (size_t)C * 4, s);
    if (rc == NRLDPC_OK) rc = h2d_async(d_off.p, off.data(), (size_t)C * 8, s);
    if (rc == NRLDPC_OK) rc = h2d_async(d_in.p, llr_g, (size_t)total * esz, s);
    if (rc == NRLDPC_OK && cur && d_cur.p) rc = h2d_async(d_cur.p, cur, nsoft, s);
    if (rc != NRLDPC_OK) { cudaStreamSynchronize(s); return rc; }
    mark();
    uint8_t *sm = d_small.as<uint8_t>();
    rc = nrldpc_sch_decode(d_in.p, in_f64, C, bgn, Zc, Ncb, k0, Qm, K_apo, d_E.as<int32_t>(), d_off.as<long long>(),
                               cur_dev, soft_dev, max_iter, alpha, beta, A, nullptr, d_tb.as<int8_t>(), sm,
                               sm + o_cb, sm + o_st, reinterpret_cast<int32_t *>(sm + o_it), s);
    if (rc != NRLDPC_OK) { cudaStreamSynchronize(s); return rc; }
    mark();
    void *h_small = nullptr;
    if (cudaError_t ea = pinned_pool().get(nsmall, &h_small)) { cudaStreamSynchronize(s); return cuda_fail(ea, "cudaHostAlloc(results)"); }
    cudaError_t e = cudaMemcpyAsync(h_small, sm, nsmall, cudaMemcpyDeviceToHost, s);
    if (e == cudaSuccess && soft_copy) {
        rc = d2h_sync(soft, d_soft.p, nsoft, s);
        if (rc != NRLDPC_OK) { cudaStreamSynchronize(s); pinned_pool().put(h_small); return rc; }
    }
    if (e == cudaSuccess) rc = d2h_sync(tbblk, d_tb.p, (size_t)A, s);  // synchronises the stream
    else cudaStreamSynchronize(s);
    if (e == cudaSuccess && rc == NRLDPC_OK) {
        const uint8_t *h = (const uint8_t *)h_small;
        *tb_err = h[0];
        if (cb_err) memcpy(cb_err, h + o_cb, (size_t)C);
        if (status) memcpy(status, h + o_st, (size_t)C);
        if (iters) memcpy(iters, h + o_it, (size_t)C * 4);
    }
    pinned_pool().put(h_small);
    mark();
    if (trace && nt == 5)
        fprintf(stderr, "nrldpc trace sch_decode_host C=%d: setup %.0f us, h2d %.0f us, launches %.0f us, d2h+sync %.0f us (soft %s)\n", C,
                tt[1] - tt[0], tt[2] - tt[1], tt[3] - tt[2], tt[4] - tt[3], soft ? (soft_copy ? "copied" : "zero-copy") : "none");
    if (e != cudaSuccess) return cuda_fail(e, "cudaMemcpyAsync(results)");
    return rc;
}

// ------------------------------------------------------------------ transmit side

namespace {
// d_acc (two words) receives {TB CRC remainder, non-binary input flag}
int sch_segment_impl(const int8_t *d_trblk, int A, int C, int K, int8_t *d_cbs, uint32_t *d_acc, cudaStream_t s)
{
    const int Ltb = A > 3824 ? 24 : 16, Lcb = C > 1 ? 24 : 0;
    const long long B = (long long)A + Ltb;
    if (A < 1 || C < 1 || C > kMaxCbs || B % C || B / C + Lcb > K || !d_trblk || !d_cbs) {
        set_error("sch_segment: bad argument (A=%d C=%d K=%d)", A, C, K);
        return NRLDPC_EINVAL;
    }
    const TbGeom *g = get_tb_geometry(C, (int)(B / C), Lcb, Ltb, -Ltb);
    NRLDPC_CUDA(cudaMemsetAsync(d_acc, 0, 8, s));
    tb_crc_partial_kernel<<<C, kTbThreads, 0, s>>>(*g, d_trblk, A, d_acc);
    tb_segment_kernel<<<C, kTbThreads, 0, s>>>(*g, d_trblk, A, K, d_acc, d_cbs);
    NRLDPC_CUDA(cudaGetLastError());
    return NRLDPC_OK;
}
// after a host-side synchronisation: the reference's assertion on non-binary input
int check_binary(const uint32_t *h_acc)
{
    if (h_acc[1]) { set_error("transport block bits must be 0 or 1"); return NRLDPC_EINVAL; }
    return NRLDPC_OK;
}
}  // namespace

extern "C" int nrldpc_sch_segment(const int8_t *d_trblk, int A, int C, int K, int8_t *d_cbs, void *stream)
{
    cudaStream_t s = (cudaStream_t)stream;
    ScratchBuf acc;
    NRLDPC_CUDA(acc.alloc(8, s));
    return sch_segment_impl(d_trblk, A, C, K, d_cbs, acc.as<uint32_t>(), s);
}

extern "C" int nrldpc_sch_segment_host(const int8_t *trblk, int A, int C, int K, int8_t *cbs)
{
    if (!trblk || !cbs || A < 1 || C < 1 || K < 1) { set_error("sch_segment: bad argument"); return NRLDPC_EINVAL; }
    cudaStream_t s;
    if (int rc = host_stream(&s)) return rc;
    ScratchBuf d_tb, d_cbs, d_acc;
    NRLDPC_CUDA(d_tb.alloc((size_t)A, s));
    NRLDPC_CUDA(d_cbs.alloc((size_t)C * K, s));
    NRLDPC_CUDA(d_acc.alloc(8, s));
    NRLDPC_TRY_SYNC(h2d_async(d_tb.p, trblk, (size_t)A, s));
    int rc = sch_segment_impl(d_tb.as<int8_t>(), A, C, K, d_cbs.as<int8_t>(), d_acc.as<uint32_t>(), s);
    uint32_t h_acc[2] = {0, 0};
    if (rc == NRLDPC_OK && cudaMemcpyAsync(h_acc, d_acc.p, 8, cudaMemcpyDeviceToHost, s) != cudaSuccess) rc = cuda_fail(cudaGetLastError(), "cudaMemcpyAsync");
    if (rc == NRLDPC_OK) rc = d2h_sync(cbs, d_cbs.p, (size_t)C * K, s);
    else cudaStreamSynchronize(s);
    return rc == NRLDPC_OK ? check_binary(h_acc) : rc;
}

extern "C" int nrldpc_encode_ratematch(int8_t *d_cbs, int C, int bgn, int Zc, int fix_fillers, int Ncb, int k0, int Qm,
                                       const int32_t *d_E, const long long *d_goff, int8_t *d_g, void *stream)
{
    const QcCfg *c = get_cfg(bgn, Zc);
    if (!c) return NRLDPC_EINVAL;
    if (C < 0 || !d_cbs || !d_E || !d_goff || !d_g) { set_error("encode_ratematch: bad argument"); return NRLDPC_EINVAL; }
    if (C == 0) return NRLDPC_OK;
    cudaStream_t s = (cudaStream_t)stream;
    ScratchBuf dn;
    NRLDPC_CUDA(dn.alloc((size_t)C * c->N, s));
    if (int rc = launch_encode(*c, d_cbs, C, fix_fillers, dn.as<int8_t>(), s)) return rc;
    return nrldpc_ratematch(dn.as<int8_t>(), C, c->N, Ncb, k0, Qm, d_E, d_goff, d_g, s);
}

namespace {
int encode_chain_host(const int8_t *trblk, int A, int8_t *cbs, int C, int bgn, int Zc, int fix_fillers, int Ncb, int k0, int Qm,
                      const int32_t *E, int8_t *g)
{
    const QcCfg *c = get_cfg(bgn, Zc);
    if (!c) return NRLDPC_EINVAL;
    if (C <= 0 || !E || !g || (!trblk && !cbs)) { set_error("sch_encode: bad argument"); return NRLDPC_EINVAL; }
    std::vector<long long> off;
    long long total = 0;
    if (int rc = host_offsets("sch_encode", E, C, Qm, &off, &total)) return rc;
    cudaStream_t s;
    if (int rc = host_stream(&s)) return rc;
    ScratchBuf d_tb, d_cbs, d_E, d_off, d_g, d_acc;
    uint32_t h_acc[2] = {0, 0};
    NRLDPC_CUDA(d_cbs.alloc((size_t)C * c->K, s));
    NRLDPC_CUDA(d_E.alloc((size_t)C * 4, s));
    NRLDPC_CUDA(d_off.alloc((size_t)C * 8, s));
    NRLDPC_CUDA(d_g.alloc((size_t)total, s));
    NRLDPC_TRY_SYNC(h2d_async(d_E.p, E, (size_t)C * 4, s));
    NRLDPC_TRY_SYNC(h2d_async(d_off.p, off.data(), (size_t)C * 8, s));
    int rc = NRLDPC_OK;
    if (trblk) {
        NRLDPC_CUDA(d_tb.alloc((size_t)A, s));
        rc = h2d_async(d_tb.p, trblk, (size_t)A, s);
        NRLDPC_CUDA(d_acc.alloc(8, s));
        if (rc == NRLDPC_OK) rc = sch_segment_impl(d_tb.as<int8_t>(), A, C, c->K, d_cbs.as<int8_t>(), d_acc.as<uint32_t>(), s);
        if (rc == NRLDPC_OK && cudaMemcpyAsync(h_acc, d_acc.p, 8, cudaMemcpyDeviceToHost, s) != cudaSuccess) rc = cuda_fail(cudaGetLastError(), "cudaMemcpyAsync");
    } else {
        rc = h2d_async(d_cbs.p, cbs, (size_t)C * c->K, s);
    }
    if (rc == NRLDPC_OK && trblk) {
        // the codeblocks come from tb_segment_kernel: the fillers are the tail [K_apo, K) of every row, so the rate matching
        // runs inside the encoder's store and dn never exists
        const int Lcb = C > 1 ? 24 : 0, Ltb = A > 3824 ? 24 : 16;
        EncRmArgs rm;
        rm.g = d_g.as<int8_t>(); rm.E = d_E.as<int32_t>(); rm.goff = d_off.as<long long>();
        rm.Ncb = Ncb; rm.k0 = k0; rm.Qm = Qm;
        rm.F0 = (A + Ltb) / C + Lcb - 2 * Zc; rm.F1 = c->K - 2 * Zc;
        if (Ncb <= 0 || Ncb > c->N || k0 < 0 || k0 >= Ncb || Qm <= 0) { set_error("sch_encode: bad rate-matching arguments"); rc = NRLDPC_EINVAL; }
        else rc = launch_encode(*c, d_cbs.as<int8_t>(), C, 1, d_g.as<int8_t>() /* unused */, s, &rm);
    } else if (rc == NRLDPC_OK) {  // caller-provided codeblocks may carry any -1 pattern: encoder, then the scanning rate matcher
        rc = nrldpc_encode_ratematch(d_cbs.as<int8_t>(), C, bgn, Zc, fix_fillers, Ncb, k0, Qm, d_E.as<int32_t>(), d_off.as<long long>(),
                                     d_g.as<int8_t>(), s);
    }
    if (rc == NRLDPC_OK && !trblk && fix_fillers) {  // the reference's in-place side effect on the caller's cbs (nr_ldpc_encode.py:32-35)
        cudaError_t e = cudaMemcpyAsync(cbs, d_cbs.p, (size_t)C * c->K, cudaMemcpyDeviceToHost, s);
        if (e != cudaSuccess) rc = cuda_fail(e, "cudaMemcpyAsync(cbs)");
    }
    if (rc == NRLDPC_OK) rc = d2h_sync(g, d_g.p, (size_t)total, s);
    else cudaStreamSynchronize(s);
    return rc == NRLDPC_OK ? check_binary(h_acc) : rc;
}
}  // namespace

extern "C" int nrldpc_encode_ratematch_host(int8_t *cbs, int C, int bgn, int Zc, int fix_fillers, int Ncb, int k0, int Qm,
                                            const int32_t *E, int8_t *g)
{
    return encode_chain_host(nullptr, 0, cbs, C, bgn, Zc, fix_fillers, Ncb, k0, Qm, E, g);
}

extern "C" int nrldpc_sch_encode_host(const int8_t *trblk, int A, int C, int bgn, int Zc, int Ncb, int k0, int Qm, const int32_t *E,
                                      int8_t *g)
{
    if (!trblk) { set_error("sch_encode: bad argument"); return NRLDPC_EINVAL; }
    return encode_chain_host(trblk, A, nullptr, C, bgn, Zc, 1, Ncb, k0, Qm, E, g);
}
