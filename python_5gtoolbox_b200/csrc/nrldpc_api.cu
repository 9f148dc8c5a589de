// nrldpc_api.cu -- the extern "C" boundary of libnrldpc_b200.so (see include/nrldpc_b200.h).
#include <cuda_fp16.h>

#include <array>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <map>
#include <mutex>
#include <vector>

#include "nrldpc_common.cuh"

namespace nrldpc {

static thread_local char g_err[512] = "";

void set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int cuda_fail(cudaError_t e, const char *what)
{
    set_error("CUDA error %d (%s) in %s", (int)e, cudaGetErrorString(e), what);
    if (e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver) return NRLDPC_ENODEV;
    if (e == cudaErrorMemoryAllocation) return NRLDPC_ENOMEM;
    return NRLDPC_ECUDA;
}

// (bgn, Zc) -> quasi-cyclic tables, built once
const QcCfg *get_cfg(int bgn, int Zc)
{
    static std::mutex mu;
    static std::map<int, QcCfg *> cache;
    std::lock_guard<std::mutex> lk(mu);
    const int key = bgn * 1024 + Zc;
    auto it = cache.find(key);
    if (it != cache.end()) return it->second;
    QcCfg *c = new QcCfg;
    if (build_qc_cfg(bgn, Zc, c) != NRLDPC_OK) {
        delete c;
        set_error("invalid (bgn=%d, Zc=%d): bgn must be 1|2 and Zc a TS 38.212 lifting size", bgn, Zc);
        return nullptr;
    }
    cache[key] = c;
    return c;
}

cudaError_t lib_mempool(cudaMemPool_t *pool)
{
    constexpr int kMaxDev = 64;
    static std::mutex mu;
    static cudaMemPool_t pools[kMaxDev] = {};
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev < 0 || dev >= kMaxDev) return cudaErrorInvalidDevice;
    std::lock_guard<std::mutex> lk(mu);
    if (!pools[dev]) {
        cudaMemPoolProps props = {};
        props.allocType = cudaMemAllocationTypePinned;
        props.handleTypes = cudaMemHandleTypeNone;
        props.location.type = cudaMemLocationTypeDevice;
        props.location.id = dev;
        cudaMemPool_t p;
        e = cudaMemPoolCreate(&p, &props);
        if (e != cudaSuccess) return e;
        unsigned long long keep = 1ull << 30;  // freed scratch beyond 1 GiB goes back to the driver at the next synchronisation
        cudaMemPoolSetAttribute(p, cudaMemPoolAttrReleaseThreshold, &keep);
        pools[dev] = p;
    }
    *pool = pools[dev];
    return cudaSuccess;
}

using DevBuf = ScratchBuf;

// CSR -> CSC (edges of every column in ascending row order = the reference's B lists,
// py5gphy/ldpc/nr_ldpc_decode.py:88-91)
static void csr_to_csc(int M, int Nv, const int32_t *rowptr, const int32_t *colidx, std::vector<int32_t> &cptr,
                       std::vector<int32_t> &cedge, std::vector<int32_t> &crow)
{
    const int E = rowptr[M];
    cptr.assign(Nv + 1, 0);
    cedge.resize(E ? E : 1);
    crow.resize(E ? E : 1);
    for (int e = 0; e < E; ++e) cptr[colidx[e] + 1]++;
    for (int n = 0; n < Nv; ++n) cptr[n + 1] += cptr[n];
    std::vector<int32_t> pos(cptr.begin(), cptr.end() - 1);
    for (int m = 0; m < M; ++m)
        for (int e = rowptr[m]; e < rowptr[m + 1]; ++e) {
            const int q = pos[colidx[e]]++;
            cedge[q] = e;
            crow[q] = m;
        }
}

static int check_csr(int M, int Nv, const int32_t *rowptr, const int32_t *colidx)
{
    if (M <= 0 || Nv <= 0 || !rowptr || !colidx || rowptr[0] != 0) { set_error("bad CSR matrix"); return NRLDPC_EINVAL; }
    for (int m = 0; m < M; ++m)
        if (rowptr[m + 1] < rowptr[m]) { set_error("bad CSR rowptr"); return NRLDPC_EINVAL; }
    for (int e = 0; e < rowptr[M]; ++e)
        if (colidx[e] < 0 || colidx[e] >= Nv) { set_error("bad CSR column index"); return NRLDPC_EINVAL; }
    return NRLDPC_OK;
}

template <typename T>
static int soft_csr_host(const T *llr, int B, int M, int Nv, const int32_t *rowptr, const int32_t *colidx, int prepend,
                         int max_iter, int algo, double alpha, double beta, int early_term, int8_t *ck,
                         uint8_t *status, int32_t *iters)
{
    if (B < 0 || max_iter < 0 || (algo != NRLDPC_ALGO_MINSUM && algo != NRLDPC_ALGO_BP) || !llr || !ck) {
        set_error("decode_csr: bad argument");
        return NRLDPC_EINVAL;
    }
    if (int rc = check_csr(M, Nv, rowptr, colidx)) return rc;
    if (B == 0) return NRLDPC_OK;
    const int E = rowptr[M], Nin = Nv - prepend;
    std::vector<int32_t> cptr, cedge, crow;
    csr_to_csc(M, Nv, rowptr, colidx, cptr, cedge, crow);
    DevBuf d_llr, d_rp, d_ci, d_cp, d_ce, d_work, d_ck, d_st, d_it;
    // bound the workspace: process the batch in chunks
    const size_t per_cb = ((size_t)E + 2 * (size_t)Nv) * sizeof(T);
    int chunk = (int)std::max<size_t>(1, std::min<size_t>((size_t)B, ((size_t)2 << 30) / per_cb));
    NRLDPC_CUDA(d_llr.alloc((size_t)chunk * Nin * sizeof(T)));
    NRLDPC_CUDA(d_rp.alloc((size_t)(M + 1) * 4));
    NRLDPC_CUDA(d_ci.alloc((size_t)(E ? E : 1) * 4));
    NRLDPC_CUDA(d_cp.alloc((size_t)(Nv + 1) * 4));
    NRLDPC_CUDA(d_ce.alloc((size_t)(E ? E : 1) * 4));
    NRLDPC_CUDA(d_work.alloc((size_t)chunk * per_cb));
    NRLDPC_CUDA(d_ck.alloc((size_t)chunk * Nv));
    NRLDPC_CUDA(d_st.alloc((size_t)chunk));
    NRLDPC_CUDA(d_it.alloc((size_t)chunk * 4));
    NRLDPC_CUDA(cudaMemcpy(d_rp.p, rowptr, (size_t)(M + 1) * 4, cudaMemcpyHostToDevice));
    NRLDPC_CUDA(cudaMemcpy(d_ci.p, colidx, (size_t)E * 4, cudaMemcpyHostToDevice));
    NRLDPC_CUDA(cudaMemcpy(d_cp.p, cptr.data(), (size_t)(Nv + 1) * 4, cudaMemcpyHostToDevice));
    NRLDPC_CUDA(cudaMemcpy(d_ce.p, cedge.data(), (size_t)E * 4, cudaMemcpyHostToDevice));
    for (int b0 = 0; b0 < B; b0 += chunk) {
        const int nb = std::min(chunk, B - b0);
        NRLDPC_CUDA(cudaMemcpy(d_llr.p, llr + (size_t)b0 * Nin, (size_t)nb * Nin * sizeof(T), cudaMemcpyHostToDevice));
        if (int rc = launch_soft_csr<T>(d_llr.as<T>(), nb, M, Nv, E, d_rp.as<int32_t>(), d_ci.as<int32_t>(),
                                        d_cp.as<int32_t>(), d_ce.as<int32_t>(), prepend, max_iter, algo, (T)alpha,
                                        (T)beta, early_term, d_work.as<T>(), d_ck.as<int8_t>(), d_st.as<uint8_t>(),
                                        d_it.as<int32_t>(), 0))
            return rc;
        NRLDPC_CUDA(cudaMemcpy(ck + (size_t)b0 * Nv, d_ck.p, (size_t)nb * Nv, cudaMemcpyDeviceToHost));
        if (status) NRLDPC_CUDA(cudaMemcpy(status + b0, d_st.p, (size_t)nb, cudaMemcpyDeviceToHost));
        if (iters) NRLDPC_CUDA(cudaMemcpy(iters + b0, d_it.p, (size_t)nb * 4, cudaMemcpyDeviceToHost));
    }
    return NRLDPC_OK;
}

static int bf_csr_host(const double *llr, int B, int M, int Nv, const int32_t *rowptr, const int32_t *colidx, int prepend,
                       int max_iter, int8_t *ck, uint8_t *status, int32_t *iters)
{
    if (B < 0 || max_iter < 0 || !llr || !ck) { set_error("decode_bf: bad argument"); return NRLDPC_EINVAL; }
    if (int rc = check_csr(M, Nv, rowptr, colidx)) return rc;
    if (B == 0) return NRLDPC_OK;
    const int E = rowptr[M], Nin = Nv - prepend;
    std::vector<int32_t> cptr, cedge, crow;
    csr_to_csc(M, Nv, rowptr, colidx, cptr, cedge, crow);
    DevBuf d_llr, d_rp, d_ci, d_cp, d_cr, d_work, d_ck, d_st, d_it;
    NRLDPC_CUDA(d_llr.alloc((size_t)B * Nin * 8));
    NRLDPC_CUDA(d_rp.alloc((size_t)(M + 1) * 4));
    NRLDPC_CUDA(d_ci.alloc((size_t)(E ? E : 1) * 4));
    NRLDPC_CUDA(d_cp.alloc((size_t)(Nv + 1) * 4));
    NRLDPC_CUDA(d_cr.alloc((size_t)(E ? E : 1) * 4));
    NRLDPC_CUDA(d_work.alloc((size_t)B * M * 4));
    NRLDPC_CUDA(d_ck.alloc((size_t)B * Nv));
    NRLDPC_CUDA(d_st.alloc((size_t)B));
    NRLDPC_CUDA(d_it.alloc((size_t)B * 4));
    NRLDPC_CUDA(cudaMemcpy(d_llr.p, llr, (size_t)B * Nin * 8, cudaMemcpyHostToDevice));
    NRLDPC_CUDA(cudaMemcpy(d_rp.p, rowptr, (size_t)(M + 1) * 4, cudaMemcpyHostToDevice));
    NRLDPC_CUDA(cudaMemcpy(d_ci.p, colidx, (size_t)E * 4, cudaMemcpyHostToDevice));
    NRLDPC_CUDA(cudaMemcpy(d_cp.p, cptr.data(), (size_t)(Nv + 1) * 4, cudaMemcpyHostToDevice));
    NRLDPC_CUDA(cudaMemcpy(d_cr.p, crow.data(), (size_t)E * 4, cudaMemcpyHostToDevice));
    if (int rc = launch_bf_csr(d_llr.as<double>(), B, M, Nv, E, d_rp.as<int32_t>(), d_ci.as<int32_t>(),
                               d_cp.as<int32_t>(), d_cr.as<int32_t>(), prepend, max_iter, d_work.as<int32_t>(),
                               d_ck.as<int8_t>(), d_st.as<uint8_t>(), d_it.as<int32_t>(), 0))
        return rc;
    NRLDPC_CUDA(cudaMemcpy(ck, d_ck.p, (size_t)B * Nv, cudaMemcpyDeviceToHost));
    if (status) NRLDPC_CUDA(cudaMemcpy(status, d_st.p, (size_t)B, cudaMemcpyDeviceToHost));
    if (iters) NRLDPC_CUDA(cudaMemcpy(iters, d_it.p, (size_t)B * 4, cudaMemcpyDeviceToHost));
    return NRLDPC_OK;
}

// Mixed-(bgn, Zc) batches: every group is one launch; the groups run on a small pool of side streams that fork
// from and join back into the caller's stream, so that transport blocks of a few codeblocks each (a persistent CTA
// per codeblock fills one SM) share the GPU instead of running one after the other.
namespace {
constexpr int kSideStreams = 8;
struct SideStreams {
    cudaStream_t s[kSideStreams] = {};
    cudaEvent_t fork = nullptr, join[kSideStreams] = {};
};
std::mutex g_side_mu;
std::map<int, SideStreams> g_side;

template <class F>
int fork_join(cudaStream_t caller, int ngroups, F &&launch_group)
{
    if (ngroups <= 0) return NRLDPC_OK;
    if (ngroups == 1) return launch_group(0, caller);
    int dev = 0;
    NRLDPC_CUDA(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lk(g_side_mu);
    SideStreams &p = g_side[dev];
    if (!p.fork) {
        NRLDPC_CUDA(cudaEventCreateWithFlags(&p.fork, cudaEventDisableTiming));
        for (int i = 0; i < kSideStreams; ++i) {
            NRLDPC_CUDA(cudaStreamCreateWithFlags(&p.s[i], cudaStreamNonBlocking));
            NRLDPC_CUDA(cudaEventCreateWithFlags(&p.join[i], cudaEventDisableTiming));
        }
    }
    const int used = std::min(ngroups, kSideStreams);
    NRLDPC_CUDA(cudaEventRecord(p.fork, caller));
    for (int i = 0; i < used; ++i) NRLDPC_CUDA(cudaStreamWaitEvent(p.s[i], p.fork, 0));
    int rc = NRLDPC_OK;
    for (int g = 0; g < ngroups && rc == NRLDPC_OK; ++g) rc = launch_group(g, p.s[g % kSideStreams]);
    for (int i = 0; i < used; ++i) {  // always join, also after a failed launch
        cudaEventRecord(p.join[i], p.s[i]);
        cudaStreamWaitEvent(caller, p.join[i], 0);
    }
    return rc;
}
}  // namespace


// IEEE half -> float, 8 values per thread (the values are exact in fp32: the decoder sees what the caller stored)
__global__ void f16_to_f32_kernel(const uint4 *__restrict__ in, float4 *__restrict__ out, size_t n8, const __half *tail_in,
                                  float *tail_out, int ntail)
{
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n8; i += (size_t)gridDim.x * blockDim.x) {
        const uint4 v = in[i];
        const __half2 *h = reinterpret_cast<const __half2 *>(&v);
        const float2 a = __half22float2(h[0]), b = __half22float2(h[1]), c = __half22float2(h[2]), d = __half22float2(h[3]);
        out[2 * i] = make_float4(a.x, a.y, b.x, b.y);
        out[2 * i + 1] = make_float4(c.x, c.y, d.x, d.y);
    }
    if (blockIdx.x == 0 && (int)threadIdx.x < ntail) tail_out[threadIdx.x] = __half2float(tail_in[threadIdx.x]);
}

}  // namespace nrldpc

using namespace nrldpc;

extern "C" {

int nrldpc_version(void) { return 100; }
const char *nrldpc_last_error(void) { return g_err; }

int nrldpc_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

int nrldpc_find_ils(int Zc) { return find_ils(Zc); }

int nrldpc_dims(int bgn, int Zc, int *K, int *N, int *Nfull, int *M)
{
    const QcCfg *c = get_cfg(bgn, Zc);
    if (!c) return NRLDPC_EINVAL;
    if (K) *K = c->K;
    if (N) *N = c->N;
    if (Nfull) *Nfull = c->Nfull;
    if (M) *M = c->M;
    return NRLDPC_OK;
}

int nrldpc_build_csr(int bgn, int Zc, int32_t *rowptr, int32_t *colidx)
{
    if (!get_cfg(bgn, Zc) || !rowptr || !colidx) return NRLDPC_EINVAL;
    return build_csr(bgn, Zc, rowptr, colidx);
}

int nrldpc_encode(int8_t *d_ck, int B, int bgn, int Zc, int fix_fillers, int8_t *d_dn, void *stream)
{
    const QcCfg *c = get_cfg(bgn, Zc);
    if (!c) return NRLDPC_EINVAL;
    if (B < 0 || !d_ck || !d_dn) { set_error("encode: bad argument"); return NRLDPC_EINVAL; }
    return launch_encode(*c, d_ck, B, fix_fillers, d_dn, (cudaStream_t)stream);
}

int nrldpc_encode_packed(const uint32_t *d_ck_words, int B, int bgn, int Zc, uint32_t *d_dn_words, void *stream)
{
    const QcCfg *c = get_cfg(bgn, Zc);
    if (!c) return NRLDPC_EINVAL;
    if (B < 0 || !d_ck_words || !d_dn_words) { set_error("encode_packed: bad argument"); return NRLDPC_EINVAL; }
    return launch_encode_packed(*c, d_ck_words, B, d_dn_words, (cudaStream_t)stream);
}

int nrldpc_encode_host(int8_t *ck, int B, int bgn, int Zc, int fix_fillers, int8_t *dn)
{
    const QcCfg *c = get_cfg(bgn, Zc);
    if (!c) return NRLDPC_EINVAL;
    if (B < 0 || !ck || !dn) { set_error("encode: bad argument"); return NRLDPC_EINVAL; }
    if (B == 0) return NRLDPC_OK;
    DevBuf d_ck, d_dn;
    NRLDPC_CUDA(d_ck.alloc((size_t)B * c->K));
    NRLDPC_CUDA(d_dn.alloc((size_t)B * c->N));
    NRLDPC_CUDA(cudaMemcpy(d_ck.p, ck, (size_t)B * c->K, cudaMemcpyHostToDevice));
    if (int rc = launch_encode(*c, d_ck.as<int8_t>(), B, fix_fillers, d_dn.as<int8_t>(), 0)) return rc;
    NRLDPC_CUDA(cudaMemcpy(dn, d_dn.p, (size_t)B * c->N, cudaMemcpyDeviceToHost));
    if (fix_fillers) NRLDPC_CUDA(cudaMemcpy(ck, d_ck.p, (size_t)B * c->K, cudaMemcpyDeviceToHost));
    return NRLDPC_OK;
}

int nrldpc_decode_minsum(const float *d_llr, int B, int bgn, int Zc, int max_iter, float alpha, float beta,
                         int early_term, int8_t *d_ck, uint32_t *d_info_packed, uint8_t *d_status, int32_t *d_iters,
                         void *stream)
{
    const QcCfg *c = get_cfg(bgn, Zc);
    if (!c) return NRLDPC_EINVAL;
    if (B < 0 || max_iter < 0 || !d_llr) { set_error("decode_minsum: bad argument"); return NRLDPC_EINVAL; }
    return launch_decode_minsum(*c, d_llr, B, max_iter, alpha, beta, early_term, d_ck, d_info_packed, d_status,
                                d_iters, (cudaStream_t)stream);
}

int nrldpc_decode_minsum_groups(int ngroups, const float *const *d_llr, const int *B, const int *bgn, const int *Zc,
                                int max_iter, float alpha, float beta, int early_term, int8_t *const *d_ck,
                                uint32_t *const *d_info_packed, uint8_t *const *d_status, int32_t *const *d_iters,
                                void *stream)
{
    if (ngroups < 0 || max_iter < 0 || (ngroups && (!d_llr || !B || !bgn || !Zc))) {
        set_error("decode_minsum_groups: bad argument");
        return NRLDPC_EINVAL;
    }
    std::vector<const QcCfg *> cfg(ngroups);
    for (int g = 0; g < ngroups; ++g) {
        if (!(cfg[g] = get_cfg(bgn[g], Zc[g]))) return NRLDPC_EINVAL;
        if (B[g] < 0 || (B[g] && !d_llr[g])) { set_error("decode_minsum_groups: bad group %d", g); return NRLDPC_EINVAL; }
    }
    return fork_join((cudaStream_t)stream, ngroups, [&](int g, cudaStream_t s) {
        return launch_decode_minsum(*cfg[g], d_llr[g], B[g], max_iter, alpha, beta, early_term, d_ck ? d_ck[g] : nullptr,
                                    d_info_packed ? d_info_packed[g] : nullptr, d_status ? d_status[g] : nullptr,
                                    d_iters ? d_iters[g] : nullptr, s);
    });
}

int nrldpc_encode_groups(int ngroups, int8_t *const *d_ck, const int *B, const int *bgn, const int *Zc, int fix_fillers,
                         int8_t *const *d_dn, void *stream)
{
    if (ngroups < 0 || (ngroups && (!d_ck || !d_dn || !B || !bgn || !Zc))) {
        set_error("encode_groups: bad argument");
        return NRLDPC_EINVAL;
    }
    std::vector<const QcCfg *> cfg(ngroups);
    for (int g = 0; g < ngroups; ++g) {
        if (!(cfg[g] = get_cfg(bgn[g], Zc[g]))) return NRLDPC_EINVAL;
        if (B[g] < 0 || (B[g] && (!d_ck[g] || !d_dn[g]))) { set_error("encode_groups: bad group %d", g); return NRLDPC_EINVAL; }
    }
    return fork_join((cudaStream_t)stream, ngroups, [&](int g, cudaStream_t s) {
        return launch_encode(*cfg[g], d_ck[g], B[g], fix_fillers, d_dn[g], s);
    });
}

int nrldpc_decode_minsum_geometry(int bgn, int Zc, int *cbs_per_cta, int *threads, int *smem_bytes)
{
    const QcCfg *c = get_cfg(bgn, Zc);
    if (!c) return NRLDPC_EINVAL;
    return decode_minsum_geometry(*c, cbs_per_cta, threads, smem_bytes);
}

// Host-buffer entry point: a ring of kHostStages chunks so that the H2D copy of chunk i+1 and the D2H copy of chunk
// i-1 overlap the decode of chunk i.  The stages' device buffers, pinned result mirrors and streams are created once
// per device and reused (grow-only) by later calls; every device has its own ring and lock.  The path is bound by the
// host link (101 KB of LLRs per codeblock), so the chunks are kept small: the only exposed time is the first chunk's
// copy and the last chunk's decode.  Pinned caller memory goes straight to the DMA engine; pageable caller memory
// (what a NumPy caller hands over) is staged through pinned slots by copy threads (h2d_async), and results destined
// for pageable memory land in the stage's pinned mirror first and are copied out when the stage is next touched --
// a cudaMemcpyAsync to / from pageable memory would block the host until the stage's kernel is done and serialise
// the ring.
namespace {
struct HostStage {
    void *buf[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};  // device: llr, ck, info, status, iters, half-precision llr
    size_t cap[6] = {0, 0, 0, 0, 0, 0};
    void *mir[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};  // pinned host mirrors of the four result buffers
    size_t mcap[5] = {0, 0, 0, 0, 0};
    cudaStream_t s = nullptr;
    // results of the chunk in flight that still have to be copied out of the mirrors
    bool busy = false;
    struct Out { void *dst; int which; size_t bytes; } out[4];
    int nout = 0;
    int ensure(int which, size_t bytes)
    {
        if (bytes <= cap[which]) return NRLDPC_OK;
        if (buf[which]) { NRLDPC_CUDA(cudaFree(buf[which])); buf[which] = nullptr; cap[which] = 0; }
        NRLDPC_CUDA(cudaMalloc(&buf[which], bytes));
        cap[which] = bytes;
        return NRLDPC_OK;
    }
    int ensure_mirror(int which, size_t bytes)
    {
        if (bytes <= mcap[which]) return NRLDPC_OK;
        if (mir[which]) { NRLDPC_CUDA(cudaFreeHost(mir[which])); mir[which] = nullptr; mcap[which] = 0; }
        NRLDPC_CUDA(cudaHostAlloc(&mir[which], bytes, cudaHostAllocPortable));
        mcap[which] = bytes;
        return NRLDPC_OK;
    }
    // wait for the chunk in flight and hand its mirrored results to the caller
    int drain()
    {
        if (!busy) return NRLDPC_OK;
        busy = false;
        NRLDPC_CUDA(cudaStreamSynchronize(s));
        for (int i = 0; i < nout; ++i) host_copy(out[i].dst, mir[out[i].which], out[i].bytes);
        nout = 0;
        return NRLDPC_OK;
    }
};
constexpr int kHostStages = 3;
constexpr size_t kHostChunkBytes = (size_t)32 << 20;
struct HostPipe {
    std::mutex mu;
    std::array<HostStage, kHostStages> st;
};
HostPipe *host_pipe()
{
    static std::mutex mu;
    static std::map<int, HostPipe *> per_device;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return nullptr;
    std::lock_guard<std::mutex> lk(mu);
    auto it = per_device.find(dev);
    if (it == per_device.end()) it = per_device.emplace(dev, new HostPipe).first;
    return it->second;
}
bool host_pinned(const void *p)
{
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return at.type == cudaMemoryTypeHost || at.type == cudaMemoryTypeManaged;
}

}  // namespace

static int decode_minsum_host_impl(const void *llr, bool f16, int B, int bgn, int Zc, int max_iter, float alpha, float beta,
                                   int early_term, int8_t *ck, uint32_t *info_packed, uint8_t *status, int32_t *iters);

int nrldpc_decode_minsum_host(const float *llr, int B, int bgn, int Zc, int max_iter, float alpha, float beta,
                              int early_term, int8_t *ck, uint32_t *info_packed, uint8_t *status, int32_t *iters)
{
    return decode_minsum_host_impl(llr, false, B, bgn, Zc, max_iter, alpha, beta, early_term, ck, info_packed, status, iters);
}

int nrldpc_decode_minsum_host_f16(const uint16_t *llr_f16, int B, int bgn, int Zc, int max_iter, float alpha, float beta,
                                  int early_term, int8_t *ck, uint32_t *info_packed, uint8_t *status, int32_t *iters)
{
    return decode_minsum_host_impl(llr_f16, true, B, bgn, Zc, max_iter, alpha, beta, early_term, ck, info_packed, status, iters);
}

static int decode_minsum_host_impl(const void *llr, bool f16, int B, int bgn, int Zc, int max_iter, float alpha, float beta,
                                   int early_term, int8_t *ck, uint32_t *info_packed, uint8_t *status, int32_t *iters)
{
    const QcCfg *c = get_cfg(bgn, Zc);
    if (!c) return NRLDPC_EINVAL;
    if (B < 0 || max_iter < 0 || !llr) { set_error("decode_minsum: bad argument"); return NRLDPC_EINVAL; }
    if (B == 0) return NRLDPC_OK;
    HostPipe *pipe = host_pipe();
    if (!pipe) return cuda_fail(cudaGetLastError(), "cudaGetDevice");
    std::lock_guard<std::mutex> lk(pipe->mu);
    HostStage *st = pipe->st.data();
    const size_t llr_bytes = (size_t)c->N * 4, nwords = (size_t)(c->K + 31) / 32;
    int chunk = (int)std::max<size_t>(1, kHostChunkBytes / llr_bytes);
    if (chunk > B) chunk = (B + 1) / 2 > 256 ? (B + 1) / 2 : B;  // two chunks still overlap copy and compute
    const int nstage = std::min(kHostStages, (B + chunk - 1) / chunk);
    // where the four results go: {caller pointer, bytes per codeblock, pinned?}
    struct Res { char *p; size_t per; bool pinned; } res[5] = {
        {nullptr, 0, true}, {(char *)ck, (size_t)c->Nfull, false}, {(char *)info_packed, nwords * 4, false},
        {(char *)status, 1, false}, {(char *)iters, 4, false}};
    for (int w = 1; w < 5; ++w) if (res[w].p) res[w].pinned = host_pinned(res[w].p);
    int rc = NRLDPC_OK;
    auto fail = [&](int code) {  // no copy may still be in flight into the caller's buffers when we return
        for (int i = 0; i < nstage; ++i) { if (st[i].s) cudaStreamSynchronize(st[i].s); st[i].busy = false; st[i].nout = 0; }
        return code;
    };
    for (int i = 0; i < nstage && rc == NRLDPC_OK; ++i) {
        if (!st[i].s && cudaStreamCreateWithFlags(&st[i].s, cudaStreamNonBlocking) != cudaSuccess) rc = cuda_fail(cudaGetLastError(), "cudaStreamCreate");
        if (rc == NRLDPC_OK) rc = st[i].ensure(0, (size_t)chunk * llr_bytes);
        if (rc == NRLDPC_OK && f16) rc = st[i].ensure(5, (size_t)chunk * llr_bytes / 2);
        for (int w = 1; w < 5 && rc == NRLDPC_OK; ++w) {
            if (w <= 2 && !res[w].p) continue;  // ck / info are optional kernel outputs; status and iters are always produced
            rc = st[i].ensure(w, (size_t)chunk * res[w].per);
            if (rc == NRLDPC_OK && res[w].p && !res[w].pinned) rc = st[i].ensure_mirror(w, (size_t)chunk * res[w].per);
        }
    }
    if (rc != NRLDPC_OK) return fail(rc);
    // Chunk schedule.  The path is bound by the host link, so what is exposed is the first chunk's copy (nothing to decode
    // yet) and the last chunk's decode (nothing left to copy): both are kept to one wave of codeblocks (one per SM) when the
    // batch is large enough, and the chunks in between are whole waves.
    int sms = 0, dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = 148;
    int edge = 0;  // codeblocks of the first and of the last chunk (0: uniform chunks)
    if (chunk >= 2 * sms && B >= 4 * chunk) { edge = sms; chunk -= chunk % sms; }
    int k = 0;
    for (int b0 = 0, nb = 0; b0 < B; b0 += nb, ++k) {
        HostStage &S = st[k % nstage];
        nb = std::min(chunk, B - b0);
        if (edge) {
            if (b0 == 0) nb = edge;
            else if (B - b0 > edge) nb = std::min(chunk, B - b0 - edge);
        }
        if ((rc = S.drain()) != NRLDPC_OK) return fail(rc);
        if (!f16) {
            if ((rc = h2d_async(S.buf[0], (const float *)llr + (size_t)b0 * c->N, (size_t)nb * llr_bytes, S.s)) != NRLDPC_OK) return fail(rc);
        } else {  // half the bytes over the host link, widened on the device (one HBM pass, ~1 % of the decode time)
            const size_t n = (size_t)nb * c->N;
            if ((rc = h2d_async(S.buf[5], (const uint16_t *)llr + (size_t)b0 * c->N, n * 2, S.s)) != NRLDPC_OK) return fail(rc);
            const size_t n8 = n / 8;
            const int grid = (int)std::min<size_t>(148 * 8, std::max<size_t>(1, (n8 + 255) / 256));
            f16_to_f32_kernel<<<grid, 256, 0, S.s>>>((const uint4 *)S.buf[5], (float4 *)S.buf[0], n8, (const __half *)S.buf[5] + n8 * 8,
                                                    (float *)S.buf[0] + n8 * 8, (int)(n - n8 * 8));
            if (cudaGetLastError() != cudaSuccess) return fail(cuda_fail(cudaPeekAtLastError(), "f16_to_f32_kernel"));
        }
        if ((rc = launch_decode_minsum(*c, (const float *)S.buf[0], nb, max_iter, alpha, beta, early_term,
                                       ck ? (int8_t *)S.buf[1] : nullptr, info_packed ? (uint32_t *)S.buf[2] : nullptr,
                                       (uint8_t *)S.buf[3], (int32_t *)S.buf[4], S.s)) != NRLDPC_OK)
            return fail(rc);
        for (int w = 1; w < 5; ++w) {
            if (!res[w].p) continue;
            const size_t bytes = (size_t)nb * res[w].per;
            char *dst = res[w].p + (size_t)b0 * res[w].per;
            cudaError_t e = cudaMemcpyAsync(res[w].pinned ? (void *)dst : S.mir[w], S.buf[w], bytes, cudaMemcpyDeviceToHost, S.s);
            if (e != cudaSuccess) return fail(cuda_fail(e, "cudaMemcpyAsync(D2H)"));
            if (!res[w].pinned) S.out[S.nout++] = {dst, w, bytes};
        }
        S.busy = true;
    }
    for (int i = 0; i < nstage; ++i)
        if ((rc = st[i].drain()) != NRLDPC_OK) return fail(rc);
    return NRLDPC_OK;
}

int nrldpc_decode_csr_host(const void *llr, int is_f64, int B, int M, int Nv, const int32_t *rowptr,
                           const int32_t *colidx, int max_iter, int algo, double alpha, double beta, int early_term,
                           int8_t *ck, uint8_t *status, int32_t *iters)
{
    if (is_f64)
        return soft_csr_host<double>((const double *)llr, B, M, Nv, rowptr, colidx, 0, max_iter, algo, alpha, beta,
                                     early_term, ck, status, iters);
    return soft_csr_host<float>((const float *)llr, B, M, Nv, rowptr, colidx, 0, max_iter, algo, alpha, beta,
                                early_term, ck, status, iters);
}

int nrldpc_decode_soft_ref_host(const void *llr, int is_f64, int B, int bgn, int Zc, int max_iter, int algo,
                                double alpha, double beta, int early_term, int8_t *ck, uint8_t *status, int32_t *iters)
{
    const QcCfg *c = get_cfg(bgn, Zc);
    if (!c) return NRLDPC_EINVAL;
    const int nnz = (bgn == 1 ? 316 : 197) * Zc;
    std::vector<int32_t> rowptr(c->M + 1), colidx(nnz);
    build_csr(bgn, Zc, rowptr.data(), colidx.data());
    if (is_f64)
        return soft_csr_host<double>((const double *)llr, B, c->M, c->Nfull, rowptr.data(), colidx.data(), 2 * Zc,
                                     max_iter, algo, alpha, beta, early_term, ck, status, iters);
    return soft_csr_host<float>((const float *)llr, B, c->M, c->Nfull, rowptr.data(), colidx.data(), 2 * Zc, max_iter,
                                algo, alpha, beta, early_term, ck, status, iters);
}

int nrldpc_decode_bf_csr_host(const double *llr, int B, int M, int Nv, const int32_t *rowptr, const int32_t *colidx,
                              int max_iter, int8_t *ck, uint8_t *status, int32_t *iters)
{
    return bf_csr_host(llr, B, M, Nv, rowptr, colidx, 0, max_iter, ck, status, iters);
}

int nrldpc_decode_bf(const void *d_llr, int is_f64, int B, int bgn, int Zc, int max_iter, int8_t *d_ck,
                     uint8_t *d_status, int32_t *d_iters, void *stream)
{
    const QcCfg *c = get_cfg(bgn, Zc);
    if (!c) return NRLDPC_EINVAL;
    if (B < 0 || max_iter < 0 || !d_llr || !d_ck) { set_error("decode_bf: bad argument"); return NRLDPC_EINVAL; }
    return launch_bf_qc(*c, d_llr, is_f64, B, max_iter, d_ck, d_status, d_iters, (cudaStream_t)stream);
}

int nrldpc_decode_bf_host(const double *llr, int B, int bgn, int Zc, int max_iter, int8_t *ck, uint8_t *status,
                          int32_t *iters)
{
    const QcCfg *c = get_cfg(bgn, Zc);
    if (!c) return NRLDPC_EINVAL;
    if (B < 0 || max_iter < 0 || !llr || !ck) { set_error("decode_bf: bad argument"); return NRLDPC_EINVAL; }
    if (B == 0) return NRLDPC_OK;
    // chunks of <= 256 MB of LLRs on this thread's stream: staged copy in, kernel, one copy of the decisions and one of
    // {iters, status} out per chunk (a per-codeblock caller pays one synchronisation, not four blocking copies)
    cudaStream_t s;
    if (int rc = host_stream(&s)) return rc;
    const size_t per_cb = (size_t)c->N * 8;
    const int chunk = (int)std::max<size_t>(1, std::min<size_t>((size_t)B, ((size_t)256 << 20) / per_cb));
    DevBuf d_llr, d_ck, d_small;
    NRLDPC_CUDA(d_llr.alloc((size_t)chunk * per_cb, s));
    NRLDPC_CUDA(d_ck.alloc((size_t)chunk * c->Nfull, s));
    NRLDPC_CUDA(d_small.alloc((size_t)chunk * 5, s));   // iters int32 [chunk], then status uint8 [chunk]
    std::vector<uint8_t> h_small((size_t)chunk * 5);
    for (int b0 = 0; b0 < B; b0 += chunk) {
        const int nb = std::min(chunk, B - b0);
        int32_t *d_it = d_small.as<int32_t>();
        uint8_t *d_st = d_small.as<uint8_t>() + (size_t)chunk * 4;
        int rc = h2d_async(d_llr.p, llr + (size_t)b0 * c->N, (size_t)nb * per_cb, s);
        if (rc == NRLDPC_OK) rc = launch_bf_qc(*c, d_llr.p, 1, nb, max_iter, d_ck.as<int8_t>(), d_st, d_it, s);
        if (rc == NRLDPC_OK && (status || iters) &&
            cudaMemcpyAsync(h_small.data(), d_small.p, (size_t)chunk * 5, cudaMemcpyDeviceToHost, s) != cudaSuccess)
            rc = cuda_fail(cudaGetLastError(), "cudaMemcpyAsync(D2H)");
        if (rc == NRLDPC_OK) rc = d2h_sync(ck + (size_t)b0 * c->Nfull, d_ck.p, (size_t)nb * c->Nfull, s);  // synchronises the stream
        else cudaStreamSynchronize(s);
        if (rc != NRLDPC_OK) return rc;
        if (iters) memcpy(iters + b0, h_small.data(), (size_t)nb * 4);
        if (status) memcpy(status + b0, h_small.data() + (size_t)chunk * 4, (size_t)nb);
    }
    return NRLDPC_OK;
}

}  // extern "C"
