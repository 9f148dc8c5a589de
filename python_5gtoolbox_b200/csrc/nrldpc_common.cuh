// nrldpc_common.cuh -- shared definitions of the B200 NR-LDPC kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>
#include "../../include/nrldpc_b200.h"

namespace nrldpc {

constexpr int kMaxRows = 46;    // BG1 row-blocks
constexpr int kMaxEdges = 316;  // BG1 non-null entries
constexpr int kMaxCore = 26;    // BG1 systematic + core-parity column-blocks (degree > 1)
constexpr int kMaxCoreEdges = 274;

// Quasi-cyclic description of one (bgn, Zc) parity-check matrix, passed to the kernels BY VALUE as a
// __grid_constant__ parameter so that the tables live in the constant bank (warp-uniform LDC reads).
// Semantics of an entry follow py5gphy/ldpc/ldpc_info.py:126-137: check r of row-block i touches
// variable (r + P) mod Zc of column-block j, P = V(i,j) mod Zc.
struct QcCfg {
    int bgn, Zc, iLS;
    int nrows, ncols, kb, ncore;  // 46/42, 68/52, 22/10, kb+4
    int K, N, Nfull, M;
    int tiles;    // r-tiles of 32 lanes per codeblock: ceil(Zc/32)
    int lanes;    // lanes one codeblock occupies inside a warp tile: 32, or next pow2 >= Zc when Zc < 32
    int lanes_log2;
    int per;      // codeblocks sharing one warp tile = 32 / lanes
    // CSR over row-blocks.  Edge word: col | shift << 8  (shift already reduced mod Zc).
    // In rows >= 4 the LAST edge is the degree-1 extension column kb+i (shift 0).
    uint16_t rowptr[kMaxRows + 1];
    uint32_t edge[kMaxEdges];
    // CSC over the ncore core column-blocks, ascending row-block (the reference's summation order,
    // py5gphy/ldpc/nr_ldpc_decode.py:126).  Entry: row | k << 6 | bitpos << 11 | back << 16 where
    // k = position of the edge inside its row, bitpos = deg(row)-1-k = its bit in the sign word,
    // back = (Zc - shift) mod Zc so that r = (c + back) mod Zc.
    uint16_t colptr[kMaxCore + 1];
    uint32_t centry[kMaxCoreEdges];
    // processing order of row-blocks / core column-blocks (degree descending => balanced tail)
    uint8_t cn_order[kMaxRows];
    uint8_t vn_order[kMaxCore];
    // shared-memory record layout: float2 magnitudes for every check row, then sign/index words:
    // 32-bit for row-blocks of degree > 12 (BG1 rows 0-3), 16-bit otherwise.
    uint8_t wide[kMaxRows];       // 1 if the row-block uses 32-bit sign words
    uint32_t bits_off[kMaxRows];  // byte offset of the row-block's sign words / Zc  (i.e. in units of Zc bytes)
    int bits_bytes_per_zc;        // total sign-word bytes per codeblock / Zc
};

// Host-side construction (nrldpc_tables.cu)
int build_qc_cfg(int bgn, int Zc, QcCfg *cfg);
int find_ils(int Zc);
int build_csr(int bgn, int Zc, int32_t *rowptr, int32_t *colidx);

// (bgn, Zc) -> quasi-cyclic tables, built once; nullptr (and the error text set) for an invalid pair (nrldpc_api.cu)
const QcCfg *get_cfg(int bgn, int Zc);

// host <-> device staging shared by the host-buffer entry points (nrldpc_sch.cu): pinned memory goes straight to the
// DMA engine, pageable memory through a ring of pinned slots filled by copy threads
void host_copy(void *dst, const void *src, size_t n);
int h2d_async(void *dst, const void *src, size_t n, cudaStream_t s);
int d2h_sync(void *dst, const void *src, size_t n, cudaStream_t s);
int host_stream(cudaStream_t *s);

// error plumbing (nrldpc_api.cu)
void set_error(const char *fmt, ...);
int cuda_fail(cudaError_t e, const char *what);

#define NRLDPC_CUDA(call)                                   \
    do {                                                    \
        cudaError_t e__ = (call);                           \
        if (e__ != cudaSuccess) return cuda_fail(e__, #call); \
    } while (0)

// The library's own stream-ordered memory pool on the current device (nrldpc_api.cu): keeps up to 1 GiB of freed
// scratch memory for reuse and gives the rest back at the next synchronisation.  The device's default pool, which
// the embedding application (PyTorch, ...) may use, is never touched.
cudaError_t lib_mempool(cudaMemPool_t *pool);

// Temporary device buffer: stream-ordered allocation from lib_mempool(), so a per-codeblock call of an unchanged
// reference script does not pay a cudaMalloc / cudaFree pair per buffer.  Freed on the stream it was allocated on.
struct ScratchBuf {
    void *p = nullptr;
    cudaStream_t s = nullptr;
    cudaError_t alloc(size_t n, cudaStream_t stream = nullptr)
    {
        cudaMemPool_t pool;
        cudaError_t e = lib_mempool(&pool);
        if (e != cudaSuccess) return e;
        s = stream;
        return cudaMallocFromPoolAsync(&p, n ? n : 1, pool, s);
    }
    ~ScratchBuf() { if (p) cudaFreeAsync(p, s); }
    template <typename T> T *as() { return static_cast<T *>(p); }
};

// kernel launchers (device pointers)
// Rate matching fused into the encoder's store (nrldpc_encode.cu): g != nullptr -> the encoder writes the rate-matched,
// concatenated sequence g (codeblock b: E[b] bytes at g + goff[b]) instead of dn; fillers = the dn positions [F0, F1).
struct EncRmArgs {
    int8_t *g = nullptr;
    const int32_t *E = nullptr;
    const long long *goff = nullptr;
    int Ncb = 0, k0 = 0, Qm = 1, F0 = 0, F1 = 0;
};
int launch_encode(const QcCfg &cfg, int8_t *d_ck, int B, int fix_fillers, int8_t *d_dn, cudaStream_t s,
                  const EncRmArgs *rm = nullptr);
// bit-packed codeblocks in and out (K/32 and N/32 words per codeblock, no fillers), lifting sizes that are a multiple of 32
int launch_encode_packed(const QcCfg &cfg, const uint32_t *d_ck_words, int B, uint32_t *d_dn_words, cudaStream_t s);
struct RrArgs;  // nrldpc_raterecover.cuh
// rr != nullptr: the LLR load is the rate recovery of a transport block (d_llr unused, early_term must be 1)
int launch_decode_minsum(const QcCfg &cfg, const float *d_llr, int B, int max_iter, float alpha, float beta,
                         int early_term, int8_t *d_ck, uint32_t *d_info, uint8_t *d_status, int32_t *d_iters,
                         cudaStream_t s, const RrArgs *rr = nullptr);
int decode_minsum_geometry(const QcCfg &cfg, int *G, int *threads, int *smem);

template <typename T>
int launch_soft_csr(const T *d_llr, int B, int M, int Nv, int E, const int32_t *d_rowptr, const int32_t *d_colidx,
                    const int32_t *d_cptr, const int32_t *d_cedge, int prepend, int max_iter, int algo, T alpha,
                    T beta, int early_term, T *d_work, int8_t *d_ck, uint8_t *d_status, int32_t *d_iters,
                    cudaStream_t s);
int launch_bf_csr(const double *d_llr, int B, int M, int Nv, int E, const int32_t *d_rowptr, const int32_t *d_colidx,
                  const int32_t *d_cptr, const int32_t *d_cedge, int prepend, int max_iter, int32_t *d_work,
                  int8_t *d_ck, uint8_t *d_status, int32_t *d_iters, cudaStream_t s);

int launch_bf_qc(const QcCfg &cfg, const void *d_llr, int is_f64, int B, int max_iter, int8_t *d_ck,
                 uint8_t *d_status, int32_t *d_iters, cudaStream_t s);

}  // namespace nrldpc
