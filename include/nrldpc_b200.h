/*
 * nrldpc_b200.h -- C ABI of libnrldpc_b200.so: batched 5G NR LDPC (TS 38.212 5.3.2) on B200 (sm_100a).
 *
 * This is the drop-in boundary for the LDPC hot path of xu753x/python_5gtoolbox.  The reference has
 * no FFI (it is pure Python); each entry point below names the reference function it replaces
 * (file:line under the reference root).  INTEGRATION.md shows the ctypes binding a maintainer adds.
 *
 * Conventions
 *   - extern "C", plain pointers and sizes; every function returns 0 on success, a negative
 *     NRLDPC_E* code on failure (never throws); nrldpc_last_error() gives the text for this thread.
 *   - `_host` functions take HOST buffers and are synchronous (H2D, kernels, D2H inside).
 *   - the others take DEVICE buffers plus a cudaStream_t passed as void* (NULL = default stream) and
 *     are asynchronous on that stream; the caller owns every buffer.
 *   - batches are row-major [B, len]; one (bgn, Zc) per call (a transport block has one Zc,
 *     py5gphy/ldpc/ldpc_info.py:62-69; mixed-Zc workloads issue one call per group/stream).
 *   - bit arrays are one int8 per bit (the reference's dtype) unless the name says `packed`.
 *   - there is no CPU fallback anywhere: without a CUDA device every compute entry point fails.
 */
#ifndef NRLDPC_B200_H
#define NRLDPC_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NRLDPC_OK 0
#define NRLDPC_EINVAL (-1)  /* bad argument (the reference would raise AssertionError) */
#define NRLDPC_ECUDA (-2)   /* CUDA runtime error, see nrldpc_last_error() */
#define NRLDPC_ENOMEM (-3)
#define NRLDPC_ENODEV (-4)  /* no CUDA device */

/* algo ids for the soft decoders */
#define NRLDPC_ALGO_MINSUM 0 /* 'min-sum' family: alpha=1,beta=0 plain; alpha<1 NMS; beta>0 OMS; both mixed */
#define NRLDPC_ALGO_BP 1     /* 'BP' (sum-product) -- generic kernels only */

int nrldpc_version(void);
const char *nrldpc_last_error(void);
int nrldpc_device_count(void);

/* ldpc_info.find_iLS (py5gphy/ldpc/ldpc_info.py:81-97): set index 0..7, or 255 for an invalid Zc. */
int nrldpc_find_ils(int Zc);

/* K, N, N' = N + 2Zc and M for (bgn, Zc) (py5gphy/ldpc/nr_ldpc_decode.py:26-31); any pointer may be NULL. */
int nrldpc_dims(int bgn, int Zc, int *K, int *N, int *Nfull, int *M);

/*
 * ldpc_info.getH in sparse form (py5gphy/ldpc/ldpc_info.py:99-139): CSR of the lifted parity-check
 * matrix, rows and columns ascending.  Host buffers: rowptr[M+1], colidx[nnz*Zc].  Returns #edges.
 */
int nrldpc_build_csr(int bgn, int Zc, int32_t *rowptr, int32_t *colidx);

/* ------------------------------------------------------------------ encoder */
/*
 * nr_ldpc_encode.encode_ldpc(ck, bgn) for B codeblocks (py5gphy/ldpc/nr_ldpc_encode.py:8-115).
 *   ck [B,K] int8 in {0,1,-1}; -1 = filler.  If fix_fillers != 0 the fillers at k >= 2Zc are
 *      overwritten with 0 in place, which is the reference's side effect (:32-35).
 *   dn [B,N] int8 in {0,1,-1}; -1 at the filler positions (:31-37).
 */
int nrldpc_encode(int8_t *d_ck, int B, int bgn, int Zc, int fix_fillers, int8_t *d_dn, void *stream);
int nrldpc_encode_host(int8_t *ck, int B, int bgn, int Zc, int fix_fillers, int8_t *dn);
/*
 * The same encoder on bit-packed codeblocks -- SURVEY 8(d)'s algorithmic K/8 + N/8 bytes per codeblock instead of the
 * reference's byte per bit (nr_ldpc_encode.py:8-50 without the filler handling of :32-37: a packed bit cannot be -1).
 *   ck_words [B, K/32] uint32, bit k of a codeblock at word k / 32, bit k % 32 (the layout of info_packed below);
 *   dn_words [B, N/32] uint32, same layout.  Zc must be a multiple of 32 and both arrays 16-byte aligned (NRLDPC_EINVAL otherwise).
 * Used by the device-resident Monte-Carlo chain (bits -> CRC -> encode -> AWGN never leave the GPU).
 */
int nrldpc_encode_packed(const uint32_t *d_ck_words, int B, int bgn, int Zc, uint32_t *d_dn_words, void *stream);

/* ------------------------------------------------------------------ min-sum decoder (hot path) */
/*
 * nr_ldpc_decode.nr_decode_ldpc(LLRin, Zc, bgn, L, 'min-sum', alpha, beta) for B codeblocks
 * (py5gphy/ldpc/nr_ldpc_decode.py:11-49 -> decode_ldpc :51-143 -> _min_sum_process :178-227):
 * flooding schedule, fp32 arithmetic in the reference's operation order, LLR > 0 <=> bit 0,
 * the 2Zc punctured columns start at LLR 0.
 *   llr         [B,N] float32 channel LLRs
 *   max_iter    L
 *   early_term  1 = reference semantics (syndrome check before every iteration, :107-114);
 *               0 = always run exactly max_iter iterations (throughput mode, not in the reference)
 *   ck          [B,N'] int8 hard decisions, or NULL            (blkandcrc = ck[:, 0:K], :47)
 *   info_packed [B, ceil(K/32)] uint32, bit k of the codeblock at word k/32 bit k%32, or NULL
 *   status      [B] uint8 1 = all parity checks satisfied (:114,:140-143), or NULL
 *   iters       [B] int32 number of check-node passes executed, or NULL
 */
int nrldpc_decode_minsum(const float *d_llr, int B, int bgn, int Zc, int max_iter, float alpha, float beta,
                         int early_term, int8_t *d_ck, uint32_t *d_info_packed, uint8_t *d_status,
                         int32_t *d_iters, void *stream);
int nrldpc_decode_minsum_host(const float *llr, int B, int bgn, int Zc, int max_iter, float alpha, float beta,
                              int early_term, int8_t *ck, uint32_t *info_packed, uint8_t *status, int32_t *iters);
/*
 * The same with the channel LLRs stored as IEEE half-precision floats (the format an 8- to 11-bit soft demapper fills
 * without loss): half the bytes cross the host link -- which is what bounds the host-buffer path at R = 1/3 -- and are
 * widened to fp32 on the device.  Arithmetic and results are those of nrldpc_decode_minsum_host on the widened values.
 */
int nrldpc_decode_minsum_host_f16(const uint16_t *llr_f16, int B, int bgn, int Zc, int max_iter, float alpha, float beta,
                                  int early_term, int8_t *ck, uint32_t *info_packed, uint8_t *status, int32_t *iters);

/*
 * Mixed-(bgn, Zc) batch: group g is nrldpc_decode_minsum / nrldpc_encode on B[g] codeblocks of (bgn[g], Zc[g]) with
 * its own buffers (arrays of ngroups device pointers, host-resident; d_ck / d_info_packed / d_status / d_iters may
 * be NULL as a whole).  This is what the per-codeblock loops of DLSCHDecode / ULSCH_decoding over several transport
 * blocks of one slot amount to (py5gphy/nr_pdsch/nr_dlsch_decode.py:83-98, nr_pusch/nr_ulsch_decode.py:86-92; one Zc
 * per transport block, ldpc_info.get_cbs_info :62-69).  The groups run concurrently on internal streams that fork
 * from and join back into `stream`: asynchronous to the host, ordered like a single launch on `stream`.
 */
int nrldpc_decode_minsum_groups(int ngroups, const float *const *d_llr, const int *B, const int *bgn, const int *Zc,
                                int max_iter, float alpha, float beta, int early_term, int8_t *const *d_ck,
                                uint32_t *const *d_info_packed, uint8_t *const *d_status, int32_t *const *d_iters,
                                void *stream);
int nrldpc_encode_groups(int ngroups, int8_t *const *d_ck, const int *B, const int *bgn, const int *Zc, int fix_fillers,
                         int8_t *const *d_dn, void *stream);

/* Launch geometry the hot kernel uses for (bgn, Zc): codeblocks per CTA, threads, dynamic smem bytes. */
int nrldpc_decode_minsum_geometry(int bgn, int Zc, int *cbs_per_cta, int *threads, int *smem_bytes);

/* ------------------------------------------------------------------ generic-H decoders */
/*
 * nr_ldpc_decode.decode_ldpc(LLRin, H, L, algo, alpha, beta) for an arbitrary parity-check matrix
 * given as CSR (py5gphy/ldpc/nr_ldpc_decode.py:51-143), B codeblocks, same H.
 *   llr [B,Nv] (fp32 if is_f64 == 0, else fp64 -- the fp64 mode reproduces the reference's
 *   float64 arithmetic bit for bit for algo = min-sum), ck [B,Nv] int8.
 *   Host variant only: it owns the temporary device workspace.
 */
int nrldpc_decode_csr_host(const void *llr, int is_f64, int B, int M, int Nv, const int32_t *rowptr,
                           const int32_t *colidx, int max_iter, int algo, double alpha, double beta,
                           int early_term, int8_t *ck, uint8_t *status, int32_t *iters);

/* The same kernels on the 5G matrix of (bgn, Zc): llr [B,N], 2Zc zeros prepended (:43), ck [B,N']. */
int nrldpc_decode_soft_ref_host(const void *llr, int is_f64, int B, int bgn, int Zc, int max_iter, int algo,
                                double alpha, double beta, int early_term, int8_t *ck, uint8_t *status,
                                int32_t *iters);

/*
 * ldpc_decoder_bit_flipping.ldpc_decoder_BF(LLRin, H, L) on a CSR H
 * (py5gphy/ldpc/ldpc_decoder_bit_flipping.py:5-73).  llr [B,Nv] float64 (only its sign is used),
 * ck [B,Nv] int8 0/1.
 */
int nrldpc_decode_bf_csr_host(const double *llr, int B, int M, int Nv, const int32_t *rowptr,
                              const int32_t *colidx, int max_iter, int8_t *ck, uint8_t *status, int32_t *iters);
/*
 * nr_decode_ldpc(..., algo='BF') (py5gphy/ldpc/nr_ldpc_decode.py:43,65-67) on the quasi-cyclic bit-flipping kernel
 * (state of a codeblock resident in shared memory): llr [B,N] (float32, or float64 when is_f64), ck [B,N'] int8,
 * status[b] = 1 when a zero syndrome was reached, iters[b] = the iteration index at which it was (else max_iter).
 * nrldpc_decode_bf takes device pointers and a stream; nrldpc_decode_bf_host host buffers (synchronous).
 */
int nrldpc_decode_bf(const void *d_llr, int is_f64, int B, int bgn, int Zc, int max_iter, int8_t *d_ck,
                     uint8_t *d_status, int32_t *d_iters, void *stream);
int nrldpc_decode_bf_host(const double *llr, int B, int bgn, int Zc, int max_iter, int8_t *ck, uint8_t *status,
                          int32_t *iters);

/*
 * nr_decode_ldpc(..., algo='BP') (py5gphy/ldpc/nr_ldpc_decode.py:51-143 with _BP_process :145-176, including the
 * one-zero quirk :164-170 and the +-38.14 clip :158-163) on the quasi-cyclic sum-product kernel: float64 arithmetic,
 * per-edge messages in an L2-resident workspace of persistent CTAs, posteriors in shared memory.
 * llr [B,N] float32 or float64 (is_f64), ck [B,N'] int8, status / iters as for the min-sum decoder.
 */
int nrldpc_decode_bp(const void *d_llr, int is_f64, int B, int bgn, int Zc, int max_iter, int early_term, int8_t *d_ck,
                     uint8_t *d_status, int32_t *d_iters, void *stream);
int nrldpc_decode_bp_host(const void *llr, int is_f64, int B, int bgn, int Zc, int max_iter, int early_term, int8_t *ck,
                          uint8_t *status, int32_t *iters);

/* ------------------------------------------------------------------ Monte-Carlo helpers (device) */
/*
 * BPSK + AWGN + LLR of nr_ldpc_decode.for_test_5g_ldpc_encoder (py5gphy/ldpc/nr_ldpc_decode.py:252-257)
 * with a counter-based Philox4x32-10 generator instead of NumPy's global RNG:
 *   llr = 2 * ((1 - 2 dn) + sigma * n) / sigma^2,  sigma = 10^(-snr_db/20),  n ~ N(0,1).
 *   dn [B,N] int8 (a -1 filler is sent as LLR 0), llr [B,N] float32.
 */
int nrldpc_awgn_llr(const int8_t *d_dn, long long count, float snr_db, unsigned long long seed,
                    unsigned long long offset, float *d_llr, void *stream);
/* Row form: row j of [rows, cols] draws from the counter range of id = first_id + j*id_stride, so a
 * codeblock's noise depends only on its global id -- any sharding of a Monte-Carlo point over GPUs
 * sees the same codeblocks. */
int nrldpc_awgn_llr_rows(const int8_t *d_dn, long long rows, long long cols, float snr_db, unsigned long long seed,
                         long long first_id, long long id_stride, float *d_llr, void *stream);
int nrldpc_random_bits_rows(int8_t *d_bits, long long rows, long long cols, unsigned long long seed,
                            long long first_id, long long id_stride, void *stream);
/* Uniform random bits (Philox), int8 0/1. */
int nrldpc_random_bits(int8_t *d_bits, long long count, unsigned long long seed, unsigned long long offset,
                       void *stream);
/*
 * Error counters of the Monte-Carlo drivers (scripts/internal/sim_ldpc_internal.py:61-62):
 * counters[0] += B, [1] += codeblocks whose first K decisions differ from ref, [2] += differing bits,
 * [3] += sum of iters (if iters != NULL).  ref [B,ref_stride], got [B,got_stride] int8; counters int64[4].
 */
int nrldpc_count_errors(const int8_t *d_ref, long long ref_stride, const int8_t *d_got, long long got_stride,
                        int B, int K, const int32_t *d_iters, long long *d_counters, void *stream);

/*
 * Bit-packed twins of the four helpers above (word layout of nrldpc_encode_packed; row j of a [rows, row_words] array
 * starts at word j * row_words, row_words >= ceil(cols / 32), the bits beyond `cols` are written as 0).  Same Philox
 * counters, so bit k of a row / the noise on it are those of the byte-per-bit functions for the same (seed, id):
 *   random_bits_packed_rows   cols random bits per row;
 *   crc_attach_packed         the CRC of the first A bits of every row stored at bits A .. A+L-1, in place
 *                             (crc.nr_crc_encode, py5gphy/crc/crc.py:4-41); returns L;
 *   awgn_llr_packed_rows      llr [rows, cols] float32 from packed dn (no fillers);
 *   count_errors_packed       counters as nrldpc_count_errors, ref and got packed (got = info_packed of the decoder).
 */
int nrldpc_random_bits_packed_rows(uint32_t *d_words, long long rows, long long cols, long long row_words,
                                   unsigned long long seed, long long first_id, long long id_stride, void *stream);
int nrldpc_crc_attach_packed(uint32_t *d_words, int B, int A, int poly_id, long long row_words, void *stream);
int nrldpc_awgn_llr_packed_rows(const uint32_t *d_dn_words, long long rows, long long cols, long long row_words, float snr_db,
                                unsigned long long seed, long long first_id, long long id_stride, float *d_llr, void *stream);
int nrldpc_count_errors_packed(const uint32_t *d_ref_words, long long ref_row_words, const uint32_t *d_got_words,
                               long long got_row_words, int B, int K, const int32_t *d_iters, long long *d_counters,
                               void *stream);

/* ------------------------------------------------------------------ CRC (callers' side of the path) */
/*
 * crc.nr_crc_encode(blk, poly) / crc.nr_crc_decode(blkandcrc, poly) with mask = 0
 * (py5gphy/crc/crc.py:4-41, :43-88; polynomials :96-106), B blocks of A payload bits, one int8 per bit.
 *   poly_id: 0 '6', 1 '11', 2 '16', 3 '24A', 4 '24B', 5 '24C'.  Returns the CRC length L (> 0) or an error.
 *   encode: in [B,A] -> out [B,A+L].   check: in [B,A+L] -> err [B] uint8 (1 = CRC error).
 */
int nrldpc_crc_encode(const int8_t *d_in, int B, int A, int poly_id, int8_t *d_out, void *stream);
int nrldpc_crc_check(const int8_t *d_in, int B, int A, int poly_id, uint8_t *d_err, void *stream);
int nrldpc_crc_encode_host(const int8_t *in, int B, int A, int poly_id, int8_t *out);
int nrldpc_crc_check_host(const int8_t *in, int B, int A, int poly_id, uint8_t *err);

/* ------------------------------------------------------------------ rate matching / recovery (callers' side) */
/*
 * nr_ldpc_ratematch.ratematch_ldpc(dn, Ncb, E, k0, Qm) for the B codeblocks of a transport block
 * (py5gphy/ldpc/nr_ldpc_ratematch.py:64-97) followed by code block concatenation
 * (py5gphy/nr_pdsch/nr_dlsch.py:66-68): bit selection from the circular buffer of length Ncb starting at k0,
 * skipping the -1 fillers, repeating when E exceeds the buffer, then the Qm-row bit interleaver.
 *   dn [B,N] int8 in {0,1,-1};  E [B] int32 output length of every codeblock (a multiple of Qm);
 *   goff [B] int64 offset of codeblock b inside g (host variant: the running sum of E);  g int8.
 */
int nrldpc_ratematch(const int8_t *d_dn, int B, int N, int Ncb, int k0, int Qm, const int32_t *d_E,
                     const long long *d_goff, int8_t *d_g, void *stream);
int nrldpc_ratematch_host(const int8_t *dn, int B, int N, int Ncb, int k0, int Qm, const int32_t *E, int8_t *g);

/*
 * nr_ldpc_raterecover.raterecover_ldpc(LLr_fe, Ncb, N, k0, Qm, Zc, K_apo, K) for B codeblocks
 * (py5gphy/ldpc/nr_ldpc_raterecover.py:6-65): de-interleave, put the E received LLRs back on the circular
 * buffer (repeated positions are averaged: float64 sum in arrival order / count, :41-63), fillers
 * [F0,F1) = [K_apo-2Zc, K-2Zc) get 10*max|LLr_fe| (:30,:64), everything else 0.
 *   llr_g : the concatenated received LLRs, float32 (in_f64 = 0) or float64; arithmetic is float64.
 *   out [B,N] float64 (out_f64 = 1, the reference's dtype) or float32 (what the fp32 decoder consumes).
 */
int nrldpc_raterecover(const void *d_llr_g, int in_f64, int B, int N, int Ncb, int k0, int Qm, int F0, int F1,
                       const int32_t *d_E, const long long *d_goff, void *d_out, int out_f64, void *stream);
int nrldpc_raterecover_host(const void *llr_g, int in_f64, int B, int N, int Ncb, int k0, int Qm, int Zc, int K_apo,
                            int K, const int32_t *E, void *out, int out_f64);

/*
 * HARQ soft combining of DLSCHDecode / ULSCH_decoding (py5gphy/nr_pdsch/nr_dlsch_decode.py:80-87,
 * py5gphy/nr_pusch/nr_ulsch_decode.py:81-88): out = a + c where either is 0, (a + c) / 2 elsewhere; float64.
 */
int nrldpc_harq_combine(const double *d_new, const double *d_cur, long long count, double *d_out, void *stream);
int nrldpc_harq_combine_host(const double *nw, const double *cur, long long count, double *out);

/* ------------------------------------------------------------------ whole transport blocks (callers' side, fused) */
/*
 * Pinned (page-locked, device-mapped) host memory from the library's pool: buffers the `_host` entry points can DMA
 * from / into directly, and that the decoder can store its float64 soft buffer into over PCIe while it iterates.
 * Pageable buffers are accepted everywhere; they are staged through a ring of pinned slots by copy threads.
 */
int nrldpc_host_alloc(size_t bytes, void **p);
int nrldpc_host_free(void *p);

/*
 * De-rate-matching + HARQ combining of the C codeblocks of a transport block without a decoder behind it:
 * nr_ldpc_raterecover.raterecover_ldpc (py5gphy/ldpc/nr_ldpc_raterecover.py:6-65) followed by the combining loop of
 * DLSCHDecode / ULSCH_decoding (py5gphy/nr_pdsch/nr_dlsch_decode.py:74-87, py5gphy/nr_pusch/nr_ulsch_decode.py:75-88).
 *   llr_g  concatenated received LLRs (float32, or float64 when in_f64), E [C], goff [C] offsets inside llr_g,
 *   cur    [C,N] float64 soft buffer of the earlier transmissions, or NULL (first transmission / HARQ off),
 *   soft   [C,N] float64 out (new_LLr_dns), llr32 [C,N] float32 out (the decoder's input); either may be NULL.
 */
int nrldpc_sch_recover(const void *d_llr_g, int in_f64, int C, int N, int Ncb, int k0, int Qm, int F0, int F1,
                       const int32_t *d_E, const long long *d_goff, const double *d_cur, double *d_soft, float *d_llr32,
                       void *stream);
int nrldpc_sch_recover_host(const void *llr_g, int in_f64, int C, int N, int Ncb, int k0, int Qm, int Zc, int K_apo, int K,
                            const int32_t *E, const double *cur, double *soft);

/*
 * DLSCHDecode / ULSCH_decoding with algo = 'min-sum' after the parameter arithmetic
 * (py5gphy/nr_pdsch/nr_dlsch_decode.py:56-107, py5gphy/nr_pusch/nr_ulsch_decode.py:56-108) in two launches:
 *   1. the min-sum decoder, whose LLR load is the rate recovery and HARQ combining above (per codeblock: received
 *      sequence -> fp32 row that stays in L2 + the float64 row of `soft`), reference early termination;
 *   2. CB CRC24B of every codeblock when C > 1 (:93-98, computed and ignored by the reference), the cbz payload bits
 *      of every codeblock into the transport block (:101-102) and the TB CRC (:105).
 *   K_apo = cbz + L of ldpc_info.get_cbs_info (fillers are [K_apo, K)); A = TBSize; C * cbz = A + (24 | 16).
 *   tbblk [A] int8, tb_err[0] = 1 when the TB CRC fails (status = not tb_err), cb_err [C], status [C] = LDPC parity
 *   ok, iters [C]; ck [C,N'] = the codeblocks' hard decisions (device variant: NULL = internal workspace).
 * Host variant: synchronous; when `soft` is pinned host memory (nrldpc_host_alloc) the decoder stores into it directly.
 */
int nrldpc_sch_decode(const void *d_llr_g, int in_f64, int C, int bgn, int Zc, int Ncb, int k0, int Qm, int K_apo,
                      const int32_t *d_E, const long long *d_goff, const double *d_cur, double *d_soft, int max_iter,
                      float alpha, float beta, int A, int8_t *d_ck, int8_t *d_tbblk, uint8_t *d_tb_err, uint8_t *d_cb_err,
                      uint8_t *d_status, int32_t *d_iters, void *stream);
int nrldpc_sch_decode_host(const void *llr_g, int in_f64, int C, int bgn, int Zc, int Ncb, int k0, int Qm, int K_apo,
                           const int32_t *E, const double *cur, double *soft, int max_iter, float alpha, float beta, int A,
                           int8_t *tbblk, uint8_t *tb_err, uint8_t *cb_err, uint8_t *status, int32_t *iters);

/*
 * Transmit side.  nrldpc_sch_segment: TB CRC attachment (crc.nr_crc_encode with '24A' | '16',
 * py5gphy/nr_pdsch/nr_dlsch.py:29-35) + nr_ldpc_cbsegment.ldpc_cbsegment (py5gphy/ldpc/nr_ldpc_cbsegment.py:7-33):
 * trblk int8 [A] -> cbs int8 [C,K] with CB CRC24B when C > 1 and -1 fillers.
 * nrldpc_encode_ratematch: the per-codeblock loop of DLSCHEncode / ULSCH_encoding_ratematch (nr_dlsch.py:53-72,
 * py5gphy/nr_pusch/nr_ulsch.py:49-66): encode_ldpc, ratematch_ldpc and code block concatenation, cbs [C,K] -> g [sum E]
 * (fix_fillers: the -1 fillers of cbs become 0 in place like encode_ldpc does).
 * nrldpc_sch_encode_host: both, trblk -> g, nothing but the two ends crossing PCIe.
 */
int nrldpc_sch_segment(const int8_t *d_trblk, int A, int C, int K, int8_t *d_cbs, void *stream);
int nrldpc_sch_segment_host(const int8_t *trblk, int A, int C, int K, int8_t *cbs);
int nrldpc_encode_ratematch(int8_t *d_cbs, int C, int bgn, int Zc, int fix_fillers, int Ncb, int k0, int Qm,
                            const int32_t *d_E, const long long *d_goff, int8_t *d_g, void *stream);
int nrldpc_encode_ratematch_host(int8_t *cbs, int C, int bgn, int Zc, int fix_fillers, int Ncb, int k0, int Qm,
                                 const int32_t *E, int8_t *g);
int nrldpc_sch_encode_host(const int8_t *trblk, int A, int C, int bgn, int Zc, int Ncb, int k0, int Qm, const int32_t *E,
                           int8_t *g);

#ifdef __cplusplus
}
#endif
#endif /* NRLDPC_B200_H */
