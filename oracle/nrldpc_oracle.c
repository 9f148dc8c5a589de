/*
 * oracle/nrldpc_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * A plain-C, CPU restatement of the reference's LDPC hot path (xu753x/python_5gtoolbox,
 * py5gphy/ldpc).  It exists only so that tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / `--impl reference` legs have something to check the CUDA path against and
 * to time on host cores.  Nothing under python_5gtoolbox_b200/ may import, link or call it.
 *
 * Parity status: PINNED.  tests/test_oracle_golden.py checks every function here against
 *   (1) the tests/golden npz fixtures -- outputs of the unmodified reference Python run in the build
 *       container by tools/gen_golden.py (committed with its script), and
 *   (2) the known-answer vectors of SURVEY.md Appendix C (also produced by the live reference),
 * and tests/test_oracle_live.py re-checks it against /root/reference when that mount exists.
 * The MATLAB .mat vectors the reference's own (unshipped) tests used are not in the mount.
 *
 * Every function cites the reference lines it follows (paths relative to the reference root).
 * The "real" type is double (what the reference computes in) or float (what the CUDA
 * kernel computes in; used to debug the kernel bit-for-bit).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "../include/nrldpc_bg_tables.inc"

/* ------------------------------------------------------------------ tables */

/* py5gphy/ldpc/ldpc_info.py:81-97 -- lifting-size set index, 255 if Zc is not a 38.212 lifting size */
int oracle_find_ils(int Zc)
{
    static const int sets[8][8] = {
        {2, 4, 8, 16, 32, 64, 128, 256}, {3, 6, 12, 24, 48, 96, 192, 384},
        {5, 10, 20, 40, 80, 160, 320, 0}, {7, 14, 28, 56, 112, 224, 0, 0},
        {9, 18, 36, 72, 144, 288, 0, 0},  {11, 22, 44, 88, 176, 352, 0, 0},
        {13, 26, 52, 104, 208, 0, 0, 0},  {15, 30, 60, 120, 240, 0, 0, 0}};
    for (int s = 0; s < 8; ++s)
        for (int k = 0; k < 8; ++k)
            if (sets[s][k] == Zc && Zc > 0) return s;
    return 255;
}

typedef struct {
    int rows, cols, nnz, kb; /* kb = number of systematic column-blocks (22 / 10) */
    const unsigned short *rowptr;
    const unsigned char *col;
    const unsigned short *shift; /* for the chosen iLS */
} bg_t;

static int get_bg(int bgn, int iLS, bg_t *g)
{
    if (iLS < 0 || iLS > 7) return -1;
    if (bgn == 1) {
        g->rows = NRLDPC_BG1_ROWS; g->cols = NRLDPC_BG1_COLS; g->nnz = NRLDPC_BG1_NNZ; g->kb = 22;
        g->rowptr = nrldpc_bg1_rowptr; g->col = nrldpc_bg1_col; g->shift = nrldpc_bg1_shift[iLS];
    } else if (bgn == 2) {
        g->rows = NRLDPC_BG2_ROWS; g->cols = NRLDPC_BG2_COLS; g->nnz = NRLDPC_BG2_NNZ; g->kb = 10;
        g->rowptr = nrldpc_bg2_rowptr; g->col = nrldpc_bg2_col; g->shift = nrldpc_bg2_shift[iLS];
    } else return -1;
    return 0;
}

/* Base-graph entry V(i,j) or -1.  py5gphy/ldpc/ldpc_info.py:124-127 (`BG[i,j] > -1`, `Pij = BG[i,j] % Zc`). */
static int bg_entry(const bg_t *g, int i, int j)
{
    for (int e = g->rowptr[i]; e < g->rowptr[i + 1]; ++e)
        if (g->col[e] == j) return g->shift[e];
    return -1;
}

/*
 * CSR form of H = getH(Zc, bgn, iLS): py5gphy/ldpc/ldpc_info.py:99-139.
 * Row m = i*Zc + r of block (i,j) has its single 1 at column j*Zc + (r + Pij) % Zc (:129-134).
 * Columns inside a row come out ascending, like np.where(H[m,:]==1) in nr_ldpc_decode.py:80-83.
 * Returns the number of edges; colidx must hold nnz*Zc ints, rowptr rows*Zc+1.
 */
int oracle_build_csr(int bgn, int Zc, int *rowptr, int *colidx)
{
    bg_t g;
    int iLS = oracle_find_ils(Zc);
    if (iLS > 7 || get_bg(bgn, iLS, &g)) return -1;
    int e = 0;
    for (int i = 0; i < g.rows; ++i)
        for (int r = 0; r < Zc; ++r) {
            rowptr[i * Zc + r] = e;
            for (int k = g.rowptr[i]; k < g.rowptr[i + 1]; ++k) {
                int P = g.shift[k] % Zc;
                colidx[e++] = g.col[k] * Zc + (r + P) % Zc;
            }
        }
    rowptr[g.rows * Zc] = e;
    return e;
}

/* Dense int8 H exactly as ldpc_info.getH returns it (for tests of the getH mirror). */
int oracle_get_h_dense(int bgn, int Zc, int8_t *H)
{
    bg_t g;
    int iLS = oracle_find_ils(Zc);
    if (iLS > 7 || get_bg(bgn, iLS, &g)) return -1;
    size_t ncol = (size_t)g.cols * Zc;
    memset(H, 0, (size_t)g.rows * Zc * ncol);
    for (int i = 0; i < g.rows; ++i)
        for (int k = g.rowptr[i]; k < g.rowptr[i + 1]; ++k) {
            int P = g.shift[k] % Zc;
            for (int r = 0; r < Zc; ++r) H[((size_t)i * Zc + r) * ncol + g.col[k] * Zc + (r + P) % Zc] = 1;
        }
    return 0;
}

/* ------------------------------------------------------------------ encoder */

/* y[r] ^= x[(r + P) % Zc]  == (circulant block with shift P) @ x over GF(2) */
static void xor_rot(int8_t *y, const int8_t *x, int P, int Zc)
{
    for (int r = 0; r < Zc; ++r) y[r] ^= x[(r + P) % Zc];
}

/*
 * encode_ldpc(ck, bgn): py5gphy/ldpc/nr_ldpc_encode.py:8-50 with _gen_ldpc_parity_bit :52-115
 * (the choose_opt==1 branch; :56-80 is dead code).
 *   ck : K values in {0,1,-1}; -1 = filler.  MUTATED like the reference (:32-35): fillers at
 *        k in [2Zc,K) become 0.
 *   dn : N values in {0,1,-1}; -1 where the input held a filler (:31-37).
 */
int oracle_encode(int8_t *ck, int bgn, int Zc, int8_t *dn)
{
    bg_t g;
    int iLS = oracle_find_ils(Zc);
    if (iLS > 7 || get_bg(bgn, iLS, &g)) return -1;
    const int K = g.kb * Zc, N = (g.cols - 2) * Zc;
    const int ncw = g.kb + 4; /* systematic + 4 core-parity column-blocks */

    memset(dn, -1, (size_t)N);
    for (int k = 2 * Zc; k < K; ++k) {
        if (ck[k] == -1) ck[k] = 0;
        else dn[k - 2 * Zc] = ck[k];
    }

    int8_t *cw = (int8_t *)calloc((size_t)g.cols * Zc, 1); /* [ck ; pc ; pe] */
    int8_t *L1 = (int8_t *)calloc((size_t)4 * Zc, 1);
    int8_t *L2 = (int8_t *)calloc((size_t)Zc, 1);
    memcpy(cw, ck, (size_t)K);

    /* :92-94  L1 = A @ ck reshaped (4,Zc);  L2 = sum_i L1[i] mod 2 */
    for (int i = 0; i < 4; ++i)
        for (int e = g.rowptr[i]; e < g.rowptr[i + 1]; ++e)
            if (g.col[e] < g.kb) xor_rot(L1 + i * Zc, cw + g.col[e] * Zc, g.shift[e] % Zc, Zc);
    for (int i = 0; i < 4; ++i)
        for (int r = 0; r < Zc; ++r) L2[r] ^= L1[i * Zc + r];

    int8_t *pc1 = cw + (g.kb + 0) * Zc, *pc2 = cw + (g.kb + 1) * Zc;
    int8_t *pc3 = cw + (g.kb + 2) * Zc, *pc4 = cw + (g.kb + 3) * Zc;
    /* :95-96 / :101-102  zeros_idx = column of the 1 in row 0 of block B(1,0) (BG1) or B(2,0) (BG2);
     * pc1 = np.roll(L2, zeros_idx), i.e. pc1[r] = L2[(r - zeros_idx) mod Zc]. */
    int s = bg_entry(&g, bgn == 1 ? 1 : 2, g.kb) % Zc;
    for (int r = 0; r < Zc; ++r) pc1[r] = L2[((r - s) % Zc + Zc) % Zc];
    /* :97-98 / :103-104  pc2 = L1[0] + B(0,0) pc1 ; pc4 = L1[3] + B(3,0) pc1 */
    memcpy(pc2, L1 + 0 * Zc, (size_t)Zc);
    xor_rot(pc2, pc1, bg_entry(&g, 0, g.kb) % Zc, Zc);
    memcpy(pc4, L1 + 3 * Zc, (size_t)Zc);
    xor_rot(pc4, pc1, bg_entry(&g, 3, g.kb) % Zc, Zc);
    if (bgn == 1) { /* :99  pc3 = L1[2] + B(2,3) pc4 */
        memcpy(pc3, L1 + 2 * Zc, (size_t)Zc);
        xor_rot(pc3, pc4, bg_entry(&g, 2, g.kb + 3) % Zc, Zc);
    } else { /* :105  pc3 = L1[1] + B(1,1) pc2 */
        memcpy(pc3, L1 + 1 * Zc, (size_t)Zc);
        xor_rot(pc3, pc2, bg_entry(&g, 1, g.kb + 1) % Zc, Zc);
    }
    /* :90,:110-112  pe = C @ [ck;pc] mod 2 with C = H[4Zc:, 0:K+4Zc] */
    for (int i = 4; i < g.rows; ++i) {
        int8_t *pe = cw + (g.kb + i) * Zc;
        for (int e = g.rowptr[i]; e < g.rowptr[i + 1]; ++e)
            if (g.col[e] < ncw) xor_rot(pe, cw + g.col[e] * Zc, g.shift[e] % Zc, Zc);
    }
    /* nr_ldpc_encode.py:47-48  dn[K-2Zc : N] = wn */
    memcpy(dn + (K - 2 * Zc), cw + K, (size_t)(N + 2 * Zc - K));
    free(cw); free(L1); free(L2);
    return 0;
}

int oracle_encode_batch(int8_t *ck, int B, int bgn, int Zc, int8_t *dn, int nthreads)
{
    bg_t g;
    int iLS = oracle_find_ils(Zc);
    if (iLS > 7 || get_bg(bgn, iLS, &g)) return -1;
    const long K = (long)g.kb * Zc, N = (long)(g.cols - 2) * Zc;
    int rc = 0;
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel for schedule(dynamic, 1)
    for (int b = 0; b < B; ++b)
        if (oracle_encode(ck + b * K, bgn, Zc, dn + b * N)) rc = -1;
    return rc;
}

/* ------------------------------------------------------------------ soft decoders (two precisions) */

#define REAL double
#define SUF(x) x##_f64
#include "nrldpc_oracle_soft.inc"
#undef REAL
#undef SUF
#define REAL float
#define SUF(x) x##_f32
#include "nrldpc_oracle_soft.inc"
#undef REAL
#undef SUF

/* ------------------------------------------------------------------ bit flipping */

/*
 * ldpc_decoder_BF(LLRin, H, L): py5gphy/ldpc/ldpc_decoder_bit_flipping.py:5-73 on a CSR H.
 *   hard decision LLR>0 -> 0, LLR<0 -> 1, LLR==0 stays 0 (:41-43; the value 0.0 is kept);
 *   per iteration S = H ck mod 2 (:47), success when S == 0 (:54-56);
 *   En = (2S-1) @ H (:61); flip every bit whose En equals max(En) (:67-70); False after L (:73).
 * ck_out holds 0/1 (the reference returns them as float64 0.0/1.0).
 */
int oracle_bf_csr(const double *llr, int M, int Nv, const int *rowptr, const int *colidx, int L,
                  int8_t *ck_out, int *status, int *iters)
{
    int8_t *ck = ck_out;
    int8_t *S = (int8_t *)malloc((size_t)M);
    int *En = (int *)malloc(sizeof(int) * (size_t)Nv);
    for (int n = 0; n < Nv; ++n) ck[n] = llr[n] < 0 ? 1 : 0;
    *status = 0; *iters = L;
    for (int it = 0; it < L; ++it) {
        int any = 0;
        for (int m = 0; m < M; ++m) {
            int p = 0;
            for (int e = rowptr[m]; e < rowptr[m + 1]; ++e) p ^= ck[colidx[e]];
            S[m] = (int8_t)p; any |= p;
        }
        if (!any) { *status = 1; *iters = it; break; }
        memset(En, 0, sizeof(int) * (size_t)Nv);
        for (int m = 0; m < M; ++m)
            for (int e = rowptr[m]; e < rowptr[m + 1]; ++e) En[colidx[e]] += 2 * S[m] - 1;
        int mx = En[0];
        for (int n = 1; n < Nv; ++n) if (En[n] > mx) mx = En[n];
        for (int n = 0; n < Nv; ++n) if (En[n] == mx) ck[n] = (int8_t)(1 - ck[n]);
    }
    free(S); free(En);
    return 0;
}

/* nr_decode_ldpc(..., algo='BF'): py5gphy/ldpc/nr_ldpc_decode.py:11-49,65-67 */
int oracle_decode_bf(const double *llr, int bgn, int Zc, int L, int8_t *ck, int *status, int *iters)
{
    bg_t g;
    int iLS = oracle_find_ils(Zc);
    if (iLS > 7 || get_bg(bgn, iLS, &g)) return -1;
    int M = g.rows * Zc, Nv = g.cols * Zc, N = (g.cols - 2) * Zc;
    int *rowptr = (int *)malloc(sizeof(int) * (size_t)(M + 1));
    int *colidx = (int *)malloc(sizeof(int) * (size_t)g.nnz * Zc);
    double *full = (double *)calloc((size_t)Nv, sizeof(double));
    oracle_build_csr(bgn, Zc, rowptr, colidx);
    memcpy(full + 2 * Zc, llr, sizeof(double) * (size_t)N); /* :43 prepend 2Zc zeros */
    int rc = oracle_bf_csr(full, M, Nv, rowptr, colidx, L, ck, status, iters);
    free(rowptr); free(colidx); free(full);
    return rc;
}

/* ------------------------------------------------------------------ CRC */

/*
 * nr_crc_encode(blk, poly): py5gphy/crc/crc.py:4-41 with the polynomials of :96-106
 * (bit arrays WITHOUT the leading x^L term).  poly_id: 0='6' 1='11' 2='16' 3='24A' 4='24B' 5='24C'.
 * out holds A + L bits.  Returns L.  (mask = 0 only: the LDPC path never masks.)
 */
int oracle_crc_encode(const int8_t *blk, int A, int poly_id, int8_t *out)
{
    static const char *polys[6] = {"100001", "11000100001", "0001000000100001",
                                   "100001100100110011111011", "100000000000000001100011",
                                   "101100101011000100010111"};
    if (poly_id < 0 || poly_id > 5) return -1;
    const char *p = polys[poly_id];
    int L = (int)strlen(p);
    int8_t rem[24];
    memcpy(out, blk, (size_t)A);
    memset(out + A, 0, (size_t)L);
    /* :28-33 long division, one message bit per step */
    memcpy(rem, out, (size_t)L);
    for (int idx = 0; idx < A; ++idx) {
        int first = rem[0];
        memmove(rem, rem + 1, (size_t)(L - 1));
        rem[L - 1] = out[idx + L];
        if (first) for (int t = 0; t < L; ++t) rem[t] ^= (int8_t)(p[t] - '0');
    }
    memcpy(out + A, rem, (size_t)L);
    return L;
}

/* ------------------------------------------------------------------ rate matching / recovery, HARQ combining */

/*
 * ratematch_ldpc(dn, Ncb, E, k0, Qm): py5gphy/ldpc/nr_ldpc_ratematch.py:64-97, statement for statement.
 * dn[N] in {0,1,-1}, fe[E].
 */
int oracle_ratematch(const int8_t *dn, int N, int Ncb, int E, int k0, int Qm, int8_t *fe)
{
    if (N < Ncb || Qm <= 0 || E % Qm) return -1;
    int8_t *ek = (int8_t *)malloc((size_t)(E > 0 ? E : 1));
    int k = 0, j = 0;
    while (k < E) { /* :82-87 bit selection, fillers skipped */
        int8_t v = dn[(k0 + j) % Ncb];
        if (v != -1) ek[k++] = v;
        ++j;
    }
    /* :90-93 bit interleaving: d1 = ek.reshape(Qm, E/Qm); fe = d1.T.reshape(E) */
    int cols = E / Qm;
    for (int q = 0; q < Qm; ++q)
        for (int e = 0; e < cols; ++e) fe[e * Qm + q] = ek[q * cols + e];
    free(ek);
    return 0;
}

/*
 * raterecover_ldpc(LLr_fe, Ncb, N, k0, Qm, Zc, K_apo, K): py5gphy/ldpc/nr_ldpc_raterecover.py:6-65,
 * statement for statement (float64).  out[N].
 */
int oracle_raterecover(const double *fe, int E, int Ncb, int N, int k0, int Qm, int Zc, int K_apo, int K, double *out)
{
    if (Qm <= 0 || E % Qm || E <= 0) return -1;
    int cols = E / Qm;
    double *ek = (double *)malloc(sizeof(double) * (size_t)E);
    double mx = 0.0;
    for (int e = 0; e < cols; ++e) /* :25-28 d1 = fe.reshape(E/Qm, Qm); ek = d1.T.reshape(E) */
        for (int q = 0; q < Qm; ++q) ek[q * cols + e] = fe[e * Qm + q];
    for (int e = 0; e < E; ++e) if (fabs(fe[e]) > mx) mx = fabs(fe[e]);
    double max_llr = mx * 10; /* :30 */
    int F0 = K_apo - 2 * Zc, F1 = K - 2 * Zc; /* :34 filler positions of dn */
    int size = Ncb - (F1 - F0);               /* :38 */
    int rep_num = (E + size - 1) / size;      /* :39 */
    double *tmp = (double *)calloc((size_t)rep_num * Ncb, sizeof(double));
    double *rep = (double *)malloc(sizeof(double) * (size_t)Ncb);
    for (int p = 0; p < Ncb; ++p) rep[p] = 10000;
    int rep_idx = -1, k = 0, j = 0;
    while (k < E) { /* :47-62 */
        int pos = (k0 + j) % Ncb;
        if (pos == k0) ++rep_idx;
        if (!(pos >= F0 && pos < F1)) tmp[(size_t)rep_idx * Ncb + pos] = ek[k++];
        if (rep[pos] == 10000) rep[pos] = 1; else rep[pos] += 1;
        ++j;
    }
    for (int p = 0; p < N; ++p) out[p] = 0.0;
    for (int p = 0; p < Ncb; ++p) { /* :65 np.sum(tmp_buf, axis=0) / rep_buf: rows added in order */
        double s = tmp[p];
        for (int r = 1; r < rep_num; ++r) s += tmp[(size_t)r * Ncb + p];
        out[p] = s / rep[p];
    }
    for (int p = F0; p < F1; ++p) if (p >= 0 && p < N) out[p] = max_llr; /* :66 */
    free(ek); free(tmp); free(rep);
    return 0;
}

/* HARQ combining loop of DLSCHDecode: py5gphy/nr_pdsch/nr_dlsch_decode.py:80-87 */
void oracle_harq_combine(const double *nw, const double *cur, long n, double *out)
{
    for (long m = 0; m < n; ++m)
        out[m] = (nw[m] == 0 || cur[m] == 0) ? nw[m] + cur[m] : (nw[m] + cur[m]) / 2;
}

/* torchrun exports OMP_NUM_THREADS=1: the CPU-baseline legs of bench.py ask for every host core explicitly */
void oracle_set_num_threads(int n)
{
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

int oracle_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
