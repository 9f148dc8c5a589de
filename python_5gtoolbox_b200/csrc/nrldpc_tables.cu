// nrldpc_tables.cu -- host-side construction of the quasi-cyclic tables from TS 38.212 Tables 5.3.2-2/-3.
// Mirrors the semantics of py5gphy/ldpc/ldpc_info.py (find_iLS :81-97, getH :99-139).
#include <algorithm>
#include <cstring>
#include <vector>

#include "nrldpc_common.cuh"
#include "../../include/nrldpc_bg_tables.inc"

namespace nrldpc {

int find_ils(int Zc)
{
    // TS 38.212 Table 5.3.2-1: Z = a * 2^j, a in {2,3,5,7,9,11,13,15}
    static const int a[8] = {2, 3, 5, 7, 9, 11, 13, 15};
    static const int jmax[8] = {7, 7, 6, 5, 5, 5, 4, 4};
    if (Zc <= 0) return 255;
    for (int s = 0; s < 8; ++s)
        for (int j = 0; j <= jmax[s]; ++j)
            if ((a[s] << j) == Zc) return s;
    return 255;
}

struct BgView {
    int rows, cols, nnz, kb;
    const unsigned short *rowptr;
    const unsigned char *col;
    const unsigned short *shift;
};

static bool bg_view(int bgn, int iLS, BgView *g)
{
    if (iLS < 0 || iLS > 7) return false;
    if (bgn == 1) {
        *g = {NRLDPC_BG1_ROWS, NRLDPC_BG1_COLS, NRLDPC_BG1_NNZ, 22, nrldpc_bg1_rowptr, nrldpc_bg1_col, nrldpc_bg1_shift[iLS]};
    } else if (bgn == 2) {
        *g = {NRLDPC_BG2_ROWS, NRLDPC_BG2_COLS, NRLDPC_BG2_NNZ, 10, nrldpc_bg2_rowptr, nrldpc_bg2_col, nrldpc_bg2_shift[iLS]};
    } else {
        return false;
    }
    return true;
}

int build_csr(int bgn, int Zc, int32_t *rowptr, int32_t *colidx)
{
    BgView g;
    if (!bg_view(bgn, find_ils(Zc), &g)) return NRLDPC_EINVAL;
    int e = 0;
    for (int i = 0; i < g.rows; ++i)
        for (int r = 0; r < Zc; ++r) {
            rowptr[i * Zc + r] = e;
            for (int k = g.rowptr[i]; k < g.rowptr[i + 1]; ++k)
                colidx[e++] = g.col[k] * Zc + (r + g.shift[k] % Zc) % Zc;
        }
    rowptr[g.rows * Zc] = e;
    return e;
}

int build_qc_cfg(int bgn, int Zc, QcCfg *c)
{
    BgView g;
    int iLS = find_ils(Zc);
    if (!bg_view(bgn, iLS, &g)) return NRLDPC_EINVAL;
    std::memset(c, 0, sizeof(*c));
    c->bgn = bgn; c->Zc = Zc; c->iLS = iLS;
    c->nrows = g.rows; c->ncols = g.cols; c->kb = g.kb; c->ncore = g.kb + 4;
    c->K = g.kb * Zc; c->N = (g.cols - 2) * Zc; c->Nfull = g.cols * Zc; c->M = g.rows * Zc;
    c->tiles = (Zc + 31) / 32;
    int lanes = 32, lg = 5;
    if (Zc < 32) { lanes = 1; lg = 0; while (lanes < Zc) { lanes <<= 1; ++lg; } }
    c->lanes = lanes; c->lanes_log2 = lg; c->per = 32 / lanes;

    for (int i = 0; i <= g.rows; ++i) c->rowptr[i] = g.rowptr[i];
    for (int e = 0; e < g.nnz; ++e) c->edge[e] = (uint32_t)g.col[e] | ((uint32_t)(g.shift[e] % Zc) << 8);
    // structural facts the kernels rely on
    for (int i = 0; i < g.rows; ++i) {
        int last = g.rowptr[i + 1] - 1;
        for (int e = g.rowptr[i]; e <= last; ++e) {
            bool ext = g.col[e] >= c->ncore;
            if (ext != (i >= 4 && e == last)) return NRLDPC_EINVAL;
            if (ext && (g.col[e] != g.kb + i || g.shift[e] % Zc != 0)) return NRLDPC_EINVAL;
        }
    }
    // CSC over the core columns, ascending row-block
    int n = 0;
    for (int j = 0; j < c->ncore; ++j) {
        c->colptr[j] = (uint16_t)n;
        for (int i = 0; i < g.rows; ++i)
            for (int e = g.rowptr[i]; e < g.rowptr[i + 1]; ++e)
                if (g.col[e] == j) {
                    int k = e - g.rowptr[i], deg = g.rowptr[i + 1] - g.rowptr[i];
                    int back = (Zc - g.shift[e] % Zc) % Zc;
                    c->centry[n++] = (uint32_t)i | ((uint32_t)k << 6) | ((uint32_t)(deg - 1 - k) << 11) | ((uint32_t)back << 16);
                }
    }
    c->colptr[c->ncore] = (uint16_t)n;

    std::vector<int> ord(g.rows);
    for (int i = 0; i < g.rows; ++i) ord[i] = i;
    std::stable_sort(ord.begin(), ord.end(), [&](int a, int b) {
        return (g.rowptr[a + 1] - g.rowptr[a]) > (g.rowptr[b + 1] - g.rowptr[b]);
    });
    for (int i = 0; i < g.rows; ++i) c->cn_order[i] = (uint8_t)ord[i];
    std::vector<int> cord(c->ncore);
    for (int j = 0; j < c->ncore; ++j) cord[j] = j;
    std::stable_sort(cord.begin(), cord.end(), [&](int a, int b) {
        return (c->colptr[a + 1] - c->colptr[a]) > (c->colptr[b + 1] - c->colptr[b]);
    });
    for (int j = 0; j < c->ncore; ++j) c->vn_order[j] = (uint8_t)cord[j];

    // sign-word layout: 32-bit words first (keeps them 4-byte aligned), then the 16-bit ones
    uint32_t off = 0;
    for (int pass = 0; pass < 2; ++pass)
        for (int i = 0; i < g.rows; ++i) {
            int deg = g.rowptr[i + 1] - g.rowptr[i];
            bool wide = deg > 12;
            if (deg > 27) return NRLDPC_EINVAL;
            if (wide == (pass == 0)) {
                c->wide[i] = wide;
                c->bits_off[i] = off;
                off += wide ? 4 : 2;
            }
        }
    c->bits_bytes_per_zc = (int)off;
    return NRLDPC_OK;
}

}  // namespace nrldpc
