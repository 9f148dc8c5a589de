"""Mirror of py5gphy/ldpc/ldpc_info.py: get_cbs_info, find_iLS, getH, gen_ldpc_para.

The base-graph tables come from python_5gtoolbox_b200/data/bg_tables.npz (converted from the
reference's tables by tools/gen_tables.py; same numbers as TS 38.212 Tables 5.3.2-2/-3).  These are
host-side integer helpers in the reference too ("keep as Python")."""
import math
import os

import numpy as np

_LIFT_SIZES = sorted(a << j for a, jm in ((2, 7), (3, 7), (5, 6), (7, 5), (9, 5), (11, 5), (13, 4), (15, 4))
                     for j in range(jm + 1))
_TABLES = None


def _tables():
    global _TABLES
    if _TABLES is None:
        path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "data", "bg_tables.npz")
        with np.load(path) as z:
            _TABLES = {1: z["BG1"], 2: z["BG2"]}
    return _TABLES


def base_graph(bgn, iLS):
    """int16 [46,68] / [42,52] shift table, -1 = null (what the reference loads at ldpc_info.py:110-113)."""
    return _tables()[bgn][iLS]


def get_cbs_info(B, bgn):
    """(C, cbz, L, F, K, Zc) of TS 38.212 5.2.2 -- py5gphy/ldpc/ldpc_info.py:5-78 (same asserts)."""
    Kcb = 8448 if bgn == 1 else 3840
    if B <= Kcb:
        L, C = 0, 1
    else:
        L = 24
        C = int(math.ceil(B / (Kcb - L)))
    Bd = B + C * L
    assert (B % C) == 0      # ldpc_info.py:41
    assert (Bd % C) == 0     # ldpc_info.py:46
    cbz, Kd = B // C, Bd // C
    if bgn == 1:
        Kb = 22
    else:
        Kb = 10 if B > 640 else 9 if B > 560 else 8 if B > 192 else 6
    Zc = next(z for z in _LIFT_SIZES if z * Kb >= Kd)
    K = (22 if bgn == 1 else 10) * Zc
    return C, cbz, L, K - Kd, K, Zc


def find_iLS(Zc):
    """Set index of TS 38.212 Table 5.3.2-1; 255 for an invalid Zc -- py5gphy/ldpc/ldpc_info.py:81-97."""
    for iLS, a in enumerate((2, 3, 5, 7, 9, 11, 13, 15)):
        q, j = Zc, 0
        while q > a and q % 2 == 0:
            q //= 2
            j += 1
        if q == a and Zc in _LIFT_SIZES:
            return iLS
    return 255


class TaggedH(np.ndarray):
    """Dense int8 H that remembers which (bgn, Zc) it was expanded from, so decode_ldpc can route a
    5G matrix to the quasi-cyclic kernel without re-deriving the structure from 461 MB of zeros."""
    nrldpc_tag = None

    def __array_finalize__(self, obj):
        self.nrldpc_tag = None  # views / copies / slices are no longer the full 5G matrix


def getH(Zc, bgn, iLS):
    """Dense int8 parity-check matrix 46Zc x 68Zc / 42Zc x 52Zc -- py5gphy/ldpc/ldpc_info.py:99-139.
    Block (i,j) with V = BG[i,j] > -1 has its 1 of row r at column (r + V mod Zc) mod Zc (:126-137)."""
    BG = base_graph(bgn, iLS)
    nr, nc = BG.shape
    H = np.zeros((nr * Zc, nc * Zc), "i1")
    r = np.arange(Zc)
    for i, j in zip(*np.nonzero(BG > -1)):
        H[i * Zc + r, j * Zc + (r + int(BG[i, j]) % Zc) % Zc] = 1
    H = H.view(TaggedH)
    if find_iLS(Zc) == iLS:
        H.nrldpc_tag = (bgn, Zc)
        # the tag routes decode_ldpc / ldpc_decoder_BF to the quasi-cyclic kernels, which never look at the entries:
        # the tagged matrix is read-only so that it cannot drift away from its tag (an in-place edit raises; edit a
        # .copy(), which carries no tag and is decoded entry by entry on the generic kernels)
        H.flags.writeable = False
    return H


def gen_ldpc_para(N, bgn):
    """(H, K, Zc) -- py5gphy/ldpc/ldpc_info.py:141-156"""
    if bgn == 1:
        Zc = N // 66
        K = 22 * Zc
    else:
        Zc = N // 50
        K = 10 * Zc
    iLS = find_iLS(Zc)
    assert iLS < 8
    return getH(Zc, bgn, iLS), K, Zc
