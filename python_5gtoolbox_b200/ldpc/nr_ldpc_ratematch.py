"""Mirror of py5gphy/ldpc/nr_ldpc_ratematch.py: get_Er_ldpc, get_k0 (host integer helpers, as in the
reference) and ratematch_ldpc on the CUDA rate-matching kernel (csrc/nrldpc_ratematch.cu)."""
import math

import numpy as np

from .. import engine


def get_Er_ldpc(G, C, Qm, NL):
    """Rate-matching output length of each of the C codeblocks, TS 38.212 5.4.2.1 --
    py5gphy/ldpc/nr_ldpc_ratematch.py:5-28 (no CBGTI): the first C - (G/(NL Qm) mod C) codeblocks get the
    floor share, the others the ceil share."""
    unit = NL * Qm
    per_cb = G / (unit * C)
    n_floor = C - ((G / unit) % C)
    return [unit * (math.floor(per_cb) if j <= n_floor - 1 else math.ceil(per_cb)) for j in range(C)]


_K0_NUM = {1: (0, 17, 33, 56), 2: (0, 13, 25, 43)}  # TS 38.212 Table 5.4.2.1-2


def get_k0(Ncb, bgn, rv, Zc):
    """Starting position of redundancy version rv in the circular buffer --
    py5gphy/ldpc/nr_ldpc_ratematch.py:30-61."""
    assert rv in [0, 1, 2, 3]
    assert bgn in [1, 2]
    if rv == 0:
        return 0
    den = (66 if bgn == 1 else 50) * Zc
    return math.floor(_K0_NUM[bgn][rv] * Ncb / den) * Zc


def ratematch_ldpc(dn, Ncb, E, k0, Qm):
    """fe = ratematch_ldpc(dn, Ncb, E, k0, Qm): bit selection + bit interleaving of one codeblock --
    py5gphy/ldpc/nr_ldpc_ratematch.py:64-97.  Returns int8 [E]."""
    dn = np.asarray(dn)
    N = dn.size
    assert N >= Ncb
    assert E % Qm == 0   # the reference's reshape(Qm, E // Qm) raises otherwise
    return engine.ratematch_batch(np.ascontiguousarray(dn, np.int8).reshape(1, N), Ncb, [E], k0, Qm)
