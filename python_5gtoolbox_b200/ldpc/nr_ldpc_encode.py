"""Mirror of py5gphy/ldpc/nr_ldpc_encode.py: encode_ldpc(ck, bgn) on the CUDA encoder."""
import numpy as np

from .. import engine
from . import ldpc_info


def encode_ldpc(ck, bgn):
    """dn = encode_ldpc(ck, bgn) -- TS 38.212 5.3.2, py5gphy/ldpc/nr_ldpc_encode.py:8-50.

    ck: K-long code block, -1 = filler; like the reference (:32-35) the caller's array is modified in
    place (fillers at k >= 2Zc become 0).  Returns the N-long int8 sequence with -1 at filler positions."""
    assert bgn in [1, 2]
    K = ck.size
    Zc = K // 22 if bgn == 1 else K // 10
    assert ldpc_info.find_iLS(Zc) < 8
    assert K == (22 if bgn == 1 else 10) * Zc
    buf = np.ascontiguousarray(ck, np.int8).reshape(1, K)
    owns = buf.base is not ck and buf is not ck and not np.shares_memory(buf, ck)
    dn = engine.encode_batch(buf, bgn, Zc, fix_fillers=True)[0]
    if owns:  # caller passed another dtype / a strided view: write the side effect back
        fill = (np.arange(K) >= 2 * Zc) & (np.asarray(ck) == -1)
        ck[fill] = 0
    return dn
