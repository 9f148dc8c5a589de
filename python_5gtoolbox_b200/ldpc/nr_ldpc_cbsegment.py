"""Mirror of py5gphy/ldpc/nr_ldpc_cbsegment.py: code block segmentation + CB-CRC attachment, TS 38.212 5.2.2."""
import numpy as np

from .. import crc
from . import ldpc_info


def ldpc_cbsegment(inbits, bgn):
    """cbs, Zc = ldpc_cbsegment(inbits, bgn) -- py5gphy/ldpc/nr_ldpc_cbsegment.py:7-33: C x K int8 code
    blocks, each the cbz payload bits (+ CRC24B when C > 1) followed by -1 fillers.  The C CRCs are one
    batched call of the CUDA CRC kernel."""
    inbits = np.asarray(inbits)
    B = inbits.size
    assert bgn in [1, 2]
    C, cbz, L, F, K, Zc = ldpc_info.get_cbs_info(B, bgn)
    cbs = np.full((C, K), -1, 'i1')
    if C == 1:
        cbs[0, 0:cbz] = inbits
    else:
        cbs[:, 0:cbz + L] = crc.nr_crc_encode_batch(inbits.reshape(C, cbz), '24B')
    return cbs, Zc
