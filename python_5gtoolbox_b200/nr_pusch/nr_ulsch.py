"""Mirror of py5gphy/nr_pusch/nr_ulsch.py (UL-SCH without UCI): CRC + segmentation, then LDPC encoding and
rate matching of all codeblocks in one batched device pass."""
import numpy as np

from .. import crc, sch
from ..ldpc import nr_ldpc_cbsegment


def ULSCH_Crc_CodeBlockSegment(trblk, TBSize, coderateby1024):
    """(cbs, Zc, bgn) -- TS 38.212 6.2.1-6.2.3, py5gphy/nr_pusch/nr_ulsch.py:13-35."""
    assert len(trblk) == TBSize
    A = TBSize
    blkandcrc = crc.nr_crc_encode(np.asarray(trblk), sch.tb_crc_poly(A))
    bgn = sch.select_bgn(A, coderateby1024)
    cbs, Zc = nr_ldpc_cbsegment.ldpc_cbsegment(blkandcrc, bgn)
    return cbs, Zc, bgn


def ULSCH_encoding_ratematch(cbs, Zc, bgn, Qm, G_ULSCH, num_of_layers, rv):
    """g_seq -- TS 38.212 6.2.4-6.2.6 with I_LBRM = 0 (Ncb = N), py5gphy/nr_pusch/nr_ulsch.py:37-68."""
    return sch.encode_ratematch(cbs, Zc, bgn, Qm, G_ULSCH, num_of_layers, rv, Ncb=None)
