"""Mirror of py5gphy/nr_pdsch/nr_dlsch.py: DLSCHEncode with all codeblocks of the transport block in one
batched device pass (CRC -> segmentation -> LDPC encode -> rate match -> concatenation)."""
import numpy as np

from .. import crc, sch
from ..ldpc import nr_ldpc_cbsegment


def DLSCHEncode(trblk, TBSize, Qm, coderateby1024, num_of_layers, rv, TBS_LBRM, G):
    """g_seq = DLSCHEncode(trblk, TBSize, Qm, coderateby1024, num_of_layers, rv, TBS_LBRM, G) -- TS 38.212
    7.2, py5gphy/nr_pdsch/nr_dlsch.py:12-74.  Returns the int8 sequence after code block concatenation."""
    assert len(trblk) == TBSize
    A = TBSize
    blkandcrc = crc.nr_crc_encode(np.asarray(trblk), sch.tb_crc_poly(A))          # 7.2.1
    bgn = sch.select_bgn(A, coderateby1024)                                      # 7.2.2
    cbs, Zc = nr_ldpc_cbsegment.ldpc_cbsegment(blkandcrc, bgn)                   # 7.2.3
    C = cbs.shape[0]
    N = (66 if bgn == 1 else 50) * Zc
    return sch.encode_ratematch(cbs, Zc, bgn, Qm, G, num_of_layers, rv, Ncb=sch.lbrm_ncb(TBS_LBRM)(C, N))  # 7.2.4-7.2.6
