"""Mirror of py5gphy/nr_pdsch/nr_dlsch.py: DLSCHEncode with all codeblocks of the transport block in one
batched device pass (CRC -> segmentation -> LDPC encode -> rate match -> concatenation)."""
import numpy as np

from .. import crc, sch
from ..ldpc import nr_ldpc_cbsegment


def DLSCHEncode(trblk, TBSize, Qm, coderateby1024, num_of_layers, rv, TBS_LBRM, G):
    """g_seq = DLSCHEncode(trblk, TBSize, Qm, coderateby1024, num_of_layers, rv, TBS_LBRM, G) -- TS 38.212
    7.2, py5gphy/nr_pdsch/nr_dlsch.py:12-74.  Returns the int8 sequence after code block concatenation."""
    return sch.dlsch_encode(np.asarray(trblk), TBSize, Qm, coderateby1024, num_of_layers, rv, sch.lbrm_ncb(TBS_LBRM), G)  # 7.2.1-7.2.6
