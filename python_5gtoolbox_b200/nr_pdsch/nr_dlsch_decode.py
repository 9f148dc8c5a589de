"""Mirror of py5gphy/nr_pdsch/nr_dlsch_decode.py: DLSCHDecode, batched over the codeblocks."""
import numpy as np

from .. import sch


def DLSCHDecode(LLr, TBSize, Qm, coderateby1024, num_of_layers, rv, TBS_LBRM, LDPC_decoder_config, HARQ_on=False,
                current_LLr_dns=np.array([]), *, soft_buffer=True):
    """(status, tbblk, new_LLr_dns) = DLSCHDecode(...) -- py5gphy/nr_pdsch/nr_dlsch_decode.py:13-109:
    de-rate matching, optional HARQ combining with current_LLr_dns, LDPC decoding with
    LDPC_decoder_config = {"L", "algo", "alpha", "beta"}, CB CRC (ignored) and TB CRC (= status).
    Keyword-only extra (reference-preserving default): soft_buffer=False skips the float64 [C, N] soft buffer (returns an
    empty array in its place) for callers that do not keep HARQ state -- it is 2/3 of the call's time on PCIe."""
    return sch.sch_decode(LLr, np.asarray(LLr).size, TBSize, Qm, coderateby1024, num_of_layers, rv, sch.lbrm_ncb(TBS_LBRM),
                          LDPC_decoder_config, HARQ_on, current_LLr_dns, soft_buffer)
