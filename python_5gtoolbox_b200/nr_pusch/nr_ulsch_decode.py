"""Mirror of py5gphy/nr_pusch/nr_ulsch_decode.py: ULSCH_decoding, batched over the codeblocks."""
import numpy as np

from .. import sch


def ULSCH_decoding(g_ulsch_LLr, TBSize, coderateby1024, Qm, G_ULSCH, num_of_layers, rv, LDPC_decoder_config, HARQ_on=False,
                   current_LLr_dns=np.array([]), *, soft_buffer=True):
    """(status, tbblk, new_LLr_dns) -- py5gphy/nr_pusch/nr_ulsch_decode.py:13-110.  As in the reference the
    codeblock lengths come from G_ULSCH (:45) and the circular buffer is unlimited (Ncb = N, :37-42)."""
    return sch.sch_decode(g_ulsch_LLr, G_ULSCH, TBSize, Qm, coderateby1024, num_of_layers, rv, lambda C, N: N,
                          LDPC_decoder_config, HARQ_on, current_LLr_dns, soft_buffer)
