"""Shared-channel transport-block chain around the LDPC hot path, batched over the codeblocks of a
transport block: everything between the transport-block bits and the rate-matched
sequence stays on the device; a whole transport block is ONE C-ABI call each way (nrldpc_sch_encode_host,
nrldpc_sch_decode_host) with no tensor library in between.

The reference-facing functions live in python_5gtoolbox_b200/nr_pdsch and nr_pusch; they keep the
reference's signatures, return types and quirks (CB-CRC result ignored, LDPC status ignored, success =
TB CRC only; py5gphy/nr_pdsch/nr_dlsch_decode.py:91-109)."""
import math

import numpy as np

from . import crc, engine
from .ldpc import ldpc_info, nr_ldpc_cbsegment, nr_ldpc_ratematch


def select_bgn(A, coderateby1024):
    """LDPC base graph selection, TS 38.212 7.2.2 / 6.2.2 -- py5gphy/nr_pdsch/nr_dlsch.py:38-43."""
    if (A <= 292) or ((A <= 3824) and (coderateby1024 <= 0.67 * 1024)) or (coderateby1024 <= 0.25 * 1024):
        return 2
    return 1


def tb_crc_poly(A):
    return '24A' if A > 3824 else '16'


def encode_ratematch(cbs, Zc, bgn, Qm, G, num_of_layers, rv, Ncb=None):
    """7.2.4-7.2.6 / 6.2.4-6.2.6 for all codeblocks at once: LDPC encode, rate match, concatenate
    (py5gphy/nr_pdsch/nr_dlsch.py:46-72, py5gphy/nr_pusch/nr_ulsch.py:37-68).  cbs int8 [C,K] (NumPy,
    fillers -1; modified in place like encode_ldpc does) -> g_seq int8 [G] (NumPy).  One C-ABI call."""
    C, K = cbs.shape
    N = (66 if bgn == 1 else 50) * Zc
    Ncb = N if Ncb is None else Ncb
    Er_list = nr_ldpc_ratematch.get_Er_ldpc(G, C, Qm, num_of_layers)
    k0 = nr_ldpc_ratematch.get_k0(Ncb, bgn, rv, Zc)
    if cbs.dtype == np.int8 and cbs.flags.c_contiguous:
        g = engine.encode_ratematch_host(cbs, bgn, Zc, Ncb, k0, Qm, Er_list)
    else:
        tmp = np.ascontiguousarray(cbs, np.int8)
        g = engine.encode_ratematch_host(tmp, bgn, Zc, Ncb, k0, Qm, Er_list)
        cbs[...] = tmp          # the reference's in-place filler side effect (nr_ldpc_encode.py:32-35)
    if g.size == G:
        return g
    g_seq = np.zeros(G, 'i1')
    g_seq[:g.size] = g
    return g_seq


def dlsch_encode(trblk, TBSize, Qm, coderateby1024, num_of_layers, rv, Ncb_of, G):
    """DLSCHEncode in one C-ABI call: TB CRC, segmentation + CB CRC, LDPC encoding, rate matching and concatenation
    stay on the device (py5gphy/nr_pdsch/nr_dlsch.py:12-74)."""
    assert len(trblk) == TBSize
    A = TBSize
    B = A + (24 if A > 3824 else 16)
    bgn = select_bgn(A, coderateby1024)
    C, cbz, L, F, K, Zc = ldpc_info.get_cbs_info(B, bgn)
    N = (66 if bgn == 1 else 50) * Zc
    Ncb = Ncb_of(C, N)
    Er_list = nr_ldpc_ratematch.get_Er_ldpc(G, C, Qm, num_of_layers)
    k0 = nr_ldpc_ratematch.get_k0(Ncb, bgn, rv, Zc)
    g = engine.sch_encode_host(trblk, C, bgn, Zc, Ncb, k0, Qm, Er_list)
    if g.size == G:
        return g
    g_seq = np.zeros(G, 'i1')
    g_seq[:g.size] = g
    return g_seq


def sch_decode(LLr, G, TBSize, Qm, coderateby1024, num_of_layers, rv, Ncb_of, LDPC_decoder_config, HARQ_on, current_LLr_dns,
               soft_buffer=True):
    """De-rate-matching, HARQ combining, LDPC decoding, CB/TB CRC for all codeblocks of a transport block
    (py5gphy/nr_pdsch/nr_dlsch_decode.py:13-109, py5gphy/nr_pusch/nr_ulsch_decode.py:13-110).
    G sizes the codeblocks (get_Er_ldpc), Ncb_of(C, N) gives the circular-buffer length.  Returns (tb_ok, tbblk int8[A],
    new_LLr_dns float64[C,N]).  algo 'min-sum': ONE C-ABI call (nrldpc_sch_decode_host: the decoder recovers and
    combines its own LLRs and writes new_LLr_dns straight into pinned host memory)."""
    A = TBSize
    B = A + (24 if A > 3824 else 16)
    bgn = select_bgn(A, coderateby1024)
    C, cbz, L, F, K, Zc = ldpc_info.get_cbs_info(B, bgn)
    K_apo = cbz + L
    N = (66 if bgn == 1 else 50) * Zc
    Ncb = Ncb_of(C, N)
    k0 = nr_ldpc_ratematch.get_k0(Ncb, bgn, rv, Zc)
    Er_list = nr_ldpc_ratematch.get_Er_ldpc(G, C, Qm, num_of_layers)
    used = int(sum(Er_list))
    x = np.asarray(LLr).reshape(-1)[:used]
    cur = np.asarray(current_LLr_dns)
    cur = cur if (HARQ_on and cur.size != 0) else None
    cfg = LDPC_decoder_config
    algo = cfg["algo"]
    if algo == 'min-sum':
        r = engine.sch_decode_host(x, Er_list, bgn, Zc, Ncb, k0, Qm, K_apo, A, cfg["L"], cfg["alpha"], cfg["beta"], cur=cur,
                                   want_soft=soft_buffer)
        return r["tb_err"] == 0, r["tbblk"], (r["soft"] if soft_buffer else np.array([]))
    new_LLr_dns = engine.sch_recover_host(x, Er_list, bgn, Zc, Ncb, k0, Qm, K_apo, cur=cur)
    if algo == 'BF':
        ck, _, _ = engine.decode_bf_batch(new_LLr_dns, Zc, bgn, cfg["L"])
        blocks = ck[:, :cbz].astype(np.float64)
    else:
        ck, _, _ = engine.decode_bp_batch(new_LLr_dns, Zc, bgn, cfg["L"])
        blocks = ck[:, :cbz]
    tbblkandcrc = np.zeros(B)
    tbblkandcrc[:C * cbz] = blocks.reshape(-1)
    tbblk, tbcrc_error = crc.nr_crc_decode(tbblkandcrc, tb_crc_poly(A))
    return tbcrc_error == 0, tbblk, new_LLr_dns


def lbrm_ncb(TBS_LBRM):
    """Circular buffer length with limited-buffer rate matching (I_LBRM = 1, DL-SCH) --
    py5gphy/nr_pdsch/nr_dlsch.py:60-62."""
    return lambda C, N: min(N, math.floor(TBS_LBRM / (C * 2 / 3)))
