"""Shared-channel transport-block chain around the LDPC hot path, batched over the codeblocks of a
transport block: everything between the transport-block bits and the rate-matched
sequence stays on the device, one kernel launch per stage instead of one Python iteration per codeblock.

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


def _device():
    import torch
    assert torch.cuda.is_available(), "python_5gtoolbox_b200 has no CPU fallback: a CUDA device is required"
    return torch.device("cuda", torch.cuda.current_device())


def encode_ratematch(cbs, Zc, bgn, Qm, G, num_of_layers, rv, Ncb=None):
    """7.2.4-7.2.6 / 6.2.4-6.2.6 for all codeblocks at once: LDPC encode, rate match, concatenate
    (py5gphy/nr_pdsch/nr_dlsch.py:46-72, py5gphy/nr_pusch/nr_ulsch.py:37-68).  cbs int8 [C,K] (NumPy,
    fillers -1; modified in place like encode_ldpc does) -> g_seq int8 [G] (NumPy)."""
    import torch
    C, K = cbs.shape
    N = (66 if bgn == 1 else 50) * Zc
    Ncb = N if Ncb is None else Ncb
    Er_list = nr_ldpc_ratematch.get_Er_ldpc(G, C, Qm, num_of_layers)
    k0 = nr_ldpc_ratematch.get_k0(Ncb, bgn, rv, Zc)
    dev = _device()
    d_cbs = torch.from_numpy(np.ascontiguousarray(cbs, np.int8)).to(dev)
    dn = engine.encode_batch(d_cbs, bgn, Zc, fix_fillers=True)
    g = engine.ratematch_batch(dn, Ncb, Er_list, k0, Qm)
    cbs[...] = d_cbs.cpu().numpy()          # the reference's in-place filler side effect (nr_ldpc_encode.py:32-35)
    g_seq = np.zeros(G, 'i1')
    g_seq[:g.numel()] = g.cpu().numpy()
    return g_seq


def sch_decode(LLr, G, TBSize, Qm, coderateby1024, num_of_layers, rv, Ncb_of, LDPC_decoder_config, HARQ_on, current_LLr_dns):
    """De-rate-matching, HARQ combining, LDPC decoding, CB/TB CRC for all codeblocks of a transport block
    (py5gphy/nr_pdsch/nr_dlsch_decode.py:13-109, py5gphy/nr_pusch/nr_ulsch_decode.py:13-110).
    G sizes the codeblocks (get_Er_ldpc), Ncb_of(C, N) gives the circular-buffer length.  Returns (tb_ok, tbblk int8[A], new_LLr_dns float64[C,N])."""
    import torch
    LLr = np.asarray(LLr)
    A = TBSize
    B = A + (24 if A > 3824 else 16)
    bgn = select_bgn(A, coderateby1024)
    C, cbz, L, F, K, Zc = ldpc_info.get_cbs_info(B, bgn)
    K_apo = cbz + L
    N = (66 if bgn == 1 else 50) * Zc
    Ncb = Ncb_of(C, N)
    k0 = nr_ldpc_ratematch.get_k0(Ncb, bgn, rv, Zc)
    Er_list = nr_ldpc_ratematch.get_Er_ldpc(G, C, Qm, num_of_layers)
    dev = _device()
    used = int(sum(Er_list))
    x = LLr.reshape(-1)[:used]
    x = np.ascontiguousarray(x, np.float32 if x.dtype == np.float32 else np.float64)
    llr_dn = engine.raterecover_batch(torch.from_numpy(x).to(dev), Er_list, Ncb, N, k0, Qm, Zc, K_apo, K, out_f64=True)
    cur = np.asarray(current_LLr_dns)
    if HARQ_on and cur.size != 0:
        llr_dn = engine.harq_combine(llr_dn, torch.from_numpy(np.ascontiguousarray(cur, np.float64)).to(dev))
    # the [C,N] float64 soft buffer goes back to the caller (HARQ state): copy it through pinned memory while
    # the decoder runs
    new_host = torch.empty(llr_dn.shape, dtype=torch.float64, pin_memory=True)
    new_host.copy_(llr_dn, non_blocking=True)
    cfg = LDPC_decoder_config
    algo = cfg["algo"]
    if algo == 'min-sum':
        res = engine.decode_batch(llr_dn.to(torch.float32), Zc, bgn, cfg["L"], cfg["alpha"], cfg["beta"], True)
        ck = res["ck"][:, :K_apo].contiguous()
        if C > 1:
            engine.crc_check_device(ck, '24B')   # computed and ignored, like the reference (nr_dlsch_decode.py:93-98)
        # transport block = the codeblocks' payloads back to back (C * cbz == B); TB CRC on the device
        tb = ck[:, :cbz].reshape(1, C * cbz).contiguous()
        tb_err = engine.crc_check_device(tb, tb_crc_poly(A))
        tbblk = tb[0, :A].cpu().numpy()
        ok = int(tb_err.cpu()[0]) == 0   # synchronises: the pinned copy of the soft buffer is complete too
        torch.cuda.synchronize()
        return ok, tbblk, new_host.numpy()
    elif algo == 'BF':
        torch.cuda.synchronize()
        new_LLr_dns = new_host.numpy()
        ck, _, _ = engine.decode_bf_batch(new_LLr_dns, Zc, bgn, cfg["L"])
        blocks = ck[:, :cbz].astype(np.float64)
    else:
        torch.cuda.synchronize()
        new_LLr_dns = new_host.numpy()
        ck, _, _ = engine.decode_ref_batch(new_LLr_dns, Zc, bgn, cfg["L"], algo, cfg["alpha"], cfg["beta"], True, f64=True)
        blocks = ck[:, :cbz]
    torch.cuda.synchronize()
    new_LLr_dns = new_host.numpy()
    tbblkandcrc = np.zeros(B)
    tbblkandcrc[:C * cbz] = blocks.reshape(-1)
    tbblk, tbcrc_error = crc.nr_crc_decode(tbblkandcrc, tb_crc_poly(A))
    return tbcrc_error == 0, tbblk, new_LLr_dns


def lbrm_ncb(TBS_LBRM):
    """Circular buffer length with limited-buffer rate matching (I_LBRM = 1, DL-SCH) --
    py5gphy/nr_pdsch/nr_dlsch.py:60-62."""
    return lambda C, N: min(N, math.floor(TBS_LBRM / (C * 2 / 3)))
