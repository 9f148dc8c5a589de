"""Mirror of py5gphy/ldpc/nr_ldpc_decode.py on the CUDA decoders.

nr_decode_ldpc / decode_ldpc keep the reference's positional signature; extras are keyword-only with
reference-preserving defaults:
  precision="fp32"  the quasi-cyclic hot kernel (fp32 arithmetic in the reference's operation order);
            "fp64"  the generic kernel in float64: bit-identical to the reference's arithmetic, slow.
"""
import numpy as np

from .. import crc, engine
from . import ldpc_decoder_bit_flipping, ldpc_info, nr_ldpc_encode


def nr_decode_ldpc(LLRin, Zc, bgn, L, algo='min-sum', alpha=1, beta=0, *, precision="fp32"):
    """(blkandcrc, ck, status) -- py5gphy/ldpc/nr_ldpc_decode.py:11-49."""
    assert bgn in [1, 2]
    assert algo in ['BF', 'BP', 'min-sum']
    N, K = (Zc * 66, Zc * 22) if bgn == 1 else (Zc * 50, Zc * 10)
    assert N == LLRin.size
    assert ldpc_info.find_iLS(Zc) < 8
    llr = np.asarray(LLRin).reshape(1, N)
    if algo == 'BF':
        ck, st, _ = engine.decode_bf_batch(llr, Zc, bgn, L)
        ck = ck[0].astype(np.float64)  # the reference's BF returns float64 0.0/1.0
    elif algo == 'BP':
        ck, st, _ = engine.decode_bp_batch(llr.astype(np.float64), Zc, bgn, L)   # quasi-cyclic sum-product kernel, float64
        ck = ck[0]
    elif precision == "fp64":
        ck, st, _ = engine.decode_ref_batch(llr, Zc, bgn, L, algo, alpha, beta, True, f64=True)
        ck = ck[0]
    else:
        res = engine.decode_batch(llr.astype(np.float32), Zc, bgn, L, alpha, beta, True)
        ck, st = res["ck"][0], res["status"]
    return ck[0:K], ck, bool(st[0])


def decode_ldpc(LLRin, H, L, algo='min-sum', alpha=1, beta=0, *, precision=None):
    """(ck, status) -- py5gphy/ldpc/nr_ldpc_decode.py:51-143 for any dense 0/1 matrix H.

    A matrix produced by this package's ldpc_info.getH carries a (bgn, Zc) tag and goes to the
    quasi-cyclic kernel (fp32) when its first 2Zc LLRs are the punctured zeros; anything else runs on
    the generic CSR kernel in float64 (the reference's arithmetic)."""
    if algo == "BF":
        return ldpc_decoder_bit_flipping.ldpc_decoder_BF(LLRin, H, L)
    M, N = H.shape
    assert LLRin.size == N
    llr = np.asarray(LLRin).reshape(1, N)
    tag = getattr(H, "nrldpc_tag", None)
    if tag is not None and algo == 'min-sum' and precision in (None, "fp32") and not np.any(llr[0, :2 * tag[1]]):
        bgn, Zc = tag
        res = engine.decode_batch(llr[:, 2 * Zc:].astype(np.float32), Zc, bgn, L, alpha, beta, True)
        return res["ck"][0], bool(res["status"][0])
    rowptr, colidx = ldpc_decoder_bit_flipping._csr_of(H)
    ck, st, _ = engine.decode_csr_batch(llr, rowptr, colidx, N, L, algo, alpha, beta, True, f64=precision != "fp32")
    return ck[0], bool(st[0])


def for_test_5g_ldpc_encoder(Zc, bgn, snr_db, crcpoly='24A'):
    """(blkandcrc, dn, LLRin) -- py5gphy/ldpc/nr_ldpc_decode.py:229-260: random bits -> CRC -> LDPC
    encode -> BPSK + AWGN -> LLR.  The two draws from NumPy's GLOBAL legacy RNG (randint :247, normal
    :253) are kept on the host in the reference's order so that np.random.seed(s) reproduces the
    reference's inputs; CRC and encoding run on the GPU."""
    assert bgn in [1, 2]
    assert crcpoly in ['24A', '24B', '16']
    K, N = (Zc * 22, Zc * 66) if bgn == 1 else (Zc * 10, Zc * 50)
    crc_len = 24 if crcpoly in ['24A', '24B'] else 16
    inbits = np.random.randint(2, size=K - crc_len)
    blkandcrc = crc.nr_crc_encode(inbits, crcpoly)
    dn = nr_ldpc_encode.encode_ldpc(blkandcrc, bgn)
    en = 1 - 2 * dn
    fn = en + np.random.normal(0, 10 ** (-snr_db / 20), dn.size)
    noise_power = 10 ** (-snr_db / 10)
    LLRin = 2 * fn / noise_power
    return blkandcrc, dn, LLRin
