"""Mirror of py5gphy/ldpc/ldpc_decoder_bit_flipping.py: ldpc_decoder_BF on the CUDA bit-flipping kernel."""
import numpy as np

from .. import engine


def _csr_of(H):
    H = np.asarray(H)
    rows, cols = np.nonzero(H)
    rowptr = np.zeros(H.shape[0] + 1, np.int64)
    np.add.at(rowptr, rows + 1, 1)
    return np.cumsum(rowptr).astype(np.int32), cols.astype(np.int32)


def ldpc_decoder_BF(LLRin, H, L):
    """(ck, status) -- py5gphy/ldpc/ldpc_decoder_bit_flipping.py:5-73.  ck is float64 0.0/1.0 like the
    reference (it copies LLRin, :41)."""
    M, N = H.shape
    assert LLRin.size == N
    tag = getattr(H, "nrldpc_tag", None)
    llr = np.asarray(LLRin, np.float64).reshape(1, N)
    if tag is not None:
        bgn, Zc = tag
        # the kernel prepends the punctured zeros itself; LLR == 0 there decodes to 0 either way,
        # but only if the caller really passed zeros for the first 2Zc entries
        if not np.any(llr[0, :2 * Zc]):
            ck, st, _ = engine.decode_bf_batch(llr[:, 2 * Zc:], Zc, bgn, L)
            return ck[0].astype(np.float64), bool(st[0])
    rowptr, colidx = _csr_of(H)
    ck, st, _ = engine.decode_bf_csr_batch(llr, rowptr, colidx, N, L)
    return ck[0].astype(np.float64), bool(st[0])


def for_test_ldpc_encoder(K, H, snr_db):
    """(dn, LLRin) -- py5gphy/ldpc/ldpc_decoder_bit_flipping.py:75-97: systematic encoding of a generic
    H = [H1 | H2] by solving H2 w = H1 c, BPSK + AWGN LLRs.  Test-input generator on tiny matrices; the
    global NumPy RNG is drawn in the reference's order (randint, then normal)."""
    H = np.asarray(H)
    ck = np.random.randint(2, size=K)
    M, N = H.shape
    H1, H2 = H[:, 0:N - M].astype(np.float64), H[:, N - M:N].astype(np.float64)
    wn = (np.round(np.linalg.solve(H2, -H1 @ ck.T)) % 2).astype("i1")
    dn = np.concatenate((ck, wn))
    en = 1 - 2 * dn
    fn = en + np.random.normal(0, 10 ** (-snr_db / 20), dn.size)
    noise_power = 10 ** (-snr_db / 10)
    return dn, 2 * fn / noise_power
