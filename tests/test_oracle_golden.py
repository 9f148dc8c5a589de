"""CPU: pin the oracle (oracle/nrldpc_oracle.c) against outputs of the unmodified reference:
tests/golden/*.npz (tools/gen_golden.py) and the SURVEY Appendix C known-answer vectors."""
import numpy as np
import pytest

from tests.conftest import hex_to_bits
from tests.golden import kat_appendix_c as KAT


def test_encode_golden_all_lifting_sizes(oracle, enc_golden):
    assert len(enc_golden) == 2 * 51 * 2
    seen = set()
    for g in enc_golden:
        ck = g["ck"].copy()
        dn = oracle.encode_ldpc(ck, g["bgn"])
        assert np.array_equal(dn, g["dn"]), (g["bgn"], g["Zc"])
        assert np.array_equal(ck, g["ck_after"])  # the in-place filler side effect
        seen.add((g["bgn"], g["Zc"]))
    assert len(seen) == 102


def test_decode_golden(oracle, dec_golden):
    n = {"min-sum": 0, "BP": 0, "BF": 0}
    for g in dec_golden:
        llr = g["llr"].astype(np.float64)
        blk, ck, st, it = oracle.nr_decode_ldpc(llr, g["Zc"], g["bgn"], g["L"], g["algo"], g["alpha"], g["beta"])
        tag = (g["bgn"], g["Zc"], g["algo"], g["seed"])
        if g["algo"] == "BP":  # libm tanh/atanh vs NumPy: decisions must still agree on these vectors
            assert np.array_equal(ck, g["ck"]) and st == g["status"], tag
        else:
            assert np.array_equal(ck, g["ck"]), tag
            assert st == g["status"], tag
            if g["algo"] == "min-sum":
                assert it == g["iters"], tag
        n[g["algo"]] += 1
    assert n["min-sum"] >= 100 and n["BP"] >= 10 and n["BF"] >= 10


def test_decode_golden_has_both_outcomes(dec_golden):
    ms = [g for g in dec_golden if g["algo"] == "min-sum"]
    assert any(g["status"] for g in ms) and any(not g["status"] for g in ms)
    assert any(g["Zc"] == 384 and g["bgn"] == 1 for g in ms)


def test_crc_golden(oracle):
    import os
    from tests.conftest import GOLDEN
    with np.load(os.path.join(GOLDEN, "crc_golden.npz")) as z:
        keys = [k for k in z.files if k.startswith("in_")]
        assert len(keys) == 24
        for k in keys:
            _, poly, n = k.split("_")
            assert np.array_equal(oracle.nr_crc_encode(z[k], poly), z[f"out_{poly}_{n}"]), k


@pytest.mark.parametrize("case", KAT.ENC, ids=lambda c: f"bg{c[0]}z{c[1]}")
def test_kat_encode(oracle, case):
    bgn, Zc, K, F, hin, N, hdn, fidx = case
    ck = hex_to_bits(hin, K).copy()
    if F:
        ck[K - F:] = -1
    dn = oracle.encode_ldpc(ck, bgn)
    want = hex_to_bits(hdn, N).copy()
    assert sorted(np.nonzero(dn == -1)[0].tolist()) == fidx
    dn0 = dn.copy()
    dn0[dn0 == -1] = 0
    assert np.array_equal(dn0, want)


@pytest.mark.parametrize("case", KAT.DEC, ids=lambda c: f"{c[0]}-bg{c[1]}z{c[2]}L{c[3]}")
def test_kat_decode(oracle, case):
    algo, bgn, Zc, L, alpha, beta, hin, flips, status, iters, hck = case
    K, N, Nf, M = oracle.dims(bgn, Zc)
    dn = oracle.encode_ldpc(hex_to_bits(hin, K).copy(), bgn)
    llr = 4.0 * (1 - 2 * dn.astype(np.float64))
    llr[flips] *= -0.5
    for dtype in (np.float64, np.float32):  # dyadic values: fp32 must match bit for bit too
        blk, ck, st, it = oracle.nr_decode_ldpc(llr, Zc, bgn, L, algo, alpha or 1, beta or 0, dtype=dtype)
        assert np.array_equal(ck, hex_to_bits(hck, Nf)) and st == status
        if iters is not None:
            assert it == iters
