"""Rows either side of the LDPC path: rate matching / recovery, segmentation, HARQ combining
and the whole DL-SCH / UL-SCH transport-block chain, against golden vectors produced by the unmodified
reference (tools/gen_golden_sch.py) and against the C oracle.

Bars: rate matching, segmentation and the encoder chain are bit-exact; rate recovery and HARQ combining
are float64-exact (same operation order as the reference); the decode chain returns the reference's
status and, when the TB CRC passes, its transport block."""
import os

import numpy as np
import pytest

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def rm_golden():
    with np.load(os.path.join(GOLD, "ratematch_golden.npz")) as z:
        return {k: z[k] for k in z.files}


@pytest.fixture(scope="module")
def sch_golden():
    with np.load(os.path.join(GOLD, "sch_golden.npz")) as z:
        d = {k: z[k] for k in z.files}
    cases = {}
    for k, v in d.items():
        name, key = k.split("__")
        cases.setdefault(name, {})[key] = v
    return cases


def _rm_cases(g):
    n = 0
    while f"meta_{n}" in g:
        bgn, Zc, F, Ncb, rv, k0, Qm, E = (int(x) for x in g[f"meta_{n}"])
        yield n, bgn, Zc, F, Ncb, rv, k0, Qm, E
        n += 1


# ------------------------------------------------------------------ CPU: oracle and host helpers vs the reference

def test_oracle_ratematch_raterecover_golden(oracle, rm_golden):
    cnt = 0
    for n, bgn, Zc, F, Ncb, rv, k0, Qm, E in _rm_cases(rm_golden):
        K, N = ((22, 66) if bgn == 1 else (10, 50))
        K, N = K * Zc, N * Zc
        assert np.array_equal(oracle.ratematch_ldpc(rm_golden[f"dn_{n}"], Ncb, E, k0, Qm), rm_golden[f"fe_{n}"]), n
        rec = oracle.raterecover_ldpc(rm_golden[f"llr_{n}"].astype(np.float64), Ncb, N, k0, Qm, Zc, K - F, K)
        assert np.array_equal(rec, rm_golden[f"rec_{n}"]), n
        cnt += 1
    assert cnt == 48


def test_ratematch_parameter_helpers(rm_golden, sch_golden):
    from python_5gtoolbox_b200.ldpc import nr_ldpc_ratematch as RM
    from python_5gtoolbox_b200 import sch
    for n, bgn, Zc, F, Ncb, rv, k0, Qm, E in _rm_cases(rm_golden):
        assert RM.get_k0(Ncb, bgn, rv, Zc) == k0
    # TS 38.212 5.4.2.1: the shares differ by at most NL*Qm and add up to G
    for G, C, Qm, NL in [(13500, 2, 6, 1), (1000 * 24, 7, 8, 3), (9600, 5, 4, 2), (600, 1, 2, 1), (32 * 32699, 115, 8, 4)]:
        Er = RM.get_Er_ldpc(G, C, Qm, NL)
        assert len(Er) == C and all(e % (Qm * NL) == 0 for e in Er) and max(Er) - min(Er) in (0, Qm * NL)
        assert sum(Er) == G and sorted(Er) == Er
    with pytest.raises(AssertionError):
        RM.get_k0(100, 1, 4, 2)
    assert sch.select_bgn(292, 900) == 2 and sch.select_bgn(3824, 686) == 2 and sch.select_bgn(3825, 686) == 1
    assert sch.select_bgn(9000, 256) == 2 and sch.select_bgn(9000, 257) == 1
    assert sch.tb_crc_poly(3824) == '16' and sch.tb_crc_poly(3825) == '24A'


def test_oracle_harq_combine(oracle):
    a = np.array([0.0, 1.5, -2.0, 0.0, 3.0, -0.0])
    c = np.array([2.0, 0.0, 4.0, 0.0, -3.0, 1.0])
    assert np.array_equal(oracle.harq_combine(a, c), np.array([2.0, 1.5, 1.0, 0.0, 0.0, 1.0]))


# ------------------------------------------------------------------ GPU: kernels and the batched chain

@pytest.fixture(scope="module")
def eng():
    from python_5gtoolbox_b200 import engine, _lib
    assert _lib.lib().nrldpc_device_count() > 0, "no CUDA device (there is no CPU fallback)"
    return engine


@pytest.mark.gpu
def test_ratematch_raterecover_golden(eng, rm_golden):
    from python_5gtoolbox_b200.ldpc import nr_ldpc_ratematch as RM, nr_ldpc_raterecover as RR
    for n, bgn, Zc, F, Ncb, rv, k0, Qm, E in _rm_cases(rm_golden):
        K, N = ((22, 66) if bgn == 1 else (10, 50))
        K, N = K * Zc, N * Zc
        fe = RM.ratematch_ldpc(rm_golden[f"dn_{n}"], Ncb, E, k0, Qm)
        assert fe.dtype == np.int8 and np.array_equal(fe, rm_golden[f"fe_{n}"]), n
        for llr in (rm_golden[f"llr_{n}"], rm_golden[f"llr_{n}"].astype(np.float64)):   # float32 and float64 inputs
            rec = RR.raterecover_ldpc(llr, Ncb, N, k0, Qm, Zc, K - F, K)
            assert rec.dtype == np.float64 and np.array_equal(rec, rm_golden[f"rec_{n}"]), n


@pytest.mark.gpu
def test_ratematch_batched_device_vs_oracle(eng, oracle):
    """A transport block's worth of codeblocks in one launch, on device tensors: mixed E (floor / ceil
    shares), fillers, LBRM, repetition; arbitrary (non-contiguous) filler patterns in the selection."""
    import torch
    rng = np.random.default_rng(3)
    for bgn, Zc, C, Qm, NL, Ncb_frac, rv, scale in [(1, 384, 12, 8, 4, 1.0, 0, 0.36), (1, 208, 5, 6, 1, 0.8, 2, 0.5),
                                                      (2, 64, 7, 2, 2, 1.0, 3, 2.3), (2, 10, 3, 1, 1, 0.9, 1, 1.0)]:
        K, N = ((22, 66) if bgn == 1 else (10, 50))
        K, N = K * Zc, N * Zc
        F = int(rng.integers(1, Zc))
        Ncb = N if Ncb_frac == 1.0 else int(N * Ncb_frac)
        from python_5gtoolbox_b200.ldpc import nr_ldpc_ratematch as RM
        k0 = RM.get_k0(Ncb, bgn, rv, Zc)
        G = int(C * N * scale) // (Qm * NL) * (Qm * NL) + Qm * NL * (C // 2)
        Er = RM.get_Er_ldpc(G, C, Qm, NL)
        dn = rng.integers(0, 2, (C, N)).astype("i1")
        dn[:, K - F - 2 * Zc:K - 2 * Zc] = -1
        dn[0, rng.integers(0, Ncb, 7)] = -1   # the reference skips ANY -1, wherever it is
        g = eng.ratematch_batch(torch.from_numpy(dn).cuda(), Ncb, Er, k0, Qm).cpu().numpy()
        ref = np.concatenate([oracle.ratematch_ldpc(dn[c], Ncb, Er[c], k0, Qm) for c in range(C)])
        assert np.array_equal(g, ref), (bgn, Zc)
        assert np.array_equal(eng.ratematch_batch(dn, Ncb, Er, k0, Qm), ref)
        llr = rng.normal(0, 5, sum(Er)).astype(np.float32)
        off = np.concatenate([[0], np.cumsum(Er)])
        want = np.stack([oracle.raterecover_ldpc(llr[off[c]:off[c + 1]].astype(np.float64), Ncb, N, k0, Qm, Zc, K - F, K) for c in range(C)])
        got = eng.raterecover_batch(torch.from_numpy(llr).cuda(), Er, Ncb, N, k0, Qm, Zc, K - F, K, out_f64=True)
        assert np.array_equal(got.cpu().numpy(), want), (bgn, Zc)
        got32 = eng.raterecover_batch(torch.from_numpy(llr).cuda(), Er, Ncb, N, k0, Qm, Zc, K - F, K, out_f64=False)
        assert np.array_equal(got32.cpu().numpy(), want.astype(np.float32))
        a, c = want, np.roll(want, 1, axis=1) * (rng.random(want.shape) > 0.3)
        assert np.array_equal(eng.harq_combine(torch.from_numpy(a).cuda(), torch.from_numpy(c).cuda()).cpu().numpy(), oracle.harq_combine(a, c))
        assert np.array_equal(eng.harq_combine(a, c), oracle.harq_combine(a, c))
    # round trip: what was rate matched comes back on its own position (no repetition, no noise)
    bgn, Zc, Qm = 1, 96, 4
    K, N = 22 * Zc, 66 * Zc
    dn = rng.integers(0, 2, (1, N)).astype("i1")
    E = 4000
    fe = eng.ratematch_batch(dn, N, [E], 0, Qm)
    rec = eng.raterecover_batch((1.0 - 2.0 * fe).astype(np.float32), [E], N, N, 0, Qm, Zc, K, K)
    assert np.array_equal(rec[0, :E], 1.0 - 2.0 * dn[0, :E]) and not rec[0, E:].any()


@pytest.mark.gpu
def test_cbsegment_golden(eng, rm_golden):
    from python_5gtoolbox_b200.ldpc import nr_ldpc_cbsegment
    n = 0
    while f"seg_in_{n}" in rm_golden:
        bgn, Zc = (int(x) for x in rm_golden[f"seg_meta_{n}"])
        cbs, z = nr_ldpc_cbsegment.ldpc_cbsegment(rm_golden[f"seg_in_{n}"], bgn)
        assert z == Zc and cbs.dtype == np.int8 and np.array_equal(cbs, rm_golden[f"seg_out_{n}"]), n
        n += 1
    assert n == 8
    with pytest.raises(AssertionError):
        nr_ldpc_cbsegment.ldpc_cbsegment(np.zeros(8449, "i1"), 1)   # B not divisible by C (ldpc_info.py:41)


@pytest.mark.gpu
def test_sch_chain_golden(eng, sch_golden):
    """DLSCHEncode / DLSCHDecode / ULSCH_* with the reference's signatures on the reference's inputs."""
    from python_5gtoolbox_b200.nr_pdsch import nr_dlsch, nr_dlsch_decode
    from python_5gtoolbox_b200.nr_pusch import nr_ulsch, nr_ulsch_decode
    assert len(sch_golden) == 9
    for name, d in sch_golden.items():
        A, R, Qm, NL, TBS_LBRM, G, ntx = (int(x) for x in d["meta"])
        cfg = {"L": int(d["cfg"][0]), "algo": ["min-sum", "BP", "BF"][int(d["cfg"][1])], "alpha": float(d["cfg"][2]), "beta": float(d["cfg"][3])}
        trblk = d["trblk"]
        cur = np.array([])
        for t, rv in enumerate(int(x) for x in d["rvs"]):
            if str(d["link"]) == "dl":
                g = nr_dlsch.DLSCHEncode(trblk, A, Qm, R, NL, rv, TBS_LBRM, G)
            else:
                cbs, Zc, bgn = nr_ulsch.ULSCH_Crc_CodeBlockSegment(trblk, A, R)
                assert np.array_equal(cbs, d["cbs"]), name
                g = nr_ulsch.ULSCH_encoding_ratematch(cbs, Zc, bgn, Qm, G, NL, rv)
            assert g.dtype == np.int8 and np.array_equal(g, d[f"g_{t}"]), (name, t)
            llr = d[f"llr_{t}"].astype(np.float64)
            if str(d["link"]) == "dl":
                st, tb, new = nr_dlsch_decode.DLSCHDecode(llr, A, Qm, R, NL, rv, TBS_LBRM, cfg, HARQ_on=ntx > 1, current_LLr_dns=cur)
            else:
                st, tb, new = nr_ulsch_decode.ULSCH_decoding(llr, A, R, Qm, G, NL, rv, cfg, HARQ_on=ntx > 1, current_LLr_dns=cur)
            assert new.dtype == np.float64 and np.array_equal(new, d[f"llrdn_{t}"]), (name, t)   # float64-exact
            assert bool(st) == bool(d[f"status_{t}"]), (name, t)
            assert tb.shape == d[f"tbblk_{t}"].shape
            if st:
                assert np.array_equal(tb, trblk) and np.array_equal(tb, d[f"tbblk_{t}"]), (name, t)
            else:   # a failed block: fp32 vs the reference's float64 may differ in a few of the wrong bits
                assert np.mean(tb != d[f"tbblk_{t}"]) < 0.02, (name, t)
            cur = new
