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


# ------------------------------------------------------------------ the fused transport-block entry points (round 2)

def _tb_case(rng, bgn, Zc, C, Qm, NL, rv, ncb_frac, rate_scale, F):
    """A synthetic transport block laid out like get_cbs_info would: C codeblocks of cbz payload bits."""
    from python_5gtoolbox_b200.ldpc import nr_ldpc_ratematch as RM
    K, N = ((22, 66) if bgn == 1 else (10, 50))
    K, N = K * Zc, N * Zc
    Lcb = 24 if C > 1 else 0
    K_apo = K - F
    cbz = K_apo - Lcb
    B = C * cbz
    Ltb = 24 if B - 24 > 3824 else 16
    A = B - Ltb
    assert (A > 3824) == (Ltb == 24)
    Ncb = N if ncb_frac == 1.0 else int(N * ncb_frac)
    k0 = RM.get_k0(Ncb, bgn, rv, Zc)
    G = int(C * N * rate_scale) // (Qm * NL) * (Qm * NL) + Qm * NL * (C // 2)
    Er = RM.get_Er_ldpc(G, C, Qm, NL)
    return dict(bgn=bgn, Zc=Zc, C=C, Qm=Qm, K=K, N=N, K_apo=K_apo, cbz=cbz, A=A, Ltb=Ltb, Ncb=Ncb, k0=k0, Er=Er, G=G)


_TB_CASES = [  # bgn, Zc, C, Qm, NL, rv, Ncb fraction, E/N scale, fillers
    (1, 384, 5, 8, 4, 0, 1.0, 0.40, 16),      # specialised kernel, high rate, fillers
    (1, 384, 3, 2, 1, 2, 0.85, 1.30, 0),      # LBRM + repetition (E > Ncb)
    (1, 208, 4, 6, 1, 0, 1.0, 0.55, 40),      # specialised kernel, 16 * odd lifting size
    (2, 144, 2, 4, 2, 1, 1.0, 0.70, 8),       # BG2 specialised
    (2, 64, 6, 2, 2, 3, 0.9, 0.60, 20),       # table-driven kernel, several codeblocks per CTA
    (1, 12, 9, 1, 1, 0, 1.0, 0.90, 3),        # Zc < 32: several codeblocks per warp; Qm = 1
    (2, 10, 1, 2, 1, 0, 1.0, 1.00, 5),        # single codeblock: no CB CRC, 16-bit TB CRC
]


def _make_tb(eng, oracle, rng, t, snr_db, first=None):
    """Random transport block -> (trblk, g, llr float32).  Encoded by the oracle-checked chain of engine calls."""
    from python_5gtoolbox_b200 import crc
    trblk = rng.integers(0, 2, t["A"]).astype("i1") if first is None else first
    blk = crc.nr_crc_encode(trblk, '24A' if t["Ltb"] == 24 else '16')
    cbs = np.full((t["C"], t["K"]), -1, "i1")
    if t["C"] == 1:
        cbs[0, :t["cbz"]] = blk
    else:
        cbs[:, :t["K_apo"]] = crc.nr_crc_encode_batch(blk.reshape(t["C"], t["cbz"]), '24B')
    dn = eng.encode_batch(cbs.copy(), t["bgn"], t["Zc"])
    g = np.concatenate([oracle.ratematch_ldpc(dn[c], t["Ncb"], t["Er"][c], t["k0"], t["Qm"]) for c in range(t["C"])])
    sigma = 10 ** (-snr_db / 20)
    llr = (2 * ((1 - 2 * g.astype("f8")) + rng.normal(0, sigma, g.size)) / sigma ** 2).astype("f4")
    return trblk, cbs, g, llr


@pytest.mark.gpu
@pytest.mark.parametrize("case", _TB_CASES)
def test_fused_sch_decode_equals_staged_chain(eng, oracle, case):
    """nrldpc_sch_decode_host (rate recovery + HARQ combining inside the decoder's LLR load, CB/TB CRC kernel) against
    the same chain done stage by stage: oracle rate recovery / combining (float64-exact), the fp32 decoder on the
    rounded soft buffer, oracle CRCs."""
    rng = np.random.default_rng(hash(case) & 0xffff)
    t = _tb_case(rng, *case)
    bgn, Zc, C, N, K = t["bgn"], t["Zc"], t["C"], t["N"], t["K"]
    snr = 7.0 if case[7] < 0.6 else 2.0   # E/N = 0.4 is a rate-0.8 code
    trblk, cbs, g, llr = _make_tb(eng, oracle, rng, t, snr)
    off = np.concatenate([[0], np.cumsum(t["Er"])])
    cur = None
    for tx, dt in enumerate((np.float32, np.float64, np.float32)):   # 1st tx, then two HARQ combinations
        x = llr.astype(dt) * (1.0 if tx == 0 else rng.uniform(0.5, 1.5))
        want = np.stack([oracle.raterecover_ldpc(x[off[c]:off[c + 1]].astype(np.float64), t["Ncb"], N, t["k0"], t["Qm"], Zc,
                                                 t["K_apo"], K) for c in range(C)])
        if cur is not None:
            want = oracle.harq_combine(want, cur)
        r = eng.sch_decode_host(x, t["Er"], bgn, Zc, t["Ncb"], t["k0"], t["Qm"], t["K_apo"], t["A"], 12, 0.8, 0.1, cur=cur)
        assert r["soft"].dtype == np.float64 and np.array_equal(r["soft"], want), (case, tx)      # float64-exact
        d = eng.decode_batch(want.astype(np.float32), Zc, bgn, 12, 0.8, 0.1, True)
        assert np.array_equal(r["status"], d["status"]) and np.array_equal(r["iters"], d["iters"]), (case, tx)
        tb = d["ck"][:, :t["cbz"]].reshape(-1)
        assert np.array_equal(r["tbblk"], tb[:t["A"]]), (case, tx)
        _, tb_err = oracle.crc_decode(tb, '24A' if t["Ltb"] == 24 else '16')
        assert r["tb_err"] == tb_err, (case, tx)
        if C > 1:
            cb_err = [oracle.crc_decode(d["ck"][c, :t["K_apo"]], '24B')[1] for c in range(C)]
            assert r["cb_err"].tolist() == cb_err, (case, tx)
        else:
            assert r["cb_err"].tolist() == [0]
        if tx == 0:
            if case[5] == 0:   # rv 0 carries the systematic bits: a clean first transmission decodes
                assert r["tb_err"] == 0 and np.array_equal(r["tbblk"], trblk), case
            soft_sep = eng.sch_recover_host(x, t["Er"], bgn, Zc, t["Ncb"], t["k0"], t["Qm"], t["K_apo"])
            assert np.array_equal(soft_sep, want)
        else:
            assert np.array_equal(eng.sch_recover_host(x, t["Er"], bgn, Zc, t["Ncb"], t["k0"], t["Qm"], t["K_apo"], cur=cur), want)
        cur = np.array(want)   # pageable copy: the next call stages it through the pinned ring
    # a corrupted codeblock: TB CRC and that codeblock's CRC fail, the others stay clean
    bad = llr.copy()
    bad[off[C - 1]:off[C]] = rng.normal(0, 1, t["Er"][C - 1])
    r = eng.sch_decode_host(bad, t["Er"], bgn, Zc, t["Ncb"], t["k0"], t["Qm"], t["K_apo"], t["A"], 12, 0.8, 0.1)
    assert r["tb_err"] == 1 and not r["status"][C - 1]
    if C > 1 and case[5] == 0:   # (a first transmission with rv != 0 has no systematic bits: nothing decodes)
        assert r["cb_err"][C - 1] == 1 and not r["cb_err"][:C - 1].any()


@pytest.mark.gpu
def test_fused_sch_decode_device_entry(eng, oracle):
    """nrldpc_sch_decode on device buffers and a caller's stream, ck handed back."""
    import ctypes
    import torch
    from python_5gtoolbox_b200 import _lib
    rng = np.random.default_rng(12)
    t = _tb_case(rng, 1, 384, 4, 4, 2, 0, 1.0, 0.5, 24)
    trblk, cbs, g, llr = _make_tb(eng, oracle, rng, t, 6.0)   # E/N = 0.5: a rate-2/3 code
    C, N, Nf = t["C"], t["N"], t["N"] + 2 * t["Zc"]
    dev = torch.device("cuda")
    E = torch.tensor(t["Er"], dtype=torch.int32, device=dev)
    goff = torch.tensor(np.concatenate([[0], np.cumsum(t["Er"])[:-1]]), dtype=torch.int64, device=dev)
    x = torch.from_numpy(llr).to(dev)
    soft = torch.empty((C, N), dtype=torch.float64, device=dev)
    ck = torch.empty((C, Nf), dtype=torch.int8, device=dev)
    tb = torch.empty(t["A"], dtype=torch.int8, device=dev)
    flags = torch.full((1 + 2 * C,), 7, dtype=torch.uint8, device=dev)
    iters = torch.empty(C, dtype=torch.int32, device=dev)
    st = torch.cuda.Stream()
    st.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(st):
        _lib.check(_lib.lib().nrldpc_sch_decode(x.data_ptr(), 0, C, 1, 384, t["Ncb"], t["k0"], t["Qm"], t["K_apo"], E.data_ptr(),
                                                goff.data_ptr(), None, soft.data_ptr(), 10, 0.8, 0.0, t["A"], ck.data_ptr(),
                                                tb.data_ptr(), flags.data_ptr(), flags.data_ptr() + 1, flags.data_ptr() + 1 + C,
                                                iters.data_ptr(), ctypes.c_void_p(st.cuda_stream)), "sch_decode")
    st.synchronize()
    assert flags[0].item() == 0 and np.array_equal(tb.cpu().numpy(), trblk)
    assert not flags[1:1 + C].any() and flags[1 + C:].all()
    want = eng.raterecover_batch(llr, t["Er"], t["Ncb"], N, t["k0"], t["Qm"], 384, t["K_apo"], t["K"], out_f64=True)
    assert np.array_equal(soft.cpu().numpy(), want)
    assert np.array_equal(ck.cpu().numpy()[:, :t["K_apo"]], cbs[:, :t["K_apo"]])


@pytest.mark.gpu
@pytest.mark.parametrize("case", _TB_CASES)
def test_fused_sch_encode_equals_staged_chain(eng, oracle, case):
    rng = np.random.default_rng(7 + (hash(case) & 0xfff))
    t = _tb_case(rng, *case)
    trblk, cbs, g, _ = _make_tb(eng, oracle, rng, t, 5.0)
    seg = eng.sch_segment_host(trblk, t["C"], t["K"])
    assert seg.dtype == np.int8 and np.array_equal(seg, cbs), case                      # TB CRC + segmentation + CB CRC
    work = cbs.copy()
    g2 = eng.encode_ratematch_host(work, t["bgn"], t["Zc"], t["Ncb"], t["k0"], t["Qm"], t["Er"])
    assert np.array_equal(g2, g), case
    fixed = cbs.copy()
    fixed[:, 2 * t["Zc"]:][fixed[:, 2 * t["Zc"]:] == -1] = 0
    assert np.array_equal(work, fixed), case                                            # encode_ldpc's in-place side effect
    g3 = eng.sch_encode_host(trblk, t["C"], t["bgn"], t["Zc"], t["Ncb"], t["k0"], t["Qm"], t["Er"])
    assert np.array_equal(g3, g), case
    with pytest.raises(AssertionError):                                                  # crc.nr_crc_encode asserts 0/1 input (:18-20)
        bad = trblk.copy()
        bad[t["A"] // 2] = 2
        eng.sch_encode_host(bad, t["C"], t["bgn"], t["Zc"], t["Ncb"], t["k0"], t["Qm"], t["Er"])


@pytest.mark.gpu
def test_pinned_pool_and_pageable_host_path(eng):
    """Pinned blocks are recycled; the host-buffer decoder gives the same results from pageable NumPy memory (staged by
    the copy threads) as from pinned memory, also for batches of several pipeline chunks."""
    a = eng.pinned_empty((3, 1000), np.float64)
    p = a.ctypes.data
    a[:] = 1.5
    del a
    b = eng.pinned_empty((3000,), np.float64)
    assert b.ctypes.data == p                      # same size class -> the block came back from the pool
    import torch
    bgn, Zc, B = 1, 384, 700                       # 700 codeblocks = 71 MB of LLRs: three 32 MiB pipeline chunks
    K, N, Nf, M = eng.dims(bgn, Zc)
    ck = eng.random_bits(B, K, seed=5, device="cuda")
    dn = eng.encode_batch(ck, bgn)
    llr_d = eng.awgn_llr(dn, 1.0, seed=6)
    llr = llr_d.cpu().numpy()                      # pageable
    pin = eng.pinned_empty(llr.shape, np.float32)
    pin[...] = llr
    r_dev = eng.decode_batch(llr_d, Zc, bgn, 10, 0.8, 0.0, True, want_info=True)
    for src in (llr, pin):
        r = eng.decode_batch(src, Zc, bgn, 10, 0.8, 0.0, True, want_info=True)
        assert np.array_equal(r["ck"], r_dev["ck"].cpu().numpy())
        assert np.array_equal(r["info"].view(np.int32), r_dev["info"].cpu().numpy())
        assert np.array_equal(r["iters"], r_dev["iters"].cpu().numpy())
        assert np.array_equal(r["status"], r_dev["status"].cpu().numpy().astype(bool))


@pytest.mark.gpu
def test_early_termination_queue_under_graph_capture(eng):
    """The dynamic codeblock queue of the early-termination kernels keeps host-side state per launch: a launch captured
    into a CUDA graph must not use it (a replay would reuse the ticket base).  Replays and interleaved eager launches
    give the eager results."""
    import torch
    bgn, Zc, B = 1, 384, 600                      # > 2 x 148 codeblocks: the eager launch uses the ticket queue
    K, N, Nf, M = eng.dims(bgn, Zc)
    ck = eng.random_bits(B, K, seed=9, device="cuda")
    llr = eng.awgn_llr(eng.encode_batch(ck, bgn), 0.7, seed=10)
    ref = eng.decode_batch(llr, Zc, bgn, 10, 0.8, 0.0, True)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    st = torch.cuda.Stream()
    st.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(st):
        eng.decode_batch(llr, Zc, bgn, 10, 0.8, 0.0, True)   # warm-up outside the capture (function attributes)
        st.synchronize()
        with torch.cuda.graph(g, stream=st):
            out = eng.decode_batch(llr, Zc, bgn, 10, 0.8, 0.0, True)
    for rep in range(3):
        for k in out:
            if out[k] is not None:
                out[k].fill_(3)
        g.replay()
        eager = eng.decode_batch(llr, Zc, bgn, 10, 0.8, 0.0, True)   # shares the device with the replay
        torch.cuda.synchronize()
        for k in ("ck", "status", "iters"):
            assert torch.equal(out[k], ref[k]), (rep, k)
            assert torch.equal(eager[k], ref[k]), (rep, k)


@pytest.mark.gpu
def test_half_precision_host_llrs(eng):
    """nrldpc_decode_minsum_host_f16: half-precision LLRs over the host link, widened on the device -- the results are those of
    the fp32 entry point on the widened values, for batches with a tail that is not a multiple of 8 values."""
    import torch
    for bgn, Zc, B in [(1, 384, 350), (2, 7, 33)]:
        K, N, Nf, M = eng.dims(bgn, Zc)
        ck = eng.random_bits(B, K, seed=15, device="cuda")
        llr = eng.awgn_llr(eng.encode_batch(ck, bgn, Zc), 1.5, seed=16).cpu().numpy()
        h = llr.astype(np.float16)
        want = eng.decode_batch(h.astype(np.float32), Zc, bgn, 10, 0.8, 0.0, True, want_info=True)
        got = eng.decode_batch(h, Zc, bgn, 10, 0.8, 0.0, True, want_info=True)
        for k in ("ck", "info", "status", "iters"):
            assert np.array_equal(got[k], want[k]), (bgn, Zc, k)


@pytest.mark.gpu
def test_sch_entry_points_reject_bad_arguments_and_fall_back_to_pageable(eng, monkeypatch):
    """Inconsistent transport-block layouts raise AssertionError (the reference asserts in get_cbs_info / reshape); beyond
    the pinned-memory cap results come back in pageable memory and are the same."""
    rng = np.random.default_rng(3)
    t = _tb_case(rng, 1, 384, 3, 4, 1, 0, 1.0, 0.5, 8)
    llr = rng.normal(0, 4, int(sum(t["Er"]))).astype("f4")
    with pytest.raises(AssertionError):   # K_apo does not match A / C
        eng.sch_decode_host(llr, t["Er"], 1, 384, t["Ncb"], t["k0"], t["Qm"], t["K_apo"] + 1, t["A"], 5, 0.8, 0.0)
    with pytest.raises(AssertionError):   # E not a multiple of Qm
        bad = list(t["Er"]); bad[0] += 1; bad[1] -= 1
        eng.sch_decode_host(llr, bad, 1, 384, t["Ncb"], t["k0"], t["Qm"], t["K_apo"], t["A"], 5, 0.8, 0.0)
    with pytest.raises(AssertionError):   # k0 outside the circular buffer
        eng.sch_decode_host(llr, t["Er"], 1, 384, t["Ncb"], t["Ncb"], t["Qm"], t["K_apo"], t["A"], 5, 0.8, 0.0)
    with pytest.raises(AssertionError):   # not a 5G lifting size
        eng.sch_decode_host(llr, t["Er"], 1, 385, t["Ncb"], t["k0"], t["Qm"], t["K_apo"], t["A"], 5, 0.8, 0.0)
    a = eng.sch_decode_host(llr, t["Er"], 1, 384, t["Ncb"], t["k0"], t["Qm"], t["K_apo"], t["A"], 5, 0.8, 0.0)
    from python_5gtoolbox_b200 import engine
    monkeypatch.setattr(engine, "_PINNED_LIMIT", 0)
    b = eng.sch_decode_host(llr, t["Er"], 1, 384, t["Ncb"], t["k0"], t["Qm"], t["K_apo"], t["A"], 5, 0.8, 0.0)
    for k in ("tbblk", "soft", "status", "iters", "cb_err"):
        assert np.array_equal(a[k], b[k]), k
    assert a["tb_err"] == b["tb_err"]


@pytest.mark.gpu
def test_ticket_queue_survives_ring_wraparound(eng):
    """The early-termination kernels draw a per-launch ticket counter from a ring of 4096 whose start values the host tracks
    (no device state is reset between launches).  More launches than the ring has slots, on two streams: every launch
    must decode every codeblock exactly once -- identical outputs throughout."""
    import torch
    bgn, Zc, B = 1, 384, 2 * 148 + 17               # >= 2 codeblocks per persistent CTA: the ticket queue is in use
    K, N, Nf, M = eng.dims(bgn, Zc)
    ck = eng.random_bits(B, K, seed=31, device="cuda")
    llr = eng.awgn_llr(eng.encode_batch(ck, bgn), 2.0, seed=32)
    ref = eng.decode_batch(llr, Zc, bgn, 6, 0.8, 0.0, True, want_ck=False, want_info=True)
    torch.cuda.synchronize()
    want = (ref["info"].clone(), ref["iters"].clone(), ref["status"].clone())
    side = torch.cuda.Stream()
    last = []
    for k in range(4300):
        if k % 2:
            with torch.cuda.stream(side):
                r = eng.decode_batch(llr, Zc, bgn, 6, 0.8, 0.0, True, want_ck=False, want_info=True)
        else:
            r = eng.decode_batch(llr, Zc, bgn, 6, 0.8, 0.0, True, want_ck=False, want_info=True)
        if k % 430 == 0 or k >= 4296:
            last.append(r)
    torch.cuda.synchronize()
    for r in last:
        assert torch.equal(r["info"], want[0]) and torch.equal(r["iters"], want[1]) and torch.equal(r["status"], want[2])


@pytest.mark.gpu
def test_dlsch_decode_without_soft_buffer(eng, sch_golden):
    """The keyword-only extra soft_buffer=False (default True = the reference's return value): same status and bits."""
    from python_5gtoolbox_b200.nr_pdsch import nr_dlsch_decode
    n = 0
    for name, d in sch_golden.items():
        if str(d["link"]) != "dl" or ["min-sum", "BP", "BF"][int(d["cfg"][1])] != "min-sum":
            continue
        A, R, Qm, NL, TBS_LBRM, G, ntx = (int(x) for x in d["meta"])
        cfg = {"L": int(d["cfg"][0]), "algo": "min-sum", "alpha": float(d["cfg"][2]), "beta": float(d["cfg"][3])}
        llr = d["llr_0"].astype(np.float64)
        rv = int(d["rvs"][0])
        st, tb, new = nr_dlsch_decode.DLSCHDecode(llr, A, Qm, R, NL, rv, TBS_LBRM, cfg)
        st2, tb2, new2 = nr_dlsch_decode.DLSCHDecode(llr, A, Qm, R, NL, rv, TBS_LBRM, cfg, soft_buffer=False)
        assert st2 == st and np.array_equal(tb2, tb) and new2.size == 0 and new.size > 0
        n += 1
    assert n >= 2


@pytest.mark.gpu
def test_fused_sch_chain_fuzz(eng, oracle):
    """Random transport-block layouts through the fused entry points against the staged chain (oracle rate matching /
    recovery / combining, the fp32 decoder on the rounded soft buffer): every lifting-size family, Qm, redundancy version,
    limited buffers that cut into the fillers, heavy repetition, single codeblocks, HARQ combining."""
    from python_5gtoolbox_b200.ldpc import nr_ldpc_ratematch as RM
    from tests.conftest import ZLIST
    rng = np.random.default_rng(20261019)
    n = 0
    for trial in range(160):
        bgn = int(rng.integers(1, 3))
        Zc = int(rng.choice(ZLIST if trial % 4 else [384, 352, 208, 144, 128, 72, 28, 12]))
        if Zc > 128 and trial % 4:
            Zc = int(rng.choice([z for z in ZLIST if z <= 128]))   # keep most trials small
        C = int(rng.integers(1, 6))
        Qm = int(rng.choice([1, 2, 4, 6, 8]))
        NL = int(rng.integers(1, 3))
        rv = int(rng.integers(0, 4))
        K, N = ((22, 66) if bgn == 1 else (10, 50))
        K, N = K * Zc, N * Zc
        Lcb = 24 if C > 1 else 0
        F = int(rng.integers(0, max(1, min(2 * Zc, K - Lcb - 2 * Zc - 32))))
        K_apo = K - F
        cbz = K_apo - Lcb
        B = C * cbz
        if B <= 40:
            continue
        Ltb = 24 if B - 24 > 3824 else 16
        A = B - Ltb
        if (A > 3824) != (Ltb == 24) or A < 1:
            continue
        ncb_frac = float(rng.choice([1.0, 1.0, 0.9, 0.6, 0.35]))
        Ncb = N if ncb_frac == 1.0 else max(int(N * ncb_frac), 4 * Zc)
        k0 = RM.get_k0(Ncb, bgn, rv, Zc)
        scale = float(rng.choice([0.3, 0.6, 1.0, 1.7, 3.2]))
        G = max(1, int(C * N * scale) // (Qm * NL)) * (Qm * NL) + Qm * NL * int(rng.integers(0, C))
        Er = RM.get_Er_ldpc(G, C, Qm, NL)
        if min(Er) <= 0 or max(Er) >= 1 << 24:
            continue
        t = dict(bgn=bgn, Zc=Zc, C=C, Qm=Qm, K=K, N=N, K_apo=K_apo, cbz=cbz, A=A, Ltb=Ltb, Ncb=Ncb, k0=k0, Er=Er, G=G)
        trblk, cbs, g, llr = _make_tb(eng, oracle, rng, t, 4.0)
        # transmit side: TB CRC + segmentation + encoder with the rate matcher in its store
        assert np.array_equal(eng.sch_segment_host(trblk, C, K), cbs), t
        assert np.array_equal(eng.sch_encode_host(trblk, C, bgn, Zc, Ncb, k0, Qm, Er), g), t
        # receive side, first transmission then one combination
        off = np.concatenate([[0], np.cumsum(Er)])
        cur = None
        for tx in range(2):
            x = llr if tx == 0 else (llr * rng.uniform(0.5, 1.5)).astype(np.float64)
            want = np.stack([oracle.raterecover_ldpc(np.asarray(x[off[c]:off[c + 1]], np.float64), Ncb, N, k0, Qm, Zc, K_apo, K) for c in range(C)])
            if cur is not None:
                want = oracle.harq_combine(want, cur)
            r = eng.sch_decode_host(x, Er, bgn, Zc, Ncb, k0, Qm, K_apo, A, 8, 0.75, 0.0, cur=cur)
            assert np.array_equal(r["soft"], want), (t, tx)
            d = eng.decode_batch(want.astype(np.float32), Zc, bgn, 8, 0.75, 0.0, True)
            assert np.array_equal(r["status"], d["status"]) and np.array_equal(r["iters"], d["iters"]), (t, tx)
            tb = d["ck"][:, :cbz].reshape(-1)
            assert np.array_equal(r["tbblk"], tb[:A]), (t, tx)
            assert r["tb_err"] == oracle.crc_decode(tb, '24A' if Ltb == 24 else '16')[1], (t, tx)
            cur = np.array(want)
        n += 1
    assert n >= 100
