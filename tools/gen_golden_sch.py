#!/usr/bin/env python3
"""Generate tests/golden/ratematch_golden.npz and tests/golden/sch_golden.npz by running the UNMODIFIED
reference (xu753x/python_5gtoolbox) live -- the callers either side of the LDPC path.

    python tools/gen_golden_sch.py [--jobs 6]

  ratematch_golden.npz  nr_ldpc_ratematch.{get_Er_ldpc, get_k0, ratematch_ldpc} and
                        nr_ldpc_raterecover.raterecover_ldpc on random parameters (fillers, LBRM-limited
                        circular buffers, every rv, puncturing and repetition), ldpc_cbsegment
  sch_golden.npz        DLSCHEncode / DLSCHDecode and ULSCH_Crc_CodeBlockSegment + ULSCH_encoding_ratematch /
                        ULSCH_decoding on whole transport blocks: BG1 and BG2, one and several codeblocks,
                        repetition, LBRM, HARQ retransmission with soft combining, BPSK/AWGN LLRs rounded to
                        float32 (the reference's demappers produce float32, SURVEY 8(a) precision note)
"""
import argparse
import multiprocessing as mp
import os
import sys

import numpy as np

REF = "/root/reference"


def _enter_ref():
    os.chdir(REF)
    if REF not in sys.path:
        sys.path.insert(0, REF)


def rm_case(seed):
    _enter_ref()
    from py5gphy.ldpc import nr_ldpc_ratematch as RM, nr_ldpc_raterecover as RR
    rng = np.random.default_rng(seed)
    bgn = int(rng.integers(1, 3))
    Zc = int(rng.choice([2, 3, 5, 8, 12, 20, 36, 52, 96]))
    K = (22 if bgn == 1 else 10) * Zc
    N = (66 if bgn == 1 else 50) * Zc
    F = int(rng.integers(0, Zc)) if seed % 3 else 0
    K_apo = K - F
    dn = rng.integers(0, 2, N).astype("i1")
    if F:
        dn[K_apo - 2 * Zc:K - 2 * Zc] = -1
    Ncb = N if seed % 2 else int(rng.integers(max(K, N // 2), N + 1))
    rv = int(rng.integers(0, 4))
    k0 = RM.get_k0(Ncb, bgn, rv, Zc)
    Qm = int(rng.choice([1, 2, 4, 6, 8]))
    E = Qm * int(rng.integers(max(1, N // (4 * Qm)), (5 * N) // (2 * Qm) + 1))
    fe = RM.ratematch_ldpc(dn, Ncb, E, k0, Qm)
    llr = rng.normal(0, 4, E).astype("f4").astype("f8")
    if seed % 5 == 0:
        llr[rng.integers(0, E, 4)] = 0
    out = RR.raterecover_ldpc(llr, Ncb, N, k0, Qm, Zc, K_apo, K)
    return np.array([bgn, Zc, F, Ncb, rv, k0, Qm, E]), dn, fe, llr.astype("f4"), out


TB_CASES = [
    # name, link, A, coderateby1024, Qm, NL, rv list, TBS_LBRM (None = unlimited), G, snr_db, decoder config
    ("dl_bg2_small", "dl", 120, 300, 2, 1, [0], None, 480, 2.0, dict(L=8, algo="min-sum", alpha=0.8, beta=0.0)),
    ("dl_bg2_rep", "dl", 200, 200, 2, 1, [0], None, 3000, -4.0, dict(L=8, algo="min-sum", alpha=0.8, beta=0.3)),
    ("dl_bg2_mid", "dl", 1000, 500, 4, 1, [0], None, 2400, 3.0, dict(L=8, algo="min-sum", alpha=0.8, beta=0.0)),
    ("dl_bg1_lbrm_harq", "dl", 4000, 800, 4, 2, [0, 2], 5000, 5200, 2.5, dict(L=6, algo="min-sum", alpha=0.8, beta=0.0)),
    ("dl_bg1_2cb", "dl", 9000, 700, 6, 1, [0], None, 13500, 7.0, dict(L=6, algo="min-sum", alpha=0.8, beta=0.0)),
    ("ul_bg2_harq", "ul", 552, 400, 2, 2, [0, 3], None, 1600, 1.0, dict(L=8, algo="min-sum", alpha=0.7, beta=0.0)),
    ("ul_bg1_2cb", "ul", 8500, 750, 4, 1, [0], None, 11600, 4.0, dict(L=6, algo="min-sum", alpha=1.0, beta=0.5)),
    ("ul_bg2_bf", "ul", 96, 250, 2, 1, [0], None, 600, 6.0, dict(L=8, algo="BF", alpha=1, beta=0)),
    ("dl_bg2_bp", "dl", 150, 350, 2, 1, [0], None, 500, 2.0, dict(L=6, algo="BP", alpha=1, beta=0)),
]


def tb_case(case):
    name, link, A, R, Qm, NL, rvs, lbrm, G, snr, cfg = case
    _enter_ref()
    from py5gphy.nr_pdsch import nr_dlsch, nr_dlsch_decode
    from py5gphy.nr_pusch import nr_ulsch, nr_ulsch_decode
    rng = np.random.default_rng(sum(ord(c) for c in name))
    trblk = rng.integers(0, 2, A).astype("i1")
    TBS_LBRM = lbrm if lbrm is not None else 10 ** 9
    d = {"meta": np.array([A, R, Qm, NL, TBS_LBRM, G, len(rvs)], "i8"), "link": np.array(link), "trblk": trblk,
         "cfg": np.array([cfg["L"], {"min-sum": 0, "BP": 1, "BF": 2}[cfg["algo"]], cfg["alpha"], cfg["beta"]], "f8"),
         "rvs": np.array(rvs)}
    cur = np.array([])
    sigma = 10 ** (-snr / 20)
    for t, rv in enumerate(rvs):
        if link == "dl":
            g = nr_dlsch.DLSCHEncode(trblk, A, Qm, R, NL, rv, TBS_LBRM, G)
        else:
            cbs, Zc, bgn = nr_ulsch.ULSCH_Crc_CodeBlockSegment(trblk, A, R)
            d["cbs"] = cbs.copy()
            g = nr_ulsch.ULSCH_encoding_ratematch(cbs, Zc, bgn, Qm, G, NL, rv)
        llr = (2 * ((1 - 2 * g.astype("f8")) + rng.normal(0, sigma, G)) / sigma ** 2).astype("f4")
        if link == "dl":
            st, tb, new = nr_dlsch_decode.DLSCHDecode(llr.astype("f8"), A, Qm, R, NL, rv, TBS_LBRM, cfg, HARQ_on=len(rvs) > 1,
                                                      current_LLr_dns=cur)
        else:
            st, tb, new = nr_ulsch_decode.ULSCH_decoding(llr.astype("f8"), A, R, Qm, G, NL, rv, cfg, HARQ_on=len(rvs) > 1,
                                                         current_LLr_dns=cur)
        cur = new
        d[f"g_{t}"] = g
        d[f"llr_{t}"] = llr
        d[f"status_{t}"] = np.array(bool(st))
        d[f"tbblk_{t}"] = np.asarray(tb)
        d[f"llrdn_{t}"] = new
    print("tb case", name, "done", [bool(d[f"status_{t}"]) for t in range(len(rvs))], flush=True)
    return name, d


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--jobs", type=int, default=6)
    a = ap.parse_args()
    out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
    with mp.Pool(a.jobs) as pool:
        res = pool.map(rm_case, range(100, 148))
        d = {}
        for n, (meta, dn, fe, llr, rec) in enumerate(res):
            d[f"meta_{n}"], d[f"dn_{n}"], d[f"fe_{n}"], d[f"llr_{n}"], d[f"rec_{n}"] = meta, dn, fe, llr, rec
        # segmentation
        _enter_ref()
        from py5gphy.ldpc import nr_ldpc_cbsegment
        rng = np.random.default_rng(5)
        for n, (B, bgn) in enumerate([(40, 2), (300, 2), (3840, 2), (3848, 2), (8448, 1), (8472, 1), (25296, 1), (1000, 1)]):
            bits = rng.integers(0, 2, B).astype("i1")
            cbs, Zc = nr_ldpc_cbsegment.ldpc_cbsegment(bits, bgn)
            d[f"seg_in_{n}"], d[f"seg_meta_{n}"], d[f"seg_out_{n}"] = bits, np.array([bgn, Zc]), cbs
        np.savez_compressed(os.path.join(out, "ratematch_golden.npz"), **d)
        print("ratematch goldens:", len(res), flush=True)
        res = pool.map(tb_case, TB_CASES, chunksize=1)
    d = {}
    for name, dd in res:
        for k, v in dd.items():
            d[f"{name}__{k}"] = v
    np.savez_compressed(os.path.join(out, "sch_golden.npz"), **d)
    print("sch goldens:", len(res))


if __name__ == "__main__":
    main()
