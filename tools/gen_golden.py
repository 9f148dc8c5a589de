#!/usr/bin/env python3
"""Generate tests/golden/*.npz by running the UNMODIFIED reference (xu753x/python_5gtoolbox) live.

    python tools/gen_golden.py [/root/reference] [--jobs 6] [--quick]

The reference is pure Python/NumPy and imports as-is when cwd is its root (tables are opened by
relative path, py5gphy/ldpc/ldpc_info.py:110).  /root/reference does not exist on the GPU box, so
the outputs are committed as small fixtures; this script is the record of how they were made.

Files written
  tests/golden/encode_golden.npz  all 51 lifting sizes x {BG1,BG2} x 2 codeblocks (random payload,
                                  random filler count):  nr_ldpc_encode.encode_ldpc(ck, bgn)
  tests/golden/decode_golden.npz  nr_ldpc_decode.nr_decode_ldpc on float32-representable LLRs from
                                  for_test_5g_ldpc_encoder: ck, status and the iteration count
                                  (calls to _min_sum_process / _BP_process divided by M)
  tests/golden/crc_golden.npz     crc.nr_crc_encode for the six polynomials
"""
import argparse
import multiprocessing as mp
import os
import sys
import time

import numpy as np

ZLIST = [2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16, 18, 20, 22, 24, 26, 28, 30, 32, 36, 40, 44, 48, 52,
         56, 60, 64, 72, 80, 88, 96, 104, 112, 120, 128, 144, 160, 176, 192, 208, 224, 240, 256, 288, 320, 352, 384]

REF = "/root/reference"


def _enter_ref():
    os.chdir(REF)
    if REF not in sys.path:
        sys.path.insert(0, REF)


def dims(bgn, Zc):
    return ((22, 66, 68, 46) if bgn == 1 else (10, 50, 52, 42)), Zc


def enc_case(args):
    bgn, Zc, seed = args
    _enter_ref()
    from py5gphy.ldpc import nr_ldpc_encode
    rng = np.random.default_rng(seed)
    K = (22 if bgn == 1 else 10) * Zc
    ck = rng.integers(0, 2, K).astype("i1")
    F = int(rng.integers(0, Zc)) if seed % 2 else 0
    if F:
        ck[K - F:] = -1
    ck_in = ck.copy()
    dn = nr_ldpc_encode.encode_ldpc(ck, bgn)
    return bgn, Zc, F, ck_in, ck, dn


def dec_case(args):
    bgn, Zc, snr, L, algo, alpha, beta, seed = args
    _enter_ref()
    from py5gphy.ldpc import nr_ldpc_decode
    cnt = [0]
    for name in ("_min_sum_process", "_BP_process"):
        orig = getattr(nr_ldpc_decode, name)

        def wrap(*a, _o=orig, **k):
            cnt[0] += 1
            return _o(*a, **k)
        setattr(nr_ldpc_decode, name, wrap)
    np.random.seed(seed)
    K = (22 if bgn == 1 else 10) * Zc
    crcpoly = "24A" if K > 24 + 8 else "16"  # BG2 Zc=2,3 have K=20,30: too short for a 24-bit CRC
    blk, dn, llr = nr_ldpc_decode.for_test_5g_ldpc_encoder(Zc, bgn, snr, crcpoly)
    llr32 = llr.astype("f4")
    t0 = time.time()
    b, ck, st = nr_ldpc_decode.nr_decode_ldpc(llr32.astype("f8"), Zc, bgn, L, algo, alpha, beta)
    M = (46 if bgn == 1 else 42) * Zc
    iters = cnt[0] // M if algo != "BF" else -1
    assert algo == "BF" or cnt[0] % M == 0
    print(f"  dec bgn{bgn} Zc{Zc} snr{snr} L{L} {algo} a{alpha} b{beta}: status={st} iters={iters} "
          f"biterr={int((np.asarray(b) != blk).sum())} {time.time() - t0:.1f}s", flush=True)
    return args, blk.astype("i1"), llr32, np.asarray(ck).astype("i1"), bool(st), iters


def dec_cases(quick):
    cases = []
    seed = 1000
    small = [2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 15, 16, 18, 22, 26, 30, 32, 36, 40, 52, 64]
    params = [(10, 0.8, 0.0), (16, 1.0, 0.5), (32, 0.8, 0.3), (10, 1.0, 0.0), (8, 0.7, 0.0), (16, 0.7, 0.3)]
    for bgn in (1, 2):
        base = 0.0 if bgn == 1 else -2.5
        for zi, Zc in enumerate(small):
            for t in range(2):
                L, a, b = params[(zi * 2 + t) % len(params)]
                snr = base + [-0.5, 0.5, 1.5][(zi + t) % 3]
                cases.append((bgn, Zc, snr, L, "min-sum", a, b, seed)); seed += 1
    for bgn in (1, 2):  # BP and BF, small sizes only
        for Zc in (2, 5, 8, 12, 16):
            cases.append((bgn, Zc, 0.5 if bgn == 1 else -2.0, 10, "BP", 1, 0, seed)); seed += 1
            cases.append((bgn, Zc, 4.5 if bgn == 1 else 3.0, 16, "BF", 1, 0, seed)); seed += 1
    if not quick:
        for bgn in (1, 2):
            base = -0.5 if bgn == 1 else -3.0
            for zi, Zc in enumerate([96, 128, 176, 208, 240, 256, 320, 352]):
                L, a, b = params[zi % len(params)]
                cases.append((bgn, Zc, base + 0.5 * (zi % 3), L, "min-sum", a, b, seed)); seed += 1
        cases += [(1, 384, -3.0, 10, "min-sum", 0.8, 0.0, 5001), (1, 384, 1.0, 10, "min-sum", 0.8, 0.0, 5002),
                  (1, 384, -0.6, 10, "min-sum", 0.7, 0.0, 5003), (1, 384, -0.2, 16, "min-sum", 0.8, 0.3, 5004),
                  (2, 384, -3.4, 10, "min-sum", 0.8, 0.0, 5005), (2, 384, -2.8, 10, "min-sum", 1.0, 0.5, 5006)]
    return cases


def main():
    global REF
    ap = argparse.ArgumentParser()
    ap.add_argument("ref", nargs="?", default=REF)
    ap.add_argument("--jobs", type=int, default=6)
    ap.add_argument("--quick", action="store_true")
    a = ap.parse_args()
    REF = a.ref
    repo = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = os.path.join(repo, "tests", "golden")
    os.makedirs(out, exist_ok=True)

    # CRC
    _enter_ref()
    from py5gphy.crc import crc
    rng = np.random.default_rng(77)
    d = {}
    for poly in ["6", "11", "16", "24A", "24B", "24C"]:
        for n in (1, 31, 100, 1000):
            blk = rng.integers(0, 2, n).astype("i1")
            d[f"in_{poly}_{n}"] = blk
            d[f"out_{poly}_{n}"] = crc.nr_crc_encode(blk, poly)
    np.savez_compressed(os.path.join(out, "crc_golden.npz"), **d)

    with mp.Pool(a.jobs) as pool:
        res = pool.map(enc_case, [(bgn, Zc, 10 * Zc + bgn * 2 + t) for bgn in (1, 2) for Zc in ZLIST for t in range(2)])
        d = {}
        for n, (bgn, Zc, F, ck_in, ck_after, dn) in enumerate(res):
            d[f"meta_{n}"] = np.array([bgn, Zc, F])
            d[f"ck_{n}"] = ck_in
            d[f"ckafter_{n}"] = ck_after
            d[f"dn_{n}"] = dn
        np.savez_compressed(os.path.join(out, "encode_golden.npz"), **d)
        print("encode goldens:", len(res), flush=True)

        cases = dec_cases(a.quick)
        big = [c for c in cases if c[1] >= 320]
        rest = [c for c in cases if c[1] < 320]
        res = pool.map(dec_case, rest, chunksize=1)
    with mp.Pool(min(a.jobs, 4)) as pool:  # ~8 GB RSS each at Zc=384 BG1
        res += pool.map(dec_case, big, chunksize=1)
    d = {}
    for n, (args, blk, llr32, ck, st, iters) in enumerate(res):
        bgn, Zc, snr, L, algo, alpha, beta, seed = args
        d[f"cfg_{n}"] = np.array([bgn, Zc, snr, L, {"min-sum": 0, "BP": 1, "BF": 2}[algo], alpha, beta, seed], "f8")
        d[f"blk_{n}"] = np.packbits(blk)
        d[f"llr_{n}"] = llr32
        d[f"ck_{n}"] = np.packbits(ck)
        d[f"res_{n}"] = np.array([int(st), iters])
    np.savez_compressed(os.path.join(out, "decode_golden.npz"), **d)
    print("decode goldens:", len(res))


if __name__ == "__main__":
    main()
