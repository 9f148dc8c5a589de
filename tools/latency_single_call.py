#!/usr/bin/env python3
"""Per-call latency of the reference-signature functions on ONE codeblock / transport block (what an
unchanged reference script pays per call): python tools/latency_single_call.py"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from python_5gtoolbox_b200.ldpc import nr_ldpc_decode, nr_ldpc_encode  # noqa: E402
from python_5gtoolbox_b200.nr_pusch import nr_ulsch, nr_ulsch_decode  # noqa: E402


def timeit(f, n=30):
    f()
    f()
    t0 = time.perf_counter()
    for _ in range(n):
        f()
    return (time.perf_counter() - t0) / n * 1e3


rng = np.random.default_rng(0)
for bgn, Zc in [(1, 208), (1, 384), (2, 28)]:
    K = (22 if bgn == 1 else 10) * Zc
    ck = rng.integers(0, 2, K).astype("i1")
    dn = nr_ldpc_encode.encode_ldpc(ck.copy(), bgn)
    llr = (1 - 2 * dn.astype("f8")) * 4 + rng.normal(0, 2.0, dn.size)
    print(f"BG{bgn} Zc={Zc}: encode_ldpc {timeit(lambda: nr_ldpc_encode.encode_ldpc(ck.copy(), bgn)):.3f} ms, "
          f"nr_decode_ldpc(L=32, mixed) {timeit(lambda: nr_ldpc_decode.nr_decode_ldpc(llr, Zc, bgn, 32, 'min-sum', 0.8, 0.3)):.3f} ms")
# the PUSCH example's transport block: TBS 4224, 20 PRB MCS5 2 layers -> BG1, C=1, Zc=208
A, R, Qm, NL, G = 4224, 378, 4, 2, 20 * 12 * 8 * 4 * 2 // 2
trblk = rng.integers(0, 2, A).astype("i1")
cfg = {"L": 32, "algo": "min-sum", "alpha": 0.8, "beta": 0.3}
cbs, Zc, bgn = nr_ulsch.ULSCH_Crc_CodeBlockSegment(trblk, A, R)
g = nr_ulsch.ULSCH_encoding_ratematch(cbs, Zc, bgn, Qm, G, NL, 0)
llr = ((1 - 2 * g.astype("f8")) * 4 + rng.normal(0, 2.0, G)).astype("f4").astype("f8")
st, tb, _ = nr_ulsch_decode.ULSCH_decoding(llr, A, R, Qm, G, NL, 0, cfg)
assert st and np.array_equal(tb, trblk)
print(f"UL-SCH TB (A={A}, Zc={Zc}, C={cbs.shape[0]}, G={G}): CRC+segment {timeit(lambda: nr_ulsch.ULSCH_Crc_CodeBlockSegment(trblk, A, R)):.3f} ms, "
      f"encode+ratematch {timeit(lambda: nr_ulsch.ULSCH_encoding_ratematch(cbs.copy(), Zc, bgn, Qm, G, NL, 0)):.3f} ms, "
      f"ULSCH_decoding {timeit(lambda: nr_ulsch_decode.ULSCH_decoding(llr, A, R, Qm, G, NL, 0, cfg)):.3f} ms")
# the per-codeblock loop of scripts/sim_ldpc_decoder_bf.py (Zc = 10, BG1): test vector + bit-flipping decode per pass
from python_5gtoolbox_b200 import crc  # noqa: E402
np.random.seed(1)
blk, dn, llr = nr_ldpc_decode.for_test_5g_ldpc_encoder(10, 1, 4.0)
bits = rng.integers(0, 2, 196).astype("i1")
print(f"Zc=10 BG1: for_test_5g_ldpc_encoder {timeit(lambda: nr_ldpc_decode.for_test_5g_ldpc_encoder(10, 1, 4.0), 200):.3f} ms "
      f"(crc.nr_crc_encode {timeit(lambda: crc.nr_crc_encode(bits, '24A'), 200):.3f}, encode_ldpc {timeit(lambda: nr_ldpc_encode.encode_ldpc(blk.copy(), 1), 200):.3f}), "
      f"nr_decode_ldpc BF L=16 {timeit(lambda: nr_ldpc_decode.nr_decode_ldpc(llr, 10, 1, 16, 'BF'), 200):.3f} ms, "
      f"min-sum L=16 {timeit(lambda: nr_ldpc_decode.nr_decode_ldpc(llr, 10, 1, 16, 'min-sum', 0.8, 0.0), 200):.3f} ms")
