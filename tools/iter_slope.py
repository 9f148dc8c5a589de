#!/usr/bin/env python3
"""Per-iteration time and fixed per-codeblock cost of the fixed-iteration decoder: time(L) for several L, least squares.
python tools/iter_slope.py [B]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from python_5gtoolbox_b200 import _lib, engine  # noqa: E402

if os.environ.get("NRLDPC_SO"):
    _lib.SO_PATH = os.path.abspath(os.environ["NRLDPC_SO"])
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
bgn, Zc = 1, 384
ck = engine.random_bits(B, 22 * Zc, seed=1, device="cuda")
llr = engine.awgn_llr(engine.encode_batch(ck, bgn), 1.0, seed=2)
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
pts = []
for L in (1, 2, 5, 10, 15, 20):
    best = 1e9
    for rep in range(3):
        ev0.record()
        engine.decode_batch(llr, Zc, bgn, L, 0.8, 0.0, False, want_ck=False, want_info=True)
        ev1.record()
        torch.cuda.synchronize()
        best = min(best, ev0.elapsed_time(ev1))
    pts.append((L, best))
    print(f"L={L:2d}: {best:8.3f} ms  = {best * 1e3 * 148 / B:7.2f} us per codeblock and SM")
n = len(pts); sx = sum(p[0] for p in pts); sy = sum(p[1] for p in pts); sxx = sum(p[0] ** 2 for p in pts); sxy = sum(p[0] * p[1] for p in pts)
slope = (n * sxy - sx * sy) / (n * sxx - sx * sx); icpt = (sy - slope * sx) / n
print(f"per iteration {slope * 1e3 * 148 / B:.3f} us, fixed {icpt * 1e3 * 148 / B:.3f} us per codeblock ({icpt / (icpt + 10 * slope) * 100:.1f} % of a 10-iteration decode)")
