#!/usr/bin/env python3
"""Does the early-termination kernel slow down when the persistent CTAs fall out of lockstep?  Same decoder call on
(a) B different codeblocks (CTAs leave the iteration loop at different times) and (b) one codeblock replicated B times
(every CTA does the same number of iterations at the same time), early termination on, BG1 Zc=384, +1 dB.
python tools/et_lockstep_probe.py [B]   (NRLDPC_SO=... for a variant build)"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from python_5gtoolbox_b200 import _lib, engine  # noqa: E402

if os.environ.get("NRLDPC_SO"):
    _lib.SO_PATH = os.path.abspath(os.environ["NRLDPC_SO"])
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
bgn, Zc = 1, 384
ck = engine.random_bits(B, 22 * Zc, seed=1, device="cuda")
dn = engine.encode_batch(ck, bgn)
llr = engine.awgn_llr(dn, 1.0, seed=2)
r = engine.decode_batch(llr, Zc, bgn, 10, 0.8, 0.0, True, want_ck=False, want_info=True)
it = r["iters"]
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
cases = [("mixed", llr)]
for target in (6, 7, 8, 9):
    idx = (it == target).nonzero()
    if idx.numel():
        cases.append((f"replicated(it={target})", llr[int(idx[0])].unsqueeze(0).repeat(B, 1).contiguous()))
for name, x in cases:
    ts = []
    for i in range(3):
        ev0.record()
        r = engine.decode_batch(x, Zc, bgn, 10, 0.8, 0.0, True, want_ck=False, want_info=True)
        ev1.record()
        torch.cuda.synchronize()
        ts.append(ev0.elapsed_time(ev1))
    mi = float(r["iters"].float().mean())
    print(f"{name:20s} {min(ts):8.3f} ms  iters={mi:.2f}  ms per (iteration+1 check pass)={min(ts) / (mi + 1):.3f}")
