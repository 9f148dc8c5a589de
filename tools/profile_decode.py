#!/usr/bin/env python3
"""Small fixed workload for ncu: BG1 Zc=384, B codeblocks, 10 iterations, no early termination
(the bench's step at reduced batch).  python tools/profile_decode.py [B] [launches] [early_term] [snr]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from python_5gtoolbox_b200 import _lib, engine  # noqa: E402

if os.environ.get("NRLDPC_SO"):  # kernel experiments: time an alternative build of the library
    _lib.SO_PATH = os.path.abspath(os.environ["NRLDPC_SO"])

B = int(sys.argv[1]) if len(sys.argv) > 1 else 592
n = int(sys.argv[2]) if len(sys.argv) > 2 else 3
et = bool(int(sys.argv[3])) if len(sys.argv) > 3 else False
snr = float(sys.argv[4]) if len(sys.argv) > 4 else 1.0
bgn, Zc = 1, 384
ck = engine.random_bits(B, 22 * Zc, seed=1, device="cuda")
dn = engine.encode_batch(ck, bgn)
llr = engine.awgn_llr(dn, snr, seed=2)
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for i in range(n):
    ev0.record()
    r = engine.decode_batch(llr, Zc, bgn, 10, 0.8, 0.0, et, want_ck=False, want_info=True)
    ev1.record()
    torch.cuda.synchronize()
    print(f"launch {i}: {ev0.elapsed_time(ev1):.3f} ms, {B * 8448 / ev0.elapsed_time(ev1) / 1e6:.3f} Gbit/s, "
          f"ok={float(r['status'].float().mean()):.3f} iters={float(r['iters'].float().mean()):.2f} "
          f"sum={int(r['info'].to(torch.int64).sum()) & 0xffffffffffff:x}")
