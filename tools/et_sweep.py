#!/usr/bin/env python3
"""Early-termination timing sweep of the BG1 Zc=384 decoder: python tools/et_sweep.py [B]  (NRLDPC_SO=... for a variant build)"""
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
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for snr in (1.0, 8.0):
    llr = engine.awgn_llr(dn, snr, seed=2)
    for et in (True, False):
        for L in (1, 2, 4, 8, 10, 20):
            ts = []
            for i in range(3):
                ev0.record()
                r = engine.decode_batch(llr, Zc, bgn, L, 0.8, 0.0, et, want_ck=False, want_info=True)
                ev1.record()
                torch.cuda.synchronize()
                ts.append(ev0.elapsed_time(ev1))
            print(f"snr={snr} et={int(et)} L={L}: {min(ts):.3f} ms ok={float(r['status'].float().mean()):.3f} "
                  f"iters={float(r['iters'].float().mean()):.2f}")
