#!/usr/bin/env python3
"""Early-termination decoder throughput at an operating point (+1 dB BG1 / -1.5 dB BG2, codeblocks leave at different
iterations) for the given sizes: python tools/et_plus1_sweep.py bgn:Zc ..."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from python_5gtoolbox_b200 import engine  # noqa: E402

for a in sys.argv[1:]:
    bgn, Zc = (int(x) for x in a.split(":"))
    K, N, Nf, M = engine.dims(bgn, Zc)
    B = max(256, min(1 << 16, (1 << 31) // (N * 16)))
    ck = engine.random_bits(B, K, seed=1, device="cuda")
    llr = engine.awgn_llr(engine.encode_batch(ck, bgn), 1.0 if bgn == 1 else -1.5, seed=2)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e9
    for i in range(3):
        e0.record()
        r = engine.decode_batch(llr, Zc, bgn, 10, 0.8, 0.0, True, want_ck=False, want_info=True)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    print(f"BG{bgn} Zc={Zc:3d} B={B:6d}: {best:8.3f} ms  {B * K / best / 1e6:7.3f} Gbit/s info  ok={float(r['status'].float().mean()):.3f}"
          f"  iters={float(r['iters'].float().mean()):.2f}")
