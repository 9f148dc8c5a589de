#!/usr/bin/env python3
"""profile_decode.py for any (bgn, Zc): B codeblocks, 10 iterations at +1 dB, for ncu.
python tools/profile_decode_zc.py BGN ZC [B] [launches] [early_term]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from python_5gtoolbox_b200 import engine  # noqa: E402

bgn, Zc = int(sys.argv[1]), int(sys.argv[2])
B = int(sys.argv[3]) if len(sys.argv) > 3 else 592
n = int(sys.argv[4]) if len(sys.argv) > 4 else 2
et = bool(int(sys.argv[5])) if len(sys.argv) > 5 else False
K, N, Nf, M = engine.dims(bgn, Zc)
ck = engine.random_bits(B, K, seed=1, device="cuda")
llr = engine.awgn_llr(engine.encode_batch(ck, bgn), 1.0, seed=2)
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for i in range(n):
    ev0.record()
    r = engine.decode_batch(llr, Zc, bgn, 10, 0.8, 0.0, et, want_ck=False, want_info=True)
    ev1.record()
    torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1)
    print(f"BG{bgn} Zc={Zc} launch {i}: {ms:.3f} ms, {B * K / ms / 1e6:.3f} Gbit/s, ok={float(r['status'].float().mean()):.3f}")
