#!/usr/bin/env python3
"""Encoder workload for ncu / timing: B codeblocks of (bgn, Zc).  python tools/profile_encode.py [B] [reps] [bgn] [Zc]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from python_5gtoolbox_b200 import engine  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
bgn = int(sys.argv[3]) if len(sys.argv) > 3 else 1
Zc = int(sys.argv[4]) if len(sys.argv) > 4 else 384
kb, nb = (22, 66) if bgn == 1 else (10, 50)
ck = engine.random_bits(B, kb * Zc, seed=1, device="cuda")
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for i in range(reps):
    e0.record()
    dn = engine.encode_batch(ck, bgn)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print(f"encode {i}: {ms:.3f} ms, {B * kb * Zc / ms / 1e6:.1f} Gbit/s info, {B * (kb + nb) * Zc / ms / 1e6:.1f} GB/s int8 in+out (BG{bgn} Zc={Zc})")
