#!/usr/bin/env python3
"""Encoder workload for ncu / timing: BG1 Zc=384, B codeblocks.  python tools/profile_encode.py [B] [reps]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from python_5gtoolbox_b200 import engine  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
bgn, Zc = 1, 384
ck = engine.random_bits(B, 22 * Zc, seed=1, device="cuda")
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for i in range(reps):
    e0.record()
    dn = engine.encode_batch(ck, bgn)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print(f"encode {i}: {ms:.3f} ms, {B * 22 * Zc / ms / 1e6:.1f} Gbit/s info, {B * (22 + 66) * Zc / ms / 1e6:.1f} GB/s int8 in+out")
