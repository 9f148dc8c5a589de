#!/usr/bin/env python3
"""Bit-packed encoder workload for ncu / timing.  python tools/profile_encode_packed.py [B] [reps] [bgn] [Zc]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from python_5gtoolbox_b200 import engine  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 18
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
bgn = int(sys.argv[3]) if len(sys.argv) > 3 else 1
Zc = int(sys.argv[4]) if len(sys.argv) > 4 else 384
kb, nb = (22, 66) if bgn == 1 else (10, 50)
ck = engine.random_bits_packed(B, kb * Zc, seed=1, device="cuda")
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for i in range(reps):
    e0.record()
    dn = engine.encode_packed(ck, bgn, Zc)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print(f"encode_packed {i}: {ms:.3f} ms, {B * kb * Zc / ms / 1e6:.1f} Gbit/s info, {B * (kb + nb) * Zc / 8 / ms / 1e6:.1f} GB/s packed in+out "
          f"(BG{bgn} Zc={Zc}, {B} codeblocks)")
