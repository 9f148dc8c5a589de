#!/usr/bin/env python3
"""Encoder throughput across lifting sizes (device-resident int8 bits, 65 536 codeblocks): python tools/bench_encode_sweep.py [bgn:Zc ...]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from python_5gtoolbox_b200 import engine  # noqa: E402

cases = [tuple(int(x) for x in a.split(":")) for a in sys.argv[1:]] or [(1, 384), (1, 240), (1, 208), (1, 176), (1, 144), (1, 112), (1, 96),
                                                                         (1, 80), (1, 48), (1, 40), (2, 384), (2, 208)]
for bgn, Zc in cases:
    K, N, Nf, M = engine.dims(bgn, Zc)
    B = 1 << 16
    ck = engine.random_bits(B, K, seed=1, device="cuda")
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(5):   # warm-up: clocks ramp up during the first launches of a process
        engine.encode_batch(ck, bgn, Zc, fix_fillers=False)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(5):
        dn = engine.encode_batch(ck, bgn, Zc, fix_fillers=False)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(f"encode BG{bgn} Zc={Zc:3d}: {ms:7.3f} ms  {B * (K + N) / ms / 1e9:5.2f} TB/s int8 in+out = {B * (K + N) / ms / 1e6 / 6552.3:.2f} of the HBM peak; "
          f"packed (8d) bytes: {B * (K + N) / 8 / ms / 1e6 / 6552.3:.3f}")
