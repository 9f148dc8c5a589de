#!/usr/bin/env python3
"""Decoder throughput across lifting sizes (device-resident LLRs, 10 fixed iterations, NMS alpha=0.8):
python tools/bench_zc_sweep.py [--et] [bgn:Zc ...]
--et: the reference's early termination at -3 dB, where nothing converges (10 iterations + 11 syndrome checks): the mode
the specialised kernels for Zc < 144 are built for (fixed-iteration runs go to the table-driven kernel there)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from python_5gtoolbox_b200 import engine  # noqa: E402

ET = "--et" in sys.argv
cases = [tuple(int(x) for x in a.split(":")) for a in sys.argv[1:] if a != "--et"] or [(1, 384), (1, 352), (1, 320), (1, 208), (1, 176), (1, 72),
                                                                         (1, 40), (1, 12), (1, 2), (2, 384), (2, 288), (2, 208), (2, 28), (2, 8)]
for bgn, Zc in cases:
    K, N, Nf, M = engine.dims(bgn, Zc)
    B = max(256, min(1 << 16, (1 << 31) // (N * 16)))
    ck = engine.random_bits(B, K, seed=1, device="cuda")
    dn = engine.encode_batch(ck, bgn)
    llr = engine.awgn_llr(dn, -3.0 if ET else (1.0 if bgn == 1 else -1.5), seed=2)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e9
    for i in range(3):
        e0.record()
        r = engine.decode_batch(llr, Zc, bgn, 10, 0.8, 0.0, ET, want_ck=False, want_info=True)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    G, nt, smem = engine.decode_geometry(bgn, Zc)
    print(f"BG{bgn} Zc={Zc:3d} B={B:6d} cbs/CTA={G:2d} threads={nt:4d} smem={smem:6d}: {best:8.3f} ms  {B * K / best / 1e6:7.3f} Gbit/s info  "
          f"{B * (316 if bgn == 1 else 197) * Zc * 10 / best / 1e6:8.1f} G edge-iter/s  ok={float(r['status'].float().mean()):.3f}"
          f"  iters={float(r['iters'].float().mean()):.2f}{'  (early termination on)' if ET else ''}")
