#!/usr/bin/env python3
"""Early-termination decoder: dynamic codeblock queue vs fixed stride (run once plain, once with NRLDPC_STATIC_QUEUE=1).
python tools/et_queue_ab.py [B] [bgn] [Zc]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from python_5gtoolbox_b200 import engine  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
bgn = int(sys.argv[2]) if len(sys.argv) > 2 else 1
Zc = int(sys.argv[3]) if len(sys.argv) > 3 else 384
K, N, Nf, M = engine.dims(bgn, Zc)
ck = engine.random_bits(B, K, seed=1, device="cuda")
dn = engine.encode_batch(ck, bgn)
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
mode = "static" if os.environ.get("NRLDPC_STATIC_QUEUE") else "dynamic"
o = 0.0 if bgn == 1 else -2.6
for snr, L, alpha, beta in [(-0.6 + o, 32, 0.8, 0.3), (-0.45 + o, 32, 0.8, 0.3), (-0.3 + o, 32, 0.8, 0.3), (-0.15 + o, 32, 0.8, 0.3), (0.0 + o, 32, 0.8, 0.3), (0.2 + o, 32, 0.8, 0.3), (1.0 + o, 10, 0.8, 0.0)]:
    llr = engine.awgn_llr(dn, snr, seed=2)
    ts = []
    for i in range(4):
        ev0.record()
        r = engine.decode_batch(llr, Zc, bgn, L, alpha, beta, True, want_ck=False, want_info=True)
        ev1.record()
        torch.cuda.synchronize()
        ts.append(ev0.elapsed_time(ev1))
    it = r["iters"].float()
    print(f"{mode} BG{bgn} Zc={Zc} B={B} snr={snr:+.2f} L={L}: {min(ts):.3f} ms  {B * K / min(ts) / 1e6:.2f} Gbit/s  bler={1 - float(r['status'].float().mean()):.4f} "
          f"iters mean={float(it.mean()):.2f} std={float(it.std()):.2f}  checksum={int(r['info'].long().sum())}")
