#!/usr/bin/env python3
"""Bit-flipping decoder: word-parallel kernel (Zc % 32 == 0) vs bf_qc_kernel (NRLDPC_BF_NO_WORDS=1), timing + checksums.
python tools/bf_words_ab.py [B]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from python_5gtoolbox_b200 import engine  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
mode = "bf_qc" if os.environ.get("NRLDPC_BF_NO_WORDS") else "bf_words"
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for bgn, Zc, snr, L in [(1, 384, 7.0, 20), (1, 384, 10.0, 20), (2, 384, 6.0, 20), (1, 256, 9.0, 20), (2, 64, 8.0, 30), (1, 32, 9.0, 30), (1, 96, 9.0, 20)]:
    K, N, Nf, M = engine.dims(bgn, Zc)
    ck = engine.random_bits(B, K, seed=1, device="cuda")
    dn = engine.encode_batch(ck, bgn, Zc)
    llr = engine.awgn_llr(dn, snr, seed=2)
    ts = []
    for i in range(4):
        ev0.record()
        r = engine.decode_bf_batch(llr, Zc, bgn, L)
        ev1.record()
        torch.cuda.synchronize()
        ts.append(ev0.elapsed_time(ev1))
    c, s, it = r
    wsum = int((c.long() * (torch.arange(c.shape[1], device=c.device) % 251 + 1)).sum())
    print(f"{mode} BG{bgn} Zc={Zc} snr={snr} L={L} B={B}: {min(ts):.3f} ms  {B * K / min(ts) / 1e6:.1f} Gbit/s info  ok={float(s.float().mean()):.4f} "
          f"iters={float(it.float().mean()):.3f}  ck_checksum={wsum} it_sum={int(it.sum())} st_sum={int(s.sum())}")
