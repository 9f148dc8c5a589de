#!/usr/bin/env python3
"""Early-termination decoder, BG1 Zc=384 (and BG2 / other sizes), L = 10, NMS 0.8, at several SNRs: ms and Gbit/s of info bits.
python tools/et_ab.py [B]   (NRLDPC_SO=... for another build of the library)"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from python_5gtoolbox_b200 import _lib, engine  # noqa: E402

if os.environ.get("NRLDPC_SO"):
    _lib.SO_PATH = os.path.abspath(os.environ["NRLDPC_SO"])
B0 = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for bgn, Zc, scale in ((1, 384, 1.0), (2, 384, 1.0), (1, 256, 1.5), (1, 128, 2.0), (1, 64, 4.0)):
    B = int(B0 * scale)
    K, N, Nf, M = engine.dims(bgn, Zc)
    ck = engine.random_bits(B, K, seed=1, device="cuda")
    dn = engine.encode_batch(ck, bgn)
    for snr in ((-3.0, 0.0, 1.0, 2.0, 4.0) if bgn == 1 else (-3.0, -1.5, 0.0)):
        llr = engine.awgn_llr(dn, snr, seed=2)
        ts = []
        for i in range(4):
            ev0.record()
            r = engine.decode_batch(llr, Zc, bgn, 10, 0.8, 0.0, True, want_ck=False, want_info=True)
            ev1.record()
            torch.cuda.synchronize()
            ts.append(ev0.elapsed_time(ev1))
        t = min(ts[1:])
        chk = int(r["info"].to(torch.int64).sum()) & 0xffffffffffff
        print(f"BG{bgn} Zc={Zc} snr={snr:+.1f}: {t:8.3f} ms {B * K / t / 1e6:7.3f} Gbit/s  ok={float(r['status'].float().mean()):.3f} "
              f"iters={float(r['iters'].float().mean()):.2f} sum={chk:x} it_sum={int(r['iters'].sum())}")
