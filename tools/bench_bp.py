#!/usr/bin/env python3
"""Sum-product ('BP') decoder at BG1 Zc=384, L=10, early termination, +1 dB: the quasi-cyclic kernel on device-resident
LLRs against the generic CSR kernel behind its host entry point (what algo='BP' ran on in round 1).  python tools/bench_bp.py"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from python_5gtoolbox_b200 import _lib, engine  # noqa: E402

if os.environ.get("NRLDPC_SO"):  # kernel experiments: time an alternative build of the library
    _lib.SO_PATH = os.path.abspath(os.environ["NRLDPC_SO"])

bgn, Zc, L = 1, 384, 10
K, N, Nf, M = engine.dims(bgn, Zc)
for snr in (1.0, -3.0):
    B = 4096
    ck = engine.random_bits(B, K, seed=1, device="cuda")
    llr = engine.awgn_llr(engine.encode_batch(ck, bgn), snr, seed=2)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    engine.decode_bp_batch(llr, Zc, bgn, L)
    torch.cuda.synchronize()
    ev0.record()
    c, s, it = engine.decode_bp_batch(llr, Zc, bgn, L)
    ev1.record()
    torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1)
    Bg = 128
    h = llr[:Bg].cpu().numpy().astype(np.float64)
    engine.decode_ref_batch(h[:8], Zc, bgn, L, "BP", 1, 0, True, f64=True)
    t0 = time.perf_counter()
    cg, sg, ig = engine.decode_ref_batch(h, Zc, bgn, L, "BP", 1, 0, True, f64=True)
    tg = time.perf_counter() - t0
    same = np.array_equal(cg, c[:Bg].cpu().numpy()) and np.array_equal(ig, it[:Bg].cpu().numpy())
    print(f"snr {snr:+.1f} dB: QC BP kernel {B} codeblocks in {ms:.2f} ms = {B * K / ms / 1e6:.3f} Gbit/s info (mean iters {float(it.float().mean()):.2f}, "
          f"ok {float(s.float().mean()):.3f}); generic CSR path {Bg} codeblocks in {tg * 1e3:.1f} ms = {Bg * K / tg / 1e9:.4f} Gbit/s; "
          f"ratio {(B / ms * 1e3) / (Bg / tg):.0f}x; identical outputs on the common codeblocks: {same}")
