#!/usr/bin/env python3
"""Two launches of the quasi-cyclic sum-product kernel at BG1 Zc=384 (4 codeblocks per SM, +1 dB, L=10, early
termination), for ncu.  python tools/profile_bp.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from python_5gtoolbox_b200 import _lib, engine  # noqa: E402

if os.environ.get("NRLDPC_SO"):  # kernel experiments: time an alternative build of the library
    _lib.SO_PATH = os.path.abspath(os.environ["NRLDPC_SO"])

bgn, Zc, L, B = 1, 384, 10, 592
K, N, Nf, M = engine.dims(bgn, Zc)
ck = engine.random_bits(B, K, seed=1, device="cuda")
llr = engine.awgn_llr(engine.encode_batch(ck, bgn), 1.0, seed=2)
for _ in range(2):
    c, s, it = engine.decode_bp_batch(llr, Zc, bgn, L)
torch.cuda.synchronize()
print("ok", float(s.float().mean()), float(it.float().mean()))
