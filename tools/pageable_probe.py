#!/usr/bin/env python3
"""Host-buffer decode (nrldpc_decode_minsum_host) from pinned vs pageable memory: Gbit/s of info bits and GB/s of LLRs.
Environment knobs of the staging path: NRLDPC_COPY_THREADS, NRLDPC_COPY_NT, NRLDPC_STAGE_MB."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from python_5gtoolbox_b200 import engine  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
bgn, Zc = 1, 384
K, N, Nf, M = engine.dims(bgn, Zc)
ck = engine.random_bits(B, K, seed=5, device="cuda")
llr = engine.awgn_llr(engine.encode_batch(ck, bgn), 1.0, seed=6).cpu().numpy()
pin = engine.pinned_empty(llr.shape, np.float32)
pin[...] = llr
out = {}
for name, src in (("pinned", pin), ("pageable", llr)):
    engine.decode_batch(src, Zc, bgn, 10, 0.8, 0.0, False, want_ck=False, want_info=True)
    t0 = time.perf_counter()
    for _ in range(3):
        engine.decode_batch(src, Zc, bgn, 10, 0.8, 0.0, False, want_ck=False, want_info=True)
    dt = (time.perf_counter() - t0) / 3
    out[name] = B * K / dt / 1e9
    print(f"{name:9s}: {out[name]:.2f} Gbit/s info, {B * N * 4 / dt / 1e9:.1f} GB/s of LLRs", end="   ")
print(f"ratio {out['pageable'] / out['pinned']:.2f}  threads={os.environ.get('NRLDPC_COPY_THREADS', 'auto')} "
      f"nt={os.environ.get('NRLDPC_COPY_NT', '1')} stage_mb={os.environ.get('NRLDPC_STAGE_MB', '8')}")
