#!/usr/bin/env python3
"""Where the time of one DLSCHDecode call goes (BASELINE config #4, 115 codeblocks of BG1 Zc=384): the Python mirror, the
C-ABI call alone, and -- with NRLDPC_TRACE=1 -- the stages inside the C call.  python tools/tb_latency_probe.py"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from python_5gtoolbox_b200 import engine, sch  # noqa: E402
from python_5gtoolbox_b200.ldpc import ldpc_info, nr_ldpc_ratematch  # noqa: E402
from python_5gtoolbox_b200.nr_pdsch import nr_dlsch, nr_dlsch_decode  # noqa: E402

rng = np.random.default_rng(4)
cfg = {"L": 10, "algo": "min-sum", "alpha": 0.8, "beta": 0.0}
A, R, Qm, NL, G = 966896, 948, 8, 4, 273 * 12 * 12 * 8 * 4
TBS_LBRM = 10 ** 9
trblk = rng.integers(0, 2, A).astype("i1")
g = nr_dlsch.DLSCHEncode(trblk, A, Qm, R, NL, 0, TBS_LBRM, G)
sigma = 10 ** (-6.5 / 20)
llr = (2 * ((1 - 2 * g.astype("f4")) + rng.normal(0, sigma, G).astype("f4")) / sigma ** 2).astype("f4")


def timeit(f, reps=20):
    for _ in range(4):   # like the timed loop: the previous result stays alive during the next call (pinned pool warm-up)
        r = f()
    t0 = time.perf_counter()
    for _ in range(reps):
        r = f()
    return (time.perf_counter() - t0) / reps * 1e3, r


ms, r = timeit(lambda: nr_dlsch_decode.DLSCHDecode(llr, A, Qm, R, NL, 0, TBS_LBRM, cfg))
print(f"DLSCHDecode (Python mirror, pageable float32 LLRs): {ms:.3f} ms  ok={r[0]}")
C, cbz, L, F, K, Zc = ldpc_info.get_cbs_info(A + 24, 1)
Er = nr_ldpc_ratematch.get_Er_ldpc(G, C, Qm, NL)
ms, _ = timeit(lambda: (ldpc_info.get_cbs_info(A + 24, 1), nr_ldpc_ratematch.get_Er_ldpc(G, C, Qm, NL)))
print(f"  parameter arithmetic (get_cbs_info + get_Er_ldpc): {ms:.3f} ms")
ms, rr = timeit(lambda: engine.sch_decode_host(llr, Er, 1, Zc, 66 * Zc, 0, Qm, cbz + L, A, 10, 0.8, 0.0))
print(f"  engine.sch_decode_host, pageable LLRs: {ms:.3f} ms   mean iterations {rr['iters'].mean():.2f}")
pin = engine.pinned_empty(llr.shape, np.float32)
pin[...] = llr
ms, _ = timeit(lambda: engine.sch_decode_host(pin, Er, 1, Zc, 66 * Zc, 0, Qm, cbz + L, A, 10, 0.8, 0.0))
print(f"  engine.sch_decode_host, pinned LLRs: {ms:.3f} ms")
ms, _ = timeit(lambda: engine.sch_decode_host(pin, Er, 1, Zc, 66 * Zc, 0, Qm, cbz + L, A, 10, 0.8, 0.0, want_soft=False))
print(f"  engine.sch_decode_host, pinned LLRs, no soft buffer out: {ms:.3f} ms")
import ctypes  # noqa: E402
from python_5gtoolbox_b200 import _lib  # noqa: E402
E = np.ascontiguousarray(Er, np.int32)
soft = engine.pinned_empty((C, 66 * Zc), np.float64)
tb = engine.pinned_empty((A,), np.int8)
small, iters = np.empty(1 + 2 * C, np.uint8), np.empty(C, np.int32)
Lb = _lib.lib()
raw = lambda so: Lb.nrldpc_sch_decode_host(pin.ctypes.data, 0, C, 1, Zc, 66 * Zc, 0, Qm, cbz + L, E.ctypes.data, None, so, 10, 0.8, 0.0, A,
                                           tb.ctypes.data, small.ctypes.data, small.ctypes.data + 1, small.ctypes.data + 1 + C, iters.ctypes.data)
ms, _ = timeit(lambda: raw(soft.ctypes.data))
print(f"  raw ctypes call, preallocated pinned buffers: {ms:.3f} ms")
ms, _ = timeit(lambda: raw(None))
print(f"  raw ctypes call, no soft buffer: {ms:.3f} ms")
ms, _ = timeit(lambda: engine.pinned_empty((C, 66 * Zc), np.float64))
print(f"  pinned_empty of the soft buffer: {ms:.4f} ms")
ms, _ = timeit(lambda: nr_dlsch.DLSCHEncode(trblk, A, Qm, R, NL, 0, TBS_LBRM, G))
print(f"DLSCHEncode: {ms:.3f} ms")
