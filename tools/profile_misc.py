#!/usr/bin/env python3
"""One pass over the kernels either side of the decoder at BASELINE config #4 / #5 sizes, for ncu:
rate matching, rate recovery (stand-alone and fused into the decoder), HARQ combining, CB/TB CRC kernels, segmentation,
Philox bits, CRC attach, encoder, AWGN, error counters.  python tools/profile_misc.py"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from python_5gtoolbox_b200 import engine  # noqa: E402
from python_5gtoolbox_b200.nr_pdsch import nr_dlsch, nr_dlsch_decode  # noqa: E402
from python_5gtoolbox_b200.ldpc import ldpc_info, nr_ldpc_ratematch  # noqa: E402

rng = np.random.default_rng(4)
cfg = {"L": 10, "algo": "min-sum", "alpha": 0.8, "beta": 0.0}
A, R, Qm, NL, G, LBRM = 966896, 948, 8, 4, 273 * 12 * 12 * 8 * 4, 10 ** 9
trblk = rng.integers(0, 2, A).astype("i1")
for _ in range(2):
    g = nr_dlsch.DLSCHEncode(trblk, A, Qm, R, NL, 0, LBRM, G)          # tb_crc_partial, tb_segment, encode, ratematch
sigma = 10 ** (-6.5 / 20)
llr = (2 * ((1 - 2 * g.astype("f4")) + rng.normal(0, sigma, G).astype("f4")) / sigma ** 2).astype("f4")
for _ in range(2):
    st, tb, new = nr_dlsch_decode.DLSCHDecode(llr, A, Qm, R, NL, 0, LBRM, cfg)   # fused decoder (rate recovery inside), tb_finish
st2, tb2, new2 = nr_dlsch_decode.DLSCHDecode(llr, A, Qm, R, NL, 0, LBRM, cfg, HARQ_on=True, current_LLr_dns=new)   # + HARQ combining
assert st and st2 and np.array_equal(tb, trblk)
C, cbz, L, F, K, Zc = ldpc_info.get_cbs_info(A + 24, 1)
Er = nr_ldpc_ratematch.get_Er_ldpc(G, C, Qm, NL)
N = 66 * Zc
x = torch.from_numpy(llr).cuda()
for _ in range(2):
    soft = engine.raterecover_batch(x, Er, N, N, 0, Qm, Zc, cbz + L, K, out_f64=True)     # stand-alone rate recovery, float64 out
    f32 = engine.raterecover_batch(x, Er, N, N, 0, Qm, Zc, cbz + L, K, out_f64=False)
    comb = engine.harq_combine(soft, soft.clone())
# Monte-Carlo chain stages at 16384 codeblocks (BASELINE config #5)
import ctypes  # noqa: E402
from python_5gtoolbox_b200 import _lib  # noqa: E402
L_ = _lib.lib()
mm, K1 = 16384, 22 * 384
s = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
for _ in range(2):
    bits = torch.empty((mm, K1 - 24), dtype=torch.int8, device="cuda")
    _lib.check(L_.nrldpc_random_bits_rows(bits.data_ptr(), mm, K1 - 24, 1, 0, 1, s), "rb")
    blk = torch.empty((mm, K1), dtype=torch.int8, device="cuda")
    _lib.check(L_.nrldpc_crc_encode(bits.data_ptr(), mm, K1 - 24, 3, blk.data_ptr(), s), "crc")
    dn = engine.encode_batch(blk, 1, 384, fix_fillers=False)
    l2 = torch.empty(dn.shape, dtype=torch.float32, device="cuda")
    _lib.check(L_.nrldpc_awgn_llr_rows(dn.data_ptr(), mm, dn.shape[1], 1.0, 1, 0, 1, l2.data_ptr(), s), "awgn")
    r = engine.decode_batch(l2[:592], 384, 1, 10, 0.8, 0.0, True)
    c = engine.count_errors(blk[:592], r["ck"], K1, r["iters"])
torch.cuda.synchronize()
print("ok", c.tolist())
