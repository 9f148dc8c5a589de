"""Small invocation of every kernel family, meant to run under compute-sanitizer (memcheck / racecheck /
synccheck) on a GPU box:  compute-sanitizer --tool racecheck python tools/sanitizer_cases.py
(compute-sanitizer is closed on this round's pool -- rc 86 -- so only the plain run is recorded: it exits 0.)

Batches are tiny (the tools slow kernels down 10-100x); no result is checked here beyond the decoder
converging on noiseless-ish input -- parity is the job of tests/.
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch  # noqa: E402

from python_5gtoolbox_b200 import engine  # noqa: E402

dev = "cuda:0"
L = 6


def one(bgn, Zc, B, snr_db=2.0):
    K, N, Nf, M = engine.dims(bgn, Zc)
    ck = engine.random_bits(B, K, seed=Zc, device=dev)
    dn = engine.encode_batch(ck, bgn)
    llr = engine.awgn_llr(dn, snr_db, seed=Zc + 1)
    out = []
    for et in (True, False):
        r = engine.decode_batch(llr, Zc, bgn, L, 0.8, 0.0, et)
        out.append(int(r["status"].sum().item()))
    r = engine.decode_batch(llr, Zc, bgn, L, 0.8, 0.3, True, want_ck=False, want_info=True)
    bf = engine.decode_bf_batch(llr, Zc, bgn, 4)
    ref = engine.decode_ref_batch(llr[:1].cpu().numpy(), Zc, bgn, 2, "min-sum", 0.8, 0.0, True, f64=True)
    torch.cuda.synchronize()
    del bf, ref
    print(f"BG{bgn} Zc={Zc} B={B}: converged {out[0]}/{B} (early term), {out[1]}/{B} (fixed)", flush=True)


# spec kernel + word encoder/BF (384), spec kernel + table encoder/bf_qc (208), table-driven decoder (12, 2: many
# codeblocks per CTA), BG2 instances
for bgn, Zc, B in ((1, 384, 3), (2, 384, 2), (1, 208, 2), (2, 160, 2), (1, 96, 5), (1, 12, 37), (2, 2, 70)):
    one(bgn, Zc, B)

# mixed-Zc groups on side streams
groups = []
for bgn, Zc, B in ((1, 384, 2), (2, 352, 3), (1, 28, 9)):
    K, N, Nf, M = engine.dims(bgn, Zc)
    ck = engine.random_bits(B, K, seed=7 * Zc, device=dev)
    dn = engine.encode_batch(ck, bgn)
    groups.append((engine.awgn_llr(dn, 2.0, seed=Zc + 3), Zc, bgn))
engine.decode_groups(groups, L, 0.8, 0.0, True)
torch.cuda.synchronize()
print("mixed groups ok", flush=True)

# CRC attach / check on device
for poly, n in (("24A", 8424), ("24B", 3000), ("16", 500)):
    blk = engine.random_bits(5, n, seed=n, device=dev)
    enc = engine.crc_encode_device(blk, poly)
    chk = engine.crc_check_device(enc, poly)
    torch.cuda.synchronize()
print("crc ok", flush=True)

# round 2: sum-product kernel, fused transport-block chain (rate recovery inside the decoder, CB/TB CRC kernels, fused
# rate-matching store of the encoder) on a specialised and a table-driven lifting size, half-precision host LLRs
import numpy as np  # noqa: E402
from python_5gtoolbox_b200.nr_pdsch import nr_dlsch, nr_dlsch_decode  # noqa: E402

for bgn, Zc in ((1, 64), (2, 12)):
    K, N, Nf, M = engine.dims(bgn, Zc)
    ck = engine.random_bits(3, K, seed=5, device=dev)
    llr = engine.awgn_llr(engine.encode_batch(ck, bgn), 2.0, seed=6)
    engine.decode_bp_batch(llr, Zc, bgn, 4)
    engine.decode_batch(llr.cpu().numpy().astype(np.float16), Zc, bgn, 4, 0.8, 0.0, True)
torch.cuda.synchronize()
print("bp / f16 ok", flush=True)
rng = np.random.default_rng(1)
cfg = {"L": 6, "algo": "min-sum", "alpha": 0.8, "beta": 0.0}
for A, R, Qm, NL, G in ((30000, 700, 4, 2, 48000), (2000, 300, 2, 1, 9000)):
    trblk = rng.integers(0, 2, A).astype("i1")
    g = nr_dlsch.DLSCHEncode(trblk, A, Qm, R, NL, 0, 10 ** 9, G)
    x = (4.0 * (1 - 2 * g.astype("f4"))).astype("f4")
    st, tb, new = nr_dlsch_decode.DLSCHDecode(x, A, Qm, R, NL, 0, 10 ** 9, cfg)
    st2, tb2, new2 = nr_dlsch_decode.DLSCHDecode(x, A, Qm, R, NL, 0, 10 ** 9, cfg, HARQ_on=True, current_LLr_dns=new)
    assert st and st2 and np.array_equal(tb, trblk)
print("transport-block chain ok", flush=True)
