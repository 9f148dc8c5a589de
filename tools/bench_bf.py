"""Bit-flipping decoder throughput: quasi-cyclic shared-memory kernel (nrldpc_decode_bf) vs the generic CSR kernel
(nrldpc_decode_bf_csr_host) on the same BG1 Zc=384 batch.  Run on a GPU box:  python tools/bench_bf.py [B]"""
import json, sys, time
import numpy as np, torch
sys.path.insert(0, ".")
from python_5gtoolbox_b200 import engine

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
out = []
for bgn, Zc, snr in [(1, 384, 7.0), (1, 384, 4.0), (2, 208, 5.0), (1, 12, 7.0)]:
    K, N, Nf, M = engine.dims(bgn, Zc)
    ck = engine.random_bits(B, K, seed=1, device="cuda")
    dn = engine.encode_batch(ck, bgn, Zc)
    llr = engine.awgn_llr(dn, snr, seed=2)
    L = 20
    for _ in range(2):
        r = engine.decode_bf_batch(llr, Zc, bgn, L)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        r = engine.decode_bf_batch(llr, Zc, bgn, L)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    rec = {"bgn": bgn, "Zc": Zc, "snr_db": snr, "L": L, "codeblocks": B, "qc_ms": ms,
           "qc_info_gbit_s": B * K / ms / 1e6, "mean_iters": float(r[2].float().mean()), "ok_frac": float(r[1].float().mean()),
           "llr_gb_s": B * N * 4 / ms / 1e6}
    nb = min(B, 512)
    h = llr[:nb].double().cpu().numpy()
    rp, ci = engine.csr(Zc, bgn)
    full = np.concatenate([np.zeros((nb, 2 * Zc)), h], axis=1)
    engine.decode_bf_csr_batch(full[:8], rp, ci, Nf, L)
    t0 = time.perf_counter(); g = engine.decode_bf_csr_batch(full, rp, ci, Nf, L); t1 = time.perf_counter()
    rec["generic_csr_host_ms_per_cb"] = (t1 - t0) * 1e3 / nb
    t0 = time.perf_counter(); q = engine.decode_bf_batch(h, Zc, bgn, L); t1 = time.perf_counter()
    rec["qc_host_ms_per_cb"] = (t1 - t0) * 1e3 / nb
    assert np.array_equal(g[0], q[0]) and np.array_equal(g[2], q[2]) and np.array_equal(q[0], r[0][:nb].cpu().numpy())
    out.append(rec)
    print(json.dumps(rec), flush=True)
