import sys, torch
sys.path.insert(0, ".")
from python_5gtoolbox_b200 import engine
bgn, Zc, B = 1, int(sys.argv[1]) if len(sys.argv) > 1 else 208, int(sys.argv[2]) if len(sys.argv) > 2 else 592
K, N, Nf, M = engine.dims(bgn, Zc)
ck = engine.random_bits(B, K, seed=1, device="cuda"); dn = engine.encode_batch(ck, bgn); llr = engine.awgn_llr(dn, 1.0, seed=2)
for i in range(3):
    r = engine.decode_batch(llr, Zc, bgn, 10, 0.8, 0.0, False, want_ck=False, want_info=True); torch.cuda.synchronize()
