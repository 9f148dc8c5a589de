import sys
sys.path.insert(0, ".")
import torch
from python_5gtoolbox_b200 import engine
B, bgn, Zc = 4096, 1, 384
K, N, Nf, M = engine.dims(bgn, Zc)
ck = engine.random_bits(B, K, seed=1, device="cuda")
dn = engine.encode_batch(ck, bgn, Zc)
llr = engine.awgn_llr(dn, 7.0, seed=2)
for i in range(3):
    r = engine.decode_bf_batch(llr, Zc, bgn, 20)
torch.cuda.synchronize()
