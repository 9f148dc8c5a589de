import sys, ctypes, torch
sys.path.insert(0, ".")
from python_5gtoolbox_b200 import engine, _lib
L_ = _lib.lib(); dev = torch.device("cuda"); Zc, bgn, mm = 384, 1, 16384
K = 22 * Zc; A = K - 24
def ev(): return torch.cuda.Event(enable_timing=True)
for rep in range(5):
    t = [ev() for _ in range(8)]
    s = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    t[0].record()
    bits = torch.empty((mm, A), dtype=torch.int8, device=dev)
    _lib.check(L_.nrldpc_random_bits_rows(bits.data_ptr(), mm, A, 1, 0, 1, s), "rb"); t[1].record()
    blk = torch.empty((mm, K), dtype=torch.int8, device=dev)
    _lib.check(L_.nrldpc_crc_encode(bits.data_ptr(), mm, A, 3, blk.data_ptr(), s), "crc"); t[2].record()
    dn = engine.encode_batch(blk.clone(), bgn, Zc); t[3].record()
    llr = torch.empty(dn.shape, dtype=torch.float32, device=dev)
    _lib.check(L_.nrldpc_awgn_llr_rows(dn.data_ptr(), mm, dn.shape[1], 1.0, 1, 0, 1, llr.data_ptr(), s), "awgn"); t[4].record()
    r = engine.decode_batch(llr, Zc, bgn, 10, 0.8, 0.0, True); t[5].record()
    c = engine.count_errors(blk, r["ck"], K, r["iters"]); t[6].record()
    torch.cuda.synchronize()
    names = ["random_bits", "crc", "clone+encode", "awgn", "decode(ET,+ck out)", "count_errors"]
    print(rep, {n: round(t[i].elapsed_time(t[i + 1]), 3) for i, n in enumerate(names)})
# the bit-packed chain (same codeblocks, same noise)
for rep in range(5):
    t = [ev() for _ in range(8)]
    t[0].record()
    blk = engine.random_bits_packed(mm, A, 1, dev, first_id=0, row_words=K // 32); t[1].record()
    engine.crc_attach_packed(blk, A, "24A"); t[2].record()
    dn = engine.encode_packed(blk, bgn, Zc); t[3].record()
    llr = engine.awgn_llr_packed(dn, dn.shape[1] * 32, 1.0, 1, first_id=0); t[4].record()
    r = engine.decode_batch(llr, Zc, bgn, 10, 0.8, 0.0, True, want_ck=False, want_info=True); t[5].record()
    c2 = engine.count_errors_packed(blk, r["info"], K, r["iters"]); t[6].record()
    torch.cuda.synchronize()
    names = ["random_bits_packed", "crc_attach_packed", "encode_packed", "awgn_packed", "decode(ET, info out)", "count_errors_packed"]
    print(rep, {n: round(t[i].elapsed_time(t[i + 1]), 3) for i, n in enumerate(names)}, "same counters:", c.tolist() == c2.tolist())
