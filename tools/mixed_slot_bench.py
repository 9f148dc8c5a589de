#!/usr/bin/env python3
"""BASELINE.json config #4, mixed-Zc part: the transport blocks of one slot (one (bgn, Zc) per transport block) decoded
as one launch per transport block back to back on one stream vs nrldpc_decode_minsum_groups (concurrent side streams).
python tools/mixed_slot_bench.py"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from python_5gtoolbox_b200 import engine  # noqa: E402

slots = {
    "16 UEs, 4-12 codeblocks each": [(1, 384, 12), (1, 352, 8), (2, 384, 6), (1, 320, 10), (2, 352, 4), (1, 288, 9), (2, 288, 5),
                                     (1, 208, 7), (2, 208, 4), (1, 176, 6), (2, 176, 4), (1, 384, 8), (1, 256, 6), (2, 320, 5),
                                     (1, 144, 4), (2, 96, 4)],
    "1 large + 3 small transport blocks": [(1, 384, 115), (2, 352, 3), (1, 208, 2), (2, 176, 1)],
    "64 single-codeblock transport blocks": [(1 + (i % 2), [384, 352, 320, 288, 208, 176, 72, 28][i % 8], 1) for i in range(64)],
}
for name, tbs in slots.items():
    groups = []
    for n, (bgn, Zc, C) in enumerate(tbs):
        K, N, Nf, M = engine.dims(bgn, Zc)
        ck = engine.random_bits(C, K, seed=10 + n, device="cuda")
        groups.append((engine.awgn_llr(engine.encode_batch(ck, bgn, Zc), 2.0, seed=50 + n), Zc, bgn))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def seq():
        return [engine.decode_batch(l, Zc, bgn, 10, 0.8, 0.0, True, want_ck=False, want_info=True) for l, Zc, bgn in groups]

    def par():
        return engine.decode_groups(groups, 10, 0.8, 0.0, True, want_ck=False, want_info=True)

    res = {}
    for label, fn in (("one_stream_ms", seq), ("groups_ms", par)):
        for _ in range(3):
            out = fn()
        torch.cuda.synchronize()
        e0.record()
        for _ in range(10):
            out = fn()
        e1.record()
        torch.cuda.synchronize()
        res[label] = e0.elapsed_time(e1) / 10
        res[label + "_checksum"] = int(sum(int(o["info"].long().sum()) for o in out))
    assert res["one_stream_ms_checksum"] == res["groups_ms_checksum"]
    bits = sum(engine.dims(b, z)[0] * c for b, z, c in tbs)
    print(json.dumps({"slot": name, "transport_blocks": len(tbs), "codeblocks": sum(c for _, _, c in tbs), "info_bits": bits,
                      "one_stream_ms": round(res["one_stream_ms"], 4), "groups_ms": round(res["groups_ms"], 4),
                      "speedup": round(res["one_stream_ms"] / res["groups_ms"], 2),
                      "groups_gbit_s": round(bits / res["groups_ms"] / 1e6, 2)}), flush=True)
