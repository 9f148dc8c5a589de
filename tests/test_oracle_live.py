"""CPU, build container only: the oracle against the LIVE reference at /root/reference (skipped on the
GPU box, where that mount does not exist).  tests/golden/*.npz are frozen samples of the same check."""
import os
import subprocess
import sys

import numpy as np
import pytest

REF = "/root/reference"
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "py5gphy")), reason="reference mount absent")

_CHILD = r"""
import sys, os, json
import numpy as np
sys.path.insert(0, sys.argv[1]); os.chdir(sys.argv[1]); sys.path.insert(0, sys.argv[2])
from py5gphy.ldpc import nr_ldpc_decode, nr_ldpc_encode
from oracle import oracle as O
cnt = [0]
orig = nr_ldpc_decode._min_sum_process
def wrap(*a, **k):
    cnt[0] += 1
    return orig(*a, **k)
nr_ldpc_decode._min_sum_process = wrap
np.random.seed(2024)
bad = 0; n = 0
for bgn, Zc, snr, L, al, be in [(1, 3, 0.5, 10, 0.8, 0), (2, 4, -2.0, 12, 1, 0.5), (1, 6, 0.0, 16, 0.8, 0.3),
                                (2, 9, -2.5, 8, 0.7, 0), (1, 10, 0.0, 10, 1, 0), (2, 14, -2.2, 10, 0.9, 0.1)]:
    for t in range(4):
        K = (22 if bgn == 1 else 10) * Zc
        blk, dn, llr = nr_ldpc_decode.for_test_5g_ldpc_encoder(Zc, bgn, snr, "24A" if K > 32 else "16")
        ck2 = blk.copy(); assert np.array_equal(O.encode_ldpc(ck2, bgn), dn)
        llr = llr.astype("f4").astype("f8")
        cnt[0] = 0
        b1, c1, s1 = nr_ldpc_decode.nr_decode_ldpc(llr, Zc, bgn, L, "min-sum", al, be)
        it1 = cnt[0] // ((46 if bgn == 1 else 42) * Zc)
        b2, c2, s2, it2 = O.nr_decode_ldpc(llr, Zc, bgn, L, "min-sum", al, be)
        n += 1
        bad += not (np.array_equal(c1, c2) and s1 == s2 and it1 == it2)
        # bit flipping on the same codeword at a high SNR
        b1, c1, s1 = nr_ldpc_decode.nr_decode_ldpc(llr * 3, Zc, bgn, 8, "BF")
        b2, c2, s2, _ = O.nr_decode_ldpc(llr * 3, Zc, bgn, 8, "BF")
        bad += not (np.array_equal(c1, c2) and s1 == s2)
print(json.dumps({"n": n, "bad": int(bad)}))
"""


def test_oracle_matches_live_reference():
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-c", _CHILD, REF, root], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    import json
    res = json.loads(r.stdout.strip().splitlines()[-1])
    assert res["n"] == 24 and res["bad"] == 0
