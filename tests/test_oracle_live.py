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


_CHILD_RM = r"""
import sys, os, json
import numpy as np
sys.path.insert(0, sys.argv[1]); os.chdir(sys.argv[1]); sys.path.insert(0, sys.argv[2])
from py5gphy.ldpc import nr_ldpc_ratematch as RM, nr_ldpc_raterecover as RR
from py5gphy.nr_pdsch import nr_dlsch
from oracle import oracle as O
import python_5gtoolbox_b200 as P
from python_5gtoolbox_b200.ldpc import nr_ldpc_ratematch as MRM
rng = np.random.default_rng(11); bad = 0; n = 0
for trial in range(40):
    bgn = int(rng.integers(1, 3)); Zc = int(rng.choice([2, 4, 7, 11, 24]))
    K = (22 if bgn == 1 else 10) * Zc; N = (66 if bgn == 1 else 50) * Zc
    F = int(rng.integers(0, Zc)); dn = rng.integers(0, 2, N).astype("i1")
    if F: dn[K - F - 2 * Zc:K - 2 * Zc] = -1
    Ncb = N if trial % 2 else int(rng.integers(max(K, N // 2), N + 1))
    rv = int(rng.integers(0, 4)); k0 = RM.get_k0(Ncb, bgn, rv, Zc)
    Qm = int(rng.choice([1, 2, 4, 6, 8])); E = Qm * int(rng.integers(max(1, N // (3 * Qm)), (2 * N) // Qm + 1))
    bad += not np.array_equal(RM.ratematch_ldpc(dn, Ncb, E, k0, Qm), O.ratematch_ldpc(dn, Ncb, E, k0, Qm))
    llr = rng.normal(0, 4, E).astype("f4").astype("f8")
    bad += not np.array_equal(RR.raterecover_ldpc(llr, Ncb, N, k0, Qm, Zc, K - F, K), O.raterecover_ldpc(llr, Ncb, N, k0, Qm, Zc, K - F, K))
    bad += MRM.get_k0(Ncb, bgn, rv, Zc) != k0
    C = int(rng.integers(1, 60)); NL = int(rng.integers(1, 5)); G = Qm * NL * int(rng.integers(C, 9000))
    bad += MRM.get_Er_ldpc(G, C, Qm, NL) != RM.get_Er_ldpc(G, C, Qm, NL)
    n += 1
# install() rebinds the reference's own module attributes (and uninstall() restores them)
orig = nr_dlsch.DLSCHEncode
names = P.install()
rebound = nr_dlsch.DLSCHEncode is not orig and RM.ratematch_ldpc.__module__.startswith("python_5gtoolbox_b200")
P.uninstall()
restored = nr_dlsch.DLSCHEncode is orig and RM.ratematch_ldpc.__module__.startswith("py5gphy")
print(json.dumps({"n": n, "bad": int(bad), "rebound": bool(rebound), "restored": bool(restored), "names": len(names)}))
"""


def test_ratematch_oracle_and_overlay_against_live_reference():
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-c", _CHILD_RM, REF, root], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    import json
    res = json.loads(r.stdout.strip().splitlines()[-1])
    assert res["n"] == 40 and res["bad"] == 0 and res["rebound"] and res["restored"] and res["names"] >= 19
