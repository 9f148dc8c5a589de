import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")
ZLIST = [2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16, 18, 20, 22, 24, 26, 28, 30, 32, 36, 40, 44, 48, 52,
         56, 60, 64, 72, 80, 88, 96, 104, 112, 120, 128, 144, 160, 176, 192, 208, 224, 240, 256, 288, 320, 352, 384]


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as O
    O.build()
    return O


@pytest.fixture(scope="session")
def enc_golden():
    with np.load(os.path.join(GOLDEN, "encode_golden.npz")) as z:
        n = sum(1 for k in z.files if k.startswith("meta_"))
        return [dict(bgn=int(z[f"meta_{i}"][0]), Zc=int(z[f"meta_{i}"][1]), F=int(z[f"meta_{i}"][2]),
                     ck=z[f"ck_{i}"], ck_after=z[f"ckafter_{i}"], dn=z[f"dn_{i}"]) for i in range(n)]


@pytest.fixture(scope="session")
def dec_golden():
    out = []
    with np.load(os.path.join(GOLDEN, "decode_golden.npz")) as z:
        n = sum(1 for k in z.files if k.startswith("cfg_"))
        for i in range(n):
            bgn, Zc, snr, L, algo, alpha, beta, seed = z[f"cfg_{i}"]
            bgn, Zc, L = int(bgn), int(Zc), int(L)
            Nf = (68 if bgn == 1 else 52) * Zc
            K = (22 if bgn == 1 else 10) * Zc
            out.append(dict(bgn=bgn, Zc=Zc, snr=float(snr), L=L, algo=["min-sum", "BP", "BF"][int(algo)],
                            alpha=float(alpha), beta=float(beta), seed=int(seed), llr=z[f"llr_{i}"],
                            blk=np.unpackbits(z[f"blk_{i}"])[:K].astype("i1"),
                            ck=np.unpackbits(z[f"ck_{i}"])[:Nf].astype("i1"),
                            status=bool(z[f"res_{i}"][0]), iters=int(z[f"res_{i}"][1])))
    return out


def hex_to_bits(h, n):
    """SURVEY Appendix C convention: hex, MSB first, first bit = index 0, right-padded to a nibble."""
    bits = np.array([(int(c, 16) >> (3 - k)) & 1 for c in h for k in range(4)], "i1")
    return bits[:n]
