"""CPU: the C-ABI library loads and exports every symbol include/nrldpc_b200.h declares (no compute
calls), and the host-side mirrors of py5gphy/ldpc/ldpc_info.py behave like the reference."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def so():
    from python_5gtoolbox_b200 import _lib
    if not os.path.exists(_lib.SO_PATH):
        _lib.build()
    return _lib


def test_header_symbols_exported(so):
    hdr = open(os.path.join(ROOT, "include", "nrldpc_b200.h")).read()
    declared = sorted(set(re.findall(r"\b(nrldpc_[a-z0-9_]+)\s*\(", hdr)))
    assert len(declared) >= 20
    L = ctypes.CDLL(so.SO_PATH)
    for name in declared:
        assert hasattr(L, name), f"{name} declared in include/nrldpc_b200.h but not exported"
    assert set(so.exported_symbols()) == set(declared)  # the ctypes binding covers the whole header


def test_no_device_paths_fail_loudly(so):
    L = so.lib()
    assert L.nrldpc_version() >= 100
    assert L.nrldpc_find_ils(384) == 1 and L.nrldpc_find_ils(17) == 255
    K, N, Nf, M = (ctypes.c_int() for _ in range(4))
    assert L.nrldpc_dims(1, 384, K, N, Nf, M) == 0
    assert (K.value, N.value, Nf.value, M.value) == (8448, 25344, 26112, 17664)
    assert L.nrldpc_dims(3, 384, K, N, Nf, M) == -1 and b"bgn" in L.nrldpc_last_error()
    if L.nrldpc_device_count() == 0:
        # no CPU fallback: a compute call without a GPU must return an error, not a result
        ck = np.zeros((1, 44), np.int8)
        dn = np.zeros((1, 132), np.int8)
        assert L.nrldpc_encode_host(ck.ctypes.data, 1, 1, 2, 1, dn.ctypes.data) < 0
        from python_5gtoolbox_b200 import engine, NrLdpcError
        with pytest.raises(NrLdpcError):
            engine.decode_batch(np.zeros((1, 132), np.float32), 2, 1, 4)
        with pytest.raises(NrLdpcError):
            engine.decode_bf_batch(np.ones((1, 132)), 2, 1, 4)   # the bit-flipping kernel too
        assert engine.bind_host_to_device(0) is None            # no device: affinity left alone
    # argument checks of the mixed-(bgn, Zc) entry points come before any CUDA call
    one = (ctypes.c_int * 1)(1)
    bad_zc = (ctypes.c_int * 1)(17)
    ptrs = (ctypes.c_void_p * 1)(None)
    assert L.nrldpc_decode_minsum_groups(1, ptrs, one, one, bad_zc, 4, 1.0, 0.0, 1, None, None, None, None, None) == -1
    assert L.nrldpc_decode_minsum_groups(0, None, None, None, None, 4, 1.0, 0.0, 1, None, None, None, None, None) == 0
    assert L.nrldpc_encode_groups(-1, None, None, None, None, 1, None, None) == -1


def test_csr_matches_oracle(so, oracle):
    from python_5gtoolbox_b200 import engine
    for bgn, Zc in [(1, 2), (2, 3), (1, 12), (2, 52), (1, 384)]:
        rp, ci = engine.csr(Zc, bgn)
        rp2, ci2 = oracle.csr(Zc, bgn)
        assert np.array_equal(rp, rp2) and np.array_equal(ci, ci2)


def test_ldpc_info_mirror(oracle):
    from python_5gtoolbox_b200.ldpc import ldpc_info
    from tests.conftest import ZLIST
    assert ldpc_info._LIFT_SIZES == ZLIST
    for Zc in range(0, 400):
        assert ldpc_info.find_iLS(Zc) == oracle.find_iLS(Zc)
    for bgn, Zc in [(1, 2), (2, 7), (1, 20), (2, 44)]:
        H = ldpc_info.getH(Zc, bgn, ldpc_info.find_iLS(Zc))
        assert H.dtype == np.int8 and np.array_equal(np.asarray(H), oracle.getH(Zc, bgn))
        assert H.nrldpc_tag == (bgn, Zc) and H[:, :].nrldpc_tag is None
    # TS 38.212 5.2.2 examples (values cross-checked against the live reference when this was written)
    assert ldpc_info.get_cbs_info(98400, 1) == (12, 8200, 24, 224, 8448, 384)
    assert ldpc_info.get_cbs_info(16896, 1) == (3, 5632, 24, 680, 6336, 288)
    assert ldpc_info.get_cbs_info(100, 2) == (1, 100, 0, 80, 180, 18)
    assert ldpc_info.get_cbs_info(600, 2) == (1, 600, 0, 120, 720, 72)
    with pytest.raises(AssertionError):
        ldpc_info.get_cbs_info(3841, 2)
    H, K, Zc = ldpc_info.gen_ldpc_para(66 * 4, 1)
    assert (K, Zc) == (88, 4) and H.shape == (184, 272)


def test_table_checksums():
    """SURVEY Appendix B crc32 of every base-graph table."""
    import zlib
    from python_5gtoolbox_b200.ldpc import ldpc_info
    want = {(1, 0): "b68fc6c3", (1, 1): "f5948ae1", (1, 6): "6f5f6990", (2, 0): "094b454c", (2, 2): "bf5a7060", (2, 7): "ff802df5"}
    for (bgn, s), crc in want.items():
        assert "%08x" % zlib.crc32(ldpc_info.base_graph(bgn, s).astype("<i2").tobytes()) == crc


def test_bench_reference_arm_contract():
    """CPU: `bench.py --impl reference` (the oracle port on the host cores) prints ONE JSON line with the contract's keys,
    honours --steps inside its time budget and needs no GPU."""
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, REF_BUDGET_S="60", CUDA_VISIBLE_DEVICES="")
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "1"],
                       capture_output=True, text=True, timeout=600, env=env, cwd=root)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "decoded_info_gbit_per_s" and d["unit"] == "Gbit/s" and d["higher_is_better"]
    assert d["steps"] == 2 and d["value"] > 0 and d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"] == {"value": d["value"], "unit": "Gbit/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "BG1 Zc=384" in d["config"]["workload"]
