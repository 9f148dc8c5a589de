"""ctypes front-end of oracle/liboracle_nrldpc.so -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference` legs may import
this.  Each wrapper mirrors the reference call it restates (file:line under the reference root).
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "liboracle_nrldpc.so")
_lib = None

ALGO = {"min-sum": 0, "BP": 1}
POLY = {"6": 0, "11": 1, "16": 2, "24A": 3, "24B": 4, "24C": 5}


def build(force=False):
    """Compile the C restatement with the committed Makefile (gcc, seconds)."""
    srcs = [os.path.join(_HERE, f) for f in ("nrldpc_oracle.c", "nrldpc_oracle_soft.inc")]
    srcs.append(os.path.join(_HERE, "..", "include", "nrldpc_bg_tables.inc"))
    if force or not os.path.exists(_SO) or any(os.path.getmtime(s) > os.path.getmtime(_SO) for s in srcs):
        subprocess.run(["make", "-C", _HERE], check=True, capture_output=True)


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_SO)
    return _lib


def _p(a, t):
    return a.ctypes.data_as(ctypes.POINTER(t))


def dims(bgn, Zc):
    """(K, N, N', M) -- py5gphy/ldpc/nr_ldpc_decode.py:26-31"""
    if bgn == 1:
        return 22 * Zc, 66 * Zc, 68 * Zc, 46 * Zc
    return 10 * Zc, 50 * Zc, 52 * Zc, 42 * Zc


def find_iLS(Zc):
    """py5gphy/ldpc/ldpc_info.py:81-97"""
    return int(lib().oracle_find_ils(int(Zc)))


def getH(Zc, bgn):
    """py5gphy/ldpc/ldpc_info.py:99-139 (dense int8)"""
    K, N, Nv, M = dims(bgn, Zc)
    H = np.zeros((M, Nv), np.int8)
    assert lib().oracle_get_h_dense(bgn, Zc, _p(H, ctypes.c_int8)) == 0
    return H


def csr(Zc, bgn):
    K, N, Nv, M = dims(bgn, Zc)
    nnz = (316 if bgn == 1 else 197) * Zc
    rowptr = np.zeros(M + 1, np.int32)
    colidx = np.zeros(nnz, np.int32)
    assert lib().oracle_build_csr(bgn, Zc, _p(rowptr, ctypes.c_int), _p(colidx, ctypes.c_int)) == nnz
    return rowptr, colidx


def dense_to_csr(H):
    H = np.asarray(H)
    rows, cols = np.nonzero(H)
    rowptr = np.zeros(H.shape[0] + 1, np.int32)
    np.add.at(rowptr, rows + 1, 1)
    return np.cumsum(rowptr).astype(np.int32), cols.astype(np.int32)


def encode_ldpc(ck, bgn):
    """py5gphy/ldpc/nr_ldpc_encode.py:8-50 -- mutates ck (fillers -> 0) like the reference."""
    assert ck.dtype == np.int8 and ck.flags.c_contiguous
    K = ck.size
    Zc = K // 22 if bgn == 1 else K // 10
    _, N, _, _ = dims(bgn, Zc)
    dn = np.empty(N, np.int8)
    assert lib().oracle_encode(_p(ck, ctypes.c_int8), bgn, Zc, _p(dn, ctypes.c_int8)) == 0
    return dn


def encode_batch(ck, bgn, Zc, nthreads=0):
    ck = np.ascontiguousarray(ck, np.int8)
    B = ck.shape[0]
    _, N, _, _ = dims(bgn, Zc)
    dn = np.empty((B, N), np.int8)
    assert lib().oracle_encode_batch(_p(ck, ctypes.c_int8), B, bgn, Zc, _p(dn, ctypes.c_int8), nthreads) == 0
    return dn


def decode_batch(llr, Zc, bgn, L, algo="min-sum", alpha=1.0, beta=0.0, early_term=1, dtype=np.float64, nthreads=0):
    """B x nr_decode_ldpc (py5gphy/ldpc/nr_ldpc_decode.py:11-49) -> ck int8[B,N'], status bool[B], iters int32[B]."""
    llr = np.ascontiguousarray(np.atleast_2d(llr), dtype)
    B = llr.shape[0]
    K, N, Nv, M = dims(bgn, Zc)
    assert llr.shape[1] == N
    ck = np.empty((B, Nv), np.int8)
    st = np.zeros(B, np.int32)
    it = np.zeros(B, np.int32)
    if dtype == np.float64:
        fn, ct = lib().oracle_decode_soft_batch_f64, ctypes.c_double
    else:
        fn, ct = lib().oracle_decode_soft_batch_f32, ctypes.c_float
    rc = fn(_p(llr, ct), B, bgn, Zc, int(L), ALGO[algo], ct(alpha), ct(beta), int(early_term),
            _p(ck, ctypes.c_int8), _p(st, ctypes.c_int), _p(it, ctypes.c_int), nthreads)
    assert rc == 0
    return ck, st.astype(bool), it


def nr_decode_ldpc(LLRin, Zc, bgn, L, algo="min-sum", alpha=1, beta=0, dtype=np.float64, early_term=1):
    """py5gphy/ldpc/nr_ldpc_decode.py:11-49 -> (blkandcrc, ck, status, iters)"""
    K = dims(bgn, Zc)[0]
    if algo == "BF":
        llr = np.ascontiguousarray(LLRin, np.float64)
        ck = np.empty(dims(bgn, Zc)[2], np.int8)
        st, it = ctypes.c_int(0), ctypes.c_int(0)
        assert lib().oracle_decode_bf(_p(llr, ctypes.c_double), bgn, Zc, int(L), _p(ck, ctypes.c_int8),
                                      ctypes.byref(st), ctypes.byref(it)) == 0
        return ck[:K], ck, bool(st.value), it.value
    ck, st, it = decode_batch(np.asarray(LLRin)[None, :], Zc, bgn, L, algo, alpha, beta, early_term, dtype)
    return ck[0, :K], ck[0], bool(st[0]), int(it[0])


def decode_ldpc(LLRin, H, L, algo="min-sum", alpha=1, beta=0, dtype=np.float64, early_term=1):
    """py5gphy/ldpc/nr_ldpc_decode.py:51-143 on an arbitrary dense H -> (ck, status, iters)"""
    rowptr, colidx = dense_to_csr(H)
    M, Nv = np.asarray(H).shape
    ck = np.empty(Nv, np.int8)
    st, it = ctypes.c_int(0), ctypes.c_int(0)
    if algo == "BF":
        llr = np.ascontiguousarray(LLRin, np.float64)
        lib().oracle_bf_csr(_p(llr, ctypes.c_double), M, Nv, _p(rowptr, ctypes.c_int), _p(colidx, ctypes.c_int),
                            int(L), _p(ck, ctypes.c_int8), ctypes.byref(st), ctypes.byref(it))
    else:
        llr = np.ascontiguousarray(LLRin, dtype)
        if dtype == np.float64:
            fn, ct = lib().oracle_soft_csr_f64, ctypes.c_double
        else:
            fn, ct = lib().oracle_soft_csr_f32, ctypes.c_float
        fn(_p(llr, ct), M, Nv, _p(rowptr, ctypes.c_int), _p(colidx, ctypes.c_int), int(L), ALGO[algo],
           ct(alpha), ct(beta), int(early_term), _p(ck, ctypes.c_int8), ctypes.byref(st), ctypes.byref(it))
    return ck, bool(st.value), it.value


def nr_crc_encode(blk, poly):
    """py5gphy/crc/crc.py:4-41 (mask=0)"""
    blk = np.ascontiguousarray(blk, np.int8)
    out = np.empty(blk.size + 24, np.int8)
    L = lib().oracle_crc_encode(_p(blk, ctypes.c_int8), blk.size, POLY[poly.upper()], _p(out, ctypes.c_int8))
    assert L > 0
    return out[: blk.size + L].copy()


def crc_decode(blkandcrc, poly):
    """(blk, err) of py5gphy/crc/crc.py:43-88 (mask=0): the long division of the whole block leaves a zero remainder
    exactly when the CRC of the first A bits equals the last L bits."""
    x = np.ascontiguousarray(blkandcrc, np.int8)
    L = {"6": 6, "11": 11, "16": 16, "24A": 24, "24B": 24, "24C": 24}[poly.upper()]
    A = x.size - L
    return x[:A].copy(), int(not np.array_equal(nr_crc_encode(x[:A], poly), x))


def ratematch_ldpc(dn, Ncb, E, k0, Qm):
    """py5gphy/ldpc/nr_ldpc_ratematch.py:64-97"""
    dn = np.ascontiguousarray(dn, np.int8)
    fe = np.empty(E, np.int8)
    assert lib().oracle_ratematch(_p(dn, ctypes.c_int8), dn.size, int(Ncb), int(E), int(k0), int(Qm), _p(fe, ctypes.c_int8)) == 0
    return fe


def raterecover_ldpc(LLr_fe, Ncb, N, k0, Qm, Zc, K_apo, K):
    """py5gphy/ldpc/nr_ldpc_raterecover.py:6-65 (float64)"""
    fe = np.ascontiguousarray(LLr_fe, np.float64)
    out = np.empty(N, np.float64)
    assert lib().oracle_raterecover(_p(fe, ctypes.c_double), fe.size, int(Ncb), int(N), int(k0), int(Qm), int(Zc),
                                    int(K_apo), int(K), _p(out, ctypes.c_double)) == 0
    return out


def harq_combine(new, cur):
    """py5gphy/nr_pdsch/nr_dlsch_decode.py:80-87"""
    a = np.ascontiguousarray(new, np.float64)
    c = np.ascontiguousarray(cur, np.float64)
    out = np.empty_like(a)
    lib().oracle_harq_combine(_p(a, ctypes.c_double), _p(c, ctypes.c_double), ctypes.c_long(a.size), _p(out, ctypes.c_double))
    return out


def num_threads():
    return int(lib().oracle_num_threads())


def set_num_threads(n):
    """Override OMP_NUM_THREADS (torchrun sets it to 1) for the CPU-baseline legs of bench.py."""
    lib().oracle_set_num_threads(int(n))
