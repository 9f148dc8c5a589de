"""Batched LDPC entry points over the C ABI (include/nrldpc_b200.h).

NumPy arrays go through the ``*_host`` entry points (H2D / kernels / D2H inside, synchronous);
CUDA ``torch.Tensor`` arguments are used in place through their ``data_ptr()`` on torch's current
stream (PyTorch is only the allocator / stream provider).  One (bgn, Zc) per call, like one
transport block in the reference (py5gphy/ldpc/ldpc_info.py:62-69).
"""
import ctypes
import os

import numpy as np

from . import _lib

ALGO_MINSUM, ALGO_BP = 0, 1


def dims(bgn, Zc):
    """(K, N, N', M) of py5gphy/ldpc/nr_ldpc_decode.py:26-31."""
    assert bgn in [1, 2]
    assert _lib.lib().nrldpc_find_ils(int(Zc)) < 8
    K, N, Nf, M = (ctypes.c_int() for _ in range(4))
    _lib.check(_lib.lib().nrldpc_dims(bgn, int(Zc), K, N, Nf, M), "dims")
    return K.value, N.value, Nf.value, M.value


def find_iLS(Zc):
    return _lib.lib().nrldpc_find_ils(int(Zc))


def csr(Zc, bgn):
    """Sparse getH (py5gphy/ldpc/ldpc_info.py:99-139): (rowptr int32[M+1], colidx int32[E])."""
    K, N, Nf, M = dims(bgn, Zc)
    E = (316 if bgn == 1 else 197) * Zc
    rowptr = np.empty(M + 1, np.int32)
    colidx = np.empty(E, np.int32)
    n = _lib.check(_lib.lib().nrldpc_build_csr(bgn, int(Zc), rowptr.ctypes.data, colidx.ctypes.data), "build_csr")
    assert n == E
    return rowptr, colidx


def _is_torch(x):
    return type(x).__module__.startswith("torch")


def _stream_ptr():
    import torch
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


# ------------------------------------------------------------------ encoder

def encode_batch(ck, bgn, Zc=None, fix_fillers=True):
    """nr_ldpc_encode.encode_ldpc for B codeblocks (py5gphy/ldpc/nr_ldpc_encode.py:8-50).

    ck: int8 [B,K] with -1 fillers (NumPy array or CUDA tensor); mutated in place (fillers -> 0) when
    fix_fillers, which is the reference's side effect.  Returns dn int8 [B,N] of the same kind.
    """
    assert bgn in [1, 2]
    assert ck.ndim == 2
    B, K = ck.shape
    if Zc is None:
        Zc = K // 22 if bgn == 1 else K // 10
    Kx, N, Nf, M = dims(bgn, Zc)
    assert K == Kx
    L = _lib.lib()
    if _is_torch(ck):
        import torch
        assert ck.is_cuda and ck.dtype == torch.int8 and ck.is_contiguous()
        dn = torch.empty((B, N), dtype=torch.int8, device=ck.device)
        with torch.cuda.device(ck.device):
            _lib.check(L.nrldpc_encode(ck.data_ptr(), B, bgn, Zc, int(fix_fillers), dn.data_ptr(), _stream_ptr()), "encode")
        return dn
    assert ck.dtype == np.int8 and ck.flags.c_contiguous
    dn = np.empty((B, N), np.int8)
    _lib.check(L.nrldpc_encode_host(ck.ctypes.data, B, bgn, Zc, int(fix_fillers), dn.ctypes.data), "encode")
    return dn


# ------------------------------------------------------------------ min-sum decoder

def decode_geometry(bgn, Zc):
    """(codeblocks per CTA, threads, dynamic shared memory bytes) of the kernel decode_batch launches for (bgn, Zc)."""
    G, nt, smem = (ctypes.c_int() for _ in range(3))
    _lib.check(_lib.lib().nrldpc_decode_minsum_geometry(int(bgn), int(Zc), G, nt, smem), "decode_geometry")
    return G.value, nt.value, smem.value


def decode_batch(llr, Zc, bgn, L, alpha=1.0, beta=0.0, early_term=True, want_ck=True, want_info=False):
    """nr_decode_ldpc(..., 'min-sum', alpha, beta) for B codeblocks in fp32
    (py5gphy/ldpc/nr_ldpc_decode.py:11-49,51-143,178-227).

    llr: float32 [B,N], NumPy (host path) or CUDA tensor (device path, async on the current stream).
    Returns dict(ck=int8[B,N'] | None, info=uint32[B,ceil(K/32)] | None, status=bool/uint8[B], iters=int32[B]).
    """
    assert bgn in [1, 2]
    K, N, Nf, M = dims(bgn, Zc)
    assert llr.ndim == 2 and llr.shape[1] == N
    B = llr.shape[0]
    nw = (K + 31) // 32
    Lb = _lib.lib()
    if _is_torch(llr):
        import torch
        assert llr.is_cuda and llr.dtype == torch.float32 and llr.is_contiguous()
        dev = llr.device
        ck = torch.empty((B, Nf), dtype=torch.int8, device=dev) if want_ck else None
        info = torch.empty((B, nw), dtype=torch.int32, device=dev) if want_info else None
        status = torch.empty((B,), dtype=torch.uint8, device=dev)
        iters = torch.empty((B,), dtype=torch.int32, device=dev)
        with torch.cuda.device(dev):
            _lib.check(Lb.nrldpc_decode_minsum(llr.data_ptr(), B, bgn, int(Zc), int(L), float(alpha), float(beta),
                                               int(bool(early_term)), ck.data_ptr() if want_ck else None,
                                               info.data_ptr() if want_info else None, status.data_ptr(),
                                               iters.data_ptr(), _stream_ptr()), "decode_minsum")
        return dict(ck=ck, info=info, status=status, iters=iters)
    f16 = llr.dtype == np.float16   # half-precision LLRs cross the host link as they are and are widened on the device
    llr = np.ascontiguousarray(llr, np.float16 if f16 else np.float32)
    ck = np.empty((B, Nf), np.int8) if want_ck else None
    info = np.empty((B, nw), np.uint32) if want_info else None
    status = np.empty(B, np.uint8)
    iters = np.empty(B, np.int32)
    _lib.check((Lb.nrldpc_decode_minsum_host_f16 if f16 else Lb.nrldpc_decode_minsum_host)(llr.ctypes.data, B, bgn, int(Zc), int(L), float(alpha), float(beta),
                                            int(bool(early_term)), ck.ctypes.data if want_ck else None,
                                            info.ctypes.data if want_info else None, status.ctypes.data,
                                            iters.ctypes.data), "decode_minsum")
    return dict(ck=ck, info=info, status=status.astype(bool), iters=iters)


def _ptr_array(tensors):
    return (ctypes.c_void_p * len(tensors))(*[t.data_ptr() if t is not None else None for t in tensors])


def decode_groups(groups, L, alpha=1.0, beta=0.0, early_term=True, want_ck=True, want_info=False):
    """Mixed-(bgn, Zc) batch on the device: groups = [(llr float32 CUDA [B_g, N_g], Zc_g, bgn_g), ...], e.g. the
    transport blocks of one slot (one Zc per transport block, py5gphy/ldpc/ldpc_info.py:62-69).  One launch per group,
    concurrent on the library's side streams, ordered on the current stream.  Returns a list of decode_batch dicts."""
    import torch
    n = len(groups)
    if n == 0:
        return []
    dev = groups[0][0].device
    outs, Bs, bgs, zs = [], [], [], []
    for llr, Zc, bgn in groups:
        assert bgn in [1, 2]
        K, N, Nf, M = dims(bgn, Zc)
        assert llr.is_cuda and llr.device == dev and llr.dtype == torch.float32 and llr.is_contiguous()
        assert llr.ndim == 2 and llr.shape[1] == N
        B = llr.shape[0]
        outs.append(dict(ck=torch.empty((B, Nf), dtype=torch.int8, device=dev) if want_ck else None,
                         info=torch.empty((B, (K + 31) // 32), dtype=torch.int32, device=dev) if want_info else None,
                         status=torch.empty((B,), dtype=torch.uint8, device=dev),
                         iters=torch.empty((B,), dtype=torch.int32, device=dev)))
        Bs.append(B), bgs.append(bgn), zs.append(int(Zc))
    ia = lambda v: (ctypes.c_int * n)(*v)
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().nrldpc_decode_minsum_groups(
            n, _ptr_array([g[0] for g in groups]), ia(Bs), ia(bgs), ia(zs), int(L), float(alpha), float(beta),
            int(bool(early_term)), _ptr_array([o["ck"] for o in outs]) if want_ck else None,
            _ptr_array([o["info"] for o in outs]) if want_info else None, _ptr_array([o["status"] for o in outs]),
            _ptr_array([o["iters"] for o in outs]), _stream_ptr()), "decode_minsum_groups")
    return outs


def encode_groups(groups, fix_fillers=True):
    """Mixed-(bgn, Zc) encode on the device: groups = [(ck int8 CUDA [B_g, K_g], Zc_g, bgn_g), ...] -> [dn int8 [B_g, N_g]]."""
    import torch
    n = len(groups)
    if n == 0:
        return []
    dev = groups[0][0].device
    dns, Bs, bgs, zs = [], [], [], []
    for ck, Zc, bgn in groups:
        K, N, Nf, M = dims(bgn, Zc)
        assert ck.is_cuda and ck.device == dev and ck.dtype == torch.int8 and ck.is_contiguous() and ck.shape[1] == K
        dns.append(torch.empty((ck.shape[0], N), dtype=torch.int8, device=dev))
        Bs.append(ck.shape[0]), bgs.append(bgn), zs.append(int(Zc))
    ia = lambda v: (ctypes.c_int * n)(*v)
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().nrldpc_encode_groups(n, _ptr_array([g[0] for g in groups]), ia(Bs), ia(bgs), ia(zs),
                                                   int(bool(fix_fillers)), _ptr_array(dns), _stream_ptr()), "encode_groups")
    return dns


def decode_ref_batch(llr, Zc, bgn, L, algo="min-sum", alpha=1.0, beta=0.0, early_term=True, f64=True):
    """The generic (CSR) kernels on the 5G matrix; f64=True reproduces the reference's float64
    arithmetic exactly for 'min-sum'.  NumPy in / out.  Returns (ck int8[B,N'], status bool[B], iters int32[B])."""
    K, N, Nf, M = dims(bgn, Zc)
    llr = np.ascontiguousarray(llr, np.float64 if f64 else np.float32)
    assert llr.ndim == 2 and llr.shape[1] == N
    B = llr.shape[0]
    ck = np.empty((B, Nf), np.int8)
    status = np.empty(B, np.uint8)
    iters = np.empty(B, np.int32)
    a = {"min-sum": ALGO_MINSUM, "BP": ALGO_BP}[algo]
    _lib.check(_lib.lib().nrldpc_decode_soft_ref_host(llr.ctypes.data, int(f64), B, bgn, int(Zc), int(L), a, float(alpha),
                                                      float(beta), int(bool(early_term)), ck.ctypes.data,
                                                      status.ctypes.data, iters.ctypes.data), "decode_soft_ref")
    return ck, status.astype(bool), iters


def decode_bp_batch(llr, Zc, bgn, L, early_term=True):
    """nr_decode_ldpc(..., algo='BP') for B codeblocks on the quasi-cyclic sum-product kernel, float64 arithmetic
    (py5gphy/ldpc/nr_ldpc_decode.py:145-176).  llr float32 / float64 [B,N]: CUDA tensor (device path, asynchronous on the
    current stream, CUDA tensors out) or NumPy (host path).  Returns (ck int8[B,N'], status, iters)."""
    K, N, Nf, M = dims(bgn, Zc)
    if _is_torch(llr):
        import torch
        assert llr.is_cuda and llr.dtype in (torch.float32, torch.float64) and llr.is_contiguous()
        assert llr.ndim == 2 and llr.shape[1] == N
        B, dev = llr.shape[0], llr.device
        ck = torch.empty((B, Nf), dtype=torch.int8, device=dev)
        status = torch.empty((B,), dtype=torch.uint8, device=dev)
        iters = torch.empty((B,), dtype=torch.int32, device=dev)
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().nrldpc_decode_bp(llr.data_ptr(), int(llr.dtype == torch.float64), B, bgn, int(Zc), int(L),
                                                   int(bool(early_term)), ck.data_ptr(), status.data_ptr(), iters.data_ptr(),
                                                   _stream_ptr()), "decode_bp")
        return ck, status, iters
    llr = np.atleast_2d(np.asarray(llr))
    llr = np.ascontiguousarray(llr, np.float32 if llr.dtype == np.float32 else np.float64)
    assert llr.shape[1] == N
    B = llr.shape[0]
    ck = np.empty((B, Nf), np.int8)
    status = np.empty(B, np.uint8)
    iters = np.empty(B, np.int32)
    _lib.check(_lib.lib().nrldpc_decode_bp_host(llr.ctypes.data, int(llr.dtype == np.float64), B, bgn, int(Zc), int(L),
                                                int(bool(early_term)), ck.ctypes.data, status.ctypes.data, iters.ctypes.data), "decode_bp")
    return ck, status.astype(bool), iters


def decode_csr_batch(llr, rowptr, colidx, Nv, L, algo="min-sum", alpha=1.0, beta=0.0, early_term=True, f64=True):
    """decode_ldpc on an arbitrary CSR H (py5gphy/ldpc/nr_ldpc_decode.py:51-143)."""
    llr = np.ascontiguousarray(np.atleast_2d(llr), np.float64 if f64 else np.float32)
    rowptr = np.ascontiguousarray(rowptr, np.int32)
    colidx = np.ascontiguousarray(colidx, np.int32)
    B, M = llr.shape[0], rowptr.size - 1
    assert llr.shape[1] == Nv
    ck = np.empty((B, Nv), np.int8)
    status = np.empty(B, np.uint8)
    iters = np.empty(B, np.int32)
    a = {"min-sum": ALGO_MINSUM, "BP": ALGO_BP}[algo]
    _lib.check(_lib.lib().nrldpc_decode_csr_host(llr.ctypes.data, int(f64), B, M, int(Nv), rowptr.ctypes.data,
                                                 colidx.ctypes.data, int(L), a, float(alpha), float(beta),
                                                 int(bool(early_term)), ck.ctypes.data, status.ctypes.data,
                                                 iters.ctypes.data), "decode_csr")
    return ck, status.astype(bool), iters


def decode_bf_batch(llr, Zc, bgn, L):
    """nr_decode_ldpc(..., algo='BF') for B codeblocks (py5gphy/ldpc/ldpc_decoder_bit_flipping.py:5-73)."""
    K, N, Nf, M = dims(bgn, Zc)
    if _is_torch(llr):
        # device path: float32 / float64 CUDA tensor in, CUDA tensors out, async on the current stream
        import torch
        assert llr.is_cuda and llr.dtype in (torch.float32, torch.float64) and llr.is_contiguous()
        assert llr.ndim == 2 and llr.shape[1] == N
        B, dev = llr.shape[0], llr.device
        ck = torch.empty((B, Nf), dtype=torch.int8, device=dev)
        status = torch.empty((B,), dtype=torch.uint8, device=dev)
        iters = torch.empty((B,), dtype=torch.int32, device=dev)
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().nrldpc_decode_bf(llr.data_ptr(), int(llr.dtype == torch.float64), B, bgn, int(Zc),
                                                   int(L), ck.data_ptr(), status.data_ptr(), iters.data_ptr(),
                                                   _stream_ptr()), "decode_bf")
        return ck, status, iters
    llr = np.ascontiguousarray(np.atleast_2d(llr), np.float64)
    assert llr.shape[1] == N
    B = llr.shape[0]
    ck = np.empty((B, Nf), np.int8)
    status = np.empty(B, np.uint8)
    iters = np.empty(B, np.int32)
    _lib.check(_lib.lib().nrldpc_decode_bf_host(llr.ctypes.data, B, bgn, int(Zc), int(L), ck.ctypes.data,
                                                status.ctypes.data, iters.ctypes.data), "decode_bf")
    return ck, status.astype(bool), iters


def decode_bf_csr_batch(llr, rowptr, colidx, Nv, L):
    llr = np.ascontiguousarray(np.atleast_2d(llr), np.float64)
    rowptr = np.ascontiguousarray(rowptr, np.int32)
    colidx = np.ascontiguousarray(colidx, np.int32)
    B, M = llr.shape[0], rowptr.size - 1
    assert llr.shape[1] == Nv
    ck = np.empty((B, Nv), np.int8)
    status = np.empty(B, np.uint8)
    iters = np.empty(B, np.int32)
    _lib.check(_lib.lib().nrldpc_decode_bf_csr_host(llr.ctypes.data, B, M, int(Nv), rowptr.ctypes.data,
                                                    colidx.ctypes.data, int(L), ck.ctypes.data, status.ctypes.data,
                                                    iters.ctypes.data), "decode_bf_csr")
    return ck, status.astype(bool), iters


# ------------------------------------------------------------------ host placement for the host-buffer entry points

def bind_host_to_device(device_index=0):
    """Pin this process to the CPU cores next to CUDA device `device_index` (NVML's ideal affinity, intersected with
    the cores the container allows), so that pinned host buffers allocated afterwards land on the GPU's NUMA node
    and the host-buffer entry points do not pull LLRs across the socket interconnect.  One process per GPU, called
    before any pinned allocation.  Returns the core list now in force, or None when nothing was changed (no NVML,
    no sched_setaffinity, or the ideal cores are outside the allowed set)."""
    import os
    if not hasattr(os, "sched_setaffinity"):
        return None
    try:
        import pynvml
        import torch
        pynvml.nvmlInit()
        pr = torch.cuda.get_device_properties(device_index)
        bus = "%08x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
        h = pynvml.nvmlDeviceGetHandleByPciBusId(bus)
        ncpu = max(os.cpu_count() or 1, max(os.sched_getaffinity(0)) + 1)
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        ideal = {64 * w + b for w, x in enumerate(words) for b in range(64) if (int(x) >> b) & 1}
        want = sorted(ideal & set(os.sched_getaffinity(0)))
        if not want:
            return None
        os.sched_setaffinity(0, want)
        return want
    except Exception:
        return None


# ------------------------------------------------------------------ device-side Monte-Carlo helpers (torch tensors)

def random_bits(B, n, seed, device, offset=0):
    import torch
    out = torch.empty((B, n), dtype=torch.int8, device=device)
    with torch.cuda.device(device):
        _lib.check(_lib.lib().nrldpc_random_bits(out.data_ptr(), B * n, int(seed), int(offset), _stream_ptr()), "random_bits")
    return out


def awgn_llr(dn, snr_db, seed, offset=0, out=None):
    """BPSK + AWGN + LLR of for_test_5g_ldpc_encoder (py5gphy/ldpc/nr_ldpc_decode.py:252-257) on device."""
    import torch
    assert dn.is_cuda and dn.dtype == torch.int8 and dn.is_contiguous()
    if out is None:
        out = torch.empty(dn.shape, dtype=torch.float32, device=dn.device)
    with torch.cuda.device(dn.device):
        _lib.check(_lib.lib().nrldpc_awgn_llr(dn.data_ptr(), dn.numel(), float(snr_db), int(seed), int(offset),
                                              out.data_ptr(), _stream_ptr()), "awgn_llr")
    return out


def count_errors(ref, got, K, iters=None, counters=None):
    """counters int64[4] += {codeblocks, block errors, bit errors, iterations} (sim_ldpc_internal.py:61-62)."""
    import torch
    B = ref.shape[0]
    if counters is None:
        counters = torch.zeros(4, dtype=torch.int64, device=ref.device)
    with torch.cuda.device(ref.device):
        _lib.check(_lib.lib().nrldpc_count_errors(ref.data_ptr(), ref.stride(0), got.data_ptr(), got.stride(0), B, int(K),
                                                  iters.data_ptr() if iters is not None else None,
                                                  counters.data_ptr(), _stream_ptr()), "count_errors")
    return counters


# ---- bit-packed twins (SURVEY 8(d)'s K/8 + N/8 bytes per codeblock): int32 tensors of little-endian 32-bit words, bit k
# of a row at word k // 32, bit k % 32 -- the layout of decode_batch(..., want_info=True)["info"]

def random_bits_packed(rows, cols, seed, device, first_id=0, row_words=None):
    """`cols` Philox bits per row, packed; bit k of row j equals random_bits_rows' byte (same seed and id)."""
    import torch
    row_words = (cols + 31) // 32 if row_words is None else int(row_words)
    out = torch.empty((rows, row_words), dtype=torch.int32, device=device)
    with torch.cuda.device(device):
        _lib.check(_lib.lib().nrldpc_random_bits_packed_rows(out.data_ptr(), rows, cols, row_words, int(seed), int(first_id), 1,
                                                             _stream_ptr()), "random_bits_packed_rows")
    return out


def crc_attach_packed(words, A, poly):
    """crc.nr_crc_encode on packed rows, in place: the CRC of bits 0..A-1 goes to bits A..A+L-1.  Returns L."""
    poly_id = {"6": 0, "11": 1, "16": 2, "24A": 3, "24B": 4, "24C": 5}[poly]
    with __import__("torch").cuda.device(words.device):
        return _lib.check(_lib.lib().nrldpc_crc_attach_packed(words.data_ptr(), words.shape[0], int(A), poly_id, words.stride(0),
                                                              _stream_ptr()), "crc_attach_packed")


def encode_packed(ck_words, bgn, Zc):
    """encode_ldpc on packed codeblocks [B, K/32] -> [B, N/32] (no fillers, Zc a multiple of 32)."""
    import torch
    K, N, _, _ = dims(bgn, Zc)
    assert ck_words.is_cuda and ck_words.dtype == torch.int32 and ck_words.is_contiguous() and ck_words.shape[1] * 32 == K
    out = torch.empty((ck_words.shape[0], N // 32), dtype=torch.int32, device=ck_words.device)
    with torch.cuda.device(ck_words.device):
        _lib.check(_lib.lib().nrldpc_encode_packed(ck_words.data_ptr(), ck_words.shape[0], bgn, Zc, out.data_ptr(), _stream_ptr()),
                   "encode_packed")
    return out


def awgn_llr_packed(dn_words, cols, snr_db, seed, first_id=0, out=None):
    """awgn_llr_rows on packed dn: llr [rows, cols] float32, the same noise as the byte-per-bit function."""
    import torch
    rows = dn_words.shape[0]
    if out is None:
        out = torch.empty((rows, cols), dtype=torch.float32, device=dn_words.device)
    with torch.cuda.device(dn_words.device):
        _lib.check(_lib.lib().nrldpc_awgn_llr_packed_rows(dn_words.data_ptr(), rows, int(cols), dn_words.stride(0), float(snr_db),
                                                          int(seed), int(first_id), 1, out.data_ptr(), _stream_ptr()),
                   "awgn_llr_packed_rows")
    return out


def count_errors_packed(ref_words, got_words, K, iters=None, counters=None):
    import torch
    if counters is None:
        counters = torch.zeros(4, dtype=torch.int64, device=ref_words.device)
    with torch.cuda.device(ref_words.device):
        _lib.check(_lib.lib().nrldpc_count_errors_packed(ref_words.data_ptr(), ref_words.stride(0), got_words.data_ptr(),
                                                         got_words.stride(0), ref_words.shape[0], int(K),
                                                         iters.data_ptr() if iters is not None else None,
                                                         counters.data_ptr(), _stream_ptr()), "count_errors_packed")
    return counters


# ------------------------------------------------------------------ rate matching / recovery, HARQ, CRC (callers' side)

def _offsets(E_list):
    E = np.ascontiguousarray(E_list, np.int32).reshape(-1)
    off = np.zeros(E.size, np.int64)
    if E.size > 1:
        off[1:] = np.cumsum(E[:-1], dtype=np.int64)
    return E, off, int(E.sum(dtype=np.int64))


def ratematch_batch(dn, Ncb, E_list, k0, Qm):
    """ratematch_ldpc for the B codeblocks of a transport block + code block concatenation
    (py5gphy/ldpc/nr_ldpc_ratematch.py:64-97, py5gphy/nr_pdsch/nr_dlsch.py:66-68).
    dn int8 [B,N] (NumPy or CUDA tensor), E_list the B output lengths -> g int8 [sum(E)] of the same kind."""
    assert dn.ndim == 2
    B, N = dn.shape
    E, off, total = _offsets(E_list)
    assert E.size == B and N >= Ncb and Qm in [1, 2, 4, 6, 8] and not np.any(E % Qm)
    L = _lib.lib()
    if _is_torch(dn):
        import torch
        assert dn.is_cuda and dn.dtype == torch.int8 and dn.is_contiguous()
        g = torch.empty((total,), dtype=torch.int8, device=dn.device)
        dE, doff = torch.from_numpy(E).to(dn.device), torch.from_numpy(off).to(dn.device)
        with torch.cuda.device(dn.device):
            _lib.check(L.nrldpc_ratematch(dn.data_ptr(), B, N, int(Ncb), int(k0), int(Qm), dE.data_ptr(), doff.data_ptr(),
                                          g.data_ptr(), _stream_ptr()), "ratematch")
        return g
    dn = np.ascontiguousarray(dn, np.int8)
    g = np.empty(total, np.int8)
    _lib.check(L.nrldpc_ratematch_host(dn.ctypes.data, B, N, int(Ncb), int(k0), int(Qm), E.ctypes.data, g.ctypes.data), "ratematch")
    return g


def raterecover_batch(llr_g, E_list, Ncb, N, k0, Qm, Zc, K_apo, K, out_f64=True):
    """raterecover_ldpc for B codeblocks (py5gphy/ldpc/nr_ldpc_raterecover.py:6-65): llr_g = the
    concatenated received LLRs (float32 or float64; NumPy or CUDA tensor) -> [B,N] float64 (the
    reference's dtype) or float32 (out_f64=False: what the fp32 decoder consumes)."""
    E, off, total = _offsets(E_list)
    B = E.size
    assert llr_g.ndim == 1 and llr_g.shape[0] == total and not np.any(E % Qm) and np.all(E > 0)
    L = _lib.lib()
    if _is_torch(llr_g):
        import torch
        assert llr_g.is_cuda and llr_g.is_contiguous() and llr_g.dtype in (torch.float32, torch.float64)
        out = torch.empty((B, N), dtype=torch.float64 if out_f64 else torch.float32, device=llr_g.device)
        dE, doff = torch.from_numpy(E).to(llr_g.device), torch.from_numpy(off).to(llr_g.device)
        with torch.cuda.device(llr_g.device):
            _lib.check(L.nrldpc_raterecover(llr_g.data_ptr(), int(llr_g.dtype == torch.float64), B, int(N), int(Ncb), int(k0),
                                            int(Qm), int(K_apo - 2 * Zc), int(K - 2 * Zc), dE.data_ptr(), doff.data_ptr(),
                                            out.data_ptr(), int(out_f64), _stream_ptr()), "raterecover")
        return out
    in64 = np.asarray(llr_g).dtype != np.float32
    x = np.ascontiguousarray(llr_g, np.float64 if in64 else np.float32)
    out = np.empty((B, N), np.float64 if out_f64 else np.float32)
    _lib.check(L.nrldpc_raterecover_host(x.ctypes.data, int(in64), B, int(N), int(Ncb), int(k0), int(Qm), int(Zc), int(K_apo),
                                         int(K), E.ctypes.data, out.ctypes.data, int(out_f64)), "raterecover")
    return out


def harq_combine(new, cur):
    """HARQ soft combining of DLSCHDecode / ULSCH_decoding (py5gphy/nr_pdsch/nr_dlsch_decode.py:80-87), float64."""
    L = _lib.lib()
    if _is_torch(new):
        import torch
        assert new.is_cuda and cur.is_cuda and new.dtype == torch.float64 and cur.dtype == torch.float64
        assert new.is_contiguous() and cur.is_contiguous() and new.shape == cur.shape
        out = torch.empty_like(new)
        with torch.cuda.device(new.device):
            _lib.check(L.nrldpc_harq_combine(new.data_ptr(), cur.data_ptr(), new.numel(), out.data_ptr(), _stream_ptr()), "harq_combine")
        return out
    a = np.ascontiguousarray(new, np.float64)
    c = np.ascontiguousarray(cur, np.float64)
    assert a.shape == c.shape
    out = np.empty_like(a)
    _lib.check(L.nrldpc_harq_combine_host(a.ctypes.data, c.ctypes.data, a.size, out.ctypes.data), "harq_combine")
    return out


_POLY_ID = {"6": 0, "11": 1, "16": 2, "24A": 3, "24B": 4, "24C": 5}
_POLY_LEN = {"6": 6, "11": 11, "16": 16, "24A": 24, "24B": 24, "24C": 24}


def crc_encode_device(blk, poly):
    """nr_crc_encode for B blocks on CUDA tensors: int8 [B,A] -> int8 [B,A+L] (py5gphy/crc/crc.py:4-41)."""
    import torch
    key = str(poly).upper()
    assert key in _POLY_ID and blk.is_cuda and blk.dtype == torch.int8 and blk.is_contiguous() and blk.ndim == 2
    B, A = blk.shape
    out = torch.empty((B, A + _POLY_LEN[key]), dtype=torch.int8, device=blk.device)
    with torch.cuda.device(blk.device):
        _lib.check(_lib.lib().nrldpc_crc_encode(blk.data_ptr(), B, A, _POLY_ID[key], out.data_ptr(), _stream_ptr()), "crc_encode")
    return out


def crc_check_device(blkandcrc, poly):
    """nr_crc_decode's error flag for B blocks on CUDA tensors: int8 [B,A+L] -> uint8 [B] (py5gphy/crc/crc.py:43-88)."""
    import torch
    key = str(poly).upper()
    assert key in _POLY_ID and blkandcrc.is_cuda and blkandcrc.dtype == torch.int8 and blkandcrc.is_contiguous()
    B, n = blkandcrc.shape
    err = torch.empty((B,), dtype=torch.uint8, device=blkandcrc.device)
    with torch.cuda.device(blkandcrc.device):
        _lib.check(_lib.lib().nrldpc_crc_check(blkandcrc.data_ptr(), B, n - _POLY_LEN[key], _POLY_ID[key], err.data_ptr(),
                                               _stream_ptr()), "crc_check")
    return err


# ------------------------------------------------------------------ whole transport blocks (fused chain, host buffers)

class _PinnedBlock:
    """A block of the library's pinned host memory pool (nrldpc_host_alloc); returned to the pool when the last
    NumPy view of it is garbage-collected."""
    __slots__ = ("ptr", "nbytes", "__weakref__")
    live_bytes = 0   # pinned bytes currently owned by NumPy arrays handed to callers

    def __init__(self, nbytes):
        p = ctypes.c_void_p()
        _lib.check(_lib.lib().nrldpc_host_alloc(max(int(nbytes), 1), ctypes.byref(p)), "host_alloc")
        self.ptr, self.nbytes = p.value, int(nbytes)
        _PinnedBlock.live_bytes += self.nbytes

    @property
    def __array_interface__(self):
        return {"shape": (self.nbytes,), "typestr": "|u1", "data": (self.ptr, False), "version": 3}

    def __del__(self):
        try:
            _PinnedBlock.live_bytes -= self.nbytes
            _lib.lib().nrldpc_host_free(self.ptr)
        except Exception:   # interpreter shutdown
            pass


# A caller that keeps every returned soft buffer alive (16 HARQ processes x 23 MB is fine, a list of 10 000 results is not)
# must not be able to page-lock the machine's memory: beyond this many live bytes results come back in ordinary (pageable)
# NumPy memory, which every entry point accepts (staged copy).  NRLDPC_PINNED_LIMIT_MB overrides the 4 GiB default.
_PINNED_LIMIT = int(os.environ.get("NRLDPC_PINNED_LIMIT_MB", "4096")) << 20


def pinned_empty(shape, dtype):
    """np.empty on pinned host memory from the library's pool: the host-buffer entry points DMA from / into it directly,
    and the fused transport-block decoder stores its float64 soft buffer into it while it iterates."""
    dtype = np.dtype(dtype)
    n = int(np.prod(shape)) * dtype.itemsize
    if n == 0 or _PinnedBlock.live_bytes + n > _PINNED_LIMIT:
        return np.empty(shape, dtype)
    return np.asarray(_PinnedBlock(n)).view(dtype).reshape(shape)


def _llr_seq(x):
    """The received sequence as the C ABI takes it: contiguous float32 or float64 (arithmetic is float64 either way)."""
    x = np.asarray(x).reshape(-1)
    return np.ascontiguousarray(x, np.float32 if x.dtype == np.float32 else np.float64)


def sch_decode_host(llr_g, E_list, bgn, Zc, Ncb, k0, Qm, K_apo, A, L, alpha, beta, cur=None, want_soft=True):
    """Rate recovery + HARQ combining + min-sum decoding + CB/TB CRC of one transport block in two launches
    (nrldpc_sch_decode_host; py5gphy/nr_pdsch/nr_dlsch_decode.py:56-107).  NumPy in / out.
    Returns dict(tbblk int8[A], tb_err int, cb_err uint8[C], status bool[C], iters int32[C], soft float64[C,N] | None)."""
    E = np.ascontiguousarray(E_list, np.int32).reshape(-1)
    C = E.size
    K, N, Nf, M = dims(bgn, Zc)
    x = _llr_seq(llr_g)
    assert x.size == int(E.sum(dtype=np.int64))
    if cur is not None:
        cur = np.ascontiguousarray(cur, np.float64)
        assert cur.shape == (C, N)
    soft = pinned_empty((C, N), np.float64) if want_soft else None
    tbblk = pinned_empty((A,), np.int8)
    small = np.empty(1 + 2 * C, np.uint8)
    iters = np.empty(C, np.int32)
    _lib.check(_lib.lib().nrldpc_sch_decode_host(
        x.ctypes.data, int(x.dtype == np.float64), C, bgn, int(Zc), int(Ncb), int(k0), int(Qm), int(K_apo), E.ctypes.data,
        cur.ctypes.data if cur is not None else None, soft.ctypes.data if soft is not None else None, int(L), float(alpha),
        float(beta), int(A), tbblk.ctypes.data, small.ctypes.data, small.ctypes.data + 1, small.ctypes.data + 1 + C,
        iters.ctypes.data), "sch_decode")
    return dict(tbblk=tbblk, tb_err=int(small[0]), cb_err=small[1:1 + C], status=small[1 + C:].astype(bool), iters=iters, soft=soft)


def sch_recover_host(llr_g, E_list, bgn, Zc, Ncb, k0, Qm, K_apo, cur=None):
    """Rate recovery + HARQ combining of a transport block's codeblocks -> float64 [C,N] (new_LLr_dns)."""
    E = np.ascontiguousarray(E_list, np.int32).reshape(-1)
    C = E.size
    K, N, Nf, M = dims(bgn, Zc)
    x = _llr_seq(llr_g)
    assert x.size == int(E.sum(dtype=np.int64))
    if cur is not None:
        cur = np.ascontiguousarray(cur, np.float64)
        assert cur.shape == (C, N)
    soft = pinned_empty((C, N), np.float64)
    _lib.check(_lib.lib().nrldpc_sch_recover_host(x.ctypes.data, int(x.dtype == np.float64), C, N, int(Ncb), int(k0), int(Qm), int(Zc),
                                                  int(K_apo), K, E.ctypes.data, cur.ctypes.data if cur is not None else None,
                                                  soft.ctypes.data), "sch_recover")
    return soft


def sch_segment_host(trblk, C, K):
    """TB CRC attachment + code block segmentation + CB CRC: int8 [A] -> cbs int8 [C,K] with -1 fillers
    (py5gphy/nr_pdsch/nr_dlsch.py:29-46, py5gphy/ldpc/nr_ldpc_cbsegment.py:7-33)."""
    t = np.ascontiguousarray(trblk, np.int8).reshape(-1)
    cbs = np.empty((C, K), np.int8)
    _lib.check(_lib.lib().nrldpc_sch_segment_host(t.ctypes.data, t.size, int(C), int(K), cbs.ctypes.data), "sch_segment")
    return cbs


def encode_ratematch_host(cbs, bgn, Zc, Ncb, k0, Qm, E_list, fix_fillers=True):
    """LDPC encoding + rate matching + concatenation of a transport block's codeblocks: cbs int8 [C,K] (fillers -1,
    set to 0 in place when fix_fillers, like encode_ldpc) -> int8 [sum E]."""
    E = np.ascontiguousarray(E_list, np.int32).reshape(-1)
    assert cbs.dtype == np.int8 and cbs.flags.c_contiguous and cbs.ndim == 2 and cbs.shape[0] == E.size
    K, N, Nf, M = dims(bgn, Zc)
    assert cbs.shape[1] == K
    g = pinned_empty((int(E.sum(dtype=np.int64)),), np.int8)   # DMA target: no staged copy on the way back
    _lib.check(_lib.lib().nrldpc_encode_ratematch_host(cbs.ctypes.data, E.size, bgn, int(Zc), int(bool(fix_fillers)), int(Ncb), int(k0),
                                                       int(Qm), E.ctypes.data, g.ctypes.data), "encode_ratematch")
    return g


def sch_encode_host(trblk, C, bgn, Zc, Ncb, k0, Qm, E_list):
    """TB CRC + segmentation + LDPC encoding + rate matching + concatenation, device-resident in between:
    trblk int8 [A] -> int8 [sum E] (py5gphy/nr_pdsch/nr_dlsch.py:12-74)."""
    E = np.ascontiguousarray(E_list, np.int32).reshape(-1)
    t = np.ascontiguousarray(trblk, np.int8).reshape(-1)
    g = pinned_empty((int(E.sum(dtype=np.int64)),), np.int8)   # DMA target: no staged copy on the way back
    _lib.check(_lib.lib().nrldpc_sch_encode_host(t.ctypes.data, t.size, int(C), bgn, int(Zc), int(Ncb), int(k0), int(Qm),
                                                 E.ctypes.data, g.ctypes.data), "sch_encode")
    return g
