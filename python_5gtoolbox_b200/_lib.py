"""ctypes loader of libnrldpc_b200.so (the C ABI declared in include/nrldpc_b200.h)."""
import ctypes
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
# NRLDPC_SO=path: kernel experiments load another build of the library (tools/build_variant.sh, same-box A/B runs)
SO_PATH = os.path.abspath(os.environ["NRLDPC_SO"]) if os.environ.get("NRLDPC_SO") else os.path.join(_HERE, "libnrldpc_b200.so")
_lib = None

EINVAL, ECUDA, ENOMEM, ENODEV = -1, -2, -3, -4


class NrLdpcError(RuntimeError):
    """A CUDA-side failure.  Bad arguments raise AssertionError instead, like the reference does."""


def build(force=False, verbose=False):
    """Compile every CUDA source for sm_100a into python_5gtoolbox_b200/libnrldpc_b200.so (in-tree)."""
    src = os.path.join(_HERE, "csrc")
    if force:
        subprocess.run(["make", "-C", src, "clean"], check=True, capture_output=True)
    r = subprocess.run(["make", "-C", src, "-j8"], capture_output=True, text=True)
    if verbose or r.returncode:
        print(r.stdout[-4000:], r.stderr[-4000:])
    if r.returncode:
        raise RuntimeError("nvcc build of libnrldpc_b200.so failed")
    return SO_PATH


def _declare(L):
    c = ctypes
    p = c.c_void_p
    i, f, d, ll, ull = c.c_int, c.c_float, c.c_double, c.c_longlong, c.c_ulonglong
    ip = c.POINTER(c.c_int)
    sigs = {
        "nrldpc_version": (i, []),
        "nrldpc_last_error": (c.c_char_p, []),
        "nrldpc_device_count": (i, []),
        "nrldpc_find_ils": (i, [i]),
        "nrldpc_dims": (i, [i, i, ip, ip, ip, ip]),
        "nrldpc_build_csr": (i, [i, i, p, p]),
        "nrldpc_encode": (i, [p, i, i, i, i, p, p]),
        "nrldpc_encode_host": (i, [p, i, i, i, i, p]),
        "nrldpc_encode_packed": (i, [p, i, i, i, p, p]),
        "nrldpc_random_bits_packed_rows": (i, [p, ll, ll, ll, ull, ll, ll, p]),
        "nrldpc_crc_attach_packed": (i, [p, i, i, i, ll, p]),
        "nrldpc_awgn_llr_packed_rows": (i, [p, ll, ll, ll, f, ull, ll, ll, p, p]),
        "nrldpc_count_errors_packed": (i, [p, ll, p, ll, i, i, p, p, p]),
        "nrldpc_decode_minsum": (i, [p, i, i, i, i, f, f, i, p, p, p, p, p]),
        "nrldpc_decode_minsum_host": (i, [p, i, i, i, i, f, f, i, p, p, p, p]),
        "nrldpc_decode_minsum_host_f16": (i, [p, i, i, i, i, f, f, i, p, p, p, p]),
        "nrldpc_decode_minsum_geometry": (i, [i, i, ip, ip, ip]),
        "nrldpc_decode_minsum_groups": (i, [i, p, p, p, p, i, f, f, i, p, p, p, p, p]),
        "nrldpc_encode_groups": (i, [i, p, p, p, p, i, p, p]),
        "nrldpc_decode_csr_host": (i, [p, i, i, i, i, p, p, i, i, d, d, i, p, p, p]),
        "nrldpc_decode_soft_ref_host": (i, [p, i, i, i, i, i, i, d, d, i, p, p, p]),
        "nrldpc_decode_bf_csr_host": (i, [p, i, i, i, p, p, i, p, p, p]),
        "nrldpc_decode_bf_host": (i, [p, i, i, i, i, p, p, p]),
        "nrldpc_decode_bf": (i, [p, i, i, i, i, i, p, p, p, p]),
        "nrldpc_awgn_llr": (i, [p, ll, f, ull, ull, p, p]),
        "nrldpc_random_bits": (i, [p, ll, ull, ull, p]),
        "nrldpc_awgn_llr_rows": (i, [p, ll, ll, f, ull, ll, ll, p, p]),
        "nrldpc_random_bits_rows": (i, [p, ll, ll, ull, ll, ll, p]),
        "nrldpc_count_errors": (i, [p, ll, p, ll, i, i, p, p, p]),
        "nrldpc_crc_encode": (i, [p, i, i, i, p, p]),
        "nrldpc_crc_check": (i, [p, i, i, i, p, p]),
        "nrldpc_crc_encode_host": (i, [p, i, i, i, p]),
        "nrldpc_crc_check_host": (i, [p, i, i, i, p]),
        "nrldpc_ratematch": (i, [p, i, i, i, i, i, p, p, p, p]),
        "nrldpc_ratematch_host": (i, [p, i, i, i, i, i, p, p]),
        "nrldpc_raterecover": (i, [p, i, i, i, i, i, i, i, i, p, p, p, i, p]),
        "nrldpc_raterecover_host": (i, [p, i, i, i, i, i, i, i, i, i, p, p, i]),
        "nrldpc_harq_combine": (i, [p, p, ll, p, p]),
        "nrldpc_harq_combine_host": (i, [p, p, ll, p]),
        "nrldpc_decode_bp": (i, [p, i, i, i, i, i, i, p, p, p, p]),
        "nrldpc_decode_bp_host": (i, [p, i, i, i, i, i, i, p, p, p]),
        "nrldpc_host_alloc": (i, [c.c_size_t, c.POINTER(p)]),
        "nrldpc_host_free": (i, [p]),
        "nrldpc_sch_recover": (i, [p, i, i, i, i, i, i, i, i, p, p, p, p, p, p]),
        "nrldpc_sch_recover_host": (i, [p, i, i, i, i, i, i, i, i, i, p, p, p]),
        "nrldpc_sch_decode": (i, [p, i, i, i, i, i, i, i, i, p, p, p, p, i, f, f, i, p, p, p, p, p, p, p]),
        "nrldpc_sch_decode_host": (i, [p, i, i, i, i, i, i, i, i, p, p, p, i, f, f, i, p, p, p, p, p]),
        "nrldpc_sch_segment": (i, [p, i, i, i, p, p]),
        "nrldpc_sch_segment_host": (i, [p, i, i, i, p]),
        "nrldpc_encode_ratematch": (i, [p, i, i, i, i, i, i, i, p, p, p, p]),
        "nrldpc_encode_ratematch_host": (i, [p, i, i, i, i, i, i, i, p, p]),
        "nrldpc_sch_encode_host": (i, [p, i, i, i, i, i, i, i, p, p]),
    }
    lenient = bool(os.environ.get("NRLDPC_SO"))   # kernel experiments load older / partial builds of the library
    for name, (res, args) in sigs.items():
        if lenient and not hasattr(L, name):
            continue
        fn = getattr(L, name)  # AttributeError here = the .so does not export what the header declares
        fn.restype = res
        fn.argtypes = args
    return sigs


def lib():
    """The loaded library.  Fails loudly when the CUDA extension has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(SO_PATH):
            raise NrLdpcError(
                f"{SO_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc, sm_100a).  python_5gtoolbox_b200 has no CPU fallback.")
        L = ctypes.CDLL(SO_PATH)
        _declare(L)
        _lib = L
    return _lib


def exported_symbols():
    return sorted(_declare(ctypes.CDLL(SO_PATH)).keys())


def check(rc, what=""):
    if rc >= 0:
        return rc
    msg = lib().nrldpc_last_error().decode()
    if rc == EINVAL:
        raise AssertionError(f"{what}: {msg}")
    raise NrLdpcError(f"{what}: {msg} (code {rc})")
