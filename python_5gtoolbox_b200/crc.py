"""Mirror of py5gphy/crc/crc.py nr_crc_encode / nr_crc_decode (:4-41, :43-88) on the CUDA CRC kernel.

One int8 per bit, like the reference.  The optional RNTI mask (:35-38) is a <=24-bit XOR done on the
host around the device call."""
import numpy as np

from . import _lib

POLY_ID = {"6": 0, "11": 1, "16": 2, "24A": 3, "24B": 4, "24C": 5}
POLY_LEN = {"6": 6, "11": 11, "16": 16, "24A": 24, "24B": 24, "24C": 24}


def _mask_bits(mask, L):
    bits = np.array([(mask >> (23 - i)) & 1 for i in range(24)], np.int8)  # MSB first (:36)
    return bits[24 - L:]


def _poly(poly):
    key = str(poly).upper()
    assert key in POLY_ID  # the reference asserts in _get_crcpoly (:107-108)
    return key


def nr_crc_encode_batch(blk, poly):
    key = _poly(poly)
    blk = np.ascontiguousarray(np.atleast_2d(blk), np.int8)
    assert (not np.any(blk < 0)) and (not np.any(blk > 1))
    B, A = blk.shape
    out = np.empty((B, A + POLY_LEN[key]), np.int8)
    _lib.check(_lib.lib().nrldpc_crc_encode_host(blk.ctypes.data, B, A, POLY_ID[key], out.ctypes.data), "crc_encode")
    return out


def nr_crc_encode(blk, poly, mask=0):
    """blkandcrc = nr_crc_encode(blk, poly, mask=0) -- py5gphy/crc/crc.py:4-41"""
    key = _poly(poly)
    blk = np.asarray(blk)
    out = nr_crc_encode_batch(blk.reshape(1, -1), key)[0]
    if mask:
        L = POLY_LEN[key]
        out[blk.size:] ^= _mask_bits(mask, L)
    return out


def nr_crc_decode(blkandcrc, poly, mask=0):
    """blk, err = nr_crc_decode(blkandcrc, poly, mask=0) -- py5gphy/crc/crc.py:43-88"""
    key = _poly(poly)
    x = np.asarray(blkandcrc)
    assert (not np.any(x < 0)) and (not np.any(x > 1))
    x = np.ascontiguousarray(x, np.int8).reshape(1, -1).copy()
    L = POLY_LEN[key]
    A = x.shape[1] - L
    if mask:
        x[0, A:] ^= _mask_bits(mask, L)  # dividing the unmasked word == masking the remainder (linear)
    err = np.empty(1, np.uint8)
    _lib.check(_lib.lib().nrldpc_crc_check_host(x.ctypes.data, 1, A, POLY_ID[key], err.ctypes.data), "crc_check")
    return np.asarray(blkandcrc).astype("i1")[0:A], int(err[0])
