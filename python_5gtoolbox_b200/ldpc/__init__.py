"""Drop-in mirrors of the reference's py5gphy/ldpc modules (same names, signatures, error behaviour)."""
from . import ldpc_info, nr_ldpc_encode, nr_ldpc_decode, ldpc_decoder_bit_flipping  # noqa: F401
from . import nr_ldpc_ratematch, nr_ldpc_raterecover, nr_ldpc_cbsegment  # noqa: F401
