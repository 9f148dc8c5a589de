"""install(): overlay this package's LDPC hot path onto an importable reference ``py5gphy``.

The reference's callers do ``from py5gphy.ldpc import nr_ldpc_decode`` and then call
``nr_ldpc_decode.nr_decode_ldpc(...)`` (py5gphy/nr_pdsch/nr_dlsch_decode.py:7,91;
scripts/internal/sim_ldpc_internal.py:7,50-58), so rebinding the functions inside the reference's own
module objects routes every caller through the CUDA path with their files byte-identical."""
import importlib

_REBIND = {
    "py5gphy.ldpc.nr_ldpc_encode": ("nr_ldpc_encode", ["encode_ldpc"]),
    "py5gphy.ldpc.nr_ldpc_decode": ("nr_ldpc_decode", ["nr_decode_ldpc", "decode_ldpc", "for_test_5g_ldpc_encoder"]),
    "py5gphy.ldpc.ldpc_decoder_bit_flipping": ("ldpc_decoder_bit_flipping", ["ldpc_decoder_BF"]),
    "py5gphy.ldpc.ldpc_info": ("ldpc_info", ["getH", "find_iLS", "gen_ldpc_para", "get_cbs_info"]),
    "py5gphy.ldpc.nr_ldpc_ratematch": ("nr_ldpc_ratematch", ["get_Er_ldpc", "get_k0", "ratematch_ldpc"]),
    "py5gphy.ldpc.nr_ldpc_raterecover": ("nr_ldpc_raterecover", ["raterecover_ldpc"]),
    "py5gphy.ldpc.nr_ldpc_cbsegment": ("nr_ldpc_cbsegment", ["ldpc_cbsegment"]),
}
# the callers either side of the path: whole-transport-block functions, batched
_REBIND_SCH = {
    "py5gphy.nr_pdsch.nr_dlsch": ("nr_pdsch.nr_dlsch", ["DLSCHEncode"]),
    "py5gphy.nr_pdsch.nr_dlsch_decode": ("nr_pdsch.nr_dlsch_decode", ["DLSCHDecode"]),
    "py5gphy.nr_pusch.nr_ulsch": ("nr_pusch.nr_ulsch", ["ULSCH_Crc_CodeBlockSegment", "ULSCH_encoding_ratematch"]),
    "py5gphy.nr_pusch.nr_ulsch_decode": ("nr_pusch.nr_ulsch_decode", ["ULSCH_decoding"]),
}
_saved = {}


def install(sch=True):
    """Rebind the reference's LDPC entry points (and, with sch=True, the whole-transport-block DL-SCH /
    UL-SCH functions) to the CUDA drop-ins.  Returns the list of rebound names."""
    done = []
    plan = [("python_5gtoolbox_b200.ldpc." + mine, modname, names) for modname, (mine, names) in _REBIND.items()]
    if sch:
        plan += [("python_5gtoolbox_b200." + mine, modname, names) for modname, (mine, names) in _REBIND_SCH.items()]
    for mine, modname, names in plan:
        ref = importlib.import_module(modname)
        new = importlib.import_module(mine)
        for n in names:
            _saved.setdefault((modname, n), getattr(ref, n))
            setattr(ref, n, getattr(new, n))
            done.append(f"{modname}.{n}")
    return done


def uninstall():
    for (modname, n), fn in _saved.items():
        setattr(importlib.import_module(modname), n, fn)
    _saved.clear()
