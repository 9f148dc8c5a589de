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
# the Monte-Carlo driver the LDPC scripts call (scripts/sim_ldpc_decoder.py:45-47, NMS_/OMS_/mixed_MS_*search*.py): the
# batched one, same signature, stopping rule and pickle layout
_REBIND_SIM = {
    "scripts.internal.sim_ldpc_internal": ("sim", ["run_ldpc_simulation"]),
}
_saved = {}


def rebound_names(sch=True, sim=True):
    """[(reference module, attribute, this package's module)] that install() rebinds."""
    plan = [(modname, n, "python_5gtoolbox_b200.ldpc." + mine) for modname, (mine, names) in _REBIND.items() for n in names]
    if sch:
        plan += [(modname, n, "python_5gtoolbox_b200." + mine) for modname, (mine, names) in _REBIND_SCH.items() for n in names]
    if sim:
        plan += [(modname, n, "python_5gtoolbox_b200." + mine) for modname, (mine, names) in _REBIND_SIM.items() for n in names]
    return plan


def resolve(modname, name):
    """The drop-in this package binds to `modname.name` (the function a replayed reference call goes to)."""
    for m, n, mine in rebound_names():
        if (m, n) == (modname, name):
            return getattr(importlib.import_module(mine), n)
    raise KeyError(f"{modname}.{name} is not a rebound name")


def install(sch=True, sim=True, stub_matplotlib=True):
    """Rebind the reference's LDPC entry points (with sch=True also the whole-transport-block DL-SCH / UL-SCH functions,
    with sim=True also the scripts' Monte-Carlo driver run_ldpc_simulation, when the reference's ``scripts`` package is
    importable) to the CUDA drop-ins.  stub_matplotlib: a box without matplotlib gets a do-nothing stand-in, because
    the reference's script modules import it at module level.  Returns the list of rebound names."""
    if stub_matplotlib:
        from . import _mpl_stub
        _mpl_stub.install()
    done = []
    for modname, n, mine in rebound_names(sch, sim):
        try:
            ref = importlib.import_module(modname)
        except ImportError:
            if modname.startswith("scripts."):
                continue   # py5gphy importable without the scripts directory: nothing to rebind there
            raise
        new = importlib.import_module(mine)
        _saved.setdefault((modname, n), getattr(ref, n))
        setattr(ref, n, getattr(new, n))
        done.append(f"{modname}.{n}")
    return done


def uninstall():
    for (modname, n), fn in _saved.items():
        setattr(importlib.import_module(modname), n, fn)
    _saved.clear()
