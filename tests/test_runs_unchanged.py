"""SURVEY 8 row (g) -- "nr_dlsch/nr_ulsch, the PDSCH/PUSCH scripts and the parameter-search scripts run unchanged".

tests/golden/calls_golden.npz holds call sequences recorded from the UNMODIFIED reference on the CPU of the build
container (tools/gen_golden_calls.py): the reference's own scripts (mixed-MS search, bit-flipping simulation, PUSCH
throughput example) and one PDSCH slot through Pdsch.process -> RX_process, with a recorder around every function
python_5gtoolbox_b200.install() rebinds.  Here every recorded top-level call is replayed, in order and under the same
NumPy seed, through the function install() binds in its place, and the returns (and in-place mutations) are compared:
bit-exact for integer results, float64-exact for LLRs and soft buffers, pickle equality for run_ldpc_simulation.
The fp32 min-sum decoder against the reference's float64: status always equal, bits equal whenever the codeblock
converged, a few differing bits tolerated on a non-converged codeblock."""
import json
import os
import pickle

import numpy as np
import pytest

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "calls_golden.npz")


def _load():
    with np.load(GOLD, allow_pickle=False) as z:
        meta = json.loads(str(z["__meta__"]))
        arrays = {k: z[k] for k in z.files if k != "__meta__"}
    return meta, arrays


def _unpack(o, arrays):
    if isinstance(o, dict):
        if "nd" in o:
            return arrays[o["nd"]].copy()
        if "seq" in o:
            v = [_unpack(x, arrays) for x in o["seq"]]
            return tuple(v) if o.get("tuple") else v
        if "dict" in o:
            return {k: _unpack(v, arrays) for k, v in o["dict"].items()}
    return o


def test_call_goldens_cover_the_rebound_surface():
    """CPU: the recording exists, names every scenario and only functions that install() rebinds."""
    from python_5gtoolbox_b200 import overlay
    meta, arrays = _load()
    assert set(meta) == {"mixed_ms", "bf", "sim", "pusch", "pdsch"}
    known = {f"{m}.{n}" for m, n, _ in overlay.rebound_names()}
    seen = {c["name"] for s in meta.values() for c in s["calls"]}
    assert seen <= known
    for want in ("py5gphy.ldpc.nr_ldpc_decode.nr_decode_ldpc", "py5gphy.ldpc.nr_ldpc_decode.for_test_5g_ldpc_encoder",
                 "scripts.internal.sim_ldpc_internal.run_ldpc_simulation", "py5gphy.nr_pdsch.nr_dlsch.DLSCHEncode",
                 "py5gphy.nr_pdsch.nr_dlsch_decode.DLSCHDecode", "py5gphy.nr_pusch.nr_ulsch.ULSCH_encoding_ratematch",
                 "py5gphy.nr_pusch.nr_ulsch_decode.ULSCH_decoding"):
        assert want in seen, want
    for name, fn in [(f"{m}.{n}", overlay.resolve(m, n)) for m, n, _ in overlay.rebound_names()]:
        assert callable(fn), name


def _same(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return a.shape == b.shape and np.array_equal(a, b)


def _check(name, got, want, ctx):
    short = name.rsplit(".", 1)[1]
    if short == "for_test_5g_ldpc_encoder":
        for g, w, what in zip(got, want, ("blkandcrc", "dn", "LLRin")):
            assert _same(g, w), (ctx, what)                       # same NumPy draws, same arithmetic: exact
    elif short == "nr_decode_ldpc":
        blk, ck, st = got
        wblk, wck, wst = want
        assert bool(st) == bool(wst), ctx
        assert np.asarray(ck).dtype == np.asarray(wck).dtype and np.asarray(ck).shape == np.asarray(wck).shape, ctx
        if wst or np.asarray(wck).dtype != np.int8:                # converged, or bit flipping (integer arithmetic): exact
            assert _same(ck, wck) and _same(blk, wblk), ctx
        else:
            assert np.mean(np.asarray(ck) != np.asarray(wck)) < 0.01, ctx
    elif short in ("DLSCHDecode", "ULSCH_decoding"):
        st, tb, new = got
        wst, wtb, wnew = want
        assert bool(st) == bool(wst), ctx
        assert np.asarray(new).dtype == np.float64 and _same(new, wnew), ctx      # float64-exact soft buffer
        assert np.asarray(tb).shape == np.asarray(wtb).shape, ctx
        if wst:
            assert _same(tb, wtb), ctx
        else:   # a block that fails in both: the hard bits of codeblocks that never converge are chaotic in the last bits
            assert np.mean(np.asarray(tb) != np.asarray(wtb)) < 0.25, ctx   # of fp32 vs float64 -- only a sanity bound
    elif short == "run_ldpc_simulation":
        pass   # compared through its pickle by the caller
    elif isinstance(want, (tuple, list)):
        assert len(got) == len(want), ctx
        for i, (g, w) in enumerate(zip(got, want)):
            if isinstance(w, np.ndarray):
                assert _same(g, w), (ctx, i)
            else:
                assert g == w, (ctx, i)
    elif isinstance(want, np.ndarray):
        assert _same(got, want), ctx
    else:
        assert got == want, ctx


@pytest.mark.gpu
@pytest.mark.parametrize("scenario", ["mixed_ms", "bf", "sim", "pusch", "pdsch"])
def test_replay_reference_call_sequences(scenario, tmp_path):
    from python_5gtoolbox_b200 import overlay, _lib
    assert _lib.lib().nrldpc_device_count() > 0, "no CUDA device (there is no CPU fallback)"
    meta, arrays = _load()
    sc = meta[scenario]
    np.random.seed(sc["seed"])
    n = 0
    for i, c in enumerate(sc["calls"]):
        mod, fname = c["name"].rsplit(".", 1)
        fn = overlay.resolve(mod, fname)
        args, kwargs, want = _unpack(c["args"], arrays), _unpack(c["kwargs"], arrays), _unpack(c["ret"], arrays)
        ctx = (scenario, i, fname)
        if fname == "run_ldpc_simulation":
            np.random.seed(sc["seed"])                          # the recorder seeds before every call of this scenario
            out = str(tmp_path / f"{i}.pickle")
            fn(*args[:-1], out, **kwargs, rng="numpy", shard="codeblock", verbose=False)
            with open(out, "rb") as f:
                got = pickle.load(f)
            assert got[0] == want[0] and got[1] == want[1], ctx           # sim_config, labels
            assert got[2] == want[2], (ctx, got[2], want[2])              # the BLER table, value for value
        else:
            got = fn(*args, **kwargs)
            _check(c["name"], got, want, ctx)
        for j, after in enumerate(c["after"]):                  # in-place side effects (encode_ldpc's filler fix, ...)
            if after is not None:
                assert _same(args[j], _unpack(after, arrays)), (ctx, "argument", j)
        n += 1
    assert n == len(sc["calls"]) and n > 0


def test_matplotlib_stub_and_workdir(tmp_path):
    """CPU: the matplotlib stand-in accepts the reference's plotting calls and does not break module-protocol probes (a stub
    that answered `__file__` / `__spec__` made `import torch` fail on the GPU box); the scratch directory of
    run_reference_script mirrors a reference root, with a writable out/ that holds copies of the stored results."""
    import importlib
    import sys
    from python_5gtoolbox_b200 import _mpl_stub, run_reference_script
    had = "matplotlib" in sys.modules
    stubbed = _mpl_stub.install()
    import matplotlib.pyplot as plt
    if stubbed:
        fig = plt.figure()
        plt.plot([1, 2], [3, 4], marker=".", label="x")
        plt.savefig("nowhere.png")
        plt.close(fig)
        import matplotlib
        with pytest.raises(AttributeError):
            matplotlib.__file__
        assert importlib.util.find_spec("torch") is not None
        import torch  # noqa: F401  (must import with the stub in place)
        if not had:
            del sys.modules["matplotlib"], sys.modules["matplotlib.pyplot"]
    ref = tmp_path / "ref"
    (ref / "py5gphy").mkdir(parents=True)
    (ref / "scripts").mkdir()
    (ref / "out").mkdir()
    (ref / "out" / "stored.pickle").write_bytes(b"x")
    wd = run_reference_script.make_workdir(str(ref), str(tmp_path / "wd"))
    assert os.path.islink(os.path.join(wd, "py5gphy")) and os.path.islink(os.path.join(wd, "scripts"))
    assert not os.path.islink(os.path.join(wd, "out")) and os.path.isfile(os.path.join(wd, "out", "stored.pickle"))
    with open(os.path.join(wd, "out", "new.pickle"), "wb") as f:   # writable, and the reference's own out/ is untouched
        f.write(b"y")
    assert sorted(os.listdir(ref / "out")) == ["stored.pickle"]
