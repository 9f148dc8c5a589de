#!/usr/bin/env python3
"""Call-sequence goldens: run the UNMODIFIED reference (scripts and classes) on the CPU of the build container with a
recorder around every function python_5gtoolbox_b200.install() rebinds, and store every top-level call -- arguments
before the call, arrays mutated by the call, return values -- in tests/golden/calls_golden.npz.  A `-m gpu` test
(tests/test_runs_unchanged.py) replays the calls through the drop-ins and compares.

    python tools/gen_golden_calls.py [--ref /root/reference] [--only NAME ...]

Scenarios (NumPy's global RNG is seeded at the start of each, the reference draws from it):
  mixed_ms    scripts/mixed_MS_ldpc_search_best_pair.py as shipped (sim_flag = 1), stopped after the first codeblocks
  bf          scripts/sim_ldpc_decoder_bf.py as shipped, stopped after the first codeblocks
  sim         scripts.internal.sim_ldpc_internal.run_ldpc_simulation, the call the search scripts make, on a reduced
              grid (whole call, its pickle is the return value)
  pusch       scripts/NR_PUSCH_throughput_example.py as shipped (20 PRB, MCS5, 2 layers, Rayleigh, MMSE-IRC ...),
              stopped after the first slots: NrPUSCH.process -> RX_process
  pdsch       one PDSCH slot through Pdsch.process -> RX_process (AWGN), built from the default configuration files
Only depth-0 calls are recorded: what the reference does inside a rebound function is its own business.
"""
import argparse
import copy
import io
import json
import os
import pickle
import runpy
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


class Stop(Exception):
    pass


class Recorder:
    def __init__(self):
        self.depth = 0
        self.calls = []
        self.limits = {}
        self.counts = {}
        self.transparent = set()   # names that are run but not recorded: the calls they make stay top-level

    def wrap(self, fn, qual):
        def w(*a, **k):
            if self.depth or qual in self.transparent:
                return fn(*a, **k)
            before = copy.deepcopy((a, k))
            self.depth += 1
            t0 = time.time()
            try:
                r = fn(*a, **k)
            finally:
                self.depth -= 1
            self.calls.append(dict(name=qual, args=before[0], kwargs=before[1], after=copy.deepcopy(a), ret=copy.deepcopy(r),
                                   seconds=time.time() - t0))
            self.counts[qual] = self.counts.get(qual, 0) + 1
            lim = self.limits.get(qual)
            if lim is not None and self.counts[qual] >= lim:
                raise Stop()
            return r
        return w


def workdir(ref):
    """A writable directory laid out like the reference root (the scripts use relative paths and write into out/)."""
    d = tempfile.mkdtemp(prefix="refrun_")
    for name in os.listdir(ref):
        if name != "out":
            os.symlink(os.path.join(ref, name), os.path.join(d, name))
    os.mkdir(os.path.join(d, "out"))
    return d


def arm(rec):
    import importlib
    from python_5gtoolbox_b200 import overlay, _mpl_stub
    _mpl_stub.install()
    originals = []
    for modname, n, _ in overlay.rebound_names():
        m = importlib.import_module(modname)
        originals.append((m, n, getattr(m, n)))
        setattr(m, n, rec.wrap(getattr(m, n), f"{modname}.{n}"))
    return originals


def disarm(originals):
    for m, n, fn in originals:
        setattr(m, n, fn)


def run_script(modname, rec, seed, limits, transparent=()):
    rec.limits, rec.counts, rec.transparent = dict(limits), {}, set(transparent)
    np.random.seed(seed)
    start = len(rec.calls)
    try:
        runpy.run_module(modname, run_name="__main__")
    except Stop:
        pass
    return rec.calls[start:]


def scenario_sim(rec, seed):
    from scripts.internal import sim_ldpc_internal
    rec.limits, rec.counts, rec.transparent = {}, {}, set()
    out = []
    # (SNRs chosen so that the reference's stopping rule ends every point at its first or second checkpoint: minutes, not hours)
    for args in [(12, 2, '24A', ['mixed-MS'], [], [], [[0.7, 0.5], [0.8, 0.3]], [32], [-2.5], "out/calls_sim_a.pickle"),
                 (10, 1, '24A', ['BF', 'NMS'], [0.8], [], [], [16], [0.5], "out/calls_sim_b.pickle")]:
        np.random.seed(seed)
        start = len(rec.calls)
        sim_ldpc_internal.run_ldpc_simulation(*args)
        with open(args[-1], "rb") as f:
            rec.calls[start]["ret"] = pickle.load(f)   # the function returns None: its result is the pickle
        out += rec.calls[start:]
    return out


def scenario_pdsch(rec, seed):
    """One PDSCH slot, AWGN: waveform generation with Pdsch.process inside, receive low-PHY, LS + DFT channel estimation,
    Pdsch.RX_process (MMSE-IRC) -- the steps of scripts/internal/sim_pdsch_throughput_internal.py, whose own import of the
    unshipped tests package keeps it from being imported here."""
    from scripts.internal import default_config_files
    from py5gphy.common import nr_slot
    from py5gphy.nr_pdsch import nr_pdsch
    from py5gphy.nr_waveform import nr_dl_waveform
    from py5gphy.nr_lowphy import rx_lowphy_process
    from py5gphy.channel_model import nr_channel_model, AWGN_channel_model
    from py5gphy.channel_estimate import nr_channel_estimation
    rec.limits, rec.counts, rec.transparent = {}, {}, set()
    np.random.seed(seed)
    start = len(rec.calls)
    cfg = default_config_files.read_DL_default_config_files()
    wf, car, pd = cfg["DL_waveform_config"], cfg["DL_carrier_config"], cfg["pdsch_config"]
    Nt, Nr, BW, scs = 2, 2, 20, 30
    prb = nr_slot.get_carrier_prb_size(scs, BW)
    fs = nr_slot.get_FFT_IFFT_size(prb) * scs * 1000 * 2
    wf.update(numofslots=1, startSFN=0, startslot=0, samplerate_in_mhz=fs / 1e6)
    car.update(BW=BW, scs=scs, num_of_ant=Nt, Nr=Nr, maxMIMO_layers=Nt)
    pd.update(mcs_index=11, num_of_layers=Nt, rv=[0], data_source=[], mcs_table="256QAM", precoding_matrix=np.empty(0),
              StartSymbolIndex=2, NrOfSymbols=12)
    pd["ResAlloType1"]["RBSize"] = 40
    pd["ResAlloType1"]["RBStart"] = 0
    pd["DMRS"]["nNIDnSCID"] = 1
    pd["DMRS"]["NumCDMGroupsWithoutData"] = 1
    pd["DMRS"]["DMRSAddPos"] = 1
    pd["codebook"]["enable"] = "False"
    car.update(PCI=1)
    carrier_freq = car["carrier_frequency_in_mhz"] * 1e6
    cm_cfg = nr_channel_model.gen_channel_model_config("AWGN", ["customized", "uniform", "DL", [0, 0]], Nt, Nr, 0, 0, 0, [], 0, np.empty(0), 0)
    p = nr_pdsch.Pdsch(pd, car)
    chan = AWGN_channel_model.AWGNChannelModel(cm_cfg, -25, carrier_freq, fs, scs, 0)
    Dm = chan.gen_Dm(1)
    _, _, dl_wave, _ = nr_dl_waveform.gen_dl_waveform(wf, car, nrPdsch_list=[p], Dm=Dm)
    rx = chan.filter(dl_wave)
    _, rx_fd = rx_lowphy_process.waveform_rx_processing(rx, car, fs)
    rx_slot = rx_fd[:, 0:prb * 12 * 14]
    H_LS, RS_info = p.H_LS_est(rx_slot, 0)
    ce = nr_channel_estimation.NrChannelEstimation(H_LS, RS_info, {"CE_algo": "DFT_symmetric", "L_symm_left_in_ns": 1400, "L_symm_right_in_ns": 1200,
                                                                   "eRB": 4, "enable_TO_comp": True, "enable_FO_est": False, "enable_FO_comp": False})
    H, cov = ce.channel_est(0)
    st, tb, new = p.RX_process(rx_slot, 0, {"algo": "MMSE-IRC"}, H, cov, {"L": 32, "algo": "min-sum", "alpha": 0.8, "beta": 0.3}, ce)
    print("pdsch slot: status", st, "TBSize", p.info["TBSize"])
    return rec.calls[start:]


# ---- serialisation: JSON structure + arrays in an npz

def pack(obj, arrays, prefix):
    if isinstance(obj, np.ndarray):
        key = f"{prefix}"
        arrays[key] = obj
        return {"nd": key}
    if isinstance(obj, (np.bool_,)):
        return bool(obj)
    if isinstance(obj, (np.integer,)):
        return int(obj)
    if isinstance(obj, (np.floating,)):
        return float(obj)
    if isinstance(obj, (list, tuple)):
        return {"seq": [pack(x, arrays, f"{prefix}_{i}") for i, x in enumerate(obj)], "tuple": isinstance(obj, tuple)}
    if isinstance(obj, dict):
        return {"dict": {str(k): pack(v, arrays, f"{prefix}_{k}") for k, v in obj.items()}}
    if obj is None or isinstance(obj, (bool, int, float, str)):
        return obj
    raise TypeError(f"cannot store {type(obj)}")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", default="/root/reference")
    ap.add_argument("--only", nargs="*")
    ap.add_argument("--out", default=os.path.join(ROOT, "tests", "golden", "calls_golden.npz"))
    args = ap.parse_args()
    wd = workdir(args.ref)
    os.chdir(wd)
    sys.path.insert(0, wd)
    rec = Recorder()
    originals = arm(rec)
    dec = "py5gphy.ldpc.nr_ldpc_decode.nr_decode_ldpc"
    scen = {
        "mixed_ms": lambda: run_script("scripts.mixed_MS_ldpc_search_best_pair", rec, 101, {dec: 16},
                                       transparent=["scripts.internal.sim_ldpc_internal.run_ldpc_simulation"]),
        "bf": lambda: run_script("scripts.sim_ldpc_decoder_bf", rec, 102, {dec: 24}),
        "sim": lambda: scenario_sim(rec, 103),
        "pusch": lambda: run_script("scripts.NR_PUSCH_throughput_example", rec, 104, {"py5gphy.nr_pusch.nr_ulsch_decode.ULSCH_decoding": 3}),
        "pdsch": lambda: scenario_pdsch(rec, 105),
    }
    seeds = {"mixed_ms": 101, "bf": 102, "sim": 103, "pusch": 104, "pdsch": 105}
    arrays, meta = {}, {}
    if args.only and os.path.exists(args.out):   # keep the other scenarios of an earlier run
        with np.load(args.out, allow_pickle=False) as z:
            meta = json.loads(str(z["__meta__"]))
            arrays = {k: z[k] for k in z.files if k != "__meta__"}
    for name, fn in scen.items():
        if args.only and name not in args.only:
            continue
        t0 = time.time()
        buf = io.StringIO()
        calls = fn()
        for k in [k for k in arrays if k.startswith(name + "_")]:
            del arrays[k]
        meta[name] = {"seed": seeds[name], "calls": [
            {"name": c["name"], "seconds": round(c["seconds"], 3),
             "args": pack(c["args"], arrays, f"{name}_{i}_a"), "kwargs": pack(c["kwargs"], arrays, f"{name}_{i}_k"),
             "after": pack(c["after"], arrays, f"{name}_{i}_m"), "ret": pack(c["ret"], arrays, f"{name}_{i}_r")}
            for i, c in enumerate(calls)]}
        print(f"{name}: {len(calls)} top-level calls recorded in {time.time() - t0:.0f} s:",
              {n: sum(1 for c in calls if c['name'] == n) for n in sorted({c['name'] for c in calls})}, flush=True)
        finish(meta, arrays, args.out, [name])   # saved after every scenario: hours of CPU work are not lost to a later failure
    disarm(originals)


def finish(meta, arrays, out, fresh):
    # `after` duplicates `args` unless the call mutated an array: drop the unchanged ones
    for name, m in meta.items():
        if name not in fresh:
            continue
        for c in m["calls"]:
            def same(a, b):
                if isinstance(a, dict) and "nd" in a:
                    return isinstance(b, dict) and "nd" in b and np.array_equal(arrays[a["nd"]], arrays[b["nd"]], equal_nan=False) \
                        and arrays[a["nd"]].dtype == arrays[b["nd"]].dtype
                return True
            seq_a, seq_m = c["args"]["seq"], c["after"]["seq"]
            keep = []
            for j, (x, y) in enumerate(zip(seq_a, seq_m)):
                if isinstance(y, dict) and "nd" in y:
                    if same(x, y):
                        del arrays[y["nd"]]
                        keep.append(None)
                    else:
                        keep.append(y)
                else:
                    keep.append(None)
            c["after"] = keep
    used = set()

    def walk(o):
        if isinstance(o, dict):
            if "nd" in o:
                used.add(o["nd"])
            for v in o.values():
                walk(v)
        elif isinstance(o, list):
            for v in o:
                walk(v)
    walk(meta)
    for k in [k for k in arrays if k not in used]:
        del arrays[k]
    np.savez_compressed(out, __meta__=np.array(json.dumps(meta)), **arrays)
    print("wrote", out, f"{os.path.getsize(out) / 1e6:.2f} MB", flush=True)


if __name__ == "__main__":
    main()
