"""Batched Monte-Carlo driver (python_5gtoolbox_b200/sim.py): host logic on CPU, the sharded N>1 path
with world_size-2 gloo, and (GPU) the whole thing against the oracle on identical NumPy seeds."""
import os
import pickle
import socket
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class OracleBackend:
    """CPU stand-in for CudaBackend used ONLY to test the driver's sharding / stopping logic."""
    reduce_device = None

    def count_failures(self, Zc, bgn, snr_db, crcpoly, algo, L, alpha, beta, n, take, rng, seed, offset):
        from oracle import oracle as O
        assert rng == "numpy"
        K, N = ((22, 66) if bgn == 1 else (10, 50))
        K, N = K * Zc, N * Zc
        crc_len = 24 if crcpoly in ['24A', '24B'] else 16
        sigma = 10 ** (-snr_db / 20)
        fails = 0
        want = set(take)
        for b in range(n):
            inb = np.random.randint(2, size=K - crc_len)
            nz = np.random.normal(0, sigma, N)
            if b not in want:
                continue
            blk = O.nr_crc_encode(inb.astype("i1"), crcpoly)
            dn = O.encode_ldpc(blk.copy(), bgn)
            llr = 2 * ((1 - 2 * dn) + nz) / 10 ** (-snr_db / 10)
            out, ck, st, it = O.nr_decode_ldpc(llr, Zc, bgn, L, algo, alpha, beta)   # un-rounded float64, like the reference's scripts
            fails += not np.array_equal(out, blk)
        return fails


def test_plan_and_stopping_rule():
    from python_5gtoolbox_b200 import sim
    plan = sim.test_plan(['BP', 'min-sum', 'NMS', 'OMS', 'mixed-MS'], [0.7], [0.5], [[0.8, 0.3], [0.7, 0.3]], [16])
    # the labels stored in the reference's pickles (out/ldpc_decode_result_opt_2.pickle)
    assert [p[0] for p in plan] == ['BP L=16', 'min-sum L=16', 'NMS-alpha=0.7-L=16', 'OMS-beta=0.5-L=16',
                                    'mixed-MS-[alpha,beta]=[0.8,0.3]-L=16', 'mixed-MS-[alpha,beta]=[0.7,0.3]-L=16']
    assert sim.stop_now(1000, 50) and not sim.stop_now(1000, 49)
    assert sim.stop_now(2000, 25) and not sim.stop_now(2000, 24) and not sim.stop_now(1500, 900)
    assert sim.stop_now(4000, 10) and not sim.stop_now(4000, 9)
    assert sim.stop_now(10000, 0)


_WORKER = r"""
import os, sys, pickle
import numpy as np
sys.path.insert(0, sys.argv[1])
rank, world, port, out = int(sys.argv[2]), int(sys.argv[3]), sys.argv[4], sys.argv[5]
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
from python_5gtoolbox_b200 import sim
from tests.test_sim_driver import OracleBackend
np.random.seed(77)
shard = sys.argv[6]
res = sim.run_ldpc_simulation(2, 2, '16', ['NMS', 'OMS'], [0.8], [0.3], [], [6], [-1.0, 2.0] if shard == "codeblock" else [-2.0, -1.0],
                              out if rank == 0 else None, backend=OracleBackend(), verbose=False, shard=shard)
if rank == 0:
    print("RESULT", res[2])
if world > 1:
    dist.destroy_process_group()
"""


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_sharded_counters_match_single_rank(tmp_path):
    """world_size 1 vs 2 (gloo, CPU): identical BLER tables, i.e. identical summed counters and the
    same stopping decisions; the pickle has the reference's layout."""
    tables = {}
    for shard in ("codeblock", "point"):   # every W-th codeblock of every point / one grid point per rank
        outs = []
        for world in (1, 2):
            port, out = str(_free_port()), str(tmp_path / f"{shard}{world}.pickle")
            procs = [subprocess.Popen([sys.executable, "-c", _WORKER, ROOT, str(r), str(world), port, out, shard],
                                      stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True) for r in range(world)]
            for p in procs:
                so, se = p.communicate(timeout=900)
                assert p.returncode == 0, se[-3000:]
            with open(out, "rb") as f:
                outs.append(pickle.load(f))
        assert all(o == outs[0] for o in outs), shard   # the result does not depend on the number of ranks
        tables[shard] = outs[0]
    assert tables["codeblock"][:2] == tables["point"][:2]   # same labels; the point mode draws other (equivalent) inputs at
    # low SNR, where every point stops at the first checkpoint
    outs = [tables["codeblock"]]
    sim_config, labels, table = outs[0]
    assert sim_config == {'Zc': 2, 'bgn': 2} and labels == ['NMS-alpha=0.8-L=6', 'OMS-beta=0.3-L=6']
    assert len(table) == 2 and all(len(row) == 2 for row in table)
    assert table[0][0] > table[0][1]  # BLER falls with SNR


@pytest.mark.gpu
def test_gpu_driver_matches_oracle_driver(tmp_path):
    """Same NumPy seed: the CUDA backend and the oracle backend must produce the same BLER table."""
    from python_5gtoolbox_b200 import sim
    args = (3, 1, '24A', ['min-sum', 'mixed-MS', 'BF'], [], [], [[0.8, 0.3]], [8], [0.0, 3.0])
    np.random.seed(5)
    a = sim.run_ldpc_simulation(*args, str(tmp_path / "a.pickle"), verbose=False)
    np.random.seed(5)
    b = sim.run_ldpc_simulation(*args, None, backend=OracleBackend(), verbose=False)
    assert a == b
    with open(tmp_path / "a.pickle", "rb") as f:
        assert pickle.load(f) == list(a)


@pytest.mark.gpu
def test_gpu_driver_device_rng_bler_in_ci():
    """Device (Philox) generation: BLER of BG1 Zc=12 mixed-MS(0.8,0.3) L=32 at 0 dB against the
    reference's shipped table value 0.00375 (out/ldpc_decode_result_opt.pickle) and at
    -1 dB against 0.275, within a generous binomial interval."""
    from python_5gtoolbox_b200 import sim
    _, _, table = sim.run_ldpc_simulation(12, 1, '24A', ['mixed-MS'], [], [], [[0.8, 0.3]], [32], [-1.0, 0.0], None,
                                          rng="device", verbose=False)
    assert 0.20 < table[0][0] < 0.36
    assert table[0][1] < 0.012


def test_bler_curve_sharding_single_reduce():
    """bler_curve: contiguous shares cover every codeblock once for any world size, and the counters of a
    sharded run (emulated ranks, summed like the all-reduce does) equal the single-rank ones."""
    from python_5gtoolbox_b200 import sim
    for n in (0, 1, 7, 1000, 10 ** 6 + 3):
        for world in (1, 2, 3, 8):
            spans = [sim.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n and all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1

    def fake(p, snr_db, lo, hi):   # per-codeblock deterministic "results": a function of (point, index) only
        idx = np.arange(lo, hi)
        be = ((idx * 2654435761 + p) % 17 == 0)
        return [hi - lo, int(be.sum()), int((be * (idx % 5 + 1)).sum()), int((idx % 9 + 1).sum())]

    one = sim.bler_curve(12, 1, [0.0, 1.0], 5000, 10, point_counters=fake)
    parts = []
    for r in range(4):
        lo, hi = sim.shard_range(5000, r, 4)
        parts.append([fake(p, s, lo, hi) for p, s in enumerate([0.0, 1.0])])
    summed = np.sum(np.array(parts), axis=0)
    for p in range(2):
        assert [one[p]["codeblocks"], one[p]["block_errors"], one[p]["bit_errors"]] == summed[p][:3].tolist()
        assert one[p]["bler"] == summed[p][1] / 5000 and abs(one[p]["mean_iters"] - summed[p][3] / 5000) < 1e-12


@pytest.mark.gpu
def test_gpu_bler_curve_fixed_count():
    """Device BLER curve, fixed count per point: monotone in SNR, consistent counters, reproducible (same
    seed -> identical counters, whatever the chunking of the work)."""
    from python_5gtoolbox_b200 import sim
    a = sim.bler_curve(28, 1, [-1.0, 0.0, 1.0], 6000, 16, 0.8, 0.3)
    b = sim.bler_curve(28, 1, [-1.0, 0.0, 1.0], 6000, 16, 0.8, 0.3)
    assert a == b
    assert all(r["codeblocks"] == 6000 for r in a)
    assert a[0]["bler"] > a[1]["bler"] >= a[2]["bler"] and a[0]["bler"] > 0.05 and a[2]["bler"] < 0.01
    assert all(r["bit_errors"] >= r["block_errors"] for r in a) and a[0]["mean_iters"] > a[2]["mean_iters"]
    # shipped value: mixed (0.8,0.3) L=32 Zc=28 BG1 at -1 dB: 0.095 (out/mixed_MS_search_pair_ZC28)
    c = sim.bler_curve(28, 1, [-1.0], 4000, 32, 0.8, 0.3)
    assert abs(c[0]["bler"] - 0.095) < 4 * (0.095 * 0.905 / 400 + 0.095 * 0.905 / 4000) ** 0.5 + 0.005


# BLER values of the reference's shipped tables (out/NMS_search_alpha_*, out/OMS_search_beta_*):
# (Zc, bgn, algo, parameter, L, snr_db, shipped BLER).  The reference's trial count per point is not stored
# (granularity suggests 400-2000): the check uses n_ref = 400 for its binomial interval.
SHIPPED_BLER = [
    (12, 1, 'NMS', 0.5, 32, -0.5, 0.19), (12, 1, 'NMS', 0.7, 32, -0.5, 0.085), (12, 1, 'NMS', 0.9, 32, -0.5, 0.405),
    (72, 1, 'NMS', 0.5, 32, -0.5, 0.02), (72, 1, 'NMS', 0.9, 32, -0.5, 0.35), (28, 2, 'NMS', 0.3, 32, -0.5, 0.425),
    (208, 1, 'NMS', 0.9, 32, -0.5, 0.145), (384, 1, 'NMS', 0.9, 32, -0.5, 0.115), (384, 1, 'NMS', 0.5, 32, -0.5, 0.05),
    (12, 1, 'OMS', 0.3, 16, -0.5, 0.245), (40, 1, 'OMS', 0.9, 16, -0.5, 0.285), (176, 1, 'OMS', 0.1, 16, -0.5, 0.685),
]


@pytest.mark.gpu
def test_gpu_bler_matches_shipped_tables():
    """BLER-vs-SNR points of the reference's own parameter-search runs, re-run with device-generated inputs
    (the reference's stopping rule: 1000-10000 codeblocks per point), inside a 4-sigma binomial interval."""
    from python_5gtoolbox_b200 import sim
    for Zc, bgn, algo, par, L, snr, ref in SHIPPED_BLER:
        alpha, beta = ([par], []) if algo == 'NMS' else ([], [par])
        _, _, table = sim.run_ldpc_simulation(Zc, bgn, '24A', [algo], alpha, beta, [], [L], [snr], None, rng="device",
                                              verbose=False)
        p = table[0][0]
        n = 10000 if p < 0.0025 else (4000 if p < 0.00625 else (2000 if p < 0.025 else 1000))  # the stopping rule's n
        tol = 4 * (ref * (1 - ref) / 400 + p * (1 - p) / n) ** 0.5 + 0.005
        assert abs(p - ref) <= tol, (Zc, bgn, algo, par, p, ref, tol)


@pytest.mark.gpu
def test_mc_chain_counters_identical_on_two_gpus_nccl():
    """the same seeds sharded over 1 and 2 GPUs (NCCL) give identical summed counters.  Runs
    bench.py --workload mc under torch.distributed.run when two devices are visible."""
    import json
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two CUDA devices")
    lines = []
    for world in (1, 2):
        cmd = [sys.executable, os.path.join(ROOT, "bench.py"), "--workload", "mc", "--steps", "1", "--warmup", "1", "--mc-codeblocks", "30000"]
        if world > 1:
            cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr", "127.0.0.1",
                   "--master-port", str(_free_port())] + cmd[1:]
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
        assert r.returncode == 0, r.stderr[-3000:]
        lines.append(json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1]))
    assert lines[0]["n_gpus"] == 1 and lines[1]["n_gpus"] == 2
    assert lines[0]["config"]["counters"] == lines[1]["config"]["counters"]
    assert all(c["codeblocks"] == 30000 for c in lines[0]["config"]["counters"])
