#!/usr/bin/env python3
"""BASELINE config #3 on several GPUs: the NMS alpha search of scripts/NMS_ldpc_search_best_alpha.py (its grid: alpha in
{0.1, 0.3, 0.5, 0.7, 0.9}, L = 32, -0.5 dB, the 1000/2000/4000/10000 stopping rule) through sim.run_ldpc_simulation, one
(alpha, SNR) grid point per rank (shard="point"), NCCL all-reduce of the counters at the end.

    python tools/param_search_multi.py OUT.json [Zc ...]                                   (1 GPU)
    python -m torch.distributed.run --nproc-per-node N ... tools/param_search_multi.py OUT.json [Zc ...]

Every rank count must produce the same tables (the streams are seeded per grid point): compare the JSON files."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
from python_5gtoolbox_b200 import sim  # noqa: E402

out = sys.argv[1]
zcs = [int(z) for z in sys.argv[2:]] or [28, 72, 384]
world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
res = {}
for rng in ("numpy", "device"):
    for Zc in zcs:
        if rng == "numpy" and Zc > 100:
            continue   # the host RNG paces this mode: 10^4 x 25 344 normals per point
        np.random.seed(1234)
        torch.cuda.synchronize()
        t0 = time.time()
        cfg, labels, table = sim.run_ldpc_simulation(Zc, 1, '24A', ['NMS'], [0.1, 0.3, 0.5, 0.7, 0.9], [], [], [32], [-0.5], None,
                                                     rng=rng, shard="point", verbose=False)
        torch.cuda.synchronize()
        res[f"{rng} Zc={Zc}"] = {"labels": labels, "bler": table, "seconds": round(time.time() - t0, 3)}
if rank == 0:
    print(json.dumps({k: (v["bler"], v["seconds"]) for k, v in res.items()}))
    with open(out, "w") as f:
        json.dump({"world": world, "results": res}, f, indent=1)
if world > 1:
    dist.destroy_process_group()
