#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
N=${1:-2}
python tools/param_search_multi.py gpurun_out/param_search_1gpu.json 2>&1 | tail -2
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29544 tools/param_search_multi.py gpurun_out/param_search_${N}gpu.json 2>&1 | tail -2
python - <<PY
import json
a=json.load(open("gpurun_out/param_search_1gpu.json"))["results"]; b=json.load(open("gpurun_out/param_search_${N}gpu.json"))["results"]
for k in a:
    print(k, "identical tables:", a[k]["bler"] == b[k]["bler"], "seconds 1 GPU", a[k]["seconds"], "-> $N GPUs", b[k]["seconds"], a[k]["bler"])
PY
