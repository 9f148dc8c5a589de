#!/bin/bash
# the driver's own commands at N GPUs: reference arm first, then ours (no extra flags)
cd "$GRAFT_REPO_ROOT" || exit 1
N=${1:-8}
mkdir -p gpurun_out
L="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29577"
s=$(date +%s); timeout 600 $L bench.py --impl reference --gpus $N --steps 20 --warmup 3 > gpurun_out/driver_ref_${N}gpu.json 2>/dev/null; echo "ref rc=$? $(( $(date +%s) - s )) s"; cut -c1-160 gpurun_out/driver_ref_${N}gpu.json
s=$(date +%s); timeout 900 $L bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/driver_bench_${N}gpu.json 2> gpurun_out/driver_bench_${N}gpu.err; echo "bench rc=$? $(( $(date +%s) - s )) s"
python - <<PY
import json
d=json.load(open("gpurun_out/driver_bench_${N}gpu.json")); e=d["e2e"]
print("n_gpus",d["n_gpus"],"value",round(d["value"],2),"e2e",round(e["value"],2),"pageable",round(e["pageable_value"],2),"f16",round(e["half_precision_llr_input"]["value"],2),"link",round(e["host_link_ceiling"]["h2d_gb_per_s_all_ranks"],1),d["clocks"])
print("TB",d["config"]["transport_block"]["DLSCHDecode_ms"], "agreement", [a["agreement"] for a in d["config"]["fp32_vs_float64_reference_agreement"]])
PY
tail -2 gpurun_out/driver_bench_${N}gpu.err
