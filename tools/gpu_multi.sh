#!/bin/bash
# multi-GPU: headline bench + Monte-Carlo chain at N GPUs (N = $1)
cd "$GRAFT_REPO_ROOT" || exit 1
N=${1:-2}
mkdir -p gpurun_out
if [ "$N" = "2" ]; then
  timeout 600 python -m pytest tests/test_sim_driver.py -m gpu -x -q -k nccl > gpurun_out/r2g_nccl_test.log 2>&1; tail -3 gpurun_out/r2g_nccl_test.log
fi
for n in 1 $N; do
  if [ "$n" = "1" ]; then L="python"; else L="python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511"; fi
  timeout 900 $L bench.py --gpus $n --workload mc --steps 2 --warmup 1 --mc-codeblocks 1000000 > gpurun_out/r2g_mc_${n}gpu.json 2> gpurun_out/r2g_mc_${n}gpu.err; echo "mc $n rc=$?"; cut -c1-400 gpurun_out/r2g_mc_${n}gpu.json
done
L="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512"
for rep in 1 2 3; do
  timeout 900 $L bench.py --gpus $N --steps 10 --warmup 3 --no-extra > gpurun_out/r2g_bench_${N}gpu_run$rep.json 2> gpurun_out/r2g_bench_${N}gpu.err; echo "bench $N run $rep rc=$?"
  python - <<PY
import json
d=json.load(open("gpurun_out/r2g_bench_${N}gpu_run$rep.json"))
print("value",round(d["value"],3),"e2e",round(d["e2e"]["value"],3),"pageable",round(d["e2e"]["pageable_value"],3),"link",d["e2e"]["host_link_ceiling"])
PY
done
tail -3 gpurun_out/r2g_bench_${N}gpu.err
