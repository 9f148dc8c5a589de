cd "$GRAFT_REPO_ROOT" || exit 1
N=${1:-8}
mkdir -p gpurun_out
L="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $L bench.py --gpus $N --workload mc --steps 2 --warmup 1 --mc-codeblocks 1000000 > gpurun_out/r2g_mc_${N}gpu.json 2> gpurun_out/r2g_mc_${N}gpu.err; echo "mc $N rc=$?"; cut -c1-200 gpurun_out/r2g_mc_${N}gpu.json
bash tools/gpu_driver_like.sh $N
