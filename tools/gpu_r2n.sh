#!/bin/bash
# 2-GPU: the driver's own bench command (no --no-extra), both arms; sanitizer attempt; full tests
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
L="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533"
s=$(date +%s); timeout 900 $L bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r2n_bench_2gpu_full.json 2> gpurun_out/r2n_bench_2gpu_full.err; echo "bench N=2 rc=$? in $(( $(date +%s) - s )) s"
cut -c1-300 gpurun_out/r2n_bench_2gpu_full.json
s=$(date +%s); timeout 600 $L bench.py --impl reference --gpus 2 --steps 20 --warmup 3 > gpurun_out/r2n_ref_2gpu.json 2>/dev/null; echo "ref N=2 rc=$? in $(( $(date +%s) - s )) s"; cut -c1-200 gpurun_out/r2n_ref_2gpu.json
s=$(date +%s); timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r2n_bench_1gpu_full.json 2> gpurun_out/r2n_bench_1gpu_full.err; echo "bench N=1 rc=$? in $(( $(date +%s) - s )) s"
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/r2n_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2n_pytest.log
timeout 300 python tools/sanitizer_cases.py > gpurun_out/r2n_sanitizer_plain.log 2>&1; echo "sanitizer cases plain rc=$?"; tail -3 gpurun_out/r2n_sanitizer_plain.log
timeout 900 compute-sanitizer --tool memcheck python tools/sanitizer_cases.py > gpurun_out/r2n_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -5 gpurun_out/r2n_memcheck.log
