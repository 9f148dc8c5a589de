#!/bin/bash
# ncu captures of round 2 (run each workload once WITHOUT ncu first)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
NCU="ncu --clock-control none"
python tools/profile_decode.py 592 2 > gpurun_out/r2n_decode_plain.log 2>&1 || exit 1
python tools/profile_misc.py > gpurun_out/r2n_misc_plain.log 2>&1 || { tail -5 gpurun_out/r2n_misc_plain.log; exit 1; }
python bench.py --steps 2 --warmup 1 --no-extra --no-cpu --batch 16384 > gpurun_out/r2n_bench_plain.json 2>/dev/null || exit 1
# 1. launch list of the bench command
$NCU --metrics gpu__time_duration.sum -c 400 --csv --log-file gpurun_out/r2_launches_bench.csv python bench.py --steps 2 --warmup 1 --no-extra --no-cpu --batch 16384 > gpurun_out/r2n_bench_ncu.log 2>&1
# 2. the headline kernel, full set
$NCU --set full --import-source on -k regex:decode_spec_kernel -s 1 -c 1 -o gpurun_out/prof_r2_decode -f python tools/profile_decode.py 592 2 > gpurun_out/r2n_decode_ncu.log 2>&1
# 3. the kernels either side of the path: launch list, then full set of one launch of each
$NCU --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum -c 200 --csv --log-file gpurun_out/r2_launches_misc.csv python tools/profile_misc.py > gpurun_out/r2n_misc_list.log 2>&1
$NCU --set full --import-source on --kernel-id '::regex:.*:2' -o gpurun_out/prof_r2_misc -f python tools/profile_misc.py > gpurun_out/r2n_misc_ncu.log 2>&1
ls -la gpurun_out/*.ncu-rep
tail -3 gpurun_out/r2n_misc_ncu.log
