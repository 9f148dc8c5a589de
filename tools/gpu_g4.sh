cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
NCU="ncu --clock-control none"
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/g4_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/g4_pytest.log
python tools/profile_decode.py 592 2 > gpurun_out/r2n_decode_plain.log 2>&1 || exit 1
python bench.py --steps 2 --warmup 1 --no-extra --no-cpu --batch 16384 > gpurun_out/r2n_bench_plain.json 2>/dev/null || exit 1
$NCU --metrics gpu__time_duration.sum -c 400 --csv --log-file gpurun_out/r2_launches_bench.csv python bench.py --steps 2 --warmup 1 --no-extra --no-cpu --batch 16384 > gpurun_out/r2n_bench_ncu.log 2>&1
$NCU --set full --import-source on -k regex:decode_spec_kernel -s 1 -c 1 -o gpurun_out/prof_r2_decode -f python tools/profile_decode.py 592 2 > gpurun_out/r2n_decode_ncu.log 2>&1
ls -la gpurun_out/*.ncu-rep
timeout 600 python tools/bench_zc_sweep.py 1:384 1:352 1:320 1:288 1:256 1:208 2:384 2:288 > gpurun_out/g4_sweep_fixed.log 2>&1; cat gpurun_out/g4_sweep_fixed.log | cut -c1-120
timeout 600 python tools/bench_zc_sweep.py --et 1:384 1:256 1:208 1:176 2:384 > gpurun_out/g4_sweep_et.log 2>&1; cat gpurun_out/g4_sweep_et.log | cut -c1-120
timeout 600 python tools/iter_slope.py 32768 | tail -1
