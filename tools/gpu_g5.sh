cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/g5_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/g5_pytest.log
timeout 600 python tools/bench_zc_sweep.py --et 1:384 1:256 1:208 1:176 1:144 1:128 1:120 1:112 1:104 1:96 1:88 1:80 1:72 1:64 1:60 1:56 1:52 1:48 1:44 1:40 1:36 1:32 1:28 1:12 2:384 2:208 2:128 2:96 2:72 2:64 2:40 2:32 2:28 2:12 > gpurun_out/g5_sweep_et.log 2>&1
timeout 300 python tools/mc_stage_times.py 2>&1 | tail -2
timeout 300 python tools/pdsch_slot_bench.py 2>&1 | tail -4
