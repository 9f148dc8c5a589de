cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "packed or encode" > gpurun_out/g1_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/g1_pytest.log
timeout 300 python tools/mc_stage_times.py > gpurun_out/g1_mc_stage.log 2>&1; tail -4 gpurun_out/g1_mc_stage.log
for z in "1 384" "2 384" "1 128" "1 32"; do timeout 120 python tools/profile_encode_packed.py 262144 3 $z 2>&1 | tail -1; done
timeout 120 python tools/profile_encode.py 65536 3 1 384 2>&1 | tail -1
timeout 120 python tools/profile_encode.py 65536 3 1 208 2>&1 | tail -1
timeout 300 ncu --clock-control none --set full --import-source on --kernel-name regex:encode_words_kernel --launch-skip 1 --launch-count 1 -o gpurun_out/prof_r2_encode_packed -f python tools/profile_encode_packed.py 262144 2 > gpurun_out/g1_ncu.log 2>&1; tail -2 gpurun_out/g1_ncu.log
