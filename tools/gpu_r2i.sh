#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/r2i_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2i_pytest.log
tail -5 gpurun_out/r2i_pytest.log
REF=$GRAFT_REPO_ROOT/build/ref_tmp
run() {  # module, timeout
  local t0=$(date +%s.%N)
  timeout $2 python -m python_5gtoolbox_b200.run_reference_script $1 --ref $REF --workdir /tmp/refrun_$1 > gpurun_out/r2i_script_$1.log 2>&1
  local rc=$?
  echo "$1: rc=$rc wall=$(echo "$(date +%s.%N) - $t0" | bc) s" | tee -a gpurun_out/r2i_scripts.log
  tail -3 gpurun_out/r2i_script_$1.log | cut -c1-200
}
: > gpurun_out/r2i_scripts.log
run scripts.mixed_MS_ldpc_search_best_pair 900
run scripts.sim_ldpc_decoder_bf 900
run scripts.sim_ldpc_decoder 300
run scripts.NMS_ldpc_search_best_alpha 300
run scripts.OMS_ldpc_search_best_beta 300
run scripts.NR_PUSCH_throughput_example 600
cp /tmp/refrun_scripts.mixed_MS_ldpc_search_best_pair/out/mixed_MS_search_pair_ZC12_bgn1.pickle gpurun_out/r2i_mixed_ZC12_bgn1.pickle 2>/dev/null
python - <<'PY'
import pickle, glob
for f in sorted(glob.glob("/tmp/refrun_scripts.mixed_MS_ldpc_search_best_pair/out/mixed_MS_search_pair_*.pickle")):
    print(f.split("/")[-1], pickle.load(open(f, "rb")))
f = "/tmp/refrun_scripts.sim_ldpc_decoder_bf/out/ldpc_decode_result_BF.pickle"
print("BF", pickle.load(open(f, "rb")))
PY
echo "== lifting-size sweep"; python tools/bench_zc_sweep.py 1:384 1:208 1:176 1:160 1:144 1:128 1:120 1:112 1:104 1:96 1:88 1:80 1:72 1:64 1:60 1:56 1:52 1:48 1:44 1:40 1:36 1:32 1:28 1:12 2:384 2:208 2:128 2:72 2:40 2:32 2:28 2:12 2>&1 | tee gpurun_out/r2i_zc_sweep.log
