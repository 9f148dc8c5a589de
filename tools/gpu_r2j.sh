#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/r2j_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2j_pytest.log
tail -5 gpurun_out/r2j_pytest.log
REF=$GRAFT_REPO_ROOT/build/ref_tmp
: > gpurun_out/r2j_scripts.log
for m in scripts.mixed_MS_ldpc_search_best_pair scripts.sim_ldpc_decoder_bf; do
  s=$(date +%s%N)
  timeout 1200 python -m python_5gtoolbox_b200.run_reference_script $m --ref $REF --workdir /tmp/refrun_$m > gpurun_out/r2j_script_$m.log 2>&1; rc=$?
  e=$(date +%s%N)
  echo "$m: rc=$rc wall=$(( (e - s) / 1000000 )) ms" | tee -a gpurun_out/r2j_scripts.log
  grep "finish test" gpurun_out/r2j_script_$m.log | head -40 | cut -c1-170
done
python - <<'PY'
import pickle, glob
for f in sorted(glob.glob("/tmp/refrun_scripts.mixed_MS_ldpc_search_best_pair/out/mixed_MS_search_pair_*.pickle")):
    print(f.split("/")[-1], pickle.load(open(f, "rb")))
PY
echo "== sweep, fixed 10 iterations"; python tools/bench_zc_sweep.py 1:176 1:160 1:144 1:128 2:160 2:144 2>&1 | tee gpurun_out/r2j_zc_fixed.log
echo "== sweep, early termination at -3 dB"; python tools/bench_zc_sweep.py --et 1:384 1:208 1:160 1:144 1:128 1:112 1:96 1:80 1:72 1:64 1:56 1:48 1:40 1:32 1:28 1:12 2:384 2:128 2:72 2:40 2:32 2:28 2>&1 | tee gpurun_out/r2j_zc_et.log
echo "== same, table-driven kernel"; NRLDPC_NO_SPEC=1 python tools/bench_zc_sweep.py --et 1:128 1:112 1:96 1:80 1:72 1:64 1:56 1:48 1:40 1:32 2:128 2:72 2:40 2:32 2>&1 | tee gpurun_out/r2j_zc_et_tab.log
