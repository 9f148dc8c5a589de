cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "graded or host" > gpurun_out/g2_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/g2_pytest.log
timeout 900 python bench.py --steps 10 --warmup 3 --no-extra > gpurun_out/g2_bench.json 2> gpurun_out/g2_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open("gpurun_out/g2_bench.json")); e=d["e2e"]
print("value",round(d["value"],3),"e2e",round(e["value"],3),"pageable",round(e["pageable_value"],3),"f16",round(e["half_precision_llr_input"]["value"],3),"ceiling",round(e["host_link_ceiling"]["e2e_gbit_per_s_at_that_rate"],3))
PY
