#!/bin/bash
# what the driver runs at round end: GPU tests, smoke, the bench (both arms)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/final_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/final_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 600 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/final_ref_1gpu.json 2>/dev/null; echo "ref rc=$?"
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/final_bench_1gpu.json 2> gpurun_out/final_bench_1gpu.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open("gpurun_out/final_bench_1gpu.json")); e=d["e2e"]
print("value",round(d["value"],3),"e2e",round(e["value"],3),"pageable",round(e["pageable_value"],3),"f16",round(e["half_precision_llr_input"]["value"],3),d["clocks"])
print("TB",d["config"]["transport_block"]["DLSCHDecode_ms"],d["config"]["transport_block"]["DLSCHEncode_ms"])
for o in d["config"]["other_kernels"]: print(o["kernel"],round(o["ms"],3))
PY
python tools/pdsch_slot_bench.py 2>&1 | tail -4
python tools/bench_encode_sweep.py 1:384 1:208 2:384 1:240 1:384 2>&1 | tail -5
