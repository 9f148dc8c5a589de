cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/g3_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/g3_pytest.log
timeout 600 python tools/bench_zc_sweep.py --et 1:384 1:256 1:208 1:176 1:144 1:128 1:120 1:112 1:104 1:96 1:88 1:80 1:72 1:64 1:60 1:56 1:52 1:48 1:44 1:40 1:36 1:32 1:28 1:12 2:384 2:208 2:128 2:96 2:72 2:64 2:40 2:32 2:28 2:12 > gpurun_out/g3_sweep_et.log 2>&1
timeout 600 python tools/bench_zc_sweep.py 1:384 1:352 1:320 1:288 1:256 1:240 1:224 1:208 1:192 1:176 1:160 1:144 1:128 1:72 1:40 1:28 1:12 1:2 2:384 2:288 2:208 2:144 2:128 2:72 2:28 2:8 > gpurun_out/g3_sweep_fixed.log 2>&1
tail -2 gpurun_out/g3_sweep_fixed.log
timeout 600 python bench.py --steps 20 --warmup 3 --no-extra > gpurun_out/g3_bench.json 2> gpurun_out/g3_bench.err; echo "bench rc=$?"
python -c "
import json
d=json.load(open('gpurun_out/g3_bench.json')); e=d['e2e']
print('value',round(d['value'],3),'e2e',round(e['value'],3),'pageable',round(e['pageable_value'],3))"
