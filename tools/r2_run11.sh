mkdir -p gpurun_out
T=${1:-r2k}
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_dropin.py -q 2>&1 | tail -30 > gpurun_out/${T}_pytest.log
( time timeout 900 python bench.py > gpurun_out/${T}_bench_full.json 2> gpurun_out/${T}_bench_full.err ) 2> gpurun_out/${T}_bench_full.time
( time timeout 600 python bench.py --impl reference > gpurun_out/${T}_bench_ref.json 2> gpurun_out/${T}_bench_ref.err ) 2> gpurun_out/${T}_bench_ref.time
tail -15 gpurun_out/${T}_pytest.log
tail -3 gpurun_out/${T}_bench_full.err; cat gpurun_out/${T}_bench_full.time gpurun_out/${T}_bench_ref.time
python - <<PY
import json
d=json.load(open("gpurun_out/${T}_bench_full.json"))
print("value", d["value"], "ms/step", d["ms_per_step"])
print("e2e", json.dumps(d["e2e"])[:1500])
print("ldpc", json.dumps(d["ldpc"])[:900])
print("other", json.dumps(d["other_configs"])[:1500])
print("cpu", d["cpu_baseline"])
r=json.load(open("gpurun_out/${T}_bench_ref.json")); print("ref", r["value"], r["cpu_baseline"])
PY
