mkdir -p gpurun_out
T=${1:-r3s}
timeout 900 python bench.py > gpurun_out/${T}_bench_1gpu.json 2> gpurun_out/${T}_bench_1gpu.err
python - <<PY
import json
d=json.loads(open("gpurun_out/${T}_bench_1gpu.json").read().strip().splitlines()[-1])
print("value", d["value"], "ms/step", d["ms_per_step"], "e2e", d["e2e"]["value"])
print("ldpc", d["ldpc"]["kernel_ms"], d["ldpc"]["kernel_ms_per_step"], d["ldpc"]["roofline"]["frac"], "demod", d["roofline"]["kernel_ms"], d["roofline"]["frac"])
print(d["clocks"])
PY
