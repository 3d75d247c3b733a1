mkdir -p gpurun_out
T=${1:-r4z}
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/${T}_smoke.log 2>&1; tail -2 gpurun_out/${T}_smoke.log
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -4 > gpurun_out/${T}_pytest_gpu.log; cat gpurun_out/${T}_pytest_gpu.log
timeout 900 python bench.py > gpurun_out/${T}_bench_1gpu.json 2> gpurun_out/${T}_bench_1gpu.err
timeout 600 python bench.py --impl reference > gpurun_out/${T}_bench_reference_arm.json 2> gpurun_out/${T}_bench_reference_arm.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${T}_launches.csv python bench.py --steps 2 --warmup 1 --no-extra --cpu-frames 0 > gpurun_out/${T}_launches_bench.log 2>&1
python - <<PY
import json
d=json.loads(open("gpurun_out/${T}_bench_1gpu.json").read().strip().splitlines()[-1])
print("value", d["value"], "e2e", d["e2e"]["value"], "ldpc", d["ldpc"]["kernel_ms"], d["ldpc"]["roofline"]["frac"], "demod", d["roofline"]["kernel_ms"], d["roofline"]["frac"], d["roofline"]["frac_of_dram_bytes"])
r=json.loads(open("gpurun_out/${T}_bench_reference_arm.json").read().strip().splitlines()[-1]); print("ref", r["value"], r["config"]==d["config"], r["cpu_baseline"]["cores"])
PY
wc -l gpurun_out/${T}_bench_1gpu.json gpurun_out/${T}_launches.csv
