mkdir -p gpurun_out
T=${1:-r2z}
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -6 > gpurun_out/${T}_pytest_gpu.log
timeout 900 python bench.py > gpurun_out/${T}_bench_1gpu.json 2> gpurun_out/${T}_bench_1gpu.err
timeout 600 python bench.py --impl reference > gpurun_out/${T}_bench_reference_arm.json 2> gpurun_out/${T}_bench_reference_arm.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${T}_launches.csv python bench.py --steps 2 --warmup 1 --no-extra --cpu-frames 0 > gpurun_out/${T}_launches_bench.log 2>&1
timeout 600 compute-sanitizer --tool memcheck python -c "
import numpy as np, mercury_b200 as mb
ts = mb.TelecomSystemB200(0)
for cfg, it in ((8, 50), (0, 50), (16, 20), (12, 50), (100, 50)):
    g = ts.load_configuration(cfg, it)
    if cfg < 100:
        x, pl = mb.synth_frames(cfg, 37, seed=3, esn0_db=mb.THRESH_DB[cfg] + (0.5 if cfg < 15 else 10.0))
    else:
        x = (np.random.default_rng(1).standard_normal((5, g['Nsymb'], 272, 2)) * 0.1).astype(np.float32).view(np.complex64)[..., 0]
    p, s, _ = ts.demod_decode_batch(x)
    print(cfg, int((s['message_decoded'] == 1).sum()), 'decoded of', len(s))
" > gpurun_out/${T}_compute_sanitizer_memcheck.log 2>&1
timeout 1500 python tools/run_baseline_configs.py --out gpurun_out/${T}_baseline_configs.json > gpurun_out/${T}_baseline_configs.log 2> gpurun_out/${T}_baseline_configs.err
tail -4 gpurun_out/${T}_pytest_gpu.log; tail -4 gpurun_out/${T}_compute_sanitizer_memcheck.log; tail -3 gpurun_out/${T}_bench_1gpu.err
python - <<PY
import json
d=json.loads(open("gpurun_out/${T}_bench_1gpu.json").read().strip().splitlines()[-1])
print("value", d["value"], "e2e", d["e2e"]["value"], "ldpc roofline", d["ldpc"]["roofline"]["frac"], "demod", d["roofline"]["frac"], d["roofline"]["frac_of_dram_bytes"])
PY
grep -c . gpurun_out/${T}_launches.csv; grep "threshold" -A8 gpurun_out/${T}_baseline_configs.log | cut -c1-250
