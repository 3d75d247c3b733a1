mkdir -p gpurun_out
T=${1:-r4g}
for v in ms16 ms8 ms4; do
for cfg in 100 102; do
MERCURY_B200_SO=$PWD/tuning/libmb_$v.so timeout 300 python tools/bench_mfsk.py --config $cfg --cpu-frames 0 > gpurun_out/${T}_mfsk_${cfg}_$v.json 2> gpurun_out/${T}_mfsk_${cfg}_$v.err
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${T}_mfsk_${cfg}_$v.json").read().strip().splitlines()[-1])
    print("$v", $cfg, {k:d[k] for k in d if k in ("value","ms_per_step")}, d.get("roofline",{}).get("frac"), d.get("roofline",{}).get("kernel_ms"), d.get("integrity"))
except Exception as e: print("$v", $cfg, "failed", e)
PY
done; done
