mkdir -p gpurun_out
T=${1:-r2u}
run() { name=$1; shift; env "$@" timeout 300 python bench.py --no-e2e --no-extra --cpu-frames 0 --steps 5 $BARGS > gpurun_out/${T}_bench_$name.json 2> gpurun_out/${T}_bench_$name.err; }
for cfg in 8 0 9; do BARGS="--config $cfg"
run m${cfg}_w8 X=1
run m${cfg}_w7 MERCURY_B200_SO=$PWD/tuning/libmb_w7.so
run m${cfg}_w6 MERCURY_B200_SO=$PWD/tuning/libmb_w6.so
run m${cfg}_w4 MERCURY_B200_SO=$PWD/tuning/libmb_w4.so
done
python - <<PY
import json,glob
for f in sorted(glob.glob("gpurun_out/${T}_bench_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1]); r, l = d["roofline"], d["ldpc"]
        print(f, f"ldpc {l['kernel_ms']:.3f} ms it {l['mean_iterations']:.2f} | mism {d['integrity']['payload_mismatches_among_decoded']}")
    except Exception as e:
        print(f, "failed", e)
PY
