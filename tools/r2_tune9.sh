mkdir -p gpurun_out
T=${1:-r4a}
run() { name=$1; shift; env "$@" timeout 300 python bench.py --no-e2e --no-extra --cpu-frames 0 --steps 5 $BARGS > gpurun_out/${T}_bench_$name.json 2> gpurun_out/${T}_bench_$name.err; }
for cfg in 0 1 2 3; do BARGS="--config $cfg"
for v in g8 c4; do run m${cfg}_$v MERCURY_B200_SO=$PWD/tuning/libmb_$v.so; done
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
