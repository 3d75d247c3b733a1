mkdir -p gpurun_out
T=${1:-r4d}
VARS=${2:-"base sc12 va vb vc vd"}
run() { name=$1; shift; env "$@" timeout 300 python bench.py --no-e2e --no-extra --cpu-frames 0 --steps 5 $BARGS > gpurun_out/${T}_bench_$name.json 2> gpurun_out/${T}_bench_$name.err; }
for cfg in ${3:-0 8 10 13 15 16}; do BARGS="--config $cfg"
for v in $VARS; do run m${cfg}_$v MERCURY_B200_SO=$PWD/tuning/libmb_$v.so; done
done
python - <<PY
import json,glob
for f in sorted(glob.glob("gpurun_out/${T}_bench_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1]); r, l = d["roofline"], d["ldpc"]
        print(f, f"demod {r['kernel_ms']:.4f} ms frac {r['frac']:.3f} | ldpc {l['kernel_ms']:.3f} | mism {d['integrity']['payload_mismatches_among_decoded']} fer {d['integrity']['fer']:.3f}")
    except Exception as e:
        print(f, "failed", e)
PY
