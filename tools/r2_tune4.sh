mkdir -p gpurun_out
T=${1:-r3g}
run() { name=$1; shift; env "$@" timeout 300 python bench.py --no-e2e --no-extra --cpu-frames 0 --steps 5 $BARGS > gpurun_out/${T}_bench_$name.json 2> gpurun_out/${T}_bench_$name.err; }
for cfg in 8 13; do BARGS="--config $cfg"
for c in 0 8 16 32 64 256; do run m${cfg}_ct$c MERCURY_B200_CHEAP_TEST=$c; done
done
for cfg in 0 3 5; do BARGS="--config $cfg"
for c in 0 16 64; do run m${cfg}_ct$c MERCURY_B200_CHEAP_TEST=$c; done
done
BARGS="--config 12 --esn0 30"; run m12_30dB X=1
BARGS="--config 8 --esn0 30"; run m8_30dB X=1
python - <<PY
import json,glob
for f in sorted(glob.glob("gpurun_out/${T}_bench_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1]); r, l = d["roofline"], d["ldpc"]
        print(f, f"ldpc {l['kernel_ms']:.3f} ms it {l['mean_iterations']:.2f} | mism {d['integrity']['payload_mismatches_among_decoded']}")
    except Exception as e:
        print(f, "failed", e)
PY
