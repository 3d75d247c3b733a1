mkdir -p gpurun_out
T=${1:-r3u}
run() { name=$1; shift; env "$@" timeout 300 python bench.py --no-e2e --no-extra --cpu-frames 0 --steps 5 $BARGS > gpurun_out/${T}_bench_$name.json 2> gpurun_out/${T}_bench_$name.err; }
for cfg in 8 0; do BARGS="--config $cfg"
run m${cfg}_g8 MERCURY_B200_SO=$PWD/tuning/libmb_g8.so
for vc in 1200,0,200 1000,20,200 1500,0,200 1200,0,300 800,40,200; do run m${cfg}_m8_vc$vc MERCURY_B200_SO=$PWD/tuning/libmb_m8.so MERCURY_B200_LDPC_VCOST=$vc; done
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
MERCURY_B200_LDPC_VCOST=1200,0,200 bash tools/r2_timing.sh ${T} 8
