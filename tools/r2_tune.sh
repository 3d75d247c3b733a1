mkdir -p gpurun_out
T=${1:-r2t}
run() { name=$1; shift; env "$@" timeout 300 python bench.py --no-e2e --no-extra --cpu-frames 0 --steps 5 $BARGS > gpurun_out/${T}_bench_$name.json 2> gpurun_out/${T}_bench_$name.err; }
for cfg in 8 0 3; do BARGS="--config $cfg"
run m${cfg}_base X=1
run m${cfg}_lchsmem MERCURY_B200_SO=$PWD/tuning/libmb_lchsmem.so
done
for cfg in 8 9; do BARGS="--config $cfg"
run m${cfg}_costA MERCURY_B200_LDPC_COST=26,25,15,30,40
run m${cfg}_costB MERCURY_B200_LDPC_COST=30,25,25,45,80
run m${cfg}_costC MERCURY_B200_LDPC_COST=20,25,5,37,55
run m${cfg}_costD MERCURY_B200_LDPC_COST=26,22,30,34,70
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
