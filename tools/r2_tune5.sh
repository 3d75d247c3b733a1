mkdir -p gpurun_out
T=${1:-r3m}
MERCURY_B200_SO=$PWD/tuning/libmb_h8.so timeout 900 python -m pytest tests/test_gpu_parity.py -x -q > gpurun_out/${T}_pytest.log 2>&1; tail -3 gpurun_out/${T}_pytest.log
run() { name=$1; shift; env "$@" timeout 300 python bench.py --no-e2e --no-extra --cpu-frames 0 --steps 5 $BARGS > gpurun_out/${T}_bench_$name.json 2> gpurun_out/${T}_bench_$name.err; }
for cfg in 8 0 12 3; do BARGS="--config $cfg"
for v in g8 h8; do run m${cfg}_$v MERCURY_B200_SO=$PWD/tuning/libmb_$v.so; done
done
BARGS="--config 8"
for vc in 300,25,180 900,25,180 600,25,100 600,25,260 1200,25,180; do run m8_h8_vc$vc MERCURY_B200_SO=$PWD/tuning/libmb_h8.so MERCURY_B200_LDPC_VCOST=$vc; done
python - <<PY
import json,glob
for f in sorted(glob.glob("gpurun_out/${T}_bench_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1]); r, l = d["roofline"], d["ldpc"]
        print(f, f"ldpc {l['kernel_ms']:.3f} ms it {l['mean_iterations']:.2f} | mism {d['integrity']['payload_mismatches_among_decoded']}")
    except Exception as e:
        print(f, "failed", e)
PY
bash tools/r2_timing.sh ${T} 8
