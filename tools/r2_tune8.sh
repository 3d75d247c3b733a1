mkdir -p gpurun_out
T=${1:-r3v}
NEW=${2:-z8}
MERCURY_B200_SO=$PWD/tuning/libmb_$NEW.so timeout 900 python -m pytest tests/test_gpu_parity.py -x -q > gpurun_out/${T}_pytest.log 2>&1; tail -3 gpurun_out/${T}_pytest.log
run() { name=$1; shift; env "$@" timeout 300 python bench.py --no-e2e --no-extra --cpu-frames 0 --steps 5 $BARGS > gpurun_out/${T}_bench_$name.json 2> gpurun_out/${T}_bench_$name.err; }
for v in g8 $NEW; do
BARGS="--config 8"; run m8_$v MERCURY_B200_SO=$PWD/tuning/libmb_$v.so
BARGS="--config 0"; run m0_$v MERCURY_B200_SO=$PWD/tuning/libmb_$v.so
BARGS="--config 16 --iters 20 --esn0 30"; run m16_30dB_$v MERCURY_B200_SO=$PWD/tuning/libmb_$v.so
BARGS="--config 16 --iters 20 --esn0 25"; run m16_25dB_$v MERCURY_B200_SO=$PWD/tuning/libmb_$v.so
BARGS="--config 15"; run m15_$v MERCURY_B200_SO=$PWD/tuning/libmb_$v.so
done
python - <<PY
import json,glob
for f in sorted(glob.glob("gpurun_out/${T}_bench_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1]); r, l = d["roofline"], d["ldpc"]
        print(f, f"ldpc {l['kernel_ms']:.3f} ms it {l['mean_iterations']:.2f} | mism {d['integrity']['payload_mismatches_among_decoded']} fer {d['integrity']['fer']:.4f}")
    except Exception as e:
        print(f, "failed", e)
PY
