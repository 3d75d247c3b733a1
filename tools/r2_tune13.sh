mkdir -p gpurun_out
T=${1:-r4j}
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_mfsk.py -x -q 2>&1 | tail -3
run() { name=$1; shift; env "$@" timeout 300 python bench.py --no-e2e --no-extra --cpu-frames 0 --steps 3 $BARGS > gpurun_out/${T}_bench_$name.json 2> gpurun_out/${T}_bench_$name.err; }
BARGS="--config 16 --iters 20 --esn0 18"; run m16_18dB X=1
BARGS="--config 14"; run m14 X=1
BARGS="--config 8"; run m8 X=1
python - <<PY
import json,glob
for f in sorted(glob.glob("gpurun_out/${T}_bench_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1]); r, l = d["roofline"], d["ldpc"]
        print(f, f"ldpc {l['kernel_ms']:.3f} ms it {l['mean_iterations']:.2f} | mism {d['integrity']['payload_mismatches_among_decoded']} fer {d['integrity']['fer']:.3f}")
    except Exception as e:
        print(f, "failed", e)
PY
