mkdir -p gpurun_out
T=${1:-r2j}
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -30 > gpurun_out/${T}_pytest.log
run() { # name, env...
  name=$1; shift
  env "$@" timeout 300 python bench.py --no-e2e --cpu-frames 0 --steps 5 $BARGS > gpurun_out/${T}_bench_$name.json 2> gpurun_out/${T}_bench_$name.err
}
for cfg in 8 0 3 9 13 12; do
BARGS="--config $cfg"
run m${cfg} X=1
done
BARGS="--config 16 --iters 20"
run m16 X=1
BARGS="--config 8 --decoder minsum"
run m8_minsum X=1
cat gpurun_out/${T}_pytest.log | tail -15
python - <<PY
import json,glob
for f in sorted(glob.glob("gpurun_out/${T}_bench_*.json")):
    try:
        d = json.load(open(f)); r, l = d["roofline"], d["ldpc"]
        print(f, f"value {d['value']:.4g} | demod {r['kernel_ms']:.3f} ms | ldpc {l['kernel_ms']:.3f} ms it {l['mean_iterations']:.2f} | mism {d['integrity']['payload_mismatches_among_decoded']} fer {d['integrity']['fer']}")
    except Exception as e:
        print(f, "failed", e)
PY
