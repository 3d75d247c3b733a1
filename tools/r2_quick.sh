mkdir -p gpurun_out
T=${1:-r2q}
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -4 > gpurun_out/${T}_pytest.log
run() { name=$1; shift; env "$@" timeout 300 python bench.py --no-e2e --no-extra --cpu-frames 0 --steps 5 $BARGS > gpurun_out/${T}_bench_$name.json 2> gpurun_out/${T}_bench_$name.err; }
for cfg in 8 0 3 9 12 13; do BARGS="--config $cfg"; run m${cfg} X=1; done
BARGS="--config 16 --iters 20 --esn0 18"; run m16_18dB X=1
BARGS="--config 16 --iters 20 --esn0 30"; run m16_30dB X=1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:mb_ldpc -s 1 -c 1 -o gpurun_out/${T}_prof_ldpc \
    python bench.py --batch 16384 --steps 1 --warmup 1 --no-e2e --no-extra --cpu-frames 0 > gpurun_out/${T}_prof_ldpc.log 2>&1
cat gpurun_out/${T}_pytest.log | tail -3
python - <<PY
import json,glob
for f in sorted(glob.glob("gpurun_out/${T}_bench_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1]); r, l = d["roofline"], d["ldpc"]
        print(f, f"value {d['value']:.4g} | demod {r['kernel_ms']:.3f} ms frac {r['frac']:.3f} | ldpc {l['kernel_ms']:.3f} ms it {l['mean_iterations']:.2f} | mism {d['integrity']['payload_mismatches_among_decoded']} fer {d['integrity']['fer']:.4f}")
    except Exception as e:
        print(f, "failed", e)
PY
