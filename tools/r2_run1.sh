mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -30 > gpurun_out/r2a_pytest.log
for s in -1 3 4; do
  MERCURY_B200_LDPC_SET=$s timeout 300 python bench.py --no-e2e --cpu-frames 0 --steps 5 > gpurun_out/r2a_bench_m8_set$s.json 2> gpurun_out/r2a_bench_m8_set$s.err
done
for s in -1 5 4; do
  MERCURY_B200_LDPC_SET=$s timeout 300 python bench.py --config 9 --no-e2e --cpu-frames 0 --steps 5 > gpurun_out/r2a_bench_m9_set$s.json 2> gpurun_out/r2a_bench_m9_set$s.err
done
timeout 400 ncu --set full --clock-control none --import-source on -k regex:mb_ldpc -s 1 -c 1 -o gpurun_out/r2a_prof \
    python bench.py --batch 16384 --steps 1 --warmup 1 --no-e2e --cpu-frames 0 > gpurun_out/r2a_prof.log 2>&1
cat gpurun_out/r2a_pytest.log | tail -15
python - <<PY
import json,glob
for f in sorted(glob.glob("gpurun_out/r2a_bench_*.json")):
    try:
        d = json.load(open(f)); r, l = d["roofline"], d["ldpc"]
        print(f, f"value {d['value']:.4g} | demod {r['kernel_ms']:.3f} ms | ldpc {l['kernel_ms']:.3f} ms it {l['mean_iterations']:.2f} | mism {d['integrity']['payload_mismatches_among_decoded']} fer {d['integrity']['fer']}")
    except Exception as e:
        print(f, "failed", e)
PY
tail -3 gpurun_out/r2a_bench_m8_set-1.err
