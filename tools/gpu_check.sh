#!/bin/bash
# usage (under gpurun): tools/gpu_check.sh <tag> [bench args...]   -> gpurun_out/<tag>_{pytest.log,bench.json,prof.ncu-rep}
# GPU parity tests, one bench line (kernel-only, no e2e / cpu legs) and one ncu --set full capture of the demod + ldpc kernels.
tag=$1; shift
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/${tag}_pytest.log
timeout 300 python bench.py --no-e2e --cpu-frames 0 "$@" > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
timeout 400 ncu --set full --clock-control none --import-source on -k regex:mb_ -s 2 -c 2 -o gpurun_out/${tag}_prof \
    python bench.py --batch 16384 --steps 1 --warmup 1 --no-e2e --cpu-frames 0 "$@" > gpurun_out/${tag}_prof.log 2>&1
tail -c 400 gpurun_out/${tag}_pytest.log
python - <<PY
import json
try:
    d = json.load(open("gpurun_out/${tag}_bench.json"))
    r, l = d["roofline"], d["ldpc"]
    print(f"value {d['value']:.4g} frames/s | demod {r['kernel_ms']:.3f} ms frac {r['frac']:.3f} | ldpc {l['kernel_ms']:.3f} ms {l['frames_per_s']:.4g} f/s it {l['mean_iterations']:.2f} | mism {d['integrity']['payload_mismatches_among_decoded']} fer {d['integrity']['fer']}")
except Exception as e:
    print("bench failed:", e)
PY
tail -3 gpurun_out/${tag}_bench.err
