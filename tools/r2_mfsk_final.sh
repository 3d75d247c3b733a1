mkdir -p gpurun_out
T=${1:-r4h}
timeout 900 python -m pytest -m gpu -q tests/test_gpu_mfsk.py tests/test_gpu_dropin.py 2>&1 | tail -3
for cfg in 100 102; do timeout 600 python tools/bench_mfsk.py --config $cfg > gpurun_out/${T}_bench_mfsk_$cfg.json 2> gpurun_out/${T}_bench_mfsk_$cfg.err; tail -c 600 gpurun_out/${T}_bench_mfsk_$cfg.json; echo; done
