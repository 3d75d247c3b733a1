mkdir -p gpurun_out
T=${1:-r3w}
timeout 1500 python -m pytest tests -m gpu -q ${2:-} 2>&1 | tail -8 > gpurun_out/${T}_pytest_gpu.log; cat gpurun_out/${T}_pytest_gpu.log
