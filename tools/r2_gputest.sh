mkdir -p gpurun_out
T=${1:-r3w}
shift
timeout 1500 python -m pytest -m gpu -q "$@" 2>&1 | tail -25 > gpurun_out/${T}_pytest_gpu.log; cat gpurun_out/${T}_pytest_gpu.log
