mkdir -p gpurun_out
T=${1:-r2o}
N=${2:-2}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/${T}_bench_${N}gpu.json 2> gpurun_out/${T}_bench_${N}gpu.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus $N --steps 3 --warmup 1 > gpurun_out/${T}_bench_ref_${N}gpu.json 2> gpurun_out/${T}_bench_ref_${N}gpu.err
g++ -std=c++14 -O2 -I include -I /usr/local/cuda/include tests/cpp/multi_gpu_host.cpp -o /tmp/multi_gpu_host -L mercury_b200 -lmercury_b200 -L/usr/local/cuda/lib64 -lcudart -lnccl -pthread -Wl,-rpath,$PWD/mercury_b200
/tmp/multi_gpu_host mercury_b200/data/ldpc_tables.bin $N 65536 3 8 i16 > gpurun_out/${T}_cpp_host_${N}gpu_i16.json 2> gpurun_out/${T}_cpp_host_${N}gpu.err
/tmp/multi_gpu_host mercury_b200/data/ldpc_tables.bin $N 65536 3 8 c64 > gpurun_out/${T}_cpp_host_${N}gpu_c64.json 2>> gpurun_out/${T}_cpp_host_${N}gpu.err
tail -3 gpurun_out/${T}_bench_${N}gpu.err; cat gpurun_out/${T}_cpp_host_${N}gpu_*.json; tail -2 gpurun_out/${T}_cpp_host_${N}gpu.err
python - <<PY
import json
d=json.loads(open("gpurun_out/${T}_bench_${N}gpu.json").read().strip().splitlines()[-1])
print("value", d["value"], "n_gpus", d["n_gpus"], d["config"]["workload"])
print("integrity", d["integrity"])
e=d["e2e"]; print("e2e", e["value"], "c64", e["complex64"]["value"], "ceiling", e["h2d_ceiling"])
r=json.loads(open("gpurun_out/${T}_bench_ref_${N}gpu.json").read().strip().splitlines()[-1]); print("ref", r["value"], r["config"]==d["config"])
PY
