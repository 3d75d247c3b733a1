mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 --no-extra > gpurun_out/n2_stdout.json 2> gpurun_out/n2_stderr.log
wc -l gpurun_out/n2_stdout.json; head -c 120 gpurun_out/n2_stdout.json; echo; grep -c "NCCL version" gpurun_out/n2_stderr.log
