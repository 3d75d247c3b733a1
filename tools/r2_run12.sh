mkdir -p gpurun_out
T=${1:-r2l}
( time timeout 1500 python tools/run_baseline_configs.py --out gpurun_out/${T}_baseline_configs.json > gpurun_out/${T}_baseline_configs.log 2> gpurun_out/${T}_baseline_configs.err ) 2> gpurun_out/${T}_baseline_configs.time
tail -60 gpurun_out/${T}_baseline_configs.log | cut -c1-330
tail -5 gpurun_out/${T}_baseline_configs.err; cat gpurun_out/${T}_baseline_configs.time
