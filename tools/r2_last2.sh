mkdir -p gpurun_out
T=${1:-r4x}
bash tools/r2_last.sh $T
for cfg in 8 13 15 16; do
timeout 400 ncu --set full --clock-control none --import-source on -k regex:mb_demod -s 1 -c 1 -o gpurun_out/${T}_prof_demod_m$cfg \
    python bench.py --config $cfg --batch 16384 --steps 1 --warmup 1 --no-e2e --no-extra --cpu-frames 0 > gpurun_out/${T}_prof_demod_m$cfg.log 2>&1
done
timeout 400 ncu --set full --clock-control none --import-source on -k regex:mb_ldpc -s 1 -c 1 -o gpurun_out/${T}_prof_ldpc \
    python bench.py --batch 16384 --steps 1 --warmup 1 --no-e2e --no-extra --cpu-frames 0 > gpurun_out/${T}_prof_ldpc.log 2>&1
timeout 1500 python tools/run_baseline_configs.py --out gpurun_out/${T}_baseline_configs.json > gpurun_out/${T}_baseline_configs.log 2> gpurun_out/${T}_baseline_configs.err
tail -3 gpurun_out/${T}_baseline_configs.log | cut -c1-200
