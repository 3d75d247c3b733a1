mkdir -p gpurun_out
T=${1:-r3p}
CFG=${2:-8}
timeout 400 ncu --set full --clock-control none --import-source on -k regex:mb_ldpc -s 1 -c 1 -o gpurun_out/${T}_prof_ldpc \
    python bench.py --config $CFG --batch 16384 --steps 1 --warmup 1 --no-e2e --no-extra --cpu-frames 0 > gpurun_out/${T}_prof_ldpc.log 2>&1
tail -2 gpurun_out/${T}_prof_ldpc.log
