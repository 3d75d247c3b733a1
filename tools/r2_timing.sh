mkdir -p gpurun_out
T=${1:-r3h}
for cfg in ${2:-8}; do
MERCURY_B200_SO=$PWD/tuning/libmb_t8.so timeout 300 python bench.py --config $cfg --no-e2e --no-extra --cpu-frames 0 --steps 1 --warmup 1 > gpurun_out/${T}_timing_m$cfg.log 2>&1
echo mode $cfg; grep "^T cta" gpurun_out/${T}_timing_m$cfg.log | tail -16 | sort -k3,3n -k5,5n
done
