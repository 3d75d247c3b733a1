mkdir -p gpurun_out
T=${1:-r3h}
for v in t8 t8d; do
MERCURY_B200_SO=$PWD/tuning/libmb_$v.so timeout 300 python bench.py --config 8 --no-e2e --no-extra --cpu-frames 0 --steps 1 --warmup 1 > gpurun_out/${T}_timing_$v.log 2>&1
echo $v; grep "^T cta" gpurun_out/${T}_timing_$v.log | tail -16 | sort -k3,3n -k5,5n
done
