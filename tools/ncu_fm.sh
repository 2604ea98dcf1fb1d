# usage: bash tools/ncu_fm.sh TAG [lib.so]  -- one --set full capture of both mate-search kernels at c4, t = 300
TAG=$1
if [ -n "$2" ]; then export GNX_B200_LIB=$PWD/$2; fi
B="python bench.py --workload c4 --presteps 300 --steps 1 --warmup 3 --no-cpu-baseline --e2e-steps 0"
timeout 300 $B > gpurun_out/plain_$TAG.log 2>&1 || exit 1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_find_mates --launch-skip 606 --launch-count 2 -o gpurun_out/r02_fm_$TAG $B > gpurun_out/ncu_$TAG.log 2>&1
tail -3 gpurun_out/ncu_$TAG.log
