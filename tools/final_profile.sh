# Round-end profiling pass (run under gpurun, one workload per call: the two full reports
# together exceed gpurun's 64 MiB return limit):  bash tools/final_profile.sh c2|c4
# Every ncu command runs only after the same command exited 0 without ncu.
set -x
W=${1:-c2}
B="python bench.py --workload $W --steps 1 --warmup 3 --no-cpu-baseline --e2e-steps 0"
timeout 300 $B > gpurun_out/plain_$W.log 2>&1 || exit 1
if [ "$W" = c2 ]; then
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01_launches_final.csv $B > gpurun_out/ncu_launches.log 2>&1
fi
# the timed step: 4-5 setup launches + 3 warm-up steps x 23 launches come first
timeout 1200 ncu --set full --clock-control none --import-source on -s 73 -c 25 -o gpurun_out/r01_full_$W $B > gpurun_out/ncu_full_$W.log 2>&1
ls -la gpurun_out/
