# Round-end profiling pass (run under gpurun, one workload per call: two full reports together
# exceed gpurun's 64 MiB return limit):  bash tools/final_profile.sh c2|c4 [ROUND_TAG]
# Every ncu command runs only after the same command exited 0 without ncu.  The profiled window is
# ONE step after the warm-up (c4: after 300 simulated steps), bracketed by cudaProfilerStart/Stop.
set -x
W=${1:-c2}
R=${2:-r02}
PRE=0; if [ "$W" = c4 ]; then PRE=300; fi
B="python bench.py --workload $W --presteps $PRE --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 0 --c4-presteps 0 --ncu-window"
timeout 600 $B > gpurun_out/plain_$W.log 2>&1 || exit 1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/${R}_launches_$W.csv $B > gpurun_out/ncu_launches_$W.log 2>&1
timeout 1500 ncu --set full --clock-control none --import-source on --profile-from-start off -o gpurun_out/${R}_full_$W $B > gpurun_out/ncu_full_$W.log 2>&1
ls -la gpurun_out/ | grep "${R}_"
