# usage: run_std.sh TAG  -- gpu tests + C4 + C2 device-only benches
TAG=$1
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 200 python bench.py --workload c4 --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 0 > gpurun_out/bench_c4_$TAG.json 2> gpurun_out/bench_c4_$TAG.err
timeout 200 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --e2e-steps 0 > gpurun_out/bench_c2_$TAG.json 2> gpurun_out/bench_c2_$TAG.err
python - <<PY
import json
for f in ['gpurun_out/bench_c4_$TAG.json','gpurun_out/bench_c2_$TAG.json']:
    d=json.load(open(f))
    print(f, round(d['ms_per_step'],4), '%.4g'%d['value'])
    print('   '+'  '.join('%s=%.0f'%(k['kernel'].replace('scan_','s_'),k['ms_per_launch']*1e3) for k in d['kernels']))
PY
