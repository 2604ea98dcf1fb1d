# usage: run_fm.sh TAG [extra pytest -k] -- mate-search parity tests + c4 (t=300, t=1000) + c2 (t=600) device-only benches
TAG=$1
timeout 900 python -m pytest tests/test_cuda_mates.py tests/test_cuda_multistep.py tests/test_cuda_parity.py tests/test_cuda_fullsize.py -m gpu -x -q 2>&1 | tail -3
for T in 300 1000; do
timeout 300 python bench.py --workload c4 --presteps $T --steps 60 --warmup 5 --no-cpu-baseline --e2e-steps 0 > gpurun_out/${TAG}_c4_$T.json 2> gpurun_out/${TAG}_c4_$T.err
done
timeout 200 python bench.py --presteps 600 --steps 200 --warmup 5 --no-cpu-baseline --e2e-steps 0 --c4-presteps 0 > gpurun_out/${TAG}_c2.json 2> gpurun_out/${TAG}_c2.err
python - <<PY
import json
for f in ['gpurun_out/${TAG}_c4_300.json','gpurun_out/${TAG}_c4_1000.json','gpurun_out/${TAG}_c2.json']:
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f, 'failed', e); continue
    print(f, round(d['ms_per_step'],4), '%.4g'%d['value'])
    print('   '+'  '.join('%s=%.0f'%(k['kernel'].replace('scan_','s_'),k['ms_per_launch']*1e3) for k in d['kernels']))
PY
