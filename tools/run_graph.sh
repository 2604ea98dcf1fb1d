timeout 800 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
for tag in graph nograph; do
  if [ $tag = nograph ]; then export GNX_NO_GRAPH=1; else unset GNX_NO_GRAPH; fi
  timeout 200 python bench.py --steps 200 --warmup 5 --no-cpu-baseline --e2e-steps 0 > gpurun_out/b_c2_$tag.json 2> gpurun_out/b_c2_$tag.err
  timeout 200 python bench.py --workload c3 --steps 200 --warmup 5 --no-cpu-baseline --e2e-steps 0 > gpurun_out/b_c3_$tag.json 2> gpurun_out/b_c3_$tag.err
  timeout 200 python bench.py --workload c3 --replicates 1 --steps 200 --warmup 5 --no-cpu-baseline --e2e-steps 0 > gpurun_out/b_c3r1_$tag.json 2> gpurun_out/b_c3r1_$tag.err
done
python - <<PY
import json
for t in ['graph','nograph']:
    for w in ['c2','c3','c3r1']:
        try:
            d=json.load(open('gpurun_out/b_%s_%s.json'%(w,t)))
            print(t, w, 'ms/step %.4f'%d['ms_per_step'], 'value %.4g'%d['value'], d['config']['replicates'], d['clocks']['samples'] if d.get('clocks') else None)
        except Exception as e:
            print(t, w, 'FAILED', e, open('gpurun_out/b_%s_%s.err'%(w,t)).read()[-400:])
PY
