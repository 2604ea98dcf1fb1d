# usage: bash tools/sweep_variants.sh TAG lib1.so lib2.so ...   (run under gpurun; c4 at t=300 and evolved c2)
TAG=$1; shift
for L in default "$@"; do
  if [ "$L" = default ]; then unset GNX_B200_LIB; N=default; else export GNX_B200_LIB=$PWD/$L; N=$(basename $L .so); fi
  timeout 300 python bench.py --workload c4 --presteps 300 --steps 60 --no-cpu-baseline --e2e-steps 0 > gpurun_out/sw_${TAG}_${N}_c4.json 2>/dev/null
  timeout 300 python bench.py --workload c2 --presteps 600 --steps 200 --no-cpu-baseline --e2e-steps 0 --c4-presteps 0 > gpurun_out/sw_${TAG}_${N}_c2.json 2>/dev/null
done
python - <<PY
import json, glob
for f in sorted(glob.glob('gpurun_out/sw_${TAG}_*.json')):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f, 'failed'); continue
    k = {r['kernel']: r['ms_per_launch'] * 1e3 for r in d['kernels']}
    fm = next((r for r in d['kernels'] if r['kernel'] == 'k_find_mates'), {}).get('launches', {})
    print('%-44s ms/step %.4f  find_mates %.1f (light %.1f dense %.1f) move_key %.1f regrid %.1f death %.1f newborns %.1f' % (f.split('/')[-1], d['ms_per_step'], k.get('k_find_mates', 0), fm.get('thread_per_focal_ms', 0) * 1e3, fm.get('crowded_cells_ms', 0) * 1e3, k.get('k_move_key', 0), k.get('k_regrid', 0), k.get('k_death', 0), k.get('k_newborns', 0)))
PY
