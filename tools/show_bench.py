"""Print a bench.py JSON line as a readable table (diagnostic)."""
import json, sys
def show(d, tag=''):
    print(tag, 'value %.4g  ms/step %.4f  beta %.4f' % (d['value'], d['ms_per_step'], d.get('births_per_individual', d.get('config', {}).get('births_per_individual', 0))))
    r = d.get('roofline') or {}
    print('  roofline', r.get('kernel'), 'frac %.3f' % r.get('frac', 0), ' gsk', {k: (round(v, 4) if isinstance(v, float) else v) for k, v in (r.get('genotype_streaming_kernel') or {}).items()})
    if 'whole_step' in d: print('  whole_step', d['whole_step'])
    for k in d['kernels']:
        print('   %-22s %8.1f us share %.3f frac %s dram/alg %s' % (k['kernel'], k['ms_per_launch'] * 1e3, k.get('share_of_kernel_time_sum', k.get('share_of_step', 0)),
              '%.3f' % k['frac_of_hbm_peak'] if 'frac_of_hbm_peak' in k else '  -  ',
              '%.2f' % (k['dram_bytes_ncu'] / k['algorithmic_bytes']) if 'dram_bytes_ncu' in k and 'algorithmic_bytes' in k else '-'))
for f in sys.argv[1:]:
    d = json.loads(open(f).read().strip().splitlines()[-1])
    show(d, f)
    if d.get('e2e'): print('  e2e', d['e2e']['value'], ' launches', d.get('gpu_launches'), ' clocks', d.get('clocks'))
    if d.get('c4'): show(d['c4'], '  [c4 block]')
