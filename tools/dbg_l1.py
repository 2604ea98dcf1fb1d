import sys, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), 'tests'))
import numpy as np
from parity_util import synthetic_case, run_device_step
from oracle import step_oracle as so
for L in (1, 64):
    arch, prm, state, draws = synthetic_case(L=L, n=1200, loci_per_trait=min(20, max(1, L // 3)), n_traits=2 if L >= 4 else 1, seed=L)
    new_o, im_o = so.step(state, arch, prm, draws)
    out = run_device_step(arch, prm, state, draws, staged=True)
    alive_o = ~so.mortality(im_o['death_p'], draws['death_u'][:len(im_o['death_p'])])
    a = out['alive'].astype(bool)
    bad = np.nonzero(a != alive_o)[0]
    print('L', L, 'n_pre', len(a), 'mismatch', bad[:10], 'p dev', out['death_p'][bad[:5]], 'p orc', im_o['death_p'][bad[:5]], 'u', draws['death_u'][bad[:5]])
    print('  len new', len(out['new']['idx']), len(new_o['idx']), 'idx equal', np.array_equal(out['new']['idx'], new_o['idx']))
    if not np.array_equal(out['new']['idx'], new_o['idx']):
        k = np.nonzero(out['new']['idx'][:min(len(out['new']['idx']), len(new_o['idx']))] != new_o['idx'][:min(len(out['new']['idx']), len(new_o['idx']))])[0]
        print('  first diff pos', k[:5], out['new']['idx'][k[:5]], new_o['idx'][k[:5]])
        print('  gslot etc: counters', out['counters'])
