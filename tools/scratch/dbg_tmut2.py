import sys, numpy as np
sys.path.insert(0, 'tests'); sys.path.insert(0, '.')
from golden_io import load_case
from parity_util import run_device_step
from oracle import step_oracle as so
from oracle import draws as od
z, arch, prm, state, draws = load_case('tmut')
new_o, im_o = so.step(state, arch, prm, draws)
arch2 = dict(arch, mutation=im_o['mutation'], traits=im_o['mutation']['traits'])
rng = np.random.default_rng(77)
n1 = len(new_o['x'])
d2 = od.make_draws(rng, dict(prm, move_distr=('wald', 1.0, 1.0), disp_distr=('wald', 0.8, 1.0)), n1, 2 * n1 + 64,
                   len(arch['paths']), max_tries=24)
nm = 4
d2.update(mut_n=np.array([nm], np.int32), mut_type_u=rng.random(nm),
          mut_ind_R=rng.integers(0, 2**32, nm, dtype=np.uint64).astype(np.uint32),
          mut_homol_u=rng.random(nm), mut_s=rng.gamma(0.2, 0.2, nm), mut_alpha=rng.normal(0, 0.15, nm))
state2 = dict(new_o)
new2_o, im2_o = so.step(state2, arch2, prm, d2)
for staged in (True, False):
    out2 = run_device_step(arch2, prm, state2, d2, staged=staged)
    print('staged', staged, 'idx equal', np.array_equal(out2['new']['idx'], new2_o['idx']))
    if staged:
        pre = im2_o['pre']
        print('B', out2['B'], im2_o['B'], 'pairs eq', np.array_equal(out2['pairs'], im2_o['pairs']))
        dz = np.abs(out2['pre']['z'] - pre['z']).max(axis=1)
        print('z diff at', np.nonzero(dz > 1e-12)[0][:20])
        df = np.abs(out2['fit_all'] - im2_o['fit_all'])
        print('fit diff at', np.nonzero(df > 1e-9)[0][:20])
        dp = np.abs(out2['death_p'] - im2_o['death_p'])
        print('death_p diff at', np.nonzero(dp > 1e-9)[0][:20])
        print('log dev', out2['mut_log']); print('log ora', im2_o['mut_log'])
        print('n0', n1)
    g_o = new2_o['g']; g_d = out2['new']['g']
    if g_o.shape == g_d.shape:
        bad = np.nonzero((g_o != g_d).any(axis=(1, 2)))[0]
        print('g rows differ for', bad[:10], 'cols', [np.nonzero((g_o[b] != g_d[b]).any(axis=1))[0] for b in bad[:5]])
    else:
        print('g shapes', g_o.shape, g_d.shape)
