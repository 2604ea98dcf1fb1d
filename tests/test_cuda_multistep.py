"""GPU vs oracle OVER TIME and AT SIZE, with injected draws (bit-exact integer / index work):

  * 24 consecutive time steps, every step fed the same injected draws as the oracle stepping
    beside it, launched through gnx_step so that from the second step on ONE captured CUDA graph
    is re-launched (pointer-stable draw buffers): slot recycling / free list, the double-buffered
    state, newborn ids and the graph re-launch are compared against the oracle's state after
    every step -- unsexed/fixed births, sexed/Poisson births with max_age, and with neutral +
    deleterious mutation accumulating;
  * one step at the full size of BASELINE.json configs[1] (1,048,576 individuals, 100 loci, 2 traits,
    1024x1024) and one at a 200,000-individual configs[3]-shaped case (1000 loci, conductance
    surfaces for movement and dispersal, clumped so that row ranges of > 64 candidates occur).
"""
import numpy as np
import pytest

from parity_util import synthetic_case, make_device

pytestmark = pytest.mark.gpu


def _device_draws(d, arch):
    d = dict(d)
    if arch.get('move_surf') is None:
        d.pop('move_choice', None)
        d.pop('disp_choice', None)
    else:
        d.pop('move_dir', None)
        d.pop('disp_dir', None)
    return d


def _compare_state(dev_state, st, burn=False):
    assert np.array_equal(dev_state['idx'], st['idx'])
    assert np.array_equal(dev_state['age'], st['age'])
    assert np.array_equal(dev_state['sex'], st['sex'])
    np.testing.assert_allclose(dev_state['x'], st['x'], rtol=0, atol=1e-9)
    np.testing.assert_allclose(dev_state['y'], st['y'], rtol=0, atol=1e-9)
    if not burn:
        assert np.array_equal(dev_state['g'], st['g'])
        if st['z'] is not None and st['z'].size:
            np.testing.assert_allclose(dev_state['z'], st['z'], rtol=1e-12, atol=1e-15)
        if st.get('fit') is not None:
            np.testing.assert_allclose(dev_state['fit'], st['fit'], rtol=1e-6)
    assert dev_state['max_ind_idx'] == st['max_ind_idx']


def _walk(arch, prm, state, n_steps, rows, seed, max_tries=24, mut_rate=None, check_every=1):
    from oracle import step_oracle as so
    from oracle import draws as od
    rng = np.random.default_rng(seed)
    dprm = dict(prm, move_distr=('wald', 1.0, 1.0), disp_distr=('wald', 1.0, 1.0))
    dev = make_device(arch, prm, capacity=rows, disp_tries=max_tries)
    recs_o = []
    try:
        dev.upload(state['x'], state['y'], state['age'], state['sex'], state['idx'], g=state['g'], z=state['z'],
                   max_ind_idx=state['max_ind_idx'])
        mut = arch.get('mutation')
        if mut is not None:
            dev.set_mutation(mut['mu_neut'], mut['mu_delet'], mut['mutables'], mut['nonneut_loci'],
                             mut['delet_loci'], mut['delet_s'], mut.get('s_shape', 0.2), mut.get('s_scale', 0.2))
        for t in range(n_steps):
            n = len(state['x'])
            assert 0 < n < rows // 2
            # same array sizes every step: the device re-uses its draw buffers, the graph stays valid
            d = od.make_draws(rng, dprm, rows // 2, rows // 2, len(arch['paths']), max_tries=max_tries)
            if arch.get('move_surf') is not None:
                A = arch['move_surf'].shape[2]
                d['move_choice'] = rng.integers(0, A, rows // 2).astype(np.int32)
                d['disp_choice'] = rng.integers(0, A, (rows // 2, max_tries)).astype(np.int32)
            if mut is not None:
                nm = int(rng.integers(0, 4))
                d.update(mut_n=np.array([nm], np.int32), mut_type_u=rng.random(4),
                         mut_ind_R=rng.integers(0, 2 ** 32, 4, dtype=np.uint64).astype(np.uint32),
                         mut_homol_u=rng.random(4), mut_s=rng.gamma(0.2, 0.2, 4))
            state, im = so.step(state, arch, prm, d)
            if mut is not None:
                arch = dict(arch, mutation=im['mutation'])
            recs_o.append((len(state['x']), im['B'], im['n_deaths']))
            dev.set_draws(_device_draws(d, arch))
            dev.step(1)
            if (t + 1) % check_every == 0 or t == n_steps - 1:
                dev.sync()
                _compare_state(dev.download(genomes=True), state)
        dev.sync()
        recs = dev.step_records()
        assert [(r['Nt'], r['n_births'], r['n_deaths']) for r in recs] == recs_o
        out = dev.download(genomes=True, e=True)
        graph = (dev.graph_launch_count, dev.graph_capture_count)
        mstate = dev.read_mutations()[1] if mut is not None else None
    finally:
        dev.close()
    return out, state, graph, arch, mstate


def test_24_steps_unsexed_fixed_births_graph_relaunch():
    arch, prm, state, _ = synthetic_case(L=200, n=2500, n_traits=2, loci_per_trait=12, dim=(50, 40), seed=3)
    prm = dict(prm, b=0.4)
    out, st, (launches, captures), _, _ = _walk(arch, prm, state, 24, 12000, seed=11, check_every=4)
    _compare_state(out, st)
    assert np.array_equal(out['e'][:, :-1], st['e'])
    # steps 2..24 ran as ONE captured graph re-launched (the first step of a context runs un-captured)
    assert launches == 23 and captures == 1, (launches, captures)
    assert st['max_ind_idx'] > 2500 + 24 * 200                 # thousands of births, ids keep ascending
    assert len(np.unique(out['idx'])) == len(out['idx'])


def test_24_steps_sexed_poisson_maxage():
    arch, prm, state, _ = synthetic_case(L=96, n=2000, n_traits=1, loci_per_trait=9, dim=(45, 45), seed=4)
    prm = dict(prm, sex=True, n_births_fixed=False, lam=2, b=0.9, max_age=6, mating_radius=2.5)
    out, st, (launches, captures), _, _ = _walk(arch, prm, state, 24, 16000, seed=12, check_every=6)
    _compare_state(out, st)
    assert launches == 23 and captures == 1, (launches, captures)
    assert out['age'].max() <= 6


def test_20_steps_with_accumulating_mutations():
    arch, prm, state, _ = synthetic_case(L=300, n=1800, n_traits=1, loci_per_trait=10, dim=(40, 40), seed=6)
    prm = dict(prm, b=0.4)
    tl = np.sort(np.asarray(arch['traits'][0]['loci']))
    mutables = [int(v) for v in np.random.default_rng(2).permutation(np.setdiff1d(np.arange(300), tl))]
    arch['mutation'] = dict(mu_neut=1e-5, mu_delet=2e-5, mutables=mutables, nonneut_loci=tl.astype(np.int64),
                            delet_loci=np.zeros(0, np.int64), delet_s=np.zeros(0), s_shape=0.2, s_scale=0.2)
    out, st, _, arch_end, mstate = _walk(arch, prm, state, 20, 10000, seed=13, check_every=5)
    _compare_state(out, st)
    m = arch_end['mutation']
    assert len(m['delet_loci']) > 3
    assert np.array_equal(mstate['delet_loci'], m['delet_loci'])
    assert np.array_equal(mstate['delet_s'], m['delet_s'])
    assert np.array_equal(mstate['nonneut_loci'], m['nonneut_loci'])
    assert mstate['n_mutables'] == len(m['mutables'])


def _one_big_step(arch, prm, state, draws, cap, max_tries):
    from oracle import step_oracle as so
    new_o, im_o = so.step(state, arch, prm, draws)
    dev = make_device(arch, prm, capacity=cap, disp_tries=max_tries)
    try:
        dev.upload(state['x'], state['y'], state['age'], state['sex'], state['idx'], g=state['g'], z=state['z'],
                   max_ind_idx=state['max_ind_idx'])
        dev.set_draws(_device_draws(draws, arch))
        dev.step(1)                                            # the production (fused) step
        dev.sync()
        out = dev.download(genomes=True, e=True)
        rec = dev.step_records()[-1]
    finally:
        dev.close()
    _compare_state(out, new_o)
    assert np.array_equal(out['e'][:, :-1], new_o['e'])
    assert (rec['Nt'], rec['n_births'], rec['n_deaths']) == (len(new_o['x']), im_o['B'], im_o['n_deaths'])
    return im_o


def test_full_size_c2_step_matches_oracle():
    """BASELINE.json configs[1] at full size: one injected-draw step, every survivor's id, age, sex,
    position, genotype, phenotype and fitness against the oracle (the oracle takes ~10-20 s here)."""
    from geonomics_b200 import workloads, genome_pack
    from oracle import step_oracle as so
    from oracle import draws as od
    cfg = dict(workloads.CONFIGS['c2'])
    w = workloads.build(cfg, cfg['seed'])
    n, L = cfg['N'], w['L']
    g = genome_pack.unpack_genomes(workloads.random_packed_genomes(n, L, cfg['seed'] + 1), L)
    pw = w['prm']
    arch = dict(land_dim=w['land_dim'], rasters=w['rasters'], K=w['rasters'][0] * pw['K_factor'], ww=None,
                traits=w['gen_arch']['traits'], dom=np.zeros(L, np.int8), paths=w['gen_arch']['paths'],
                move_surf=None, disp_surf=None)
    arch['ww'] = round(0.1 * max(cfg['dim']))
    prm = dict(b=pw['b'], R=pw['R'], lam=1, n_births_fixed=True, mating_radius=2.0, d_min=0.0, d_max=1.0,
               sex=False, sex_ratio_p=0.5, max_age=None, direction_mu=0.0, direction_kappa=0.0)
    state = dict(x=w['pop']['x'], y=w['pop']['y'], age=w['pop']['age'], sex=w['pop']['sex'], idx=w['pop']['idx'],
                 g=g, z=so.phenotype(g, arch['traits']), max_ind_idx=n - 1)
    rng = np.random.default_rng(77)
    births_cap = n // 3
    draws = od.make_draws(rng, dict(prm, move_distr=('wald', 1.0, 1.0), disp_distr=('wald', 1.0, 1.0)), n, births_cap,
                          len(arch['paths']), max_tries=12)
    im = _one_big_step(arch, prm, state, draws, int(1.5 * n), 12)
    assert im['B'] > 150000 and im['n_deaths'] > 100000


def test_c4_shaped_clumped_step_matches_oracle():
    """configs[3] shape: 200,000 individuals, 1000 loci (2 x 50 trait loci), conductance-surface
    movement and dispersal (float16 direction tables, injected column choices), 30 % of the
    population in tight clumps so that 3-cell row ranges far beyond 64 candidates occur."""
    from geonomics_b200 import workloads, genome_pack
    from geonomics_b200.api import _make_conductance_surface
    from oracle import step_oracle as so
    from oracle import draws as od
    cfg = workloads.scaled(dict(workloads.CONFIGS['c4']), 200000)
    w = workloads.build(cfg, 2024)
    n, L = cfg['N'], w['L']
    X, Y = w['land_dim']
    rng = np.random.default_rng(99)
    x, y = w['pop']['x'].copy(), w['pop']['y'].copy()
    k = int(0.3 * n)
    centres = rng.uniform(10, X - 10, (40, 2))
    which = rng.integers(0, 40, k)
    x[:k] = np.clip(centres[which, 0] + rng.normal(0, 1.5, k), 0, X - 0.001)
    y[:k] = np.clip(centres[which, 1] + rng.normal(0, 1.5, k), 0, Y - 0.001)
    g = genome_pack.unpack_genomes(workloads.random_packed_genomes(n, L, 5), L)
    pw = w['prm']
    np.random.seed(8)
    A = 6
    tab = _make_conductance_surface(w['rasters'][0], mixture=True, approx_len=A, vm_distr_kappa=12)
    arch = dict(land_dim=w['land_dim'], rasters=w['rasters'], K=w['rasters'][0] * pw['K_factor'], ww=None,
                traits=w['gen_arch']['traits'], dom=np.zeros(L, np.int8), paths=w['gen_arch']['paths'],
                move_surf=tab, disp_surf=tab)
    arch['ww'] = round(0.1 * max(cfg['dim']))
    prm = dict(b=pw['b'], R=pw['R'], lam=1, n_births_fixed=True, mating_radius=2.0, d_min=0.0, d_max=1.0,
               sex=False, sex_ratio_p=0.5, max_age=None, direction_mu=0.0, direction_kappa=0.0)
    state = dict(x=x, y=y, age=w['pop']['age'], sex=w['pop']['sex'], idx=w['pop']['idx'], g=g,
                 z=so.phenotype(g, arch['traits']), max_ind_idx=n - 1)
    births_cap = n // 2
    tries = 12
    draws = od.make_draws(rng, dict(prm, move_distr=('wald', 1.0, 1.0), disp_distr=('wald', 1.0, 1.0)), n, births_cap,
                          len(arch['paths']), max_tries=tries)
    draws['move_choice'] = rng.integers(0, A, n).astype(np.int32)
    draws['disp_choice'] = rng.integers(0, A, (births_cap, tries)).astype(np.int32)
    im = _one_big_step(arch, prm, state, draws, int(1.6 * n), tries)
    assert im['n_nbrs'].max() > 200                           # the > 64-candidate (recount) path ran
    assert im['B'] > 20000
