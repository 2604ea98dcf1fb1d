"""GPU parity of a13 mutation (ops/mutation.py:169-206, use_tskit=False): the device's
bookkeeping, genotype edits, phenotypes and deleterious fitness against the vectors recorded
from the reference (tests/golden/step_mut.npz) and the oracle; plus free-running (Philox)
consistency checks."""
import numpy as np
import pytest

from golden_io import load_case
from parity_util import run_device_step, make_device, synthetic_case

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def mut_step():
    from oracle import step_oracle as so
    z, arch, prm, state, draws = load_case('mut')
    new_o, im_o = so.step(state, arch, prm, draws)
    out = run_device_step(arch, prm, state, draws, staged=True)
    return z, arch, draws, out, new_o, im_o


def test_bookkeeping_matches_reference(mut_step):
    z, arch, draws, out, new_o, im_o = mut_step
    st = out['mutation']
    assert st['n_mutables'] == len(z['out_mut_mutables'])
    assert np.array_equal(st['nonneut_loci'], z['out_mut_nonneut_loci'])
    assert np.array_equal(st['delet_loci'], z['out_mut_delet_loci'])
    assert np.array_equal(st['delet_s'], z['out_mut_delet_s'])          # bit-exact: min(injected gamma, 1)


def test_log_matches_oracle(mut_step):
    z, arch, draws, out, new_o, im_o = mut_step
    log_d, log_o = out['mut_log'], im_o['mut_log']
    assert len(log_d) == len(log_o) == int(draws['mut_n'][0])
    for a, b in zip(log_d, log_o):
        for k in ('individual', 'locus', 'row', 'homologue', 'type'):
            assert a[k] == b[k], (k, a, b)
        assert a['s'] == b['s']


def test_mutated_genotypes_phenotypes_fitness(mut_step):
    z, arch, draws, out, new_o, im_o = mut_step
    # survivors' genotypes (incl. the rows written by mutation.py:117) against the reference
    assert np.array_equal(out['new']['g'], z['out_g'])
    np.testing.assert_allclose(out['pre']['z'], z['pre_z'], rtol=1e-12, atol=1e-15)
    # fitness of everyone alive before mortality includes prod(1 - s * dosage) (selection.py:78-94)
    np.testing.assert_allclose(out['fit_all'], z['fit_all'], rtol=1e-6)
    np.testing.assert_allclose(out['death_p'], z['death_p'], rtol=1e-6, atol=1e-9)
    assert out['records'][-1]['n_deaths'] == int(z['out_n_deaths'])


@pytest.fixture(scope='module')
def tmut_step():
    """use_tskit = True with neutral + deleterious + TRAIT mutation and tskit rows, recorded from the reference
    (tests/golden/step_tmut.npz: the reference runs on the functional table shim of oracle/ref_shims.py)."""
    from oracle import step_oracle as so
    z, arch, prm, state, draws = load_case('tmut')
    new_o, im_o = so.step(state, arch, prm, draws)
    tsk = dict(node0=z['in_nodes'][:, 0], node1=z['in_nodes'][:, 1], next_node_id=int(z['tsk_in_rows'][0]),
               next_individual_row=int(z['tsk_in_rows'][2]), edge_capacity=len(z['tsk_edge_left']) + 1000)
    out = run_device_step(arch, prm, state, draws, staged=True, tskit=tsk)
    return z, arch, draws, out, new_o, im_o


def test_tskit_layout_bookkeeping_matches_reference(tmut_step):
    z, arch, draws, out, new_o, im_o = tmut_step
    st = out['mutation']
    assert st['n_mutables'] == len(z['out_mut_mutables'])
    assert np.array_equal(st['nonneut_loci'], z['out_mut_nonneut_loci'])
    assert np.array_equal(st['delet_loci'], z['out_mut_delet_loci'])
    assert np.array_equal(st['delet_s'], z['out_mut_delet_s'])
    # the index arrays as the reference leaves them (genome.py:416-437 shifts loci_idxs inside the mutated
    # trait only, genome.py:779-782 shifts nothing): stale on purpose
    assert np.array_equal(out['mut_delet_loci_idxs'], z['out_mut_delet_loci_idxs'])
    for t, tr in enumerate(out['mut_traits']):
        assert np.array_equal(tr['loci'], z['out_trait%i_loci' % t])
        assert np.array_equal(tr['alpha'], z['out_trait%i_alpha' % t])          # bit-exact: clip(injected normal)
        assert np.array_equal(tr['loci_idxs'], z['out_trait%i_loci_idxs' % t])
    log_d, log_o = out['mut_log'], im_o['mut_log']
    assert len(log_d) == len(log_o) == int(draws['mut_n'][0])
    assert {r['type'] for r in log_d} == {'neut', 'delet', 't0', 't1'}
    for a, b in zip(log_d, log_o):
        for k in ('individual', 'locus', 'row', 'homologue', 'type', 's', 'alpha'):
            assert a[k] == b[k], (k, a, b)


def test_tskit_layout_genotypes_phenotypes_fitness(tmut_step):
    z, arch, draws, out, new_o, im_o = tmut_step
    # genotype ROWS of the survivors (gametes through the subsetters, rows inserted by the mutations)
    assert np.array_equal(out['new']['g'], z['out_g'])
    # nothing but the non-neutral loci is ever set in the by-locus device rows
    mask = np.ones(out['new']['g_loci'].shape[1], bool)
    mask[z['out_mut_nonneut_loci']] = False
    assert not out['new']['g_loci'][:, mask, :].any()
    # phenotypes through the (stale) loci_idxs, deleterious fitness through delet_loci_idxs
    np.testing.assert_allclose(out['pre']['z'], z['pre_z'], rtol=1e-12, atol=1e-15)
    np.testing.assert_allclose(out['fit_all'], z['fit_all'], rtol=1e-6)
    np.testing.assert_allclose(out['death_p'], z['death_p'], rtol=1e-6, atol=1e-9)
    assert np.array_equal(out['new']['idx'], z['out_idx'])
    np.testing.assert_allclose(out['new']['z'], z['out_z'], rtol=1e-12, atol=1e-15)


def test_tskit_layout_next_step_uses_patched_subsetters(tmut_step):
    """The subsetters the mutations edited (genome.py:133-160) are what the NEXT step's gametes go through."""
    z, arch, draws, out, new_o, im_o = tmut_step
    from oracle import step_oracle as so
    from oracle import draws as od
    arch2 = dict(arch, mutation=im_o['mutation'], traits=im_o['mutation']['traits'])
    assert np.array_equal(arch2['mutation']['subsetters'], z['out_subsetters'])
    rng = np.random.default_rng(77)
    n1 = len(new_o['x'])
    prm = load_case('tmut')[2]
    d2 = od.make_draws(rng, dict(prm, move_distr=('wald', 1.0, 1.0), disp_distr=('wald', 0.8, 1.0)), n1, 2 * n1 + 64,
                       len(arch['paths']), max_tries=24)
    nm = 4
    d2.update(mut_n=np.array([nm], np.int32), mut_type_u=rng.random(nm),
              mut_ind_R=rng.integers(0, 2**32, nm, dtype=np.uint64).astype(np.uint32),
              mut_homol_u=rng.random(nm), mut_s=rng.gamma(0.2, 0.2, nm), mut_alpha=rng.normal(0, 0.15, nm))
    state2 = dict(new_o)
    new2_o, im2_o = so.step(state2, arch2, prm, d2)
    out2 = run_device_step(arch2, prm, state2, d2, staged=False)
    assert np.array_equal(out2['new']['idx'], new2_o['idx'])
    assert np.array_equal(out2['new']['g'], new2_o['g'])
    np.testing.assert_allclose(out2['new']['z'], new2_o['z'], rtol=1e-12, atol=1e-15)
    np.testing.assert_allclose(out2['new']['fit'], new2_o['fit'], rtol=1e-6)


def test_tskit_rows_of_the_step_match_reference(tmut_step):
    """species.py:692-736 (individuals / nodes / edges rows per birth) and mutation.py:44-58 (mutations rows)."""
    z, arch, draws, out, new_o, im_o = tmut_step
    rows = out['tskit_rows']
    assert rows['first_node_id'] == int(z['tsk_in_rows'][0])
    assert rows['first_individual_row'] == int(z['tsk_in_rows'][2])
    assert np.array_equal(rows['left'], z['tsk_edge_left'])
    assert np.array_equal(rows['right'], z['tsk_edge_right'])
    assert np.array_equal(rows['parent'], z['tsk_edge_parent'])
    assert np.array_equal(rows['child'], z['tsk_edge_child'])
    assert np.array_equal(rows['idx'], z['tsk_ind_idx'])
    assert np.array_equal(rows['node_individual'], z['tsk_node_individual'])
    loc = z['tsk_ind_location']                       # [x, y, z0, z1, fit]: z at birth, before the mutations
    np.testing.assert_allclose(rows['x'], loc[:, 0], rtol=0, atol=1e-9)
    np.testing.assert_allclose(rows['y'], loc[:, 1], rtol=0, atol=1e-9)
    np.testing.assert_allclose(rows['z'], loc[:, 2:4], rtol=1e-12)
    # mutations table: site, node of the mutated homologue, in event order
    assert np.array_equal([r['locus'] for r in out['mut_log']], z['tsk_mut_site'])
    assert np.array_equal([r['node'] for r in out['mut_log']], z['tsk_mut_node'])


def test_trait_mutation_needs_tskit_layout():
    """use_tskit = False: the reference raises in Trait._add_locus (genome.py:430); rejected here."""
    from geonomics_b200._lib import GnxError
    arch, prm, state, draws = synthetic_case(L=64, n=300, n_traits=1, loci_per_trait=4, seed=44, max_tries=24)
    dev = make_device(arch, prm, capacity=2000, seed=7)
    try:
        with pytest.raises(GnxError):
            dev.set_mutation(1e-5, 0.0, [5, 9], np.sort(np.asarray(arch['traits'][0]['loci'])),
                             trait_mus=[1e-5], trait_alpha_distr=[(0.0, 0.1, None)])
    finally:
        dev.close()


def test_trait_less_deleterious_selection():
    """mu_delet > 0 switches selection on even without traits (species.py:449-451)."""
    from oracle import step_oracle as so
    arch, prm, state, draws = synthetic_case(L=200, n=1200, n_traits=0, loci_per_trait=0, seed=31, max_tries=24)
    rng = np.random.default_rng(5)
    arch['mutation'] = dict(mu_neut=1e-5, mu_delet=3e-5, mutables=[int(v) for v in rng.permutation(200)[:150]],
                            nonneut_loci=np.zeros(0, np.int64), delet_loci=np.array([3, 77]),
                            delet_s=np.array([0.3, 0.05]), s_shape=0.2, s_scale=0.2)
    nm = 5
    draws.update(mut_n=np.array([nm], np.int32), mut_type_u=rng.random(nm),
                 mut_ind_R=rng.integers(0, 2**32, nm, dtype=np.uint64).astype(np.uint32),
                 mut_homol_u=rng.random(nm), mut_s=rng.gamma(0.2, 0.2, nm))
    new_o, im_o = so.step(state, arch, prm, draws)
    out = run_device_step(arch, prm, state, draws, staged=True)
    assert im_o['fit_all'] is not None and (im_o['fit_all'] < 1).any()
    np.testing.assert_allclose(out['fit_all'], im_o['fit_all'], rtol=1e-6)
    np.testing.assert_allclose(out['death_p'], im_o['death_p'], rtol=1e-6, atol=1e-9)
    assert np.array_equal(out['new']['g'], new_o['g'])
    assert np.array_equal(out['new']['idx'], new_o['idx'])
    assert np.array_equal(out['mutation']['delet_loci'], im_o['mutation']['delet_loci'])


def test_free_running_mutation_statistics():
    """Philox path: the number of mutations per step is Binomial(B*L, mu_tot); each one pops a
    distinct mutable locus; deleterious ones appear in the tables in ascending order."""
    arch, prm, state, draws = synthetic_case(L=300, n=3000, n_traits=1, loci_per_trait=10, seed=41, max_tries=24)
    mu_neut, mu_delet = 2e-6, 2e-6
    trait_loci = np.asarray(arch['traits'][0]['loci'])
    mutables = [int(v) for v in np.random.default_rng(9).permutation(np.setdiff1d(np.arange(300), trait_loci))]
    dev = make_device(arch, prm, capacity=12000, seed=123)
    try:
        dev.upload(state['x'], state['y'], state['age'], state['sex'], state['idx'], g=state['g'], z=state['z'],
                   max_ind_idx=state['max_ind_idx'])
        dev.set_draws(None)
        dev.set_mutation(mu_neut, mu_delet, mutables, np.sort(trait_loci))
        steps = 40
        dev.step(steps)
        dev.sync()
        log, st = dev.read_mutations()
        births = sum(r['n_births'] for r in dev.step_records()[-steps:])
    finally:
        dev.close()
    expect = births * 300 * (mu_neut + mu_delet)
    assert expect > 20
    assert abs(len(log) - expect) < 6 * np.sqrt(expect), (len(log), expect)
    loci = [r['locus'] for r in log]
    assert len(set(loci)) == len(loci)                      # infinite sites
    assert loci == mutables[::-1][:len(loci)]               # popped from the end, in order
    assert st['n_mutables'] == len(mutables) - len(log)
    n_del = sum(r['type'] == 'delet' for r in log)
    assert 0 < n_del < len(log)
    assert len(st['delet_loci']) == n_del and np.all(np.diff(st['delet_loci']) > 0)
    assert np.all(np.diff(st['nonneut_loci']) > 0) and len(st['nonneut_loci']) == len(trait_loci) + n_del
    s = np.array([r['s'] for r in log if r['type'] == 'delet'])
    assert np.all((s > 0) & (s <= 1))


def test_mutables_exhausted_is_an_error():
    from geonomics_b200._lib import GnxError
    arch, prm, state, draws = synthetic_case(L=64, n=800, n_traits=1, loci_per_trait=4, seed=43, max_tries=24)
    dev = make_device(arch, prm, capacity=4000, seed=7)
    try:
        dev.upload(state['x'], state['y'], state['age'], state['sex'], state['idx'], g=state['g'], z=state['z'],
                   max_ind_idx=state['max_ind_idx'])
        dev.set_draws(None)
        dev.set_mutation(1e-3, 0.0, [5, 9], np.sort(np.asarray(arch['traits'][0]['loci'])))
        dev.step(3)
        with pytest.raises(GnxError) as ei:
            dev.sync()
            dev.read_mutations()
        assert ei.value.code == -6
    finally:
        dev.close()
