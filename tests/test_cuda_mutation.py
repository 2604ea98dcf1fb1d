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
