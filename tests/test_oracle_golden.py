"""Pins the numpy oracle (oracle/step_oracle.py) against vectors recorded from the
unmodified reference (tests/golden/make_golden.py).  CPU only."""
import numpy as np
import pytest

from oracle import step_oracle as so
from golden_io import case_names, load_case

CASES = case_names()


@pytest.fixture(scope='module', params=CASES)
def case(request):
    z, arch, prm, state, draws = load_case(request.param)
    new, im = so.step(state, arch, prm, draws, burn=prm.get('burn', False))
    return request.param, z, arch, prm, state, draws, new, im


def test_cases_present():
    assert {'base', 'sexed', 'surf', 'burn', 'mut', 'pan', 'nearest', 'invdist'} <= set(CASES)


def test_age_and_movement_bit_exact(case):
    _, z, arch, prm, state, draws, new, im = case
    assert np.array_equal(im['mv_age'], z['mv_age'])
    assert np.array_equal(im['mv_x'], z['mv_x'])
    assert np.array_equal(im['mv_y'], z['mv_y'])
    assert np.array_equal(im['mv_e'], z['mv_e'])
    cx, cy = so.cells(im['mv_x'], im['mv_y'])
    assert np.array_equal(np.stack([cx, cy], axis=1), z['mv_cells'])


def test_neighbour_sets_match_reference(case):
    _, z, arch, prm, state, draws, new, im = case
    assert np.array_equal(im['n_nbrs'], z['n_nbrs'])
    if prm['mating_radius'] is None:
        pytest.skip('panmixia: no neighbour search')
    if prm.get('choose_nearest'):
        pytest.skip('nearest-neighbour mode: the reference never lists neighbour sets (cKDTree.query)')
    nb = so.neighbor_lists_bruteforce(z['mv_x'], z['mv_y'], arch['land_dim'], prm['mating_radius'])
    ip, ix = z['ref_nbr_indptr'], z['ref_nbr_indices']
    for i in range(len(nb)):
        assert np.array_equal(np.sort(nb[i]), ix[ip[i]:ip[i + 1]])


def test_pairs(case):
    _, z, arch, prm, state, draws, new, im = case
    assert np.array_equal(im['pairs'], z['pairs'])
    # set-wise against the raw reference pair list (hash-ordered ids, mating.py:63,109-113)
    ids = z['in_idx']
    ref = {frozenset(map(int, p)) for p in z['ref_pairs_ids']}
    mine = {frozenset(map(int, ids[p])) for p in im['pairs']}
    assert ref == mine
    assert len(z['ref_pairs_ids']) == len(im['pairs'])


def test_births_and_gametes_bit_exact(case):
    _, z, arch, prm, state, draws, new, im = case
    assert np.array_equal(im['nb'], z['nb'])
    assert im['B'] == int(z['B'])
    assert np.array_equal(im['mid_x'], z['mid_x'])
    assert np.array_equal(im['mid_y'], z['mid_y'])
    assert np.array_equal(im['disp_tries'], z['disp_tries'])
    pre = im['pre']
    assert np.array_equal(pre['idx'], z['pre_idx'])
    if not prm.get('burn'):
        assert np.array_equal(pre['g'], z['pre_g'])
    assert np.array_equal(pre['sex'], z['pre_sex'])
    assert np.array_equal(pre['age'], z['pre_age'])
    assert np.array_equal(pre['x'], z['pre_x'])
    assert np.array_equal(pre['y'], z['pre_y'])


def test_phenotype(case):
    _, z, arch, prm, state, draws, new, im = case
    if prm.get('burn'):
        pytest.skip('no genomes during burn-in')
    np.testing.assert_allclose(im['pre']['z'], z['pre_z'], rtol=1e-12, atol=0)


def test_density_and_d_rasters(case):
    _, z, arch, prm, state, draws, new, im = case
    np.testing.assert_allclose(im['n_pairs_rast'], z['n_pairs_rast'], rtol=1e-12, atol=1e-14)
    np.testing.assert_allclose(im['N_rast'], z['N_rast'], rtol=1e-12, atol=1e-14)
    np.testing.assert_allclose(im['d_rast'], z['d_rast'], rtol=1e-10, atol=1e-14)


def test_restated_clough_tocher_matches_scipy(case):
    _, z, arch, prm, state, draws, new, im = case
    dgs = so.DensityGridStack(arch['land_dim'], arch['ww'])
    pre = im['pre']
    vals = dgs.vals(pre['x'], pre['y'])
    mine = np.clip(so.ct_density_restated(dgs, vals), 0, None)
    np.testing.assert_allclose(mine, z['N_rast'], rtol=1e-9, atol=1e-12)


def test_fitness_and_death_probs(case):
    _, z, arch, prm, state, draws, new, im = case
    if not prm.get('burn'):
        np.testing.assert_allclose(im['fit_all'], z['fit_all'], rtol=1e-12)
    np.testing.assert_allclose(im['death_p'], z['death_p'], rtol=1e-10, atol=1e-15)


def test_mortality_and_final_state(case):
    _, z, arch, prm, state, draws, new, im = case
    assert im['n_deaths'] == int(z['out_n_deaths'])
    assert len(new['x']) == int(z['out_Nt'])
    assert np.array_equal(new['idx'], z['out_idx'])
    assert np.array_equal(new['x'], z['out_x'])
    assert np.array_equal(new['y'], z['out_y'])
    assert np.array_equal(new['age'], z['out_age'])
    assert np.array_equal(new['sex'], z['out_sex'])
    if not prm.get('burn'):
        assert np.array_equal(new['g'], z['out_g'])
        np.testing.assert_allclose(new['z'], z['out_z'], rtol=1e-12)
        np.testing.assert_allclose(new['fit'], z['out_fit'], rtol=1e-12)
    assert new['max_ind_idx'] == int(z['out_max_ind_idx'])


def test_mutation_bookkeeping(case):
    name, z, arch, prm, state, draws, new, im = case
    if arch.get('mutation') is None:
        pytest.skip('no mutation in this case')
    m = im['mutation']
    assert np.array_equal(m['mutables'], z['out_mut_mutables'])
    assert np.array_equal(m['nonneut_loci'], z['out_mut_nonneut_loci'])
    assert np.array_equal(m['delet_loci'], z['out_mut_delet_loci'])
    assert np.array_equal(m['delet_s'], z['out_mut_delet_s'])
    assert len(im['mut_log']) == int(draws['mut_n'][0])
    if not m.get('tskit_layout'):
        assert {r['type'] for r in im['mut_log']} == {'neut', 'delet'}      # the case exercises both
        return
    # use_tskit = True with trait mutation (tmut): every table the reference edits, as it leaves them
    assert {r['type'] for r in im['mut_log']} == {'neut', 'delet', 't0', 't1'}
    assert np.array_equal(m['delet_loci_idxs'], z['out_mut_delet_loci_idxs'])
    for t, tr in enumerate(m['traits']):
        assert np.array_equal(tr['loci'], z['out_trait%i_loci' % t])
        assert np.array_equal(tr['alpha'], z['out_trait%i_alpha' % t])
        assert np.array_equal(tr['loci_idxs'], z['out_trait%i_loci_idxs' % t])
    assert np.array_equal(m['subsetters'], z['out_subsetters'])
    # the index arrays really are stale in this case (the reference shifts loci_idxs only inside the mutated trait)
    nn = z['out_mut_nonneut_loci']
    assert any(not np.array_equal(nn[tr['loci_idxs']], tr['loci']) for tr in m['traits'])
    # mutations-table rows (mutation.py:44-58): site = locus, in event order
    assert np.array_equal([r['locus'] for r in im['mut_log']], z['tsk_mut_site'])


def test_pack_roundtrip(case):
    _, z, arch, prm, state, draws, new, im = case
    g = z['in_g']
    if g.shape[1] == 0:
        pytest.skip('no genotype rows')
    p = so.pack_genomes(g)
    assert p.shape[2] % 4 == 0
    assert np.array_equal(so.unpack_genomes(p, g.shape[1]), g)
