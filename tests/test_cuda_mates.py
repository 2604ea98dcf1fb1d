"""GPU: neighbour scan / mate choice against the oracle for all three choice modes, sparse
and dense (row ranges > 32 candidates -> second-walk path) populations, and edge cases
(isolated individuals, coincident positions, landscape corners)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _run(x, y, dim, radius, b, mode, seed):
    from geonomics_b200.device import DeviceSpecies
    from oracle import step_oracle as so
    n = len(x)
    rng = np.random.default_rng(seed)
    rasters = np.ones((1, dim[1], dim[0]))
    prm = dict(b=b, R=0.5, lam=1, n_births_fixed=True, mating_radius=radius, d_min=0, d_max=1, sex=False,
               K_layer=0, K_factor=1.0, move=False, choose_nearest=(mode == 'nearest'),
               inverse_dist=(mode == 'inverse'))
    draws = dict(mate_R=rng.integers(0, 1 << 32, n, dtype=np.uint64).astype(np.uint32), mate_u=rng.random(n),
                 mate_inv_u=rng.random(n))
    dev = DeviceSpecies(dim, rasters, prm, None, capacity=2 * n + 64)
    try:
        dev.set_debug(True)
        dev.set_burn(True)
        dev.upload(x, y)
        dev.set_draws(draws)
        dev.stage('bin_cells')
        dev.stage('find_mates')
        dev.stage('dedup_pairs')
        dev.sync()
        mate = dev.read('MATE', n)
        nn = dev.read('N_NBRS', n)
        P = dev.counters()['P']
        pairs = dev.read('PAIRS', 2 * P).reshape(-1, 2)
        perm = dev.read('PERM', n)
    finally:
        dev.close()
    pairs_o, nn_o, mate_o = so.find_mates_radius(x, y, dim, radius, b, draws['mate_R'], draws['mate_u'],
                                                 choose_nearest=(mode == 'nearest'),
                                                 inverse_dist=(mode == 'inverse'), inv_u=draws['mate_inv_u'])
    rank, perm_o, _ = so.canonical_rank(x, y, dim, radius)
    assert np.array_equal(perm, perm_o)                       # stable counting sort
    assert np.array_equal(nn, nn_o)
    assert np.array_equal(mate, mate_o)
    assert np.array_equal(pairs, pairs_o)
    return nn


@pytest.mark.parametrize('mode', ['random', 'nearest', 'inverse'])
def test_sparse_population(mode):
    rng = np.random.default_rng(1)
    dim = (64, 48)
    n = 600
    x = rng.uniform(0, dim[0] - 0.001, n)
    y = rng.uniform(0, dim[1] - 0.001, n)
    nn = _run(x, y, dim, 2.0, 0.7, mode, 2)
    assert nn.max() < 32 and (nn == 0).any()


@pytest.mark.parametrize('mode', ['random', 'nearest', 'inverse'])
def test_dense_population_overflow_path(mode):
    rng = np.random.default_rng(3)
    dim = (20, 20)
    n = 2500
    x = rng.uniform(0, dim[0] - 0.001, n)
    y = rng.uniform(0, dim[1] - 0.001, n)
    nn = _run(x, y, dim, 3.5, 0.5, mode, 4)
    assert nn.max() > 100                                      # row ranges far beyond 32 candidates


@pytest.mark.parametrize('mode', ['random', 'nearest', 'inverse'])
def test_medium_density_second_mask_path(mode):
    """Row ranges of 33-64 candidates (clumped, evolved populations): the second 32-bit mask."""
    rng = np.random.default_rng(8)
    dim = (20, 20)
    n = 1400                                                   # ~14 per mating-grid cell, ~42 per 3-cell run
    x = rng.uniform(0, dim[0] - 0.001, n)
    y = rng.uniform(0, dim[1] - 0.001, n)
    nn = _run(x, y, dim, 2.0, 0.6, mode, 9)
    assert 33 < nn.max() < 110 and np.median(nn) > 25


def test_crowded_cells_beyond_mask_capacity():
    """More than 2048 candidates in a 3x3 block: k_find_mates_dense recounts for the pick."""
    rng = np.random.default_rng(21)
    dim = (12, 12)
    n = 4200
    x = rng.uniform(0, dim[0] - 0.001, n)
    y = rng.uniform(0, dim[1] - 0.001, n)
    nn = _run(x, y, dim, 3.5, 0.5, 'random', 22)
    assert nn.max() > 1000


def test_clumped_population_uses_both_kernels():
    """A density gradient from empty to crowded: sparse cells stay with the thread-per-focal
    kernel, crowded ones (row range > 64 or focals x candidates >= 4096) go to the warp-per-
    cell kernel; odd focal counts, cells on the landscape edge and in the corners included."""
    rng = np.random.default_rng(23)
    dim = (48, 40)
    n = 9000
    x = dim[0] * rng.beta(0.6, 2.5, n) * 0.9999
    y = dim[1] * rng.beta(2.5, 0.6, n) * 0.9999
    nn = _run(x, y, dim, 2.0, 0.4, 'random', 24)
    assert nn.max() > 200 and (nn < 10).sum() > 100


def test_edge_cases():
    dim = (10, 10)
    # corners, an exactly coincident couple, a neighbour at exactly distance r, an isolate
    x = np.array([0.0, 9.999, 0.0, 9.999, 5.0, 5.0, 2.0, 4.0, 7.5])
    y = np.array([0.0, 0.0, 9.999, 9.999, 5.0, 5.0, 2.0, 2.0, 1.0])
    _run(x, y, dim, 2.0, 1.0, 'random', 5)
    _run(x, y, dim, 2.0, 1.0, 'nearest', 6)
    _run(x, y, dim, 2.0, 1.0, 'inverse', 7)


def test_empty_and_single():
    from geonomics_b200.device import DeviceSpecies
    rasters = np.ones((1, 8, 8))
    prm = dict(b=1.0, R=0.5, lam=1, n_births_fixed=True, mating_radius=2.0, K_layer=0, K_factor=1.0, move=False)
    for n in (0, 1):
        dev = DeviceSpecies((8, 8), rasters, prm, None, capacity=64)
        try:
            dev.set_burn(True)
            dev.upload(np.full(n, 3.0), np.full(n, 3.0))
            dev.step(2)
            dev.sync()
            recs = dev.step_records()
            assert [r['n_births'] for r in recs] == [0, 0]
            assert dev.population_size() <= n
        finally:
            dev.close()


@pytest.mark.parametrize('sexed,b', [(False, 0.3), (True, 0.8), (False, 1.0)])
def test_panmixia_full_step(sexed, b):
    """mating_radius = None (species.py:2178-2194): Wright-Fisher style random pairing."""
    from oracle import step_oracle as so
    from parity_util import synthetic_case, run_device_step, compare_step
    arch, prm, state, draws = synthetic_case(L=64, n=1500, loci_per_trait=8, seed=31, max_tries=24)
    prm = dict(prm, mating_radius=None, b=b, sex=sexed)
    new_o, im_o = so.step(state, arch, prm, draws)
    expected = 1500 * b * (0.25 if sexed else 1.0)
    assert abs(len(im_o['pairs']) - expected) < 5 * np.sqrt(expected)
    assert np.all(im_o['pairs'][:, 0] != im_o['pairs'][:, 1])
    if sexed:
        assert np.all(state['sex'][im_o['pairs'][:, 0]] == 0) and np.all(state['sex'][im_o['pairs'][:, 1]] == 1)
    out = run_device_step(arch, prm, state, draws, staged=True)
    compare_step(out, new_o, im_o)
