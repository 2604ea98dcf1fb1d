"""Free-running (Philox) randomness of the CUDA path against the distributions the reference
draws from: the reference's MT19937 stream cannot be reproduced, so the
samplers are checked by Kolmogorov-Smirnov / chi-square tests against scipy's exact CDFs.

  movement direction  numpy vonmises(mu, kappa)                         movement.py:55
  movement distance   numpy wald(mean, scale) / lognormal / scipy levy  movement.py:62-70
  births per pair     clip(poisson(lambda), 1, inf)                     mating.py:120-126
  deleterious s       min(gamma(shape, scale), 1)                       genome.py:690-693

Also: exhausting the preallocated capacity is an error (GNX_ERR_CAPACITY), not a silent clamp.
"""
import numpy as np
import pytest
from scipy import stats

from parity_util import make_device, synthetic_case

pytestmark = pytest.mark.gpu
ALPHA = 1e-4
N = 150000


def _moved(move_distr, mu=0.0, kappa=0.0, seed=1):
    """One free-running movement stage of N individuals parked at the centre of a landscape far
    larger than any step; returns (distance, direction)."""
    dim = (4000, 4000)
    arch, prm, state, draws = synthetic_case(L=32, n=2000, n_traits=0, loci_per_trait=0, dim=(40, 40), seed=3)
    arch = dict(arch, land_dim=dim, rasters=np.ones((arch['rasters'].shape[0],) + dim[::-1]), K=np.ones(dim[::-1]))
    arch['ww'] = round(0.1 * max(dim))
    prm = dict(prm, direction_mu=mu, direction_kappa=kappa)
    dev = make_device(arch, prm, capacity=N + 1000, seed=seed, extra=dict(move_distr=move_distr))
    try:
        x0 = np.full(N, dim[0] / 2.0)
        y0 = np.full(N, dim[1] / 2.0)
        dev.set_burn(True)
        dev.upload(x0, y0, np.zeros(N, np.int32), np.zeros(N, np.int8), np.arange(N, dtype=np.int64), max_ind_idx=N - 1)
        dev.set_draws(None)
        dev.stage('move')
        dev.sync()
        x, y = dev.read('X', N), dev.read('Y', N)
    finally:
        dev.close()
    dx, dy = x - x0, y - y0
    return np.hypot(dx, dy), np.arctan2(dy, dx)


def test_wald_distance_and_uniform_direction():
    d, th = _moved(('wald', 1.0, 1.0))
    # numpy wald(mean, scale) = inverse Gaussian: scipy invgauss(mu=mean/scale, scale=scale)
    assert stats.kstest(d, stats.invgauss(1.0, scale=1.0).cdf).pvalue > ALPHA
    # vonmises(mu, kappa < 1e-8) = uniform on (-pi, pi)   (numpy legacy, movement.py:55)
    assert stats.kstest(th, stats.uniform(-np.pi, 2 * np.pi).cdf).pvalue > ALPHA
    assert abs(np.corrcoef(d, np.cos(th))[0, 1]) < 0.01          # distance and direction independent


def test_wald_other_parameters():
    d, _ = _moved(('wald', 0.6, 2.5), seed=2)
    assert stats.kstest(d, stats.invgauss(0.6 / 2.5, scale=2.5).cdf).pvalue > ALPHA


def test_lognormal_distance_and_von_mises_direction():
    d, th = _moved(('lognormal', 0.0, 0.5), mu=0.7, kappa=2.0, seed=3)
    assert stats.kstest(d, stats.lognorm(0.5, scale=np.exp(0.0)).cdf).pvalue > ALPHA
    # direction ~ von Mises(0.7, 2): compare on the circle relative to mu so the CDF has no wrap
    rel = np.angle(np.exp(1j * (th - 0.7)))
    assert stats.kstest(rel, stats.vonmises(2.0).cdf).pvalue > ALPHA


def test_levy_distance():
    d, _ = _moved(('levy', 0.0, 0.05), seed=4)
    # heavy tail: the landscape clamps steps beyond 2000 cells; compare below that
    keep = d < 1500
    assert keep.mean() > 0.99
    c = stats.levy(0.0, 0.05)
    u = c.cdf(d[keep]) / c.cdf(1500)
    assert stats.kstest(u, 'uniform').pvalue > ALPHA


def test_poisson_births_clipped_at_one():
    arch, prm, state, draws = synthetic_case(L=32, n=60000, n_traits=0, loci_per_trait=0, dim=(250, 250), seed=9,
                                             max_tries=24)
    lam = 1.7
    prm = dict(prm, n_births_fixed=False, lam=lam, b=0.9)
    dev = make_device(arch, prm, capacity=400000, seed=11)
    try:
        dev.upload(state['x'], state['y'], state['age'], state['sex'], state['idx'], g=state['g'],
                   max_ind_idx=state['max_ind_idx'])
        dev.set_draws(None)
        for st in ('age_step', 'move', 'bin_cells', 'find_mates', 'dedup_pairs'):
            dev.stage(st)
        dev.sync()
        P = dev.counters()['P']
        nb = dev.read('NB', P)
    finally:
        dev.close()
    assert P > 10000 and nb.min() >= 1
    kmax = 8
    obs = np.array([(nb == k).sum() for k in range(1, kmax)] + [(nb >= kmax).sum()], dtype=float)
    pm = stats.poisson(lam)
    exp = np.array([pm.cdf(1)] + [pm.pmf(k) for k in range(2, kmax)] + [pm.sf(kmax - 1)]) * P
    assert stats.chisquare(obs, exp).pvalue > ALPHA


def test_deleterious_s_is_clipped_gamma():
    arch, prm, state, draws = synthetic_case(L=4000, n=3000, n_traits=0, loci_per_trait=0, seed=13, max_tries=24)
    mutables = [int(v) for v in np.random.default_rng(2).permutation(4000)]
    dev = make_device(arch, prm, capacity=12000, seed=17)
    try:
        dev.upload(state['x'], state['y'], state['age'], state['sex'], state['idx'], g=state['g'],
                   max_ind_idx=state['max_ind_idx'])
        dev.set_draws(None)
        shape, scale = 0.7, 0.4
        dev.set_mutation(0.0, 3e-4, mutables, np.zeros(0, np.int64), delet_s_shape=shape, delet_s_scale=scale,
                         log_capacity=8192)
        dev.step(30)
        dev.sync()
        log, st = dev.read_mutations(max_rows=8192)
    finally:
        dev.close()
    s = np.array([r['s'] for r in log])
    assert len(s) > 800 and np.all(s <= 1.0)
    g = stats.gamma(shape, scale=scale)
    below = s < 1.0
    assert abs((~below).mean() - g.sf(1.0)) < 0.02
    assert stats.kstest(g.cdf(s[below]) / g.cdf(1.0), 'uniform').pvalue > ALPHA


def test_capacity_overflow_is_reported():
    from geonomics_b200._lib import GnxError
    arch, prm, state, draws = synthetic_case(L=64, n=3000, n_traits=0, loci_per_trait=0, seed=21, max_tries=24)
    dev = make_device(arch, dict(prm, b=0.9), capacity=3050, seed=5)      # room for 50 births only
    try:
        dev.upload(state['x'], state['y'], state['age'], state['sex'], state['idx'], g=state['g'],
                   max_ind_idx=state['max_ind_idx'])
        dev.set_draws(None)
        dev.step(1)
        with pytest.raises(GnxError) as ei:
            dev.sync()
        assert ei.value.code == -3
    finally:
        dev.close()
