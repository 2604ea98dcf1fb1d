"""Burn-in control (sim/burnin.py, structs/community.py:107-131): the host tests (CPU) and the
device-side per-cell count statistic (GPU)."""
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))

from geonomics_b200 import burnin  # noqa: E402


def test_mackinnon_p_values_reproduce_the_published_critical_values():
    # MacKinnon's table, constant only, asymptotic: 1 % -3.43, 5 % -2.86, 10 % -2.57
    assert abs(burnin.mackinnonp(-3.43) - 0.01) < 1.5e-3
    assert abs(burnin.mackinnonp(-2.86) - 0.05) < 2e-3
    assert abs(burnin.mackinnonp(-2.57) - 0.10) < 3e-3
    # the two polynomials meet at tau_star; the p-value is monotone in the statistic
    eps = 1e-9
    assert abs(burnin.mackinnonp(-1.61 - eps) - burnin.mackinnonp(-1.61 + eps)) < 2e-3
    ts = np.linspace(-18, 2.7, 400)
    ps = np.array([burnin.mackinnonp(t) for t in ts])
    assert np.all(np.diff(ps) >= -1e-12) and ps[0] < 1e-12 and ps[-1] > 0.99
    assert burnin.mackinnonp(3.0) == 1.0 and burnin.mackinnonp(-19.0) == 0.0


def test_adfuller_lag0_statistic_is_the_dickey_fuller_t_ratio():
    """With no lagged differences selected the statistic is the plain OLS t-ratio of rho in
    dx_t = a + rho x_{t-1} + e_t; checked against the normal equations written out by hand."""
    rng = np.random.default_rng(3)
    x = np.zeros(60)
    for t in range(1, 60):
        x[t] = 0.2 * x[t - 1] + rng.normal()
    stat, p, lag, nobs = burnin.adfuller(x)
    dx, lvl = np.diff(x), x[:-1]
    dx, lvl = dx[lag:], lvl[lag:]
    cols = [lvl] + [np.diff(x)[lag - k:len(np.diff(x)) - k] for k in range(1, lag + 1)] + [np.ones(len(dx))]
    X = np.column_stack(cols)
    beta = np.linalg.solve(X.T @ X, X.T @ dx)
    res = dx - X @ beta
    s2 = res @ res / (len(dx) - X.shape[1])
    t_ratio = beta[0] / np.sqrt(s2 * np.linalg.inv(X.T @ X)[0, 0])
    assert nobs == len(dx) and abs(stat - t_ratio) < 1e-9 * max(1, abs(t_ratio))


def test_adfuller_separates_stationary_series_from_random_walks():
    rng = np.random.default_rng(11)
    n_stat = n_walk = 0
    for _ in range(40):
        e = rng.normal(size=120)
        ar = np.zeros(120)
        for t in range(1, 120):
            ar[t] = 0.3 * ar[t - 1] + e[t]
        n_stat += burnin.adfuller(ar)[1] < 0.05
        n_walk += burnin.adfuller(np.cumsum(e))[1] < 0.05
    assert n_stat >= 36          # power against phi = 0.3 at n = 120 is ~1
    assert n_walk <= 6           # size 5 %
    with pytest.raises(ValueError):
        burnin.adfuller(np.ones(30))
    with pytest.raises(ValueError):
        burnin.adfuller(np.arange(3.0))


def test_adfuller_against_statsmodels_when_present():
    sm = pytest.importorskip('statsmodels.tsa.stattools')
    if not hasattr(sm.adfuller, '__code__'):
        pytest.skip('statsmodels is shimmed')
    rng = np.random.default_rng(5)
    for n in (30, 75, 200):
        x = np.cumsum(rng.normal(size=n)) * 0.3 + rng.normal(size=n)
        want = sm.adfuller(x)
        got = burnin.adfuller(x)
        assert abs(got[0] - want[0]) < 1e-8 and abs(got[1] - want[1]) < 1e-8 and got[2] == want[2]


def test_t_threshold_is_the_paired_test_on_the_two_halves():
    from scipy.stats import ttest_rel
    rng = np.random.default_rng(2)
    Nt = list(rng.normal(1000, 10, 37))
    a, b = Nt[-30:-15], Nt[-15:]
    assert burnin.test_t_threshold(Nt, 29) == (ttest_rel(a, b)[1] > 0.05)      # odd burn_T is rounded up


class _FakeDev:
    def __init__(self, series):
        self.series = list(series)

    def burnin_cell_stats(self):
        return self.series.pop(0)


def test_spatial_tester_logic():
    rng = np.random.default_rng(0)
    noise = [(rng.normal(0, 1e-3), 0.3 + rng.normal(0, 1e-2)) for _ in range(80)]
    st = burnin.SpatialTester(_FakeDev([(1.0, 1.2)] + noise))
    dev = _FakeDev(noise)
    for _ in range(5):
        st.update(dev)
    assert st.run_test(30) is False                 # too few samples: ADF cannot run (None -> not burned)
    for _ in range(60):
        st.update(dev)
    assert st.run_test(30) is True                  # white noise around a constant: stationary


@pytest.mark.gpu
def test_device_cell_stats_match_the_reference_tally():
    """gnx_burnin_cell_stats vs SpatialTester.update restated with collections.Counter (burnin.py:41-58),
    on a non-square landscape (the reference's [i, j] quirk included)."""
    from collections import Counter
    from geonomics_b200.device import DeviceSpecies
    rng = np.random.default_rng(9)
    from geonomics_b200 import workloads
    n = 3000
    cfg = dict(workloads.CONFIGS['c2'], dim=(24, 17), N=n, L=40, n_paths=50, n_traits=1, loci_per_trait=5)
    w = workloads.build(cfg, 3)
    X, Y = w['land_dim']
    dev = DeviceSpecies(w['land_dim'], w['rasters'], w['prm'], w['gen_arch'], capacity=3 * n, seed=5)
    try:
        dev.set_burn(True)
        dev.upload(rng.uniform(0, X - 1e-3, n), rng.uniform(0, Y - 1e-3, n), np.zeros(n, np.int32),
                   np.zeros(n, np.int8), np.arange(n))
        counts = np.zeros((X, Y))
        for rep in range(3):
            st = dev.download(genomes=False)
            tally = Counter((int(x), int(y)) for x, y in zip(st['x'], st['y']))
            new = np.zeros((X, Y))
            for i in range(X):
                for j in range(Y):
                    new[i, j] = tally.get((j, i), 0)
            diff = new - counts
            counts = new
            m, s = dev.burnin_cell_stats()
            assert abs(m - np.mean(diff)) < 1e-12 and abs(s - np.std(diff)) < 1e-10
            dev.step(2)
    finally:
        dev.close()
