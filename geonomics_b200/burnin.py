"""Burn-in stationarity tests (sim/burnin.py, structs/community.py:107-131).

The reference decides that a Community has burned in when, after at least `burn_T` steps, three
tests pass for every species: an augmented Dickey-Fuller test and a paired t-test on the
population-size series `Nt`, and the same two tests on the mean and the standard deviation of the
per-landscape-cell change of the individual counts (`SpatialTester`).  Here the per-cell counts
and their change are accumulated on the device (`gnx_burnin_cell_stats`), so a burn-in step costs
two 8-byte integers of device->host traffic; the tests themselves are a few hundred numbers of
host arithmetic.

`adfuller` restates `statsmodels.tsa.stattools.adfuller(x)` with its defaults (regression 'c',
autolag 'AIC'), the only form the reference calls (burnin.py:75, 94): statsmodels is a dependency
of the reference that is absent from this image, so the restatement is pinned by its own
properties (tests/test_burnin.py: the 5 % / 10 % critical values of MacKinnon's table, the
continuity of the two p-value polynomials, stationary against unit-root series) and not against
statsmodels itself -- parity unpinned for this function.
"""
import numpy as np
from scipy.stats import norm, ttest_rel

# MacKinnon (1994) response-surface approximation of the Dickey-Fuller p-value, constant only, one
# series (statsmodels.tsa.adfvalues: tau_star_c[0], tau_min_c[0], tau_max_c[0], tau_c_smallp[0],
# tau_c_largep[0] after their scaling vectors)
_TAU_STAR, _TAU_MIN, _TAU_MAX = -1.61, -18.83, 2.74
_TAU_SMALLP = np.array([2.1659, 1.4412, 3.8269e-2])
_TAU_LARGEP = np.array([1.7339, 9.3202e-1, -1.2745e-1, -1.0368e-2])


def mackinnonp(teststat):
    if teststat > _TAU_MAX:
        return 1.0
    if teststat < _TAU_MIN:
        return 0.0
    coef = _TAU_SMALLP if teststat <= _TAU_STAR else _TAU_LARGEP
    return float(norm.cdf(np.polyval(coef[::-1], teststat)))


def _lagmat_in(xdiff, maxlag):
    """lagmat(xdiff[:, None], maxlag, trim='both', original='in'): row t = [dx_t, dx_{t-1}, ..., dx_{t-maxlag}]."""
    n = len(xdiff)
    return np.column_stack([xdiff[maxlag - k:n - k] for k in range(maxlag + 1)])


def _ols(y, X):
    """params, ssr, rank and (X'X)^-1 by the pseudo-inverse, as statsmodels' OLS.fit(method='pinv')."""
    pinv = np.linalg.pinv(X)
    params = pinv @ y
    resid = y - X @ params
    return params, float(resid @ resid), int(np.linalg.matrix_rank(X)), pinv @ pinv.T


def adfuller(x):
    """(adf statistic, p-value, used lag, nobs) of the augmented Dickey-Fuller unit-root test."""
    x = np.asarray(x, dtype=np.float64)
    if x.ndim != 1:
        raise ValueError('x must be one-dimensional')
    if x.max() == x.min():
        raise ValueError('Invalid input, x is constant')
    nobs = len(x)
    ntrend = 1
    maxlag = int(np.ceil(12.0 * np.power(nobs / 100.0, 1 / 4.0)))
    maxlag = min(nobs // 2 - ntrend - 1, maxlag)
    if maxlag < 0:
        raise ValueError('sample size is too short to use selected regression component')
    xdiff = np.diff(x)
    xdall = _lagmat_in(xdiff, maxlag)
    nobs = xdall.shape[0]
    xdall[:, 0] = x[-nobs - 1:-1]                     # the level replaces the unlagged difference
    xdshort = xdiff[-nobs:]
    # lag order by AIC over the common sample: columns [const, level, dx_{t-1}, ..., dx_{t-maxlag}]
    full = np.column_stack([np.ones(nobs), xdall])
    startlag = 2
    best = None
    for lag in range(startlag, startlag + maxlag + 1):
        _, ssr, rank, _ = _ols(xdshort, full[:, :lag])
        llf = -nobs / 2.0 * np.log(2 * np.pi) - nobs / 2.0 * np.log(ssr / nobs) - nobs / 2.0
        aic = -2.0 * llf + 2.0 * rank
        if best is None or (aic, lag) < best:
            best = (aic, lag)
    usedlag = best[1] - startlag
    xdall = _lagmat_in(xdiff, usedlag)
    nobs = xdall.shape[0]
    xdall[:, 0] = x[-nobs - 1:-1]
    xdshort = xdiff[-nobs:]
    X = np.column_stack([xdall[:, :usedlag + 1], np.ones(nobs)])
    params, ssr, rank, xtx_inv = _ols(xdshort, X)
    scale = ssr / (nobs - rank)
    adfstat = params[0] / np.sqrt(scale * xtx_inv[0, 0])
    return float(adfstat), mackinnonp(adfstat), usedlag, nobs


def test_adf_threshold(Nt, num_timesteps_back, alpha=0.05):
    """burnin.py:93-95."""
    return adfuller(Nt[-num_timesteps_back:])[1] < alpha


def test_t_threshold(Nt, num_timesteps_back, alpha=0.05):
    """burnin.py:98-103."""
    num_timesteps_back += num_timesteps_back % 2
    return ttest_rel(Nt[int(-num_timesteps_back):int(-num_timesteps_back / 2)],
                     Nt[int(-num_timesteps_back / 2):])[1] > alpha


test_adf_threshold.__test__ = False          # not pytest tests, whatever their names
test_t_threshold.__test__ = False


class SpatialTester:
    """burnin.py:21-90 with the counts kept on the device: `stats` holds, per update, the mean and
    the standard deviation of the change of every landscape cell's individual count."""

    def __init__(self, dev):
        self.stats = {'mean': [], 'std': []}
        self.update(dev)

    def update(self, dev):
        m, s = dev.burnin_cell_stats()
        self.stats['mean'].append(m)
        self.stats['std'].append(s)

    def run_test(self, num_timesteps_back, alpha=0.05):
        results = []
        for data in self.stats.values():
            try:
                adf_res = adfuller(data[-num_timesteps_back:])[1] < alpha
            except ValueError:
                adf_res = None
            try:
                ttest_res = ttest_rel(data[int(-num_timesteps_back):int(-num_timesteps_back / 2)],
                                      data[int(-num_timesteps_back / 2):])[1] > alpha
            except ValueError:
                ttest_res = None
            results.append(adf_res and ttest_res)
        return bool(np.all(results))
