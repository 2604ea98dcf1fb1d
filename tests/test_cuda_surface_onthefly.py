"""GPU: the on-the-fly conductance-surface direction sampler (GNX_SURF_ONTHEFLY, the mode the
10M-individual target config runs) against the distribution the reference TABULATES per cell
(utils/spatial.py:365-461 `_make_conductance_surface`):

  * per-cell chi-square of the sampled directions against the analytic von Mises mixture (or
    unimodal von Mises) the builder draws from -- interior, edge and corner cells and a cell
    whose whole neighbourhood is zero (uniform weights / mean of all eight directions);
  * two-sample chi-square against the table the unmodified reference built for the same raster
    (tests/golden/surface_tables.npz, made by tests/golden/make_surface_golden.py);
  * exact quantisation: every step is (half(cos h), half(sin h)) * distance for a float16
    direction h in [-pi, pi] -- the values the reference gets from its float16 table
    (spatial.py:447, movement.py:75-76), with numpy's portable float16 cos/sin.

Both call sites are covered: movement (movement.py:45, k_move) and natal dispersal
(movement.py:105-108, k_newborns).
"""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
DIRS = np.array([-3 * np.pi / 4, -np.pi / 2, -np.pi / 4, np.pi, 0, 3 * np.pi / 4, np.pi / 2, np.pi / 4])
OFFS = [(-1, -1), (-1, 0), (-1, 1), (0, -1), (0, 1), (1, -1), (1, 0), (1, 1)]
NBINS = 48
CELLS = [(2, 3), (0, 3), (0, 0), (6, 5), (4, 1), (4, 0), (3, 5)]      # (row i, col j)
DIST = 0.25


def _golden():
    return np.load(os.path.join(HERE, 'golden', 'surface_tables.npz'))


def _neigh(rast, i, j):
    Y, X = rast.shape
    return np.array([rast[i + di, j + dj] if 0 <= i + di < Y and 0 <= j + dj < X else 0.0 for di, dj in OFFS])


def _expected_bins(rast, i, j, mixture, kappa):
    """Bin probabilities of the direction (wrapped onto [-pi, pi)) the builder samples for cell (i, j)."""
    from scipy.special import i0
    nv = _neigh(rast, i, j)
    if mixture:
        w = nv / nv.sum() if nv.sum() > 0 else np.full(8, 0.125)        # spatial.py:414-418
        locs = DIRS
    else:
        locs = np.array([DIRS[nv == nv.max()].mean()])                  # spatial.py:376-381
        w = np.array([1.0])
    m = 400 * NBINS
    th = -np.pi + (np.arange(m) + 0.5) * (2 * np.pi / m)
    pdf = sum(wk * np.exp(kappa * np.cos(th - lk)) for wk, lk in zip(w, locs)) / (2 * np.pi * i0(kappa))
    return pdf.reshape(NBINS, -1).sum(axis=1) * (2 * np.pi / m)


def _hist(theta):
    th = np.mod(np.asarray(theta, dtype=np.float64) + np.pi, 2 * np.pi) - np.pi
    return np.bincount(np.minimum((th + np.pi) / (2 * np.pi) * NBINS, NBINS - 1).astype(int), minlength=NBINS)


def _half_lookup():
    """(half cos, half sin) -> direction, for every float16 direction in [-pi, pi] with the oracle's
    (numpy-portable) float16 cos/sin semantics."""
    from oracle import step_oracle as so
    h = np.arange(65536, dtype=np.uint16).view(np.float16)
    h = h[np.isfinite(h) & (np.abs(h.astype(np.float64)) <= 3.1427)]
    c, s = so._cos_sin(h)
    return {(float(a), float(b)): float(d) for a, b, d in zip(c, s, h)}


def _directions_from_steps(dx, dy, lookup):
    c, s = dx / DIST, dy / DIST
    key = list(zip(c.tolist(), s.tolist()))
    missing = [k for k in key if k not in lookup]
    assert not missing, '%d of %d steps are not (half cos, half sin) of a float16 direction in [-pi, pi]: %r' % (
        len(missing), len(key), missing[:3])
    return np.array([lookup[k] for k in key])


def _device(rast, mixture, n_cap, kappa):
    from geonomics_b200.device import DeviceSpecies
    Y, X = rast.shape
    surf = dict(layer=0, mixture=mixture, kappa=kappa)          # no table -> GNX_SURF_ONTHEFLY
    prm = dict(b=1.0, R=0.5, lam=1, n_births_fixed=True, mating_radius=1.0, d_min=0, d_max=1, sex=False, K_layer=0,
               K_factor=1.0, move=True, move_surf=surf, disp_surf=surf)
    return DeviceSpecies((X, Y), rast[None].copy(), prm, None, capacity=n_cap, seed=97, disp_tries_injected=1)


def _chi2(obs, expected_p):
    from scipy.stats import chisquare
    n = obs.sum()
    e = expected_p / expected_p.sum() * n
    # merge low-expectation bins into one so the chi-square approximation holds
    low = e < 8
    o2 = np.concatenate([obs[~low], [obs[low].sum()]]) if low.any() else obs
    e2 = np.concatenate([e[~low], [e[low].sum()]]) if low.any() else e
    if e2[-1] < 1e-9 and o2[-1] == 0:
        o2, e2 = o2[:-1], e2[:-1]
    return chisquare(o2, e2 * o2.sum() / e2.sum()).pvalue


def _two_sample(a, b):
    from scipy.stats import chi2_contingency
    keep = (a + b) >= 10
    tab = np.stack([np.concatenate([a[keep], [a[~keep].sum()]]), np.concatenate([b[keep], [b[~keep].sum()]])])
    tab = tab[:, tab.sum(axis=0) > 0]
    return chi2_contingency(tab).pvalue


@pytest.mark.parametrize('mixture', [True, False])
def test_movement_directions_match_reference_distribution(mixture):
    z = _golden()
    rast, kappa = z['rast'], float(z['kappa'])
    table = z['table_mix' if mixture else 'table_uni']
    lookup = _half_lookup()
    M = 40000
    n = M * len(CELLS)
    x = np.concatenate([np.full(M, j + 0.5) for i, j in CELLS])
    y = np.concatenate([np.full(M, i + 0.5) for i, j in CELLS])
    dev = _device(rast, mixture, n + 64, kappa)
    try:
        dev.set_burn(True)
        dev.upload(x, y)
        dev.set_draws(dict(move_dist=np.full(n, DIST)))
        dev.stage('move')
        dev.sync()
        nx, ny = dev.read('X', n), dev.read('Y', n)
    finally:
        dev.close()
    for k, (i, j) in enumerate(CELLS):
        sl = slice(k * M, (k + 1) * M)
        theta = _directions_from_steps(nx[sl] - x[sl], ny[sl] - y[sl], lookup)
        assert np.abs(theta).max() <= 3.1427
        obs = _hist(theta)
        p = _chi2(obs, _expected_bins(rast, i, j, mixture, kappa))
        assert p > 1e-4, ('analytic', mixture, (i, j), p)
        p2 = _two_sample(obs, _hist(table[i, j].astype(np.float64)))
        assert p2 > 1e-4, ('reference table', mixture, (i, j), p2)


@pytest.mark.parametrize('mixture', [True, False])
def test_dispersal_directions_match_reference_distribution(mixture):
    """Natal dispersal (movement.py:98-141): M coincident parents in one cell mate among themselves
    (distance 0 <= radius), every offspring disperses from the common midpoint."""
    z = _golden()
    rast, kappa = z['rast'], float(z['kappa'])
    table = z['table_mix' if mixture else 'table_uni']
    lookup = _half_lookup()
    M = 6000
    for (i, j) in [(2, 3), (0, 0), (4, 1)]:
        x = np.full(M, j + 0.5)
        y = np.full(M, i + 0.5)
        dev = _device(rast, mixture, 3 * M, kappa)
        try:
            dev.set_burn(True)
            dev.upload(x, y)
            dev.set_draws(dict(disp_dist=np.full((2 * M, 1), DIST)))
            for st in ('bin_cells', 'find_mates', 'dedup_pairs', 'make_offspring'):
                dev.stage(st)
            dev.sync()
            c = dev.counters()
            B = c['B']
            assert B > M // 3
            ox, oy = dev.read('X', M + B)[M:], dev.read('Y', M + B)[M:]
        finally:
            dev.close()
        theta = _directions_from_steps(ox - (j + 0.5), oy - (i + 0.5), lookup)
        obs = _hist(theta)
        p = _chi2(obs, _expected_bins(rast, i, j, mixture, kappa))
        assert p > 1e-4, ('analytic', mixture, (i, j), p)
        p2 = _two_sample(obs, _hist(table[i, j].astype(np.float64)))
        assert p2 > 1e-4, ('reference table', mixture, (i, j), p2)
