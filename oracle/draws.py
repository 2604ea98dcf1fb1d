"""TEST INFRASTRUCTURE ONLY -- draw generation for free-running oracle steps.

Produces, with numpy's own samplers (the ones the reference calls),
one time step's worth of draws in the layout oracle.step_oracle.step() consumes.  Used by
bench.py's cpu_baseline / --impl reference legs and by the statistical tests."""
import numpy as np


def make_draws(rng, prm, n, cap, n_paths, max_tries=8):
    d = {}
    mu, kappa = prm.get('direction_mu', 0.0), prm.get('direction_kappa', 0.0)
    # numpy legacy vonmises(mu, 0) == pi*(2U-1): movement.py:55
    d['move_dir'] = rng.vonmises(mu, kappa, n) if kappa > 0 else rng.uniform(-np.pi, np.pi, n)
    kind, p1, p2 = prm.get('move_distr', ('wald', 1.0, 1.0))
    d['move_dist'] = _dist(rng, kind, p1, p2, n)
    d['mate_R'] = rng.integers(0, 1 << 32, n, dtype=np.uint64).astype(np.uint32)
    d['mate_u'] = rng.random(n)
    d['poisson'] = rng.poisson(prm['lam'], n)
    d['recomb_keys'] = rng.integers(0, n_paths, 2 * cap)
    d['start_homs'] = rng.integers(0, 2, (cap, 2))
    d['disp_dir'] = rng.uniform(-np.pi, np.pi, (cap, max_tries))
    kind, p1, p2 = prm.get('disp_distr', ('wald', 1.0, 1.0))
    d['disp_dist'] = _dist(rng, kind, p1, p2, (cap, max_tries))
    d['sex_u'] = rng.random(cap)
    d['sex_redraw_u'] = rng.random(cap)
    d['death_u'] = rng.random(n + cap)
    d['pan_u'] = rng.random(n)
    d['pan_R'] = rng.integers(0, 1 << 32, (n, 2), dtype=np.uint64).astype(np.uint32)
    return d


def _dist(rng, kind, p1, p2, size):
    if kind == 'wald':
        return rng.wald(p1, p2, size)
    if kind == 'lognormal':
        return rng.lognormal(p1, p2, size)
    z = rng.standard_normal(size)          # scipy levy.rvs(loc, scale) = loc + scale / Z^2
    return p1 + p2 / (z * z)
