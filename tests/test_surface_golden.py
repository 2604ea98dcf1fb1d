"""CPU: the reference-built conductance-surface tables (tests/golden/surface_tables.npz, made by the
unmodified reference's `_make_conductance_surface`, utils/spatial.py:365-461) follow the analytic
von Mises mixture the GPU's on-the-fly sampler is tested against
(tests/test_cuda_surface_onthefly.py) -- i.e. the yardstick itself is pinned to the reference."""
import numpy as np

import test_cuda_surface_onthefly as sf


def test_reference_table_follows_the_analytic_mixture():
    z = sf._golden()
    rast, kappa = z['rast'], float(z['kappa'])
    for name, mix in (('table_mix', True), ('table_uni', False)):
        for (i, j) in sf.CELLS:
            p = sf._chi2(sf._hist(z[name][i, j].astype(np.float64)), sf._expected_bins(rast, i, j, mix, kappa))
            assert p > 1e-5, (name, (i, j), p)


def test_reference_table_is_float16_on_the_wrapped_circle():
    z = sf._golden()
    for name in ('table_mix', 'table_uni'):
        assert z[name].dtype == np.float16
        assert np.abs(z[name].astype(np.float64)).max() <= 3.1427      # scipy wraps rvs onto [-pi, pi)


def test_host_table_builder_matches_reference_distribution():
    """geonomics_b200.api._make_conductance_surface (TABLE mode input) against the reference's table."""
    from geonomics_b200.api import _make_conductance_surface
    z = sf._golden()
    rast, kappa = z['rast'], float(z['kappa'])
    for name, mix in (('table_mix', True), ('table_uni', False)):
        np.random.seed(5)
        tab = _make_conductance_surface(rast, mixture=mix, approx_len=6000, vm_distr_kappa=kappa)
        assert tab.dtype == np.float16 and np.abs(tab.astype(np.float64)).max() <= 3.1427
        for (i, j) in sf.CELLS:
            obs = sf._hist(tab[i, j].astype(np.float64))
            assert sf._chi2(obs, sf._expected_bins(rast, i, j, mix, kappa)) > 1e-5, (name, i, j)
            assert sf._two_sample(obs, sf._hist(z[name][i, j].astype(np.float64))) > 1e-5, (name, i, j)
