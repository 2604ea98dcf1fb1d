"""The drop-in adapter (geonomics_b200/dropin.py) against the UNMODIFIED reference: everything
up to the device call -- parameter / architecture / population extraction from a real
`geonomics.Species` -- runs here on CPU whenever /root/reference exists (build container only;
the GPU box has no reference, so this test skips there)."""
import contextlib
import io
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
pytestmark = pytest.mark.skipif(not os.path.isdir('/root/reference/geonomics'),
                                reason='reference sources not present')


@pytest.fixture(scope='module')
def ref_model():
    sys.path.insert(0, os.path.join(HERE, 'golden'))
    from oracle import ref_shims
    gnx = ref_shims.install()
    import make_golden as mg
    p = mg.build_params(gnx, 'base')
    with contextlib.redirect_stdout(io.StringIO()):
        mod = gnx.make_model(p, name='dropin_test')
        mod.walk(10000, 'burn', verbose=False)
        mod.walk(2, 'main', verbose=False)
    return mg, mod


def test_species_to_device_args_matches_reference_state(ref_model):
    from geonomics_b200 import dropin
    mg, mod = ref_model
    spp, land = mod.comm[0], mod.land
    a = dropin.species_to_device_args(spp, land)
    ref = mg.capture_arch(spp, land)
    assert a['land_dim'] == tuple(land.dim)
    assert np.array_equal(a['rasters'], ref['rasters'])
    ga = a['gen_arch']
    assert ga['L'] == spp.gen_arch.L and np.array_equal(ga['paths'], ref['paths'])
    assert np.array_equal(ga['dom'], ref['dom'])
    assert len(ga['traits']) == int(ref['n_traits'])
    for t, tr in enumerate(ga['traits']):
        assert np.array_equal(tr['loci'], ref['trait%i_loci' % t])
        assert np.array_equal(tr['alpha'], ref['trait%i_alpha' % t])
        assert tr['phi'] == float(ref['trait%i_phi' % t]) and tr['gamma'] == float(ref['trait%i_gamma' % t])
    prm = a['prm']
    assert prm['b'] == spp.b and prm['R'] == spp.R and prm['mating_radius'] == spp.mating_radius
    assert prm['move_distr'][0] == 'wald' and prm['disp_distr'][0] == 'wald'
    assert prm['density_grid_window_width'] == float(ref['ww'])
    assert a['capacity'] >= 2 * len(spp)


def test_population_arrays_feed_the_oracle(ref_model):
    """The SoA the adapter uploads is a valid oracle state: one oracle step runs on it and keeps
    the invariants (ids ascending = species order; genotypes in {0, 1})."""
    from geonomics_b200 import dropin
    from oracle import step_oracle as so
    from oracle import draws as od
    mg, mod = ref_model
    spp, land = mod.comm[0], mod.land
    p = dropin.population_arrays(spp)
    st = mg.capture_state(spp)
    for k in ('x', 'y', 'age', 'idx', 'g'):
        assert np.array_equal(p[k], st[k])
    assert np.all(np.diff(p['idx']) > 0)
    a = dropin.species_to_device_args(spp, land)
    arch = dict(land_dim=a['land_dim'], rasters=a['rasters'], K=np.asarray(spp.K, dtype=np.float64),
                ww=a['prm']['density_grid_window_width'], traits=a['gen_arch']['traits'],
                dom=a['gen_arch']['dom'], paths=a['gen_arch']['paths'], move_surf=None, disp_surf=None)
    prm = dict(a['prm'], burn=False)
    n = len(p['x'])
    draws = od.make_draws(np.random.default_rng(1), prm, n, 2 * n + 64, len(arch['paths']), 24)
    state = dict(p, z=st['z'], max_ind_idx=int(spp.max_ind_idx))
    new, im = so.step(state, arch, prm, draws)
    assert len(new['x']) == n + im['B'] - im['n_deaths']
    assert set(np.unique(new['g'])) <= {0, 1}
