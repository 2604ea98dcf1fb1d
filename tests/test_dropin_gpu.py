"""The drop-in on a GPU, against the UNMODIFIED reference (oracle/_ref on the GPU box, /root/reference in
the build container): a reference Model is built and burned in by the reference's own code on the CPU,
`dropin.attach` moves its Species to the device, and the reference's own `Model.walk` then drives the
device through the swapped queue entries (model.py:603-667)."""
import contextlib
import io
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.join(HERE, 'golden'))
from oracle import ref_shims  # noqa: E402

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(ref_shims.reference_root() is None, reason='reference package not installed')]


def _ref_model(case='base', burn=True, tweak=None):
    gnx = ref_shims.install()
    import make_golden as mg
    p = mg.build_params(gnx, case)
    if tweak:
        tweak(p)
    with contextlib.redirect_stdout(io.StringIO()):
        mod = gnx.make_model(p, name='dropin_gpu')
        if burn:
            mod.walk(10000, 'burn', verbose=False)
    return gnx, mg, mod


def _quiet_walk(mod, T, mode='main'):
    with contextlib.redirect_stdout(io.StringIO()):
        mod.walk(T, mode, verbose=False)


def test_reference_walk_drives_the_device_and_state_syncs_back():
    from geonomics_b200 import dropin
    gnx, mg, mod = _ref_model()
    spp, land = mod.comm[0], mod.land
    n0, nt0, t0 = len(spp), len(spp.Nt), mod.t
    dev = dropin.attach(spp, land, seed=5)
    try:
        launches0 = dev.launch_count
        _quiet_walk(mod, 20)
        assert mod.t == t0 + 20 and len(spp.Nt) == nt0 + 20
        assert dev.launch_count - launches0 >= 20 * 15          # the steps ran as device kernels
        N = np.array([n0] + spp.Nt[nt0:])
        nb, nd = np.array(spp.n_births[-20:]), np.array(spp.n_deaths[-20:])
        assert np.array_equal(N[1:], N[:-1] + nb - nd)
        dropin.sync_to_host(spp)
        assert len(spp) == spp.Nt[-1]
        ids = np.array(list(spp.keys()))
        assert np.all(np.diff(ids) > 0) and ids.max() <= spp.max_ind_idx
        # the reference's own ops, run on the synced Individuals, agree with what the device holds
        from geonomics.ops import selection as rsel
        fit_ref = rsel._calc_fitness(spp)
        fit_dev = np.array([i.fit for i in spp.values()])
        np.testing.assert_allclose(fit_dev, fit_ref, rtol=1e-6)
        for trait_num in range(len(spp.gen_arch.traits)):
            for ind in list(spp.values())[::37]:
                z_ref = rsel._calc_phenotype(ind, spp.gen_arch, trait_num)
                assert abs(ind.z[trait_num] - z_ref) <= 1e-12 * max(1.0, abs(z_ref))
        e = np.array([i.e for i in spp.values()])
        x, y = spp._get_x(), spp._get_y()
        for l in range(len(land)):
            assert np.array_equal(e[:, l], land[l].rast[np.int32(y), np.int32(x)])
        assert x.min() >= 0 and x.max() <= land.dim[0] - 0.001
        ages = np.array([i.age for i in spp.values()])
        assert ages.min() == 0 and set(np.unique(np.stack([i.g for i in spp.values()]))) <= {0, 1}
        # and the reference can take over again from the synced state
        dropin.detach(spp)
        _quiet_walk(mod, 2)
        assert len(spp.Nt) == nt0 + 22 and len(spp) == spp.Nt[-1]
    finally:
        dropin.detach(spp)


def test_injected_step_through_the_dropin_matches_the_oracle():
    from geonomics_b200 import dropin
    from oracle import step_oracle as so
    from oracle import draws as od
    gnx, mg, mod = _ref_model()
    spp, land = mod.comm[0], mod.land
    st = mg.capture_state(spp)
    a = dropin.species_to_device_args(spp, land)
    arch = dict(land_dim=a['land_dim'], rasters=a['rasters'], K=np.asarray(spp.K, dtype=np.float64),
                ww=a['prm']['density_grid_window_width'], traits=a['gen_arch']['traits'],
                dom=a['gen_arch']['dom'], paths=a['gen_arch']['paths'], move_surf=None, disp_surf=None)
    prm = dict(a['prm'], burn=False)
    n = len(st['x'])
    draws = od.make_draws(np.random.default_rng(17), prm, n, 2 * n + 64, len(arch['paths']), 12)
    state = dict(st, max_ind_idx=int(spp.max_ind_idx))
    new_o, im_o = so.step(state, arch, prm, draws)
    dev = dropin.attach(spp, land, seed=1, disp_tries_injected=12)
    try:
        dev.set_draws(draws)
        _quiet_walk(mod, 1)
        dropin.sync_to_host(spp)
        got = mg.capture_state(spp)
        assert spp.Nt[-1] == len(new_o['x']) and spp.n_births[-1] == im_o['B']
        for k in ('idx', 'age', 'g'):                            # integer / index work: bit-exact
            assert np.array_equal(got[k], new_o[k]), k
        for k in ('x', 'y'):                                     # cos/sin of the injected direction: 1e-9 abs
            np.testing.assert_allclose(got[k], new_o[k], rtol=0, atol=1e-9)
        np.testing.assert_allclose(got['z'], new_o['z'], rtol=1e-12)
        np.testing.assert_allclose(got['fit'], new_o['fit'], rtol=1e-6)
    finally:
        dropin.detach(spp)


def test_attach_before_burn_in_runs_the_whole_model_on_the_device():
    """attach at model creation.  The burn-in runs in device burn mode, the
    reference's own _set_genomes_and_tables assigns genomes on the host, and the wrapped method
    uploads them and switches selection (and mutation) on."""
    from geonomics_b200 import dropin

    def tweak(p):
        g = p['comm']['species']['spp_0']['gen_arch']
        g['mu_neut'] = 3e-5                                     # ~9 mutations in 15 steps, within the
        g['mu_delet'] = 0                                       # infinite-sites budget for T = 20
        p['model']['T'] = 20
    gnx, mg, mod = _ref_model(burn=False, tweak=tweak)
    spp, land = mod.comm[0], mod.land
    dev = dropin.attach(spp, land, seed=9, eager=True)          # eager: the burn-in tests read the individuals
    try:
        _quiet_walk(mod, 10000, 'burn')
        assert mod.comm.burned and spp.burned
        assert all(i.g is not None for i in spp.values())       # assigned by the reference after burn-in
        assert spp.mutate and spp.gen_arch._mutables is not None
        n_mutables0 = len(spp.gen_arch._mutables)
        _quiet_walk(mod, 15)
        assert len(spp) == spp.Nt[-1] > 0
        z = np.array([i.z for i in spp.values()])
        assert np.isfinite(z).all() and z.std() > 0             # selection-phase phenotypes from real genomes
        fit = np.array([i.fit for i in spp.values()])
        assert fit.min() > 0 and fit.max() <= 1 and fit.std() > 0
        assert len(spp.gen_arch._mutables) < n_mutables0         # neutral mutations consumed loci (mutation.py:62-86)
    finally:
        dropin.detach(spp)
