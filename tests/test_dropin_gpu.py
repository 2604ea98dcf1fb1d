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


def test_use_tskit_species_with_trait_mutation_on_the_device():
    """gen_arch.use_tskit = True (genotype ROWS per non-neutral locus, tskit tables filled per birth) with neutral,
    deleterious and trait mutation: the reference's own Model.walk drives the device; its TableCollection (the
    functional shim of oracle/ref_shims.py), gen_arch bookkeeping, subsetters and Individuals are kept as the
    reference's own code would leave them -- checked with the reference's own ops, and by handing the model back."""
    from geonomics_b200 import dropin
    def fewer_mutations(p):                                # ~1.3 per step: some steps have none
        g = p['comm']['species']['spp_0']['gen_arch']
        g['mu_neut'] *= 0.3
        g['mu_delet'] *= 0.3
        for tr in g['traits'].values():
            tr['mu'] *= 0.3
    gnx, mg, mod = _ref_model('tmut', tweak=fewer_mutations)
    spp, land = mod.comm[0], mod.land
    assert spp.gen_arch.use_tskit and spp.mutate and spp.burned
    _quiet_walk(mod, 2)                                    # a couple of reference steps first: tables already filled
    tc = spp._tc
    L = spp.gen_arch.L
    rows0 = (tc.nodes.num_rows, tc.edges.num_rows, tc.individuals.num_rows, tc.mutations.num_rows)
    n_mutables0 = len(spp.gen_arch._mutables)
    nt0 = len(spp.Nt)
    dev = dropin.attach(spp, land, seed=11, eager=True)
    try:
        steps = 0
        kinds = set()
        ga = spp.gen_arch
        seen = len(ga.nonneut_loci), len(ga.delet_loci)
        while True:
            before = len(ga._mutables)
            _quiet_walk(mod, 1)
            steps += 1
            used = before - len(ga._mutables)
            now = len(ga.nonneut_loci), len(ga.delet_loci)
            d_nn, d_dl = now[0] - seen[0], now[1] - seen[1]
            seen = now
            kinds |= ({'delet'} if d_dl else set()) | ({'trait'} if d_nn - d_dl else set()) | \
                ({'neut'} if used - d_nn else set())
            if n_mutables0 - len(ga._mutables) >= 10 and used == 0:
                break                                      # end on a step without mutation (see the z check below)
            assert steps < 100
        births = int(np.sum(spp.n_births[nt0:]))
        n_muts = n_mutables0 - len(ga._mutables)
        assert births > 500 and n_muts >= 10
        assert kinds == {'neut', 'delet', 'trait'}
        # tables: one individuals row, two nodes rows per birth; one mutations row per mutation
        assert tc.individuals.num_rows == rows0[2] + births
        assert tc.nodes.num_rows == rows0[0] + 2 * births
        assert tc.mutations.num_rows == rows0[3] + n_muts
        node_ind = np.array(tc.nodes.column('individual'))
        meta = tc.individuals.column('metadata')
        for ind in spp.values():
            n0_, n1_ = ind._nodes_tab_ids[0], ind._nodes_tab_ids[1]
            assert node_ind[n0_] == node_ind[n1_] == ind._individuals_tab_id
            assert int.from_bytes(meta[ind._individuals_tab_id], 'little') == ind.idx
        # the edges of every new node tile [0, L) (species.py:738-760 checks the same)
        child = np.array(tc.edges.column('child')[rows0[1]:])
        left = np.array(tc.edges.column('left')[rows0[1]:])
        right = np.array(tc.edges.column('right')[rows0[1]:])
        parent = np.array(tc.edges.column('parent')[rows0[1]:])
        assert set(np.unique(child)) == set(range(rows0[0], rows0[0] + 2 * births))
        span = np.zeros(2 * births)
        np.add.at(span, child - rows0[0], right - left)
        assert np.allclose(span, L)
        assert parent.min() >= 0 and np.all(parent < child)
        # mutations sit on nodes of individuals born in their step, at popped loci
        m_site = np.array(tc.mutations.column('site')[rows0[3]:])
        m_node = np.array(tc.mutations.column('node')[rows0[3]:])
        assert len(set(m_site)) == n_muts and np.all(m_node >= rows0[0])
        # genotype rows = non-neutral loci; bookkeeping arrays the reference's own ops index with
        g = np.stack([i.g for i in spp.values()])
        assert g.shape[1:] == (len(ga.nonneut_loci), 2)
        for k in range(ga.recombinations._n):
            assert len(ga.recombinations._subsetters[k]) == 2 * len(ga.nonneut_loci)
        from geonomics.ops import selection as rsel
        fit_ref = rsel._calc_fitness(spp)                  # traits via ind.z, deleterious loci via delet_loci_idxs
        fit_dev = np.array([i.fit for i in spp.values()])
        np.testing.assert_allclose(fit_dev, fit_ref, rtol=1e-6)
        newborn = [i for i in spp.values() if i.age == 0]
        assert len(newborn) > 20
        for trait_num in range(len(ga.traits)):
            for ind in newborn:                            # born under the current tables (no mutation this step)
                z_ref = rsel._calc_phenotype(ind, ga, trait_num)
                assert abs(ind.z[trait_num] - z_ref) <= 1e-12 * max(1.0, abs(z_ref))
        # hand the model back: the reference's own step runs on what the device left (subsetters, node ids, tables)
        dropin.detach(spp)
        _quiet_walk(mod, 2)
        assert len(spp) == spp.Nt[-1]
        assert tc.nodes.num_rows == rows0[0] + 2 * (births + int(np.sum(spp.n_births[-2:])))
    finally:
        dropin.detach(spp)
