"""GPU tests of the reference-shaped API (geonomics_b200.api) running on libgnxb200.so:
make_model -> walk('burn') -> walk('main'), state views, environmental change, invariants,
and per-individual phenotype / fitness checked against the oracle formulas."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
PARAMS = os.path.join(HERE, 'data', 'params_small.py')


@pytest.fixture(scope='module')
def model():
    from geonomics_b200 import api
    mod = api.make_model(PARAMS)
    mod.walk(10000, 'burn')
    return mod


def test_burn_in_then_main(model):
    from oracle import step_oracle as so
    mod = model
    spp = mod.comm[0]
    assert mod.comm.burned and spp.burned
    assert len(spp.Nt) == mod.burn_t + 1 >= mod.burn_T
    n_burn = len(spp.Nt)
    # genomes were assigned after burn-in with the parameterised starting frequency
    g = mod.get_genotypes()
    assert g.shape == (len(spp), 60, 2)
    assert abs(g.mean() - 0.5) < 0.01
    rast1_before = mod.land[1].rast.copy()
    mod.walk(12, 'main')
    assert mod.t == 11 and len(spp.Nt) == n_burn + 12
    assert len(spp.n_births) == len(spp.Nt) == len(spp.n_deaths)
    N = np.array(spp.Nt)
    assert np.all(N[1:] == N[:-1] + np.array(spp.n_births[1:]) - np.array(spp.n_deaths[1:]))
    assert 0.3 * spp.K.sum() < N[-1] < 1.5 * spp.K.sum()
    # the scheduled environmental change ran (t = 4, 6, 8) on host and device rasters
    assert np.allclose(mod.land[1].rast, rast1_before[:, ::-1])
    e = mod.get_e()
    x, y = mod.get_x(), mod.get_y()
    assert np.array_equal(e[:, 1], mod.land[1].rast[y.astype(int), x.astype(int)])
    # views agree with each other and with the oracle formulas
    assert len(spp) == N[-1] == len(x)
    ids = np.array([i for i in spp])
    assert len(np.unique(ids)) == len(ids) and ids.max() <= spp.max_ind_idx
    assert np.all(np.diff(ids[ids >= 1200]) > 0)             # newborn ids ascend in species order
    g = mod.get_genotypes()
    t0 = spp.gen_arch.traits[0]
    z_o = so.phenotype(g, [dict(loci=t0.loci, alpha=t0.alpha)])
    np.testing.assert_allclose(mod.get_z(), z_o, rtol=1e-12)
    ind = spp[int(ids[3])]
    assert ind.x == x[3] and ind.age == mod.get_age()[3] and np.array_equal(ind.g, g[3])
    assert x.min() >= 0 and x.max() <= 40 - 0.001 and y.min() >= 0 and y.max() <= 40 - 0.001
    dens = spp._calc_density()
    assert dens.shape == (40, 40) and dens.min() >= 0
    # ages: everyone alive now aged by one per step since birth; newborns of the last step are 0
    assert mod.get_age().min() == 0


def test_walk_main_before_burn_raises():
    from geonomics_b200 import api
    mod = api.make_model(PARAMS)
    with pytest.raises(ValueError):
        mod.walk(1, 'main')


def test_seeded_models_are_reproducible():
    from geonomics_b200 import api
    outs = []
    for _ in range(2):
        mod = api.make_model(PARAMS)
        mod.walk(10000, 'burn')
        mod.walk(5, 'main')
        outs.append((list(mod.comm[0].Nt), mod.get_x().copy(), mod.get_genotypes().copy()))
    assert outs[0][0] == outs[1][0]
    assert np.array_equal(outs[0][1], outs[1][1])
    assert np.array_equal(outs[0][2], outs[1][2])


def test_on_device_stats_match_numpy(model):
    """sim/stats.py:399-435 formulas on the downloaded genotypes vs the device reductions."""
    mod = model
    spp = mod.comm[0]
    g = mod.get_genotypes()                       # [N, L, 2]
    N = g.shape[0]
    het_ref = np.sum(np.mean(g, axis=2) == 0.5, axis=0) / N
    f1 = np.sum(np.sum(g, axis=2), axis=0) / (2 * N)
    maf_ref = np.where(f1 > 0.5, 1 - f1, f1)
    np.testing.assert_allclose(spp._calc_het(), het_ref, rtol=0, atol=1e-15)
    np.testing.assert_allclose(spp._calc_maf(), maf_ref, rtol=0, atol=1e-15)
    np.testing.assert_allclose(spp._calc_allele_freqs(), f1, rtol=0, atol=1e-15)
    assert abs(spp._calc_het(mean=True) - het_ref.mean()) < 1e-15
    fit = mod.get_fitness()
    assert abs(spp._calc_mean_fitness() - fit.mean()) < 1e-12


def test_on_device_ld_matches_the_reference_formula(model):
    """sim/stats.py:359-392 _calc_ld, loop for loop, on the downloaded genotypes vs the device's
    chromosome-count accumulation (gnx_stats_ld); the reference's own function where it is installed."""
    mod = model
    spp = mod.comm[0]
    speciome = mod.get_genotypes()                # [N, L, 2]
    n, L = speciome.shape[0], speciome.shape[1]
    N = n * 2
    want = np.zeros([L] * 2) * np.nan
    with np.errstate(divide='ignore', invalid='ignore'):
        for i in range(L):
            for j in range(i + 1, L):
                f1_i = np.sum(speciome[:, i, :], axis=None) / (N)
                f1_j = np.sum(speciome[:, j, :], axis=None) / (N)
                f11_ij = float(np.sum(speciome[:, [i, j], :].sum(axis=1) == 2, axis=None)) / (N)
                D_1_1 = f11_ij - (f1_i * f1_j)
                r2 = (D_1_1 ** 2) / (f1_i * (1 - f1_i) * f1_j * (1 - f1_j))
                want[i, j] = want[j, i] = r2
    got = spp._calc_ld()
    assert got.shape == (L, L) and np.all(np.isnan(np.diagonal(got)))
    np.testing.assert_allclose(got, want, rtol=1e-12, atol=0, equal_nan=True)
    assert np.nanmax(got) <= 1 + 1e-12 and np.nanmin(got) >= 0
    from oracle import ref_shims
    if ref_shims.reference_root() is not None:
        ref_shims.install()
        from geonomics.sim import stats as ref_stats

        class _Spp:
            gen_arch = spp.gen_arch

            def _get_genotypes(self):
                return speciome
        with np.errstate(divide='ignore', invalid='ignore'):
            ref = ref_stats._calc_ld(_Spp())
        np.testing.assert_allclose(got, ref, rtol=1e-12, atol=0, equal_nan=True)


def test_on_device_ld_many_words():
    """L = 300 (10 word columns, 55 tiles) with loci in strong and in no linkage."""
    from geonomics_b200.device import DeviceSpecies
    from geonomics_b200 import genome_pack as gp
    rng = np.random.default_rng(8)
    n, L = 700, 300
    g = (rng.random((n, L, 2)) < 0.4).astype(np.int8)
    g[:, 37, :] = g[:, 290, :]                    # complete linkage across distant words
    g[:, 100, :] = 1 - g[:, 101, :]
    g[:, 5, :] = 0                                # a fixed locus: r^2 undefined (nan)
    from geonomics_b200 import workloads
    cfg = dict(workloads.CONFIGS['c2'], dim=(20, 20), N=n, L=L, n_paths=50)
    w = workloads.build(cfg, 3)
    X, Y = w['land_dim']
    dev = DeviceSpecies(w['land_dim'], w['rasters'], w['prm'], w['gen_arch'], capacity=3 * n, seed=1)
    try:
        dev.upload(rng.uniform(0, X - 1e-3, n), rng.uniform(0, Y - 1e-3, n), np.zeros(n, np.int32),
                   np.zeros(n, np.int8), np.arange(n), g=g)
        got = dev.ld()
    finally:
        dev.close()
    H = g.transpose(0, 2, 1).reshape(2 * n, L).astype(np.float64)
    n11 = H.T @ H
    f = np.diagonal(n11) / (2 * n)
    with np.errstate(divide='ignore', invalid='ignore'):
        D = n11 / (2 * n) - np.outer(f, f)
        want = D ** 2 / np.outer(f * (1 - f), f * (1 - f))
    want[np.arange(L), np.arange(L)] = np.nan
    np.testing.assert_allclose(got, want, rtol=1e-9, atol=1e-15, equal_nan=True)
    assert abs(got[37, 290] - 1) < 1e-12 and abs(got[100, 101] - 1) < 1e-12 and np.all(np.isnan(got[5, :]))


def test_model_with_mutation():
    """make_model -> burn -> main with mu_neut, mu_delet > 0 (a13): the device's mutation log
    is pulled into Species.mutations / GenomicArchitecture like the reference maintains them."""
    from geonomics_b200 import api
    p = api.read_parameters_file(PARAMS)
    g = p['comm']['species']['spp_0']['gen_arch']
    g['L'] = 600
    g['mu_neut'] = 1e-5
    g['mu_delet'] = 1e-5
    p['model']['T'] = 12
    mod = api.make_model(p)
    mod.walk(10000, 'burn')
    spp = mod.comm[0]
    ga = spp.gen_arch
    assert spp.mutate and len(ga._mutables) == 600 - len(ga.nonneut_loci)
    n_nonneut0 = len(ga.nonneut_loci)
    mutables0 = list(ga._mutables)
    mod.walk(12, 'main')
    births = sum(spp.n_births[-12:])
    expect = births * 600 * 2e-5
    assert expect > 10
    assert 0 < len(spp.mutations) and abs(len(spp.mutations) - expect) < 6 * np.sqrt(expect)
    assert [r['locus'] for r in spp.mutations] == mutables0[::-1][:len(spp.mutations)]
    n_del = sum(r['type'] == 'delet' for r in spp.mutations)
    assert len(ga.delet_loci) == n_del == len(ga.delet_loci_s)
    assert len(ga.nonneut_loci) == n_nonneut0 + n_del
    assert len(ga._mutables) == len(mutables0) - len(spp.mutations)
    # every logged individual was born in the step it was logged for
    assert all(r['individual'] <= spp.max_ind_idx for r in spp.mutations)
    # too many expected mutations for the genome: the reference's MutationRateError
    p2 = api.read_parameters_file(PARAMS)
    p2['comm']['species']['spp_0']['gen_arch']['mu_neut'] = 1e-3
    mod2 = api.make_model(p2)
    with pytest.raises(api.MutationRateError):
        mod2.walk(10000, 'burn')


def test_two_species_advance_independently():
    """species do not interact (community.py:37-43); each one
    owns a device context and its own stream.  (The reference's queue lambdas advance only the
    last species, model.py:613-656 -- consciously corrected here.)"""
    import copy
    from geonomics_b200 import api
    p = api.read_parameters_file(PARAMS)
    spp = p['comm']['species']
    spp['spp_1'] = copy.deepcopy(spp['spp_0'])
    spp['spp_1']['init']['N'] = int(spp['spp_0']['init']['N'] * 0.6)
    spp['spp_1']['mating']['b'] = 0.35
    mod = api.make_model(p)
    mod.walk(10000, 'burn')
    assert all(s.burned for s in mod.comm.values()) and mod.comm.burned
    a, b = mod.comm[0], mod.comm[1]
    na, nb = len(a.Nt), len(b.Nt)
    assert na == nb > 0
    mod.walk(6, 'main')
    for s in (a, b):
        assert len(s.Nt) == na + 6
        N = np.array(s.Nt)
        assert np.all(N[1:] == N[:-1] + np.array(s.n_births[1:]) - np.array(s.n_deaths[1:]))
        assert len(s) == s.Nt[-1] > 0
    # different parameters and RNG streams: the trajectories are not copies of each other
    assert list(a.Nt[-6:]) != list(b.Nt[-6:])
    assert mod.get_genotypes(0).shape[0] == len(a) and mod.get_genotypes(1).shape[0] == len(b)


def test_region_stats_and_fst_match_numpy(model):
    """Sub-population allele counts on the device and the pairwise Fst of the reference's
    validation suite (tests/validation/island/island_test.py:54-68) against numpy on the
    downloaded genotypes."""
    mod = model
    spp = mod.comm[0]
    g = mod.get_genotypes().astype(np.int64)      # [N, L, 2]
    x, y = mod.get_x(), mod.get_y()
    X, Y = mod.land.dim
    west, east = (0, X / 2, 0, Y), (X / 2, X, 0, Y)
    dev = spp._dev
    out = {}
    for name, (x0, x1, y0, y1) in (('w', west), ('e', east)):
        m = (x >= x0) & (x < x1) & (y >= y0) & (y < y1)
        st = dev.stats(region=(x0, x1, y0, y1))
        assert st['N'] == int(m.sum()) > 0
        f = g[m].sum(axis=(0, 2)) / (2.0 * m.sum())
        het = (g[m].sum(axis=2) == 1).sum(axis=0) / float(m.sum())
        np.testing.assert_allclose(st['freq'], f, rtol=0, atol=1e-15)
        np.testing.assert_allclose(st['het'], het, rtol=0, atol=1e-15)
        out[name] = (f, het)
    (f0, h0), (f1, h1) = out['w'], out['e']
    pbar = (f0 + f1) / 2
    Ht = 2 * pbar * (1 - pbar)
    with np.errstate(divide='ignore', invalid='ignore'):
        fst_ref = (Ht - (h0 + h1) / 2) / Ht
    fst_ref[f0 == f1] = np.nan
    np.testing.assert_allclose(dev.fst(west, east), fst_ref, rtol=1e-12, atol=1e-15, equal_nan=True)
    whole = dev.stats()
    assert whole['N'] == len(spp)


def _read_K(spp):
    X, Y = spp._land_dim
    return spp._dev.read('K_RAST', X * Y).reshape(Y, X)


def test_species_change_events_reach_the_device():
    """a17: demographic and life-history change events (change.py:612-742) through Model.walk:
    the device's K raster follows spp.K exactly, the changed birth probability shows in the pair
    counts, and the population tracks the new carrying capacity."""
    from geonomics_b200 import api
    p = api.read_parameters_file(PARAMS)
    spp_p = p['comm']['species'][next(iter(p['comm']['species']))]
    spp_p['change'] = {
        'dem': {0: dict(kind='custom', timesteps=[3, 12], sizes=[0.4, 1.0], start_t=None, end_t=None, rate=None,
                        interval=None, n_cycles=None, size_range=None, distr='uniform')},
        'life_hist': {'b': dict(timesteps=[20], vals=[0.6])},
    }
    for lp in p['landscape']['layers'].values():          # no landscape changer: K keeps the override
        lp.pop('change', None)
    mod = api.make_model(p)
    mod.walk(10000, 'burn')
    spp = mod.comm[0]
    K0 = spp.K.copy()
    assert spp._changer is not None and np.array_equal(_read_K(spp), K0)
    mod.walk(3, 'main')                                    # t = 0, 1, 2
    assert np.array_equal(spp.K, K0)
    N_before = spp.Nt[-1]
    mod.walk(1, 'main')                                    # the change scheduled for t = 3 runs after that step
    assert np.array_equal(spp.K, K0 * 0.4) and np.array_equal(_read_K(spp), K0 * 0.4)
    mod.walk(8, 'main')                                    # t = 4 .. 11 under 0.4 K
    assert spp.Nt[-1] < 0.75 * N_before
    mod.walk(1, 'main')                                    # t = 12: back to the base K
    assert np.array_equal(spp.K, K0) and np.array_equal(_read_K(spp), K0)
    mod.walk(8, 'main')                                    # t = 13 .. 20, one bulk call across no event
    assert spp.b == 0.6 and spp._dev.prm['b'] == 0.6       # life-history change at t = 20
    nb = np.array(spp.n_births, dtype=float)
    N = np.array(spp.Nt, dtype=float)
    mod.walk(6, 'main')
    nb, N = np.array(spp.n_births, dtype=float), np.array(spp.Nt, dtype=float)
    rate_before = (nb[-14:-6] / N[-15:-7]).mean()
    rate_after = (nb[-5:] / N[-6:-1]).mean()
    assert rate_after > 2.0 * rate_before                  # b went from 0.2 to 0.6
    assert mod.t == 26


def test_table_surface_follows_a_landscape_change():
    """a table-mode conductance surface over a layer that changes is re-built and
    re-uploaded (change.py:576-606), not left stale."""
    from geonomics_b200 import api
    rng = np.random.default_rng(5)
    dim = (12, 10)
    # conductance rises with y (a y-gradient: the reference's plain mean of the max-valued neighbours'
    # directions, spatial.py:376-381, is only a true mean direction when they do not straddle +-pi)
    lyr = api.Layer(np.tile(np.linspace(0.05, 1, dim[1])[:, None], (1, dim[0])), 'defined', 'cond', dim, idx=0)
    land = api.Landscape({0: lyr})
    spp_params = api.ParametersDict({
        'init': {'N': 300, 'K_layer': 'cond', 'K_factor': 3},
        'mating': dict(repro_age=0, sex=False, sex_ratio=1.0, R=0.5, b=0.2, n_births_distr_lambda=1,
                       n_births_fixed=True, mating_radius=2, choose_nearest_mate=False, inverse_dist_mating=False),
        'mortality': dict(max_age=None, d_min=0, d_max=1, density_grid_window_width=None),
        'movement': dict(move=True, direction_distr_mu=0, direction_distr_kappa=0,
                         movement_distance_distr_param1=0.5, movement_distance_distr_param2=1.0,
                         movement_distance_distr='wald', dispersal_distance_distr_param1=0.5,
                         dispersal_distance_distr_param2=1.0, dispersal_distance_distr='wald',
                         move_surf=dict(layer='cond', mixture=False, vm_distr_kappa=50, approx_len=64)),
    })
    np.random.seed(3)
    spp = api._make_species(land, 'spp_0', 0, spp_params, seed=11)
    spp._attach()
    try:
        # the non-mixture surface points up the gradient (direction ~ +pi/2)
        tab0 = spp._move_surf.surf.copy()
        assert abs(float(np.median(tab0[5, 5].astype(float))) - np.pi / 2) < 0.3
        spp._step(3)
        y_mean_up = spp._get_y().mean()
        land._set_raster(0, land[0].rast[::-1].copy())      # mirror: conductance now rises towards y = 0
        tab1 = spp._move_surf.surf
        assert abs(float(np.median(tab1[5, 5].astype(float))) + np.pi / 2) < 0.3
        assert np.array_equal(spp._dev._surf_tabs[0], tab1) and not np.array_equal(tab0, tab1)
        spp._step(6)
        assert spp._get_y().mean() < y_mean_up - 0.5        # the population now drifts the other way
    finally:
        spp._dev.close()
    del rng


@pytest.mark.parametrize('rand_genarch,repeat_burn,rand_comm', [(True, False, False), (False, True, False),
                                                                (False, False, True)])
def test_run_iterates_n_its(rand_genarch, repeat_burn, rand_comm):
    """Model.run (model.py:866-953) over n_its iterations with the reset rules of model.py:455-593:
    the burned-in community is the common start unless the burn-in is repeated or the community
    re-drawn; a re-drawn genomic architecture gets new genomes; the landscape changer starts over."""
    from geonomics_b200 import api
    p = api.read_parameters_file(PARAMS)
    p['model']['its'] = dict(n_its=3, rand_landscape=False, rand_comm=rand_comm, rand_genarch=rand_genarch,
                             repeat_burn=repeat_burn)
    p['model']['T'] = 10
    mod = api.make_model(p)
    land0 = mod.land[1].rast.copy()
    out = mod.run()
    assert sorted(out) == [0, 1, 2] and mod.it == 2 and mod.its == []
    for it in range(3):
        rec = out[it][0]
        assert len(rec['Nt']) == 10 and rec['Nt'].min() > 0
        N = rec['Nt']
        assert np.all(N[1:] == N[:-1] + rec['n_births'][1:] - rec['n_deaths'][1:])
    spp = mod.comm[0]
    assert spp.burned and mod.t == 9
    # every iteration saw the scheduled landscape change (t = 4..8) from the original raster again
    assert np.allclose(mod.land[1].rast, land0[:, ::-1])
    g = mod.get_genotypes()
    assert g.shape[0] == len(spp) and 0.3 < g.mean() < 0.7
    if not repeat_burn and not rand_comm:
        # same burned-in start every iteration: the first main step acts on the same individuals
        starts = [out[it][0]['Nt'][0] - out[it][0]['n_births'][0] + out[it][0]['n_deaths'][0] for it in range(3)]
        assert starts[0] == starts[1] == starts[2]
    if rand_genarch:
        # iterations drew different architectures: at least the trait loci or effect sizes moved
        assert mod.orig_comm[0]['gen_arch'] is not spp.gen_arch


def test_use_tskit_model_tables_replay_to_the_genotypes():
    """gen_arch.use_tskit = True through the host API (species.py:891-905, 956-1094, 692-736; mutation.py:44-58):
    the genotype views hold one ROW per non-neutral locus, the tables fill from the device row buffers over a
    multi-step walk, and replaying the edges (and mutations) from the starting nodes reproduces the genotype
    rows of everyone alive -- the invariant the reference checks through tskit (species.py:762-801)."""
    import copy
    from geonomics_b200 import api
    p = copy.deepcopy(api.read_parameters_file(PARAMS))
    g = p['comm']['species']['spp_0']['gen_arch']
    g.update(use_tskit=True, tskit_simp_interval=1000, L=200, r_distr_alpha=None, mu_neut=5e-6, mu_delet=8e-6,
             start_neut_zero=True)
    g['traits']['trait_0'].update(mu=8e-6, n_loci=6)
    p['model']['T'] = 25
    mod = api.make_model(p)
    mod.walk(10000, 'burn')
    spp = mod.comm[0]
    ga = spp.gen_arch
    nn0 = np.array(ga.nonneut_loci, dtype=np.int64)
    tc = spp._tc
    n0 = len(spp)
    assert tc.nodes.num_rows == 2 * n0 and tc.individuals.num_rows == n0 and tc.sites.num_rows == 200
    g0 = mod.get_genotypes()
    assert g0.shape == (n0, len(nn0), 2)                       # rows = non-neutral loci
    # haplotypes of the starting nodes at the starting non-neutral loci: node 2k + h = (individual k, homologue h)
    hap = {}
    for k in range(n0):
        hap[2 * k], hap[2 * k + 1] = g0[k, :, 0].copy(), g0[k, :, 1].copy()
    start_muts = tc.mutations.num_rows
    # starting mutations at the non-neutral sites are exactly the 1-alleles (genome.py:1137-1147)
    ms, mn = tc.mutations.site, tc.mutations.node
    col_of = {int(l): c for c, l in enumerate(nn0)}
    cnt = np.zeros((2 * n0, len(nn0)), np.int8)
    for s_, n_ in zip(ms, mn):
        if int(s_) in col_of:
            cnt[n_, col_of[int(s_)]] = 1
    assert np.array_equal(cnt, g0.transpose(0, 2, 1).reshape(2 * n0, len(nn0)))
    spp._tsk_steps = 7                                        # several drains of the device row buffers
    steps = 23
    mod.walk(steps, 'main')
    births = int(np.sum(spp.n_births[-steps:]))
    assert tc.nodes.num_rows == 2 * (n0 + births) and tc.individuals.num_rows == n0 + births
    assert len(spp.mutations) >= 3 and {'delet'} <= {r['type'] for r in spp.mutations}
    assert tc.mutations.num_rows == start_muts + len(spp.mutations)
    # node times: -t of the birth step, t = 0 .. steps - 1
    times = tc.nodes.time[2 * n0:]
    assert times.max() == 0 and times.min() == -(steps - 1) and np.all(np.diff(times) <= 0)
    # replay: every new node inherits, locus by locus, from the parent node whose edge covers the locus
    el, er, ep, ec = tc.edges.left, tc.edges.right, tc.edges.parent, tc.edges.child
    order = np.argsort(ec, kind='stable')
    el, er, ep, ec = el[order], er[order], ep[order], ec[order]
    lo = np.searchsorted(ec, np.arange(2 * n0, 2 * (n0 + births)), side='left')
    hi = np.searchsorted(ec, np.arange(2 * n0, 2 * (n0 + births)), side='right')
    pos = nn0.astype(np.float64)
    new_muts = {}
    for r in spp.mutations:
        new_muts.setdefault(int(r['node']), []).append(int(r['locus']))
    for j, node in enumerate(range(2 * n0, 2 * (n0 + births))):
        l_, r_, p_ = el[lo[j]:hi[j]], er[lo[j]:hi[j]], ep[lo[j]:hi[j]]
        assert l_[0] == 0 and r_[-1] == 200 and np.all(l_[1:] == r_[:-1])
        seg = np.searchsorted(r_, pos, side='right')          # locus l lies in [left, right)
        hap[node] = np.array([hap[int(p_[s])][c] for c, s in enumerate(seg)], dtype=np.int8)
    # everyone alive: rows of the STARTING non-neutral loci equal the replayed haplotypes of their two nodes
    nodes = spp._node_ids()
    g_now = mod.get_genotypes()
    nn_now = np.array(ga.nonneut_loci, dtype=np.int64)
    assert g_now.shape == (len(spp), len(nn_now), 2) and len(nn_now) > len(nn0)
    rows_of_start = np.searchsorted(nn_now, nn0)
    for k in range(len(spp)):
        for h in (0, 1):
            assert np.array_equal(g_now[k, rows_of_start, h], hap[int(nodes[k, h])]), (k, h)
    # rows added by non-neutral mutations: the 1-alleles descend from the mutated node only
    new_rows = np.setdiff1d(np.arange(len(nn_now)), rows_of_start)
    assert g_now[:, new_rows, :].sum() >= 0 and set(nn_now[new_rows]) == {r['locus'] for r in spp.mutations
                                                                           if r['type'] != 'neut'}
    # individuals table: metadata idx and node -> individual links of everyone alive
    ids = np.array([i for i in spp])
    ind_rows = tc.nodes.individual[nodes[:, 0]]
    assert np.array_equal(ind_rows, tc.nodes.individual[nodes[:, 1]])
    assert np.array_equal(tc.individuals.idx[ind_rows], ids)
    # trait bookkeeping came back (genome.py:416-437)
    t0 = ga.traits[0]
    assert len(t0.loci) == len(t0.alpha) == len(t0.loci_idxs) == t0.n_loci >= 6


def _haplotypes_from_tables(tc, loci):
    """Haplotype of every node at `loci`, from the tables alone: roots carry their mutations, every other node
    inherits locus by locus from the parent whose edge covers it, then adds its own mutations (what tskit's
    genotype decoding does; species.py:762-801 checks the individuals' arrays against it)."""
    n = tc.nodes.num_rows
    loci = np.asarray(loci, dtype=np.int64)
    col = {int(l): c for c, l in enumerate(loci)}
    H = np.zeros((n, len(loci)), np.int8)
    el, er, ep, ec = tc.edges.left, tc.edges.right, tc.edges.parent, tc.edges.child
    order = np.argsort(ec, kind='stable')
    el, er, ep, ec = el[order], er[order], ep[order], ec[order]
    lo = np.searchsorted(ec, np.arange(n), side='left')
    hi = np.searchsorted(ec, np.arange(n), side='right')
    muts = {}
    for s_, n_ in zip(tc.mutations.site, tc.mutations.node):
        if int(s_) in col:
            muts.setdefault(int(n_), []).append(col[int(s_)])
    for u in np.argsort(-tc.nodes.time, kind='stable'):        # oldest first: parents before children
        l_, r_, p_ = el[lo[u]:hi[u]], er[lo[u]:hi[u]], ep[lo[u]:hi[u]]
        if len(l_):
            o = np.argsort(l_)
            l_, r_, p_ = l_[o], r_[o], p_[o]
            seg = np.searchsorted(r_, loci.astype(np.float64), side='right')
            ok = (seg < len(l_))
            ok[ok] &= l_[seg[ok]] <= loci[ok]
            H[u, ok] = H[p_[seg[ok]], np.flatnonzero(ok)]
        for c in muts.get(int(u), []):
            H[u, c] = 1
    return H


def test_use_tskit_model_simplifies_at_its_interval():
    """model.py:756-768 + species.py:1107-1219: every tskit_simp_interval steps the tables are sorted and
    simplified on the current nodes (tskit where installed; here the restated algorithm of
    tables.simplify_columns) and the device continues from the simplified tables: nodes 2k, 2k + 1 in species
    order, next rows behind the retained ancestors.  The genotype rows of everyone alive must still decode from
    the tables alone."""
    import copy
    from geonomics_b200 import api
    p = copy.deepcopy(api.read_parameters_file(PARAMS))
    g = p['comm']['species']['spp_0']['gen_arch']
    g.update(use_tskit=True, tskit_simp_interval=5, L=200, r_distr_alpha=None, mu_neut=5e-6, mu_delet=8e-6,
             start_neut_zero=True)
    g['traits']['trait_0'].update(mu=8e-6, n_loci=6)
    p['model']['T'] = 25
    mod = api.make_model(p)
    mod.walk(10000, 'burn')
    spp = mod.comm[0]
    tc = spp._tc
    nn0 = np.array(spp.gen_arch.nonneut_loci, dtype=np.int64)
    n0 = len(spp)
    mod.walk(18, 'main')                                      # simplified after t = 4, 9, 14; three more steps since
    births = int(np.sum(spp.n_births[-18:]))
    n = len(spp)
    assert tc.nodes.num_rows < 2 * (n0 + births)              # history nobody descends from is gone
    assert tc.nodes.num_rows >= 2 * n and tc.individuals.num_rows >= n
    assert np.all(tc.nodes.time[tc.edges.parent] > tc.nodes.time[tc.edges.child])
    nodes = spp._node_ids()
    assert len(np.unique(nodes)) == 2 * n and nodes.max() < tc.nodes.num_rows
    # node -> individual -> idx links of everyone alive survive the renumbering
    ids = np.array([i for i in spp])
    rows = tc.nodes.individual[nodes[:, 0]]
    assert np.array_equal(rows, tc.nodes.individual[nodes[:, 1]])
    assert np.array_equal(tc.individuals.idx[rows], ids)
    # genotype rows at the starting non-neutral loci decode from the tables alone
    H = _haplotypes_from_tables(tc, nn0)
    g_now = mod.get_genotypes()
    rows_of_start = np.searchsorted(np.array(spp.gen_arch.nonneut_loci, dtype=np.int64), nn0)
    for h in (0, 1):
        assert np.array_equal(g_now[:, rows_of_start, h], H[nodes[:, h]])
    # and right after a simplification the samples are nodes 0 .. 2N - 1 in species order
    mod.walk(2, 'main')                                       # t = 19: (t + 1) % 5 == 0
    assert spp._tc_sorted_and_simplified
    n = len(spp)
    assert np.array_equal(spp._node_ids().reshape(-1), np.arange(2 * n))
