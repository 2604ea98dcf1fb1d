"""Strip domain decomposition against the undecomposed run: with the
same seed, a landscape cut into 2, 3 or 4 strips -- here as contexts of one process exchanging
through each other's device buffers, the same kernels that write over NVLink peer memory between
processes -- must hold, after every step, exactly the individuals of the single-context run:
same ids, positions, ages, genomes, phenotypes and fitness, same births / deaths / pairs."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _workload(surfaces=False, n=3000, dim=(48, 44), L=64, seed=3, sex=False, b=0.5):
    rng = np.random.default_rng(seed)
    X, Y = dim
    yy, xx = np.mgrid[0:Y, 0:X]
    lyr0 = 0.4 + 0.6 * (0.5 + 0.5 * np.cos(xx / 7.0) * np.sin(yy / 5.0)) if surfaces else np.ones((Y, X))
    rasters = np.stack([lyr0, np.tile(np.linspace(0, 1, X), (Y, 1)), np.tile(np.linspace(0, 1, Y)[:, None], (1, X))])
    loci = np.sort(rng.choice(L, 16, replace=False))
    traits = [dict(loci=loci[:8].astype(np.int64), alpha=rng.normal(0, 0.1, 8), phi=0.3, gamma=1.0, lyr_num=1,
                   univ_adv=False),
              dict(loci=loci[8:].astype(np.int64), alpha=rng.normal(0, 0.1, 8), phi=0.2, gamma=1.0, lyr_num=2,
                   univ_adv=False)]
    rates = np.full(L, 0.3)
    rates[0] = 0
    paths = (np.cumsum(rng.random((200, L)) < rates[None], axis=1) % 2).astype(np.uint8)
    ga = dict(L=L, paths=paths, traits=traits, dom=np.zeros(L, np.int8))
    prm = dict(b=b, R=0.5, lam=1, n_births_fixed=True, mating_radius=2.0, d_min=0.0, d_max=1.0, sex=sex,
               sex_ratio_p=0.5, max_age=None, K_layer=0, K_factor=n / float(lyr0.sum()), move=True,
               move_distr=('wald', 1.5, 1.0), disp_distr=('wald', 2.0, 1.0), direction_mu=0.0, direction_kappa=0.0,
               density_grid_window_width=None)
    if surfaces:
        prm['move_surf'] = dict(layer=0, mixture=True, kappa=12.0)
        prm['disp_surf'] = dict(layer=0, mixture=True, kappa=12.0)
    pop = dict(x=rng.uniform(0, X - 0.001, n), y=rng.uniform(0, Y - 0.001, n),
               age=rng.integers(0, 4, n).astype(np.int32), sex=rng.integers(0, 2, n).astype(np.int8),
               idx=np.arange(n, dtype=np.int64))
    g = (rng.random((n, L, 2)) < 0.5).astype(np.int8)
    return dim, rasters, prm, ga, pop, g


def _single(wl, steps, seed=77):
    from geonomics_b200.device import DeviceSpecies
    dim, rasters, prm, ga, pop, g = wl
    dev = DeviceSpecies(dim, rasters, prm, ga, capacity=4 * len(pop['x']), seed=seed)
    try:
        dev.upload(pop['x'], pop['y'], pop['age'], pop['sex'], pop['idx'], g=g)
        out = []
        for _ in range(steps):
            dev.step(1)
            dev.sync()
            out.append(dev.download())
        return out, dev.step_records()
    finally:
        dev.close()


def _strips(wl, world, steps, seed=77, bounds=None, device_barrier=False):
    from geonomics_b200.strips import LocalStrips
    dim, rasters, prm, ga, pop, g = wl
    st = LocalStrips(world, dim, rasters, prm, ga, capacity=4 * len(pop['x']), seed=seed, bounds=bounds,
                     device_barrier=device_barrier)
    try:
        st.upload(pop['x'], pop['y'], pop['age'], pop['sex'], pop['idx'], g=g)
        out = []
        for _ in range(steps):
            st.step(1)
            out.append(st.download())
        return out, st.step_records(), st.bounds
    finally:
        st.close()


def _same(a, b, t):
    assert np.array_equal(a['idx'], b['idx']), 'ids differ at step %d (%d vs %d individuals)' % (t, len(a['idx']), len(b['idx']))
    for k in ('x', 'y', 'age', 'sex', 'g', 'z', 'fit'):
        assert np.array_equal(a[k], b[k]), '%s differs at step %d' % (k, t)
    assert a['max_ind_idx'] == b['max_ind_idx']


@pytest.mark.parametrize('world', [2, 3, 4])
def test_strips_reproduce_the_undecomposed_run(world):
    wl = _workload()
    ref, ref_recs = _single(wl, 6)
    got, recs, bounds = _strips(wl, world, 6)
    assert len(bounds) == world + 1 and np.all(np.diff(bounds) >= 2)
    for t, (a, b) in enumerate(zip(ref, got)):
        _same(a, b, t)
    for ra, rb in zip(ref_recs, recs):
        assert (ra['Nt'], ra['n_births'], ra['n_deaths'], ra['n_pairs']) == (rb['Nt'], rb['n_births'], rb['n_deaths'],
                                                                             rb['n_pairs'])
    # the run exercised the exchanges: individuals crossed strip edges and pairs formed across them
    y0 = wl[4]['y']
    assert ref_recs[0]['n_births'] > 100 and len(ref[-1]['idx']) > 500
    del y0


@pytest.mark.timeout(600)
@pytest.mark.parametrize('world', [2, 4])
def test_strips_with_barriers_and_collectives_over_peer_memory(world):
    """gnx_strip_barrier: the barrier between the phases and the three small collectives (births all-gather,
    density-count sum, max(N)) as one-CTA kernels that write the peers' synchronisation pages and spin on their
    own -- no host round trip and no NCCL inside a step; still bit-identical to the undecomposed run."""
    wl = _workload(surfaces=True, seed=21)
    ref, ref_recs = _single(wl, 6, seed=13)
    got, recs, _ = _strips(wl, world, 6, seed=13, device_barrier=True)
    for t, (a, b) in enumerate(zip(ref, got)):
        _same(a, b, t)
    assert [(r['Nt'], r['n_births'], r['n_deaths']) for r in ref_recs] == \
        [(r['Nt'], r['n_births'], r['n_deaths']) for r in recs]


def test_strips_with_surfaces_sexes_and_uneven_rows():
    """On-the-fly conductance surfaces (the c4 movement / dispersal path), separate sexes (the
    female-male pair rule instead of the reciprocal one) and a deliberately lopsided cut."""
    wl = _workload(surfaces=True, sex=True, b=0.9, seed=11)
    ncy = int(44 / 2.0000002) + 1
    bounds = np.array([0, 2, 9, ncy], dtype=np.int32)
    ref, ref_recs = _single(wl, 5, seed=5)
    got, recs, _ = _strips(wl, 3, 5, seed=5, bounds=bounds)
    for t, (a, b) in enumerate(zip(ref, got)):
        _same(a, b, t)
    assert [r['n_births'] for r in ref_recs] == [r['n_births'] for r in recs]


def test_strips_burn_in_mode():
    """Burn-in (no genomes, no selection: species.py:825) under decomposition."""
    from geonomics_b200.device import DeviceSpecies
    from geonomics_b200.strips import LocalStrips
    dim, rasters, prm, ga, pop, g = _workload(seed=4)
    dev = DeviceSpecies(dim, rasters, prm, ga, capacity=4 * len(pop['x']), seed=9)
    st = LocalStrips(2, dim, rasters, prm, ga, capacity=4 * len(pop['x']), seed=9)
    try:
        dev.set_burn(True)
        st.set_burn(True)
        dev.upload(pop['x'], pop['y'], pop['age'], pop['sex'], pop['idx'])
        st.upload(pop['x'], pop['y'], pop['age'], pop['sex'], pop['idx'])
        dev.step(5)
        dev.sync()
        st.step(5)
        a, b = dev.download(genomes=False), st.download(genomes=False)
        assert np.array_equal(a['idx'], b['idx'])
        for k in ('x', 'y', 'age'):
            assert np.array_equal(a[k], b[k]), k
        ra, rb = dev.step_records(), st.step_records()
        assert [r['Nt'] for r in ra] == [r['Nt'] for r in rb]
    finally:
        dev.close()
        st.close()
