"""CPU: the column store of the tskit hand-off (geonomics_b200/tables.py) and the host half of the drop-in for a
use_tskit=True reference species (rows <-> loci, subsetters vs paths)."""
import contextlib
import io
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.join(HERE, 'golden'))


def _rows(first_node, first_row, nb, L, T):
    rng = np.random.default_rng(nb)
    child = np.repeat(first_node + np.arange(2 * nb), 2)
    left = np.tile([0.0, 3.5], 2 * nb)
    right = np.tile([3.5, float(L)], 2 * nb)
    return dict(idx=1000 + np.arange(nb), x=rng.random(nb), y=rng.random(nb), z=rng.random((nb, T)),
                time=np.full(nb, -2.0), left=left, right=right, parent=rng.integers(0, first_node, 4 * nb),
                child=child, first_node_id=first_node, first_individual_row=first_row)


def test_table_columns_append_births_in_reference_row_order():
    from geonomics_b200.tables import TableColumns
    tc = TableColumns(10, 2)
    assert tc.n_loc == 5                                   # x, y, z0, z1, fit (species.py:694-697)
    tc.individuals.append_columns(flags=np.ones(3, np.int32), location=np.zeros((3, 5)), idx=np.arange(3))
    tc.nodes.append_columns(flags=np.ones(6, np.int32), time=np.ones(6), population=np.zeros(6, np.int32),
                            individual=np.repeat(np.arange(3, dtype=np.int32), 2))
    r = _rows(6, 3, 4, 10, 2)
    tc.append_births(r)
    assert tc.individuals.num_rows == 7 and tc.nodes.num_rows == 14 and tc.edges.num_rows == 16
    assert np.array_equal(tc.nodes.individual[6:], np.repeat(3 + np.arange(4), 2))
    assert np.all(tc.nodes.time[6:] == -2.0) and np.all(tc.nodes.flags == 1)
    assert np.array_equal(tc.individuals.idx[3:], r['idx'])
    assert np.allclose(tc.individuals.location[3:, 2:4], r['z']) and np.isnan(tc.individuals.location[3:, 4]).all()
    with pytest.raises(AssertionError):                    # rows out of step with the device counters
        tc.append_births(_rows(99, 7, 2, 10, 2))
    many = _rows(14, 7, 3000, 10, 2)                       # growth past the first allocation
    tc.append_births(many)
    assert tc.nodes.num_rows == 14 + 6000 and tc.edges.child[-1] == 14 + 5999


def test_to_tskit_needs_tskit():
    from geonomics_b200.api import Model
    if Model._have_tskit():
        pytest.skip('tskit is installed here')
    from geonomics_b200.tables import TableColumns
    saved = sys.modules.pop('tskit', None)
    try:
        with pytest.raises(RuntimeError):
            TableColumns(10, 0).to_tskit()
    finally:
        if saved is not None:
            sys.modules['tskit'] = saved


def test_rows_loci_roundtrip():
    from geonomics_b200 import genome_pack as gp
    rng = np.random.default_rng(0)
    nn = np.array([3, 17, 40, 41, 99])
    rows = rng.integers(0, 2, (20, 5, 2)).astype(np.int8)
    g = gp.rows_to_loci(rows, nn, 128)
    assert g.shape == (20, 128, 2) and g.sum() == rows.sum()
    assert np.array_equal(gp.loci_to_rows(g, nn), rows)


@pytest.mark.skipif(not os.path.isdir('/root/reference/geonomics'), reason='reference sources not present')
def test_dropin_host_half_for_a_use_tskit_reference_species():
    from oracle import ref_shims
    gnx = ref_shims.install()
    import make_golden as mg
    from geonomics_b200 import dropin
    p = mg.build_params(gnx, 'tmut')
    with contextlib.redirect_stdout(io.StringIO()):
        mod = gnx.make_model(p, name='tables_test')
        mod.walk(10000, 'burn', verbose=False)
        mod.walk(3, 'main', verbose=False)
    spp = mod.comm[0]
    ga = spp.gen_arch
    a = dropin.species_to_device_args(spp, mod.land)
    t = a['gen_arch']['tskit']
    nn = np.asarray(ga.nonneut_loci)
    assert a['gen_arch']['paths'].shape == (ga.recombinations._n, ga.L)
    assert t['subsetters'].shape == (ga.recombinations._n, len(nn))
    # rows present from the start read the path itself; rows inserted by mutations read the homologue in FRONT
    # of the locus (genome.py:133-160), i.e. the path one locus earlier
    paths = a['gen_arch']['paths']
    for c, l in enumerate(nn):
        same = np.array_equal(t['subsetters'][:, c], paths[:, l])
        prev = l > 0 and np.array_equal(t['subsetters'][:, c], paths[:, l - 1])
        assert same or prev, l
    q = dropin.population_arrays(spp)
    assert q['g'].shape == (len(spp), ga.L, 2) and q['node0'].shape == (len(spp),)
    rows = np.stack([np.asarray(i.g, dtype=np.int8) for i in spp.values()])
    from geonomics_b200 import genome_pack as gp
    assert np.array_equal(gp.loci_to_rows(q['g'], nn), rows)
    assert all(tr['loci_idxs'].shape == tr['loci'].shape for tr in a['gen_arch']['traits'])


def test_simplify_columns_keeps_the_samples_haplotypes():
    """tables.simplify_columns (restated Kelleher et al. 2018 simplify; tskit absent, parity unpinned): on a
    random pedigree with recombination and mutation the haplotypes of the sample nodes decoded from the tables
    are the same before and after, the samples come first in the order given, parents are older than children,
    no child has overlapping edges, and the tables shrink."""
    from geonomics_b200.tables import TableColumns, simplify_columns
    rng = np.random.default_rng(0)
    L, N, G = 50, 30, 12
    tc = TableColumns(L, 0)
    tc.sites.append_columns(position=np.arange(L, dtype=float), nonneutral=np.zeros(L, np.int8))
    n_ind = [0]

    def add_inds(k, t):
        first = tc.individuals.append_columns(flags=np.ones(k, np.int32), location=np.zeros((k, 2)),
                                              idx=np.arange(n_ind[0], n_ind[0] + k))
        tc.nodes.append_columns(flags=np.ones(2 * k, np.int32), time=np.full(2 * k, float(t)),
                                population=np.zeros(2 * k, np.int32),
                                individual=np.repeat(first + np.arange(k, dtype=np.int32), 2))
        n_ind[0] += k
        return first
    add_inds(N, 1)
    hap = {i: (rng.random(L) < 0.3).astype(np.int8) for i in range(2 * N)}
    for nd, h in hap.items():
        for s_ in np.flatnonzero(h):
            tc.mutations.append_columns(site=[s_], node=[nd], time=[np.nan])
    alive = list(range(N))
    for t in range(G):
        first = add_inds(N, -t)
        kids = []
        for b in range(N):
            child = first + b
            for hom in (0, 1):
                par = alive[rng.integers(len(alive))]
                bp = np.sort(rng.choice(np.arange(1, L), size=rng.integers(0, 3), replace=False))
                lefts, rights = np.r_[0, bp - 0.5], np.r_[bp - 0.5, L]
                st = rng.integers(2)
                cn = 2 * child + hom
                h = np.zeros(L, np.int8)
                for k, (a, bq) in enumerate(zip(lefts, rights)):
                    pn = 2 * par + ((k + st) % 2)
                    tc.edges.append_columns(left=[a], right=[bq], parent=[pn], child=[cn])
                    h[int(np.ceil(a)):int(np.ceil(bq))] = hap[pn][int(np.ceil(a)):int(np.ceil(bq))]
                if rng.random() < 0.2:
                    s_ = rng.integers(L)
                    if h[s_] == 0:
                        h[s_] = 1
                        tc.mutations.append_columns(site=[s_], node=[cn], time=[-t])
                hap[cn] = h
            kids.append(child)
        alive = kids
    samples = np.array([n for i in alive for n in (2 * i, 2 * i + 1)])
    sys.path.insert(0, HERE)
    from test_model_gpu import _haplotypes_from_tables
    loci = np.arange(L)
    before = _haplotypes_from_tables(tc, loci)[samples]
    assert all(np.array_equal(before[k], hap[int(s_)]) for k, s_ in enumerate(samples))
    sizes = (tc.nodes.num_rows, tc.edges.num_rows, tc.individuals.num_rows)
    kept_idx = tc.individuals.idx[tc.nodes.individual[samples]].copy()
    simplify_columns(tc, samples)
    after = _haplotypes_from_tables(tc, loci)[:len(samples)]
    assert np.array_equal(before, after)
    assert (tc.nodes.num_rows, tc.edges.num_rows, tc.individuals.num_rows) < sizes
    assert np.all(tc.nodes.time[tc.edges.parent] > tc.nodes.time[tc.edges.child])
    assert np.array_equal(tc.individuals.idx[tc.nodes.individual[:len(samples)]], kept_idx)
    assert np.all(tc.nodes.flags[:len(samples)] == 1) and not tc.nodes.flags[len(samples):].any()
    o = np.lexsort((tc.edges.left, tc.edges.child))
    c_, l_, r_ = tc.edges.child[o], tc.edges.left[o], tc.edges.right[o]
    same = c_[1:] == c_[:-1]
    assert np.all(l_[1:][same] >= r_[:-1][same])
