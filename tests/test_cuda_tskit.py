"""GPU: tskit record buffers (individual / node / edge rows per birth, species.py:692-736,
genome.py:234-281) against the reference-pinned oracle, plus the invariant the reference
itself checks (species.py:738-801): the edges of a child's node tile [0, L) and replaying
them over the parents' genotypes reproduces the child's genotype."""
import numpy as np
import pytest

from parity_util import synthetic_case, make_device

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('L,fixed', [(100, True), (1000, True), (60, False)])
def test_tskit_rows_match_oracle(L, fixed):
    from oracle import step_oracle as so
    from oracle import tskit_oracle as to
    arch, prm, state, draws = synthetic_case(L=L, n=1200, loci_per_trait=min(10, L // 4), seed=100 + L, max_tries=24,
                                             cap=3600)
    if not fixed:
        prm = dict(prm, n_births_fixed=False, lam=2)
        draws['poisson'] = np.random.default_rng(3).poisson(2, len(draws['mate_u']))
    n0 = len(state['x'])
    new_o, im_o = so.step(state, arch, prm, draws)
    B = im_o['B']
    dev = make_device(arch, prm, capacity=4 * n0, disp_tries=draws['disp_dist'].shape[1])
    try:
        dev.upload(state['x'], state['y'], state['age'], state['sex'], state['idx'], g=state['g'],
                   z=state['z'], max_ind_idx=state['max_ind_idx'])
        dev.tskit_enable(edge_capacity=200 * B + 1000, birth_capacity=2 * B + 10)
        rng = np.random.default_rng(9)
        perm = rng.permutation(2 * n0).astype(np.int32)          # msprime-style arbitrary node ids
        node0, node1 = perm[:n0] + 50, perm[n0:] + 50
        dev.tskit_set_nodes(node0, node1, next_node_id=2 * n0 + 50, next_individual_row=n0 + 7)
        d = dict(draws)
        d.pop('move_choice', None)
        d.pop('disp_choice', None)
        dev.set_draws(d)
        dev.step(1)
        dev.sync()
        rows = dev.tskit_drain()
        again = dev.tskit_drain()
        new = dev.download()
    finally:
        dev.close()
    assert len(again['left']) == 0 and len(again['idx']) == 0          # buffers were emptied
    bps = to.breakpoints_from_paths(arch['paths'])
    keys = so.gamete_keys(im_o['nb'], draws['recomb_keys'][:2 * B])
    pre = im_o['pre']
    exp = to.offspring_rows(im_o['pairs'], im_o['nb'], keys, draws['start_homs'][:B], bps, L, node0, node1,
                            2 * n0 + 50, n0 + 7, pre['x'][n0:], pre['y'][n0:], pre['z'][n0:], pre['idx'][n0:], t=0)
    assert rows['first_node_id'] == 2 * n0 + 50 and rows['first_individual_row'] == n0 + 7
    for k in ('left', 'right', 'parent', 'child', 'idx', 'node_time', 'node_individual'):
        assert np.array_equal(rows[k], exp[k]), k
    np.testing.assert_allclose(rows['x'], exp['x'], rtol=0, atol=1e-9)
    np.testing.assert_allclose(rows['y'], exp['y'], rtol=0, atol=1e-9)
    np.testing.assert_allclose(rows['z'], exp['z'], rtol=1e-12)
    # invariants the reference asserts (species.py:738-801)
    g_par = state['g']
    node_to = {}
    for i in range(n0):
        node_to[int(node0[i])] = (i, 0)
        node_to[int(node1[i])] = (i, 1)
    og = im_o['pre']['g'][n0:]
    for o in range(min(B, 150)):
        for hom in (0, 1):
            ch = rows['first_node_id'] + 2 * o + hom
            sel = rows['child'] == ch
            le, ri, pa = rows['left'][sel], rows['right'][sel], rows['parent'][sel]
            assert le[0] == 0 and ri[-1] == L and np.all(le[1:] == ri[:-1])
            hap = np.zeros(L, np.int8)
            for a, b, p in zip(le, ri, pa):
                i, h = node_to[int(p)]
                lo, hi = int(np.ceil(a)), int(np.ceil(b))
                hap[lo:hi] = g_par[i, lo:hi, h]
            assert np.array_equal(hap, og[o, :, hom])
    # survivors keep their node ids through the compaction
    assert len(new['idx']) == len(new_o['idx'])
