"""Pins oracle/tskit_oracle.py to the reference's Recombinations segment bookkeeping
(tests/golden/seginfo.npz), CPU only."""
import os
import numpy as np
from oracle import tskit_oracle as to

Z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'seginfo.npz'))


def test_breakpoints_from_paths_match_reference():
    bps = to.breakpoints_from_paths(Z['paths'])
    ptr, pos = Z['bp_ptr'], Z['bp_pos']
    assert len(bps) == len(ptr) - 1
    for k, b in enumerate(bps):
        assert np.array_equal(b, pos[ptr[k]:ptr[k + 1]])
    assert Z['rates'][0] == 0 and all(0 not in b for b in bps)


def test_get_seg_info_matches_reference():
    bps = to.breakpoints_from_paths(Z['paths'])
    L = int(Z['L'])
    for q in range(len(Z['q_key'])):
        nodes, left, right = to.get_seg_info(bps[Z['q_key'][q]], L, int(Z['q_start'][q]), Z['q_nodes'][q])
        s, e = Z['q_ptr'][q], Z['q_ptr'][q + 1]
        assert np.array_equal(nodes, Z['q_node'][s:e])
        assert np.array_equal(left, Z['q_left'][s:e])
        assert np.array_equal(right, Z['q_right'][s:e])
        assert left[0] == 0 and right[-1] == L and np.all(right > left)
        assert np.isclose((right - left).sum(), L)           # species.py:738-763 check_haps invariant
