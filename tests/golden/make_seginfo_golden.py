#!/usr/bin/env python
"""Golden vectors for the tskit segment bookkeeping (TEST INFRASTRUCTURE; build container only).
Drives the reference's own Recombinations class (structs/genome.py:47-281; pure numpy, no tskit
call) and records, for one set of cached recombination events, the breakpoints, the
(left, right) segment arrays and the output of `_get_seg_info` for sampled keys."""
import os
import sys
import warnings
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
warnings.filterwarnings('ignore')
from oracle import ref_shims   # noqa: E402


def main():
    ref_shims.install()
    from geonomics.structs.genome import Recombinations
    np.random.seed(77)
    L, n = 60, 200
    rec = Recombinations(L, None, n, 0.08, None, None, False)
    bps, subs = rec._draw_recombination_events(use_subsetters=True, nonneutral_loci=np.arange(L), use_tskit=True)
    rec._breakpoints = bps
    rec._set_seg_info()
    paths = np.array([[1 if list(subs[k])[2 * l + 1] else 0 for l in range(L)] for k in range(n)], dtype=np.uint8)
    out = dict(L=np.int64(L), paths=paths, rates=np.asarray(rec._rates))
    bp_ptr = [0]
    bp_pos = []
    for k in range(n):
        bp_pos.extend(int(v) for v in bps[k])
        bp_ptr.append(len(bp_pos))
    out['bp_ptr'] = np.array(bp_ptr, dtype=np.int64)
    out['bp_pos'] = np.array(bp_pos, dtype=np.int64)
    rng = np.random.default_rng(5)
    q_key, q_start, q_nodes, q_ptr, q_node, q_left, q_right = [], [], [], [0], [], [], []
    for _ in range(60):
        k = int(rng.integers(0, n))
        s = int(rng.integers(0, 2))
        nodes = rng.integers(0, 1000, 2)
        seg = [*rec._get_seg_info(start_homologue=s, event_key=k, node_ids=nodes)]
        q_key.append(k)
        q_start.append(s)
        q_nodes.append(nodes)
        for nd, le, ri in seg:
            q_node.append(int(nd))
            q_left.append(float(le))
            q_right.append(float(ri))
        q_ptr.append(len(q_node))
    out.update(q_key=np.array(q_key), q_start=np.array(q_start), q_nodes=np.array(q_nodes), q_ptr=np.array(q_ptr),
               q_node=np.array(q_node), q_left=np.array(q_left), q_right=np.array(q_right))
    np.savez_compressed(os.path.join(HERE, 'seginfo.npz'), **out)
    print('wrote seginfo.npz: %d events, %d breakpoints, %d sampled queries' % (n, len(bp_pos), len(q_key)))


if __name__ == '__main__':
    main()
