"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's tskit row generation.

Restates, in numpy, what Species._do_mating writes into the tskit TableCollection for every
offspring (structs/species.py:692-736) and how Recombinations turns a cached recombination
event into parental segments (structs/genome.py:234-281).  tskit itself is absent from the
build container and is never called here: the oracle produces the *rows* (columns of the
individuals / nodes / edges tables), which is also what the CUDA path produces.

Pinned against the reference's own Recombinations class (pure numpy, importable without
tskit) by tests/golden/make_seginfo_golden.py -> tests/golden/seginfo.npz.
What is NOT pinned: TableCollection.sort()/simplify() semantics (tskit is a third-party
dependency, unpinned in requirements.txt: tskit>=0.2.3) -- parity unpinned for that boundary.
"""
import numpy as np


def breakpoints_from_paths(paths):
    """genome.py:194-199: breakpoints = positions where a crossover was drawn = loci where the
    cached path (cumsum(events) % 2) switches homologue.  rate[0] == 0 (genome.py:183), so
    a path never switches at locus 0."""
    paths = np.asarray(paths, dtype=np.int8)
    prev = np.concatenate([np.zeros((paths.shape[0], 1), np.int8), paths[:, :-1]], axis=1)
    return [np.nonzero(row)[0] for row in (paths != prev)]


def seg_info(bp, L):
    """genome.py:234-254 `_set_seg_info`: left = [0] + (bp - 0.5), right = (bp - 0.5) + [L]."""
    bp = np.asarray(bp, dtype=np.float64)
    left = np.concatenate([[0.0], bp - 0.5])
    right = np.concatenate([bp - 0.5, [float(L)]])
    return left, right


def get_seg_info(bp, L, start_homologue, node_ids):
    """genome.py:257-281 `_get_seg_info` (jitter_breakpoints=False): segment i descends from
    node_ids[(i + start_homologue) % 2]."""
    left, right = seg_info(bp, L)
    nodes = np.asarray(node_ids)[[(i + start_homologue) % 2 for i in range(len(left))]]
    return nodes, left, right


def offspring_rows(pairs, nb, keys, start_homs, bps, L, node0, node1, first_node_id, first_individual_row,
                   x, y, z, idx, t):
    """species.py:692-736 for all offspring of one step, pairs outer / offspring inner.
    keys int[B, 2] (gamete_keys of step_oracle), start_homs int[B, 2], bps = breakpoints per
    cached path, node0/node1 = parents' node ids by ordinal.  Returns dict of columns."""
    pair_of = np.repeat(np.arange(len(nb)), nb)
    B = len(pair_of)
    left, right, parent, child = [], [], [], []
    for o in range(B):
        p = pair_of[o]
        for hom in (0, 1):
            par = pairs[p, hom]
            nodes, le, ri = get_seg_info(bps[keys[o, hom]], L, start_homs[o, hom], (node0[par], node1[par]))
            left.append(le)
            right.append(ri)
            parent.append(nodes)
            child.append(np.full(len(le), first_node_id + 2 * o + hom))
    cat = (lambda a, dt: np.concatenate(a).astype(dt) if a else np.zeros(0, dt))
    return dict(left=cat(left, np.float64), right=cat(right, np.float64), parent=cat(parent, np.int32),
                child=cat(child, np.int32), idx=np.asarray(idx), x=np.asarray(x), y=np.asarray(y),
                z=np.asarray(z), time=np.full(B, -float(t)),
                node_time=np.repeat(np.full(B, -float(t)), 2),
                node_individual=np.repeat(first_individual_row + np.arange(B), 2))
