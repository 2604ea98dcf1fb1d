"""Multi-GPU plumbing (torch.distributed; NCCL over NVLink on the GPU box, gloo in CPU tests).

The per-timestep loop shards by *independent replicate iterations* (`n_its`,
sim/model.py:115-117, 936-939: every iteration resets/deep-copies landscape and community,
so iterations never exchange data).  One process per GPU advances its share of the
iterations; there is NO data-path collective.  Collectives are used only for (i) the
benchmark's barrier / max-time / sum reductions and (ii) gathering the small per-step
trajectories (Nt, births, deaths, summary statistics) to rank 0 at the end.
"""
import os

import numpy as np


def world():
    return int(os.environ.get('RANK', '0')), int(os.environ.get('WORLD_SIZE', '1')), \
        int(os.environ.get('LOCAL_RANK', '0'))


def shard_iterations(n_its, rank, world_size):
    """Iterations owned by `rank`: round-robin over 0..n_its-1 (every iteration is run exactly
    once across the job, in ascending order within a rank)."""
    return list(range(rank, n_its, world_size))


def reduce_throughput(ms, units, dist=None, device='cpu'):
    """(max over ranks of elapsed ms, sum over ranks of units processed).  Device time is
    taken per rank with CUDA events; the job's time is the slowest rank's."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return float(ms), float(units)
    import torch
    t = torch.tensor([float(ms), float(units)], dtype=torch.float64, device=device)
    tmax = t.clone()
    dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    tsum = t.clone()
    dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
    return float(tmax[0]), float(tsum[1])


def gather_trajectories(local, dist=None):
    """Gather {iteration: {name: array}} dicts from every rank onto rank 0 (merged); other
    ranks get None."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return dict(local)
    out = [None] * dist.get_world_size() if dist.get_rank() == 0 else None
    dist.gather_object(local, out, dst=0)
    if dist.get_rank() != 0:
        return None
    merged = {}
    for part in out:
        for it, v in part.items():
            assert it not in merged, 'iteration %r ran on two ranks' % it
            merged[it] = v
    return dict(sorted(merged.items()))


def run_iterations(make_model, n_its, T, dist=None, burn=True, collect=None):
    """Run `n_its` independent iterations of a model, sharded one-iteration-group per rank
    (BASELINE config 3).  `make_model(it)` builds the Model for iteration `it` (its seed and,
    with rand_genarch, its genomic architecture vary by iteration: params.py:609-625).
    Returns on rank 0: {it: {'Nt': [...], 'n_births': [...], 'n_deaths': [...], **collect(mod)}}."""
    if dist is not None and dist.is_initialized():
        rank, ws = dist.get_rank(), dist.get_world_size()
    else:
        rank, ws = 0, 1
    local = {}
    for it in shard_iterations(n_its, rank, ws):
        mod = make_model(it)
        mod.it = it
        if burn:
            mod.walk(10 ** 9, 'burn')
        mod.walk(T, 'main')
        spp = mod.comm[0]
        rec = dict(Nt=np.array(spp.Nt[-T:]), n_births=np.array(spp.n_births[-T:]),
                   n_deaths=np.array(spp.n_deaths[-T:]))
        if collect is not None:
            rec.update(collect(mod))
        local[it] = rec
    return gather_trajectories(local, dist)
