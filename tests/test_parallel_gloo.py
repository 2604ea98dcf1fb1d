"""N>1 host logic on CPU: world_size-2 gloo process group (no GPU).  Covers the replicate
sharding, the benchmark's max-time / sum-units reduction and the trajectory gather."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world_size, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank),
                      WORLD_SIZE=str(world_size))
    dist.init_process_group('gloo', rank=rank, world_size=world_size)
    from geonomics_b200 import parallel
    its = parallel.shard_iterations(7, rank, world_size)
    ms, units = parallel.reduce_throughput(10.0 + 5 * rank, 1000 * (rank + 1), dist)

    class FakeSpp:
        def __init__(self, it):
            self.Nt = [100 + it] * 5
            self.n_births = [it] * 5
            self.n_deaths = [2 * it] * 5

    class FakeModel:
        def __init__(self, it):
            self.comm = {0: FakeSpp(it)}
            self.walked = []

        def walk(self, T, mode):
            self.walked.append((T, mode))
    traj = parallel.run_iterations(lambda it: FakeModel(it), 7, 3, dist,
                                   collect=lambda m: {'it_seen': np.array([m.it])})
    q.put((rank, its, ms, units, None if traj is None else {k: {n: v.tolist() for n, v in d.items()}
                                                              for k, d in traj.items()}))
    dist.barrier()
    dist.destroy_process_group()


def test_replicate_sharding_world_size_2():
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 500)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in procs])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, its0, ms0, u0, traj0), (r1, its1, ms1, u1, traj1) = res
    assert its0 == [0, 2, 4, 6] and its1 == [1, 3, 5]
    assert sorted(its0 + its1) == list(range(7))
    assert ms0 == ms1 == 15.0            # max over ranks
    assert u0 == u1 == 3000.0            # sum over ranks
    assert traj1 is None
    assert sorted(traj0) == list(range(7))
    for it, d in traj0.items():
        assert d['Nt'] == [100 + it] * 3 and d['n_births'] == [it] * 3 and d['it_seen'] == [it]


def test_single_process_paths():
    sys.path.insert(0, ROOT)
    from geonomics_b200 import parallel
    assert parallel.shard_iterations(5, 0, 1) == [0, 1, 2, 3, 4]
    assert parallel.reduce_throughput(3.0, 7.0) == (3.0, 7.0)
    assert parallel.gather_trajectories({1: {'a': np.zeros(2)}})[1]['a'].shape == (2,)
