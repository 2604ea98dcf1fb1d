"""Host side of the strip decomposition (geonomics_b200/strips.py): the row plan, ownership and
the merge of the ranks' shares -- on CPU, including a world-size-2 gloo run in which every rank
derives its own share of one population and the shares are checked to partition it."""
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from geonomics_b200 import strips  # noqa: E402


def test_mating_grid_matches_the_library_rule():
    cs, ncx, ncy = strips.mating_grid((4096, 4096), 2.0)
    assert ncx == ncy == 2048 and abs(cs - 2.0000002) < 1e-9
    cs, ncx, ncy = strips.mating_grid((8192, 8192), 1.0)          # > 2^22 cells: the side doubles
    assert ncx * ncy <= (1 << 22) and cs > 1.9
    cs, ncx, ncy = strips.mating_grid((48, 44), 2.0)
    assert (ncx, ncy) == (24, 22)


@pytest.mark.parametrize('world', [1, 2, 3, 8])
def test_plan_rows_is_a_partition_with_two_rows_each(world):
    b = strips.plan_rows(64, world)
    assert b[0] == 0 and b[-1] == 64 and len(b) == world + 1 and np.all(np.diff(b) >= 2)
    # weights concentrate the cuts where the load is
    w = np.zeros(64)
    w[40:48] = 1.0
    bw = strips.plan_rows(64, world, w)
    assert bw[0] == 0 and bw[-1] == 64 and np.all(np.diff(bw) >= 2)
    if world == 2:
        assert 42 <= bw[1] <= 46
    with pytest.raises(AssertionError):
        strips.plan_rows(5, 3)


def test_owner_of_follows_the_kernel_row_rule():
    cs, ncx, ncy = strips.mating_grid((48, 44), 2.0)
    bounds = strips.plan_rows(ncy, 3)
    y = np.array([0.0, cs * bounds[1] - 1e-9, cs * bounds[1], 43.999, cs * bounds[2]])
    own = strips.owner_of(y, bounds, cs, ncy)
    assert list(own) == [0, 0, 1, 2, 2]
    # the last landscape row can fall into a partial cell row: clamped to the last rank
    assert strips.owner_of(np.array([43.9995]), bounds, cs, ncy)[0] == 2


def test_merge_records_and_states():
    recs = strips.merge_records([[dict(t=0, Nt=5, n_births=2, n_deaths=1, n_pairs=2)],
                                 [dict(t=0, Nt=7, n_births=1, n_deaths=3, n_pairs=1)]])
    assert recs == [dict(t=0, Nt=12, n_births=3, n_deaths=4, n_pairs=3)]
    a = dict(idx=np.array([4, 9]), x=np.array([1.0, 2.0]), z=np.array([[1.0], [2.0]]), max_ind_idx=9)
    b = dict(idx=np.array([1, 7]), x=np.array([3.0, 4.0]), z=np.array([[3.0], [4.0]]), max_ind_idx=11)
    m = strips.merge_states([a, b])
    assert list(m['idx']) == [1, 4, 7, 9] and list(m['x']) == [3.0, 1.0, 4.0, 2.0] and m['max_ind_idx'] == 11
    assert m['z'].shape == (4, 1)


def _gloo_worker(rank, world, port, out):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(5)                       # every rank holds the same host population
        y = rng.uniform(0, 43.999, 5000)
        cs, ncx, ncy = strips.mating_grid((48, 44), 2.0)
        K = np.ones((44, 48))
        bounds = strips.plan_rows(ncy, world, strips.row_weights_from_K(K, cs, ncy))
        mine = np.flatnonzero(strips.owner_of(y, bounds, cs, ncy) == rank)
        shares = [None] * world
        dist.all_gather_object(shares, (mine.tolist(), bounds.tolist()))
        if rank == 0:
            allidx = np.sort(np.concatenate([np.array(s[0], dtype=np.int64) for s in shares]))
            ok = np.array_equal(allidx, np.arange(5000)) and all(s[1] == shares[0][1] for s in shares)
            balanced = min(len(s[0]) for s in shares) > 0.35 * 5000 / world * 2 / 2
            open(out, 'w').write('ok' if ok and balanced else 'bad %s' % [len(s[0]) for s in shares])
    finally:
        dist.destroy_process_group()


def test_two_ranks_partition_one_population(tmp_path):
    import socket
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    out = str(tmp_path / 'result.txt')
    mp.spawn(_gloo_worker, args=(2, port, out), nprocs=2, join=True)
    assert open(out).read() == 'ok'
