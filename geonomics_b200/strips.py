"""Strip domain decomposition of ONE landscape over several GPUs (SURVEY.md section 8e-2,
BASELINE configs[3]): host-side plan and orchestration above the C-ABI (gnx_strip_*).

Each rank owns a horizontal strip of the mating grid (whole rows, so that "owned" is one
contiguous range of the (mating cell, id) order and the rank-major order of pairs, births and
offspring ids is the order of the undecomposed run).  The data path is in the library:
individuals are written by the sender straight into the receiver's buffer (NVLink peer memory
mapped through CUDA IPC) and the device reads every size from device memory, so a time step
is enqueued without a host round trip.  This module supplies what the library leaves to the
caller: the plan (which rows, which individuals go where at upload), the barrier between
phases and the three small collectives (births all-gather, density-count sum, max(N)).

Two transports:
  * `NcclStrips`   one process per GPU (`torch.distributed`, backend nccl): the barrier is a
                   one-element all-reduce enqueued on the context's stream, the collectives act
                   on the context's own device buffers.
  * `LocalStrips`  all ranks as contexts of ONE process on one GPU (tests; a box with fewer GPUs
                   than ranks): the same kernels write into the other contexts' buffers, the
                   host runs each phase on every context before the next and does the
                   collectives with plain device copies.
"""
import ctypes as C

import numpy as np

from . import _lib
from .device import DeviceSpecies

N_PHASES = 8


# ---------------------------------------------------------------------------------------------
# the plan (pure host logic; CPU-tested)
# ---------------------------------------------------------------------------------------------
def mating_grid(land_dim, mating_radius):
    """Cell size and shape of the mating grid exactly as the library builds it (gnx_api.cu
    `mating_grid`): square cells of side >= radius * (1 + 1e-7), doubled until <= 2^22 cells."""
    cs = float(mating_radius) * 1.0000001
    while True:
        ncx, ncy = int(land_dim[0] / cs) + 1, int(land_dim[1] / cs) + 1
        if ncx * ncy <= (1 << 22):
            return cs, ncx, ncy
        cs *= 2.0


def plan_rows(ncy, world, weights=None):
    """First mating-grid row of every rank (length world + 1).  `weights` (per-row expected load,
    e.g. the row sums of the carrying-capacity raster) balance the strips; every strip gets at
    least two rows."""
    assert ncy >= 2 * world, 'the mating grid has too few rows for %d strips' % world
    if weights is None:
        weights = np.ones(ncy)
    w = np.maximum(np.asarray(weights, dtype=np.float64), 0) + 1e-12
    cum = np.concatenate([[0.0], np.cumsum(w)])
    bounds = [0]
    for r in range(1, world):
        target = cum[-1] * r / world
        row = int(np.searchsorted(cum, target))
        row = max(row, bounds[-1] + 2)
        row = min(row, ncy - 2 * (world - r))
        bounds.append(row)
    bounds.append(ncy)
    return np.array(bounds, dtype=np.int32)


def row_weights_from_K(K, cell_size, ncy):
    """Expected individuals per mating-grid row from a carrying-capacity raster [Y, X]."""
    rows = np.minimum((np.arange(K.shape[0]) / cell_size).astype(np.int64), ncy - 1)
    return np.bincount(rows, weights=K.sum(axis=1), minlength=ncy)


def owner_of(y, bounds, cell_size, ncy):
    """Rank that owns an individual at ordinate y (its mating-grid row, clamped like the kernel)."""
    row = np.minimum(np.floor(np.asarray(y) / cell_size).astype(np.int64), ncy - 1)
    return np.searchsorted(np.asarray(bounds)[1:-1], row, side='right')


def merge_records(per_rank):
    """Species-wide step records from the ranks' shares (Nt, births, deaths and pairs add up)."""
    out = []
    for recs in zip(*per_rank):
        out.append({'t': recs[0]['t'], 'Nt': sum(r['Nt'] for r in recs), 'n_births': sum(r['n_births'] for r in recs),
                    'n_deaths': sum(r['n_deaths'] for r in recs), 'n_pairs': sum(r['n_pairs'] for r in recs)})
    return out


def merge_states(states):
    """Species order (ascending id) from the ranks' downloads."""
    ids = np.concatenate([s['idx'] for s in states])
    order = np.argsort(ids, kind='stable')
    out = {}
    for k in states[0]:
        v0 = states[0][k]
        if isinstance(v0, np.ndarray) and v0.ndim >= 1 and len(v0) == len(states[0]['idx']):
            out[k] = np.concatenate([s[k] for s in states])[order]
    out['max_ind_idx'] = max(int(s['max_ind_idx']) for s in states)
    return out


# ---------------------------------------------------------------------------------------------
# one rank
# ---------------------------------------------------------------------------------------------
class _DevArray:
    """A ctx-owned device array as a __cuda_array_interface__ object (zero-copy torch view)."""

    def __init__(self, ptr, n, typestr):
        self.__cuda_array_interface__ = {'shape': (int(n),), 'typestr': typestr, 'data': (int(ptr), False),
                                         'version': 2, 'strides': None}


class StripRank:
    """One rank's DeviceSpecies with strip decomposition enabled."""

    def __init__(self, rank, world, bounds, land_dim, rasters, prm, gen_arch, capacity, seed,
                 migrant_capacity=None, halo_capacity=None, res_ratio=(1.0, 1.0)):
        self.rank, self.world = int(rank), int(world)
        self.bounds = np.ascontiguousarray(bounds, dtype=np.int32)
        self.dev = DeviceSpecies(land_dim, rasters, prm, gen_arch, capacity=capacity, seed=seed, res_ratio=res_ratio)
        L = self.dev._L
        sc = _lib.StripConfig()
        sc.rank, sc.world = self.rank, self.world
        sc.first_rows = self.bounds.ctypes.data_as(_lib.c_int32_p)
        sc.migrant_capacity = int(migrant_capacity or max(4096, capacity // 8))
        sc.halo_capacity = int(halo_capacity or max(4096, capacity // 4))
        _lib.check(L.gnx_strip_enable(self.dev._ctx, C.byref(sc)), 'gnx_strip_enable')
        self.endpoints = _lib.StripEndpoints()
        _lib.check(L.gnx_strip_endpoints(self.dev._ctx, C.byref(self.endpoints)), 'gnx_strip_endpoints')
        b, c, n, m = C.c_void_p(), C.c_void_p(), C.c_int64(), C.c_void_p()
        _lib.check(L.gnx_strip_collective_ptrs(self.dev._ctx, C.byref(b), C.byref(c), C.byref(n), C.byref(m)),
                   'gnx_strip_collective_ptrs')
        self.births_ptr, self.counts_ptr, self.n_counts, self.nmax_ptr = b.value, c.value, int(n.value), m.value

    def connect(self, peer_rank, endpoints, same_process):
        _lib.check(self.dev._L.gnx_strip_connect(self.dev._ctx, int(peer_rank), C.byref(endpoints),
                                                 int(bool(same_process))), 'gnx_strip_connect')

    def phase(self, k):
        _lib.check(self.dev._L.gnx_strip_phase(self.dev._ctx, int(k)), 'gnx_strip_phase')

    def barrier(self, kind=0):
        """gnx_strip_barrier: the barrier (0) or a collective (1 births, 2 counts, 3 max(N)) over peer memory."""
        _lib.check(self.dev._L.gnx_strip_barrier(self.dev._ctx, int(kind)), 'gnx_strip_barrier')

    def check(self):
        _lib.check(self.dev._L.gnx_strip_check(self.dev._ctx), 'gnx_strip_check')

    def tensors(self):
        import torch
        births = torch.as_tensor(_DevArray(self.births_ptr, self.world, '<i8'), device='cuda')
        counts = torch.as_tensor(_DevArray(self.counts_ptr, self.n_counts, '<i4'), device='cuda')
        nmax = torch.as_tensor(_DevArray(self.nmax_ptr, 1, '<i8'), device='cuda')      # bits of a double >= 0
        return births, counts, nmax


# what follows each phase: the barrier / collective kind of gnx_strip_barrier (phase 7 ends the step)
PHASE_SYNC = {0: 0, 1: 0, 2: 0, 3: 1, 4: 0, 5: 2, 6: 3}


def _subset(pop, mask):
    return {k: (v[mask] if isinstance(v, np.ndarray) and v.ndim >= 1 and len(v) == len(mask) else v)
            for k, v in pop.items()}


# ---------------------------------------------------------------------------------------------
# all ranks in one process (tests / fewer GPUs than ranks)
# ---------------------------------------------------------------------------------------------
class LocalStrips:
    """`world` strips as contexts of this process on the current GPU."""

    def __init__(self, world, land_dim, rasters, prm, gen_arch, capacity, seed=0, bounds=None, device_barrier=False,
                 **kw):
        import torch
        self.torch = torch
        self.world = int(world)
        # device_barrier: the ranks synchronise through gnx_strip_barrier (no host round trip inside a step);
        # otherwise the host is the barrier and does the three collectives on the ranks' tensors
        self.device_barrier = bool(device_barrier)
        self._warm = False
        self.cs, self.ncx, self.ncy = mating_grid(land_dim, prm['mating_radius'])
        K = np.asarray(rasters)[int(prm.get('K_layer', 0))] * float(prm.get('K_factor', 1.0))
        self.bounds = plan_rows(self.ncy, world, row_weights_from_K(K, self.cs, self.ncy)) if bounds is None \
            else np.asarray(bounds, dtype=np.int32)
        self.ranks = [StripRank(r, world, self.bounds, land_dim, rasters, prm, gen_arch, capacity, seed, **kw)
                      for r in range(world)]
        for a in self.ranks:
            for b in self.ranks:
                if a is not b:
                    a.connect(b.rank, b.endpoints, same_process=True)
        self._t = [r.tensors() for r in self.ranks]

    def set_burn(self, burn):
        for r in self.ranks:
            r.dev.set_burn(burn)

    def upload(self, x, y, age=None, sex=None, idx=None, g=None, genomes_packed=None, max_ind_idx=None):
        n = len(x)
        idx = np.arange(n, dtype=np.int64) if idx is None else np.asarray(idx)
        max_ind_idx = int(idx.max()) if (max_ind_idx is None and n) else max_ind_idx
        own = owner_of(y, self.bounds, self.cs, self.ncy)
        for r in self.ranks:
            m = own == r.rank
            r.dev.upload(x[m], y[m], None if age is None else age[m], None if sex is None else sex[m], idx[m],
                         g=None if g is None else g[m],
                         genomes_packed=None if genomes_packed is None else genomes_packed[m],
                         max_ind_idx=max_ind_idx)

    def _sync(self):
        for r in self.ranks:
            r.dev.sync()

    def step(self, n=1):
        torch = self.torch
        if self.device_barrier and self._warm:
            # whole steps through gnx_step (one CUDA graph per rank and step, spin barriers inside); one step of
            # every rank is enqueued before the next, so no rank's queue fills up while a peer has nothing queued
            for _ in range(n):
                for r in self.ranks:
                    r.dev.step(1)
            return
        if self.device_barrier:
            # The first step of a process loads every kernel (CUDA loads modules lazily, and a load may wait for
            # the device to drain): with all ranks enqueued by ONE host thread, rank 0's spinning barrier would
            # wait for a rank the blocked host has not enqueued yet.  One step with the host as the barrier first.
            # (One process per GPU has no such coupling: a rank's barrier kernel is queued before its host can
            # block on the next phase's kernels.)
            self._warm = True
            self.device_barrier = False
            try:
                self.step(1)
            finally:
                self.device_barrier = True
            return self.step(n - 1) if n > 1 else None
        for _ in range(n):
            for k in range(N_PHASES):
                for r in self.ranks:
                    r.phase(k)
                self._sync()                               # the barrier: every context finished phase k
                if k == 3:                                 # births of every rank, everywhere
                    b = torch.stack([t[0][r] for r, t in enumerate(self._t)])
                    for t in self._t:
                        t[0][:self.world] = b
                elif k == 5:                               # density counts summed over the strips
                    tot = sum(t[1] for t in self._t)
                    for t in self._t:
                        t[1].copy_(tot)
                elif k == 6:                               # max(N): non-negative doubles order like their bits
                    mx = torch.stack([t[2] for t in self._t]).max()
                    for t in self._t:
                        t[2].fill_(mx)
                if k in (3, 5, 6):
                    torch.cuda.synchronize()

    def step_records(self):
        return merge_records([r.dev.step_records() for r in self.ranks])

    def download(self, **kw):
        for r in self.ranks:
            r.check()
        return merge_states([r.dev.download(**kw) for r in self.ranks])

    def population_sizes(self):
        return [r.dev.population_size() for r in self.ranks]

    def close(self):
        for r in self.ranks:
            r.dev.close()


# ---------------------------------------------------------------------------------------------
# one process per GPU
# ---------------------------------------------------------------------------------------------
class NcclStrips:
    """This process's strip of a landscape decomposed over the ranks of `torch.distributed`
    (backend nccl, one GPU each, one node: the receive buffers are shared through CUDA IPC)."""

    def __init__(self, land_dim, rasters, prm, gen_arch, capacity, seed=0, bounds=None, group=None,
                 device_barrier=None, **kw):
        import os
        import torch
        import torch.distributed as dist
        self.torch, self.dist, self.group = torch, dist, group
        # the barriers and the three small collectives of a step run as one-CTA kernels over peer memory
        # (gnx_strip_barrier); GNX_STRIP_NCCL=1 (or device_barrier=False) keeps them on NCCL for comparison
        if device_barrier is None:
            device_barrier = os.environ.get('GNX_STRIP_NCCL', '0') in ('', '0')
        self.device_barrier = bool(device_barrier)
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.cs, self.ncx, self.ncy = mating_grid(land_dim, prm['mating_radius'])
        K = np.asarray(rasters)[int(prm.get('K_layer', 0))] * float(prm.get('K_factor', 1.0))
        self.bounds = plan_rows(self.ncy, self.world, row_weights_from_K(K, self.cs, self.ncy)) if bounds is None \
            else np.asarray(bounds, dtype=np.int32)
        self.me = StripRank(self.rank, self.world, self.bounds, land_dim, rasters, prm, gen_arch, capacity, seed, **kw)
        # exchange the endpoints (raw bytes of the struct: offsets + the CUDA IPC handle)
        mine = bytes(self.me.endpoints)
        everyone = [None] * self.world
        dist.all_gather_object(everyone, mine, group=group)
        for r, raw in enumerate(everyone):
            if r != self.rank:
                self.me.connect(r, _lib.StripEndpoints.from_buffer_copy(raw), same_process=False)
        self.stream = torch.cuda.ExternalStream(self.me.dev.stream_ptr)
        self.births, self.counts, self.nmax = self.me.tensors()
        self._token = torch.zeros(1, device='cuda', dtype=torch.int32)
        self._mine = torch.zeros(1, device='cuda', dtype=torch.int64)
        dist.barrier(group=group)

    @property
    def dev(self):
        return self.me.dev

    def set_burn(self, burn):
        self.me.dev.set_burn(burn)

    def upload_owned(self, x, y, age=None, sex=None, idx=None, g=None, genomes_packed=None, max_ind_idx=None):
        """Upload this rank's share of a population every rank holds on the host."""
        idx = np.arange(len(x), dtype=np.int64) if idx is None else np.asarray(idx)
        m = owner_of(y, self.bounds, self.cs, self.ncy) == self.rank
        self.me.dev.upload(x[m], y[m], None if age is None else age[m], None if sex is None else sex[m], idx[m],
                           g=None if g is None else g[m],
                           genomes_packed=None if genomes_packed is None else genomes_packed[m],
                           max_ind_idx=int(idx.max()) if max_ind_idx is None else max_ind_idx)

    def step(self, n=1):
        torch, dist = self.torch, self.dist
        if self.device_barrier:
            self.me.dev.step(n)                            # gnx_step: eight phases + gnx_strip_barrier, one graph per step
            return
        with torch.cuda.stream(self.stream):               # collectives are ordered on the context's stream
            for _ in range(n):
                for k in range(N_PHASES):
                    self.me.phase(k)
                    if k in (0, 1, 2, 4):
                        dist.all_reduce(self._token, group=self.group)            # stream-ordered barrier
                    elif k == 3:
                        self._mine.copy_(self.births[self.rank:self.rank + 1])
                        dist.all_gather_into_tensor(self.births[:self.world], self._mine, group=self.group)
                    elif k == 5:
                        dist.all_reduce(self.counts, op=dist.ReduceOp.SUM, group=self.group)
                    elif k == 6:
                        dist.all_reduce(self.nmax, op=dist.ReduceOp.MAX, group=self.group)

    def sync(self):
        self.me.check()

    def step_records_local(self):
        return self.me.dev.step_records()

    def close(self):
        self.me.dev.close()
