"""Host side of the tskit hand-off (species.py:440-445, 692-736, 956-1094, 1107-1219; mutation.py:44-58).

The reference adds one `tskit.TableCollection` row per Python call.  Here the step kernels write
the rows of every birth into device buffers (include/gnx_b200.h, "tskit record buffering"); this
module keeps them on the host as numpy COLUMNS -- the layout `TableCollection.*.append_columns` /
`set_columns` take -- and turns them into a real `tskit.TableCollection` when tskit is importable
(`to_tskit`).  tskit and msprime are absent from the build image: nothing here imports them at module
level, and `sort_and_simplify` says so instead of falling back to anything.
"""
import numpy as np


class _Columns:
    """Append-only numpy columns with amortised growth."""

    def __init__(self, **dtypes):
        self._dt = dtypes
        self._n = 0
        self._c = {k: np.zeros((0,) + tuple(sh), dtype=dt) for k, (dt, sh) in dtypes.items()}

    @property
    def num_rows(self):
        return self._n

    def __len__(self):
        return self._n

    def append_columns(self, **cols):
        n_new = None
        for k, v in cols.items():
            v = np.asarray(v)
            n_new = len(v) if n_new is None else n_new
            assert len(v) == n_new, 'ragged append'
        if not n_new:
            return self._n
        need = self._n + n_new
        for k, (dt, sh) in self._dt.items():
            a = self._c[k]
            if len(a) < need:
                b = np.zeros((max(need, 2 * len(a), 1024),) + tuple(sh), dtype=dt)
                b[:self._n] = a[:self._n]
                self._c[k] = a = b
            a[self._n:need] = cols[k] if k in cols else self._default(k)
        first = self._n
        self._n = need
        return first

    def _default(self, k):
        return -1 if np.issubdtype(self._dt[k][0], np.integer) else 0

    def column(self, k):
        return self._c[k][:self._n]

    def __getattr__(self, k):
        c = self.__dict__.get('_c')
        if c is not None and k in c:
            return c[k][:self._n]
        raise AttributeError(k)

    def truncate(self, n=0):
        self._n = n

    def asdict(self):
        return {k: self.column(k).copy() for k in self._c}


class TableColumns:
    """The five tables the path writes, as columns.

    nodes:       flags, time, population, individual
    edges:       left, right, parent, child
    individuals: flags, location [n, 2 + T + 1] = x, y, z..., fit (species.py:694-697), idx (the reference stores
                 it as 4 little-endian metadata bytes, species.py:703-704)
    sites:       position, nonneutral (metadata 't' / 'n', species.py:994-1002); ancestral_state is '0' for all
    mutations:   site, node, time; derived_state '1', parent -1 (mutation.py:47-57, genome.py:1146-1147)
    """

    def __init__(self, sequence_length, n_traits):
        self.sequence_length = int(sequence_length)
        self.n_loc = 2 + (n_traits + 1 if n_traits else 0)
        f8, i4, i8 = np.float64, np.int32, np.int64
        self.nodes = _Columns(flags=(i4, ()), time=(f8, ()), population=(i4, ()), individual=(i4, ()))
        self.edges = _Columns(left=(f8, ()), right=(f8, ()), parent=(i4, ()), child=(i4, ()))
        self.individuals = _Columns(flags=(i4, ()), location=(f8, (self.n_loc,)), idx=(i8, ()))
        self.sites = _Columns(position=(f8, ()), nonneutral=(np.int8, ()))
        self.mutations = _Columns(site=(i4, ()), node=(i4, ()), time=(f8, ()))

    def append_births(self, rows):
        """Rows drained from the device (DeviceSpecies.tskit_drain), one or several time steps of births
        (species.py:692-736: location = [x, y] + z + [fit], nodes flags = 1, time = -t, population = 0;
        rows['time'] is -t of each birth, t counted from the step recording was enabled in)."""
        nb = len(rows['idx'])
        if nb == 0:
            return
        if self.individuals.num_rows != rows['first_individual_row'] or self.nodes.num_rows != rows['first_node_id']:
            raise AssertionError('tables out of step with the device (rows %i / %i, device %i / %i)' % (
                self.individuals.num_rows, self.nodes.num_rows, rows['first_individual_row'], rows['first_node_id']))
        loc = np.full((nb, self.n_loc), np.nan)
        loc[:, 0], loc[:, 1] = rows['x'], rows['y']
        if self.n_loc > 2:
            loc[:, 2:-1] = rows['z']
        first = self.individuals.append_columns(flags=np.zeros(nb, np.int32), location=loc, idx=rows['idx'])
        time = np.repeat(np.asarray(rows['time'], dtype=np.float64), 2)
        self.nodes.append_columns(flags=np.ones(2 * nb, np.int32), time=time,
                                           population=np.zeros(2 * nb, np.int32),
                                           individual=np.repeat(first + np.arange(nb, dtype=np.int32), 2))
        self.edges.append_columns(left=rows['left'], right=rows['right'], parent=rows['parent'], child=rows['child'])

    def to_tskit(self):
        """A real tskit.TableCollection with these rows (needs tskit)."""
        try:
            import tskit
        except ImportError as e:           # pragma: no cover - tskit is absent from the build image
            raise RuntimeError('tskit is not installed: the tables stay numpy columns (TableColumns)') from e
        tc = tskit.TableCollection(sequence_length=self.sequence_length)
        n = self.individuals.num_rows
        meta = self.individuals.idx.astype('<u4').view(np.int8)
        tc.individuals.set_columns(flags=self.individuals.flags.astype(np.uint32),
                                   location=self.individuals.location.reshape(-1),
                                   location_offset=np.arange(n + 1, dtype=np.uint64) * self.n_loc,
                                   metadata=meta, metadata_offset=np.arange(n + 1, dtype=np.uint64) * 4)
        tc.nodes.set_columns(flags=self.nodes.flags.astype(np.uint32), time=self.nodes.time,
                             population=self.nodes.population, individual=self.nodes.individual)
        tc.edges.set_columns(left=self.edges.left, right=self.edges.right, parent=self.edges.parent,
                             child=self.edges.child)
        ns = self.sites.num_rows
        tc.sites.set_columns(position=self.sites.position, ancestral_state=np.full(ns, ord('0'), np.int8),
                             ancestral_state_offset=np.arange(ns + 1, dtype=np.uint64),
                             metadata=np.where(self.sites.nonneutral == 1, ord('t'), ord('n')).astype(np.int8),
                             metadata_offset=np.arange(ns + 1, dtype=np.uint64))
        nm = self.mutations.num_rows
        tc.mutations.set_columns(site=self.mutations.site, node=self.mutations.node, time=self.mutations.time,
                                 derived_state=np.full(nm, ord('1'), np.int8),
                                 derived_state_offset=np.arange(nm + 1, dtype=np.uint64))
        return tc

    def sort_and_simplify(self, sample_nodes):
        """species.py:1107-1152: TableCollection.sort(); simplify(current nodes, filter_individuals=True,
        filter_sites=False).  These are tskit's own algorithms: run on a real TableCollection and read back."""
        tc = self.to_tskit()
        tc.sort()
        tc.simplify(np.asarray(sample_nodes, dtype=np.int32), filter_individuals=True, filter_sites=False)
        self._load(tc)
        return tc

    def _load(self, tc):                    # pragma: no cover - needs tskit
        for tab in (self.nodes, self.edges, self.individuals, self.mutations):
            tab.truncate(0)
        self.nodes.append_columns(flags=tc.nodes.flags, time=tc.nodes.time, population=tc.nodes.population,
                                  individual=tc.nodes.individual)
        self.edges.append_columns(left=tc.edges.left, right=tc.edges.right, parent=tc.edges.parent,
                                  child=tc.edges.child)
        n = tc.individuals.num_rows
        loc = tc.individuals.location.reshape(n, self.n_loc) if n else np.zeros((0, self.n_loc))
        idx = tc.individuals.metadata.view('<u4').astype(np.int64) if n else np.zeros(0, np.int64)
        self.individuals.append_columns(flags=tc.individuals.flags, location=loc, idx=idx)
        self.mutations.append_columns(site=tc.mutations.site, node=tc.mutations.node, time=tc.mutations.time)
