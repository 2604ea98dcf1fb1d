"""TEST INFRASTRUCTURE ONLY -- not part of the product path.

Import shims that let the *unmodified* reference package at /root/reference
(erthward/geonomics v1.4.9, pure Python) be imported in the build container,
where several of its import-time dependencies are absent (matplotlib, bitarray,
tskit, msprime, shapely, statsmodels, geopandas, rasterio, vcf).

Used by the golden-vector generators under tests/golden/ (build container), by the
drop-in tests that attach the GPU path to a real reference Species, and by
`bench.py --impl reference` (the unmodified reference timed on the host cores).  On the GPU
box the reference is the unmodified copy installed under oracle/_ref by build().
Nothing in geonomics_b200/ imports this file.

Why each shim is safe for the hot path (reference file:line):
  * matplotlib / geopandas / rasterio / vcf: plotting + file I/O only
    (species.py:37-46, utils/io.py:15-17, sim/data.py:20-22).
  * tskit / msprime: only reached with gen_arch.use_tskit=True (species.py:442, 978).
    `tskit.TableCollection` is a functional row store (add_row / columns / asdict /
    set_columns / clear -- what species.py:692-736, 956-1094 and mutation.py:44-58 call);
    `msprime.simulate(n, ...)` returns n current sample nodes and no ancestry (a forest of
    singletons: a valid tree sequence, and the path only reads the nodes' flags,
    species.py:1006-1009).  `sort()` / `simplify()` are the tskit C library's algorithms and
    are NOT shimmed: they raise, so a golden case must not reach the simplification cadence
    (model.py:756-768).
  * statsmodels adfuller: burn-in stationarity test only (burnin.py:17,94).
  * bitarray: pure container for recombination subsetters
    (genome.py:158-160, 220-224; mating.py:166-167) -- functional shim.
  * shapely Polygon: axis-aligned rectangle intersection areas at density-grid
    construction (spatial.py:300-313) -- functional shim.
"""
import sys
import types
from unittest.mock import MagicMock

import os

# the reference sources in the build container; on the GPU box (no /root/reference) the copy that
# __graft_entry__.build() pip-installed, unmodified, into the git-ignored oracle/_ref
REFERENCE_ROOT = '/root/reference'
INSTALLED_ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), '_ref')


def reference_root():
    if os.path.isdir(os.path.join(REFERENCE_ROOT, 'geonomics')):
        return REFERENCE_ROOT
    if os.path.isdir(os.path.join(INSTALLED_ROOT, 'geonomics')):
        return INSTALLED_ROOT
    return None


class _BitArray(list):
    def __init__(self, s=''):
        if isinstance(s, str):
            super().__init__(int(c) for c in s)
        else:
            super().__init__(s)

    def __add__(self, o):
        return _BitArray(list(self) + ([int(c) for c in o]
                                       if isinstance(o, str) else list(o)))

    def __getitem__(self, k):
        r = list.__getitem__(self, k)
        return _BitArray(r) if isinstance(k, slice) else r


class _RectPolygon:
    """Axis-aligned rectangles are all the reference ever builds."""

    def __init__(self, c):
        xs = [p[0] for p in c]
        ys = [p[1] for p in c]
        self.b = (min(xs), min(ys), max(xs), max(ys))

    def intersection(self, o):
        p = _RectPolygon.__new__(_RectPolygon)
        x0 = max(self.b[0], o.b[0])
        y0 = max(self.b[1], o.b[1])
        p.b = (x0, y0, max(x0, min(self.b[2], o.b[2])),
               max(y0, min(self.b[3], o.b[3])))
        return p

    @property
    def area(self):
        return (self.b[2] - self.b[0]) * (self.b[3] - self.b[1])


class _Table:
    """One tskit table as python lists per column; `add_row` returns the row id."""
    COLS = {}

    def __init__(self):
        self._c = {k: [] for k in self.COLS}

    def add_row(self, *a, **k):
        names = list(self.COLS)
        row = dict(self.COLS)
        for n, v in zip(names, a):
            row[n] = v
        for n, v in k.items():
            if n not in row:
                raise TypeError('%s.add_row: unknown column %r' % (type(self).__name__, n))
            row[n] = v
        for n in names:
            self._c[n].append(row[n])
        return len(self._c[names[0]]) - 1

    @property
    def num_rows(self):
        return len(self._c[next(iter(self.COLS))])

    def __len__(self):
        return self.num_rows

    def clear(self):
        for v in self._c.values():
            del v[:]

    def column(self, name):
        return self._c[name]

    def __getattr__(self, name):
        c = self.__dict__.get('_c')
        if c is not None and name in c:
            import numpy as np
            try:
                return np.array(c[name])
            except Exception:
                return list(c[name])
        raise AttributeError(name)

    def asdict(self):
        import numpy as np
        return {k: np.array(v) for k, v in self._c.items()}

    def set_columns(self, **cols):
        n = None
        for k, v in cols.items():
            if k in self._c:
                self._c[k] = list(v)
                n = len(self._c[k])
        for k in self._c:
            assert n is None or len(self._c[k]) == n, 'ragged table columns'


class _Nodes(_Table):
    COLS = dict(flags=0, time=0.0, population=-1, individual=-1, metadata=b'')


class _Edges(_Table):
    COLS = dict(left=0.0, right=0.0, parent=-1, child=-1)


class _Individuals(_Table):
    COLS = dict(flags=0, location=None, parents=None, metadata=b'')


class _Sites(_Table):
    COLS = dict(position=0.0, ancestral_state='', metadata=b'')


class _Mutations(_Table):
    COLS = dict(site=-1, node=-1, derived_state='', parent=-1, metadata=b'', time=float('nan'))


class _TableCollection:
    def __init__(self, sequence_length=0):
        self.sequence_length = sequence_length
        self.nodes = _Nodes()
        self.edges = _Edges()
        self.individuals = _Individuals()
        self.sites = _Sites()
        self.mutations = _Mutations()

    def sort(self, *a, **k):
        raise NotImplementedError('tskit.TableCollection.sort is not shimmed (tskit is absent)')

    def simplify(self, *a, **k):
        raise NotImplementedError('tskit.TableCollection.simplify is not shimmed (tskit is absent)')

    def tree_sequence(self):
        raise NotImplementedError('tskit.TableCollection.tree_sequence is not shimmed (tskit is absent)')


class _SimulatedAncestry:
    def __init__(self, n, length):
        self._n, self._length = int(n), length

    def dump_tables(self):
        tc = _TableCollection(self._length)
        for _ in range(self._n):
            tc.nodes.add_row(flags=1, time=0.0, population=0)
        return tc


def _functional_tskit():
    tk = MagicMock(name='tskit')
    tk.TableCollection = _TableCollection
    tk.NULL = -1
    tk.UNKNOWN_TIME = float('nan')
    ms = MagicMock(name='msprime')
    ms.simulate = lambda sample_size, Ne=None, length=None, **k: _SimulatedAncestry(sample_size, length)
    return tk, ms


def install():
    """Install the shims, then make `import geonomics` resolve to the reference."""
    for name in ['matplotlib', 'matplotlib.pyplot', 'matplotlib.colors',
                 'matplotlib.animation', 'matplotlib.gridspec',
                 'matplotlib.ticker', 'matplotlib.lines', 'mpl_toolkits',
                 'mpl_toolkits.mplot3d', 'mpl_toolkits.mplot3d.axes3d',
                 'mpl_toolkits.axes_grid1', 'geopandas', 'rasterio',
                 'statsmodels', 'statsmodels.api',
                 'statsmodels.tsa', 'vcf', 'nlmpy']:
        if name not in sys.modules:
            sys.modules[name] = MagicMock(name=name)
    if not isinstance(getattr(sys.modules.get('tskit'), 'TableCollection', None), type):
        sys.modules['tskit'], sys.modules['msprime'] = _functional_tskit()
    st = types.ModuleType('statsmodels.tsa.stattools')
    st.adfuller = lambda x, *a, **k: (0.0, 0.0)
    sys.modules['statsmodels.tsa.stattools'] = st
    ba = types.ModuleType('bitarray')
    ba.bitarray = _BitArray
    sys.modules['bitarray'] = ba
    sh = types.ModuleType('shapely')
    geo = types.ModuleType('shapely.geometry')
    geo.Polygon = _RectPolygon
    geo.Point = MagicMock()
    sh.geometry = geo
    sys.modules['shapely'] = sh
    sys.modules['shapely.geometry'] = geo
    root = reference_root()
    if root is None:
        raise ImportError('the reference package is neither at /root/reference nor installed in oracle/_ref')
    if root not in sys.path:
        sys.path.insert(0, root)
    import geonomics  # noqa: F401
    return geonomics
