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
  * tskit / msprime: only reached with gen_arch.use_tskit=True
    (species.py:442, 978); golden vectors use use_tskit=False.
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


def install():
    """Install the shims, then make `import geonomics` resolve to the reference."""
    for name in ['matplotlib', 'matplotlib.pyplot', 'matplotlib.colors',
                 'matplotlib.animation', 'matplotlib.gridspec',
                 'matplotlib.ticker', 'matplotlib.lines', 'mpl_toolkits',
                 'mpl_toolkits.mplot3d', 'mpl_toolkits.mplot3d.axes3d',
                 'mpl_toolkits.axes_grid1', 'geopandas', 'rasterio', 'tskit',
                 'msprime', 'statsmodels', 'statsmodels.api',
                 'statsmodels.tsa', 'vcf', 'nlmpy']:
        if name not in sys.modules:
            sys.modules[name] = MagicMock(name=name)
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
