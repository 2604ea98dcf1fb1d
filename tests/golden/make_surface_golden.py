#!/usr/bin/env python
"""Golden conductance-surface tables (TEST INFRASTRUCTURE; runs only in the build container).

Calls the *unmodified* reference's table builder `_make_conductance_surface`
(/root/reference/geonomics/utils/spatial.py:365-461) on a small raster, for the mixture and
the unimodal variant, and stores the float16 direction tables.  The GPU's on-the-fly sampler
(GNX_SURF_ONTHEFLY: the same distribution drawn per individual instead of tabulated per cell)
is tested against these tables and against the analytic mixture the builder samples
(tests/test_cuda_surface_onthefly.py).

The raster is 7 rows x 6 columns (non-square, so an x/y swap shows) and holds a 3x3 block of
zeros: cell (4, 1) has an all-zero neighbourhood (uniform mixture weights, spatial.py:415-418;
mean of all eight directions in the unimodal variant, spatial.py:376-381) and cell (4, 0) adds
the zero-embedded landscape edge to it.

Usage:  python tests/golden/make_surface_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import ref_shims   # noqa: E402

APPROX_LEN = 1500
KAPPA = 12


def surface_raster():
    rng = np.random.default_rng(404)
    r = 0.1 + 0.9 * rng.random((7, 6))
    r[3:6, 0:3] = 0.0
    r[1, 4] = 1.0                      # a unique maximum among the neighbours of (2, 3), (1, 3), ...
    return r


def main():
    ref_shims.install()
    import geonomics.utils.spatial as sp
    rast = surface_raster()
    out = dict(rast=rast, kappa=np.float64(KAPPA))
    for name, mix in (('mix', True), ('uni', False)):
        np.random.seed(2024 + int(mix))
        tab = sp._make_conductance_surface(rast, mixture=mix, approx_len=APPROX_LEN, vm_distr_kappa=KAPPA)
        assert tab.dtype == np.float16 and tab.shape == rast.shape + (APPROX_LEN,)
        out['table_' + name] = tab
    path = os.path.join(HERE, 'surface_tables.npz')
    np.savez_compressed(path, **out)
    print('wrote', path, os.path.getsize(path), 'bytes')


if __name__ == '__main__':
    main()
