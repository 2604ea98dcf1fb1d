"""CPU tests of the host-side mirror (geonomics_b200/api.py), the C-ABI surface and the
no-fallback rule.  No compute call is made without a GPU."""
import ctypes
import os
import re

import numpy as np
import pytest

import __graft_entry__ as entry
from geonomics_b200 import _lib, api
from geonomics_b200.density import DensityGridSetup

HERE = os.path.dirname(os.path.abspath(__file__))
PARAMS = os.path.join(HERE, 'data', 'params_small.py')


@pytest.fixture(scope='module')
def built():
    entry.build()
    return _lib.lib()


def test_library_exports_every_declared_symbol(built):
    hdr = open(_lib.HEADER_PATH).read()
    declared = set(re.findall(r'\b(gnx_[A-Za-z_0-9]+)\s*\(', hdr))
    assert len(declared) >= 36
    for name in declared:
        assert hasattr(built, name), 'libgnxb200.so does not export %s' % name
    assert declared == set(_lib.SIGNATURES), (declared ^ set(_lib.SIGNATURES))
    assert built.gnx_abi_version() == _lib.GNX_ABI_VERSION


def test_ctypes_structs_match_header_field_order():
    hdr = open(_lib.HEADER_PATH).read()

    def fields(struct_name):
        end = hdr.index('} %s;' % struct_name)
        body = hdr[hdr.rindex('typedef struct {', 0, end) + len('typedef struct {'):end]
        body = re.sub(r'/\*.*?\*/', '', body, flags=re.S)
        names = []
        for decl in body.split(';'):
            decl = decl.strip()
            if not decl:
                continue
            for part in decl.split(','):
                names.append(re.sub(r'\[.*\]', '', part.strip().split()[-1].lstrip('*')))
        return names
    for cname, cls in (('gnx_config_t', _lib.Config), ('gnx_trait_t', _lib.Trait), ('gnx_density_t', _lib.Density),
                       ('gnx_draws_t', _lib.Draws), ('gnx_population_t', _lib.Population),
                       ('gnx_step_record_t', _lib.StepRecord), ('gnx_strip_config_t', _lib.StripConfig),
                       ('gnx_strip_endpoints_t', _lib.StripEndpoints)):
        assert fields(cname) == [f[0] for f in cls._fields_], cname


def test_no_cpu_fallback_without_a_gpu(built):
    import torch
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    p = api.read_parameters_file(PARAMS)
    with pytest.raises(_lib.GnxError):
        api.make_model(p)


def test_parameters_file_and_landscape():
    p = api.read_parameters_file(PARAMS)
    assert p.model.T == 12 and p.model.name == 'params_small'
    assert p.comm.species.spp_0.mating.b == 0.2
    np.random.seed(3)
    land = api._make_landscape(p)
    assert land.dim == (40, 40) and len(land) == 3
    assert land[1].name == 'lyr_1' and land[1].rast.shape == (40, 40)
    assert 0 <= land[2].rast.min() and land[2].rast.max() <= 1          # landscape.py:646-648
    assert land._changer is not None
    ts = [c[0] for c in land._changer.changes]
    assert ts == [4, 6, 8]                                                # linspace(4, 8, 3)
    # the last raster of the series is the change raster itself (change.py:349-354)
    assert np.allclose(land._changer.changes[-1][2], land[1].rast[:, ::-1])
    assert np.allclose(land._changer.changes[0][2], land[1].rast + (land[1].rast[:, ::-1] - land[1].rast) / 3)


def test_species_and_genomic_architecture_setup():
    p = api.read_parameters_file(PARAMS)
    np.random.seed(5)
    land = api._make_landscape(p)
    spp = api._make_species(land, 'spp_0', 0, p.comm.species.spp_0, seed=1)
    assert len(spp) == 1200 and spp.start_N == 1200
    assert spp.b == 0.2 and spp.mating_radius == 2 and spp.K_factor == 0.75
    assert spp.sex_ratio == 0.5
    assert np.allclose(spp.K, land[0].rast * 0.75)                       # species.py:546-547
    ga = spp.gen_arch
    assert ga.L == 60 and len(ga.traits) == 1
    t = ga.traits[0]
    assert len(t.loci) == 8 and len(t.alpha) == 8 and np.all(np.diff(t.loci) > 0)
    assert np.all(np.abs(t.alpha) <= 0.25)
    assert sorted(ga.nonneut_loci) == sorted(t.loci)
    r = ga.recombinations
    assert r._rates[0] == 0 and np.all(r._rates[1:] == 0.5)               # genome.py:183
    assert r._paths.shape == (2000, 60) and np.all(r._paths[:, 0] == 0)
    sub = r._get_subsetter(3)
    assert sub.sum() == 60 and np.array_equal(sub[1::2], r._paths[3] == 1)
    # Individual views before the device is attached
    ind = spp[5]
    assert ind.idx == 5 and 0 <= ind.x < 40 and ind.age == 0


def test_density_setup_matches_reference_construction():
    from oracle import step_oracle as so
    for dim, ww in (((40, 40), None), ((50, 30), None), ((24, 24), 4)):
        d = DensityGridSetup(dim, ww)
        o = so.DensityGridStack(dim, ww)
        assert np.array_equal(d.points, o.pts)
        assert np.array_equal(d.areas, np.hstack([g['areas'].ravel() for g in o.grids]))
        assert d.colourable == 1
        assert d.lat_ni * d.lat_nj == len(d.points)
        # every lattice square is covered by exactly two triangles
        assert d.square_tri.min() >= 0 and len(np.unique(d.square_tri)) == len(d.simplices)


def test_conductance_surface_table_statistics():
    np.random.seed(11)
    rast = np.zeros((5, 5))
    rast[2, 3] = 1.0                         # only the east neighbour of cell (2, 2) is non-zero
    surf = api._make_conductance_surface(rast, mixture=True, approx_len=400, vm_distr_kappa=12)
    assert surf.dtype == np.float16 and surf.shape == (5, 5, 400)
    ang = surf[2, 2].astype(np.float64)
    assert abs(np.mean(np.cos(ang)) - 0.957) < 0.03      # E[cos] = I1(12)/I0(12) around direction 0
    assert abs(np.mean(np.sin(ang))) < 0.05


def test_landscape_change_series_is_numpy_linspace():
    """ops/change.py:349-354: every cell follows np.linspace(start, end, n_steps + 1)[1:], at the
    time steps round(linspace(start_t, end_t, n_steps)) (change.py:310) -- bit for bit."""
    from geonomics_b200 import api
    rng = np.random.default_rng(3)
    start = rng.random((5, 7))
    end = rng.random((5, 7))
    end[0, 0] = start[0, 0]                      # a cell that does not change (step == 0)

    class _Lyr:
        rast = start
    land = {0: _Lyr()}
    ch = api._LandscapeChanger(land, {0: {0: dict(change_rast=end, start_t=3, end_t=17, n_steps=6)}})
    ts = np.int64(np.round(np.linspace(3, 17, 6)))
    assert [c[0] for c in ch.changes] == list(ts)
    ref = np.stack([np.linspace(start.flat[i], end.flat[i], 7)[1:] for i in range(start.size)])
    for k, (_, lyr, rast) in enumerate(ch.changes):
        assert lyr == 0
        assert np.array_equal(rast.ravel(), ref[:, k])
    assert np.array_equal(ch.changes[-1][2], end)


def test_landscape_res_ratio_matches_reference_formula():
    """landscape.py:277-278: each axis' cell size over the larger one (movement / dispersal
    distances are scaled by it on rasters with non-square cells, movement.py:79-82)."""
    from geonomics_b200 import api
    lyr = {0: api.Layer(np.ones((4, 6)), 'x', 'l0', (6, 4))}
    assert api.Landscape(lyr, res=(1, 1))._res_ratio == (1.0, 1.0)
    assert api.Landscape(lyr, res=(2, 1))._res_ratio == (1.0, 0.5)
    assert api.Landscape(lyr, res=(30, 90))._res_ratio == (30 / 90, 1.0)
