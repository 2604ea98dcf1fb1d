#!/usr/bin/env python
"""Golden files for the genetic-data writers (TEST INFRASTRUCTURE; build container only).

Runs the UNMODIFIED reference model of make_stat_golden.py (2 layers, 1 trait, 20 loci) through its
burn-in and a few main steps, then has the reference's own `Model.write_gendata`
(sim/model.py:3342-3396 -> sim/data.py:408-544) write a VCF (all sites and segregating sites only) and
a FASTA of the whole population and of an ad hoc sample.  The sample's arrays and the file texts go
into writers.npz; tests/test_writers.py rebuilds the texts from the arrays."""
import contextlib
import io
import os
import sys
import tempfile
import warnings
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, HERE)
warnings.filterwarnings('ignore')
from oracle import ref_shims   # noqa: E402
from make_stat_golden import stat_params   # noqa: E402


def sample_arrays(spp, ids):
    inds = [spp[i] for i in ids]
    return dict(ids=np.array(ids, dtype=np.int64), x=np.array([i.x for i in inds]), y=np.array([i.y for i in inds]),
                age=np.array([i.age for i in inds], dtype=np.int64), sex=np.array([i.sex for i in inds], dtype=np.int64),
                z=np.array([i.z for i in inds], dtype=np.float64), e=np.array([i.e for i in inds]),
                g=np.stack([i.g for i in inds]).astype(np.int8))


class _Point:
    """shapely.geometry.Point as utils/io.py:176 uses it: a holder of x and y."""
    def __init__(self, x, y):
        self.x, self.y = float(x), float(y)


class _PointColumn:
    def __init__(self, pts):
        self.x = [p.x for p in pts]
        self.y = [p.y for p in pts]


class _GeoPandas:
    """geopandas (absent here) as the CSV branch of utils/io.py:165-186 uses it: a DataFrame whose geometry
    column answers .x / .y with float columns; everything written comes from pandas' own to_csv."""
    @staticmethod
    def GeoDataFrame(df, geometry):
        import pandas as pd

        class GDF(pd.DataFrame):
            @property
            def pt(self):
                return _PointColumn(list(self[geometry]))
        return GDF(df)


def main():
    gnx = ref_shims.install()
    import geonomics.sim.burnin as _b
    _b._test_t_threshold = lambda *a, **k: True
    _b.SpatialTester.run_test = lambda self, n, alpha=0.05: True
    from geonomics.sim.params import ParametersDict
    import geonomics.sim.data as data
    p = ParametersDict(stat_params())
    p['model']['seed'] = {'num': 4242}
    p['model']['name'] = 'writers'
    p['comm']['species']['spp_0']['gen_arch']['start_p_fixed'] = 0.2      # some loci fix within the run
    with contextlib.redirect_stdout(io.StringIO()):
        mod = gnx.make_model(p, name='writers')
        mod.walk(10000, 'burn', verbose=False)
        assignment_max_idx = mod.comm[0].max_ind_idx      # genomes are assigned when the burn-in ends
        mod.walk(12, 'main', verbose=False)
    spp = mod.comm[0]
    out = dict(L=np.int64(spp.gen_arch.L), assignment_max_idx=np.int64(assignment_max_idx), numpy_major=np.int64(int(np.__version__.split('.')[0])))
    tmp = tempfile.mkdtemp(prefix='gnx_writers_')

    def written(name, **kw):
        path = os.path.join(tmp, name)
        mod.write_gendata(path, **kw)
        return open(path).read()

    ids_all = sorted(spp)
    for k, v in sample_arrays(spp, ids_all).items():
        out['all_' + k] = v
    out['vcf_all_fixed'] = written('a.vcf', include_fixed_sites=True)
    out['vcf_all_seg'] = written('b.vcf', include_fixed_sites=False)
    out['fasta_all'] = written('a.fasta')
    # an ad hoc sample of 17: the ids the reference drew are recovered from its own file
    np.random.seed(99)
    txt = written('c.vcf', n=17, include_fixed_sites=True)
    ids_s = [int(v) for v in txt.split('\n')[3].split('\t')[9:]]
    np.random.seed(99)
    assert sorted(data._get_adhoc_sample(spp, 17)) == ids_s
    for k, v in sample_arrays(spp, ids_s).items():
        out['sub_' + k] = v
    out['vcf_sub_fixed'] = txt
    np.random.seed(99)
    out['fasta_sub'] = written('c.fasta', n=17)
    out['sub_seed'] = np.int64(99)
    # fixed sites (none arose in this run): the reference's formatter on the sample's genotypes with
    # locus 3 lost and locus 7 fixed
    sample = data._get_adhoc_sample(spp, None)
    ids_f = [*sample][:40]
    gts = {i: np.array(sample[i].g, copy=True) for i in ids_f}
    for gt in gts.values():
        gt[3, :] = 0
        gt[7, :] = 1
    sample = {i: sample[i] for i in ids_f}
    out['fix_ids'] = np.array(ids_f, dtype=np.int64)
    out['fix_g'] = np.stack([gts[i] for i in ids_f]).astype(np.int8)
    out['vcf_fix_fixed'] = data._format_vcf(sample, gts, spp.gen_arch, include_fixed_sites=True)
    out['vcf_fix_seg'] = data._format_vcf(sample, gts, spp.gen_arch, include_fixed_sites=False)
    # geodata CSV (model.py:3399-3446 -> utils/io.py:165-186), whole population and the same ad hoc sample
    import geonomics.utils.io as gio
    gio.gpd, gio.Point = _GeoPandas, _Point

    def written_geo(name, **kw):
        path = os.path.join(tmp, name)
        mod.write_geodata(path, **kw)
        return open(path).read()

    out['csv_all'] = written_geo('a.csv')
    np.random.seed(99)
    out['csv_sub'] = written_geo('c.csv', n=17)
    np.savez_compressed(os.path.join(HERE, 'writers.npz'), **out)
    segs = sum(1 for ln in out['vcf_all_fixed'].split('\n') if '\tSEG\t' in ln)
    print('wrote writers.npz: N %d, L %d (%d segregating), sample of %d' % (len(ids_all), out['L'], segs, len(ids_s)))


if __name__ == '__main__':
    main()
