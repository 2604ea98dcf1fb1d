"""Genetic-data writers (geonomics_b200/writers.py) against files the unmodified reference wrote
(tests/golden/writers.npz, made by tests/golden/make_writers_golden.py from Model.write_gendata,
sim/model.py:3342-3396 -> sim/data.py:408-544).  Host-side formatting only: no device call."""
import os
import re

import numpy as np
import pytest

from geonomics_b200 import writers
from geonomics_b200.api import Individual

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope='module')
def gold():
    return np.load(os.path.join(HERE, 'golden', 'writers.npz'))


def _date_of(vcf_text):
    return re.search(r'##fileDate=(\d+)', vcf_text).group(1)


def _sample(z, prefix):
    ids = z[prefix + 'ids']
    out = {}
    for k, i in enumerate(ids):
        # the reference's Individuals: z and e are lists of numpy scalars (individual.py:148-155)
        out[int(i)] = Individual(int(i), z[prefix + 'x'][k], z[prefix + 'y'][k], int(z[prefix + 'age'][k]), None,
                                 int(z[prefix + 'sex'][k]), list(z[prefix + 'e'][k]), list(z[prefix + 'z'][k]))
    return out


@pytest.mark.parametrize('prefix,key,fixed', [('all_', 'vcf_all_fixed', True), ('all_', 'vcf_all_seg', False),
                                              ('sub_', 'vcf_sub_fixed', True)])
def test_vcf_equals_reference_file(gold, prefix, key, fixed):
    ref = str(gold[key])
    got = writers.format_vcf(gold[prefix + 'ids'], gold[prefix + 'g'], int(gold['L']), include_fixed_sites=fixed,
                             date=_date_of(ref))
    assert got == ref


@pytest.mark.parametrize('key,fixed,n_rows', [('vcf_fix_fixed', True, 20), ('vcf_fix_seg', False, 18)])
def test_vcf_fixed_and_lost_sites(gold, key, fixed, n_rows):
    ref = str(gold[key])
    got = writers.format_vcf(gold['fix_ids'], gold['fix_g'], int(gold['L']), include_fixed_sites=fixed,
                             date=_date_of(ref))
    assert got == ref
    rows = [ln for ln in got.split('\n') if ln and not ln.startswith('#')]
    assert len(rows) <= n_rows and sum('\tFIX\t' in r for r in rows) == (2 if fixed else 0)
    assert np.array_equal(np.setdiff1d(np.arange(20), writers.segregating_sites(gold['fix_g']))[:2], [3, 7])


@pytest.mark.parametrize('prefix,key', [('all_', 'fasta_all'), ('sub_', 'fasta_sub')])
def test_fasta_equals_reference_file(gold, prefix, key):
    if int(np.__version__.split('.')[0]) != int(gold['numpy_major']):
        pytest.skip('the header fields are str() of numpy scalars: text recorded under another numpy major')
    # whoever was alive at the genome assignment carries a float64 array (species.py:898-905), every
    # newborn an int8 one (individual.py:104)
    got = writers.format_fasta(_sample(gold, prefix), gold[prefix + 'g'],
                               float_rows=gold[prefix + 'ids'] <= int(gold['assignment_max_idx']))
    assert got == str(gold[key])
    assert 0 < (gold['all_ids'] <= int(gold['assignment_max_idx'])).sum() < len(gold['all_ids'])


def test_fasta_float_genotypes_three_characters_per_allele(gold):
    g = gold['sub_g'].astype(np.float64)
    lines = writers.format_fasta(_sample(gold, 'sub_'), g).split('\n')
    assert lines[1] == ''.join(str(b) for b in g[0, :, 0]) and len(lines[1]) == 3 * g.shape[1]


def test_fasta_integer_genotypes_one_digit_per_allele(gold):
    g = gold['sub_g']
    txt = writers.format_fasta(_sample(gold, 'sub_'), g)
    lines = txt.split('\n')
    assert len(lines) == 4 * len(g) + 1 and lines[-1] == ''
    for k in range(len(g)):
        for h in (0, 1):
            assert lines[4 * k + 2 * h].startswith('>%d:%d;' % (gold['sub_ids'][k], h))
            assert lines[4 * k + 2 * h + 1] == ''.join(str(int(b)) for b in g[k, :, h])


@pytest.mark.parametrize('prefix,key', [('all_', 'csv_all'), ('sub_', 'csv_sub')])
def test_geodata_csv_equals_reference_file(gold, prefix, key, tmp_path):
    if int(np.__version__.split('.')[0]) != int(gold['numpy_major']):
        pytest.skip('the z and e columns are str() of lists of numpy scalars: recorded under another numpy major')
    sample = _sample(gold, prefix)
    assert writers.format_geodata_csv(sample) == str(gold[key])
    spp = _Spp(_sample(gold, 'all_'), gold['all_g'], int(gold['L']), False, int(gold['assignment_max_idx']))
    np.random.seed(int(gold['sub_seed']))
    path = writers.write_geodata(str(tmp_path / 'geo.csv'), spp, n=None if prefix == 'all_' else 17)
    assert open(path).read() == str(gold[key])
    with pytest.raises(NotImplementedError):
        writers.write_geodata(str(tmp_path / 'geo.shp'), spp)


def test_adhoc_sample_draws_like_the_reference(gold):
    # data.py:408-424: r.choice(ids, n, replace=False) on numpy's global stream, then sorted
    np.random.seed(int(gold['sub_seed']))
    assert writers.adhoc_sample_ids(gold['all_ids'], 17) == [int(i) for i in gold['sub_ids']]
    assert writers.adhoc_sample_ids(gold['all_ids'], None) == sorted(int(i) for i in gold['all_ids'])
    assert writers.adhoc_sample_ids([5, 3, 9], 10) == [3, 5, 9]


def test_empty_and_mismatched_inputs():
    txt = writers.format_vcf([], np.zeros((0, 4, 2), np.int8), 4, include_fixed_sites=True, date='20260101')
    assert txt.count('\n') == 4 + 4 and '\tFIX\t' in txt          # 3 header lines + columns + one row per locus
    assert writers.format_fasta({}, np.zeros((0, 4, 2), np.int8)) == ''
    with pytest.raises(ValueError):
        writers.format_vcf([1, 2], np.zeros((1, 4, 2), np.int8), 4)
    with pytest.raises(ValueError):
        writers.format_vcf([1], np.full((1, 4, 2), 12, np.int8), 4)


class _Spp(dict):
    """The slice of the host Species the writer reads."""
    def __init__(self, sample, g, L, use_tskit, assignment_max_idx):
        super().__init__(sample)
        self._g = g
        self._genome_assignment_max_idx = assignment_max_idx
        self.gen_arch = type('GA', (), dict(L=L, use_tskit=use_tskit))()

    def _get_genotypes(self, individs=None, all_loci=False):
        assert all_loci
        order = {i: k for k, i in enumerate(self)}
        return self._g[[order[i] for i in individs]]


@pytest.mark.parametrize('ext', ['vcf', 'fasta'])
def test_write_gendata_files(gold, tmp_path, ext):
    if ext == 'fasta' and int(np.__version__.split('.')[0]) != int(gold['numpy_major']):
        pytest.skip('text recorded under another numpy major')
    spp = _Spp(_sample(gold, 'all_'), gold['all_g'], int(gold['L']), False, int(gold['assignment_max_idx']))
    path = str(tmp_path / ('out.' + ext))
    writers.write_gendata(path, spp)
    got = open(path).read()
    ref = str(gold['vcf_all_fixed' if ext == 'vcf' else 'fasta_all'])
    if ext == 'vcf':
        got = got.replace(_date_of(got), _date_of(ref), 1)
    assert got == ref
    with pytest.raises(AssertionError):
        writers.write_gendata(str(tmp_path / 'out.txt'), spp)


@pytest.mark.gpu
def test_model_write_gendata_from_device_state(tmp_path):
    """Model.write_gendata (model.py:3342-3396) over the state downloaded from the device: the VCF decodes
    back to get_genotypes() in ascending id order, the FASTA tells the individuals alive at the genome
    assignment (float64 arrays in the reference) from those born since (int8)."""
    from geonomics_b200 import api
    mod = api.make_model(os.path.join(HERE, 'data', 'params_small.py'))
    mod.walk(10000, 'burn')
    mod.walk(3, 'main')
    spp = mod.comm[0]
    ids = np.array([i for i in spp])
    order = np.argsort(ids)
    g = mod.get_genotypes()[order]
    vcf = open(mod.write_gendata(str(tmp_path / 'pop.vcf'))).read().split('\n')
    assert [int(v) for v in vcf[3].split('\t')[9:]] == [int(i) for i in ids[order]]
    rows = [ln.split('\t') for ln in vcf[4:] if ln]
    assert [int(r[1]) for r in rows] == list(range(spp.gen_arch.L))
    got = np.array([[[int(c) for c in gt.split('|')] for gt in r[9:]] for r in rows])      # [L, N, 2]
    assert np.array_equal(np.transpose(got, (1, 0, 2)), g)
    seg = writers.segregating_sites(g)
    assert [r[7] for r in rows] == ['SEG' if l in set(seg.tolist()) else 'FIX' for l in range(spp.gen_arch.L)]
    np.random.seed(3)
    sub = open(mod.write_gendata(str(tmp_path / 'sub.vcf'), n=25, include_fixed_sites=False)).read().split('\n')
    np.random.seed(3)
    assert [int(v) for v in sub[3].split('\t')[9:]] == writers.adhoc_sample_ids([*spp], 25)
    geo = open(mod.write_geodata(str(tmp_path / 'pop.csv'))).read().split('\n')
    assert geo[0] == 'idx,z,e,age,sex,x,y' and len(geo) == len(ids) + 2
    assert [int(r.split(',')[0]) for r in geo[1:-1]] == [int(i) for i in ids[order]]
    assert [float(r.split(',')[-2]) for r in geo[1:-1]] == [float(v) for v in mod.get_x()[order]]
    fasta = open(mod.write_gendata(str(tmp_path / 'pop.fasta'))).read().split('\n')
    assert len(fasta) == 4 * len(ids) + 1
    founders = ids[order] <= spp._genome_assignment_max_idx
    assert 0 < founders.sum() < len(ids)
    for k in (0, len(ids) // 2, len(ids) - 1):
        for h in (0, 1):
            assert fasta[4 * k + 2 * h].startswith('>%d:%d;%s;' % (ids[order][k], h, str(float(spp[int(ids[order][k])].x))))
            want = ''.join((str(float(b)) if founders[k] else str(int(b))) for b in g[k, :, h])
            assert fasta[4 * k + 2 * h + 1] == want
