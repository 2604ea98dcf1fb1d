"""Genetic-data writers over the state read back from the device.

The reference stringifies one genotype at a time (sim/data.py:427-544: `_format_fasta` joins
`str(base)` per locus per homologue, `_format_vcf` formats `'%i|%i'` per individual per locus and
grows the file text by `rows = rows + ...`).  Here the genotype text is laid out as ONE byte matrix
(numpy, no per-genotype Python work) and only the short per-row prefixes are formatted in Python.
The text is the reference's, byte for byte (tests/test_writers.py pins both formats to files the
unmodified reference wrote: tests/golden/writers.npz); the date of the VCF header is the only
field that varies between runs (data.py:531-535).

Inputs are the arrays `Species._state()` holds after one download (`g` int8[N, L, 2], species
order) and the `Individual` views of the host API; nothing here touches the device.
"""
import datetime
import re

import numpy as np

_TAB, _NL, _BAR, _ZERO = 9, 10, 124, 48


def _genotype_bytes(g):
    g = np.ascontiguousarray(g)
    if g.ndim != 3 or g.shape[2] != 2:
        raise ValueError('genotypes must be [n, L, 2]')
    if g.size and (g.min() < 0 or g.max() > 9):
        raise ValueError('allele codes must be single digits')      # '%i' of the reference: one character here
    return g.astype(np.uint8)


def segregating_sites(g):
    """data.py:500-503: loci whose summed allele count over the sample is neither 0 nor 2n."""
    g = np.asarray(g)
    tot = g.sum(axis=2, dtype=np.int64).sum(axis=0)
    return np.where((tot > 0) & (tot < 2 * g.shape[0]))[0]


def format_vcf(ids, g, L, include_fixed_sites=False, date=None):
    """data.py:460-544 `_format_vcf(sample, genotypes, gen_arch, include_fixed_sites)`.

    ids: the sample's individual ids in file order; g: int[n, L, 2] their genotypes, same order.
    One row per locus: `0 <tab> locus <tab> . A T 1000 PASS SEG|FIX GT` then `a|b` per individual."""
    ids = [int(i) for i in ids]
    g8 = _genotype_bytes(g)
    n = g8.shape[0]
    if n != len(ids):
        raise ValueError("'sample' and 'genotypes' have different lengths")   # data.py:463
    segs = segregating_sites(g8)
    loci = np.arange(int(L)) if include_fixed_sites else segs
    is_seg = np.zeros(max(int(L), g8.shape[1]), dtype=bool)
    is_seg[segs] = True
    if date is None:
        now = datetime.datetime.now()
        date = '%d%s%s' % (now.year, str(now.month).zfill(2), str(now.day).zfill(2))
    head = '##fileformat=VCFv4.2\n##fileDate=%s\n##source=Geonomics\n' % date
    cols = '#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\t%s\n' % '\t'.join(str(i) for i in ids)
    # the genotype columns of every row at once: [n_rows, n, 4] = a | b <tab>, the last <tab> a newline
    body = np.empty((len(loci), n, 4), dtype=np.uint8)
    sel = g8[:, loci, :]                                   # [n, n_rows, 2]
    body[:, :, 0] = sel[:, :, 0].T + _ZERO
    body[:, :, 1] = _BAR
    body[:, :, 2] = sel[:, :, 1].T + _ZERO
    body[:, :, 3] = _TAB
    body = body.reshape(len(loci), n * 4)
    if n:
        body[:, -1] = _NL
    parts = [head.encode(), cols.encode()]
    for r, locus in enumerate(loci):
        parts.append(('0\t%i\t.\tA\tT\t1000\tPASS\t%s\tGT\t' % (locus, 'SEG' if is_seg[locus] else 'FIX')).encode())
        parts.append(body[r].tobytes() if n else b'\n')
    return b''.join(parts).decode()


def _attr_text(v):
    # data.py:446-448: str() of the attribute, brackets and blanks dropped, commas -> '|'
    return re.sub(',', '|', re.sub(r'[\[\] ]', '', str(v)))


def _sequence_lines(g8, width):
    # [n, L, 2] allele codes -> uint8[n, 2, L * width + 1]: one text line per homologue, '\n' included
    n, L = g8.shape[0], g8.shape[1]
    seq = np.empty((n, 2, L * width + 1), dtype=np.uint8)
    seq[:, :, 0:L * width:width] = np.transpose(g8, (0, 2, 1)) + _ZERO
    if width == 3:
        seq[:, :, 1:L * width:width] = 46           # '.'
        seq[:, :, 2:L * width:width] = _ZERO
    seq[:, :, L * width] = _NL
    return seq


def format_fasta(sample, g, float_rows=None):
    """data.py:427-457 `_format_fasta(sample, genotypes)`.

    sample: mapping id -> Individual view (idx, x, y, age, sex, z, e) in file order; g: [n, L, 2] in the
    same order.  Two records per individual, `>idx:hap;x;y;age;sex;z;e` then the L alleles.

    The sequence line is ''.join(str(base)) (data.py:451), so it follows the dtype of the individual's
    array in the reference: one digit per allele for int8 arrays -- every newborn (individual.py:104) and
    everyone of a use_tskit species (species.py:1418) -- and 'd.0' per allele for the float64 arrays
    that the individuals alive at the genome assignment keep (np.zeros, species.py:898-905).
    float_rows: bool[n], who prints as float64; default: everyone if g itself is floating, else no one."""
    if float_rows is None:
        float_rows = np.full(len(g), np.issubdtype(np.asarray(g).dtype, np.floating))
    float_rows = np.asarray(float_rows, dtype=bool)
    g8 = _genotype_bytes(g)
    inds = list(sample.values()) if hasattr(sample, 'values') else list(sample)
    if len(inds) != g8.shape[0] or len(float_rows) != g8.shape[0]:
        raise ValueError("'sample' and 'genotypes' have different lengths")   # data.py:438
    seq = _sequence_lines(g8, 1)
    seq_f = _sequence_lines(g8[float_rows], 3)
    row_f = np.cumsum(float_rows) - 1
    parts = []
    for k, ind in enumerate(inds):
        tail = ';'.join(_attr_text(getattr(ind, a)) for a in ('x', 'y', 'age', 'sex', 'z', 'e'))
        idx = _attr_text(ind.idx)
        lines = seq_f[row_f[k]] if float_rows[k] else seq[k]
        for hap in range(2):
            parts.append(('>%s:%i;%s\n' % (idx, hap, tail)).encode())
            parts.append(lines[hap].tobytes())
    return b''.join(parts).decode()


def adhoc_sample_ids(ids, n=None, rng=None):
    """data.py:408-424 `_get_adhoc_sample`: everyone (n None or >= the population), else n ids drawn
    without replacement through numpy's global stream (`r.choice`, as the reference), sorted."""
    ids = [int(i) for i in ids]
    if n is not None and len(ids) > n:
        choice = (rng or np.random).choice
        ids = [int(i) for i in choice(ids, size=n, replace=False)]
    return sorted(set(ids))


def write_gendata(filepath, spp, n=None, include_fixed_sites=True):
    """model.py:3342-3396 `Model.write_gendata`: '.vcf' or '.fasta' by extension."""
    ext = filepath.split('.')[-1].lower()
    assert ext in ('vcf', 'fasta'), ('Must provide valid file extension. Valid extensions '
                                     'include ".vcf" and ".fasta".')
    ids = adhoc_sample_ids([*spp], n)
    g = spp._get_genotypes(individs=ids, all_loci=True)
    if g is None:
        raise ValueError('the species has no genomes to write (no genomic architecture, or not burned in)')
    if ext == 'vcf':
        text = format_vcf(ids, g, spp.gen_arch.L, include_fixed_sites=include_fixed_sites)
    else:
        # who was alive when the genomes were assigned (ids are handed out in increasing order)
        last = getattr(spp, '_genome_assignment_max_idx', None)
        founders = np.zeros(len(ids), dtype=bool) if (spp.gen_arch.use_tskit or last is None) \
            else np.asarray(ids) <= last
        text = format_fasta({i: spp[i] for i in ids}, g, float_rows=founders)
    with open(filepath, 'w') as f:
        f.write(text)
    return filepath


def format_geodata_csv(sample):
    """utils/io.py:165-186 `_write_geopandas(..., driver='CSV')`: columns idx, z, e, age, sex as str() of the
    Individual's attributes, then x and y as floats, written by pandas (the geopandas frame of the reference
    only supplies x and y from its Point column before `to_csv`)."""
    import pandas as pd
    inds = list(sample.values()) if hasattr(sample, 'values') else list(sample)
    cols = {att: [str(getattr(ind, att)) for ind in inds] for att in ('idx', 'z', 'e', 'age', 'sex')}
    cols['x'] = np.array([ind.x for ind in inds], dtype=np.float64)
    cols['y'] = np.array([ind.y for ind in inds], dtype=np.float64)
    return pd.DataFrame(cols).to_csv(index=False)


def write_geodata(filepath, spp, n=None):
    """model.py:3399-3446 `Model.write_geodata`.  '.csv' only: the shapefile and GeoJSON drivers are
    geopandas / fiona file formats, and neither library is in this image."""
    ext = filepath.split('.')[-1].lower()
    assert ext in ('csv', 'shp', 'json'), ('Must provide valid file extension. Valid extensions '
                                           'include ".csv", ".shp", and ".json".')
    if ext != 'csv':
        raise NotImplementedError("'.%s' needs geopandas (utils/io.py:185), which is not installed" % ext)
    ids = adhoc_sample_ids([*spp], n)
    with open(filepath, 'w') as f:
        f.write(format_geodata_csv({i: spp[i] for i in ids}))
    return filepath
