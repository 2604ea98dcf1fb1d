"""Statistical parity (north_star): population-size, allele-frequency, heterozygosity and
Fst distributions over 100 replicates of the GPU path must be indistinguishable (two-sample
Kolmogorov-Smirnov) from 400 replicates of the unmodified reference, recorded in
tests/golden/stat_reference.npz by tests/golden/make_stat_golden.py."""
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, 'golden'))
REF = os.path.join(HERE, 'golden', 'stat_reference.npz')

ALPHA = 1e-3      # per-test false-alarm rate; 4 times x 6 statistics are tested


@pytest.fixture(scope='module')
def samples():
    import make_stat_golden as msg
    from geonomics_b200 import api
    ref = np.load(REF)
    n_reps = 100            # GPU replicates, against the reference's 400

    def fixed_burn(self):
        ok = all(len(s.Nt) >= self.burn_T for s in self.comm.values())
        for s in self.comm.values():
            s.burned = ok
        self.comm.burned = ok
    orig = api.Model._check_comm_burned
    api.Model._check_comm_burned = fixed_burn
    try:
        Nt, stats = [], []
        for rep in range(n_reps):
            p = api.make_params_dict(msg.stat_params(), 'stat')
            p['model']['seed'] = {'num': 5000 + rep}
            mod = api.make_model(p)
            mod.walk(10000, 'burn')
            spp = mod.comm[0]
            tl = int(spp.gen_arch.traits[0].loci[0])
            nl = int([l for l in range(spp.gen_arch.L) if l != tl][0])
            rows = []
            for t in range(msg.T):
                mod.walk(1, 'main')
                if t in msg.SAMPLE_T:
                    rows.append(msg.summarise(mod.get_x(), mod.get_genotypes(), tl, nl))
            Nt.append(spp.Nt[-msg.T:])
            stats.append(rows)
            assert mod.burn_t + 1 == msg.BURN_T
    finally:
        api.Model._check_comm_burned = orig
    return ref, np.array(Nt), np.array(stats)


@pytest.mark.gpu
def test_population_size_trajectories(samples):
    from scipy.stats import ks_2samp
    ref, Nt, stats = samples
    for t in (0, 9, 19, 29, 39):
        p = ks_2samp(ref['Nt'][:, t], Nt[:, t]).pvalue
        assert p > ALPHA, 'N at t=%d differs from the reference (KS p=%.2g; means %.1f vs %.1f)' % (
            t, p, ref['Nt'][:, t].mean(), Nt[:, t].mean())
    p = ks_2samp(ref['Nt'].mean(axis=1), Nt.mean(axis=1)).pvalue
    assert p > ALPHA


@pytest.mark.gpu
@pytest.mark.parametrize('col,name', [(1, 'trait-locus allele frequency'), (2, 'neutral allele frequency'),
                                      (3, 'heterozygosity'), (4, 'Fst west/east'), (5, 'trait-locus cline')])
def test_genetic_trajectories(samples, col, name):
    from scipy.stats import ks_2samp
    ref, Nt, stats = samples
    for k in range(stats.shape[1]):
        a, b = ref['stats'][:, k, col], stats[:, k, col]
        a, b = a[np.isfinite(a)], b[np.isfinite(b)]
        p = ks_2samp(a, b).pvalue
        assert p > ALPHA, '%s at sample %d differs from the reference (KS p=%.2g; means %.4f vs %.4f)' % (
            name, k, p, a.mean(), b.mean())


def test_reference_sample_present():
    ref = np.load(REF)
    assert ref['Nt'].shape[0] >= 50 and ref['stats'].shape[2] == 6
