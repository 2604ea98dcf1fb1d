#!/usr/bin/env python
"""Statistical golden data (TEST INFRASTRUCTURE; runs only in the build container).

Runs R independent replicates of a small model through the UNMODIFIED reference
(free-running numpy.random, seeds 1..R) and records, per replicate, the trajectories the
north_star's KS tests compare: population size, allele frequency at a trait locus and at a
neutral locus, mean expected heterozygosity, and Fst between the west and east halves of
the landscape (Hs/Ht form, tests/validation/island/island_test.py:54-68).
tests/test_statistical_parity.py runs the same model on the GPU path and KS-tests each
statistic against these samples.

Usage: python tests/golden/make_stat_golden.py [n_reps [seed_base [out.npz]]]

The committed stat_reference.npz holds 400 replicates: seeds 1000-1099 plus four blocks of 75
(seed bases 2000, 3000, 4000, 5000; `seed_blocks` in the file), generated block by block with
this script and concatenated.  (The first 100 alone happened to sit ~2 standard errors above
the long-run mean population size, which a 100-vs-100 KS test then pins on the GPU arm.)
"""
import os
import sys
import io
import contextlib
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
warnings.filterwarnings('ignore')
from oracle import ref_shims   # noqa: E402

DIM = (20, 20)
BURN_T = 30
T = 40
SAMPLE_T = (9, 19, 29, 39)


def stat_params():
    """The model, as a plain nested dict (also imported by the GPU test)."""
    xg = np.tile(np.linspace(0, 1, DIM[0]), (DIM[1], 1))
    return {
        'landscape': {
            'main': {'dim': DIM, 'res': (1, 1), 'ulc': (0, 0), 'prj': None},
            'layers': {
                'lyr_0': {'init': {'defined': {'rast': np.ones((DIM[1], DIM[0])), 'pts': None, 'vals': None,
                                               'interp_method': None}}},
                'lyr_1': {'init': {'defined': {'rast': xg, 'pts': None, 'vals': None, 'interp_method': None}}},
            },
        },
        'comm': {'species': {'spp_0': {
            'init': {'N': 500, 'K_layer': 'lyr_0', 'K_factor': 1.25},
            'mating': {'repro_age': 0, 'sex': False, 'sex_ratio': 1 / 1, 'R': 0.5, 'b': 0.3,
                       'n_births_distr_lambda': 1, 'n_births_fixed': True, 'mating_radius': 2,
                       'choose_nearest_mate': False, 'inverse_dist_mating': False},
            'mortality': {'max_age': None, 'd_min': 0, 'd_max': 1, 'density_grid_window_width': None},
            'movement': {'move': True, 'direction_distr_mu': 0, 'direction_distr_kappa': 0,
                         'movement_distance_distr_param1': 0.5, 'movement_distance_distr_param2': 1.0,
                         'movement_distance_distr': 'wald',
                         'dispersal_distance_distr_param1': 0.5, 'dispersal_distance_distr_param2': 1.0,
                         'dispersal_distance_distr': 'wald'},
            'gen_arch': {
                'gen_arch_file': None, 'L': 20, 'start_p_fixed': 0.5, 'start_neut_zero': False,
                'mu_neut': 0, 'mu_delet': 0, 'delet_alpha_distr_shape': 0.2, 'delet_alpha_distr_scale': 0.2,
                'r_distr_alpha': 0.5, 'r_distr_beta': None, 'dom': False, 'pleiotropy': False,
                'recomb_rate_custom_fn': None, 'n_recomb_paths_mem': int(1e4), 'n_recomb_paths_tot': int(1e5),
                'n_recomb_sims': 1000, 'allow_ad_hoc_recomb': False, 'jitter_breakpoints': False,
                'mut_log': False, 'use_tskit': False, 'tskit_simp_interval': 100,
                'traits': {'trait_0': {'layer': 'lyr_1', 'phi': 0.5, 'n_loci': 1, 'mu': 0,
                                       'alpha_distr_mu': 0.1, 'alpha_distr_sigma': 0, 'max_alpha_mag': None,
                                       'gamma': 1, 'univ_adv': False}},
            },
        }}},
        'model': {'T': T, 'burn_T': BURN_T, 'seed': {'num': 1},
                  'its': {'n_its': 1, 'rand_landscape': False, 'rand_comm': False, 'rand_genarch': True,
                          'repeat_burn': False}},
    }


def summarise(x, g, trait_locus, neut_locus):
    """x float[N], g int8[N, L, 2] -> (N, p_trait, p_neut, mean_het, fst_west_east, cline)."""
    n = len(x)
    p = g.mean(axis=(0, 2))
    het = float(np.mean(2 * p * (1 - p)))
    west = x < DIM[0] / 2
    fst = np.nan
    if west.sum() > 0 and (~west).sum() > 0:
        pw = g[west].mean(axis=(0, 2))
        pe = g[~west].mean(axis=(0, 2))
        hs = np.mean((2 * pw * (1 - pw) + 2 * pe * (1 - pe)) / 2)
        pt = (pw + pe) / 2
        ht = np.mean(2 * pt * (1 - pt))
        fst = float((ht - hs) / ht) if ht > 0 else np.nan
        cline = float(pe[trait_locus] - pw[trait_locus])
    else:
        cline = np.nan
    return n, float(p[trait_locus]), float(p[neut_locus]), het, fst, cline


def run_reference(n_reps, seed_base=1000, out_path=None):
    gnx = ref_shims.install()
    import geonomics.sim.burnin as _b
    # fixed-length burn-in in both arms (burn-in control is out of scope)
    _b._test_t_threshold = lambda *a, **k: True
    _b.SpatialTester.run_test = lambda self, n, alpha=0.05: True
    from geonomics.sim.params import ParametersDict
    out = []
    for rep in range(n_reps):
        p = ParametersDict(stat_params())
        p['model']['seed'] = {'num': seed_base + rep}
        p['model']['name'] = 'stat'
        with contextlib.redirect_stdout(io.StringIO()):
            mod = gnx.make_model(p, name='stat')
            mod.walk(10000, 'burn', verbose=False)
            spp = mod.comm[0]
            tl = int(spp.gen_arch.traits[0].loci[0])
            nl = int([l for l in range(spp.gen_arch.L) if l != tl][0])
            rows = []
            for t in range(T):
                mod.walk(1, 'main', verbose=False)
                if t in SAMPLE_T:
                    x = np.array([i.x for i in spp.values()])
                    g = np.stack([i.g for i in spp.values()])
                    rows.append(summarise(x, g, tl, nl))
        out.append(dict(Nt=np.array(spp.Nt[-T:]), stats=np.array(rows), burn_steps=mod.burn_t + 1))
        print('rep', rep, 'N', spp.Nt[-1], 'burn', mod.burn_t + 1, flush=True)
    np.savez_compressed(out_path or os.path.join(HERE, 'stat_reference.npz'),
                        Nt=np.stack([o['Nt'] for o in out]), stats=np.stack([o['stats'] for o in out]),
                        burn_steps=np.array([o['burn_steps'] for o in out]), sample_t=np.array(SAMPLE_T))


if __name__ == '__main__':
    # optional: seed base and output path (extra samples to size the reference's own sampling noise)
    run_reference(int(sys.argv[1]) if len(sys.argv) > 1 else 100,
                  int(sys.argv[2]) if len(sys.argv) > 2 else 1000, sys.argv[3] if len(sys.argv) > 3 else None)
