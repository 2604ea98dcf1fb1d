#!/usr/bin/env python
"""Golden-vector generator (TEST INFRASTRUCTURE; runs only in the build container).

Imports the *unmodified* reference (erthward/geonomics v1.4.9 at /root/reference, via
oracle/ref_shims.py), builds small models through its public API
(make_parameters_file -> read_parameters_file -> make_model -> walk), and records one
full main time step of the reference's own code

    Species._set_age_stage -> Species._do_movement -> Species._do_pop_dynamics
    (sim/model.py:603-667 queue order)

with every numpy.random call site on that path replaced by a
replay of pre-generated draws, capturing the inputs and outputs of each stage.  The
vectors are committed as tests/golden/*.npz; tests/test_oracle_golden.py pins the numpy
oracle to them on CPU and tests/test_cuda_parity.py checks the CUDA path against them on
the GPU box (where /root/reference does not exist).

The only change made to the reference's data flow is a *re-ordering* of the pair list
returned by Species._find_mating_pairs into the canonical order (ascending focal
ordinal, column 0 = focal; reciprocal couples kept under the smaller focal): the
reference's own order comes from Python set/frozenset hashing (mating.py:63).  The raw
reference pair set is recorded too and compared set-wise.

Usage:  python tests/golden/make_golden.py [case ...]
"""
import os
import sys
import warnings
import contextlib
import io

# numpy's float16 cos/sin (conductance-surface directions, movement.py:75-76) are
# CPU-dependent: with AVX512 dispatch they come from a low-precision vector routine.  Record
# the vectors on the portable path (half -> float -> libm -> half), i.e. what the reference
# computes on any CPU without AVX512.  Must be set before numpy is imported.
os.environ.setdefault('NPY_DISABLE_CPU_FEATURES',
                      'AVX512F AVX512CD AVX512_SKX AVX512_CLX AVX512_CNL AVX512_ICL AVX512_SPR')

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
warnings.filterwarnings('ignore')

from oracle import ref_shims           # noqa: E402
from oracle import step_oracle as so   # noqa: E402

MAX_TRIES = 6       # dispersal draws recorded per offspring


# ------------------------------------------------------------------------------------------
# cases
# ------------------------------------------------------------------------------------------
def gradient(dim, axis):
    X, Y = dim
    if axis == 'x':
        return np.tile(np.linspace(0, 1, X), (Y, 1))
    return np.tile(np.linspace(0, 1, Y)[:, None], (1, X))


def bumpy(dim, seed):
    rng = np.random.default_rng(seed)
    X, Y = dim
    jj, ii = np.meshgrid(np.arange(X), np.arange(Y))
    r = np.zeros((Y, X))
    for _ in range(6):
        kx, ky = rng.uniform(0.5, 3, 2) * 2 * np.pi / max(dim)
        r += np.cos(kx * jj + ky * ii + rng.uniform(0, 2 * np.pi))
    r = (r - r.min()) / (r.max() - r.min())
    return 0.1 + 0.9 * r


CASES = {
    # unsexed, fixed births, vonmises(0,0)+wald movement, 2 polygenic traits
    'base': dict(dim=(40, 40), N=900, K_factor=0.8, L=100, n_traits=2, trait_loci=[6, 5],
                 mating_radius=2, b=0.4, sex=False, n_births_fixed=True, lam=1,
                 move=('wald', 1.0, 1.0), disp=('wald', 0.8, 1.0), kappa=0.0, mu=0.0,
                 dom=False, max_age=None, phi=[0.1, 0.05], gamma=[1, 2], seed=11,
                 surfaces=False, main_steps=4),
    # sexed, poisson births, lognormal movement, directional vonmises, dominance,
    # max_age, non-square landscape, monogenic + polygenic traits, univ_adv trait
    'sexed': dict(dim=(50, 30), N=1100, K_factor=2.0, L=40, force_burn=12, n_traits=2, trait_loci=[1, 7],
                  mating_radius=3, b=0.9, sex=True, n_births_fixed=False, lam=2,
                  move=('lognormal', 0.0, 0.5), disp=('lognormal', -0.5, 0.3), kappa=2.0,
                  mu=0.7, dom=True, max_age=9, phi=[0.2, 0.1], gamma=[1, 1], seed=12,
                  surfaces=False, univ_adv=[False, True], main_steps=4),
    # conductance surfaces for movement and dispersal (float16 direction tables)
    'surf': dict(dim=(24, 24), N=500, K_factor=3.0, L=130, n_traits=1, trait_loci=[8],
                 mating_radius=2.5, b=0.5, sex=False, n_births_fixed=True, lam=1,
                 move=('wald', 0.6, 1.0), disp=('wald', 0.6, 1.0), kappa=0.0, mu=0.0,
                 dom=False, max_age=None, phi=[0.1], gamma=[1], seed=13,
                 surfaces=True, main_steps=3),
    # a burn-in step: no genomes, no selection (species.py:825-830, 624, 666-672)
    'burn': dict(dim=(30, 30), N=700, K_factor=1.0, L=20, n_traits=1, trait_loci=[4],
                 mating_radius=2, b=0.5, sex=False, n_births_fixed=True, lam=1,
                 move=('wald', 1.0, 1.0), disp=('wald', 0.8, 1.0), kappa=0.0, mu=0.0,
                 dom=False, max_age=None, phi=[0.1], gamma=[1], seed=14,
                 surfaces=False, main_steps=0, burn_case=6),
    # non-default density window (spatial.py:270-360 with window_width = 2.5), spatially varying
    # phi (genome.py:391-396), non-integer gamma, non-square landscape
    'misc': dict(dim=(32, 24), N=700, K_factor=1.0, L=60, n_traits=2, trait_loci=[6, 3],
                 mating_radius=2, b=0.5, sex=False, n_births_fixed=True, lam=1,
                 move=('wald', 1.0, 1.0), disp=('wald', 0.8, 1.0), kappa=0.0, mu=0.0,
                 dom=False, max_age=None, phi=['raster', 0.08], gamma=[1.5, 1], seed=19,
                 surfaces=False, main_steps=3, window_width=2.5, res=(2, 1)),
    # nearest-neighbour mating (spatial.py:194-203) and inverse-distance-weighted mate choice
    # (spatial.py:209-229)
    'nearest': dict(dim=(30, 30), N=700, K_factor=0.8, L=40, n_traits=1, trait_loci=[5],
                    mating_radius=2.5, b=0.5, sex=False, n_births_fixed=True, lam=1,
                    move=('wald', 1.0, 1.0), disp=('wald', 0.8, 1.0), kappa=0.0, mu=0.0,
                    dom=False, max_age=None, phi=[0.1], gamma=[1], seed=17,
                    surfaces=False, main_steps=3, choose_nearest=True),
    'invdist': dict(dim=(30, 30), N=700, K_factor=0.8, L=40, n_traits=1, trait_loci=[5],
                    mating_radius=2.5, b=0.5, sex=False, n_births_fixed=True, lam=1,
                    move=('wald', 1.0, 1.0), disp=('wald', 0.8, 1.0), kappa=0.0, mu=0.0,
                    dom=False, max_age=None, phi=[0.1], gamma=[1], seed=18,
                    surfaces=False, main_steps=3, inverse_dist=True),
    # Wright-Fisher style panmixia (mating_radius = None, species.py:2178-2194)
    'pan': dict(dim=(30, 30), N=600, K_factor=0.7, L=50, n_traits=1, trait_loci=[6],
                mating_radius=None, b=0.3, sex=False, n_births_fixed=True, lam=1,
                move=('wald', 1.0, 1.0), disp=('wald', 0.8, 1.0), kappa=0.0, mu=0.0,
                dom=False, max_age=None, phi=[0.1], gamma=[1], seed=16,
                surfaces=False, main_steps=3),
    # neutral + deleterious mutation (ops/mutation.py), use_tskit=False; recorded after a few
    # main steps so that earlier deleterious loci already enter the fitness
    'mut': dict(dim=(40, 40), N=900, K_factor=0.8, L=400, n_traits=2, trait_loci=[5, 4],
                mating_radius=2, b=0.4, sex=False, n_births_fixed=True, lam=1,
                move=('wald', 1.0, 1.0), disp=('wald', 0.8, 1.0), kappa=0.0, mu=0.0,
                dom=False, max_age=None, phi=[0.1, 0.05], gamma=[1, 2], seed=15,
                surfaces=False, main_steps=5, mu_neut=1e-5, mu_delet=1e-5, model_T=10, mut_n=7),
    # use_tskit=True (species.py:891-905: genotype arrays hold the non-neutral loci only) with
    # neutral + deleterious + TRAIT mutation (mutation.py:90-131, genome.py:416-437, 753-788) and
    # the tskit rows of a step (species.py:692-736, mutation.py:44-58).  The reference runs on the
    # functional table shim of oracle/ref_shims.py; free-running mutations of the earlier main
    # steps leave the index arrays (Trait.loci_idxs, delet_loci_idxs) as the reference leaves them.
    'tmut': dict(dim=(40, 40), N=900, K_factor=0.8, L=400, n_traits=2, trait_loci=[5, 4],
                 mating_radius=2, b=0.4, sex=False, n_births_fixed=True, lam=1,
                 move=('wald', 1.0, 1.0), disp=('wald', 0.8, 1.0), kappa=0.0, mu=0.0,
                 dom=False, max_age=None, phi=[0.1, 0.05], gamma=[1, 2], seed=21,
                 surfaces=False, main_steps=6, mu_neut=1e-5, mu_delet=1.5e-5, trait_mu=[2e-5, 1.5e-5],
                 model_T=10, mut_n=10, use_tskit=True),
}


def build_params(gnx, case, tmpdir='/tmp'):
    c = CASES[case]
    path = os.path.join(tmpdir, 'gnx_golden_%s.py' % case)
    layers = [{'type': 'defined', 'change': False} for _ in range(3)]
    spp = [{'movement': True, 'movement_surface': c['surfaces'],
            'dispersal_surface': c['surfaces'], 'genomes': True,
            'n_traits': c['n_traits'], 'demographic_change': 0, 'parameter_change': False}]
    gnx.make_parameters_file(path, layers=layers, species=spp)
    txt = open(path).read()
    open(path, 'w').write('import numpy as np\n' + txt)     # params.py:180 quirk
    p = gnx.read_parameters_file(path)
    dim = c['dim']
    p['landscape']['main']['dim'] = dim
    if c.get('res') is not None:
        p['landscape']['main']['res'] = c['res']          # non-square cells: movement distances scale per axis
    rasts = [bumpy(dim, c['seed']), gradient(dim, 'x'), gradient(dim, 'y')]
    for n, r in enumerate(rasts):
        p['landscape']['layers']['lyr_%i' % n]['init']['defined']['rast'] = r
    s = p['comm']['species']['spp_0']
    s['init']['N'] = c['N']
    s['init']['K_layer'] = 'lyr_0'
    s['init']['K_factor'] = c['K_factor']
    m = s['mating']
    m['sex'] = c['sex']
    m['sex_ratio'] = 0.5 if c['sex'] else 1 / 1
    m['b'] = c['b']
    m['R'] = 0.5
    m['n_births_distr_lambda'] = c['lam']
    m['n_births_fixed'] = c['n_births_fixed']
    m['mating_radius'] = c['mating_radius']
    m['choose_nearest_mate'] = bool(c.get('choose_nearest', False))
    m['inverse_dist_mating'] = bool(c.get('inverse_dist', False))
    s['mortality']['max_age'] = c['max_age']
    if c.get('window_width') is not None:
        s['mortality']['density_grid_window_width'] = c['window_width']
    mv = s['movement']
    mv['direction_distr_mu'] = c['mu']
    mv['direction_distr_kappa'] = c['kappa']
    mv['movement_distance_distr'] = c['move'][0]
    mv['movement_distance_distr_param1'] = c['move'][1]
    mv['movement_distance_distr_param2'] = c['move'][2]
    mv['dispersal_distance_distr'] = c['disp'][0]
    mv['dispersal_distance_distr_param1'] = c['disp'][1]
    mv['dispersal_distance_distr_param2'] = c['disp'][2]
    if c['surfaces']:
        for k in ('move_surf', 'disp_surf'):
            mv[k]['layer'] = 'lyr_0'
            mv[k]['approx_len'] = 120
            mv[k]['mixture'] = True
    g = s['gen_arch']
    g['L'] = c['L']
    g['use_tskit'] = bool(c.get('use_tskit', False))
    g['tskit_simp_interval'] = 10 ** 6        # sort()/simplify() are tskit's own algorithms: never reached
    g['jitter_breakpoints'] = False
    g['dom'] = c['dom']
    g['n_recomb_sims'] = 1000
    g['r_distr_alpha'] = 0.5
    g['r_distr_beta'] = 0.5 if case == 'sexed' else None
    for t in range(c['n_traits']):
        tr = g['traits']['trait_%i' % t]
        tr['layer'] = 'lyr_%i' % (1 + t % 2)
        tr['n_loci'] = c['trait_loci'][t]
        tr['phi'] = (0.02 + 0.1 * gradient(dim, 'y')) if c['phi'][t] == 'raster' else c['phi'][t]
        tr['gamma'] = c['gamma'][t]
        tr['alpha_distr_mu'] = 0.0 if c['trait_loci'][t] > 1 else 0.1
        tr['alpha_distr_sigma'] = 0.15 if c['trait_loci'][t] > 1 else 0
        tr['max_alpha_mag'] = 0.3
        tr['univ_adv'] = c.get('univ_adv', [False] * 4)[t]
        if c.get('trait_mu') is not None:
            tr['mu'] = c['trait_mu'][t]
    p['model']['T'] = c.get('model_T', 100)
    if c.get('mu_neut') is not None:
        g['mu_neut'] = c['mu_neut']
        g['mu_delet'] = c['mu_delet']
        g['delet_alpha_distr_shape'] = 0.2
        g['delet_alpha_distr_scale'] = 0.2
    p['model']['burn_T'] = c.get('force_burn', 20)
    p['model']['seed'] = {'num': c['seed']}
    return p


# ------------------------------------------------------------------------------------------
# state capture
# ------------------------------------------------------------------------------------------
def capture_state(spp):
    inds = list(spp.values())
    st = dict(
        idx=np.array([i.idx for i in inds], dtype=np.int64),
        x=np.array([i.x for i in inds], dtype=np.float64),
        y=np.array([i.y for i in inds], dtype=np.float64),
        age=np.array([i.age for i in inds], dtype=np.int32),
        sex=np.array([i.sex for i in inds], dtype=np.int8),
    )
    nt = len(spp.gen_arch.traits) if spp.gen_arch.traits is not None else 0
    if len(inds) and all(i.g is not None for i in inds):
        st['g'] = np.stack([np.int8(i.g) for i in inds])
        st['z'] = np.array([i.z for i in inds], dtype=np.float64).reshape(len(inds), -1)
    else:       # burn-in: individuals carry no genomes yet (species.py:666-672)
        st['g'] = np.zeros((len(inds), spp.gen_arch.L, 2), np.int8)
        st['z'] = np.zeros((len(inds), nt))
    st['fit'] = np.array([np.nan if i.fit is None else i.fit for i in inds], dtype=np.float64)
    if spp.gen_arch.use_tskit and len(inds) and len(inds[0]._nodes_tab_ids) == 2:
        st['nodes'] = np.array([[i._nodes_tab_ids[0], i._nodes_tab_ids[1]] for i in inds], dtype=np.int64)
        st['ind_row'] = np.array([i._individuals_tab_id for i in inds], dtype=np.int64)
    return st


def capture_arch(spp, land):
    ga = spp.gen_arch
    out = {}
    L = ga.L
    n_sims = ga.recombinations._n
    paths = np.zeros((n_sims, L), dtype=np.uint8)
    if ga.use_tskit:
        # the subsetters cover the genotype rows (non-neutral loci) only: genome.py:215-218, 133-160;
        # the full-length paths are the cumulated breakpoints (genome.py:211-212)
        nn = len(ga.nonneut_loci)
        subs = np.zeros((n_sims, nn), dtype=np.uint8)
        for k in range(n_sims):
            sub = list(ga.recombinations._subsetters[k])
            assert len(sub) == 2 * nn
            subs[k] = np.array(sub[1::2], dtype=np.uint8)
            rec_at = np.zeros(L, dtype=np.int64)
            rec_at[np.asarray(ga.recombinations._breakpoints[k], dtype=np.int64)] = 1
            paths[k] = np.cumsum(rec_at) % 2
        out['subsetters'] = subs
    else:
        for k in range(n_sims):
            sub = list(ga.recombinations._subsetters[k])
            paths[k] = np.array(sub[1::2], dtype=np.uint8)     # '10'->hom 0, '01'->hom 1
    out['paths'] = paths
    out['use_tskit'] = np.int64(bool(ga.use_tskit))
    out['dom'] = np.asarray(ga.dom, dtype=np.int8)
    out['n_traits'] = np.int64(len(ga.traits))
    for t, tr in ga.traits.items():
        out['trait%i_loci' % t] = np.asarray(tr.loci, dtype=np.int64)
        out['trait%i_alpha' % t] = np.asarray(tr.alpha, dtype=np.float64)
        out['trait%i_phi' % t] = np.asarray(tr.phi, dtype=np.float64)
        out['trait%i_gamma' % t] = np.float64(tr.gamma)
        out['trait%i_lyr' % t] = np.int64(tr.lyr_num)
        out['trait%i_univ_adv' % t] = np.int64(bool(tr.univ_adv))
        if ga.use_tskit:
            out['trait%i_loci_idxs' % t] = np.asarray(tr.loci_idxs, dtype=np.int64)
            out['trait%i_alpha_distr' % t] = np.array(
                [tr.alpha_distr_mu, tr.alpha_distr_sigma,
                 -1.0 if tr.max_alpha_mag is None else tr.max_alpha_mag], dtype=np.float64)
    out['rasters'] = np.stack([land[l].rast for l in range(len(land))]).astype(np.float64)
    out['K'] = np.asarray(spp.K, dtype=np.float64)
    out['land_dim'] = np.array(land.dim, dtype=np.int64)
    out['res_ratio'] = np.array(land._res_ratio, dtype=np.float64)
    out['ww'] = np.float64(spp._dens_grids.window_width)
    if spp._move_surf is not None:
        out['move_surf'] = np.asarray(spp._move_surf.surf)          # float16 [Y, X, A]
        out['disp_surf'] = np.asarray(spp._disp_surf.surf)
    if getattr(spp, 'mutate', False):
        out['mut_mu_neut'] = np.float64(ga.mu_neut)
        out['mut_mu_delet'] = np.float64(ga.mu_delet)
        out['mut_trait_mus'] = np.array([tr.mu for tr in ga.traits.values()], dtype=np.float64)
        out['mut_mutables'] = np.array(ga._mutables, dtype=np.int64)
        out['mut_nonneut_loci'] = np.array(ga.nonneut_loci, dtype=np.int64)
        out['mut_delet_loci'] = np.array(ga.delet_loci, dtype=np.int64)
        out['mut_delet_s'] = np.array(ga.delet_loci_s, dtype=np.float64)
        out['mut_s_shape'] = np.float64(ga.delet_alpha_distr_shape)
        out['mut_s_scale'] = np.float64(ga.delet_alpha_distr_scale)
        if ga.use_tskit:
            out['mut_delet_loci_idxs'] = np.asarray(ga.delet_loci_idxs, dtype=np.int64)
    prm = {}
    for k in ('b', 'R', 'n_births_distr_lambda', 'mating_radius', 'd_min', 'd_max',
              'direction_distr_mu', 'direction_distr_kappa'):
        v = getattr(spp, k)
        prm[k] = -1.0 if v is None else float(v)          # mating_radius None (panmixia) -> -1
    prm['choose_nearest'] = int(bool(spp.choose_nearest_mate))
    prm['inverse_dist'] = int(bool(spp.inverse_dist_mating))
    prm['sex'] = int(bool(spp.sex))
    prm['sex_ratio_p'] = float(spp.sex_ratio)
    prm['n_births_fixed'] = int(bool(spp.n_births_fixed))
    prm['max_age'] = -1 if spp.max_age is None else int(spp.max_age)
    for k, v in prm.items():
        out['prm_' + k] = np.float64(v)
    return out


# ------------------------------------------------------------------------------------------
# replay of the reference's random call sites
# ------------------------------------------------------------------------------------------
class Replay:
    def __init__(self, gnx, spp, land, draws):
        self.gnx = gnx
        self.spp = spp
        self.land = land
        self.d = draws
        self.rec = {}
        self.off = -1          # current offspring (dispersal / sex draws)
        self.tries = 0
        self.sex_phase = 0
        self.n_start = 0
        self.focals = None
        self.in_mutation = False
        self.n_mut_done = 0
        self.pan_phase = False

    # -- movement / dispersal samplers (movement.py:55-72, 111-120)
    def vonmises(self, mu, kappa, size=None):
        if size is not None:
            return self.d['move_dir'][:size].copy()
        v = self.d['disp_dir'][self.off, self.tries]
        return v

    def dist(self, *a, size=None, **k):
        if size is not None:
            return self.d['move_dist'][:size].copy()
        v = self.d['disp_dist'][self.off, self.tries]
        self.tries += 1
        return v

    def randint(self, low=0, high=None, size=None):
        n_sims = self.spp.gen_arch.recombinations._n
        if high == n_sims:                                   # species.py:625
            return self.d['recomb_keys'][:size].copy()
        # conductance-surface lookups (spatial.py:183)
        if size == 1:
            return np.array([self.d['disp_choice'][self.off, self.tries]])
        return self.d['move_choice'][:size].copy()

    def gamma(self, shape, scale=1.0, size=None):          # genome.py:691
        return float(self.d['mut_s'][self.n_mut_done])

    def normal(self, loc=0.0, scale=1.0, size=None):       # genome.py:679 _draw_trait_alpha
        assert self.in_mutation
        return np.array([self.d['mut_alpha'][self.n_mut_done]], dtype=np.float64)

    def choice(self, opts, *a, **k):
        if self.pan_phase:
            # species.py:2189: 2 * n_mates individuals with replacement; slot i (pan_u[i] < b) draws
            # the ordinals (pan_R[i, 0] * N) >> 32 and (pan_R[i, 1] * N) >> 32
            n = len(opts)
            act = np.nonzero(self.d['pan_u'][:n] < self.spp.b)[0]
            R = self.d['pan_R'][act].astype(np.uint64)
            ac = ((R * np.uint64(n)) >> np.uint64(32)).astype(np.int64)
            assert k.get('size') == 2 * len(act)
            return ac.reshape(-1)
        if 'p' in k and not self.in_mutation:
            # inverse-distance mate choice (spatial.py:222-227): the reference lists the options by
            # ascending distance; the shared convention walks them in canonical order.  Check the
            # option set and the weights, return the oracle's pick for this focal.
            i = self.focals[self.n_choice]
            self.n_choice += 1
            opts = np.asarray(opts, dtype=np.int64)
            x, y = self.xy
            d = np.sqrt((x[opts] - x[i]) ** 2 + (y[opts] - y[i]) ** 2)
            w = self.spp.mating_radius - d
            np.testing.assert_allclose(np.asarray(k['p']), w / w.sum(), rtol=1e-9, atol=1e-12)
            self.rec.setdefault('ref_opts_focal', []).append(i)
            self.rec.setdefault('ref_opts', []).append(np.sort(opts))
            m = int(self.mate_full[i])
            assert m in opts
            return m
        if self.in_mutation:
            if 'p' in k:                                   # genome.py:662 _draw_mut_types
                cdf = np.cumsum(np.asarray(k['p'], dtype=np.float64))
                cdf /= cdf[-1]
                u = self.d['mut_type_u'][:k['size']]
                return np.asarray(opts)[np.searchsorted(cdf, u, side='right')]
            # mutation.py:66 / :101 r.choice(offspring)
            return opts[int(so.choose_k(self.d['mut_ind_R'][self.n_mut_done], len(opts)))]
        # mate choice (spatial.py:241): canonical k-th neighbour
        i = self.focals[self.n_choice]
        self.n_choice += 1
        opts = np.asarray(opts, dtype=np.int64)
        self.rec.setdefault('ref_opts_focal', []).append(i)
        self.rec.setdefault('ref_opts', []).append(np.sort(opts))
        order = opts[np.argsort(self.rank[opts], kind='stable')]
        kk = int(so.choose_k(self.d['mate_R'][i], len(order)))
        return order[kk]

    def binomial(self, n=None, p=None, size=None):
        if self.pan_phase:                                   # species.py:2183 n_mates
            return int((self.d['pan_u'][:n] < p).sum())
        if self.in_mutation:
            if n == 1:                                       # mutation.py:76 / :107 homologue
                r = int(self.d['mut_homol_u'][self.n_mut_done] < 0.5)
                self.n_mut_done += 1
                return r
            return int(self.d['mut_n'][0])                   # mutation.py:172 n_muts
        if np.ndim(p) == 1:                                  # demography.py:176
            u = self.d['death_u'][:len(p)]
            self.rec['death_p'] = np.array(p, dtype=np.float64)
            return (u < p).astype(np.int64)
        if size is not None and np.ndim(size) == 0 and size == 2 and p == 0.5 and n == 1 \
                and self.in_gametes:                         # mating.py:133
            v = self.d['start_homs'][self.n_start].copy()
            self.n_start += 1
            return v
        if size is not None:                                 # species.py:2212 can_mate
            f = np.asarray(self.focals[:size], dtype=np.int64)
            return (self.d['mate_u'][f] < p).astype(np.int64)
        # scalar sex draws (species.py:660, individual.py:115)
        if self.spp.sex and self.sex_phase == 0:
            r = int(self.d['sex_u'][self.off] < p)
            self.sex_phase = 1 if r == 0 else 0
            if r:
                self.off_sex_done()
            return r
        r = int(self.d['sex_redraw_u'][self.off] < 0.5)
        self.sex_phase = 0
        return r

    def off_sex_done(self):
        pass

    def poisson(self, lam, size=None):
        return self.d['poisson'][:size].copy()


@contextlib.contextmanager
def patched(rp):
    import numpy.random as npr
    import geonomics.ops.movement as mv
    import geonomics.ops.mating as mt
    import geonomics.ops.demography as dm
    import geonomics.structs.species as sp
    saved = []

    def setp(obj, name, val):
        saved.append((obj, name, getattr(obj, name)))
        setattr(obj, name, val)
    setp(mv, '_r_vonmises', rp.vonmises)
    setp(mv, '_wald', rp.dist)
    setp(mv, '_lognormal', rp.dist)
    setp(npr, 'randint', rp.randint)
    setp(npr, 'choice', rp.choice)
    setp(npr, 'binomial', rp.binomial)
    setp(npr, 'poisson', rp.poisson)
    setp(npr, 'gamma', rp.gamma)
    setp(npr, 'normal', rp.normal)

    # mutation stage marker (species.py:808-809)
    orig_mut = sp._do_mutation

    def do_mutation(offspring, spp, log=None):
        rp.in_mutation = True
        rp.n_mut_done = 0
        ga = spp.gen_arch
        before = len(ga._mutables)
        with contextlib.redirect_stdout(io.StringIO()):
            out = orig_mut(offspring, spp, log=log)
        rp.in_mutation = False
        rp.rec['mut_count'] = before - len(ga._mutables)
        return out
    setp(sp, '_do_mutation', do_mutation)

    # dispersal wrapper: offspring counter (species.py:645-648)
    orig_disp = sp._do_dispersal

    def disp(spp, mx, my, p1, p2, **k):
        rp.off += 1
        rp.tries = 0
        rp.sex_phase = 0
        rp.rec.setdefault('mid_x', []).append(mx)
        rp.rec.setdefault('mid_y', []).append(my)
        x, y = orig_disp(spp, mx, my, p1, p2, **k)
        rp.rec.setdefault('disp_tries', []).append(rp.tries)
        return x, y
    setp(sp, '_do_dispersal', disp)

    # gamete stage marker (species.py:636 -> mating.py:186)
    orig_mating = sp._do_mating
    rp.in_gametes = False

    def do_mating(spp, pairs, nb, keys):
        rp.in_gametes = True
        rp.rec['nb'] = np.array(nb, dtype=np.int64)
        out = orig_mating(spp, pairs, nb, keys)
        rp.in_gametes = False
        return out
    setp(sp, '_do_mating', do_mating)

    # stage recorders in demography (module-level names are looked up at call time)
    for name in ('_calc_n_pairs', '_calc_dNdt', '_calc_d', '_calc_prob_death'):
        orig = getattr(dm, name)

        def wrap(*a, _orig=orig, _name=name, **k):
            out = _orig(*a, **k)
            rp.rec[_name] = np.array(out, dtype=np.float64)
            if _name == '_calc_prob_death':
                rp.rec['fit'] = np.array([i.fit for i in rp.spp.values()], dtype=np.float64)
                rp.rec['d_ind'] = np.array(a[1], dtype=np.float64)
            return out
        setp(dm, name, wrap)
    orig_mort = dm._do_mortality

    def mort(spp, death_probs):
        for k, v in capture_state(spp).items():
            rp.rec['pre_' + k] = v
        return orig_mort(spp, death_probs)
    setp(dm, '_do_mortality', mort)
    try:
        yield
    finally:
        for obj, name, val in reversed(saved):
            setattr(obj, name, val)


def make_draws(rng, cap, spp, case):
    c = CASES[case]
    n_sims = spp.gen_arch.recombinations._n
    d = {}
    mu, kappa = c['mu'], c['kappa']
    d['move_dir'] = rng.vonmises(mu, kappa, cap) if kappa > 0 else rng.uniform(-np.pi, np.pi, cap)
    kind, p1, p2 = c['move']
    d['move_dist'] = rng.wald(p1, p2, cap) if kind == 'wald' else rng.lognormal(p1, p2, cap)
    d['move_choice'] = rng.integers(0, 120, cap)
    d['mate_R'] = rng.integers(0, 2**32, cap, dtype=np.uint64).astype(np.uint32)
    d['mate_u'] = rng.random(cap)
    d['poisson'] = rng.poisson(c['lam'], cap)
    d['recomb_keys'] = rng.integers(0, n_sims, 2 * cap)
    d['start_homs'] = rng.integers(0, 2, (cap, 2))
    d['disp_dir'] = rng.uniform(-np.pi, np.pi, (cap, MAX_TRIES))
    kind, p1, p2 = c['disp']
    d['disp_dist'] = (rng.wald(p1, p2, (cap, MAX_TRIES)) if kind == 'wald'
                      else rng.lognormal(p1, p2, (cap, MAX_TRIES)))
    d['disp_choice'] = rng.integers(0, 120, (cap, MAX_TRIES))
    d['sex_u'] = rng.random(cap)
    d['sex_redraw_u'] = rng.random(cap)
    d['death_u'] = rng.random(cap)
    if c.get('inverse_dist'):
        d['mate_inv_u'] = rng.random(cap)
    if c['mating_radius'] is None:
        d['pan_u'] = rng.random(cap)
        d['pan_R'] = rng.integers(0, 2**32, (cap, 2), dtype=np.uint64).astype(np.uint32)
    if c.get('mut_n'):
        nm = c['mut_n']
        d['mut_n'] = np.array([nm], dtype=np.int32)
        d['mut_type_u'] = rng.random(nm)
        d['mut_ind_R'] = rng.integers(0, 2**32, nm, dtype=np.uint64).astype(np.uint32)
        d['mut_homol_u'] = rng.random(nm)
        d['mut_s'] = rng.gamma(0.2, 0.2, nm)
        d['mut_alpha'] = rng.normal(0.0, 0.15, nm)
    return d


def canonical_pairs_from_ids(ref_pairs_ids, ids, mate_ord):
    """reference pair list (ids, hash order) -> canonical (ordinals)."""
    pos = {int(v): k for k, v in enumerate(ids)}
    out = []
    for a, b in np.asarray(ref_pairs_ids).reshape(-1, 2):
        i, j = pos[int(a)], pos[int(b)]
        # focal = the one whose recorded mate is the other; reciprocal -> smaller ordinal
        cand = [f for f, m in ((i, j), (j, i)) if mate_ord[f] == m]
        assert cand, 'reference pair not explained by recorded mate choices'
        f = min(cand)
        out.append((f, j if f == i else i))
    out.sort()
    return np.array(out, dtype=np.int64).reshape(-1, 2)


def record_case(gnx, case, out_dir=HERE):
    c = CASES[case]
    p = build_params(gnx, case)
    if c.get('force_burn'):
        # the sexed reference population declines steadily (75 % of newborns are male:
        # species.py:660 + individual.py:110-115), so the stationarity tests never pass;
        # end the burn-in after burn_T steps instead (burn-in control is out of scope).
        import geonomics.sim.burnin as _b
        _b._test_t_threshold = lambda *a, **k: True
        _b.SpatialTester.run_test = lambda self, n, alpha=0.05: True
    with contextlib.redirect_stdout(io.StringIO()):
        mod = gnx.make_model(p, name='golden_' + case)
        if c.get('burn_case'):
            mod.walk(c['burn_case'], 'burn', verbose=False)      # stop inside the burn-in
        else:
            mod.walk(10000, 'burn', verbose=False)
            mod.walk(c['main_steps'], 'main', verbose=False)
    spp = mod.comm[0]
    land = mod.land
    print(case, 'N =', len(spp), 'L =', spp.gen_arch.L, 'burn_t =', mod.burn_t)

    rec = capture_arch(spp, land)
    st0 = capture_state(spp)
    for k, v in st0.items():
        rec['in_' + k] = v
    rec['in_max_ind_idx'] = np.int64(spp.max_ind_idx)
    N0 = len(spp)
    cap = 2 * N0 + 64
    rng = np.random.default_rng(1000 + c['seed'])
    draws = make_draws(rng, cap, spp, case)
    for k, v in draws.items():
        rec['draw_' + k] = v

    rp = Replay(gnx, spp, land, draws)
    ids0 = st0['idx']
    if spp.gen_arch.use_tskit:
        tab0 = (spp._tc.nodes.num_rows, spp._tc.edges.num_rows, spp._tc.individuals.num_rows,
                spp._tc.mutations.num_rows)

    with patched(rp):
        # ---- a1 age, a2 movement, a3 env sample (model.py queue; species.py:567-586)
        spp._set_age_stage()
        spp._do_movement(land)
        st1 = capture_state(spp)
        rec['mv_x'] = st1['x']
        rec['mv_y'] = st1['y']
        rec['mv_age'] = st1['age']
        rec['mv_e'] = np.array([i.e for i in spp.values()], dtype=np.float64)
        rec['mv_cells'] = np.asarray(spp._cells, dtype=np.int32)

        # ---- a5/a6 mate search with canonical ordering wrapper
        x, y = st1['x'], st1['y']
        panmixia = spp.mating_radius is None
        if panmixia:
            nb_lists, mate_ord = None, None
            rec['n_nbrs'] = np.full(N0, N0 - 1, dtype=np.int32)
        else:
            rank, _, _ = so.canonical_rank(x, y, land.dim, spp.mating_radius)
            rp.rank = rank
            nb_lists = so.neighbor_lists(x, y, land.dim, spp.mating_radius)
            rp.focals = [i for i, l in enumerate(nb_lists) if len(l) > 0]
            rp.n_choice = 0
            rec['n_nbrs'] = np.array([len(l) for l in nb_lists], dtype=np.int32)

            # oracle's own prediction of the mate of every focal (needed to orient pairs)
            modes = dict(choose_nearest=bool(spp.choose_nearest_mate), inverse_dist=bool(spp.inverse_dist_mating),
                         inv_u=draws.get('mate_inv_u'))
            _, _, mate_ord = so.find_mates_radius(
                x, y, land.dim, spp.mating_radius, spp.b, draws['mate_R'], draws['mate_u'],
                sex=None, nbrs=nb_lists, **modes)
            # ... and the choice itself, whatever the Bernoulli(b) draw says (inverse-distance replay)
            _, _, rp.mate_full = so.find_mates_radius(
                x, y, land.dim, spp.mating_radius, 2.0, draws['mate_R'], draws['mate_u'],
                sex=None, nbrs=nb_lists, **modes)
            rp.xy = (x, y)
        orig_find = spp._find_mating_pairs
        holder = {}

        def find_canonical():
            rp.pan_phase = panmixia
            ref_pairs = orig_find()
            rp.pan_phase = False
            holder['ref_pairs_ids'] = np.array(ref_pairs, dtype=np.int64).reshape(-1, 2)
            if panmixia:
                # the reference folds each drawn couple through a Python set (species.py:2192), which
                # may swap the two parents; canonical = draw order (oracle), matched as a multiset
                can = so.find_mates_panmixia_draws(N0, spp.b, draws['pan_u'], draws['pan_R'], None)
                ref_unordered = sorted(tuple(sorted(p)) for p in holder['ref_pairs_ids'].tolist())
                ours = sorted(tuple(sorted(p)) for p in ids0[can].tolist())
                assert ref_unordered == ours, 'panmixia pairs differ from the reference'
            elif spp.sex:
                # sexed: reference keeps (female focal, male mate) rows in focal order
                pos = {int(v): k for k, v in enumerate(ids0)}
                can = np.array([(pos[int(a)], pos[int(b)]) for a, b in
                                holder['ref_pairs_ids']], dtype=np.int64).reshape(-1, 2)
            else:
                can = canonical_pairs_from_ids(ref_pairs, ids0, mate_ord)
            holder['pairs'] = can
            return ids0[can] if len(can) else np.array([])
        spp._find_mating_pairs = find_canonical

        # ---- a7..a16: the reference's own _do_pop_dynamics
        spp._do_pop_dynamics(land)
        del spp._find_mating_pairs
        spp._set_Nt()

    # the reference's own neighbour sets (spatial.py:232-236), CSR over focal ordinals
    counts = np.zeros(N0, dtype=np.int32)
    for f, o in zip(rp.rec.get('ref_opts_focal', []), rp.rec.get('ref_opts', [])):
        counts[f] = len(o)
        assert np.array_equal(o, np.sort(nb_lists[f])), 'oracle neighbour set != reference'
    # (nearest-neighbour mode goes through cKDTree.query and never lists the neighbour sets)
    assert panmixia or c.get('choose_nearest') or np.array_equal(counts, rec['n_nbrs']), \
        'oracle neighbour counts != reference'
    rec['ref_nbr_indptr'] = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
    rec['ref_nbr_indices'] = (np.concatenate(rp.rec['ref_opts']).astype(np.int32)
                              if rp.rec.get('ref_opts') else np.zeros(0, np.int32))
    for k in list(rp.rec):
        if k.startswith('pre_'):
            rec[k] = rp.rec[k]
    rec['ref_pairs_ids'] = holder['ref_pairs_ids']
    rec['pairs'] = holder['pairs']
    rec['nb'] = rp.rec.get('nb', np.full(len(holder['pairs']), int(c['lam']), dtype=np.int64))
    rec['prm_burn'] = np.float64(1.0 if c.get('burn_case') else 0.0)
    B = int(rec['nb'].sum())
    rec['B'] = np.int64(B)
    rec['mid_x'] = np.array(rp.rec.get('mid_x', []), dtype=np.float64)
    rec['mid_y'] = np.array(rp.rec.get('mid_y', []), dtype=np.float64)
    rec['disp_tries'] = np.array(rp.rec.get('disp_tries', []), dtype=np.int32)
    assert rec['disp_tries'].max(initial=0) <= MAX_TRIES
    rec['n_pairs_rast'] = rp.rec['_calc_n_pairs']
    rec['N_rast'] = np.asarray(spp.N, dtype=np.float64)
    rec['dNdt_rast'] = rp.rec['_calc_dNdt']
    rec['d_rast'] = rp.rec['_calc_d']
    rec['death_p'] = rp.rec['death_p']
    if 'fit' in rp.rec:                      # absent in a burn-in step (no selection)
        rec['fit_all'] = rp.rec['fit']       # fitness of the N0+B individuals alive before mortality
        rec['d_ind'] = rp.rec['d_ind']
    st2 = capture_state(spp)
    for k, v in st2.items():
        rec['out_' + k] = v
    rec['out_max_ind_idx'] = np.int64(spp.max_ind_idx)
    rec['out_Nt'] = np.int64(spp.Nt[-1])
    rec['out_n_births'] = np.int64(spp.n_births[-1])
    rec['out_n_deaths'] = np.int64(spp.n_deaths[-1])
    rec['out_e'] = np.array([i.e for i in spp.values()], dtype=np.float64)
    if getattr(spp, 'mutate', False):
        ga = spp.gen_arch
        assert rp.rec['mut_count'] == c['mut_n']
        rec['out_mut_mutables'] = np.array(ga._mutables, dtype=np.int64)
        rec['out_mut_nonneut_loci'] = np.array(ga.nonneut_loci, dtype=np.int64)
        rec['out_mut_delet_loci'] = np.array(ga.delet_loci, dtype=np.int64)
        rec['out_mut_delet_s'] = np.array(ga.delet_loci_s, dtype=np.float64)
        if ga.use_tskit:
            rec['out_mut_delet_loci_idxs'] = np.asarray(ga.delet_loci_idxs, dtype=np.int64)
            for t, tr in ga.traits.items():
                rec['out_trait%i_loci' % t] = np.asarray(tr.loci, dtype=np.int64)
                rec['out_trait%i_alpha' % t] = np.asarray(tr.alpha, dtype=np.float64)
                rec['out_trait%i_loci_idxs' % t] = np.asarray(tr.loci_idxs, dtype=np.int64)
            n_sims = ga.recombinations._n
            subs = np.zeros((n_sims, len(ga.nonneut_loci)), dtype=np.uint8)
            for k in range(n_sims):
                subs[k] = np.array(list(ga.recombinations._subsetters[k])[1::2], dtype=np.uint8)
            rec['out_subsetters'] = subs
    if spp.gen_arch.use_tskit:
        # rows the step appended to the tables (species.py:692-736, mutation.py:44-58)
        tc = spp._tc
        n0, e0, i0, m0 = tab0
        rec['tsk_in_rows'] = np.array(tab0, dtype=np.int64)
        rec['tsk_node_time'] = np.array(tc.nodes.column('time')[n0:], dtype=np.float64)
        rec['tsk_node_flags'] = np.array(tc.nodes.column('flags')[n0:], dtype=np.int64)
        rec['tsk_node_individual'] = np.array(tc.nodes.column('individual')[n0:], dtype=np.int64)
        rec['tsk_edge_left'] = np.array(tc.edges.column('left')[e0:], dtype=np.float64)
        rec['tsk_edge_right'] = np.array(tc.edges.column('right')[e0:], dtype=np.float64)
        rec['tsk_edge_parent'] = np.array(tc.edges.column('parent')[e0:], dtype=np.int64)
        rec['tsk_edge_child'] = np.array(tc.edges.column('child')[e0:], dtype=np.int64)
        rec['tsk_ind_location'] = np.array(tc.individuals.column('location')[i0:], dtype=np.float64)
        rec['tsk_ind_idx'] = np.array([int.from_bytes(b, 'little') for b in tc.individuals.column('metadata')[i0:]],
                                      dtype=np.int64)
        rec['tsk_mut_site'] = np.array(tc.mutations.column('site')[m0:], dtype=np.int64)
        rec['tsk_mut_node'] = np.array(tc.mutations.column('node')[m0:], dtype=np.int64)
        rec['tsk_mut_time'] = np.array(tc.mutations.column('time')[m0:], dtype=np.float64)
        rec['tsk_t'] = np.int64(spp.t)
    # trim draw arrays to what can be consumed (keeps fixtures small)
    nmax = N0 + B + 8
    for k in list(rec):
        if k.startswith('draw_'):
            rec[k] = rec[k][:2 * nmax] if k == 'draw_recomb_keys' else rec[k][:nmax]
    path = os.path.join(out_dir, 'step_%s.npz' % case)
    np.savez_compressed(path, **rec)
    print('  wrote', path, '%.1f KB' % (os.path.getsize(path) / 1024),
          'N0 =', N0, 'pairs =', len(rec['pairs']), 'B =', B, 'deaths =', int(rec['out_n_deaths']))


if __name__ == '__main__':
    gnx = ref_shims.install()
    cases = sys.argv[1:] or list(CASES)
    for c in cases:
        record_case(gnx, c)
