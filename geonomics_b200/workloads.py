"""Synthetic workloads of BASELINE.json `configs`, built as plain
numpy inputs: landscape rasters, species parameters, genomic architecture and an initial
population.  Used by bench.py, the scale tests and the CPU baseline (same inputs, scaled).
No dataset or checkpoint is involved: everything is generated from a seed.
"""
import numpy as np

CONFIGS = {
    # configs[1]: 1M individuals, 100 non-neutral loci, 2 traits, 1024x1024
    'c2': dict(dim=(1024, 1024), N=1 << 20, L=100, n_traits=2, loci_per_trait=50, mating_radius=2.0,
               b=0.2, lam=1, R=0.5, phi=0.05, gamma=1.0, n_paths=10000, recomb_rate=0.5, seed=42,
               surfaces=False),
    # configs[2]: one replicate of the 64 x 200k replicate study, 512x512
    'c3': dict(dim=(512, 512), N=200000, L=100, n_traits=1, loci_per_trait=10, mating_radius=2.0,
               b=0.2, lam=1, R=0.5, phi=0.05, gamma=1.0, n_paths=10000, recomb_rate=0.5, seed=1000,
               surfaces=False),
    # configs[3] on ONE GPU: 10M individuals, 1000 loci (100 non-neutral), conductance-surface
    # movement + dispersal sampled on the fly, 4096x4096
    'c4': dict(dim=(4096, 4096), N=10_000_000, L=1000, n_traits=2, loci_per_trait=50, mating_radius=2.0,
               b=0.2, lam=1, R=0.5, phi=0.05, gamma=1.0, n_paths=10000, recomb_rate=0.5, seed=2024,
               surfaces=True),
    # one species of configs[4]: 500k individuals, 10 000 loci at r = 1e-3 (20 of them trait loci),
    # 1024x1024 -- the long-genome shape (1280-byte homologues); tskit recording is measured
    # separately (tests/test_cuda_tskit.py), the reference cannot run this config at all
    'c5': dict(dim=(1024, 1024), N=500_000, L=10_000, n_traits=1, loci_per_trait=20, mating_radius=2.0,
               b=0.2, lam=1, R=0.5, phi=0.05, gamma=1.0, n_paths=10000, recomb_rate=1e-3, seed=5,
               surfaces=False),
}


def scaled(cfg, n):
    """Same per-capita parameters on a smaller square landscape holding ~n individuals at
    the same density (used for the bounded CPU-baseline sample and quick tests)."""
    c = dict(cfg)
    dens = cfg['N'] / float(cfg['dim'][0] * cfg['dim'][1])
    side = max(20, int(round(np.sqrt(n / dens))))
    c['dim'] = (side, side)
    c['N'] = int(round(dens * side * side))
    return c


def smooth_field(dim, seed, n_waves=8):
    """lyr_0 of C4: sum of random 2-D cosines rescaled to [0.1, 1]."""
    rng = np.random.default_rng(seed)
    X, Y = dim
    jj = np.arange(X)[None, :]
    ii = np.arange(Y)[:, None]
    r = np.zeros((Y, X))
    for _ in range(n_waves):
        kx, ky = rng.uniform(0.5, 4, 2) * 2 * np.pi / max(dim)
        r += np.cos(kx * jj + ky * ii + rng.uniform(0, 2 * np.pi))
    r = (r - r.min()) / (r.max() - r.min())
    return 0.1 + 0.9 * r


def build(cfg, seed=None):
    """Returns dict(land_dim, rasters, prm, gen_arch, pop) for DeviceSpecies / the oracle."""
    seed = cfg['seed'] if seed is None else seed
    rng = np.random.default_rng(seed)
    X, Y = cfg['dim']
    N, L = cfg['N'], cfg['L']
    if cfg['surfaces']:
        lyr0 = smooth_field(cfg['dim'], 7)
    else:
        lyr0 = np.ones((Y, X))
    lyr1 = np.tile(np.linspace(0, 1, X), (Y, 1))
    lyr2 = np.tile(np.linspace(0, 1, Y)[:, None], (1, X))
    rasters = np.stack([lyr0, lyr1, lyr2]).astype(np.float64)
    K_factor = N / float(lyr0.sum())
    traits = []
    nn = cfg['n_traits'] * cfg['loci_per_trait']
    nonneut = np.sort(rng.choice(L, size=nn, replace=False))
    perm = rng.permutation(nn)
    for t in range(cfg['n_traits']):
        loci = np.sort(nonneut[perm[t * cfg['loci_per_trait']:(t + 1) * cfg['loci_per_trait']]])
        alpha = np.clip(rng.normal(0, 0.1, len(loci)), -0.25, 0.25)
        traits.append(dict(loci=loci.astype(np.int64), alpha=alpha, phi=cfg['phi'], gamma=cfg['gamma'],
                           lyr_num=1 + (t % 2), univ_adv=False))
    # cached recombination paths (genome.py:188-215): cumsum(Bernoulli(rate)) % 2, rate[0] = 0
    rates = np.full(L, cfg['recomb_rate'])
    rates[0] = 0
    ev = rng.random((cfg['n_paths'], L)) < rates[None, :]
    paths = (np.cumsum(ev, axis=1) % 2).astype(np.uint8)
    gen_arch = dict(L=L, paths=paths, traits=traits, dom=np.zeros(L, np.int8))
    prm = dict(b=cfg['b'], R=cfg['R'], lam=cfg['lam'], n_births_fixed=True,
               mating_radius=cfg['mating_radius'], d_min=0.0, d_max=1.0, sex=False, sex_ratio_p=0.5,
               max_age=None, K_layer=0, K_factor=K_factor, move=True, move_distr=('wald', 1.0, 1.0),
               disp_distr=('wald', 1.0, 1.0), direction_mu=0.0, direction_kappa=0.0,
               density_grid_window_width=None)
    if cfg['surfaces']:
        prm['move_surf'] = dict(layer=0, mixture=True, kappa=12.0)
        prm['disp_surf'] = dict(layer=0, mixture=True, kappa=12.0)
    x = rng.uniform(0, X - 0.001, N)
    y = rng.uniform(0, Y - 0.001, N)
    pop = dict(x=x, y=y, age=rng.integers(0, 5, N).astype(np.int32),
               sex=rng.integers(0, 2, N).astype(np.int8), idx=np.arange(N, dtype=np.int64))
    return dict(land_dim=(X, Y), rasters=rasters, prm=prm, gen_arch=gen_arch, pop=pop, L=L)


def random_packed_genomes(n, L, seed, p=0.5):
    """start_p_fixed = 0.5 genotypes, generated directly in the packed device layout
    u32[n, 2, W] (genome.py:1108-1157 assigns round(2N*p) mutated homologues per locus; here
    every bit is Bernoulli(p), which has the same expectation)."""
    from .genome_pack import words_per_hap
    W = words_per_hap(L)
    rng = np.random.default_rng(seed)
    if p == 0.5:
        g = rng.integers(0, 1 << 32, size=(n, 2, W), dtype=np.uint64).astype(np.uint32)
    else:
        bits = rng.random((n, 2, W * 32)) < p
        g = np.packbits(bits, axis=-1, bitorder='little').view(np.uint32).reshape(n, 2, W)
    # clear padding bits beyond L
    full, rem = divmod(L, 32)
    if rem:
        g[:, :, full] &= np.uint32((1 << rem) - 1)
        g[:, :, full + 1:] = 0
    else:
        g[:, :, full:] = 0
    return g
