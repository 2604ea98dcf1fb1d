"""Drop-in adapter for a *reference* `geonomics.Species` (erthward/geonomics v1.4.9).

`attach(spp, land)` moves one Species of an already-built reference Model onto the GPU and
replaces the queue entries of `Model._make_fn_queue` (sim/model.py:603-667) that act on it --
`_set_age_stage` (species.py:567), `_do_movement` (:582), `_do_pop_dynamics` (:822),
`_set_Nt` (:554) -- with calls into libgnxb200.so.  It is written against the reference's own
attribute names, so it serves a real `geonomics.Species` in a container that has both
packages; `species_to_device_args` (everything up to the device call) is exercised against the
unmodified reference in tests/test_dropin.py whenever /root/reference is present.

`Species` stays the `OrderedDict` the rest of the reference reads: `sync_to_host(spp)` rebuilds
its `Individual` values from the device state (ids in species order; x, y, age, sex, z, fit, e,
g), and is called by the replaced `_set_Nt` when `eager=True`, or by the caller before anything
outside the hot path looks at the individuals (getters species.py:1364-1499, stats, writers).

There is no CPU fallback: without libgnxb200.so / a CUDA device `attach` raises.
"""
import numpy as np

from .device import DeviceSpecies


def _paths_from_subsetters(recombinations, L):
    """genome.py:209-226: each cached recombination path is a bitarray over 2L positions that picks,
    for locus l, homologue 0 ('10') or homologue 1 ('01'); the odd positions are the path."""
    n = recombinations._n
    paths = np.zeros((n, L), dtype=np.uint8)
    for k in range(n):
        sub = list(recombinations._subsetters[k])
        paths[k] = np.array(sub[1::2], dtype=np.uint8)
    return paths


def species_to_device_args(spp, land, capacity=None):
    """Everything `DeviceSpecies(...)` needs, read off a reference Species / Landscape
    ("state the rest of the package reads")."""
    rasters = np.stack([np.asarray(land[l].rast, dtype=np.float64) for l in range(len(land))])
    prm = dict(b=float(spp.b), R=float(spp.R), lam=spp.n_births_distr_lambda,
               n_births_fixed=bool(spp.n_births_fixed), mating_radius=spp.mating_radius,
               d_min=float(spp.d_min), d_max=float(spp.d_max), sex=bool(spp.sex),
               sex_ratio_p=float(spp.sex_ratio), max_age=spp.max_age, K_layer=int(spp.K_layer),
               K_factor=float(spp.K_factor), move=bool(getattr(spp, '_move', True)),
               choose_nearest=bool(getattr(spp, 'choose_nearest_mate', False)),
               inverse_dist=bool(getattr(spp, 'inverse_dist_mating', False)),
               density_grid_window_width=spp._dens_grids.window_width)
    if prm['move']:
        prm['move_distr'] = (spp.movement_distance_distr, spp.movement_distance_distr_param1,
                             spp.movement_distance_distr_param2)
        prm['direction_mu'] = float(spp.direction_distr_mu)
        prm['direction_kappa'] = float(spp.direction_distr_kappa)
    prm['disp_distr'] = (spp.dispersal_distance_distr, spp.dispersal_distance_distr_param1,
                         spp.dispersal_distance_distr_param2)
    for nm, surf in (('move_surf', getattr(spp, '_move_surf', None)), ('disp_surf', getattr(spp, '_disp_surf', None))):
        if surf is not None:                                        # spatial.py:149-184 float16 tables
            prm[nm] = dict(table=np.asarray(surf.surf), layer=int(surf.lyr_num))
    ga = None
    if spp.gen_arch is not None:
        g = spp.gen_arch
        if g.use_tskit:
            raise NotImplementedError('use_tskit=True species: genotype arrays hold only non-neutral loci')
        traits = []
        for t in (g.traits or {}).values():
            traits.append(dict(loci=np.asarray(t.loci, dtype=np.int64), alpha=np.asarray(t.alpha, dtype=np.float64),
                               phi=t.phi, gamma=float(t.gamma), lyr_num=int(t.lyr_num), univ_adv=bool(t.univ_adv)))
        ga = dict(L=int(g.L), paths=_paths_from_subsetters(g.recombinations, g.L), traits=traits,
                  dom=np.asarray(g.dom, dtype=np.int8))
    if capacity is None:
        capacity = int(max(4096, 3.0 * float(np.sum(spp.K)), 2 * len(spp)))
    res = getattr(land, '_res_ratio', (1.0, 1.0)) if hasattr(land, '_res_ratio') else (1.0, 1.0)
    return dict(land_dim=tuple(land.dim), rasters=rasters, prm=prm, gen_arch=ga, capacity=capacity,
                res_ratio=res)


def population_arrays(spp):
    """Species (OrderedDict of Individuals, individual.py:100-124) -> SoA in species order."""
    inds = list(spp.values())
    out = dict(x=np.array([i.x for i in inds], dtype=np.float64), y=np.array([i.y for i in inds], dtype=np.float64),
               age=np.array([i.age for i in inds], dtype=np.int32),
               sex=np.array([0 if i.sex is None else i.sex for i in inds], dtype=np.int8),
               idx=np.array([i.idx for i in inds], dtype=np.int64))
    if inds and all(i.g is not None for i in inds):
        out['g'] = np.stack([np.asarray(i.g, dtype=np.int8) for i in inds])
    return out


def sync_to_host(spp):
    """Rebuild the Species' Individual objects from the device state (species order kept)."""
    from collections import OrderedDict
    dev = spp._gnx
    s = dev.download(genomes=bool(spp.burned and spp.gen_arch is not None), e=True)
    proto = type(next(iter(spp.values()))) if len(spp) else None
    new = OrderedDict()
    g = s.get('g')
    for k in range(len(s['x'])):
        idx = int(s['idx'][k])
        ind = spp.get(idx)
        if ind is None:
            ind = proto(idx=idx, x=float(s['x'][k]), y=float(s['y'][k]), age=int(s['age'][k]),
                        sex=int(s['sex'][k]) if spp.sex else None)
        ind.x, ind.y, ind.age = float(s['x'][k]), float(s['y'][k]), int(s['age'][k])
        ind.e = list(s['e'][k]) if s.get('e') is not None else ind.e
        if g is not None:
            ind.g = g[k]
            ind.z = list(s['z'][k])
            ind.fit = float(s['fit'][k])
        new[idx] = ind
    spp.clear()
    spp.update(new)
    spp._set_coords_and_cells()


def attach(spp, land, seed=0, capacity=None, eager=False):
    """Move `spp` onto the GPU and swap its queue entries (model.py:615-640)."""
    a = species_to_device_args(spp, land, capacity)
    dev = DeviceSpecies(a['land_dim'], a['rasters'], a['prm'], a['gen_arch'], capacity=a['capacity'], seed=seed,
                        res_ratio=a['res_ratio'])
    p = population_arrays(spp)
    dev.set_burn(not spp.burned)
    dev.upload(p['x'], p['y'], p['age'], p['sex'], p['idx'], g=p.get('g'), max_ind_idx=spp.max_ind_idx)
    spp._gnx = dev

    spp._set_age_stage = lambda: None                  # folded into gnx_step
    spp._do_movement = lambda land=None: None          # folded into gnx_step

    def _do_pop_dynamics(land=None):
        dev.step(1)                                    # age + move + mate + births + mortality

    def _set_Nt():
        for r in dev.step_records():                   # species.py:374-380, 554
            spp.Nt.append(int(r['Nt']))
            spp.n_births.append(int(r['n_births']))
            spp.n_deaths.append(int(r['n_deaths']))
            spp.max_ind_idx += int(r['n_births'])
            spp.extinct = r['Nt'] == 0                 # demography.py:329
        if eager:
            sync_to_host(spp)
    spp._do_pop_dynamics = _do_pop_dynamics
    spp._set_Nt = _set_Nt

    orig_set_raster = land._set_raster                 # landscape.py:353-354

    def _set_raster(lyr_num, rast):
        orig_set_raster(lyr_num, rast)
        dev.set_raster(lyr_num, rast)                  # also rescales K when lyr_num == K_layer
    land._set_raster = _set_raster
    return dev
