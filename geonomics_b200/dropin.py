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


def sync_to_host(spp, genomes=None):
    """Rebuild the Species' Individual objects from the device state (species order kept)."""
    from collections import OrderedDict
    dev = spp._gnx
    if genomes is None:
        genomes = bool(spp.burned and spp.gen_arch is not None)
    s = dev.download(genomes=genomes, e=True)
    proto = type(next(iter(spp.values()))) if len(spp) else None
    new = OrderedDict()
    g = s.get('g')
    for k in range(len(s['x'])):
        idx = int(s['idx'][k])
        ind = spp.get(idx)
        if ind is None:
            ind = proto(idx=idx, x=float(s['x'][k]), y=float(s['y'][k]), age=int(s['age'][k]), sex=1)
            # (Individual.__init__ re-draws a falsy sex, individual.py:110-115: set it afterwards)
            ind.sex = int(s['sex'][k])
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


def _set_device_mutation(spp, dev):
    """Hand the reference's mutation bookkeeping (genome.py:573-608, 1060-1104) to the device."""
    ga = spp.gen_arch
    if not getattr(spp, 'mutate', False) or ga is None:
        return
    if ga.traits is not None and any((t.mu or 0) > 0 for t in ga.traits.values()):
        raise NotImplementedError('trait mutation (Trait.mu > 0) is not supported on the device path; the reference '
                                  'itself raises for it when use_tskit=False (genome.py:416-437)')
    if ga._mutables is None:
        raise RuntimeError('spp.mutate is set but gen_arch._mutables is not: attach after burn-in, or let the '
                           'wrapped _set_genomes_and_tables run first')
    dev.set_mutation(ga.mu_neut or 0, ga.mu_delet or 0, list(ga._mutables),
                     np.asarray(ga.nonneut_loci, dtype=np.int64), np.asarray(ga.delet_loci, dtype=np.int64),
                     np.asarray(ga.delet_loci_s, dtype=np.float64), ga.delet_alpha_distr_shape,
                     ga.delet_alpha_distr_scale, log_capacity=max(int(ga.L), 16))


def _sync_mutations(spp, dev):
    """Device mutation log -> the reference's gen_arch bookkeeping (genome.py:753-788)."""
    ga = spp.gen_arch
    rows, st = dev.read_mutations(max_rows=max(int(ga.L), 16))
    ga._mutables = list(ga._mutables)[:st['n_mutables']]
    ga.nonneut_loci = st['nonneut_loci'].astype(np.int64)
    ga.neut_loci = np.array(sorted(set(range(ga.L)).difference(set(int(v) for v in ga.nonneut_loci))))
    ga.delet_loci = st['delet_loci'].astype(np.int64)
    ga.delet_loci_s = st['delet_s']
    return rows


def detach(spp, land=None):
    """Undo `attach`: restore the reference's own methods and free the device context."""
    st = spp.__dict__.pop('_gnx_attached', None)
    if st is None:
        return
    for name in ('_set_age_stage', '_do_movement', '_do_pop_dynamics', '_set_Nt', '_set_genomes_and_tables',
                 '_make_change'):
        spp.__dict__.pop(name, None)                   # instance overrides off: the class methods show again
    st['land'].__dict__.pop('_set_raster', None)
    if st.get('prev_set_raster') is not None:          # another attached species' wrapper was underneath
        st['land']._set_raster = st['prev_set_raster']
    try:
        st['dev'].close()
    finally:
        spp.__dict__.pop('_gnx', None)


def attach(spp, land, seed=0, capacity=None, eager=False, disp_tries_injected=6):
    """Move `spp` onto the GPU and swap its queue entries (model.py:615-656).

    May be called before the burn-in (the device then runs in burn mode; the reference's own
    `_set_genomes_and_tables` is wrapped so that the genomes it assigns after the burn-in are
    uploaded and selection / mutation switch on) or after it.  Species change events
    (`spp._changer`, change.py:612-742) keep running through the reference's own change functions;
    the device follows `spp.K` and the changed parameters.  Calling it again on the same species
    replaces the previous attachment."""
    if spp.gen_arch is not None and getattr(spp.gen_arch, 'use_tskit', False):
        raise NotImplementedError('use_tskit=True species: genotype arrays hold only non-neutral loci')
    detach(spp)
    a = species_to_device_args(spp, land, capacity)
    dev = DeviceSpecies(a['land_dim'], a['rasters'], a['prm'], a['gen_arch'], capacity=a['capacity'], seed=seed,
                        res_ratio=a['res_ratio'], disp_tries_injected=disp_tries_injected)
    p = population_arrays(spp)
    dev.set_burn(not spp.burned)
    dev.upload(p['x'], p['y'], p['age'], p['sex'], p['idx'], g=p.get('g') if spp.burned else None,
               max_ind_idx=spp.max_ind_idx)
    if spp.burned:
        _set_device_mutation(spp, dev)
    spp._gnx = dev
    cls = type(spp)

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
        if getattr(spp, 'mutate', False) and spp.burned and spp.gen_arch is not None:
            _sync_mutations(spp, dev)
        if eager:
            sync_to_host(spp)
    spp._do_pop_dynamics = _do_pop_dynamics
    spp._set_Nt = _set_Nt

    def _set_genomes_and_tables(burn_T, T):
        # species.py:956-1094 runs on the host objects (they hold the burned-in positions), then the
        # genomes, the main-phase flags and the mutation bookkeeping go to the device
        sync_to_host(spp, genomes=False)
        cls._set_genomes_and_tables(spp, burn_T, T)
        q = population_arrays(spp)
        dev.set_burn(False)
        dev.upload(q['x'], q['y'], q['age'], q['sex'], q['idx'], g=q.get('g'), max_ind_idx=spp.max_ind_idx)
        _set_device_mutation(spp, dev)
    spp._set_genomes_and_tables = _set_genomes_and_tables

    if getattr(spp, '_changer', None) is not None:
        mirrored = ('b', 'R', 'lam', 'n_births_fixed', 'd_min', 'd_max', 'max_age', 'sex_ratio_p', 'K_factor',
                    'choose_nearest', 'inverse_dist', 'direction_mu', 'direction_kappa', 'move_distr', 'disp_distr')

        def _make_change(verbose=False):
            # species.py:836-838 runs the reference's own change functions; whatever they changed
            # (spp.K, life-history attributes, the movement surface) is mirrored afterwards
            nxt = spp._changer.next_change
            due = nxt is not None and nxt[0] == spp.t
            surf_before = (getattr(spp, '_move_surf', None), getattr(spp, '_disp_surf', None))
            cls._make_change(spp, verbose=verbose)
            if not due:
                return
            dev.set_K(np.asarray(spp.K, dtype=np.float64))
            now = species_to_device_args(spp, land, capacity=a['capacity'])['prm']
            upd = {k: now[k] for k in mirrored if k in now and now[k] != dev.prm.get(k)}
            if upd:
                dev.set_life_history(**upd)
            surf_now = (getattr(spp, '_move_surf', None), getattr(spp, '_disp_surf', None))
            tabs = [np.asarray(sn.surf) if (sn is not None and sn is not sb) else None
                    for sn, sb in zip(surf_now, surf_before)]
            if tabs[0] is not None or tabs[1] is not None:
                dev.set_surface_tables(*tabs)          # change.py:597-606
        spp._make_change = _make_change

    prev_wrapper = land.__dict__.get('_set_raster')   # another attached species' wrapper, if any
    inner = land._set_raster                           # landscape.py:353-354 (or that wrapper)

    def _set_raster(lyr_num, rast):
        inner(lyr_num, rast)
        dev.set_raster(lyr_num, rast)                  # also recomputes K when lyr_num == K_layer
    land._set_raster = _set_raster
    spp._gnx_attached = dict(dev=dev, land=land, prev_set_raster=prev_wrapper)
    return dev
