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

from . import genome_pack as gp
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


def _tskit_paths_and_subsetters(recombinations, L, n_rows):
    """gen_arch.use_tskit = True: the subsetters cover the genotype ROWS (non-neutral loci) only
    (genome.py:215-224, edited by _update_subsetters :133-160); the full-length paths are the cumulated
    breakpoints (genome.py:211-212), which also give the tskit edges (genome.py:234-281)."""
    n = recombinations._n
    paths = np.zeros((n, L), dtype=np.uint8)
    subs = np.zeros((n, n_rows), dtype=np.uint8)
    for k in range(n):
        at = np.zeros(L, dtype=np.int64)
        at[np.asarray(recombinations._breakpoints[k], dtype=np.int64)] = 1
        paths[k] = np.cumsum(at) % 2
        sub = list(recombinations._subsetters[k])
        if n_rows:
            if len(sub) != 2 * n_rows:
                raise ValueError('recombination subsetter %i covers %i rows, the genotype arrays have %i'
                                 % (k, len(sub) // 2, n_rows))
            subs[k] = np.array(sub[1::2], dtype=np.uint8)
    return paths, subs


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
        traits = []
        for t in (g.traits or {}).values():
            traits.append(dict(loci=np.asarray(t.loci, dtype=np.int64), alpha=np.asarray(t.alpha, dtype=np.float64),
                               phi=t.phi, gamma=float(t.gamma), lyr_num=int(t.lyr_num), univ_adv=bool(t.univ_adv)))
        if g.use_tskit:
            # species.py:891-905: genotype arrays hold the non-neutral loci only; the device keeps one bit per
            # locus and reads row r as bit nonneut_loci[r] (include/gnx_b200.h, tskit_layout)
            nn = np.asarray(g.nonneut_loci, dtype=np.int64)
            paths, subs = _tskit_paths_and_subsetters(g.recombinations, int(g.L), len(nn))
            for t, tr in zip((g.traits or {}).values(), traits):
                tr['loci_idxs'] = np.asarray(t.loci_idxs, dtype=np.int64)
                tr['alpha_distr'] = (t.alpha_distr_mu, t.alpha_distr_sigma, t.max_alpha_mag)
                tr['mu'] = float(t.mu or 0)
            ga = dict(L=int(g.L), paths=paths, traits=traits, dom=np.asarray(g.dom, dtype=np.int8),
                      tskit=dict(nonneut_loci=nn, subsetters=subs))
        else:
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
    ga = spp.gen_arch
    if inds and all(i.g is not None for i in inds):
        out['g'] = np.stack([np.asarray(i.g, dtype=np.int8) for i in inds])
        if ga is not None and ga.use_tskit:
            out['g'] = gp.rows_to_loci(out['g'], ga.nonneut_loci, ga.L)
    elif inds and ga is not None and ga.use_tskit and spp.burned:
        out['g'] = np.zeros((len(inds), int(ga.L), 2), dtype=np.int8)      # species.py:893-899: no rows to carry
    if inds and ga is not None and ga.use_tskit and all(len(i._nodes_tab_ids) == 2 for i in inds):
        out['node0'] = np.array([i._nodes_tab_ids[0] for i in inds], dtype=np.int32)
        out['node1'] = np.array([i._nodes_tab_ids[1] for i in inds], dtype=np.int32)
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
    tsk = spp.gen_arch is not None and getattr(spp.gen_arch, 'use_tskit', False)
    if g is not None and tsk:
        g = gp.loci_to_rows(g, spp.gen_arch.nonneut_loci)            # species.py:891-905: rows = non-neutral loci
    born = spp.__dict__.get('_gnx_attached', {}).get('born', {})
    for k in range(len(s['x'])):
        idx = int(s['idx'][k])
        ind = spp.get(idx)
        if ind is None:
            ind = proto(idx=idx, x=float(s['x'][k]), y=float(s['y'][k]), age=int(s['age'][k]), sex=1)
            # (Individual.__init__ re-draws a falsy sex, individual.py:110-115: set it afterwards)
            ind.sex = int(s['sex'][k])
            if tsk and idx in born:                                  # species.py:699-729
                n0_, n1_, row_ = born[idx]
                ind._set_nodes_tab_ids(n0_, n1_)
                ind._individuals_tab_id = row_
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
    if tsk:
        for idx in [k for k in born if k not in new]:               # dead before the host ever saw them
            del born[idx]


def _set_device_mutation(spp, dev):
    """Hand the reference's mutation bookkeeping (genome.py:573-608, 1060-1104) to the device.  A use_tskit
    species always goes through it (tskit_layout: the trait tables read the rows Trait.loci_idxs name)."""
    ga = spp.gen_arch
    if ga is None:
        return
    tsk = bool(ga.use_tskit)
    mutate = bool(getattr(spp, 'mutate', False))
    if not mutate and not tsk:
        return
    trait_mus = [float(t.mu or 0) for t in (ga.traits or {}).values()]
    if any(m > 0 for m in trait_mus) and not tsk:
        raise NotImplementedError('trait mutation (Trait.mu > 0): the reference itself raises for it when '
                                  'use_tskit=False (genome.py:416-437)')
    if mutate and ga._mutables is None:
        raise RuntimeError('spp.mutate is set but gen_arch._mutables is not: attach after burn-in, or let the '
                           'wrapped _set_genomes_and_tables run first')
    kw = {}
    if tsk:
        nn = np.asarray(ga.nonneut_loci, dtype=np.int64)
        _, subs = _tskit_paths_and_subsetters(ga.recombinations, int(ga.L), len(nn))
        kw = dict(tskit_layout=True, trait_mus=trait_mus if mutate else None,
                  trait_alpha_distr=[(t.alpha_distr_mu, t.alpha_distr_sigma, t.max_alpha_mag)
                                     for t in (ga.traits or {}).values()],
                  trait_loci_idxs=[np.asarray(t.loci_idxs, dtype=np.int64) for t in (ga.traits or {}).values()],
                  delet_loci_idxs=np.asarray(ga.delet_loci_idxs, dtype=np.int64), subsetters=subs)
    dev.set_mutation((ga.mu_neut or 0) if mutate else 0, (ga.mu_delet or 0) if mutate else 0,
                     list(ga._mutables) if mutate else [],
                     np.asarray(ga.nonneut_loci, dtype=np.int64), np.asarray(ga.delet_loci, dtype=np.int64),
                     np.asarray(ga.delet_loci_s, dtype=np.float64), ga.delet_alpha_distr_shape or 0.2,
                     ga.delet_alpha_distr_scale or 0.2, log_capacity=max(int(ga.L), 16), **kw)


def _sync_mutations(spp, dev):
    """Device mutation log -> the reference's gen_arch bookkeeping (genome.py:416-437, 753-788) and, for a
    use_tskit species, its subsetters (genome.py:133-160) and mutations table (mutation.py:44-58)."""
    ga = spp.gen_arch
    rows, st = dev.read_mutations(max_rows=max(int(ga.L), 16))
    if not rows:
        return rows
    ga._mutables = list(ga._mutables)[:st['n_mutables']]
    ga.nonneut_loci = st['nonneut_loci'].astype(np.int64)
    ga.neut_loci = np.array(sorted(set(range(ga.L)).difference(set(int(v) for v in ga.nonneut_loci))))
    ga.delet_loci = st['delet_loci'].astype(np.int64)
    ga.delet_loci_s = st['delet_s']
    if ga.use_tskit:
        traits, di = dev.read_mutation_tables()
        ga.delet_loci_idxs = di.astype(np.int64)
        for t, tr in zip((ga.traits or {}).values(), traits):
            t.loci = tr['loci'].astype(np.int64)
            t.alpha = tr['alpha']
            t.loci_idxs = tr['loci_idxs'].astype(np.int64)
            t.n_loci = len(t.loci)
        for r in rows:
            if r['type'] != 'neut':
                ga.recombinations._update_subsetters(r['locus'], r['row'])
            spp._tc.mutations.add_row(site=r['locus'], node=r['node'], derived_state='1', time=-1 * spp.t)
    return rows


def _drain_tskit_rows(spp, dev):
    """Device row buffers -> the species' TableCollection, in the reference's row order (species.py:692-736:
    per offspring one individuals row, two nodes rows, then its edges)."""
    rows = dev.tskit_drain()
    nb = len(rows['idx'])
    if nb == 0:
        return
    tc = spp._tc
    born = spp._gnx_attached['born']
    T = rows['z'].shape[1] if rows['z'].ndim == 2 else 0
    child = rows['child']
    lo = np.searchsorted(child, rows['first_node_id'] + 2 * np.arange(nb), side='left')
    hi = np.searchsorted(child, rows['first_node_id'] + 2 * np.arange(nb) + 1, side='right')
    for k in range(nb):
        idx = int(rows['idx'][k])
        loc = [float(rows['x'][k]), float(rows['y'][k])]
        if T:
            loc = loc + [float(v) for v in rows['z'][k]] + [np.nan]      # z at birth, Individual.fit unset
        row = tc.individuals.add_row(location=loc, metadata=idx.to_bytes(length=4, byteorder='little'))
        assert row == rows['first_individual_row'] + k, 'individuals table out of step with the device'
        ids = [tc.nodes.add_row(flags=1, time=-1 * spp.t, population=0, individual=row) for _ in range(2)]
        assert ids[0] == rows['first_node_id'] + 2 * k, 'nodes table out of step with the device'
        born[idx] = (ids[0], ids[1], row)
        for e in range(lo[k], hi[k]):
            tc.edges.add_row(parent=int(rows['parent'][e]), left=float(rows['left'][e]),
                             right=float(rows['right'][e]), child=int(child[e]))


def detach(spp, land=None):
    """Undo `attach`: restore the reference's own methods and free the device context."""
    st = spp.__dict__.pop('_gnx_attached', None)
    if st is None:
        return
    for name in ('_set_age_stage', '_do_movement', '_do_pop_dynamics', '_set_Nt', '_set_genomes_and_tables',
                 '_make_change', '_sort_and_simplify_table_collection'):
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
    detach(spp)
    tsk = spp.gen_arch is not None and bool(getattr(spp.gen_arch, 'use_tskit', False))
    a = species_to_device_args(spp, land, capacity)
    dev = DeviceSpecies(a['land_dim'], a['rasters'], a['prm'], a['gen_arch'], capacity=a['capacity'], seed=seed,
                        res_ratio=a['res_ratio'], disp_tries_injected=disp_tries_injected)
    spp._gnx = dev
    spp._gnx_attached = dict(dev=dev, land=land, prev_set_raster=None, born={})
    cls = type(spp)

    def _upload_main_phase():
        # genomes (use_tskit: rows spread to their loci), node ids, mutation bookkeeping, tskit row buffers
        q = population_arrays(spp)
        dev.set_burn(False)
        dev.upload(q['x'], q['y'], q['age'], q['sex'], q['idx'], g=q.get('g'), max_ind_idx=spp.max_ind_idx)
        if tsk:
            # paths and traits may have been re-drawn / edited since the context was made
            now = species_to_device_args(spp, land, capacity=a['capacity'])['gen_arch']
            dev.set_recomb_paths(now['paths'])
            dev.set_traits(now['traits'], now['dom'])
        _set_device_mutation(spp, dev)
        if tsk:
            n_bp = max(int(np.max([len(v) for v in spp.gen_arch.recombinations._breakpoints.values()])), 1)
            births = int(max(1024, 2 * float(np.sum(spp.K))))
            dev.tskit_enable(edge_capacity=2 * (n_bp + 1) * births, birth_capacity=births)
            dev.tskit_set_nodes(q['node0'], q['node1'], spp._tc.nodes.num_rows, spp._tc.individuals.num_rows)
            spp._gnx_attached['born'].clear()

    if spp.burned:
        _upload_main_phase()
    else:
        p = population_arrays(spp)
        dev.set_burn(True)
        dev.upload(p['x'], p['y'], p['age'], p['sex'], p['idx'], g=None, max_ind_idx=spp.max_ind_idx)

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
        if tsk and spp.burned:
            _drain_tskit_rows(spp, dev)
        if eager:
            sync_to_host(spp)
    spp._do_pop_dynamics = _do_pop_dynamics
    spp._set_Nt = _set_Nt

    def _set_genomes_and_tables(burn_T, T):
        # species.py:956-1094 runs on the host objects (they hold the burned-in positions), then the
        # genomes, the main-phase flags and the mutation bookkeeping go to the device
        sync_to_host(spp, genomes=False)
        cls._set_genomes_and_tables(spp, burn_T, T)
        _upload_main_phase()
    spp._set_genomes_and_tables = _set_genomes_and_tables

    if tsk:
        def _sort_and_simplify_table_collection(*args, **kw):
            # species.py:1107-1219: tskit's own sort() / simplify() run on the host tables; the reference then
            # numbers the nodes 2k, 2k + 1 in species order (:1148-1152) and re-reads the individuals rows from
            # the metadata -- the device follows with gnx_tskit_renumber
            sync_to_host(spp)
            cls._sort_and_simplify_table_collection(spp, *args, **kw)
            q = population_arrays(spp)                 # nodes 2k, 2k + 1 now; the next rows continue from the
            dev.tskit_set_nodes(q['node0'], q['node1'], spp._tc.nodes.num_rows, spp._tc.individuals.num_rows)
            spp._gnx_attached['born'].clear()
        spp._sort_and_simplify_table_collection = _sort_and_simplify_table_collection

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
    spp._gnx_attached['prev_set_raster'] = prev_wrapper
    return dev
