"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference hot path.

Plain numpy/scipy restatement of the per-timestep individual-based update loop of
erthward/geonomics v1.4.9 (pure Python reference at /root/reference, cited below as
`file:line` relative to /root/reference/geonomics/).  It is the parity oracle for
the CUDA path in geonomics_b200/; it is never on the product path.

Parity pinning: the reference ships no golden vectors for this path (SURVEY.md
section 4), so the oracle is pinned against the reference itself: tests/golden/
holds stage-by-stage input/output vectors recorded from the unmodified reference
run in the build container with its numpy.random call sites replaced by replayed
draws (tests/golden/make_golden.py), and tests/test_oracle_golden.py checks every
function here against them.

Conventions shared with the CUDA path (DESIGN.md "Draw conventions"):
  * individuals are held structure-of-arrays in *species order* (the iteration
    order of the reference's Species OrderedDict: survivors in insertion order,
    newborns appended pairs-outer / offspring-inner, species.py:638-677);
  * random draws are injected: uniforms in [0,1) for Bernoulli events
    (event = u < p), u32 words for uniform integer choices
    (k = (R * n) >> 32), and sampler *outputs* for vonmises/wald/lognormal/
    levy/poisson (SURVEY.md section 7 "RNG");
  * where the reference's order depends on Python set/frozenset hashing
    (spatial.py:232-242, mating.py:63) a canonical order is defined here and the
    reference is matched set-wise.
"""
import numpy as np

# --------------------------------------------------------------------------------------
# genotype packing helpers (the CUDA path keeps genotypes bit-packed; the reference keeps
# an (L, 2) int8 array per individual, individual.py:102-106)
# --------------------------------------------------------------------------------------


def words_per_hap(L):
    """u32 words per packed haplotype: rows are padded to 128-bit units."""
    return 4 * ((int(L) + 127) // 128)


def pack_genomes(g):
    """g int8[N, L, 2] -> u32[N, 2, W]; locus l of homologue h is bit (l % 32) of word
    l // 32 of row [i, h]."""
    g = np.asarray(g)
    n, L, _ = g.shape
    W = words_per_hap(L)
    bits = np.zeros((n, 2, W * 32), dtype=np.uint8)
    bits[:, :, :L] = np.transpose(g, (0, 2, 1)).astype(np.uint8)
    b = bits.reshape(n, 2, W, 32).astype(np.uint64)
    weights = (np.uint64(1) << np.arange(32, dtype=np.uint64))
    return (b * weights).sum(axis=3).astype(np.uint32)


def unpack_genomes(packed, L):
    """inverse of pack_genomes -> int8[N, L, 2]."""
    packed = np.asarray(packed, dtype=np.uint32)
    n, _, W = packed.shape
    shifts = np.arange(32, dtype=np.uint32)
    bits = ((packed[:, :, :, None] >> shifts) & np.uint32(1)).astype(np.int8)
    bits = bits.reshape(n, 2, W * 32)[:, :, :L]
    return np.ascontiguousarray(np.transpose(bits, (0, 2, 1)))


def pack_paths(paths):
    """paths uint8[n_sims, L] (0/1 = homologue the cached recombination path is on at
    each locus, genome.py:209-215) -> u32[n_sims, W]."""
    paths = np.asarray(paths)
    n, L = paths.shape
    W = words_per_hap(L)
    bits = np.zeros((n, W * 32), dtype=np.uint64)
    bits[:, :L] = paths
    weights = (np.uint64(1) << np.arange(32, dtype=np.uint64))
    return (bits.reshape(n, W, 32) * weights).sum(axis=2).astype(np.uint32)


# --------------------------------------------------------------------------------------
# a1  age
# --------------------------------------------------------------------------------------
def age_step(age):
    """species.py:567-569 / individual.py:139: age += 1 for everyone."""
    return age + 1


# --------------------------------------------------------------------------------------
# a2  movement
# --------------------------------------------------------------------------------------
def _cos_sin(direction):
    """np.cos / np.sin as movement.py:75-76 evaluates them.  For a float64 direction this
    is plain double precision.  For a float16 direction (conductance-surface lookup,
    spatial.py:182-184, table dtype spatial.py:447) numpy returns float16, and *how* it
    gets there depends on the host CPU: with AVX512 dispatch numpy 2.x uses a vector
    half-precision routine that is off by up to 1.43 half-ulp on 13 % of inputs, without it
    it converts half -> float, calls libm cosf/sinf and rounds back to half.  The oracle
    (and the CUDA path) pin the portable semantics: the correctly rounded float32 result
    rounded to half; golden vectors are recorded with NPY_DISABLE_CPU_FEATURES set so the
    reference takes the same path (tests/golden/make_golden.py)."""
    direction = np.asarray(direction)
    if direction.dtype == np.float16:
        d64 = direction.astype(np.float64)
        return (np.cos(d64).astype(np.float32).astype(np.float16),
                np.sin(d64).astype(np.float32).astype(np.float16))
    return np.cos(direction), np.sin(direction)


def move(x, y, direction, distance, land_dim, res_ratio=(1, 1)):
    """movement.py:74-92."""
    c, s = _cos_sin(direction)
    dist_x = c * distance
    dist_y = s * distance
    if res_ratio[0] != 1:
        dist_x = dist_x * res_ratio[0]
    if res_ratio[1] != 1:
        dist_y = dist_y * res_ratio[1]
    new_x = np.clip(x + dist_x, 0, land_dim[0] - 0.001)
    new_y = np.clip(y + dist_y, 0, land_dim[1] - 0.001)
    return np.float64(new_x), np.float64(new_y)


def surface_directions(surf, x_cells, y_cells, choices):
    """spatial.py:182-184: surf[y, x, randint(0, approx_len)] (float16)."""
    return surf[y_cells, x_cells, choices]


# --------------------------------------------------------------------------------------
# a3 / a4  environment sample, cells
# --------------------------------------------------------------------------------------
def cells(x, y):
    """species.py:937-939: int32(floor(coords)); col 0 = x cell (j), col 1 = y cell (i)."""
    return np.int32(np.floor(x)), np.int32(np.floor(y))


def sample_env(rasters, x, y):
    """species.py:913-922: e[i, l] = land[l].rast[int(y_i), int(x_i)]."""
    cx = x.astype(np.int64)
    cy = y.astype(np.int64)
    return np.stack([np.asarray(r)[cy, cx] for r in rasters], axis=1)


# --------------------------------------------------------------------------------------
# a5 / a6  mate search
# --------------------------------------------------------------------------------------
MATING_GRID_MAX_CELLS = 1 << 22


def mating_grid(land_dim, radius):
    """Binning grid used to define the canonical neighbour order: square cells of side
    cs >= radius*(1+1e-7) (so every neighbour within `radius` lies in the 3x3 block of
    cells around the focal cell), doubled until the grid has <= 2^22 cells."""
    cs = float(radius) * 1.0000001
    while (int(land_dim[0] / cs) + 1) * (int(land_dim[1] / cs) + 1) > MATING_GRID_MAX_CELLS:
        cs *= 2.0
    ncx = int(land_dim[0] / cs) + 1
    ncy = int(land_dim[1] / cs) + 1
    return cs, ncx, ncy


def canonical_rank(x, y, land_dim, radius):
    """Position of every individual in the stable cell-sorted order (key = cy*ncx+cx,
    ties by species-order ordinal).  Neighbour lists are ordered by this rank."""
    cs, ncx, ncy = mating_grid(land_dim, radius)
    cx = np.floor(x / cs).astype(np.int64)
    cy = np.floor(y / cs).astype(np.int64)
    key = cy * ncx + cx
    perm = np.argsort(key, kind='stable')
    rank = np.empty(len(x), dtype=np.int64)
    rank[perm] = np.arange(len(x))
    return rank, perm, key


def neighbor_lists(x, y, land_dim, radius):
    """spatial.py:232-236: for each focal i, all j != i with
    (x_i-x_j)^2 + (y_i-y_j)^2 <= radius^2 (scipy cKDTree.query_ball_point closed ball,
    p=2, compared on squared distances), in canonical order."""
    from scipy.spatial import cKDTree
    n = len(x)
    rank, _, _ = canonical_rank(x, y, land_dim, radius)
    coords = np.stack([x, y], axis=1)
    if n == 0:
        return []
    tree = cKDTree(coords, leafsize=100)        # species.py:2170, spatial.py:189
    raw = tree.query_ball_point(coords, r=radius)
    out = []
    for i, l in enumerate(raw):
        l = np.array([j for j in l if j != i], dtype=np.int64)
        if len(l):
            l = l[np.argsort(rank[l], kind='stable')]
        out.append(l)
    return out


def neighbor_lists_bruteforce(x, y, land_dim, radius):
    """O(N^2) cross-check of neighbor_lists (small N only)."""
    rank, _, _ = canonical_rank(x, y, land_dim, radius)
    r2 = radius * radius
    out = []
    for i in range(len(x)):
        dx = x - x[i]
        dy = y - y[i]
        d2 = dx * dx + dy * dy
        l = np.nonzero(d2 <= r2)[0]
        l = l[l != i]
        l = l[np.argsort(rank[l], kind='stable')]
        out.append(l)
    return out


def choose_k(R, n):
    """uniform integer in [0, n) from a u32 word: (R * n) >> 32."""
    return (np.uint64(R) * np.uint64(n)) >> np.uint64(32)


def find_mates_radius(x, y, land_dim, radius, b, mate_R, mate_u, sex=None,
                      choose_nearest=False, inverse_dist=False, inv_u=None, nbrs=None):
    """species.py:2157-2215 + spatial.py:191-245 + mating.py:24-117 (radius modes).

    Returns (pairs int64[P, 2] of species-order ordinals, n_nbrs int32[N], mate int64[N]).

      * default mode: mate_i = nbrs_i[(mate_R[i] * len(nbrs_i)) >> 32]
        (np.random.choice(opts), spatial.py:241);
      * choose_nearest: nearest neighbour within radius (spatial.py:194-203);
      * inverse_dist: p_j proportional to (radius - dist_ij), zero-distance neighbours
        excluded (spatial.py:209-229); chosen by inverse CDF on inv_u[i] over the
        canonical neighbour order;
      * the pair survives iff mate_u[i] < b (species.py:2212-2214);
      * sex: keep (female focal, male mate) only (mating.py:41-55); otherwise a
        reciprocal couple {i,m},{m,i} is kept once (mating.py:62-63), canonically as
        the pair whose focal has the smaller ordinal.
    Pairs are listed in ascending focal ordinal, column 0 = focal.
    """
    n = len(x)
    if nbrs is None:
        nbrs = neighbor_lists(x, y, land_dim, radius)
    n_nbrs = np.array([len(l) for l in nbrs], dtype=np.int32)
    mate = np.full(n, -1, dtype=np.int64)
    for i in range(n):
        l = nbrs[i]
        if len(l) == 0:
            continue
        if choose_nearest:
            dx = x[l] - x[i]
            dy = y[l] - y[i]
            d2 = dx * dx + dy * dy
            m = l[np.argmin(d2)]
        elif inverse_dist:
            dx = x[l] - x[i]
            dy = y[l] - y[i]
            d = np.sqrt(dx * dx + dy * dy)
            ok = d != 0
            if not ok.any():
                continue
            w = np.where(ok, radius - d, 0.0)
            cdf = np.cumsum(w)
            k = int(np.searchsorted(cdf, inv_u[i] * cdf[-1], side='right'))
            k = min(k, len(l) - 1)
            m = l[k]
        else:
            m = l[int(choose_k(mate_R[i], len(l)))]
        if mate_u[i] < b:
            mate[i] = m
    if sex is not None:
        keep = (mate >= 0)
        keep &= (sex == 0)
        keep[keep] &= (sex[mate[keep]] == 1)
    else:
        keep = mate >= 0
        idx = np.nonzero(keep)[0]
        recip = mate[mate[idx]] == idx        # my mate also (validly) chose me
        drop = recip & (mate[idx] < idx)
        keep[idx[drop]] = False
    foc = np.nonzero(keep)[0]
    pairs = np.stack([foc, mate[foc]], axis=1).astype(np.int64) if len(foc) else \
        np.zeros((0, 2), dtype=np.int64)
    return pairs, n_nbrs, mate


def find_mates_panmixia_draws(n, b, pan_u, pan_R, sex=None):
    """species.py:2178-2194 with the draw convention shared with the CUDA path: individual i
    opens a mating slot iff pan_u[i] < b (the number of slots is then Binomial(N, b), as
    np.random.binomial(n=len(self), p=self.b) draws it); the slot's parents are
    a = (pan_R[i,0]*N) >> 32, b = (pan_R[i,1]*N) >> 32 (np.random.choice with replacement);
    selfing pairs are dropped (species.py:2192-2193); with sexes column 0 must be female and
    column 1 male (mating.py:41-55).  Pairs are listed in slot order."""
    pan_R = np.asarray(pan_R, dtype=np.uint64).reshape(-1, 2)[:n]
    a = ((pan_R[:, 0] * np.uint64(n)) >> np.uint64(32)).astype(np.int64)
    c = ((pan_R[:, 1] * np.uint64(n)) >> np.uint64(32)).astype(np.int64)
    active = np.ones(n, bool) if b >= 1 else (np.asarray(pan_u)[:n] < b)
    ok = active & (a != c)
    if sex is not None:
        ok &= (sex[a] == 0) & (sex[c] == 1)
    return np.stack([a[ok], c[ok]], axis=1)


def find_mates_panmixia(n_mates_draw_idx):
    """species.py:2178-2194 (mating_radius=None): 2*n_mates ordinals drawn with
    replacement, folded to (n_mates, 2), selfing pairs dropped; no de-dup
    (mating.py:64-65).  Canonical parent order = draw order."""
    p = np.asarray(n_mates_draw_idx, dtype=np.int64).reshape(-1, 2)
    return p[p[:, 0] != p[:, 1]]


# --------------------------------------------------------------------------------------
# a8 births bookkeeping
# --------------------------------------------------------------------------------------
def n_births(n_pairs, lam, fixed, poisson_draws=None):
    """species.py:604-609, mating.py:120-126."""
    if fixed:
        return np.full(n_pairs, int(lam), dtype=np.int64)
    return np.clip(np.asarray(poisson_draws[:n_pairs], dtype=np.int64), 1, None)


def offspring_ids(max_ind_idx, total_births):
    """species.py:614-619: ids max_ind_idx+1 ... handed out in offspring order."""
    return np.arange(max_ind_idx + 1, max_ind_idx + 1 + total_births, dtype=np.int64)


# --------------------------------------------------------------------------------------
# a9 gamete formation
# --------------------------------------------------------------------------------------
def offspring_table(nb):
    """(pair index, within-pair offspring index) for every offspring, pairs outer."""
    nb = np.asarray(nb, dtype=np.int64)
    pair_of = np.repeat(np.arange(len(nb)), nb)
    start = np.concatenate([[0], np.cumsum(nb)])[:-1]
    j_of = np.arange(int(nb.sum())) - np.repeat(start, nb)
    return pair_of, j_of, start


def gamete_keys(nb, recomb_keys):
    """mating.py:176-181, 204-209: pair p owns keys[2*start_p : 2*(start_p+n_p)]; its
    offspring j pops from the END of that slice: first pop -> parent pair[0], second pop
    -> parent pair[1].  Returns int64[B, 2]."""
    pair_of, j_of, start = offspring_table(nb)
    nb = np.asarray(nb, dtype=np.int64)
    s = 2 * start[pair_of]
    e = s + 2 * nb[pair_of]
    k0 = np.asarray(recomb_keys)[e - 1 - 2 * j_of]
    k1 = np.asarray(recomb_keys)[e - 2 - 2 * j_of]
    return np.stack([k0, k1], axis=1).astype(np.int64)


def make_gametes(g, pairs, nb, recomb_keys, start_homs, paths):
    """mating.py:130-172.  g int8[N, L, 2] (species order), pairs int[P, 2] ordinals,
    nb int[P], recomb_keys int[2B], start_homs int[B, 2] (binomial(1,.5,2) per
    offspring, mating.py:133), paths uint8[n_sims, L].

    child[l, c] = g_parent_c[l, paths[key_c][l] XOR start_homs[c]]; column c of the
    child comes from pairs[p, c] (mating.py:165-169)."""
    pair_of, j_of, _ = offspring_table(nb)
    keys = gamete_keys(nb, recomb_keys)
    B = len(pair_of)
    L = g.shape[1]
    child = np.zeros((B, L, 2), dtype=np.int8)
    loc = np.arange(L)
    for c in (0, 1):
        par = pairs[pair_of, c]
        hom = paths[keys[:, c]].astype(np.int64) ^ np.asarray(start_homs)[:, c:c + 1]
        child[:, :, c] = g[par[:, None], loc[None, :], hom]
    return child


# --------------------------------------------------------------------------------------
# a12 phenotype
# --------------------------------------------------------------------------------------
def phenotype(g, traits, dom=None):
    """selection.py:22-48.  traits: list of dicts {loci (row indices into g), alpha}.
    genotype = mean over homologues in {0, .5, 1}; with dominance
    clip(genotype*(1+dom[locus]), max 1); polygenic z = 0.5 + sum(genotype*alpha)
    (left-to-right Python sum); monogenic z = genotype[0]."""
    n = g.shape[0]
    z = np.zeros((n, len(traits)), dtype=np.float64)
    for t, tr in enumerate(traits):
        loci = np.asarray(tr['loci'], dtype=np.int64)
        alpha = np.asarray(tr['alpha'], dtype=np.float64)
        geno = g[:, loci, :].astype(np.float64).mean(axis=2)
        if dom is not None and np.any(dom):
            d = np.asarray(tr.get('dom_loci', loci), dtype=np.int64)
            geno = np.clip(geno * (1 + np.asarray(dom)[d]), None, 1)
        if len(loci) > 1:
            acc = np.zeros(n)
            for k in range(len(loci)):           # sequential, like Python's sum()
                acc = acc + geno[:, k] * alpha[k]
            z[:, t] = 0.5 + acc
        else:
            z[:, t] = geno[:, 0]
    return z


# --------------------------------------------------------------------------------------
# a11 dispersal + newborn sex
# --------------------------------------------------------------------------------------
def disperse(mid_x, mid_y, dir_draws, dist_draws, land_dim, res_ratio=(1, 1)):
    """movement.py:98-141.  dir_draws/dist_draws [B, R]: the successive sampler outputs
    each offspring consumes until its position is accepted
    (0 < x < dim_x and 0 < y < dim_y after clipping to [0, dim-0.001]).
    Returns x, y, n_tries (n_tries == R+1 marks 'draws exhausted')."""
    B, R = dir_draws.shape
    ox = np.zeros(B)
    oy = np.zeros(B)
    tries = np.zeros(B, dtype=np.int32)
    done = np.zeros(B, dtype=bool)
    for r in range(R):
        c, s = _cos_sin(dir_draws[:, r])
        dx = c * dist_draws[:, r]
        dy = s * dist_draws[:, r]
        if res_ratio[0] != 1:
            dx = dx * res_ratio[0]
        if res_ratio[1] != 1:
            dy = dy * res_ratio[1]
        nx = np.clip(mid_x + dx, 0, land_dim[0] - 0.001)
        ny = np.clip(mid_y + dy, 0, land_dim[1] - 0.001)
        ok = (nx > 0) & (nx < land_dim[0]) & (ny > 0) & (ny < land_dim[1])
        upd = ~done
        ox[upd] = nx[upd]
        oy[upd] = ny[upd]
        tries[upd] += 1
        done |= ok
    tries[~done] = R + 1
    return ox, oy, tries


def newborn_sex(sexed, sex_ratio_p, u_sex, u_redraw):
    """species.py:657-662 + individual.py:110-115: if the species is sexed,
    sex = (u_sex < sex_ratio_p); whenever that is falsy (0 or None) Individual.__init__
    re-draws sex = binomial(1, 0.5) = (u_redraw < 0.5)."""
    first = (u_sex < sex_ratio_p) if sexed else np.zeros(len(u_redraw), dtype=bool)
    return np.where(first, 1, (u_redraw < 0.5).astype(np.int8)).astype(np.int8)


# --------------------------------------------------------------------------------------
# a10 density
# --------------------------------------------------------------------------------------
def _floordiv(a, b):
    return np.floor_divide(a, b)


class DensityGridStack:
    """spatial.py:100-146 (stack), :34-97 (grid), :270-360 (construction)."""

    def __init__(self, land_dim, window_width=None):
        self.dim = tuple(land_dim)
        if window_width is None:
            window_width = round(0.1 * max(self.dim))        # spatial.py:110-111
        self.ww = window_width
        ww = window_width
        hww = ww / 2.
        self.grids = []
        for x_edge, y_edge in ((True, True), (False, False), (True, False), (False, True)):
            xs = np.arange(0, self.dim[0] + ww, ww) if x_edge else \
                np.arange(0 + hww, self.dim[0] + hww, ww)
            ys = np.arange(0, self.dim[1] + ww, ww) if y_edge else \
                np.arange(0 + hww, self.dim[1] + hww, ww)
            gj, gi = np.meshgrid(xs, ys)
            # window-landscape intersection areas (spatial.py:296-319)
            wx = np.clip(np.minimum(gj + hww, self.dim[0]) - np.maximum(gj - hww, 0), 0, None)
            wy = np.clip(np.minimum(gi + hww, self.dim[1]) - np.maximum(gi - hww, 0), 0, None)
            areas = wx * wy
            areas[areas == 0] = 0.0001
            i_cells = _floordiv(gi - hww * y_edge, ww) + y_edge
            j_cells = _floordiv(gj - hww * x_edge, ww) + x_edge
            self.grids.append(dict(x_edge=x_edge, y_edge=y_edge, gi=gi, gj=gj, areas=areas,
                                   i_cells=i_cells.astype(np.int64),
                                   j_cells=j_cells.astype(np.int64)))
        self.pts = np.vstack([np.stack([g['gi'].ravel(), g['gj'].ravel()], axis=1)
                              for g in self.grids])           # (i, j) order, spatial.py:64-65
        self.land_gj, self.land_gi = np.meshgrid(np.arange(0, self.dim[0]) + 0.5,
                                                 np.arange(0, self.dim[1]) + 0.5)

    def counts(self, x, y):
        """integer counts per grid point (spatial.py:79-92), list of 4 int64 arrays."""
        out = []
        for g in self.grids:
            xc = (_floordiv(x - g['x_edge'] * self.ww / 2., self.ww) + g['x_edge']).astype(np.int64)
            yc = (_floordiv(y - g['y_edge'] * self.ww / 2., self.ww) + g['y_edge']).astype(np.int64)
            ni, nj = g['gi'].shape
            # grid point (a, b) has cell id (i_cells[a, b], j_cells[a, b]); those are
            # a + const and b + const, so build a dense histogram and index it.
            i0 = g['i_cells'][0, 0]
            j0 = g['j_cells'][0, 0]
            h = np.zeros((ni, nj), dtype=np.int64)
            a = yc - i0
            b_ = xc - j0
            ok = (a >= 0) & (a < ni) & (b_ >= 0) & (b_ < nj)
            np.add.at(h, (a[ok], b_[ok]), 1)
            out.append(h)
        return out

    def vals(self, x, y):
        return np.hstack([(c / g['areas']).ravel() for c, g in zip(self.counts(x, y), self.grids)])

    def calc_density(self, x, y):
        """spatial.py:132-146: scipy griddata cubic from the lattice to cell centres."""
        from scipy import interpolate
        return interpolate.griddata(self.pts, self.vals(x, y), (self.land_gi, self.land_gj),
                                    method='cubic')


def species_density(dgs, x, y):
    """species.py:845-882 with set_N=True: clip >= 0."""
    return np.clip(dgs.calc_density(x, y), 0, None)


def n_pairs_raster(dgs, x, y, pairs, land_dim):
    """demography.py:60-91: density of pair midpoints, clip >= 0, NaN -> 0."""
    if len(pairs) == 0:
        return np.zeros(land_dim)            # NB reference shape quirk: (dim_x, dim_y)
    px = (x[pairs[:, 0]] + x[pairs[:, 1]]) / 2
    py = (y[pairs[:, 0]] + y[pairs[:, 1]]) / 2
    n_pairs = np.clip(dgs.calc_density(px, py), 0, None)
    n_pairs[np.isnan(n_pairs)] = 0
    return n_pairs


# ---- restated Clough-Tocher pieces (scipy/interpolate/interpnd.pyx, scipy 1.18.1; the
# ---- reference reaches them through scipy.interpolate.griddata(method='cubic'),
# ---- spatial.py:144).  Validated against scipy in tests/test_oracle_density.py; this is
# ---- the algorithm the CUDA density kernels implement.
def ct_triangulation(pts):
    from scipy.spatial import Delaunay
    tri = Delaunay(pts)                      # Qhull options 'Qbb Qc Qz Q12' as in griddata
    indptr, indices = tri.vertex_neighbor_vertices
    return dict(points=np.ascontiguousarray(tri.points),
                simplices=np.ascontiguousarray(tri.simplices, dtype=np.int32),
                neighbors=np.ascontiguousarray(tri.neighbors, dtype=np.int32),
                nbr_indptr=np.ascontiguousarray(indptr, dtype=np.int32),
                nbr_indices=np.ascontiguousarray(indices, dtype=np.int32),
                tri=tri)


def ct_gradients(T, f, maxiter=400, tol=1e-6):
    """interpnd.pyx `_estimate_gradients_2d_global`: Gauss-Seidel minimisation of the
    edge-wise second-derivative energy; one 2x2 solve per vertex per sweep, in vertex
    order, stop when the max relative change of a sweep is < tol."""
    P = T['points']
    ip = T['nbr_indptr']
    ix = T['nbr_indices']
    n = P.shape[0]
    yv = np.zeros((n, 2))
    for it in range(maxiter):
        err = 0.0
        for i in range(n):
            Q0 = Q1 = Q3 = 0.0
            s0 = s1 = 0.0
            for jj in range(ip[i], ip[i + 1]):
                j = ix[jj]
                ex = P[j, 0] - P[i, 0]
                ey = P[j, 1] - P[i, 1]
                Ln = np.sqrt(ex * ex + ey * ey)
                L3 = Ln * Ln * Ln
                f1 = f[i]
                f2 = f[j]
                df2 = -ex * yv[j, 0] - ey * yv[j, 1]
                Q0 += 4 * ex * ex / L3
                Q1 += 4 * ex * ey / L3
                Q3 += 4 * ey * ey / L3
                s0 += (6 * (f1 - f2) - 2 * df2) * ex / L3
                s1 += (6 * (f1 - f2) - 2 * df2) * ey / L3
            Q2 = Q1
            det = Q0 * Q3 - Q1 * Q2
            r0 = (Q3 * s0 - Q1 * s1) / det
            r1 = (-Q2 * s0 + Q0 * s1) / det
            change = max(abs(yv[i, 0] + r0), abs(yv[i, 1] + r1))
            yv[i, 0] = -r0
            yv[i, 1] = -r1
            change /= max(1.0, max(abs(r0), abs(r1)))
            err = max(err, change)
        if err < tol:
            return yv, it + 1
    return yv, 0


def ct_coefficients(T, f, grad):
    """interpnd.pyx `_clough_tocher_2d_single`, the part that does not depend on the
    query point: the 19 Bezier ordinates of every triangle.  Returns float64[ntri, 19] in
    the order c3000 c0300 c0030 c0003 c2100 c2010 c2001 c0210 c0201 c0021 c1200 c1020
    c1002 c0120 c0102 c0012 c1101 c1011 c0111."""
    P = T['points']
    S = T['simplices']
    NB = T['neighbors']
    nt = S.shape[0]
    out = np.zeros((nt, 19))
    for t in range(nt):
        v0, v1, v2 = S[t]
        e12x = P[v1, 0] - P[v0, 0]
        e12y = P[v1, 1] - P[v0, 1]
        e23x = P[v2, 0] - P[v1, 0]
        e23y = P[v2, 1] - P[v1, 1]
        e31x = P[v0, 0] - P[v2, 0]
        e31y = P[v0, 1] - P[v2, 1]
        f1, f2, f3 = f[v0], f[v1], f[v2]
        df12 = +(grad[v0, 0] * e12x + grad[v0, 1] * e12y)
        df21 = -(grad[v1, 0] * e12x + grad[v1, 1] * e12y)
        df23 = +(grad[v1, 0] * e23x + grad[v1, 1] * e23y)
        df32 = -(grad[v2, 0] * e23x + grad[v2, 1] * e23y)
        df31 = +(grad[v2, 0] * e31x + grad[v2, 1] * e31y)
        df13 = -(grad[v0, 0] * e31x + grad[v0, 1] * e31y)
        c3000 = f1
        c2100 = (df12 + 3 * c3000) / 3
        c2010 = (df13 + 3 * c3000) / 3
        c0300 = f2
        c1200 = (df21 + 3 * c0300) / 3
        c0210 = (df23 + 3 * c0300) / 3
        c0030 = f3
        c1020 = (df31 + 3 * c0030) / 3
        c0120 = (df32 + 3 * c0030) / 3
        c2001 = (c2100 + c2010 + c3000) / 3
        c0201 = (c1200 + c0300 + c0210) / 3
        c0021 = (c1020 + c0120 + c0030) / 3
        g = np.zeros(3)
        # barycentric transform of this triangle: c = Tinv (p - v2)
        A = np.array([[P[v0, 0] - P[v2, 0], P[v1, 0] - P[v2, 0]],
                      [P[v0, 1] - P[v2, 1], P[v1, 1] - P[v2, 1]]])
        Ainv = np.linalg.inv(A)
        for k in range(3):
            itri = NB[t, k]
            if itri == -1:
                g[k] = -0.5
                continue
            cen = P[S[itri]].sum(axis=0) / 3
            c01 = Ainv @ (cen - P[v2])
            c = np.array([c01[0], c01[1], 1 - c01[0] - c01[1]])
            if k == 0:
                g[k] = (2 * c[2] + c[1] - 1) / (2 - 3 * c[2] - 3 * c[1])
            elif k == 1:
                g[k] = (2 * c[0] + c[2] - 1) / (2 - 3 * c[0] - 3 * c[2])
            else:
                g[k] = (2 * c[1] + c[0] - 1) / (2 - 3 * c[1] - 3 * c[0])
        c0111 = (g[0] * (-c0300 + 3 * c0210 - 3 * c0120 + c0030)
                 + (-c0300 + 2 * c0210 - c0120 + c0021 + c0201)) / 2
        c1011 = (g[1] * (-c0030 + 3 * c1020 - 3 * c2010 + c3000)
                 + (-c0030 + 2 * c1020 - c2010 + c2001 + c0021)) / 2
        c1101 = (g[2] * (-c3000 + 3 * c2100 - 3 * c1200 + c0300)
                 + (-c3000 + 2 * c2100 - c1200 + c2001 + c0201)) / 2
        c1002 = (c1101 + c1011 + c2001) / 3
        c0102 = (c1101 + c0111 + c0201) / 3
        c0012 = (c1011 + c0111 + c0021) / 3
        c0003 = (c1002 + c0102 + c0012) / 3
        out[t] = [c3000, c0300, c0030, c0003, c2100, c2010, c2001, c0210, c0201, c0021,
                  c1200, c1020, c1002, c0120, c0102, c0012, c1101, c1011, c0111]
    return out


def ct_eval(T, coef, simplex, pi, pj):
    """interpnd.pyx `_clough_tocher_2d_single`, the point-dependent part, for points
    (pi, pj) lying in triangles `simplex`."""
    P = T['points']
    S = T['simplices'][simplex]
    v0, v1, v2 = S[:, 0], S[:, 1], S[:, 2]
    a00 = P[v0, 0] - P[v2, 0]
    a01 = P[v1, 0] - P[v2, 0]
    a10 = P[v0, 1] - P[v2, 1]
    a11 = P[v1, 1] - P[v2, 1]
    det = a00 * a11 - a01 * a10
    dx = pi - P[v2, 0]
    dy = pj - P[v2, 1]
    b0 = (a11 * dx - a01 * dy) / det
    b1 = (-a10 * dx + a00 * dy) / det
    b2 = 1 - b0 - b1
    minval = np.minimum(np.minimum(b0, b1), b2)
    b1_, b2_, b3_, b4_ = b0 - minval, b1 - minval, b2 - minval, 3 * minval
    c = coef[simplex]
    (c3000, c0300, c0030, c0003, c2100, c2010, c2001, c0210, c0201, c0021, c1200, c1020,
     c1002, c0120, c0102, c0012, c1101, c1011, c0111) = [c[:, k] for k in range(19)]
    b1, b2, b3, b4 = b1_, b2_, b3_, b4_
    w = (b1**3 * c3000 + 3 * b1**2 * b2 * c2100 + 3 * b1**2 * b3 * c2010 +
         3 * b1**2 * b4 * c2001 + 3 * b1 * b2**2 * c1200 +
         6 * b1 * b2 * b4 * c1101 + 3 * b1 * b3**2 * c1020 + 6 * b1 * b3 * b4 * c1011 +
         3 * b1 * b4**2 * c1002 + b2**3 * c0300 + 3 * b2**2 * b3 * c0210 +
         3 * b2**2 * b4 * c0201 + 3 * b2 * b3**2 * c0120 + 6 * b2 * b3 * b4 * c0111 +
         3 * b2 * b4**2 * c0102 + b3**3 * c0030 + 3 * b3**2 * b4 * c0021 +
         3 * b3 * b4**2 * c0012 + b4**3 * c0003)
    return w


def ct_density_restated(dgs, vals, T=None):
    """griddata(method='cubic') re-expressed with the pieces above."""
    if T is None:
        T = ct_triangulation(dgs.pts)
    grad, _ = ct_gradients(T, vals)
    coef = ct_coefficients(T, vals, grad)
    q = np.stack([dgs.land_gi.ravel(), dgs.land_gj.ravel()], axis=1)
    simplex = T['tri'].find_simplex(q)
    w = ct_eval(T, coef, simplex, q[:, 0], q[:, 1])
    w[simplex < 0] = np.nan
    return w.reshape(dgs.land_gi.shape)


# --------------------------------------------------------------------------------------
# a14 logistic density dependence
# --------------------------------------------------------------------------------------
def calc_dNdt(R, N, K):
    """demography.py:95-119."""
    with np.errstate(divide='ignore', invalid='ignore'):
        dNdt = R * (1 - (N / K)) * N
    dNdt = np.clip(dNdt, -1 * N.max(), None)
    dNdt[np.isnan(dNdt)] = -1 * N.max()
    dNdt[np.isinf(dNdt)] = -1 * N.max()
    return dNdt


def calc_d(N, n_pairs, K, R, b, lam, d_min, d_max):
    """demography.py:122-172: N_b = b*lam*n_pairs; N_d = N_b - dNdt; d = clip(N_d/N
    (NaN -> 0), d_min, d_max)."""
    dNdt = calc_dNdt(R, N, K)
    N_b = b * lam * n_pairs
    N_d = N_b - dNdt
    with np.errstate(divide='ignore', invalid='ignore'):
        d = N_d / N
    d[np.isnan(d)] = 0
    return np.clip(d, d_min, d_max)


# --------------------------------------------------------------------------------------
# a15 fitness / death probability
# --------------------------------------------------------------------------------------
def fitness(e, z, traits, cx=None, cy=None, delet=None):
    """selection.py:51-112.  traits: list of dicts {phi (scalar or raster), gamma,
    lyr_num, univ_adv}; w = clip(prod_t (1 - phi*|e^(not univ_adv) - z|^gamma), 0.001).
    delet: optional (dosage int[N, D], s float[D]) -> w *= prod(1 - dosage*s)."""
    n = z.shape[0] if z is not None and len(traits) else (len(cx) if cx is not None else 0)
    w = np.ones(n)
    if traits:
        fits = []
        for t, tr in enumerate(traits):
            phi = tr['phi']
            if np.ndim(phi) == 2:
                phi = np.asarray(phi)[cy, cx]
            ev = e[:, tr['lyr_num']] ** (not tr['univ_adv'])
            fits.append(1 - phi * (np.abs(ev - z[:, t]) ** tr['gamma']))
        w = w * np.clip(np.stack(fits).prod(axis=0), 0.001, None)
    if delet is not None:
        dosage, s = delet
        w = w * (1 - dosage * s).prod(axis=1)
    return w


def prob_death(d_ind, w):
    """selection.py:119-125."""
    return 1 - (1 - d_ind) * w


def death_probs(d_raster, x, y, w=None, age=None, max_age=None):
    """demography.py:306-321."""
    cx, cy = cells(x, y)
    p = d_raster[cy, cx].copy()
    if w is not None:
        p = prob_death(p, w)
    if max_age is not None:
        p[age > max_age] = 1
    return p


def mortality(p, u):
    """demography.py:175-180: dead_i = binomial(1, p_i) := (u_i < p_i)."""
    return u < p


# --------------------------------------------------------------------------------------
# a13 mutation (use_tskit = False)
# --------------------------------------------------------------------------------------
def mutation_type_cdf(mu_neut, mu_delet, trait_mus=()):
    """genome.py:650-663 _draw_mut_types: probs = mu / sum(mu) over ('neut', 'delet', 't0', ...);
    numpy RandomState.choice(p=...) draws searchsorted(cumsum(p) / cumsum(p)[-1], u, 'right')."""
    mus = [mu_neut, mu_delet] + list(trait_mus)
    tot = sum(mus)
    p = np.array([m / tot for m in mus], dtype=np.float64)
    cdf = p.cumsum()
    cdf /= cdf[-1]
    return cdf


def mutate(og, oz, mut, draws, traits, dom, first_id):
    """ops/mutation.py:169-206 _do_mutation on this step's B offspring (rows of og int8[B, L, 2]).

    mut: dict(mu_neut, mu_delet, mutables (list, popped from the end: mutation.py:73, :95),
              nonneut_loci, delet_loci, delet_s (ascending arrays)) -- updated copies are returned.
    draws: mut_n (binomial(B*L, mu_tot) output, :172-173), and per mutation mut_type_u,
           mut_ind_R (r.choice(offspring) := offspring[(R*B) >> 32], offspring ids descending), mut_homol_u
           (r.binomial(1, .5) := u < .5), mut_s (r.gamma output, genome.py:691).
    neutral (:62-86): pops a locus, genotypes untouched (:81-82).
    deleterious (:156-166 -> :90-131): s = min(gamma, 1); idx = _add_nonneut_locus
      (genome.py:753-788: bisect into nonneut_loci; delet_loci / delet_loci_s likewise);
      g[idx, homol] = 1 -- row idx, as written at mutation.py:117; z recomputed (:119).
    Returns (og, oz, new_mut, log rows)."""
    B, L = og.shape[0], og.shape[1]
    new = dict(mut)
    mutables = list(mut['mutables'])
    nonneut = [int(v) for v in mut['nonneut_loci']]
    dl = [int(v) for v in mut['delet_loci']]
    ds = [float(v) for v in mut['delet_s']]
    log = []
    n_muts = int(np.asarray(draws['mut_n']).reshape(-1)[0]) if B else 0
    if n_muts > 0:
        cdf = mutation_type_cdf(mut['mu_neut'], mut['mu_delet'], mut.get('trait_mus', ()))
        og = og.copy()
        oz = oz.copy()
        for m in range(n_muts):
            ti = int(np.searchsorted(cdf, draws['mut_type_u'][m], side='right'))
            assert ti in (0, 1), 'trait mutation raises in the reference when use_tskit=False (genome.py:430)'
            locus = int(mutables.pop())
            # keys_list is the DESCENDING id list (species.py:615-622), so choice index k is offspring B-1-k
            o = B - 1 - int(choose_k(np.uint32(draws['mut_ind_R'][m]), B))
            homol = int(draws['mut_homol_u'][m] < 0.5)
            row, sel = -1, 0.0
            if ti == 1:
                sel = min(float(draws['mut_s'][m]), 1.0)                     # genome.py:690-693
                import bisect
                row = bisect.bisect_left(nonneut, locus)
                nonneut.insert(row, locus)
                di = bisect.bisect_left(dl, locus)
                dl.insert(di, locus)
                ds.insert(di, sel)
                og[o, row, homol] = 1                                        # mutation.py:117
                if traits:
                    oz[o] = phenotype(og[o:o + 1], traits, dom)[0]           # species.py:929
            log.append(dict(individual=first_id + o, locus=locus, row=row, homologue=homol,
                            type=('neut', 'delet')[ti], s=sel))
    new['mutables'] = mutables
    new['nonneut_loci'] = np.array(nonneut, dtype=np.int64)
    new['delet_loci'] = np.array(dl, dtype=np.int64)
    new['delet_s'] = np.array(ds, dtype=np.float64)
    return og, oz, new, log


def draw_trait_alpha(tr, draw):
    """genome.py:666-687 _draw_trait_alpha(n=1): mu when sigma == 0, else clip(normal(mu, sigma) := the injected
    sampler output, +-max_alpha_mag); |alpha| while the trait is monogenic (n_loci read BEFORE the locus is added)."""
    mu_a, sigma, max_mag = (float(v) for v in tr['alpha_distr'])
    if sigma == 0:
        alpha = mu_a
    else:
        alpha = float(draw)
        if max_mag >= 0:
            alpha = float(np.clip(alpha, -max_mag, max_mag))
    if len(tr['loci']) == 1:
        alpha = abs(alpha)
    return alpha


def rows_traits(traits):
    """use_tskit=True: _calc_phenotype reads genotype ROWS Trait.loci_idxs (selection.py:29-30); dominance still
    indexes dom by Trait.loci (selection.py:37)."""
    return [dict(tr, loci=np.asarray(tr['loci_idxs'], dtype=np.int64), dom_loci=np.asarray(tr['loci'], dtype=np.int64))
            for tr in traits]


def mutate_tskit(g_all, z_all, n0, mut, draws, dom, first_id):
    """ops/mutation.py:169-206 with gen_arch.use_tskit = True (genotype arrays hold one ROW per non-neutral locus,
    species.py:891-905).  g_all int8[N0 + B, n_nonneut, 2]: everyone alive, the B offspring last (a non-neutral
    mutation inserts a zero row into EVERY individual, species.py:908-910 / individual.py:165-168).

    mut: the dict of `mutate` plus traits (list of dicts loci, alpha, loci_idxs, alpha_distr (mu, sigma,
         max_alpha_mag or -1) and the fitness keys), trait_mus, delet_loci_idxs, subsetters uint8[n_paths, n_nonneut]
         (Recombinations._subsetters, one homologue per genotype row) and paths uint8[n_paths, L] (as simulated).
    Per non-neutral mutation, in the reference's order: [delet: s = min(gamma, 1)] -> locus = mutables.pop() ->
    individual = choice(offspring) -> idx = _add_nonneut_locus (genome.py:753-788: nonneut_loci; delet_loci /
    delet_loci_s / delet_loci_idxs -- NO later index is shifted; or Trait._add_locus genome.py:416-437 with
    alpha = _draw_trait_alpha, loci_idxs shifted behind the insertion point of THIS trait only) -> homologue ->
    zero row at idx for everyone -> g[idx, homol] = 1 -> phenotype of the individual (species.py:929) ->
    mutations-table row (mutation.py:44-58) -> subsetters get (#breakpoints < locus) % 2 at row idx
    (genome.py:133-160).  Neutral: locus, individual, homologue, mutations-table row.
    Returns (g_all, z_all, new_mut, log rows)."""
    import bisect
    B = g_all.shape[0] - n0
    new = dict(mut)
    mutables = list(mut['mutables'])
    nonneut = [int(v) for v in mut['nonneut_loci']]
    dl = [int(v) for v in mut['delet_loci']]
    ds = [float(v) for v in mut['delet_s']]
    di = [int(v) for v in mut['delet_loci_idxs']]
    traits = [dict(tr, loci=[int(v) for v in tr['loci']], alpha=[float(v) for v in tr['alpha']],
                   loci_idxs=[int(v) for v in tr['loci_idxs']]) for tr in mut['traits']]
    subs = np.asarray(mut['subsetters'], dtype=np.uint8)
    paths = np.asarray(mut['paths'], dtype=np.uint8)
    log = []
    n_muts = int(np.asarray(draws['mut_n']).reshape(-1)[0]) if B else 0
    if n_muts > 0:
        cdf = mutation_type_cdf(mut['mu_neut'], mut['mu_delet'], mut.get('trait_mus', ()))
        g_all = g_all.copy()
        z_all = z_all.copy()
        for m in range(n_muts):
            ti = int(np.searchsorted(cdf, draws['mut_type_u'][m], side='right'))
            sel = min(float(draws['mut_s'][m]), 1.0) if ti == 1 else 0.0             # genome.py:690-693
            locus = int(mutables.pop())
            o = B - 1 - int(choose_k(np.uint32(draws['mut_ind_R'][m]), B))           # descending id list
            row, alpha = -1, 0.0
            if ti >= 1:
                row = bisect.bisect_left(nonneut, locus)
                nonneut.insert(row, locus)
                if ti == 1:
                    k = bisect.bisect_left(dl, locus)
                    dl.insert(k, locus)
                    ds.insert(k, sel)
                    di.insert(k, row)
                else:
                    tr = traits[ti - 2]
                    alpha = draw_trait_alpha(tr, draws['mut_alpha'][m])
                    k = bisect.bisect_left(tr['loci'], locus)
                    tr['loci'].insert(k, locus)
                    tr['alpha'].insert(k, alpha)
                    tr['loci_idxs'] = tr['loci_idxs'][:k] + [row] + [v + 1 for v in tr['loci_idxs'][k:]]
            homol = int(draws['mut_homol_u'][m] < 0.5)
            if ti >= 1:
                g_all = np.insert(g_all, row, 0, axis=1)                             # species.py:908-910
                g_all[n0 + o, row, homol] = 1                                        # mutation.py:117
                z_all[n0 + o] = phenotype(g_all[n0 + o:n0 + o + 1], rows_traits(traits), dom)[0]
                ins = paths[:, locus - 1] if locus > 0 else np.zeros(len(paths), np.uint8)   # bisect_left(bps, locus) % 2
                subs = np.insert(subs, row, ins, axis=1)
            log.append(dict(individual=first_id + o, locus=locus, row=row, homologue=homol,
                            type=('neut', 'delet')[ti] if ti < 2 else 't%i' % (ti - 2), s=sel, alpha=alpha))
    new['mutables'] = mutables
    new['nonneut_loci'] = np.array(nonneut, dtype=np.int64)
    new['delet_loci'] = np.array(dl, dtype=np.int64)
    new['delet_s'] = np.array(ds, dtype=np.float64)
    new['delet_loci_idxs'] = np.array(di, dtype=np.int64)
    new['traits'] = [dict(tr, loci=np.array(tr['loci'], dtype=np.int64), alpha=np.array(tr['alpha'], dtype=np.float64),
                          loci_idxs=np.array(tr['loci_idxs'], dtype=np.int64)) for tr in traits]
    new['subsetters'] = subs
    return g_all, z_all, new, log


def delet_dosage(g, delet_loci):
    """selection.py:78-94: diploid dosage at the deleterious loci (use_tskit=False: g rows are loci)."""
    return np.sum(g[:, np.asarray(delet_loci, dtype=np.int64), :], axis=2)


# --------------------------------------------------------------------------------------
# one whole main time step (model.py:603-667 queue order for one species):
#   _set_age_stage -> _do_movement (+_set_e) -> _do_pop_dynamics -> _set_Nt
# --------------------------------------------------------------------------------------
def step(state, arch, prm, draws, dgs=None, max_tries=None, burn=False):
    """state: dict(x, y, age, sex, idx, g[N,L,2], z[N,T], max_ind_idx)
    arch:  dict(land_dim, rasters[n_lyr,Y,X], K[Y,X], ww, traits[list], dom, paths,
                move_surf/disp_surf (optional float16 [Y,X,A]))
    prm:   dict(b, R, lam, n_births_fixed, mating_radius, d_min, d_max, sex, sex_ratio_p,
                max_age (None or int), choose_nearest, inverse_dist)
    draws: dict(move_dir|move_choice, move_dist, mate_R, mate_u, poisson, recomb_keys,
                start_homs, disp_dir|disp_choice [cap,R], disp_dist [cap,R], sex_u,
                sex_redraw_u, death_u)
    Returns (new_state, intermediates)."""
    land_dim = tuple(int(v) for v in arch['land_dim'])
    rasters = arch['rasters']
    traits = arch['traits']
    if dgs is None:
        dgs = DensityGridStack(land_dim, arch.get('ww'))
    im = {}
    n0 = len(state['x'])
    # a1
    age = age_step(state['age'])
    # a2
    cx0, cy0 = cells(state['x'], state['y'])
    if arch.get('move_surf') is not None:
        direction = surface_directions(arch['move_surf'], cx0, cy0, draws['move_choice'][:n0])
    else:
        direction = draws['move_dir'][:n0]
    res_ratio = tuple(arch.get('res_ratio', (1, 1)))        # landscape.py:277-278
    x, y = move(state['x'], state['y'], direction, draws['move_dist'][:n0], land_dim, res_ratio)
    im['mv_x'], im['mv_y'], im['mv_age'] = x, y, age
    im['mv_e'] = sample_env(rasters, x, y)
    # a5/a6
    sexed = bool(prm['sex'])
    if prm['mating_radius'] is None:
        pairs = find_mates_panmixia_draws(n0, prm['b'], draws['pan_u'], draws['pan_R'],
                                          state['sex'] if sexed else None)
        n_nbrs = np.full(n0, n0 - 1, dtype=np.int32)
        mate = None
    else:
        pairs, n_nbrs, mate = find_mates_radius(
            x, y, land_dim, prm['mating_radius'], prm['b'], draws['mate_R'], draws['mate_u'],
            sex=state['sex'] if sexed else None,
            choose_nearest=prm.get('choose_nearest', False),
            inverse_dist=prm.get('inverse_dist', False), inv_u=draws.get('mate_inv_u'))
    im['pairs'], im['n_nbrs'], im['mate'] = pairs, n_nbrs, mate
    # a7
    im['n_pairs_rast'] = n_pairs_raster(dgs, x, y, pairs, land_dim)
    # a8/a9/a11/a12
    nb = n_births(len(pairs), prm['lam'], bool(prm['n_births_fixed']), draws.get('poisson'))
    B = int(nb.sum())
    im['nb'], im['B'] = nb, B
    pair_of, j_of, _ = offspring_table(nb)
    ids = offspring_ids(int(state['max_ind_idx']), B)
    if B:
        mid_x = (x[pairs[pair_of, 0]] + x[pairs[pair_of, 1]]) / 2
        mid_y = (y[pairs[pair_of, 0]] + y[pairs[pair_of, 1]]) / 2
        if arch.get('disp_surf') is not None:
            R = draws['disp_choice'].shape[1]
            ddir = arch['disp_surf'][np.int64(mid_y)[:, None], np.int64(mid_x)[:, None],
                                     draws['disp_choice'][:B]]
        else:
            ddir = draws['disp_dir'][:B]
        ox, oy, tries = disperse(mid_x, mid_y, ddir, draws['disp_dist'][:B], land_dim, res_ratio)
        osex = newborn_sex(sexed, prm['sex_ratio_p'], draws['sex_u'][:B],
                           draws['sex_redraw_u'][:B])
    else:
        mid_x = mid_y = ox = oy = np.zeros(0)
        tries = np.zeros(0, np.int32)
        osex = np.zeros(0, np.int8)
    im['mid_x'], im['mid_y'], im['disp_tries'] = mid_x, mid_y, tries
    tskit_layout = (not burn) and arch.get('mutation') is not None and arch['mutation'].get('tskit_layout')
    if tskit_layout:
        # gen_arch.use_tskit = True: state['g'] holds the non-neutral ROWS; gametes through the subsetters
        # (mating.py:156-166), phenotypes through Trait.loci_idxs; the evolving tables live in arch['mutation']
        mt = arch['mutation']
        traits = mt['traits']
        og = make_gametes(state['g'], pairs, nb, draws['recomb_keys'][:2 * B], draws['start_homs'][:B],
                          mt['subsetters']) if B else np.zeros((0,) + state['g'].shape[1:], np.int8)
        oz = phenotype(og, rows_traits(traits), arch.get('dom')) if traits else np.zeros((B, 0))
        im['birth_z'] = oz.copy()                                            # individuals-table location (species.py:694-697)
        g_all = np.concatenate([state['g'], og])
        z_all = np.concatenate([state['z'], oz])
        g_all, z_all, im['mutation'], im['mut_log'] = mutate_tskit(
            g_all, z_all, n0, mt, draws, arch.get('dom'), int(state['max_ind_idx']) + 1)
        traits = im['mutation']['traits']
    elif not burn:
        og = make_gametes(state['g'], pairs, nb, draws['recomb_keys'][:2 * B],
                          draws['start_homs'][:B], arch['paths']) if B else \
            np.zeros((0,) + state['g'].shape[1:], np.int8)
        oz = phenotype(og, traits, arch.get('dom')) if traits else np.zeros((B, 0))
        if arch.get('mutation') is not None:                                 # species.py:808-809
            og, oz, im['mutation'], im['mut_log'] = mutate(
                og, oz, arch['mutation'], draws, traits, arch.get('dom'), int(state['max_ind_idx']) + 1)
        g_all = np.concatenate([state['g'], og])
        z_all = np.concatenate([state['z'], oz]) if traits else None
    else:
        g_all = z_all = None
    x_all = np.concatenate([x, ox])
    y_all = np.concatenate([y, oy])
    age_all = np.concatenate([age, np.zeros(B, dtype=age.dtype)])
    sex_all = np.concatenate([state['sex'], osex]).astype(np.int8)
    idx_all = np.concatenate([state['idx'], ids])
    im['pre'] = dict(x=x_all, y=y_all, age=age_all, sex=sex_all, idx=idx_all, g=g_all, z=z_all)
    # a10 / a14
    N = species_density(dgs, x_all, y_all)
    im['N_rast'] = N
    d = calc_d(N, im['n_pairs_rast'], arch['K'], prm['R'], prm['b'], prm['lam'],
               prm['d_min'], prm['d_max'])
    im['d_rast'] = d
    # a15
    e_all = sample_env(rasters, x_all, y_all)
    mut_now = im.get('mutation', arch.get('mutation'))
    has_delet = (not burn) and mut_now is not None and len(mut_now['delet_loci']) > 0
    # species.py:449-451: selection if there are traits or mu_delet > 0
    selection = (not burn) and (bool(traits) or (mut_now is not None and mut_now['mu_delet'] > 0))
    if selection:
        cxa, cya = cells(x_all, y_all)
        # use_tskit: rows delet_loci_idxs (selection.py:86-88), else the loci themselves
        delet_rows = mut_now['delet_loci_idxs'] if (has_delet and mut_now.get('tskit_layout')) else \
            (mut_now['delet_loci'] if has_delet else None)
        delet = (delet_dosage(g_all, delet_rows), mut_now['delet_s']) if has_delet else None
        w = fitness(e_all, z_all, traits, cxa, cya, delet=delet)
    else:
        w = None
    im['fit_all'] = w
    p = death_probs(d, x_all, y_all, w, age_all, prm.get('max_age'))
    im['death_p'] = p
    # a16
    dead = mortality(p, draws['death_u'][:len(p)])
    keep = ~dead
    new = dict(x=x_all[keep], y=y_all[keep], age=age_all[keep], sex=sex_all[keep],
               idx=idx_all[keep], g=None if g_all is None else g_all[keep],
               z=None if z_all is None else z_all[keep],
               fit=None if w is None else w[keep], e=e_all[keep],
               max_ind_idx=int(state['max_ind_idx']) + B)
    im['n_deaths'] = int(dead.sum())
    return new, im
