"""Host-side *setup* of the density-grid stack and its triangulation.

Mirrors the reference's construction (utils/spatial.py:100-130 `_DensityGridStack.__init__`,
:270-360 `_make_density_grid(s)`): four coarse grids offset by 0 or window_width/2 in x and
y, window-landscape intersection areas (0 -> 1e-4), and the Delaunay triangulation that
`scipy.interpolate.griddata(method='cubic')` (spatial.py:144) builds over the union of their
points.  The reference rebuilds that triangulation on every call; it only depends on the
landscape dimensions and the window width, so it is built once here (same Qhull, same point
order => same diagonals) and handed to the device, where counting, gradient estimation and
Clough-Tocher evaluation run every time step (csrc/gnx_kernels.cuh).
"""
import numpy as np

from . import _lib


class DensityGridSetup:
    def __init__(self, land_dim, window_width=None):
        self.dim = (int(land_dim[0]), int(land_dim[1]))
        if window_width is None:
            window_width = round(0.1 * max(self.dim))            # spatial.py:110-111
        self.ww = window_width
        ww = window_width
        hww = ww / 2.
        pts, areas = [], []
        self.grid_shape, self.grid_cell0, self.grid_edges = [], [], []
        for x_edge, y_edge in ((True, True), (False, False), (True, False), (False, True)):
            xs = np.arange(0, self.dim[0] + ww, ww) if x_edge else np.arange(0 + hww, self.dim[0] + hww, ww)
            ys = np.arange(0, self.dim[1] + ww, ww) if y_edge else np.arange(0 + hww, self.dim[1] + hww, ww)
            gj, gi = np.meshgrid(xs, ys)
            wx = np.clip(np.minimum(gj + hww, self.dim[0]) - np.maximum(gj - hww, 0), 0, None)
            wy = np.clip(np.minimum(gi + hww, self.dim[1]) - np.maximum(gi - hww, 0), 0, None)
            a = wx * wy
            a[a == 0] = 0.0001                                     # spatial.py:319
            i0 = int(np.floor_divide(gi[0, 0] - hww * y_edge, ww) + y_edge)
            j0 = int(np.floor_divide(gj[0, 0] - hww * x_edge, ww) + x_edge)
            pts.append(np.stack([gi.ravel(), gj.ravel()], axis=1))   # (i, j) order, spatial.py:64-65
            areas.append(a.ravel())
            self.grid_shape.append(gi.shape)
            self.grid_cell0.append((i0, j0))
            self.grid_edges.append((int(x_edge), int(y_edge)))
        self.points = np.ascontiguousarray(np.vstack(pts), dtype=np.float64)
        self.areas = np.ascontiguousarray(np.hstack(areas), dtype=np.float64)
        self._triangulate()

    def _triangulate(self):
        from scipy.spatial import Delaunay        # same Qhull call griddata makes
        tri = Delaunay(self.points)
        self.simplices = np.ascontiguousarray(tri.simplices, dtype=np.int32)
        self.neighbors = np.ascontiguousarray(tri.neighbors, dtype=np.int32)
        indptr, indices = tri.vertex_neighbor_vertices
        self.nbr_indptr = np.ascontiguousarray(indptr, dtype=np.int32)
        self.nbr_indices = np.ascontiguousarray(indices, dtype=np.int32)
        hww = self.ww / 2.
        P = self.points
        li = np.rint(P[:, 0] / hww).astype(np.int64)
        lj = np.rint(P[:, 1] / hww).astype(np.int64)
        if not (np.allclose(li * hww, P[:, 0]) and np.allclose(lj * hww, P[:, 1])):
            raise NotImplementedError('density lattice is not regular')
        self.lat_ni = int(li.max()) + 1
        self.lat_nj = int(lj.max()) + 1
        if self.lat_ni * self.lat_nj != len(P):
            raise NotImplementedError('density lattice has holes')
        # the two triangles that tile each lattice square
        sq = -np.ones(((self.lat_ni - 1) * (self.lat_nj - 1), 2), dtype=np.int32)
        for t, s in enumerate(self.simplices):
            si, sj = li[s].min(), lj[s].min()
            if li[s].max() - si != 1 or lj[s].max() - sj != 1:
                raise NotImplementedError('triangulation does not conform to the lattice')
            k = si * (self.lat_nj - 1) + sj
            slot = 0 if sq[k, 0] < 0 else 1
            if sq[k, slot] >= 0:
                raise NotImplementedError('more than two triangles in a lattice square')
            sq[k, slot] = t
        if (sq < 0).any():
            raise NotImplementedError('lattice square without two triangles')
        self.square_tri = np.ascontiguousarray(sq)
        # the 4 grids are independent sets of the triangulation graph <=> a Gauss-Seidel
        # sweep in vertex order equals 4 parallel phases
        sizes = [a * b for a, b in self.grid_shape]
        grp = np.repeat(np.arange(4), sizes)
        src = np.repeat(np.arange(len(P)), np.diff(self.nbr_indptr))
        self.colourable = int(not np.any(grp[src] == grp[self.nbr_indices]))

    def to_struct(self):
        d = _lib.Density()
        d.window_width = float(self.ww)
        d.n_points = len(self.points)
        d.host_points = self.points.ctypes.data_as(_lib.c_double_p)
        d.host_areas = self.areas.ctypes.data_as(_lib.c_double_p)
        for g in range(4):
            d.grid_ni[g], d.grid_nj[g] = self.grid_shape[g]
            d.grid_i0[g], d.grid_j0[g] = self.grid_cell0[g]
            d.grid_x_edge[g], d.grid_y_edge[g] = self.grid_edges[g]
        d.n_tri = len(self.simplices)
        d.host_simplices = self.simplices.ctypes.data_as(_lib.c_int32_p)
        d.host_neighbors = self.neighbors.ctypes.data_as(_lib.c_int32_p)
        d.host_nbr_indptr = self.nbr_indptr.ctypes.data_as(_lib.c_int32_p)
        d.host_nbr_indices = self.nbr_indices.ctypes.data_as(_lib.c_int32_p)
        d.lat_ni, d.lat_nj = self.lat_ni, self.lat_nj
        d.host_square_tri = self.square_tri.ctypes.data_as(_lib.c_int32_p)
        d.colourable = self.colourable
        return d
