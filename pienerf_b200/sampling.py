"""Host-side mirror of main_sample.py (AdaptiveUniformSampling): turns a trained density field into the point cloud the
simulator loads (x, y, z, vp), SURVEY.md 8f.4.  Same class name, constructor arguments (opt, model), option names
(`bound`, `density_threshold`, `sub_res`, `sub_coeff`, `hash_grid_size`, `cut`, `cut_bounds`, `workspace`, `exp_name`) and
method names as the reference; the four Warp kernels of main_sample.py:26-140 become whole-array torch operations on the
model's device (this is a once-per-asset tool, not a per-frame path), each citing the kernel it replaces.

What the reference leaves to chance is pinned here, and said so:
  * `get_sub_bgn` (main_sample.py:71-79) hands out output ranges with an atomic counter; here the ranges ascend with the
    cell index (an exclusive prefix sum) — the same points, in a reproducible order.
  * `get_grid_coords` (main_sample.py:46-61) stores each lattice point's cell at `hash_code_g(cell)`.  The lattice spans
    [-bound, bound] with res points (spacing 2*bound/(res-1)) while cells are 2*bound/res wide, so the last point of every
    row lands in cell index `res`: its slot aliases the first cell of the next row (a write race in the reference) or lies
    past the end of the array (an out-of-bounds write).  Here the point with the highest index wins a contested slot (the
    outcome of running the reference's threads in order) and writes past the end are dropped; `get_sub_grid`'s reads past
    the end of `grid_density` (main_sample.py:112-127) return 0.
"""
import os

import torch

from .ply import write_ply_xyz


def write_ply(filename, points, volumes, binary=True):
    """main_sample.py:14-23: vertex schema x, y, z, vp (all f8)."""
    if not binary:
        raise NotImplementedError("text ply output")
    write_ply_xyz(filename, points, {"vp": volumes})


class AdaptiveUniformSampling:
    """main_sample.py:142-308."""

    def __init__(self, opt, model, device=None, out_dir=None):
        self.device = torch.device(device) if device is not None else torch.device("cuda:0")
        self.dtype = torch.float32
        self.opt = opt
        self.bound = opt.bound
        self.threshold = opt.density_threshold
        self.res = opt.sub_res
        self.model = model.to(self.device)
        self.grid_size = 2 * self.bound / self.res
        root = out_dir if out_dir is not None else os.path.join(os.getcwd(), "model")         # main_sample.py:160-162 (beside the script)
        self.write_path = os.path.join(root, str(opt.workspace).split("/")[-1], opt.exp_name)

    # ---- main_sample.py:164-181
    def get_density(self, x):
        x = x.to(self.device)
        density = self.model.density(x)["sigma"]
        return 1 - torch.exp(-density / 128.0)

    def p2g(self, x):
        return torch.floor((x + self.bound) / self.grid_size)

    def g2p(self, g):
        return g * self.grid_size - self.bound

    def hash_code_g(self, g):
        return int(g[2] * self.res * self.res + g[1] * self.res + g[0])

    def hash_code(self, x):
        return self.hash_code_g(self.p2g(x))

    # ---- main_sample.py:183-204 + nerf/utils.py:355-443: volume = cell volume / points in the cell
    def get_point_volumes(self, pts):
        pts = pts.to(self.device, torch.float32)
        f32 = torch.float32
        bbmin = pts.min(dim=0).values - 1e-3 * torch.ones(3, dtype=f32, device=self.device)
        bbmax = pts.max(dim=0).values + 1e-3 * torch.ones(3, dtype=f32, device=self.device)
        hgs = self.opt.hash_grid_size
        resolution = torch.ceil((bbmax - bbmin) / hgs).to(torch.int64)
        n_grid = int(resolution[2] * resolution[1] * resolution[0])
        g = torch.floor((pts - bbmin) / torch.tensor(hgs, dtype=f32, device=self.device)).to(torch.int64)   # p2g of nerf/utils.py:389-407
        gid = g[:, 2] * resolution[1] * resolution[0] + g[:, 1] * resolution[0] + g[:, 0]
        if int(gid.max()) >= n_grid or int(gid.min()) < 0:
            raise RuntimeError("get_point_volumes: a point fell outside its own bounding grid")
        cnt = torch.bincount(gid, minlength=n_grid)
        return (hgs ** 3 / cnt.to(f32))[gid]

    def _lattice(self):
        """main_sample.py:207-227: res^3 lattice points, point (i*res + j)*res + k at (xs[k], ys[j], zs[i])."""
        o = self.opt
        if o.cut:
            cb = o.cut_bounds
            for a in (0, 2, 4):
                if cb[a] < -o.bound: cb[a] = -o.bound
            for a in (1, 3, 5):
                if cb[a] > o.bound: cb[a] = o.bound
            assert cb[0] < cb[1] and cb[2] < cb[3] and cb[4] < cb[5]
            xs = torch.linspace(cb[0], cb[1], self.res)
            ys = torch.linspace(cb[2], cb[3], self.res)
            zs = torch.linspace(cb[4], cb[5], self.res)
            x_grid, y_grid, z_grid = torch.meshgrid(zs, ys, xs, indexing="ij")
        else:
            xs = torch.linspace(-o.bound, o.bound, self.res)
            x_grid, y_grid, z_grid = torch.meshgrid(xs, xs, xs, indexing="ij")
        return torch.stack([z_grid, y_grid, x_grid], dim=-1).reshape(-1, 3).to(self.device)

    @torch.no_grad()
    def sample(self, write=True, points_tmp=None):
        """Returns (pts [n,3] f32, vols [n] f32) and, like the reference, writes <write_path>.ply.  `points_tmp` overrides
        the uniform random offsets every cell draws its boundary points from (default: torch.rand, main_sample.py:270)."""
        res = self.res
        n_grid = res ** 3
        f32 = torch.float32
        dev = self.device
        grid_pts = self._lattice()
        assert grid_pts.shape[0] > 0, "No grid points, check params!"
        grid_density = self.get_density(grid_pts).to(dev, f32)
        bound = torch.tensor(self.bound, dtype=f32, device=dev)
        gs = torch.tensor(self.grid_size, dtype=f32, device=dev)

        # get_grid_coords (main_sample.py:46-61): cell of every lattice point, stored at the cell's hash
        g = torch.floor((grid_pts + bound) / gs).to(torch.int64)
        gid = g[:, 2] * res * res + g[:, 1] * res + g[:, 0]
        ok = (gid >= 0) & (gid < n_grid)
        tid = torch.arange(n_grid, device=dev)
        winner = torch.full((n_grid,), -1, dtype=torch.int64, device=dev)
        winner.scatter_reduce_(0, gid[ok], tid[ok], reduce="amax", include_self=True)
        grid_coords = torch.zeros(n_grid, 3, dtype=torch.int64, device=dev)
        has = winner >= 0
        grid_coords[has] = g[winner[has]]

        # get_sub_grid (main_sample.py:97-139): density gradient over the cell's 8 corners -> how many points to add
        def corner_density(dx, dy, dz):
            h = (grid_coords[:, 2] + dz) * res * res + (grid_coords[:, 1] + dy) * res + (grid_coords[:, 0] + dx)
            inside = (h >= 0) & (h < n_grid)
            d = torch.zeros(n_grid, dtype=f32, device=dev)
            d[inside] = grid_density[h[inside]]
            return d
        d0 = corner_density(0, 0, 0); d1 = corner_density(0, 0, 1); d2 = corner_density(0, 1, 0); d3 = corner_density(0, 1, 1)
        d4 = corner_density(1, 0, 0); d5 = corner_density(1, 0, 1); d6 = corner_density(1, 1, 0); d7 = corner_density(1, 1, 1)
        grad_x = d4 + d5 + d6 + d7 - (d0 + d1 + d2 + d3)
        grad_y = d2 + d3 + d6 + d7 - (d0 + d1 + d4 + d5)
        grad_z = d1 + d3 + d5 + d7 - (d0 + d2 + d4 + d6)
        grad_norm = torch.sqrt(grad_x * grad_x + grad_y * grad_y + grad_z * grad_z)
        flat = grad_norm == 0
        sub_mins = grid_coords.to(f32) * gs - bound                                   # g2p(g0)
        sub_maxs = (grid_coords + 1).to(f32) * gs - bound                             # g2p(g7)
        sub_coeff = torch.tensor(self.opt.sub_coeff, dtype=f32, device=dev)
        sub_dims = ((sub_maxs - sub_mins)[:, 0] * sub_coeff * torch.tensor(float(res), dtype=f32, device=dev) * grad_norm).to(torch.int32)
        sub_dims[flat] = 0
        sub_mins[flat] = 0
        sub_maxs[flat] = 0

        # get_sub_bgn (main_sample.py:71-79), in cell order
        per_cell = sub_dims.to(torch.int64) ** 3
        sub_bgn = torch.cumsum(per_cell, 0) - per_cell
        tot = int(per_cell.sum())
        max_add = int(sub_dims.max()) ** 3
        if points_tmp is None:
            points_tmp = torch.rand((max_add, 3), dtype=f32, device=dev)              # [0,1)^3, shared by all cells
        points_tmp = points_tmp.to(dev, f32)
        assert tot > 0, "No boundary points sampled, check params!"

        # get_pnts_add (main_sample.py:81-95): cell c takes the first sub_dims[c]^3 offsets, scaled into its box
        cell = torch.repeat_interleave(torch.arange(n_grid, device=dev), per_cell)
        local = torch.arange(tot, device=dev) - sub_bgn[cell]
        pnts_add = (sub_maxs - sub_mins)[cell] * points_tmp[local] + sub_mins[cell]

        pts = torch.cat((pnts_add, grid_pts + 0.5 * 2 * self.opt.bound / float(res)), dim=0)
        density = self.get_density(pts)
        pts = pts[density > self.threshold]
        assert pts.shape[0] > 0, "No points sampled, check params!"
        vols = self.get_point_volumes(pts)
        if write:
            os.makedirs(os.path.dirname(self.write_path), exist_ok=True)
            write_ply(self.write_path + ".ply", pts.cpu().numpy(), vols.cpu().numpy())
        return pts, vols
