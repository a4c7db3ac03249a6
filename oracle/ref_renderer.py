"""oracle/ref_renderer.py -- TEST INFRASTRUCTURE / bench reference arm ONLY.

Drives the REFERENCE's own CUDA kernels (oracle/_ref/_ref_{raymarching,gridencoder,shencoder}.so, compiled
unmodified from /root/reference) through the reference's own loop structure: a restatement of
nerf/renderer.py:755-907 (rund_cuda), nerf/network.py:98-127 (fp32 nn.Linear stack, TF32 off) and
raymarching/raymarching.py wrappers, without the imports that are missing in this image (trimesh, warp, ...).
The two Warp kernels of get_pnts_in_grids (nerf/utils.py:355-443) are replaced by torch bincount / stable sort.
"""
import numpy as np
import torch
import torch.nn.functional as F

from .build_ref import load_ref


class ReferenceRenderer:
    def __init__(self, field, bits, bound=1.0, density_scale=1.0, min_near=0.2, device="cuda"):
        self.rm = load_ref("_ref_raymarching"); self.ge = load_ref("_ref_gridencoder"); self.se = load_ref("_ref_shencoder")
        if self.rm is None or self.ge is None or self.se is None:
            raise RuntimeError("oracle/_ref/*.so missing: run `python oracle/build_ref.py` where /root/reference exists")
        torch.backends.cuda.matmul.allow_tf32 = False                         # the reference never enables TF32
        dev = torch.device(device)
        self.dev = dev
        self.bound, self.density_scale, self.min_near = float(bound), float(density_scale), float(min_near)
        self.cascade = 1 + int(np.ceil(np.log2(bound))) if bound > 1 else 1
        self.H = 128
        self.emb = torch.from_numpy(field["embeddings"]).to(dev)
        self.offsets = torch.from_numpy(field["offsets"]).to(dev)
        self.S = float(np.log2(field["per_level_scale"])); self.base_res = int(field["base_resolution"])
        self.sigma_w = [torch.from_numpy(w).to(dev) for w in field["sigma_net"]]
        self.color_w = [torch.from_numpy(w).to(dev) for w in field["color_net"]]
        self.bits = torch.from_numpy(bits).to(dev)

    # nerf/network.py:98-127
    def field(self, x, d):
        B = x.shape[0]
        inp = ((x + self.bound) / (2 * self.bound)).contiguous()
        out = torch.empty(16, B, 2, device=self.dev, dtype=torch.float32)
        self.ge.grid_encode_forward(inp, self.emb, self.offsets, out, B, 3, 2, 16, self.S, self.base_res, None, 0, False, 0)
        h = out.permute(1, 0, 2).reshape(B, 32)
        h = F.relu(F.linear(h, self.sigma_w[0]), inplace=True)
        h = F.linear(h, self.sigma_w[1])
        sigma = torch.exp(h[..., 0]); geo = h[..., 1:]
        sh = torch.empty(B, 16, device=self.dev, dtype=torch.float32)
        self.se.sh_encode_forward(d.contiguous(), sh, B, 3, 4, None)
        h = torch.cat([sh, geo], dim=-1)
        h = F.relu(F.linear(h, self.color_w[0]), inplace=True)
        h = F.relu(F.linear(h, self.color_w[1]), inplace=True)
        return sigma, torch.sigmoid(F.linear(h, self.color_w[2]))

    @staticmethod
    def pnts_in_grids(p_def, bbmin, hgs, resolution):
        g = torch.floor((p_def - bbmin) / hgs).to(torch.int64)
        gid = g[:, 2] * resolution[1] * resolution[0] + g[:, 1] * resolution[0] + g[:, 0]
        n_grid = int(resolution[0] * resolution[1] * resolution[2])
        cnt = torch.bincount(gid, minlength=n_grid).to(torch.int32)
        bgn = (torch.cumsum(cnt, 0, dtype=torch.int32) - cnt).to(torch.int32)
        idx = torch.sort(gid, stable=True).indices.to(torch.int32)
        return cnt, bgn, idx, n_grid

    # nerf/renderer.py:755-907
    @torch.no_grad()
    def rund_cuda(self, rays_o, rays_d, p_def, p_ori, F_IP, dF_IP, IP_dx, dt_gamma=0.0, max_steps=1024, T_thresh=1e-2, max_iter_num=1,
                  hash_grid_size=0.06, cut=False, cut_bounds=(0.0,) * 6, num_seek_IP=1, bg_color=1.0, return_stats=False, first_noises=None):
        dev = self.dev; rm = self.rm
        rays_o = rays_o.contiguous().view(-1, 3); rays_d = rays_d.contiguous().view(-1, 3)
        N = rays_o.shape[0]
        cut_bounds = torch.tensor(cut_bounds, dtype=torch.float32, device=dev)
        bmin = p_def.min(axis=0).values; bmax = p_def.max(axis=0).values
        if cut:
            bmin = -self.bound * torch.ones(3, device=dev); bmax = self.bound * torch.ones(3, device=dev)
        bbmin = bmin - 1e-3 * torch.ones(3, device=dev); bbmax = bmax + 1e-3 * torch.ones(3, device=dev)
        resolution = torch.ceil((bbmax - bbmin) / hash_grid_size).to(torch.int32)
        aabb = torch.cat((bbmin, bbmax), dim=0)
        nears = torch.empty(N, device=dev); fars = torch.empty(N, device=dev)
        rm.near_far_from_aabb(rays_o, rays_d, aabb, N, self.min_near, nears, fars)
        weights_sum = torch.zeros(N, device=dev); depth = torch.zeros(N, device=dev); image = torch.zeros(N, 3, device=dev)
        n_vtx = p_ori.shape[0]
        pig_cnt, pig_bgn, pig_idx, n_grid = self.pnts_in_grids(p_def, bbmin, hash_grid_size, resolution)
        rays_alive = torch.arange(N, dtype=torch.int32, device=dev); rays_t = nears.clone()
        step = 0; n_samples = 0; iters = 0
        while step < max_steps:
            n_alive = rays_alive.shape[0]
            if n_alive <= 0:
                break
            n_step = max(min(N // n_alive, 8), 1)
            M = n_alive * n_step
            M += 128 - (M % 128)
            xyzs = torch.zeros(M, 3, device=dev); dirs = torch.zeros(M, 3, device=dev); deltas = torch.zeros(M, 2, device=dev)
            noises = first_noises if (first_noises is not None and step == 0) else torch.zeros(n_alive, device=dev)   # perturb if step == 0 (renderer.py:863)
            rm.march_rays_quadratic_bending(pig_cnt, pig_bgn, pig_idx, n_vtx, n_grid, p_def, p_ori, F_IP, dF_IP, max_iter_num, bbmin, bbmax,
                                            hash_grid_size, resolution, num_seek_IP, IP_dx, cut, cut_bounds, n_alive, n_step, rays_alive,
                                            rays_t, rays_o, rays_d, self.bound, dt_gamma, max_steps, self.cascade, self.H, self.bits, nears,
                                            fars, xyzs, dirs, deltas, noises)
            sigmas, rgbs = self.field(xyzs, dirs)
            sigmas = self.density_scale * sigmas
            if return_stats:
                n_samples += int((deltas[:, 0] != 0).sum()); iters += 1
            rm.composite_rays(n_alive, n_step, T_thresh, rays_alive, rays_t, sigmas, rgbs, deltas, weights_sum, depth, image)
            rays_alive = rays_alive[rays_alive >= 0]
            step += n_step
        depth_0 = depth
        image = image + (1 - weights_sum).unsqueeze(-1) * bg_color
        depth = torch.clamp(depth - nears, min=0) / (fars - nears)
        out = {"image": image, "depth": depth, "depth_0": depth_0, "weights_sum": weights_sum}
        if return_stats:
            out["n_samples"] = n_samples; out["iters"] = iters
        return out
