"""Host-side mirror of nerf/network.py `NeRFNetwork` + the two CUDA-ray render loops of nerf/renderer.py
(`run_cuda` eval branch :332-388, `rund_cuda` :755-907) over the B200 operators, plus the fused entry points.

Two ways to render a deformed frame, same inputs, same outputs:
  * `rund_cuda(...)`           the reference's wavefront loop verbatim (host-driven, per-op drop-in kernels) —
                               exists for A/B parity against the reference structure;
  * `render_deformed(...)`     ONE C-ABI call (`pn_render_deformed`): device-resident, no host sync — the product path.
State-dict keys match torch-ngp checkpoints (trainer.py:799-818): encoder.embeddings, encoder.offsets,
sigma_net.{0,1}.weight, color_net.{0,1,2}.weight, density_bitfield, aabb_train, aabb_infer.
"""
import ctypes as C
import math

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import raymarching
from ._lib import PN_IO_WEIGHTS_READY, DeformT, FieldT, FrameIoT, check, dptr, lib, stream_ptr
from .gridencoder import GridEncoder
from .shencoder import SHEncoder


import os

DEFAULT_RENDER_MODE = int(os.environ.get("PN_RENDER_MODE", "3"))   # 3 wavefront (product path), 0 fused tcgen05 kernel (see pn_render_deformed)


class NeRFNetwork(nn.Module):
    def __init__(self, encoding="hashgrid", encoding_dir="sphere_harmonics", num_layers=2, hidden_dim=64, geo_feat_dim=15,
                 num_layers_color=3, hidden_dim_color=64, bound=1, cuda_ray=True, density_scale=1, min_near=0.2,
                 density_thresh=0.01, bg_radius=-1, **kwargs):
        super().__init__()
        if encoding != "hashgrid" or encoding_dir != "sphere_harmonics":
            raise NotImplementedError("the hot path is hashgrid + sphere_harmonics (main_gui.py:24-32)")
        if bg_radius > 0:
            raise NotImplementedError("background model (bg_radius > 0) is outside the sim+render hot path")
        # nerf/renderer.py:74-113
        self.bound = bound
        self.cascade = 1 + math.ceil(math.log2(bound))
        self.grid_size = 128
        self.density_scale = density_scale
        self.min_near = min_near
        self.density_thresh = density_thresh
        self.bg_radius = bg_radius
        self.cuda_ray = cuda_ray
        aabb = torch.FloatTensor([-bound, -bound, -bound, bound, bound, bound])
        self.register_buffer("aabb_train", aabb)
        self.register_buffer("aabb_infer", aabb.clone())
        self.register_buffer("density_bitfield", torch.zeros(self.cascade * self.grid_size ** 3 // 8, dtype=torch.uint8))
        # density-grid maintenance state (renderer.py:95-106); the frame loop only reads the bitfield
        self.register_buffer("density_grid", torch.zeros(self.cascade, self.grid_size ** 3))
        self.mean_density = 0.0
        self.iter_density = 0
        # nerf/network.py:30-71
        self.num_layers, self.hidden_dim, self.geo_feat_dim = num_layers, hidden_dim, geo_feat_dim
        self.encoder = GridEncoder(input_dim=3, num_levels=16, level_dim=2, base_resolution=16, log2_hashmap_size=19,
                                   desired_resolution=2048 * bound, gridtype="hash", align_corners=False)
        self.in_dim = self.encoder.output_dim
        self.sigma_net = nn.ModuleList([
            nn.Linear(self.in_dim if l == 0 else hidden_dim, 1 + geo_feat_dim if l == num_layers - 1 else hidden_dim, bias=False)
            for l in range(num_layers)])
        self.num_layers_color, self.hidden_dim_color = num_layers_color, hidden_dim_color
        self.encoder_dir = SHEncoder(input_dim=3, degree=4)
        self.in_dim_dir = self.encoder_dir.output_dim
        self.color_net = nn.ModuleList([
            nn.Linear(self.in_dim_dir + geo_feat_dim if l == 0 else hidden_dim_color, 3 if l == num_layers_color - 1 else hidden_dim_color, bias=False)
            for l in range(num_layers_color)])
        # deformation state wired on by the frame driver (main_gui.py:50-56)
        self.p_ori = self.p_def = self.IP_F = self.IP_dF = None
        self.IP_dx = None
        self._workspace = None
        self._stats = None

    # ------------------------------------------------------------------ synthetic field loading
    def load_field(self, field):
        """Load a pienerf_b200.synthetic.make_field() dict (random-init stand-in for a checkpoint)."""
        with torch.no_grad():
            self.encoder.embeddings.copy_(torch.from_numpy(field["embeddings"]))
            for lin, w in zip(self.sigma_net, field["sigma_net"]):
                lin.weight.copy_(torch.from_numpy(w))
            for lin, w in zip(self.color_net, field["color_net"]):
                lin.weight.copy_(torch.from_numpy(w))
        return self

    # ------------------------------------------------------------------ nerf/network.py:98-127
    @torch.no_grad()
    def forward(self, x, d):
        """Per-op path: hash-grid kernel -> torch fp32 Linear stack -> SH kernel -> Linear stack."""
        x = self.encoder(x, bound=self.bound)
        h = x
        for l in range(self.num_layers):
            h = self.sigma_net[l](h)
            if l != self.num_layers - 1:
                h = F.relu(h, inplace=True)
        sigma = torch.exp(h[..., 0])                                 # trunc_exp forward (activation.py:9-11)
        geo_feat = h[..., 1:]
        d = self.encoder_dir(d)
        h = torch.cat([d, geo_feat], dim=-1)
        for l in range(self.num_layers_color):
            h = self.color_net[l](h)
            if l != self.num_layers_color - 1:
                h = F.relu(h, inplace=True)
        return sigma, torch.sigmoid(h)

    # ------------------------------------------------------------------ nerf/network.py:129-146
    @torch.no_grad()
    def density(self, x):
        """sigma and geometry features only (hash-grid kernel + the two sigma_net layers)."""
        h = self.encoder(x, bound=self.bound)
        for l in range(self.num_layers):
            h = self.sigma_net[l](h)
            if l != self.num_layers - 1:
                h = F.relu(h, inplace=True)
        return {"sigma": torch.exp(h[..., 0]), "geo_feat": h[..., 1:]}

    # ------------------------------------------------------------------ nerf/renderer.py:396-401, 455-549
    def reset_extra_state(self):
        self.density_grid.zero_()
        self.mean_density = 0.0
        self.iter_density = 0

    @torch.no_grad()
    def update_extra_state(self, decay=0.95, S=128, noise=True, generator=None):
        """Density-grid maintenance: sample sigma at (jittered) cell centres of every cascade, EMA-max into
        `density_grid`, re-threshold and pack the occupancy bitfield the march kernels read.  Full update for the
        first 16 calls, then the reference's half-random / half-occupied partial update.  `noise=False` (cell centres,
        no jitter) makes the result reproducible for the parity test; the reference always jitters."""
        if not self.cuda_ray:
            return
        dev, H = self.density_bitfield.device, self.grid_size
        tmp_grid = -torch.ones_like(self.density_grid)

        def rand_like(t):
            return torch.rand(t.shape, dtype=t.dtype, device=t.device, generator=generator)

        def sample(cas, coords, indices):
            xyzs = 2 * coords.float() / (H - 1) - 1
            bound = min(2 ** cas, self.bound)
            half_grid_size = bound / H
            cas_xyzs = xyzs * (bound - half_grid_size)
            if noise:
                cas_xyzs = cas_xyzs + (rand_like(cas_xyzs) * 2 - 1) * half_grid_size
            sig = self.density(cas_xyzs)["sigma"].reshape(-1) * self.density_scale
            tmp_grid[cas, indices] = sig

        if self.iter_density < 16:
            ar = torch.arange(H, dtype=torch.int32, device=dev)
            for xs in ar.split(S):
                for ys in ar.split(S):
                    for zs in ar.split(S):
                        xx, yy, zz = torch.meshgrid(xs, ys, zs, indexing="ij")
                        coords = torch.stack([xx.reshape(-1), yy.reshape(-1), zz.reshape(-1)], dim=-1).contiguous()
                        indices = raymarching.morton3D(coords).long()
                        for cas in range(self.cascade):
                            sample(cas, coords, indices)
        else:
            N = H ** 3 // 4
            for cas in range(self.cascade):
                coords = torch.randint(0, H, (N, 3), device=dev, generator=generator, dtype=torch.int32)
                indices = raymarching.morton3D(coords).long()
                occ = torch.nonzero(self.density_grid[cas] > 0).squeeze(-1)
                if occ.numel() > 0:
                    pick = torch.randint(0, occ.shape[0], [N], dtype=torch.long, device=dev, generator=generator)
                    occ = occ[pick]
                    occ_coords = raymarching.morton3D_invert(occ.int())
                    indices = torch.cat([indices, occ], dim=0)
                    coords = torch.cat([coords, occ_coords], dim=0)
                sample(cas, coords, indices)
        valid = (self.density_grid >= 0) & (tmp_grid >= 0)
        self.density_grid[valid] = torch.maximum(self.density_grid[valid] * decay, tmp_grid[valid])
        self.mean_density = torch.mean(self.density_grid.clamp(min=0)).item()
        self.iter_density += 1
        density_thresh = min(self.mean_density, self.density_thresh)
        self.density_bitfield = raymarching.packbits(self.density_grid, density_thresh, self.density_bitfield)

    # ------------------------------------------------------------------ nerf/trainer.py:799-818,853-906
    def load_checkpoint(self, path_or_state, strict=False):
        """torch-ngp `.pth` checkpoint (or its `model` state dict): same key names as this module, so a real chair /
        trex asset loads directly; extra keys (optimizer, ema, bg_net ...) are ignored unless strict."""
        state = torch.load(path_or_state, map_location="cpu") if isinstance(path_or_state, (str, bytes)) or hasattr(path_or_state, "read") else path_or_state
        if "model" in state and isinstance(state["model"], dict):
            if "mean_density" in state:
                self.mean_density = float(state["mean_density"])
            state = state["model"]
        own = self.state_dict()
        if "encoder.embeddings" in state and tuple(state["encoder.embeddings"].shape) != tuple(own["encoder.embeddings"].shape):
            raise ValueError(f"checkpoint hash table {tuple(state['encoder.embeddings'].shape)} does not match bound={self.bound} "
                             f"({tuple(own['encoder.embeddings'].shape)}): construct NeRFNetwork with the checkpoint's bound")
        return self.load_state_dict({k: v for k, v in state.items() if strict or k in own}, strict=strict)

    def _field_struct(self, embeddings=None):
        """pn_field_t over this module's parameters (`embeddings`: another copy of the hash table, same shape).  The fused
        kernels hard-wire the architecture of nerf/network.py:30-71 (16 x 2 hash grid -> 32-64-16 | SH(4) + 15 -> 31-64-64-3,
        bias-free): anything else is refused here instead of reading weights out of bounds."""
        shapes = [tuple(l.weight.shape) for l in self.sigma_net] + [tuple(l.weight.shape) for l in self.color_net]
        enc = self.encoder
        if (shapes != [(64, 32), (16, 64), (64, 31), (64, 64), (3, 64)] or enc.num_levels != 16 or enc.level_dim != 2 or enc.input_dim != 3
                or getattr(enc, "align_corners", False) or getattr(enc, "interp_id", 0) != 0 or getattr(enc, "gridtype_id", 0) != 0):
            raise NotImplementedError(f"the fused field kernels implement the default NeRFNetwork only (got layers {shapes}, "
                                      f"{enc.num_levels} x {enc.level_dim} grid); use forward() / rund_cuda() for other shapes")
        emb = self.encoder.embeddings.data if embeddings is None else embeddings
        if emb.shape != self.encoder.embeddings.shape:
            raise ValueError("hash-table copy has a different shape")
        f = FieldT()
        self._keep = [emb, self.encoder.offsets] + [l.weight.data for l in self.sigma_net] + [l.weight.data for l in self.color_net]
        f.embeddings, f.offsets = dptr(self._keep[0], "embeddings", torch.float32), dptr(self._keep[1], "offsets", torch.int32)
        f.S = float(np.log2(self.encoder.per_level_scale)); f.H = int(self.encoder.base_resolution); f.L = int(self.encoder.num_levels)
        f.bound = float(self.bound)
        f.w_sigma0, f.w_sigma1 = dptr(self._keep[2], "sigma_net.0", torch.float32), dptr(self._keep[3], "sigma_net.1", torch.float32)
        f.w_color0, f.w_color1, f.w_color2 = (dptr(self._keep[4], "color_net.0", torch.float32), dptr(self._keep[5], "color_net.1", torch.float32),
                                              dptr(self._keep[6], "color_net.2", torch.float32))
        return f

    @torch.no_grad()
    def forward_fused(self, x, d, mode=0):
        """Whole NeRFNetwork.forward as one kernel (pn_field_forward)."""
        x = x.to(torch.float32).contiguous().view(-1, 3); d = d.to(torch.float32).contiguous().view(-1, 3)
        M = x.shape[0]
        sigmas = torch.empty(M, dtype=torch.float32, device=x.device)
        rgbs = torch.empty(M, 3, dtype=torch.float32, device=x.device)
        f = self._field_struct()
        check(lib.pn_field_forward(C.byref(f), dptr(x), dptr(d), M, dptr(sigmas), dptr(rgbs), int(mode), stream_ptr()))
        return sigmas, rgbs

    @torch.no_grad()
    def mlp_only(self, enc, d):
        """sigma_net + color_net on pre-encoded features `enc` [M,32] and directions `d` [M,3] (network.py:105-127) through the
        frame renderer's tcgen05 pipeline (pn_mlp_forward) — the MLP pass measured by itself."""
        enc = enc.to(torch.float32).contiguous().view(-1, 32); d = d.to(torch.float32).contiguous().view(-1, 3)
        M = enc.shape[0]
        need = int(lib.pn_mlp_workspace_bytes(M))
        if getattr(self, "_mlp_ws", None) is None or self._mlp_ws.numel() < need or self._mlp_ws.device != enc.device:
            self._mlp_ws = torch.empty(need, dtype=torch.uint8, device=enc.device)
        sigmas = torch.empty(M, dtype=torch.float32, device=enc.device)
        rgbs = torch.empty(M, 3, dtype=torch.float32, device=enc.device)
        f = self._field_struct()
        check(lib.pn_mlp_forward(C.byref(f), dptr(enc), dptr(d), M, dptr(sigmas), dptr(rgbs), dptr(self._mlp_ws), need, stream_ptr()))
        return sigmas, rgbs

    # ------------------------------------------------------------------ nerf/renderer.py:332-388
    @torch.no_grad()
    def run_cuda(self, rays_o, rays_d, dt_gamma=0, bg_color=None, perturb=False, max_steps=1024, T_thresh=1e-2, fused_field=False, **kwargs):
        prefix = rays_o.shape[:-1]
        rays_o = rays_o.contiguous().view(-1, 3); rays_d = rays_d.contiguous().view(-1, 3)
        N = rays_o.shape[0]; device = rays_o.device
        nears, fars = raymarching.near_far_from_aabb(rays_o, rays_d, self.aabb_infer, self.min_near)
        if bg_color is None:
            bg_color = 1
        weights_sum = torch.zeros(N, dtype=torch.float32, device=device)
        depth = torch.zeros(N, dtype=torch.float32, device=device)
        image = torch.zeros(N, 3, dtype=torch.float32, device=device)
        rays_alive = torch.arange(N, dtype=torch.int32, device=device)
        rays_t = nears.clone()
        step = 0
        field = self.forward_fused if fused_field else self.forward
        while step < max_steps:
            n_alive = rays_alive.shape[0]
            if n_alive <= 0:
                break
            n_step = max(min(N // n_alive, 8), 1)
            xyzs, dirs, deltas = raymarching.march_rays(n_alive, n_step, rays_alive, rays_t, rays_o, rays_d, self.bound,
                                                        self.density_bitfield, self.cascade, self.grid_size, nears, fars, 128,
                                                        perturb if step == 0 else False, dt_gamma, max_steps)
            sigmas, rgbs = field(xyzs, dirs)
            sigmas = self.density_scale * sigmas
            raymarching.composite_rays(n_alive, n_step, rays_alive, rays_t, sigmas, rgbs, deltas, weights_sum, depth, image, T_thresh)
            rays_alive = rays_alive[rays_alive >= 0]
            step += n_step
        image = image + (1 - weights_sum).unsqueeze(-1) * bg_color
        depth = torch.clamp(depth - nears, min=0) / (fars - nears)
        return {"depth": depth.view(*prefix), "image": image.view(*prefix, 3), "weights_sum": weights_sum}

    # ------------------------------------------------------------------ nerf/renderer.py:755-907
    @torch.no_grad()
    def rund_cuda(self, rays_o, rays_d, dt_gamma=0, bg_color=None, perturb=False, max_steps=1024, T_thresh=1e-2, fused_field=False,
                  return_stats=False, **kwargs):
        dtype = torch.float32
        prefix = rays_o.shape[:-1]
        rays_o = rays_o.contiguous().view(-1, 3); rays_d = rays_d.contiguous().view(-1, 3)
        N = rays_o.shape[0]; device = rays_o.device
        max_iter_num = kwargs.get("max_iter_num"); hgs = kwargs.get("hash_grid_size"); bound = kwargs.get("bound", self.bound)
        cut = bool(kwargs.get("cut")); num_seek_IP = kwargs.get("num_seek_IP")
        cut_bounds = torch.tensor(kwargs.get("cut_bounds") or [0.0] * 6, dtype=dtype, device=device)
        p_def = self.p_def.contiguous().cuda(); p_ori = self.p_ori.contiguous().cuda()
        F_IP = self.IP_F.contiguous().cuda(); dF_IP = self.IP_dF.contiguous().cuda()
        bbmin, bbmax, resolution = raymarching.ip_bbox(p_def, hgs, cut, bound)      # renderer.py:782-791, one kernel
        aabb = torch.cat((bbmin, bbmax), dim=0)
        nears, fars = raymarching.near_far_from_aabb(rays_o, rays_d, aabb, self.min_near)
        if bg_color is None:
            bg_color = 1
        weights_sum = torch.zeros(N, dtype=dtype, device=device)
        depth = torch.zeros(N, dtype=dtype, device=device)
        image = torch.zeros(N, 3, dtype=dtype, device=device)
        n_vtx = p_ori.shape[0]
        n_grid = int(resolution[2] * resolution[1] * resolution[0])                  # host sync, as in the reference
        assert p_def.shape == p_ori.shape and n_vtx > 0
        pig_cnt, pig_bgn, pig_idx = raymarching.get_pnts_in_grids(n_vtx, n_grid, p_def, bbmin, bbmax, hgs, resolution)
        rays_alive = torch.arange(N, dtype=torch.int32, device=device)
        rays_t = nears.clone()
        step = 0; n_samples = 0; iters = 0
        field = self.forward_fused if fused_field else self.forward
        while step < max_steps:
            n_alive = rays_alive.shape[0]
            if n_alive <= 0:
                break
            n_step = max(min(N // n_alive, 8), 1)
            xyzs, dirs, deltas = raymarching.march_rays_quadratic_bending(
                pig_cnt, pig_bgn, pig_idx, n_vtx, n_grid, p_def, p_ori, F_IP, dF_IP, max_iter_num, bbmin, bbmax, hgs, resolution,
                num_seek_IP, self.IP_dx, cut, cut_bounds, n_alive, n_step, rays_alive, rays_t, rays_o, rays_d, self.bound,
                self.density_bitfield, self.cascade, self.grid_size, nears, fars, 128, perturb if step == 0 else False, dt_gamma, max_steps)
            sigmas, rgbs = field(xyzs, dirs)
            sigmas = self.density_scale * sigmas
            if return_stats:
                n_samples += int((deltas[:, 0] != 0).sum()); iters += 1
            raymarching.composite_rays(n_alive, n_step, rays_alive, rays_t, sigmas, rgbs, deltas, weights_sum, depth, image, T_thresh)
            rays_alive = rays_alive[rays_alive >= 0]
            step += n_step
        depth_0 = depth
        image = image + (1 - weights_sum).unsqueeze(-1) * bg_color
        depth = torch.clamp(depth - nears, min=0) / (fars - nears)
        out = {"depth": depth.view(*prefix), "image": image.view(*prefix, 3), "depth_0": depth_0.view(*prefix), "weights_sum": weights_sum}
        if return_stats:
            out["n_samples"] = n_samples; out["iters"] = iters
        return out

    # ------------------------------------------------------------------ fused frame (product path)
    def deform_struct(self, ip_state=None, dt_gamma=0, bg_color=None, max_steps=1024, T_thresh=1e-2, **kwargs):
        """pn_deform_t for one frame: `ip_state` = (p_def, p_ori, F, dF) fp32 contiguous CUDA tensors (default: the state wired
        onto the module, main_gui.py:50-56) + the flat `opt` namespace the reference passes as **vars(opt)."""
        if ip_state is None:
            ip_state = (self.p_def, self.p_ori, self.IP_F, self.IP_dF)
        keep = [t if (t.dtype == torch.float32 and t.is_contiguous()) else t.to(torch.float32).contiguous() for t in ip_state]
        d = DeformT()
        d.p_def, d.p_ori, d.F_IP, d.dF_IP = [dptr(t, "IP state", torch.float32) for t in keep]
        d.n_vtx = keep[0].shape[0]; d.IP_dx = float(self.IP_dx)
        d.density_bitfield = dptr(self.density_bitfield, "density_bitfield", torch.uint8)
        d.bound, d.cascade, d.grid_size = float(self.bound), int(self.cascade), int(self.grid_size)
        d.min_near, d.density_scale, d.dt_gamma = float(self.min_near), float(self.density_scale), float(dt_gamma)
        d.max_steps, d.T_thresh, d.max_iter_num = int(max_steps), float(T_thresh), int(kwargs.get("max_iter_num"))
        d.hgs, d.cut = float(kwargs.get("hash_grid_size")), int(bool(kwargs.get("cut")))
        cb = kwargs.get("cut_bounds") or [0.0] * 6
        for i in range(6):
            d.cut_bounds[i] = float(cb[i])
        d.num_seek_IP = int(kwargs.get("num_seek_IP"))
        d.bg_color = 1.0 if bg_color is None else float(bg_color)
        return d, keep

    def workspace_bytes(self, N, n_vtx, **kwargs):
        return int(lib.pn_render_workspace_bytes(int(N), int(n_vtx), float(kwargs.get("bound", self.bound)), float(kwargs.get("hash_grid_size"))))

    @staticmethod
    def check_stats(stats):
        """Raise on the error bits of a finished frame's stats (pn_render_deformed: stats[4]).  Synchronises."""
        st = [int(v) for v in stats.tolist()]
        if st[4] & 1:
            raise RuntimeError("render_deformed: the deformed IP bounding box needs more grid cells than the scene box allows "
                               "(diverged simulation / IPs far outside [-bound, bound]); the frame is not valid")
        if st[4] & 2:
            raise RuntimeError(f"render_deformed: {st[5]} rays were cut short (sample list and passes exhausted); "
                               "render with a larger workspace or mode 0")
        return st

    @torch.no_grad()
    def render_deformed(self, rays_o, rays_d, staged=False, dt_gamma=0, bg_color=None, perturb=False, max_steps=1024, T_thresh=1e-2,
                        mode=None, out=None, workspace=None, stats=None, io=None, embeddings=None, ip_state=None, noises=None, max_passes=None,
                        weights_ready=False, **kwargs):
        """renderer.py:587-599 -> rund_cuda semantics in one device-resident call.  Returns the same dict.
        mode 3: wavefront (default); 0: fused warp-cooperative kernel + tcgen05 MLP; 1: same with the fp32 SIMT MLP; 2: one lane
        per ray.  `workspace` / `stats` / `embeddings` / `ip_state` / `io` (_lib.FrameIoT) let a frame pipeline keep several
        frames in flight (pipeline.py); by default the module's own single workspace is used."""
        if mode is None:
            mode = DEFAULT_RENDER_MODE
        prefix = rays_o.shape[:-1]
        rays_o = rays_o.to(torch.float32).contiguous().view(-1, 3); rays_d = rays_d.to(torch.float32).contiguous().view(-1, 3)
        N = rays_o.shape[0]; device = rays_o.device
        if perturb and noises is None:                                        # raymarching.py:407-409: torch.rand per ray (spp > 1 accumulation)
            noises = torch.rand(N, dtype=torch.float32, device=device)
        if noises is not None:
            if mode != 3:
                raise NotImplementedError("perturbed ray starts are implemented by the wavefront renderer (mode 3)")
            noises = noises.to(torch.float32).contiguous().view(-1)
            if io is None:
                io = FrameIoT()
            io.noises = dptr(noises, "noises", torch.float32)
        if max_passes or weights_ready:                                       # frame pipelines: fewer launches per frame (see pn_frame_io_t)
            if io is None:
                io = FrameIoT()
            io.max_passes = int(max_passes or 0)
            io.flags = PN_IO_WEIGHTS_READY if weights_ready else 0
        d, keep = self.deform_struct(ip_state, dt_gamma=dt_gamma, bg_color=bg_color, max_steps=max_steps, T_thresh=T_thresh, **kwargs)
        need = self.workspace_bytes(N, d.n_vtx, **kwargs)
        if workspace is None:
            if self._workspace is None or self._workspace.numel() < need or self._workspace.device != device:
                self._workspace = torch.empty(need, dtype=torch.uint8, device=device)
            workspace = self._workspace
        if stats is None:
            if self._stats is None or self._stats.device != device:
                self._stats = torch.zeros(8, dtype=torch.int64, device=device)
            stats = self._stats
        if workspace.numel() < need:
            raise ValueError(f"workspace too small: {workspace.numel()} < {need} bytes")
        if out is None:
            out = {"image": torch.empty(N, 3, dtype=torch.float32, device=device), "depth": torch.empty(N, dtype=torch.float32, device=device),
                   "depth_0": torch.empty(N, dtype=torch.float32, device=device), "weights_sum": torch.empty(N, dtype=torch.float32, device=device)}
        # kernels this call enqueues: bbox, 4 x IP grid, frame setup, IP pack, 3 x neighbourhood lists, (3 per pass | 1 fused), stats
        n_pass = int(lib.pn_render_pass_count(int(max_steps)))
        if io is not None and 0 < io.max_passes < n_pass:
            n_pass = int(io.max_passes)
        fused_prep = mode != 2 and (math.ceil((2 * float(kwargs.get("bound", self.bound)) + 2e-3) / float(kwargs.get("hash_grid_size"))) + 1) ** 3 <= 64 * 1024
        self._render_launches = (3 if fused_prep else 11) + (1 + 3 * n_pass if mode == 3 else 1)
        if io is not None:
            self._render_launches += (1 if io.epoch else 0) + (1 if io.n_signal else 0) - (1 if (io.flags & PN_IO_WEIGHTS_READY and mode == 3) else 0)
        f = self._field_struct(embeddings)
        check(lib.pn_render_deformed_ex(C.byref(f), C.byref(d), dptr(rays_o), dptr(rays_d), N, dptr(out["image"]), dptr(out["depth"]),
                                        dptr(out["depth_0"]), dptr(out["weights_sum"]), dptr(workspace), workspace.numel(), dptr(stats),
                                        int(mode), C.byref(io) if io is not None else None, stream_ptr()))
        if io is not None and io.pix:
            return {"image": out["image"], "depth": out["depth"], "depth_0": out["depth_0"], "weights_sum": out["weights_sum"], "stats": stats}
        return {"image": out["image"].view(*prefix, 3), "depth": out["depth"].view(*prefix), "depth_0": out["depth_0"].view(*prefix),
                "weights_sum": out["weights_sum"], "stats": stats}
