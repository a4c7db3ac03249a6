"""Headless per-frame driver: the `NeRFSimGUI.test_step -> Trainer.test_gui -> Trainer.test_step` chain of the
reference (nerf/gui.py:556-645, nerf/trainer.py:284-329,531-602) without the window.

Order of operations per frame is the reference's, including the one-step render lag: the IP state is read
BEFORE `stepforward()` (trainer.py:303-308), so frame k shows the body after k-1 steps.
"""
import numpy as np
import torch

from . import raymarching
from .synthetic import orbit_intrinsics, orbit_pose


class Options(dict):
    """The flat `opt` namespace of get_opts.py that the reference passes as **vars(opt) into render_deformed."""
    __getattr__ = dict.__getitem__
    __setattr__ = dict.__setitem__

    @staticmethod
    def defaults(**over):
        o = Options(max_steps=1024, T_thresh=1e-2, bound=2.0, dt_gamma=1 / 128, min_near=0.2, density_thresh=10, W=1920, H=1080,
                    radius=5, fovy=50, cut=False, cut_bounds=[0.0, 2.0, -2.0, 1.0, -1.42, 0.92], num_seek_IP=1, max_iter_num=100,
                    sim_dt=1e-2, sim_dx=0.05, sim_iters=10, sim_stiff=1e5)
        o.update(over)
        o["hash_grid_size"] = 1.2 * o["sim_dx"]                       # get_opts.py:96
        o["num_seek_IP"] = max(min(3, o["num_seek_IP"]), 1)            # get_opts.py:97
        return o


class FrameDriver:
    def __init__(self, model, sim, opt, fused=True):
        self.model, self.sim, self.opt, self.fused = model, sim, opt, fused
        # main_gui.py:50-56
        IP_pos, IP_F, IP_dF = sim.get_IP_info()
        model.p_ori = IP_pos
        model.p_def = IP_pos
        model.IP_F = IP_F
        model.IP_dF = IP_dF
        model.IP_dx = sim.dx * 1.05
        self.frame = 0
        self._rays = None
        self._rays_key = None

    def rays(self, pose, intrinsics, W, H, downscale=1):
        rH, rW = int(H * downscale), int(W * downscale)
        key = (pose.tobytes(), tuple(float(v) for v in intrinsics), rW, rH)
        if self._rays_key != key:
            self._rays = raymarching.get_rays(torch.from_numpy(pose).unsqueeze(0), np.asarray(intrinsics) * downscale, rH, rW, -1)
            self._rays_key = key
        return self._rays, rH, rW

    @torch.no_grad()
    def test_gui(self, pose, intrinsics, W, H, paused=False, to_host=True, host_out=None, spp=1, downscale=1):
        """trainer.py:531-602 with render_def=True, gui_sim=True.  downscale < 1 renders int(H*downscale) x int(W*downscale) rays
        with scaled intrinsics and upsamples the result with nearest-neighbour interpolation (:537-539,577-586); spp > 1 is the GUI's
        accumulation index, used only as `perturb` (:565)."""
        self._rays_key = None                                         # the reference regenerates rays every frame
        rays, rH, rW = self.rays(pose, intrinsics, W, H, downscale)
        if not paused:                                                # trainer.py:299-308
            IP_pos, IP_F, IP_dF = self.sim.get_IP_info()
            self.model.p_def, self.model.IP_F, self.model.IP_dF = IP_pos, IP_F, IP_dF
            self.sim.stepforward()
            self.frame += 1
        render = self.model.render_deformed if self.fused else self.model.rund_cuda
        outputs = render(rays["rays_o"], rays["rays_d"], staged=True, bg_color=None, perturb=False if spp == 1 else spp, **self.opt)
        image = outputs["image"].reshape(rH, rW, 3); depth = outputs["depth"].reshape(rH, rW); depth_0 = outputs["depth_0"].reshape(rH, rW)
        if downscale != 1:                                            # trainer.py:577-586
            import torch.nn.functional as F
            image = F.interpolate(image.permute(2, 0, 1)[None], size=(H, W), mode="nearest")[0].permute(1, 2, 0).contiguous()
            depth = F.interpolate(depth[None, None], size=(H, W), mode="nearest")[0, 0]
            depth_0 = F.interpolate(depth_0[None, None], size=(H, W), mode="nearest")[0, 0]
        if to_host:                                                   # trainer.py:589-593: .cpu().numpy() of the frame
            if host_out is not None:
                host_out["image"].copy_(image, non_blocking=True); host_out["depth"].copy_(depth, non_blocking=True)
                host_out["depth_0"].copy_(depth_0, non_blocking=True)
                torch.cuda.current_stream().synchronize()
                return host_out
            return {"image": image.cpu().numpy(), "depth": depth.cpu().numpy(), "depth_0": depth_0.cpu().numpy()}
        return {"image": image, "depth": depth, "depth_0": depth_0}


def screen_to_world(x, y, depth_0, pose, intrinsics, average_depth=0.0):
    """nerf/gui.py:647-657: unproject pixel (x, y) with the rendered `depth_0` [H,W] (0 = nothing hit -> average_depth)."""
    fx, fy, cx, cy = [float(v) for v in intrinsics]
    H, W = depth_0.shape
    # the GUI indexes its [W,H]-reshaped buffer with (x, y); same pixel here as depth_0[row y, column x]
    d = float(depth_0[min(max(int(y), 0), H - 1), min(max(int(x), 0), W - 1)])
    if d == 0.0:
        d = float(average_depth)
    cam = np.array([(x - cx) / fx * d, (y - cy) / fy * d, d, 1.0])
    return (np.asarray(pose, dtype=np.float64) @ cam)[:3], d


def world_to_screen(pos, pose, intrinsics):
    """nerf/gui.py:659-667."""
    fx, fy, cx, cy = [float(v) for v in intrinsics]
    P = np.eye(4); P[:np.asarray(pose).shape[0], :] = np.asarray(pose, dtype=np.float64)
    xc, yc, zc = (np.linalg.inv(P) @ np.array([*pos, 1.0]))[:3]
    return xc / zc * fx + cx, yc / zc * fy + cy, zc


def drag_force(sim, sid, target, force_scale=1.0, limit=5e5):
    """nerf/gui.py:561-577,584-586: the mouse-drag force on IP `sid` towards world point `target` (None clears it)."""
    if sid is None:
        sim.clear_force()
        return None
    p0 = sim.get_IP_info()[0][sid].double().cpu().numpy()
    f = force_scale * 1e5 * (np.asarray(target, dtype=np.float64) - p0)
    n = np.linalg.norm(f)
    if n > limit:
        f *= limit / n
    sim.update_force(int(sid), f)
    return f


class DistFrameDriver:
    """The frame loop on N GPUs of one node (one process per GPU; N = 1 works without a process group).

    Per frame: rank 0 reads the IP state and advances the simulator (state BEFORE the step is
    rendered, trainer.py:303-308), ONE broadcast of the packed [n_ip,39] fp32 IP state, every rank renders its
    interleaved tiles with the fused kernel, ONE gather of [pixels,5] rows (rgb, depth, depth_0) to rank 0, and an
    asynchronous copy of the frame into one of two pinned host buffers (the host consumes frame k while the GPUs
    work on k+1; `wait_host()` blocks until a given frame has landed).
    """

    def __init__(self, model, sim, opt, tile=16, weights=None, overlap_sim=True):
        import torch.distributed as dist
        self.model, self.sim, self.opt, self.tile = model, sim, opt, tile
        self.world = dist.get_world_size() if dist.is_initialized() else 1
        self.rank = dist.get_rank() if dist.is_initialized() else 0
        self.dev = next(model.parameters()).device
        IP_pos, IP_F, IP_dF = sim.get_IP_info()
        model.p_ori = IP_pos
        model.p_def, model.IP_F, model.IP_dF = IP_pos, IP_F, IP_dF
        model.IP_dx = sim.dx * 1.05
        self.W, self.H = opt.W, opt.H
        from .dist import ip_state_views
        self.ipbuf = torch.zeros(sim.n_ip * 39, dtype=torch.float32, device=self.dev)
        self.ip_views = ip_state_views(self.ipbuf, sim.n_ip)
        self.copy_stream = torch.cuda.Stream(device=self.dev)
        self.sim_stream = torch.cuda.Stream(device=self.dev, priority=-1)
        self.overlap_sim = overlap_sim
        npx = self.W * self.H
        self.host = [{"image": torch.empty(npx, 3, dtype=torch.float32).pin_memory(), "depth": torch.empty(npx, dtype=torch.float32).pin_memory(),
                      "depth_0": torch.empty(npx, dtype=torch.float32).pin_memory()} for _ in range(2)] if self.rank == 0 else None
        self.copy_done = [None, None]
        self.frame_id = 0
        self.launches = 0
        self.set_partition(weights)

    def set_partition(self, weights=None):
        from .dist import tile_partition
        self.parts = tile_partition(self.H, self.W, self.world, self.tile, weights)
        self.my = torch.from_numpy(self.parts[self.rank]).to(self.dev)
        from .dist import PlanarFrameGather
        self.gathers = [PlanarFrameGather(self.parts, self.dev) for _ in range(2)]   # double-buffered with the host frames
        self._rays = None

    def rays(self, pose_host, intrinsics):
        full = raymarching.get_rays(pose_host[None] if pose_host.dim() == 2 else pose_host, intrinsics, self.H, self.W)
        self.launches += 1
        if self.world == 1:
            return full["rays_o"], full["rays_d"]
        self.launches += 2
        return full["rays_o"][:, self.my].contiguous(), full["rays_d"][:, self.my].contiguous()

    @torch.no_grad()
    def frame(self, pose_host, intrinsics, to_host=True, regenerate_rays=True, profile_events=None, paused=False):
        """One GUI frame.  Returns (render outputs of this rank, index of the pinned host buffer or None)."""
        import torch.distributed as dist
        from . import _lib
        from .dist import broadcast_ip_state
        if regenerate_rays or self._rays is None:                    # trainer.py:541-543: rays from the host pose every frame
            self._rays = self.rays(pose_host, intrinsics)
        rays_o, rays_d = self._rays
        step_done = None
        if not paused:
            pos, F, dF = self.ip_views
            if self.rank == 0:
                self.sim.get_IP_info(out=self.ip_views)                  # state BEFORE the step (trainer.py:303-306)
                self.launches += 1
            if self.world > 1:
                broadcast_ip_state(self.ipbuf)                           # the other ranks start rendering at once ...
                self.launches += 1
            if self.rank == 0:
                # ... while rank 0 advances the simulator (trainer.py:308).  The frame renders the state read BEFORE the
                # step, so the step only has to be ordered after that read: it runs on a high-priority side stream,
                # concurrently with this rank's render kernels, and the frame ends by waiting for it.
                if self.overlap_sim:
                    read_done = torch.cuda.Event(); read_done.record()
                    with torch.cuda.stream(self.sim_stream):
                        self.sim_stream.wait_event(read_done)
                        self.sim.stepforward()
                        step_done = torch.cuda.Event(); step_done.record()
                else:
                    self.sim.stepforward()
                    step_done = None
                self.launches += 2 + 4 * self.sim.iters + 1
            self.model.p_def, self.model.IP_F, self.model.IP_dF = pos, F, dF
        if profile_events is not None:                               # (start, stop[, flat list of per-pass field-kernel events])
            _lib.lib.pn_set_profile_events(_lib.vp(profile_events[0].cuda_event), _lib.vp(profile_events[1].cuda_event))
            if len(profile_events) > 2:
                arr = (_lib.vp * len(profile_events[2]))(*[e.cuda_event for e in profile_events[2]])
                _lib.lib.pn_set_profile_event_list(arr, len(profile_events[2]))
        gather_frame = self.world > 1 or to_host
        slot = self.frame_id & 1 if gather_frame else None
        gat = self.gathers[slot] if gather_frame else None
        if gather_frame and self.rank == 0 and self.copy_done[slot] is not None:
            torch.cuda.current_stream().wait_event(self.copy_done[slot])    # frame k-2's host copy has read this slot's buffers
        out = self.model.render_deformed(rays_o, rays_d, out=gat.out if gather_frame else None, **self.opt)
        self.launches += self.model._render_launches
        if profile_events is not None:
            _lib.lib.pn_set_profile_events(_lib.vp(0), _lib.vp(0))
            _lib.lib.pn_set_profile_event_list(None, 0)
        if gather_frame:
            fb = gat()                                                   # the renderer wrote straight into the send segment
            if self.world > 1:
                self.launches += 1 + (3 if self.rank == 0 else 0)
            if to_host and self.rank == 0:
                ready = torch.cuda.Event(); ready.record()
                with torch.cuda.stream(self.copy_stream):
                    self.copy_stream.wait_event(ready)
                    for k in ("image", "depth", "depth_0"):                 # trainer.py:589-593 .cpu().numpy() of the frame
                        self.host[slot][k].copy_(fb[k], non_blocking=True)
                    done = torch.cuda.Event(); done.record()
                self.copy_done[slot] = done
        if step_done is not None:
            torch.cuda.current_stream().wait_event(step_done)           # frame k is complete only when step k is
        self.frame_id += 1
        return out, slot

    def wait_host(self, slot=None):
        """Block until the async frame copies (all, or the given slot) have landed; returns the pinned buffer(s)."""
        if self.rank != 0:
            return None
        for i, ev in enumerate(self.copy_done):
            if ev is not None and (slot is None or i == slot):
                ev.synchronize()
        return self.host if slot is None else self.host[slot]

    @torch.no_grad()
    def calibrate(self, pose_host, intrinsics, frames=3):
        """Sim-aware tile weights: rank 0 also runs the simulator (S ms), so it takes a share f0 of the tiles with
        f0*R + S = (1-f0)*R/(N-1), R = whole-frame render time.  Measured here with CUDA events, agreed by broadcast."""
        import torch.distributed as dist
        if self.world == 1:
            return None
        tr, tsim = [], []
        for _ in range(frames):
            e = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
            if self.rank == 0:
                e[0].record(); self.sim.get_IP_info(); self.sim.stepforward(); e[1].record()
            rays_o, rays_d = self.rays(pose_host, intrinsics)
            e[2].record(); self.model.render_deformed(rays_o, rays_d, **self.opt); e[3].record()
            torch.cuda.synchronize()
            tr.append(e[2].elapsed_time(e[3]))
            if self.rank == 0:
                tsim.append(e[0].elapsed_time(e[1]))
        t = torch.tensor([min(tr), min(tsim) if tsim else 0.0], dtype=torch.float64, device=self.dev)
        allr = [torch.zeros_like(t) for _ in range(self.world)]
        dist.all_gather(allr, t)
        R = float(sum(x[0] for x in allr))                                    # whole frame = sum of the equal shares
        S = float(allr[0][1]) + 0.05                                          # + pack / bookkeeping on rank 0
        n = self.world
        f0 = max(0.0, min(1.0 / n, (R / (n - 1) - S) / (R * n / (n - 1))))
        w = [f0] + [(1.0 - f0) / (n - 1)] * (n - 1)
        self.set_partition(w)
        return {"render_ms": R, "sim_ms": S, "weights": w}


def build_scene(config, device="cuda", seed=0, density_scale=1.0, solver="inverse"):
    """Synthetic stand-in for `main_gui.py:20-56`: model + simulator + options for a named config."""
    from .network import NeRFNetwork
    from .simulator import Simulator
    from .synthetic import CONFIGS, make_body, make_field, occupancy_bitfield
    cfg = CONFIGS[config]
    bound = cfg["bound"]
    opt = Options.defaults(bound=bound, W=cfg["W"], H=cfg["H"], max_steps=cfg["max_steps"], T_thresh=cfg["T_thresh"], dt_gamma=cfg["dt_gamma"],
                           min_near=cfg["min_near"], max_iter_num=cfg["max_iter_num"], num_seek_IP=cfg["num_seek_IP"], sim_dx=cfg["sim_dx"],
                           radius=cfg["radius"], fovy=cfg["fovy"], cut=cfg["cut"], cut_bounds=cfg.get("cut_bounds", [0.0] * 6))
    body = make_body(cfg["body"], dx=cfg["sim_dx"], bound=bound, seed=seed)
    field = make_field(bound=bound, seed=seed)
    model = NeRFNetwork(bound=bound, cuda_ray=True, density_scale=density_scale, min_near=opt.min_near).to(device)
    model.load_field(field)
    model.density_bitfield.copy_(torch.from_numpy(occupancy_bitfield(body["pos"], 0.6 * cfg["sim_dx"], bound=bound)))
    sim = Simulator(dt=opt.sim_dt, iters=opt.sim_iters, bbox=torch.tensor([2.0 * bound] * 3), dx=opt.sim_dx, stiff=opt.sim_stiff,
                    base=torch.tensor([-bound] * 3), solver=solver, device=device)
    sim.set_points(body["pos"], body["mass"], body["mu"], body["lam"], body["pin"]).initialize()
    pose = orbit_pose(radius=cfg["radius"])
    intr = orbit_intrinsics(cfg["W"], cfg["H"], cfg["fovy"])
    return model, sim, opt, pose, intr, body, field
