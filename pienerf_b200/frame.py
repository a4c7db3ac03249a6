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
    def test_gui(self, pose, intrinsics, W, H, paused=False, to_host=True, host_out=None):
        """trainer.py:531-602 with render_def=True, gui_sim=True, spp=1, downscale=1."""
        self._rays_key = None                                         # the reference regenerates rays every frame
        rays, rH, rW = self.rays(pose, intrinsics, W, H)
        if not paused:                                                # trainer.py:299-308
            IP_pos, IP_F, IP_dF = self.sim.get_IP_info()
            self.model.p_def, self.model.IP_F, self.model.IP_dF = IP_pos, IP_F, IP_dF
            self.sim.stepforward()
            self.frame += 1
        render = self.model.render_deformed if self.fused else self.model.rund_cuda
        outputs = render(rays["rays_o"], rays["rays_d"], staged=True, bg_color=None, perturb=False, **self.opt)
        image = outputs["image"].reshape(rH, rW, 3); depth = outputs["depth"].reshape(rH, rW); depth_0 = outputs["depth_0"].reshape(rH, rW)
        if to_host:                                                   # trainer.py:589-593: .cpu().numpy() of the frame
            if host_out is not None:
                host_out["image"].copy_(image, non_blocking=True); host_out["depth"].copy_(depth, non_blocking=True)
                host_out["depth_0"].copy_(depth_0, non_blocking=True)
                torch.cuda.current_stream().synchronize()
                return host_out
            return {"image": image.cpu().numpy(), "depth": depth.cpu().numpy(), "depth_0": depth_0.cpu().numpy()}
        return {"image": image, "depth": depth, "depth_0": depth_0}


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
