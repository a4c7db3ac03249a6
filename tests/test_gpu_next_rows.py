"""SURVEY.md 8(f) rows built after the hot path: density-grid maintenance (update_extra_state / packbits), checkpoint
keys, picking + drag force, and edge cases of the wavefront renderer (no hits, tiny ray sets, K = 2 with a real
Newton inverse warp)."""
import io

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

from oracle import render_oracle as ro  # noqa: E402
from tests.util import deformed_ip_state, small_scene  # noqa: E402


def _gpu(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def _model(density_scale=20.0, seed=0, amp=0.02, W=32, H=32):
    from pienerf_b200.network import NeRFNetwork
    body, field, bits, pose, intr = small_scene(W=W, H=H, seed=seed)
    p_ori, p_def, F, dF = deformed_ip_state(body, seed=seed, amp=amp)
    model = NeRFNetwork(bound=1, density_scale=density_scale).cuda().load_field(field)
    model.density_bitfield.copy_(_gpu(bits))
    model.p_ori, model.p_def, model.IP_F, model.IP_dF, model.IP_dx = _gpu(p_ori), _gpu(p_def), _gpu(F), _gpu(dF), 0.0525
    rays_o, rays_d = ro.get_rays(pose, intr, H, W)
    return model, field, bits, pose, intr, _gpu(rays_o)[None], _gpu(rays_d)[None]


def test_update_extra_state_vs_oracle(rng):
    """renderer.py:455-549 with the jitter switched off: every cell of the density grid holds density_scale * sigma at its
    centre (morton order), the bitfield is packbits(grid, min(mean, thresh)), and a second call is the EMA max."""
    model, field, *_ = _model(density_scale=3.0)
    model.reset_extra_state()
    model.update_extra_state(noise=False)
    grid = model.density_grid.cpu().numpy()
    H = model.grid_size
    assert grid.shape == (1, H ** 3) and (grid >= 0).all()
    cells = rng.integers(0, H, size=(3000, 3))
    idx = ro.morton3D(cells[:, 0].astype(np.uint32), cells[:, 1].astype(np.uint32), cells[:, 2].astype(np.uint32)).astype(np.int64)
    half = 1.0 / H
    xyz = ((2 * cells.astype(np.float32) / np.float32(H - 1) - 1) * np.float32(1.0 - half)).astype(np.float32)
    x01 = ((xyz + np.float32(1)) * np.float32(0.5)).astype(np.float32)                # grid.py:149 with bound = 1
    enc = ro.grid_encode(x01, field["embeddings"], field["offsets"], np.log2(field["per_level_scale"]), field["base_resolution"])
    enc = enc.transpose(1, 0, 2).reshape(x01.shape[0], -1)
    h = np.maximum(enc @ field["sigma_net"][0].T, 0) @ field["sigma_net"][1].T
    want = 3.0 * np.exp(h[:, 0])
    np.testing.assert_allclose(grid[0, idx], want, rtol=2e-4, atol=1e-6)
    thresh = min(float(np.clip(grid, 0, None).mean()), model.density_thresh)
    assert abs(model.mean_density - float(np.clip(grid, 0, None).mean())) < 1e-6
    np.testing.assert_array_equal(model.density_bitfield.cpu().numpy(), ro.packbits(grid.reshape(-1), thresh))
    # EMA: same field, decay < 1 -> max(old * decay, new) == new == old
    model.update_extra_state(noise=False, decay=0.5)
    np.testing.assert_allclose(model.density_grid.cpu().numpy(), grid, rtol=1e-6)
    # the jittered (reference) variant stays within the cell: bounded change, same occupancy statistics
    model.update_extra_state(noise=True, generator=torch.Generator(device="cuda").manual_seed(0))
    assert model.iter_density == 3 and torch.isfinite(model.density_grid).all()


def test_partial_density_update_runs():
    model, *_ = _model(density_scale=3.0)
    model.reset_extra_state()
    model.update_extra_state(noise=False)
    model.iter_density = 16                                                          # renderer.py:493: partial updates from here on
    before = model.density_grid.clone()
    model.update_extra_state(generator=torch.Generator(device="cuda").manual_seed(1))
    assert model.iter_density == 17 and (model.density_grid >= before * 0.95 - 1e-6).all()


def test_checkpoint_roundtrip_same_frame():
    """torch-ngp checkpoint layout (trainer.py:799-818): {'model': state_dict, 'mean_density': ...} with our key names."""
    from pienerf_b200.network import NeRFNetwork
    model, field, bits, pose, intr, ro_, rd_ = _model()
    opt = dict(max_iter_num=1, hash_grid_size=0.06, bound=1.0, cut=False, cut_bounds=[0.0] * 6, num_seek_IP=3, dt_gamma=0.0, max_steps=256, T_thresh=1e-2)
    want = model.render_deformed(ro_, rd_, **opt)["image"].clone()
    keys = set(model.state_dict().keys())
    assert {"encoder.embeddings", "encoder.offsets", "sigma_net.0.weight", "sigma_net.1.weight", "color_net.0.weight", "color_net.1.weight",
            "color_net.2.weight", "density_grid", "density_bitfield", "aabb_train", "aabb_infer"} <= keys
    buf = io.BytesIO()
    torch.save({"model": model.state_dict(), "mean_density": 0.123, "optimizer": {"junk": 1}}, buf)
    buf.seek(0)
    other = NeRFNetwork(bound=1, density_scale=20.0).cuda()
    other.load_checkpoint(buf)
    assert other.mean_density == pytest.approx(0.123)
    other.p_ori, other.p_def, other.IP_F, other.IP_dF, other.IP_dx = model.p_ori, model.p_def, model.IP_F, model.IP_dF, model.IP_dx
    got = other.render_deformed(ro_, rd_, **opt)["image"]
    assert torch.equal(torch.nan_to_num(got), torch.nan_to_num(want))
    with pytest.raises(ValueError):
        NeRFNetwork(bound=2).cuda().load_checkpoint({"model": model.state_dict()})


def test_picking_and_drag_force():
    """gui.py:561-577,647-667: unproject the picked pixel with depth_0, drag force towards it (clamped), clear."""
    from pienerf_b200.frame import FrameDriver, Options, drag_force, screen_to_world, world_to_screen
    from pienerf_b200.simulator import Simulator
    from pienerf_b200.synthetic import make_body
    model, field, bits, pose, intr, ro_, rd_ = _model(W=48, H=48)
    body = make_body("block64", dx=0.05, bound=1.0)
    sim = Simulator(dt=1e-2, iters=5, bbox=torch.tensor([2.0, 2.0, 2.0]), dx=0.05, stiff=1e5, base=torch.tensor([-1.0, -1.0, -1.0]))
    sim.set_points(body["pos"], body["mass"], body["mu"], body["lam"], body["pin"]).initialize()
    opt = Options.defaults(bound=1.0, W=48, H=48, max_steps=256, dt_gamma=0.0, max_iter_num=1, num_seek_IP=3)
    drv = FrameDriver(model, sim, opt, fused=True)
    out = drv.test_gui(pose, intr, 48, 48, paused=True, to_host=True)
    d0 = out["depth_0"]
    ys, xs = np.nonzero(d0 > 0)
    assert ys.size > 50
    x, y = int(xs[len(xs) // 2]), int(ys[len(ys) // 2])
    p, d = screen_to_world(x, y, d0, pose, intr, average_depth=2.5)
    assert d == pytest.approx(float(d0[y, x]))
    sx, sy, sz = world_to_screen(p, pose, intr)
    assert abs(sx - x) < 1e-6 and abs(sy - y) < 1e-6 and sz == pytest.approx(d)
    assert np.linalg.norm(p) < 1.0                                                   # the picked point lies on the body, not at infinity
    _, dflt = screen_to_world(0, 0, d0, pose, intr, average_depth=2.5)               # background pixel -> average depth
    assert dflt == 2.5
    f = drag_force(sim, 5, p + np.array([0.0, 100.0, 0.0]))
    assert np.linalg.norm(f) == pytest.approx(5e5)                                   # clamped like gui.py:575-577
    assert float(sim.dof_f.abs().sum()) > 0
    v0 = sim.dof_vel.clone(); sim.stepforward()
    assert not torch.equal(sim.dof_vel, v0)
    assert drag_force(sim, None, None) is None and float(sim.dof_f.abs().sum()) == 0


def test_wavefront_edge_cases():
    model, field, bits, pose, intr, ro_, rd_ = _model()
    opt = dict(max_iter_num=1, hash_grid_size=0.06, bound=1.0, cut=False, cut_bounds=[0.0] * 6, num_seek_IP=3, dt_gamma=0.0, max_steps=256, T_thresh=1e-2)
    # rays that all look away from the body: nothing hits the IP box, every pass is empty
    away = model.render_deformed(ro_, -rd_, mode=3, **opt)
    assert int(away["stats"][0]) == 0 and int(away["stats"][1]) == 0
    assert (away["image"] == 1).all() and (away["weights_sum"] == 0).all()
    # a single ray, and a ray count that is not a multiple of anything
    for n in (1, 37):
        a = model.render_deformed(ro_[:, :n], rd_[:, :n], mode=3, **opt)
        b = model.render_deformed(ro_[:, :n], rd_[:, :n], mode=0, **opt)
        assert torch.equal(torch.nan_to_num(a["image"]), torch.nan_to_num(b["image"]))
    # K = 2 and a real Newton inverse warp (max_iter_num > 1): same emit decisions as the reference-order search
    opt2 = dict(opt, num_seek_IP=2, max_iter_num=100)
    w = {k: (v.clone() if torch.is_tensor(v) else v) for k, v in model.render_deformed(ro_, rd_, mode=3, **opt2).items()}
    l = {k: (v.clone() if torch.is_tensor(v) else v) for k, v in model.render_deformed(ro_, rd_, mode=2, **opt2).items()}
    assert int(w["stats"][0]) == int(l["stats"][0]) > 1000
    assert float((w["image"] - l["image"]).abs().max()) < 1e-3
    # a tiny per-ray cap: passes stop at max_steps samples per ray exactly like the fused kernel
    opt3 = dict(opt, max_steps=48)
    a = model.render_deformed(ro_, rd_, mode=3, **opt3); sa = int(a["stats"][0]); ia = a["image"].clone()
    b = model.render_deformed(ro_, rd_, mode=0, **opt3)
    assert sa == int(b["stats"][0]) and torch.equal(torch.nan_to_num(ia), torch.nan_to_num(b["image"]))


def test_point_sampling_on_the_gpu(tmp_path):
    """SURVEY 8f.4 point sampling (pienerf_b200/sampling.py; bit-for-bit parity with the reference's main_sample.py is the CPU test
    tests/test_sampling.py).  Here: the same torch code on cuda:0 reproduces the reference's point set up to the handful of lattice
    points whose density sits at the threshold (CUDA vs CPU expf), and it runs on a real NeRFNetwork.density."""
    import argparse
    import os
    from pienerf_b200.ply import read_ply_vertices
    from pienerf_b200.sampling import AdaptiveUniformSampling
    from tests.sampling_cases import CASES, BlobField
    G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_sampling.npz"))
    o = dict(CASES["a"]); opt = argparse.Namespace(**o)
    offsets = torch.rand((64, 3), dtype=torch.float32, generator=torch.Generator().manual_seed(5))
    cpu_p, cpu_v = AdaptiveUniformSampling(opt, BlobField(), device="cpu").sample(write=False, points_tmp=offsets)
    gpu_p, gpu_v = AdaptiveUniformSampling(opt, BlobField(), device="cuda:0").sample(write=False, points_tmp=offsets)
    assert abs(gpu_p.shape[0] - cpu_p.shape[0]) <= 0.01 * cpu_p.shape[0] and abs(gpu_p.shape[0] - G["a_points"].shape[0]) <= 0.02 * cpu_p.shape[0]
    a = {tuple(np.round(r, 5)) for r in cpu_p.numpy()}; b = {tuple(np.round(r, 5)) for r in gpu_p.cpu().numpy()}
    assert len(a & b) >= 0.98 * len(a)
    # a real field: the synthetic hash-grid + MLP of the render tests, through NeRFNetwork.density (CUDA hash-grid kernel)
    model = _model()[0]
    opt2 = argparse.Namespace(**dict(o, density_threshold=0.02, workspace="ws/field", exp_name="pts"))
    s = AdaptiveUniformSampling(opt2, model, out_dir=str(tmp_path))
    dens = s.get_density(s._lattice())
    if float(dens.max()) <= opt2.density_threshold:
        pytest.skip("synthetic field has no density above the threshold on this lattice")
    try:
        pts, vols = s.sample()
    except AssertionError as e:                       # a random-init field may have a flat density: no boundary points
        assert "No boundary points" in str(e) or "No points" in str(e)
        return
    assert float(s.get_density(pts).min()) > opt2.density_threshold and float(vols.min()) > 0
    ply = read_ply_vertices(os.path.join(str(tmp_path), "field", "pts.ply"))
    assert ply["x"].shape[0] == pts.shape[0] and np.allclose(ply["vp"], vols.cpu().numpy().astype(np.float64))
