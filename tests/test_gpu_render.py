"""Whole-frame parity: the drop-in wavefront loop and the fused device-resident renderer vs the numpy oracle's
rund_cuda, on undeformed and deformed bodies.  RGB bar from BASELINE.json: 1e-3 absolute (undeformed frame)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

from oracle import render_oracle as ro  # noqa: E402
from tests.util import deformed_ip_state, psnr, small_scene  # noqa: E402

f32 = np.float32


def _gpu(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def _setup(amp, W=40, H=40, density_scale=20.0, seed=0):
    from pienerf_b200.network import NeRFNetwork
    body, field, bits, pose, intr = small_scene(W=W, H=H, seed=seed)
    p_ori, p_def, F, dF = deformed_ip_state(body, seed=seed, amp=amp)
    model = NeRFNetwork(bound=1, density_scale=density_scale).cuda().load_field(field)
    model.density_bitfield.copy_(_gpu(bits))
    model.p_ori, model.p_def, model.IP_F, model.IP_dF, model.IP_dx = _gpu(p_ori), _gpu(p_def), _gpu(F), _gpu(dF), 0.0525
    rays_o, rays_d = ro.get_rays(pose, intr, H, W)
    return model, field, bits, (p_ori, p_def, F, dF), rays_o, rays_d


def _bad_fraction(a, b, tol):
    return float((np.abs(a - b).max(-1) > tol).mean())


@pytest.mark.parametrize("amp,K", [(0.0, 3), (0.0, 1), (0.03, 3)])
def test_frame_vs_oracle(amp, K):
    model, field, bits, (p_ori, p_def, F, dF), rays_o, rays_d = _setup(amp)
    kw = dict(dt_gamma=0.0, max_steps=256, T_thresh=1e-2)
    want = ro.rund_cuda(ro.OracleField(field), rays_o, rays_d, p_def, p_ori, F, dF, 0.0525, bits, 1.0, 1, min_near=0.2,
                        density_scale=20.0, max_iter_num=1, hash_grid_size=0.06, num_seek_IP=K, return_stats=True, **kw)
    opt = dict(max_iter_num=1, hash_grid_size=0.06, bound=1.0, cut=False, cut_bounds=[0.0] * 6, num_seek_IP=K)
    ro_, rd_ = _gpu(rays_o)[None], _gpu(rays_d)[None]
    loop = model.rund_cuda(ro_, rd_, return_stats=True, **kw, **opt)
    fused = model.render_deformed(ro_, rd_, **kw, **opt)
    fused = {k: (v.clone() if torch.is_tensor(v) else v) for k, v in fused.items()}                       # mode 0: tcgen05 MLP
    simt = {k: (v.clone() if torch.is_tensor(v) else v) for k, v in model.render_deformed(ro_, rd_, mode=1, **kw, **opt).items()}
    lane = {k: (v.clone() if torch.is_tensor(v) else v) for k, v in model.render_deformed(ro_, rd_, mode=2, **kw, **opt).items()}
    wave = {k: (v.clone() if torch.is_tensor(v) else v) for k, v in model.render_deformed(ro_, rd_, mode=3, **kw, **opt).items()}
    torch.cuda.synchronize()
    hit = want["weights_sum"] > 0
    assert hit.sum() > 100 and want["n_samples"] > 1000
    for name, got in (("loop", loop), ("fused-tc", fused), ("fused-simt", simt), ("fused-lane", lane), ("wavefront", wave)):
        img = got["image"][0].cpu().numpy(); ws = got["weights_sum"].cpu().numpy(); d0 = got["depth_0"][0].cpu().numpy()
        # knife-edge occupancy flips (fp32 FMA contraction) can move a handful of silhouette pixels
        assert _bad_fraction(img, want["image"], 1e-3) <= 0.01, (name, _bad_fraction(img, want["image"], 1e-3))
        assert np.median(np.abs(img - want["image"])[hit]) < 1e-4, name
        assert psnr(img, want["image"]) > 45, (name, psnr(img, want["image"]))
        assert _bad_fraction(ws[:, None], want["weights_sum"][:, None], 1e-3) <= 0.01
        assert _bad_fraction(d0[:, None], want["depth_0"][:, None], 2e-3) <= 0.01
        assert (img[~hit & (ws == 0)] == 1).all()
    # the two CUDA paths share every device function: they must agree far tighter than either does with numpy
    a = loop["image"][0].cpu().numpy(); b = simt["image"][0].cpu().numpy(); c = fused["image"][0].cpu().numpy()
    assert _bad_fraction(a, b, 1e-4) <= 0.002
    assert _bad_fraction(b, c, 2e-4) <= 0.002                                       # bf16x3 tensor-core MLP vs fp32 SIMT MLP
    # wavefront and fused kernels: same per-ray sample sequence, same tensor-core MLP, same compositing order
    for k in ("image", "depth", "depth_0", "weights_sum"):
        assert torch.equal(torch.nan_to_num(wave[k], nan=-7.0), torch.nan_to_num(fused[k], nan=-7.0)), k   # depth is NaN for a miss, as in the reference
    assert int(wave["stats"][0]) == int(fused["stats"][0]) and int(wave["stats"][1]) == int(fused["stats"][1])
    # mode 2 walks the IP grid in the reference's own order (march_device.cuh); the packed/flattened search of modes
    # 0/1/3 must take exactly the same emit / skip decisions
    assert int(lane["stats"][0]) == int(simt["stats"][0]) == int(wave["stats"][0])
    assert abs(int(simt["stats"][0]) - loop["n_samples"]) <= 0.002 * loop["n_samples"] + 2
    assert abs(loop["n_samples"] - want["n_samples"]) <= 0.01 * want["n_samples"]


def test_undeformed_equals_plain_renderer():
    """on an undeformed body rund_cuda == run_cuda restricted to the IP bbox (identity warp)."""
    model, field, bits, (p_ori, p_def, F, dF), rays_o, rays_d = _setup(0.0)
    kw = dict(dt_gamma=0.0, max_steps=256, T_thresh=1e-2)
    opt = dict(max_iter_num=1, hash_grid_size=0.06, bound=1.0, cut=False, cut_bounds=[0.0] * 6, num_seek_IP=3)
    ro_, rd_ = _gpu(rays_o)[None], _gpu(rays_d)[None]
    fused = model.render_deformed(ro_, rd_, **kw, **opt)
    bbmin = p_ori.min(0) - f32(1e-3); bbmax = p_ori.max(0) + f32(1e-3)
    model.aabb_infer.copy_(_gpu(np.concatenate([bbmin, bbmax])))
    plain = model.run_cuda(ro_, rd_, **kw)
    a = fused["image"][0].cpu().numpy(); b = plain["image"][0].cpu().numpy()
    assert _bad_fraction(a, b, 1e-3) <= 0.01 and psnr(a, b) > 45


def test_cut_and_dt_gamma_config():
    """trex-style options: cut box (with the reference's x-for-y typo), dt_gamma>0, bound 2 / two cascades."""
    from pienerf_b200.network import NeRFNetwork
    from pienerf_b200.synthetic import make_body, make_field, occupancy_bitfield, orbit_pose
    body = make_body("block64", dx=0.05, bound=2.0)
    field = make_field(bound=2.0, seed=1)
    bits = occupancy_bitfield(body["pos"], 0.03, bound=2.0)
    p_ori, p_def, F, dF = deformed_ip_state(body, amp=0.02)
    W = H = 32
    pose = orbit_pose(radius=4.0); intr = np.array([0.5 * H * 4.0 / 0.3 * 0.8] * 2 + [W // 2, H // 2])
    rays_o, rays_d = ro.get_rays(pose, intr, H, W)
    cb = [-0.05, 0.5, -0.5, 0.5, -0.5, 0.5]
    kw = dict(dt_gamma=1 / 128, max_steps=300, T_thresh=5e-2)
    want = ro.rund_cuda(ro.OracleField(field), rays_o, rays_d, p_def, p_ori, F, dF, 0.0525, bits, 2.0, 2, min_near=0.2, density_scale=20.0,
                        max_iter_num=1, hash_grid_size=0.06, num_seek_IP=1, cut=True, cut_bounds=cb, **kw)
    model = NeRFNetwork(bound=2, density_scale=20.0).cuda().load_field(field)
    assert model.cascade == 2
    model.density_bitfield.copy_(_gpu(bits))
    model.p_ori, model.p_def, model.IP_F, model.IP_dF, model.IP_dx = _gpu(p_ori), _gpu(p_def), _gpu(F), _gpu(dF), 0.0525
    opt = dict(max_iter_num=1, hash_grid_size=0.06, bound=2.0, cut=True, cut_bounds=cb, num_seek_IP=1)
    import functools
    for fn in (model.rund_cuda, functools.partial(model.render_deformed, mode=0), functools.partial(model.render_deformed, mode=3)):
        got = fn(_gpu(rays_o)[None], _gpu(rays_d)[None], **kw, **opt)["image"][0].cpu().numpy()
        assert _bad_fraction(got, want["image"], 1e-3) <= 0.02 and psnr(got, want["image"]) > 40


def test_full_size_properties():
    """800x800 chair-config frame (too big for numpy): size-independent properties of the fused renderer."""
    from pienerf_b200.frame import FrameDriver, build_scene
    model, sim, opt, pose, intr, body, field = build_scene("chair", density_scale=50.0)
    drv = FrameDriver(model, sim, opt, fused=True)
    out = drv.test_gui(pose, intr, opt.W, opt.H, paused=True, to_host=False)
    img = out["image"]; ws = model._workspace  # noqa: F841
    stats = model._stats.cpu().numpy()
    assert img.shape == (800, 800, 3) and torch.isfinite(img).all()
    assert 0 <= float(img.min()) and float(img.max()) <= 1 + 1e-5
    assert stats[1] > 10000 and stats[0] > stats[1]                                # rays hit, several samples each
    # idempotence: same state -> bit-identical frame (deterministic IP order, no atomics on the data path)
    again = drv.test_gui(pose, intr, opt.W, opt.H, paused=True, to_host=False)["image"]
    assert torch.equal(img, again)
    # background stays white, silhouette is centred
    assert float(img[0, 0].min()) == 1.0 and float(img[400, 400].max()) < 1.0
    # tile-sharded rendering == full-frame rendering (what the multi-GPU path relies on)
    rays, _, _ = drv.rays(pose, intr, opt.W, opt.H)
    sel = torch.arange(0, 640000, 7, device="cuda")
    part = model.render_deformed(rays["rays_o"][:, sel], rays["rays_d"][:, sel], **opt)["image"][0]
    assert torch.equal(part, img.reshape(-1, 3)[sel])


@pytest.mark.parametrize("amp,K,ds", [(0.0, 3, 20.0), (0.03, 3, 20.0), (0.03, 1, 1.0)])
def test_frame_vs_reference_gpu_renderer(amp, K, ds):
    """BASELINE.json's render bar, taken literally: RGB within 1e-3 absolute of the REFERENCE GPU renderer — the
    reference's own CUDA kernels (oracle/_ref, compiled unmodified) in its own rund_cuda loop with the fp32 nn.Linear
    MLP — on an undeformed frame (amp = 0) and on deformed ones, at 160x160 (too big for the numpy oracle)."""
    from oracle.build_ref import load_ref
    if load_ref("_ref_raymarching") is None or load_ref("_ref_gridencoder") is None or load_ref("_ref_shencoder") is None:
        pytest.skip("oracle/_ref not built")
    from oracle.ref_renderer import ReferenceRenderer
    model, field, bits, (p_ori, p_def, F, dF), rays_o, rays_d = _setup(amp, W=160, H=160, density_scale=ds)
    ref = ReferenceRenderer(field, bits, bound=1.0, density_scale=ds, min_near=0.2)
    kw = dict(dt_gamma=0.0, max_steps=512, T_thresh=1e-2)
    want = ref.rund_cuda(_gpu(rays_o), _gpu(rays_d), _gpu(p_def), _gpu(p_ori), _gpu(F), _gpu(dF), 0.0525, max_iter_num=1,
                         hash_grid_size=0.06, num_seek_IP=K, return_stats=True, **kw)
    opt = dict(max_iter_num=1, hash_grid_size=0.06, bound=1.0, cut=False, cut_bounds=[0.0] * 6, num_seek_IP=K)
    got = model.render_deformed(_gpu(rays_o)[None], _gpu(rays_d)[None], mode=3, **kw, **opt)
    img = got["image"][0]; wimg = want["image"]
    err = (img - wimg).abs().max(-1).values
    hit = want["weights_sum"] > 0
    assert int(hit.sum()) > 2000 and want["n_samples"] > 50000
    # kept samples: identical march decisions up to knife-edge occupancy flips (FMA contraction in the warp)
    assert abs(int(got["stats"][0]) - want["n_samples"]) <= 2e-3 * want["n_samples"] + 2
    assert float((err > 1e-3).float().mean()) <= 0.002, float((err > 1e-3).float().mean())
    assert float(err[hit].median()) < 2e-5
    assert float((got["weights_sum"] - want["weights_sum"]).abs().quantile(0.999)) < 1e-3
    d0 = got["depth_0"][0]
    assert float((d0 - want["depth_0"]).abs().quantile(0.998)) < 2e-3
