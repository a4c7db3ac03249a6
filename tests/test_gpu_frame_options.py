"""Frame options inside a16's line ranges and the two robustness paths ADVICE r1 asked for:
perturb (renderer.py:863, trainer.py:565), downscaled render + nearest upsample (trainer.py:537-539,577-586), a full sample
list (rays resume in the next pass; a list too small for the passes is REPORTED, not silently truncated), and an IP bounding
box beyond the scene (diverged body): flagged, no out-of-bounds access."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

from tests.test_gpu_render import _gpu, _setup  # noqa: E402

OPT = dict(max_iter_num=1, hash_grid_size=0.06, bound=1.0, cut=False, cut_bounds=[0.0] * 6, num_seek_IP=3)
KW = dict(dt_gamma=0.0, max_steps=512, T_thresh=1e-2)


def test_perturb_matches_reference_gpu_renderer():
    from oracle.build_ref import load_ref
    if any(load_ref(n) is None for n in ("_ref_raymarching", "_ref_gridencoder", "_ref_shencoder")):
        pytest.skip("oracle/_ref not built")
    from oracle.ref_renderer import ReferenceRenderer
    model, field, bits, (p_ori, p_def, F, dF), rays_o, rays_d = _setup(0.03, W=160, H=160, density_scale=20.0)
    ref = ReferenceRenderer(field, bits, bound=1.0, density_scale=20.0, min_near=0.2)
    noises = torch.rand(rays_o.shape[0], device="cuda", generator=torch.Generator(device="cuda").manual_seed(7))
    want = ref.rund_cuda(_gpu(rays_o), _gpu(rays_d), _gpu(p_def), _gpu(p_ori), _gpu(F), _gpu(dF), 0.0525, max_iter_num=1, hash_grid_size=0.06,
                         num_seek_IP=3, return_stats=True, first_noises=noises, **KW)
    plain = ref.rund_cuda(_gpu(rays_o), _gpu(rays_d), _gpu(p_def), _gpu(p_ori), _gpu(F), _gpu(dF), 0.0525, max_iter_num=1, hash_grid_size=0.06,
                          num_seek_IP=3, **KW)
    got = model.render_deformed(_gpu(rays_o)[None], _gpu(rays_d)[None], mode=3, noises=noises, **KW, **OPT)
    err = (got["image"][0] - want["image"]).abs().max(-1).values
    assert float((err > 1e-3).float().mean()) <= 0.003, float((err > 1e-3).float().mean())
    assert abs(int(got["stats"][0]) - want["n_samples"]) <= 3e-3 * want["n_samples"] + 2
    # the perturbation is visible (otherwise the comparison above proves nothing) ...
    assert float((want["image"] - plain["image"]).abs().max()) > 1e-3
    # ... and perturb=True draws its own noise
    rnd = model.render_deformed(_gpu(rays_o)[None], _gpu(rays_d)[None], mode=3, perturb=True, **KW, **OPT)
    assert torch.isfinite(rnd["image"]).all() and not torch.equal(rnd["image"], got["image"])


def test_downscaled_frame_is_nearest_upsampled():
    from pienerf_b200.frame import FrameDriver, Options
    from pienerf_b200.network import NeRFNetwork
    from pienerf_b200.simulator import Simulator
    from tests.util import small_scene
    W = H = 96
    body, field, bits, pose, intr = small_scene(kind="block64", W=W, H=H)
    model = NeRFNetwork(bound=1, density_scale=20.0).cuda().load_field(field)
    model.density_bitfield.copy_(torch.from_numpy(bits).cuda())
    sim = Simulator(dt=1e-2, iters=10, bbox=torch.tensor([2.0, 2.0, 2.0]), dx=0.05, stiff=1e5, base=torch.tensor([-1.0, -1.0, -1.0]))
    sim.set_points(body["pos"], body["mass"], body["mu"], body["lam"], body["pin"]).initialize()
    opt = Options.defaults(bound=1.0, W=W, H=H, max_steps=256, T_thresh=1e-2, dt_gamma=0.0, min_near=0.2, max_iter_num=1, num_seek_IP=3, sim_dx=0.05)
    drv = FrameDriver(model, sim, opt)
    half = drv.test_gui(pose, intr, W // 2, H // 2, paused=True, to_host=False)               # the frame the downscaled render must equal ...
    low = drv.test_gui(pose, np.asarray(intr) * 0.5, W // 2, H // 2, paused=True, to_host=False)
    up = drv.test_gui(pose, intr, W, H, paused=True, to_host=False, downscale=0.5)            # ... after intrinsics * 0.5 and nearest upsampling
    assert up["image"].shape == (H, W, 3) and up["depth_0"].shape == (H, W)
    assert torch.equal(up["image"][::2, ::2], low["image"]) and torch.equal(up["image"][1::2, 1::2], low["image"])
    assert torch.equal(up["depth"][::2, 1::2].nan_to_num(-1), low["depth"].nan_to_num(-1))
    assert not torch.equal(low["image"], half["image"])                                        # scaled intrinsics matter


def test_full_sample_list_defers_rays_and_reports_truncation():
    from pienerf_b200._lib import check, lib
    model, field, bits, state, rays_o, rays_d = _setup(0.03, W=64, H=64, density_scale=0.05)   # thin fog: long rays, many samples
    ro_, rd_ = _gpu(rays_o)[None], _gpu(rays_d)[None]
    base = {k: (v.clone() if torch.is_tensor(v) else v) for k, v in model.render_deformed(ro_, rd_, mode=3, **KW, **OPT).items()}
    st0 = model.check_stats(base["stats"])
    assert st0[6] == 0 and st0[0] > 5000
    try:
        # (a) a list that fills up in the first passes: chunks are deferred, the rays resume later, same frame bit for bit
        check(lib.pn_set_wave_capacity(max(1024, int(st0[0] * 0.6) // 256 * 256)))
        model._workspace = None
        small = model.render_deformed(ro_, rd_, mode=3, **KW, **OPT)
        st1 = model.check_stats(small["stats"])
        assert st1[6] > 0 and st1[5] == 0 and st1[0] == st0[0]
        assert torch.equal(small["image"], base["image"]) and torch.equal(small["depth_0"], base["depth_0"])
        # (b) far too small for the available passes: rays are cut short, and that is reported
        check(lib.pn_set_wave_capacity(1024))
        model._workspace = None
        tiny = model.render_deformed(ro_, rd_, mode=3, **KW, **OPT)
        st2 = [int(v) for v in tiny["stats"].tolist()]
        assert st2[5] > 0 and (st2[4] & 2)
        with pytest.raises(RuntimeError, match="cut short"):
            model.check_stats(tiny["stats"])
    finally:
        check(lib.pn_set_wave_capacity(0))
        model._workspace = None


def test_ip_bbox_beyond_the_scene_is_flagged_not_fatal():
    model, field, bits, (p_ori, p_def, F, dF), rays_o, rays_d = _setup(0.0, W=48, H=48)
    ro_, rd_ = _gpu(rays_o)[None], _gpu(rays_d)[None]
    for bad in (1.0e6, float("inf"), float("nan")):
        p = _gpu(p_def).clone()
        p[3, 1] = bad                                                       # one IP of a diverged simulation
        out = model.render_deformed(ro_, rd_, mode=3, ip_state=(p, _gpu(p_ori), _gpu(F), _gpu(dF)), **KW, **OPT)
        torch.cuda.synchronize()                                            # no illegal address
        st = [int(v) for v in out["stats"].tolist()]
        if bad == bad:                                                      # inf / huge: the clamp triggers; a NaN position is ignored by fmin/fmax
            assert st[4] & 1
            with pytest.raises(RuntimeError, match="bounding box"):
                model.check_stats(out["stats"])
    good = model.render_deformed(ro_, rd_, mode=3, **KW, **OPT)
    assert model.check_stats(good["stats"])[4] == 0


def test_non_default_architecture_is_refused_by_the_fused_kernels():
    from pienerf_b200.network import NeRFNetwork
    m = NeRFNetwork(bound=1, hidden_dim=32).cuda()
    x = torch.zeros(8, 3, device="cuda"); d = torch.zeros(8, 3, device="cuda"); d[:, 2] = 1
    with pytest.raises(NotImplementedError):
        m.forward_fused(x, d)
    s, c = m.forward(x, d)                                                   # the per-op path handles any shape
    assert s.shape == (8,) and c.shape == (8, 3)


@torch.no_grad()
def test_mlp_only_pass_matches_torch_fp32():
    """pn_mlp_forward (features in, sigma / rgb out; the kernel bench.py's `mlp_pass` times) vs the fp32 torch layers of network.py:105-127.
    Tolerance: the bf16x3 split keeps fp32-level accuracy (2e-4 relative on sigma, 1e-4 absolute on rgb, as for the fused field)."""
    import torch.nn.functional as F
    from pienerf_b200.network import NeRFNetwork
    from pienerf_b200.synthetic import make_field
    torch.backends.cuda.matmul.allow_tf32 = False
    model = NeRFNetwork(bound=1).cuda().load_field(make_field())
    g = torch.Generator(device="cuda").manual_seed(3)
    M = 100003
    enc = torch.randn(M, 32, device="cuda", generator=g) * 0.5
    d = F.normalize(torch.randn(M, 3, device="cuda", generator=g), dim=-1)
    sig, rgb = model.mlp_only(enc, d)
    h = F.relu(F.linear(enc.double(), model.sigma_net[0].weight.double()))
    h = F.linear(h, model.sigma_net[1].weight.double())
    want_sig = torch.exp(h[:, 0])
    sh = model.encoder_dir(d).double()
    c = torch.cat([sh, h[:, 1:]], dim=-1)
    c = F.relu(F.linear(c, model.color_net[0].weight.double())); c = F.relu(F.linear(c, model.color_net[1].weight.double()))
    want_rgb = torch.sigmoid(F.linear(c, model.color_net[2].weight.double()))
    assert float(((sig.double() - want_sig).abs() / want_sig.abs().clamp_min(1e-6)).max()) < 2e-4
    assert float((rgb.double() - want_rgb).abs().max()) < 1e-4
    # many tiles per CTA with producers that outrun the consumers: the stage ring must not let a producer group lap another
    # (regression: 2^21 rows hung before the consumers' software count guarded the parity-only mbarrier waits)
    big = enc.repeat(21, 1)[: 1 << 21].contiguous(); dbig = d.repeat(21, 1)[: 1 << 21].contiguous()
    sb, cb = model.mlp_only(big, dbig)
    torch.cuda.synchronize()
    assert torch.equal(sb[:M], sig) and torch.equal(cb[:M], rgb)


@pytest.mark.parametrize("K", [1, 3])
def test_fused_ip_preparation_kernel_equals_the_multi_kernel_chain(K):
    """bbox + counting sort + IP records + neighbourhood lists as one single-CTA kernel vs the ten stand-alone kernels: same frame bit for bit."""
    from pienerf_b200._lib import check, lib
    model, field, bits, state, rays_o, rays_d = _setup(0.03, W=96, H=96)
    ro_, rd_ = _gpu(rays_o)[None], _gpu(rays_d)[None]
    opt = dict(OPT, num_seek_IP=K)
    try:
        check(lib.pn_set_prep_mode(1))
        a = {k: (v.clone() if torch.is_tensor(v) else v) for k, v in model.render_deformed(ro_, rd_, mode=3, **KW, **opt).items()}
        check(lib.pn_set_prep_mode(0))
        b = model.render_deformed(ro_, rd_, mode=3, **KW, **opt)
        for name in ("image", "depth_0", "weights_sum"):
            assert torch.equal(a[name], b[name]), name
        assert a["stats"].tolist()[:4] == b["stats"].tolist()[:4] and int(a["stats"][0]) > 1000
        c = model.render_deformed(ro_, rd_, mode=0, **KW, **opt)                # the fused-frame kernel reads the same prepared structures
        assert torch.equal(c["image"], b["image"])
    finally:
        check(lib.pn_set_prep_mode(0))
