"""Whole frames at the sizes BASELINE.json names, against the REFERENCE GPU renderer (its own CUDA kernels compiled
unmodified into oracle/_ref, its rund_cuda loop, fp32 nn.Linear MLP — oracle/ref_renderer.py):

    chair      800 x 800,   2028 IPs, num_seek_IP 3, max_steps 1024, T_thresh 1e-2          (README.md:123)
    trex      1008 x 756,   8000 IPs, bound 2, --cut, dt_gamma 1/128, max_steps 300, T 5e-2  (README.md:134)
    synth1080 1920 x 1080,  4096 IPs                                                         (BASELINE.json configs[3])

each with density_scale 1 (thin fog: no early termination) and 50 (opaque, like a trained model), undeformed (0 steps) and
after 20 simulator steps with a drag force.  Bar (BASELINE.json): RGB within 1e-3 absolute.  What can differ is a sample
sitting on an occupancy-cell or IP-cell boundary (FMA contraction differs between the two builds); the tests print the
exact count of pixels off by more than 1e-3 and the kept-sample delta, and bound them.  The simulator sequences at the
same sizes are checked against the oracle in the second half of the file."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


def _ref_available():
    from oracle.build_ref import load_ref
    return all(load_ref(n) is not None for n in ("_ref_raymarching", "_ref_gridencoder", "_ref_shencoder"))


REPORT = []


@pytest.mark.parametrize("config", ["chair", "trex", "synth1080"])
@pytest.mark.parametrize("ds", [1.0, 50.0])
def test_full_size_frame_vs_reference_gpu_renderer(config, ds):
    if not _ref_available():
        pytest.skip("oracle/_ref not built")
    from oracle.ref_renderer import ReferenceRenderer
    from pienerf_b200 import raymarching
    from pienerf_b200.frame import build_scene
    model, sim, opt, pose, intr, body, field = build_scene(config, density_scale=ds)
    bits = model.density_bitfield.cpu().numpy()
    ref = ReferenceRenderer(field, bits, bound=float(opt.bound), density_scale=ds, min_near=float(opt.min_near))
    p_ori = sim.get_IP_info()[0].clone()
    model.p_ori, model.IP_dx = p_ori, sim.dx * 1.05
    rays = raymarching.get_rays(torch.from_numpy(pose).unsqueeze(0), np.asarray(intr), opt.H, opt.W, -1)
    ro, rd = rays["rays_o"][0].contiguous(), rays["rays_d"][0].contiguous()
    kw = dict(dt_gamma=float(opt.dt_gamma), max_steps=int(opt.max_steps), T_thresh=float(opt.T_thresh))
    for steps in (0, 20):
        if steps:
            sim.update_force(sim.n_ip // 3, torch.tensor([3e4, -1e4, 2e4]))
            for _ in range(steps):
                sim.stepforward()
        pos, F, dF = sim.get_IP_info()
        want = ref.rund_cuda(ro, rd, pos, p_ori, F, dF, sim.dx * 1.05, max_iter_num=int(opt.max_iter_num), hash_grid_size=float(opt.hash_grid_size),
                             cut=bool(opt.cut), cut_bounds=tuple(opt.cut_bounds), num_seek_IP=int(opt.num_seek_IP), return_stats=True, **kw)
        got = model.render_deformed(ro[None], rd[None], mode=3, ip_state=(pos, p_ori, F, dF), **opt)
        st = model.check_stats(got["stats"])                                  # raises on IP-grid overflow / truncated rays
        assert st[5] == 0
        err = (got["image"][0] - want["image"]).abs().max(-1).values
        n_bad = int((err > 1e-3).sum()); n_pix = err.numel()
        hit = want["weights_sum"] > 0
        d_samples = st[0] - want["n_samples"]
        REPORT.append(f"{config:9s} ds {ds:4.0f} steps {steps:2d}: pixels>1e-3 {n_bad:6d} / {n_pix} ({100.0 * n_bad / n_pix:.4f} %), max err {float(err.max()):.3e}, "
                      f"median err on hit pixels {float(err[hit].median()):.2e}, kept samples {st[0]} vs reference {want['n_samples']} (delta {d_samples:+d})")
        print(REPORT[-1])
        assert int(hit.sum()) > 0.02 * n_pix and want["n_samples"] > 100000
        if steps:
            assert float((pos - p_ori).abs().max()) > 1e-3                    # the body really is deformed
        assert n_bad <= 0.003 * n_pix, REPORT[-1]
        assert float(err[hit].median()) < 5e-5
        if ds == 1.0:                                                         # no early termination: both sides count exactly the marched samples
            assert abs(d_samples) <= 2e-3 * want["n_samples"] + 2
        else:                                                                 # the reference counts samples marched past a termination inside an n_step batch
            assert 0 < st[0] <= want["n_samples"]
        d0 = (got["depth_0"][0] - want["depth_0"]).abs()
        assert float(d0.quantile(0.995)) < 3e-3
    sim.clear_force()


@pytest.mark.parametrize("kind,bound", [("chair2k", 1.0), ("block4k", 1.0), ("block8k", 2.0), ("chairlike", 1.0)])
def test_full_size_step_sequence_vs_oracle(kind, bound):
    """30-step sequences (force on at step 3, off at step 6) at the IP counts of the BASELINE configs vs the fp64 oracle
    (itself pinned to the reference source by tests/test_sim_golden.py).  Bar: 1e-4 relative on IP positions / velocities."""
    from oracle.sim_oracle import OracleSimulator
    from pienerf_b200.simulator import Simulator
    from pienerf_b200.synthetic import make_body
    b = make_body(kind, bound=bound)
    o = OracleSimulator(dt=1e-2, iters=10, bbox=[2 * bound] * 3, dx=0.05, stiff=1e5, base=[-bound] * 3)
    o.initialize(b["pos"], b["mass"], b["mu"], b["lam"], b["pin"])
    s = Simulator(dt=1e-2, iters=10, bbox=torch.tensor([2.0 * bound] * 3), dx=0.05, stiff=1e5, base=torch.tensor([-bound] * 3))
    s.set_points(b["pos"], b["mass"], b["mu"], b["lam"], b["pin"]).initialize()
    assert (s.n_ip, s.n_k) == (o.n_ip, o.n_k)
    rel = lambda a, c: float(np.abs(a - c).max() / max(np.abs(c).max(), 1e-300))
    wp = wv = 0.0
    vid = s.n_ip // 3
    for i in range(30):
        if i == 3:
            s.update_force(vid, torch.tensor([2e4, 0.0, -1e4])); o.update_force(vid, [2e4, 0.0, -1e4])
        if i == 6:
            s.clear_force(); o.clear_force()
        s.stepforward(); o.stepforward()
        if i % 5 == 4 or i < 8:
            pos = s.get_IP_info()[0].cpu().numpy()
            wp = max(wp, rel(pos, o.get_IP_info()[0]))
            wv = max(wv, rel(s.dof_vel.cpu().numpy().reshape(-1), o.array("dof_vel")))
    print(f"{kind}: n_ip {s.n_ip} n_k {s.n_k}: worst rel. IP position error {wp:.2e}, DOF velocity error {wv:.2e} over 30 steps ({s.step_launches} launches / step)")
    assert wp < 1e-4 and wv < 1e-4                                            # BASELINE.json bar
    assert wp < 1e-6
