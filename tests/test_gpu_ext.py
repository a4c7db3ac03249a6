"""The compiled pybind11 torch-extension modules (pienerf_b200/ext/*.so) on the GPU: same results as the ctypes modules over the
same C-ABI, and the reference's own unmodified wrappers (gridencoder/grid.py ...) run against them when the reference tree is
present (it is not on the GPU box; tests/test_cabi.py covers the import there is no GPU for)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


@pytest.fixture(scope="module")
def ext():
    from pienerf_b200.build_ext import load_ext
    mods = {n: load_ext(n) for n in ("_gridencoder", "_shencoder", "_raymarching", "_qgmls")}
    if any(m is None for m in mods.values()):
        pytest.skip("compiled modules not built (python -m pienerf_b200.build_ext)")
    return mods


def test_encoders_equal_ctypes_modules(ext):
    from pienerf_b200 import _gridencoder, _shencoder
    from pienerf_b200.synthetic import grid_offsets
    g = torch.Generator(device="cuda").manual_seed(0)
    offsets, pls = grid_offsets()
    emb = torch.rand(int(offsets[-1]), 2, device="cuda", generator=g) * 2 - 1
    off = torch.from_numpy(offsets).cuda()
    B = 50000
    x = torch.rand(B, 3, device="cuda", generator=g)
    a = torch.empty(16, B, 2, device="cuda"); b = torch.empty_like(a)
    args = (x, emb, off, None, B, 3, 2, 16, float(np.log2(pls)), 16, None, 0, False, 0)
    _gridencoder.grid_encode_forward(*(args[:3] + (a,) + args[4:]))
    ext["_gridencoder"].grid_encode_forward(*(args[:3] + (b,) + args[4:]))
    assert torch.equal(a, b) and float(a.abs().max()) > 0
    d = torch.nn.functional.normalize(torch.randn(B, 3, device="cuda", generator=g), dim=-1)
    sa = torch.empty(B, 16, device="cuda"); sb = torch.empty_like(sa)
    _shencoder.sh_encode_forward(d, sa, B, 3, 4, None)
    ext["_shencoder"].sh_encode_forward(d, sb, B, 3, 4, None)
    assert torch.equal(sa, sb)


def test_raymarching_equal_ctypes_modules(ext):
    from pienerf_b200 import _raymarching
    g = torch.Generator(device="cuda").manual_seed(1)
    N = 4096
    o = torch.zeros(N, 3, device="cuda"); o[:, 2] = -2.0
    d = torch.nn.functional.normalize(torch.randn(N, 3, device="cuda", generator=g) * 0.2 + torch.tensor([0.0, 0.0, 1.0], device="cuda"), dim=-1)
    aabb = torch.tensor([-1.0, -1, -1, 1, 1, 1], device="cuda")
    outs = []
    for m in (_raymarching, ext["_raymarching"]):
        nears = torch.empty(N, device="cuda"); fars = torch.empty(N, device="cuda")
        m.near_far_from_aabb(o, d, aabb, N, 0.2, nears, fars)
        bits = torch.full((128 ** 3 // 8,), 255, dtype=torch.uint8, device="cuda")
        alive = torch.arange(N, dtype=torch.int32, device="cuda"); t = nears.clone()
        n_step = 4
        xyzs = torch.zeros(N * n_step, 3, device="cuda"); dirs = torch.zeros_like(xyzs); deltas = torch.zeros(N * n_step, 2, device="cuda")
        m.march_rays(N, n_step, alive, t, o, d, 1.0, 0.0, 1024, 1, 128, bits, nears, fars, xyzs, dirs, deltas, torch.zeros(N, device="cuda"))
        sig = xyzs.abs().sum(-1) * 20; rgb = torch.sigmoid(xyzs)
        ws = torch.zeros(N, device="cuda"); dep = torch.zeros(N, device="cuda"); img = torch.zeros(N, 3, device="cuda")
        m.composite_rays(N, n_step, 1e-2, alive, t, sig, rgb, deltas, ws, dep, img)
        idx = torch.empty(N, dtype=torch.int32, device="cuda")
        m.morton3D(torch.randint(0, 128, (N, 3), device="cuda", dtype=torch.int32, generator=torch.Generator(device="cuda").manual_seed(2)), N, idx)
        outs.append((nears, fars, xyzs, deltas, ws, dep, img, alive, t, idx))
    for a, b in zip(*outs):
        assert torch.equal(torch.nan_to_num(a.float()), torch.nan_to_num(b.float()))
    assert float(outs[0][4].max()) > 0


def test_qgmls_step_through_the_compiled_module(ext):
    """One stepforward through the pybind11 `_qgmls.step` equals Simulator.stepforward (ctypes) bit for bit."""
    from pienerf_b200.simulator import Simulator
    from pienerf_b200.synthetic import make_body
    b = make_body("block512")
    s = Simulator(dt=1e-2, iters=10, bbox=torch.tensor([2.0, 2.0, 2.0]), dx=0.05, stiff=1e5, base=torch.tensor([-1.0, -1.0, -1.0]), use_graph=False)
    s.set_points(b["pos"], b["mass"], b["mu"], b["lam"], b["pin"]).initialize()
    s.stepforward()
    dof0, vel0 = s.dof.clone(), s.dof_vel.clone()
    s.stepforward()
    want = (s.dof.clone(), s.dof_vel.clone())
    dof, vel = dof0.clone(), vel0.clone()
    q = ext["_qgmls"]
    scratch = torch.empty(q.step_scratch_doubles(s.n_ip, s.n_k, s.adj_slices), dtype=torch.float64, device="cuda")
    q.step(10, 1e-2, 0.05, s.IP_kernel, s.IP_mu, s.IP_lam, s.IP_dNx, s.kernel_bg, s.buffer, s.adj_slices, s.global_matrix, s.mass_matrix_invt2, None, None,
           0, s.dof_rest, s.dof_f, s.rhs_rest, s.rhs_gravity, dof, vel, scratch, 0)
    assert torch.equal(dof, want[0]) and torch.equal(vel, want[1])
    pos = torch.empty(s.n_ip, 3, device="cuda"); F = torch.empty(s.n_ip, 9, device="cuda"); dF = torch.empty(s.n_ip, 27, device="cuda")
    q.ip_info(s.IP_kernel, dof, s.IP_Nx, s.IP_dNx, s.IP_ddNx, pos, F, dF)
    p2, F2, dF2 = s.get_IP_info()
    assert torch.equal(pos, p2) and torch.equal(F, F2) and torch.equal(dF, dF2)
    with pytest.raises(RuntimeError, match="float64"):
        q.matvec3(s.global_matrix.float(), dof, vel)


def test_dropin_prefers_compiled_modules_and_renders(ext):
    """install() -> sys.modules carries the pybind11 modules; the drop-in wavefront loop (NeRFNetwork.rund_cuda over the per-op
    entry points) renders the same frame through either binding."""
    import sys

    from pienerf_b200 import dropin
    dropin.install()
    assert set(dropin.installed.values()) == {"pybind11"}
    assert sys.modules["_raymarching"].__file__.endswith("ext/_raymarching.so")
    dropin.install(compiled=False)


def test_training_entry_points_equal_ctypes_modules(ext):
    """The compiled modules' training entry points (gridencoder/shencoder/raymarching bindings.cpp of the reference) reach the
    same C-ABI kernels as the ctypes modules: deterministic outputs are bit-equal, reduction outputs agree to reduction order."""
    from pienerf_b200 import _gridencoder, _raymarching, _shencoder
    from pienerf_b200.synthetic import grid_offsets
    g = torch.Generator(device="cuda").manual_seed(3)
    N = 2048
    o = torch.zeros(N, 3, device="cuda"); o[:, 2] = -2.0
    d = torch.nn.functional.normalize(torch.randn(N, 3, device="cuda", generator=g) * 0.2 + torch.tensor([0.0, 0.0, 1.0], device="cuda"), dim=-1)
    aabb = torch.tensor([-1.0, -1, -1, 1, 1, 1], device="cuda")
    nears = torch.empty(N, device="cuda"); fars = torch.empty(N, device="cuda")
    _raymarching.near_far_from_aabb(o, d, aabb, N, 0.2, nears, fars)
    bits = torch.randint(0, 256, (128 ** 3 // 8,), dtype=torch.uint8, device="cuda", generator=g)
    noises = torch.rand(N, device="cuda", generator=g)
    outs = []
    for m in (_raymarching, ext["_raymarching"]):
        M = N * 32
        xyzs = torch.zeros(M, 3, device="cuda"); dirs = torch.zeros(M, 3, device="cuda"); deltas = torch.zeros(M, 2, device="cuda")
        rays = torch.empty(N, 3, dtype=torch.int32, device="cuda"); counter = torch.zeros(2, dtype=torch.int32, device="cuda")
        m.march_rays_train(o, d, bits, 1.0, 0.0, 64, N, 1, 128, M, nears, fars, xyzs, dirs, deltas, rays, counter, noises)
        sig = xyzs.abs().sum(-1) * 20; rgb = torch.sigmoid(xyzs)
        ws = torch.empty(N, device="cuda"); dep = torch.empty(N, device="cuda"); img = torch.empty(N, 3, device="cuda")
        m.composite_rays_train_forward(sig, rgb, deltas, rays, M, N, 1e-3, ws, dep, img)
        gs = torch.zeros(M, device="cuda"); gc = torch.zeros(M, 3, device="cuda")
        m.composite_rays_train_backward(torch.ones(N, device="cuda"), torch.ones(N, 3, device="cuda"), sig, rgb, deltas, rays, ws, img, M, N, 1e-3, gs, gc)
        outs.append((xyzs, dirs, deltas, rays, counter, ws, dep, img, gs, gc))
    for a, b in zip(*outs):
        assert torch.equal(a, b)
    assert int(outs[0][4][0]) > N and int(outs[0][4][1]) == N and float(outs[0][8].abs().max()) > 0

    offsets, pls = grid_offsets(num_levels=8, desired_resolution=256, log2_hashmap_size=15)
    emb = torch.rand(int(offsets[-1]), 2, device="cuda", generator=g) * 2 - 1
    off = torch.from_numpy(offsets).cuda()
    B = 20000; S = float(np.log2(pls))
    x = torch.rand(B, 3, device="cuda", generator=g); grad = torch.randn(8, B, 2, device="cuda", generator=g)
    out = torch.empty(8, B, 2, device="cuda"); dy_dx = torch.empty(B, 8 * 3 * 2, device="cuda")
    _gridencoder.grid_encode_forward(x, emb, off, out, B, 3, 2, 8, S, 16, dy_dx, 0, False, 0)
    res = []
    for m in (_gridencoder, ext["_gridencoder"]):
        ge = torch.zeros_like(emb); gi = torch.zeros(B, 3, device="cuda"); tv = torch.zeros_like(emb)
        m.grid_encode_backward(grad, x, emb, off, ge, B, 3, 2, 8, S, 16, dy_dx, gi, 0, False, 0)
        m.grad_total_variation(x, emb, tv, off, 1e-2, B, 3, 2, 8, S, 16, 0, False)
        res.append((ge, gi, tv))
    assert torch.equal(res[0][1], res[1][1])
    assert float((res[0][0] - res[1][0]).abs().max()) < 1e-5 * float(res[0][0].abs().max())
    assert float((res[0][2] - res[1][2]).abs().max()) < 1e-5 * float(res[0][2].abs().max())
    dn = torch.nn.functional.normalize(torch.randn(B, 3, device="cuda", generator=g), dim=-1); gsh = torch.randn(B, 36, device="cuda", generator=g)
    sh = []
    for m in (_shencoder, ext["_shencoder"]):
        y = torch.empty(B, 36, device="cuda"); j = torch.empty(B, 3 * 36, device="cuda"); gi = torch.zeros(B, 3, device="cuda")
        m.sh_encode_forward(dn, y, B, 3, 6, j)
        m.sh_encode_backward(gsh, dn, B, 3, 6, j, gi)
        sh.append((y, j, gi))
    for a, b in zip(*sh):
        assert torch.equal(a, b)
