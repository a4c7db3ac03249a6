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
