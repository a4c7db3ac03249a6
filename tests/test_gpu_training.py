"""Training-side CUDA kernels (march_rays_train, composite_rays_train fwd/bwd, grid backward / input backward /
total variation, SH Jacobian / backward) through the C-ABI against
  * tests/golden/ref_train.npz — outputs of the reference's own kernels (tests/golden/make_golden_train.py),
  * the numpy oracle (oracle/train_oracle.py) on other seeds,
  * the reference's kernels directly (oracle/_ref/*.so, when built) at sizes the oracle cannot reach,
  * size-independent properties (adjoint identity of the table gradient, autograd vs a float64 torch restatement).
Tolerances are stated per check; integer outputs (ray table, counters) are compared exactly."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

from oracle import render_oracle as ro  # noqa: E402
from oracle import train_oracle as to  # noqa: E402
from oracle.build_ref import load_ref  # noqa: E402
from pienerf_b200.synthetic import grid_offsets  # noqa: E402
from tests.util import repack_by_ray, small_scene  # noqa: E402

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_train.npz")
f32 = np.float32


def _gpu(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


@pytest.fixture(scope="module")
def G():
    if not os.path.exists(GOLD):
        pytest.skip("tests/golden/ref_train.npz not generated yet")
    return np.load(GOLD)


def _march(be, o, d, bits, bound, dt_gamma, max_steps, C, H, M, nears, fars, noises, counter0=(0, 0)):
    N = o.shape[0]
    xyzs = torch.zeros(M, 3, device="cuda"); dirs = torch.zeros(M, 3, device="cuda"); deltas = torch.zeros(M, 2, device="cuda")
    rays = torch.full((N, 3), -7, dtype=torch.int32, device="cuda")
    counter = torch.tensor(list(counter0), dtype=torch.int32, device="cuda")
    be.march_rays_train(_gpu(o), _gpu(d), _gpu(bits), float(bound), float(dt_gamma), int(max_steps), N, int(C), int(H), M, _gpu(nears), _gpu(fars),
                        xyzs, dirs, deltas, rays, counter, _gpu(noises))
    torch.cuda.synchronize()
    return xyzs.cpu().numpy(), dirs.cpu().numpy(), deltas.cpu().numpy(), rays.cpu().numpy(), counter.cpu().numpy()


@pytest.fixture(params=[0, 1], ids=["stream-stores", "cooperative-flush"])
def write_mode(request):
    """Both write passes of march_rays_train (pn_set_train_write_mode) must produce the same bytes."""
    from pienerf_b200._lib import lib
    was = lib.pn_set_train_write_mode(request.param)
    yield request.param
    lib.pn_set_train_write_mode(was)


@pytest.mark.parametrize("tag", ["mA", "mB"])
def test_march_rays_train_vs_golden(G, tag, write_mode):
    """Same compiler, same float expressions as the reference: identical ray table and bit-identical samples."""
    import pienerf_b200._raymarching as be
    bound, dt_gamma, max_steps, C, H = G[f"{tag}_par"]
    N = G[f"{tag}_o"].shape[0]
    x, d, dl, rays, counter = _march(be, G[f"{tag}_o"], G[f"{tag}_d"], G[f"{tag}_bits"], bound, dt_gamma, max_steps, C, H, N * 64,
                                     G[f"{tag}_nears"], G[f"{tag}_fars"], G[f"{tag}_noises"])
    R = G[f"{tag}_rays"]
    m = int(R[:, 2].sum())
    assert np.array_equal(rays, R)                                  # ray-ordered packing == the re-packed reference table
    assert counter.tolist() == [m, N]
    assert np.array_equal(x[:m], G[f"{tag}_xyzs"]) and np.array_equal(d[:m], G[f"{tag}_dirs"]) and np.array_equal(dl[:m], G[f"{tag}_deltas"])
    assert not x[m:].any() and not dl[m:].any()                     # nothing written past the samples produced


def test_march_rays_train_vs_oracle_other_seed_and_overflow(rng, write_mode):
    import pienerf_b200._raymarching as be
    body, field, bits, pose, intr = small_scene(W=20, H=28, seed=5)
    o, d = ro.get_rays(pose, intr, 28, 20)
    nears, fars = ro.near_far_from_aabb(o, d, np.array([-1, -1, -1, 1, 1, 1], f32), 0.2)
    noises = rng.uniform(0, 1, o.shape[0]).astype(f32)
    N = o.shape[0]
    args = (o, d, bits, 1.0, 1.0 / 256, 512, 1, 128)
    want = to.march_rays_train(*args, N * 64, nears, fars, noises)
    got = _march(be, *args, N * 64, nears, fars, noises)
    same = got[3][:, 2] == want[3][:, 2]
    assert same.mean() >= 0.99 and (want[3][:, 2] > 0).sum() > 50   # knife-edge occupancy flips under FMA contraction only
    if same.all():
        assert np.array_equal(got[3], want[3]) and np.abs(got[0] - want[0]).max() < 2e-6 and np.abs(got[2] - want[2]).max() < 2e-6
    total = int(got[4][0])
    # M too small: rays whose samples do not fit are dropped whole, the table and the counter still describe every ray,
    # and an incoming counter offsets the packing (raymarching.cu:407-419)
    M = total // 2
    x2, d2, dl2, rays2, counter2 = _march(be, *args, M, nears, fars, noises, counter0=(16, 3))
    assert np.array_equal(rays2[:, [0, 2]], got[3][:, [0, 2]]) and np.array_equal(rays2[:, 1], got[3][:, 1] + 16)
    assert counter2.tolist() == [total + 16, N + 3]
    fits = (rays2[:, 2] > 0) & (rays2[:, 1] + rays2[:, 2] <= M)
    assert fits.any() and (~fits & (rays2[:, 2] > 0)).any()
    for n in np.nonzero(fits)[0][:50]:
        a = slice(rays2[n, 1], rays2[n, 1] + rays2[n, 2]); b = slice(got[3][n, 1], got[3][n, 1] + got[3][n, 2])
        assert np.array_equal(x2[a], got[0][b]) and np.array_equal(dl2[a], got[2][b])
    assert not x2[:16].any()


@pytest.mark.parametrize("tag", ["mA", "mB"])
def test_composite_rays_train_vs_golden(G, tag):
    import pienerf_b200._raymarching as be
    R = _gpu(G[f"{tag}_rays"]); dl = _gpu(G[f"{tag}_deltas"]); sig = _gpu(G[f"{tag}_sig"]); rgb = _gpu(G[f"{tag}_rgb"])
    N = R.shape[0]; m = sig.shape[0]
    ws = torch.empty(N, device="cuda"); dep = torch.empty(N, device="cuda"); img = torch.empty(N, 3, device="cuda")
    be.composite_rays_train_forward(sig, rgb, dl, R, m, N, 1e-2, ws, dep, img)
    # same expressions, same compiler: allow 1 ulp-level differences from contraction choices only
    for got, key in ((ws, "ws"), (dep, "depth"), (img, "image")):
        assert np.abs(got.cpu().numpy() - G[f"{tag}_{key}"]).max() < 2e-6, key
    gs = torch.zeros(m, device="cuda"); gc = torch.zeros(m, 3, device="cuda")
    be.composite_rays_train_backward(_gpu(G[f"{tag}_gws"]), _gpu(G[f"{tag}_gim"]), sig, rgb, dl, R, _gpu(G[f"{tag}_ws"]), _gpu(G[f"{tag}_image"]),
                                     m, N, 1e-2, gs, gc)
    assert np.abs(gc.cpu().numpy() - G[f"{tag}_gc"]).max() < 1e-6
    assert np.abs(gs.cpu().numpy() - G[f"{tag}_gs"]).max() < 1e-5 * max(1.0, np.abs(G[f"{tag}_gs"]).max())
    assert ((gs.cpu().numpy() == 0) != (G[f"{tag}_gs"] == 0)).sum() <= 2       # the same samples are behind the termination point


def _grid_case(G, tag):
    S, H, D, C, gridtype, align, interp, L = G[f"{tag}_par"]
    return float(S), int(H), int(D), int(C), int(gridtype), bool(align), int(interp), int(L)


@pytest.mark.parametrize("tag", ["gA", "gB"])
def test_grid_training_kernels_vs_golden(G, tag):
    import pienerf_b200._gridencoder as ge
    S, H, D, C, gridtype, align, interp, L = _grid_case(G, tag)
    x = _gpu(G[f"{tag}_x"]); emb = _gpu(G[f"{tag}_emb"]); off = _gpu(G[f"{tag}_off"]); grad = _gpu(G[f"{tag}_grad"])
    B = x.shape[0]
    out = torch.empty(L, B, C, device="cuda"); dy_dx = torch.empty(B, L * D * C, device="cuda")
    ge.grid_encode_forward(x, emb, off, out, B, D, C, L, S, H, dy_dx, gridtype, align, interp)
    assert np.abs(dy_dx.cpu().numpy() - G[f"{tag}_dy_dx"]).max() < 1e-6 * max(1.0, np.abs(G[f"{tag}_dy_dx"]).max())
    gemb = torch.zeros_like(emb); gin = torch.full((B, D), 9.0, device="cuda")        # grad_inputs is overwritten, not accumulated
    ge.grid_encode_backward(grad, x, emb, off, gemb, B, D, C, L, S, H, dy_dx, gin, gridtype, align, interp)
    ref = G[f"{tag}_gemb"]
    assert np.abs(gemb.cpu().numpy() - ref).max() < 1e-5 * max(1.0, np.abs(ref).max())   # fp32 reductions, order differs
    assert np.array_equal(gemb.cpu().numpy() == 0, ref == 0)                             # the same entries are touched
    assert np.abs(gin.cpu().numpy() - G[f"{tag}_gin"]).max() < 1e-5 * max(1.0, np.abs(G[f"{tag}_gin"]).max())
    # accumulation into a non-zero grad_embeddings, and the no-Jacobian call
    g2 = torch.ones_like(emb)
    ge.grid_encode_backward(grad, x, emb, off, g2, B, D, C, L, S, H, None, None, gridtype, align, interp)
    assert np.abs(g2.cpu().numpy() - 1 - ref).max() < 2e-5 * max(1.0, np.abs(ref).max())
    tv = torch.zeros_like(emb)
    ge.grad_total_variation(x, emb, tv, off, 1e-2, B, D, C, L, S, H, gridtype, align)
    tr = G[f"{tag}_tv"]
    assert np.abs(tv.cpu().numpy() - tr).max() < 1e-5 * max(1.0, np.abs(tr).max())
    assert np.array_equal(tv.cpu().numpy() == 0, tr == 0)


@pytest.mark.parametrize("D,C,gridtype,align,interp,half", [(3, 2, 0, False, 0, True), (3, 8, 0, False, 0, False), (3, 1, 0, False, 0, False),
                                                            (4, 2, 1, False, 1, False), (2, 1, 0, True, 0, True)])
def test_grid_backward_variants_vs_oracle(rng, D, C, gridtype, align, interp, half):
    """Every vector-reduction path (f32 x1/x2/x4, f16 x1/x2) against the float64 oracle."""
    import pienerf_b200._gridencoder as ge
    L = 5
    off, s = grid_offsets(input_dim=D, num_levels=L, base_resolution=4, log2_hashmap_size=11, desired_resolution=48, align_corners=align)
    emb = rng.uniform(-1, 1, size=(int(off[-1]), C)).astype(f32)
    B = 3000
    x = rng.uniform(-0.01, 1.01, size=(B, D)).astype(f32)
    grad = (rng.normal(size=(L, B, C)) * (0.05 if half else 1.0)).astype(f32)
    dt = torch.float16 if half else torch.float32
    S = float(np.log2(s))
    e = _gpu(emb).to(dt); g = _gpu(grad).to(dt)
    out = torch.empty(L, B, C, device="cuda", dtype=dt); dy_dx = torch.empty(B, L * D * C, device="cuda", dtype=dt)
    ge.grid_encode_forward(_gpu(x), e, _gpu(off), out, B, D, C, L, S, 4, dy_dx, gridtype, align, interp)
    gemb = torch.zeros_like(e); gin = torch.zeros(B, D, device="cuda", dtype=dt)
    ge.grid_encode_backward(g, _gpu(x), e, _gpu(off), gemb, B, D, C, L, S, 4, dy_dx, gin, gridtype, align, interp)
    want = to.grid_encode_backward(g.float().cpu().numpy(), x, emb.shape[0], off, S, 4, gridtype, align, interp)
    got = gemb.float().cpu().numpy().astype(np.float64)
    tol = (2e-2 if half else 2e-5) * max(1.0, np.abs(want).max())                       # fp16 sums round at every reduction
    assert np.abs(got - want).max() < tol, np.abs(got - want).max()
    jr = dy_dx.float().cpu().numpy().reshape(B, L, D, C)
    gi = to.grid_input_backward(g.float().cpu().numpy(), jr)
    assert np.abs(gin.float().cpu().numpy() - gi).max() < (5e-2 if half else 1e-5) * max(1.0, np.abs(gi).max())
    tv = torch.zeros_like(e)
    ge.grad_total_variation(_gpu(x).to(dt), e, tv, _gpu(off), 0.5, B, D, C, L, S, 4, gridtype, align)
    wtv = to.grad_total_variation(_gpu(x).to(dt).float().cpu().numpy(), e.float().cpu().numpy(), off, 0.5, S, 4, gridtype, align)
    # fp16 tables accumulate the TV gradient in fp16 (as the reference's at::Half atomics do): ~200 roundings per coarse entry
    assert np.abs(tv.float().cpu().numpy() - wtv).max() < (5e-2 if half else 2e-5) * max(1.0, np.abs(wtv).max())


def test_grid_backward_full_size_properties_and_reference_kernel(rng):
    """Hot configuration (D=3, C=2, L=16, T=2^19) at 2^20 samples: the table gradient is the adjoint of the forward pass
    (<g, E(x) e> == <E(x)^T g, e>, checked in float64), and it agrees with the reference's own kernel to reduction-order rounding."""
    import pienerf_b200._gridencoder as ge
    off, s = grid_offsets(desired_resolution=2048)
    S = float(np.log2(s)); B = 1 << 20; L = 16
    emb = _gpu(rng.uniform(-1, 1, size=(int(off[-1]), 2)).astype(f32)); x = _gpu(rng.uniform(0, 1, size=(B, 3)).astype(f32)); o = _gpu(off)
    grad = torch.randn(L, B, 2, device="cuda")
    out = torch.empty(L, B, 2, device="cuda")
    ge.grid_encode_forward(x, emb, o, out, B, 3, 2, L, S, 16, None, 0, False, 0)
    gemb = torch.zeros_like(emb)
    ge.grid_encode_backward(grad, x, emb, o, gemb, B, 3, 2, L, S, 16, None, None, 0, False, 0)
    lhs = float((grad.double() * out.double()).sum()); rhs = float((gemb.double() * emb.double()).sum())
    assert abs(lhs - rhs) < 1e-4 * max(1.0, abs(lhs), float((grad.double() * out.double()).abs().sum()) * 1e-3), (lhs, rhs)
    rg = load_ref("_ref_gridencoder")
    if rg is None:
        pytest.skip("oracle/_ref not built")
    ref = torch.zeros_like(emb)
    rg.grid_encode_backward(grad, x, emb, o, ref, B, 3, 2, L, S, 16, None, None, 0, False, 0)
    torch.cuda.synchronize()
    scale = float(ref.abs().max())
    assert float((gemb - ref).abs().max()) < 1e-4 * scale, (float((gemb - ref).abs().max()), scale)   # coarse levels sum ~1e5 terms per entry
    tv = torch.zeros_like(emb); tvr = torch.zeros_like(emb)
    ge.grad_total_variation(x, emb, tv, o, 1e-3, B, 3, 2, L, S, 16, 0, False)
    rg.grad_total_variation(x, emb, tvr, o, 1e-3, B, 3, 2, L, S, 16, 0, False)
    assert float((tv - tvr).abs().max()) < 1e-4 * float(tvr.abs().max())


@pytest.mark.parametrize("deg", [4, 8])
def test_sh_jacobian_and_backward_vs_golden(G, deg):
    """The Jacobian comes from forward-mode differentiation of the forward polynomials, the reference tabulates closed forms:
    same values to fp32 rounding (tolerance scaled by the largest derivative of the band, ~deg^2)."""
    import pienerf_b200._shencoder as se
    dirs = _gpu(G["sh_dirs"]); B = dirs.shape[0]
    y = torch.empty(B, deg * deg, device="cuda"); j = torch.empty(B, 3 * deg * deg, device="cuda")
    se.sh_encode_forward(dirs, y, B, 3, deg, j)
    y0 = torch.empty_like(y)
    se.sh_encode_forward(dirs, y0, B, 3, deg, None)
    assert torch.equal(y, y0)                                        # outputs do not depend on whether the Jacobian is requested
    assert np.abs(y.cpu().numpy() - G[f"sh{deg}_y"]).max() < 2e-6
    jr = G[f"sh{deg}_dy_dx"]
    assert np.abs(j.cpu().numpy() - jr).max() < 5e-6 * max(1.0, np.abs(jr).max())
    gin = torch.zeros(B, 3, device="cuda")
    se.sh_encode_backward(_gpu(G[f"sh{deg}_grad"]), dirs, B, 3, deg, _gpu(jr), gin)
    assert np.abs(gin.cpu().numpy() - G[f"sh{deg}_gin"]).max() < 1e-5 * max(1.0, np.abs(G[f"sh{deg}_gin"]).max())
    se.sh_encode_backward(_gpu(G[f"sh{deg}_grad"]), dirs, B, 3, deg, _gpu(jr), gin)     # accumulates, as the reference does
    assert np.abs(gin.cpu().numpy() - 2 * G[f"sh{deg}_gin"]).max() < 2e-5 * max(1.0, np.abs(G[f"sh{deg}_gin"]).max())
    if deg == 4:
        assert np.abs(j.cpu().numpy().reshape(B, 3, 16) - to.sh_jacobian(G["sh_dirs"], 4)).max() < 2e-6


def test_autograd_wrappers_end_to_end(rng):
    """One training-shaped step through the host-side mirrors (march_rays_train -> GridEncoder / SHEncoder -> composite_rays_train
    -> loss.backward()): gradients agree with a float64 torch-autograd restatement of the compositing on the same samples, and
    with central differences for the encoders' input gradients."""
    from pienerf_b200 import raymarching as rm
    from pienerf_b200.gridencoder import GridEncoder
    from pienerf_b200.shencoder import SHEncoder
    body, field, bits, pose, intr = small_scene(W=16, H=16, seed=2)
    o, d = ro.get_rays(pose, intr, 16, 16)
    nears, fars = rm.near_far_from_aabb(_gpu(o), _gpu(d), _gpu(np.array([-1, -1, -1, 1, 1, 1], f32)), 0.2)
    counter = torch.zeros(2, dtype=torch.int32, device="cuda")
    xyzs, dirs, deltas, rays = rm.march_rays_train(_gpu(o), _gpu(d), 1.0, _gpu(bits), 1, 128, nears, fars, counter, -1, True, 128, True, 0.0, 256)
    m = int(counter[0])
    assert xyzs.shape[0] % 128 == 0 and xyzs.shape[0] >= m > 200 and rays.shape == (256, 3)
    torch.manual_seed(0)
    enc = GridEncoder(num_levels=8, desired_resolution=256, log2_hashmap_size=14).cuda()
    enc.embeddings.data.uniform_(-1, 1)
    she = SHEncoder(degree=4)
    w_sigma = torch.randn(16, device="cuda") * 0.5; w_rgb = torch.randn(16 + 16, 3, device="cuda") * 0.3
    x = xyzs.clone().requires_grad_(True); dv = dirs.clone().requires_grad_(True)
    feat = enc(x, bound=1.0)                                          # [M, 16]; inputs.requires_grad -> Jacobian path
    sh = she(dv)
    sigmas = torch.nn.functional.softplus(feat @ w_sigma) * 20
    rgbs = torch.sigmoid(torch.cat([feat, sh], -1) @ w_rgb)
    ws, depth, image = rm.composite_rays_train(sigmas, rgbs, deltas, rays, 1e-4)
    target = torch.rand(256, 3, device="cuda")
    loss = ((image - target) ** 2).sum() + 0.1 * (ws ** 2).sum()
    loss.backward()
    g_emb = enc.embeddings.grad.clone(); g_x = x.grad.clone(); g_d = dv.grad.clone()
    assert float(g_emb.abs().max()) > 0 and float(g_x.abs().max()) > 0 and float(g_d.abs().max()) > 0

    # float64 restatement of the compositing with torch autograd, on the same sigmas / rgbs
    s64 = sigmas.detach().double().requires_grad_(True); c64 = rgbs.detach().double().requires_grad_(True)
    img = torch.zeros(256, 3, dtype=torch.float64, device="cuda"); wsum = torch.zeros(256, dtype=torch.float64, device="cuda")
    rays_h = rays.cpu().numpy(); dl64 = deltas.double()
    img_rows, ws_rows = [], []
    for n, off, num in rays_h:
        if num == 0:
            img_rows.append(torch.zeros(3, dtype=torch.float64, device="cuda")); ws_rows.append(torch.zeros((), dtype=torch.float64, device="cuda")); continue
        sl = slice(int(off), int(off + num))
        alpha = 1 - torch.exp(-s64[sl] * dl64[sl, 0])
        T = torch.cumprod(torch.cat([torch.ones(1, dtype=torch.float64, device="cuda"), 1 - alpha[:-1]]), 0)
        keep = (torch.cat([torch.ones(1, dtype=torch.float64, device="cuda"), torch.cumprod(1 - alpha, 0)[:-1]]) >= 1e-4).double()   # T_thresh cut
        w = alpha * T * keep
        img_rows.append((w[:, None] * c64[sl]).sum(0)); ws_rows.append(w.sum())
    order = np.argsort(rays_h[:, 0])
    img = torch.stack([img_rows[k] for k in order]); wsum = torch.stack([ws_rows[k] for k in order])
    assert float((img - image.detach().double()).abs().max()) < 1e-5 and float((wsum - ws.detach().double()).abs().max()) < 1e-5
    loss64 = ((img - target.double()) ** 2).sum() + 0.1 * (wsum ** 2).sum()
    gs64, gc64 = torch.autograd.grad(loss64, [s64, c64])
    # our composite backward, isolated
    s32 = sigmas.detach().clone().requires_grad_(True); c32 = rgbs.detach().clone().requires_grad_(True)
    ws2, _, im2 = rm.composite_rays_train(s32, c32, deltas, rays, 1e-4)
    (((im2 - target) ** 2).sum() + 0.1 * (ws2 ** 2).sum()).backward()
    # the reference's sigma gradient drops the (tiny) contribution of samples behind the T_thresh cut: 1e-3 relative to the largest entry
    assert float((c32.grad.double() - gc64).abs().max()) < 1e-5 * max(1.0, float(gc64.abs().max()))
    assert float((s32.grad.double() - gs64).abs().max()) < 2e-3 * max(1.0, float(gs64.abs().max()))

    # encoders: input gradients vs central differences of the forward pass (float64 accumulation of a random projection)
    proj = torch.randn(m, 16, device="cuda", dtype=torch.float64); projs = torch.randn(m, 16, device="cuda", dtype=torch.float64)
    xs = xyzs[:m].clone().requires_grad_(True); ds = dirs[:m].clone().requires_grad_(True)
    (enc(xs, bound=1.0).double() * proj).sum().backward(); (she(ds).double() * projs).sum().backward()
    h, hg = 1e-3, 2e-4                                               # the grid's finest cells are 2/256 wide: keep the stencil inside one cell
    with torch.no_grad():
        for a in range(3):
            e = torch.zeros(3, device="cuda"); e[a] = hg
            fd = ((enc(xyzs[:m] + e, bound=1.0).double() - enc(xyzs[:m] - e, bound=1.0).double()) * proj).sum(1) / (2 * hg)
            e[a] = h
            err = (fd - xs.grad[:, a].double()).abs()
            assert float(err.median()) < 2e-2 * float(fd.abs().median() + 1e-3), (a, float(err.median()))   # piecewise-linear: exact except across cell faces
            fds = ((she(dirs[:m] + e).double() - she(dirs[:m] - e).double()) * projs).sum(1) / (2 * h)
            assert float((fds - ds.grad[:, a].double()).abs().max()) < 5e-3 * float(fds.abs().max())
    # total-variation gradient accumulates into embeddings.grad
    before = enc.embeddings.grad.clone()
    enc.grad_total_variation(weight=1e-3, inputs=xyzs[:m], bound=1.0)
    assert float((enc.embeddings.grad - before).abs().max()) > 0


def test_training_entry_points_validate_arguments():
    import pienerf_b200._gridencoder as ge
    import pienerf_b200._raymarching as be
    import pienerf_b200._shencoder as se
    x = torch.zeros(4, 3, device="cuda"); e = torch.zeros(8, 2, device="cuda"); o = torch.zeros(2, dtype=torch.int32, device="cuda")
    g = torch.zeros(1, 4, 2, device="cuda")
    with pytest.raises(RuntimeError, match="must be a CUDA tensor"):
        ge.grid_encode_backward(g.cpu(), x, e, o, torch.zeros_like(e), 4, 3, 2, 1, 1.0, 16, None, None, 0, False, 0)
    with pytest.raises(RuntimeError, match="dtype of embeddings"):
        ge.grid_encode_backward(g.half(), x, e, o, torch.zeros_like(e), 4, 3, 2, 1, 1.0, 16, None, None, 0, False, 0)
    with pytest.raises(NotImplementedError, match="double"):
        ge.grid_encode_backward(g.double(), x, e.double(), o, torch.zeros_like(e).double(), 4, 3, 2, 1, 1.0, 16, None, None, 0, False, 0)
    with pytest.raises(RuntimeError, match="go together"):
        ge.grid_encode_backward(g, x, e, o, torch.zeros_like(e), 4, 3, 2, 1, 1.0, 16, torch.zeros(4, 6, device="cuda"), None, 0, False, 0)
    with pytest.raises(RuntimeError, match="dtype of embeddings"):
        ge.grad_total_variation(x.half(), e, torch.zeros_like(e), o, 1.0, 4, 3, 2, 1, 1.0, 16, 0, False)
    with pytest.raises(RuntimeError, match="must be a CUDA tensor"):
        se.sh_encode_backward(torch.zeros(4, 16), x, 4, 3, 4, torch.zeros(4, 48, device="cuda"), torch.zeros(4, 3, device="cuda"))
    with pytest.raises(RuntimeError, match="must be an int tensor"):
        be.composite_rays_train_forward(torch.zeros(4, device="cuda"), x, torch.zeros(4, 2, device="cuda"), torch.zeros(2, 3, device="cuda"), 4, 2, 1e-4,
                                        torch.zeros(2, device="cuda"), torch.zeros(2, device="cuda"), torch.zeros(2, 3, device="cuda"))
    # empty inputs are a no-op
    be.composite_rays_train_forward(torch.zeros(0, device="cuda"), torch.zeros(0, 3, device="cuda"), torch.zeros(0, 2, device="cuda"),
                                    torch.zeros(0, 3, dtype=torch.int32, device="cuda"), 0, 0, 1e-4, torch.zeros(0, device="cuda"),
                                    torch.zeros(0, device="cuda"), torch.zeros(0, 3, device="cuda"))


@pytest.mark.parametrize("bound,dt_gamma", [(1.0, 0.0), (2.0, 1.0 / 128)])
def test_march_rays_train_full_frame_vs_reference_kernel(bound, dt_gamma):
    """800x800 rays through the chair-sized body against the reference's own kernel: every ray has the reference's sample count and
    its samples are bit-identical (positions, directions, deltas), for one cascade (where empty space is skipped block-wise, also
    compared with the skipping switched off) and for two.  The comparison gathers each ray's run from both packings on the GPU
    (the reference's offsets are in atomic order, ours in ray order)."""
    rr = load_ref("_ref_raymarching")
    if rr is None:
        pytest.skip("oracle/_ref not built")
    import pienerf_b200._raymarching as be
    from pienerf_b200 import raymarching as rm
    from pienerf_b200.synthetic import make_body, occupancy_bitfield, orbit_intrinsics, orbit_pose
    body = make_body("chair2k", dx=0.05, bound=1.0, seed=0)
    bits = _gpu(occupancy_bitfield(body["pos"], 0.03, bound=bound))
    C = 1 if bound <= 1 else 2
    W = H = 800
    rays = rm.get_rays(torch.from_numpy(orbit_pose(radius=2.5).astype(np.float32))[None], orbit_intrinsics(W, H, 50.0), H, W)
    o = rays["rays_o"][0].contiguous(); d = rays["rays_d"][0].contiguous(); N = o.shape[0]
    nears, fars = rm.near_far_from_aabb(o, d, torch.tensor([-bound] * 3 + [bound] * 3, device="cuda"), 0.2)
    noises = torch.rand(N, device="cuda", generator=torch.Generator(device="cuda").manual_seed(4))
    M = N * 24
    from pienerf_b200._lib import lib
    was_skip, was_mode = lib.pn_set_train_block_skip(1), lib.pn_set_train_write_mode(1)
    res = []
    for m, skip, mode in ((be, 1, 0), (rr, 1, 0), (be, 0, 0), (be, 1, 1), (be, 0, 1)):
        lib.pn_set_train_block_skip(skip); lib.pn_set_train_write_mode(mode)
        xyzs = torch.zeros(M, 3, device="cuda"); dirs = torch.zeros(M, 3, device="cuda"); deltas = torch.zeros(M, 2, device="cuda")
        rt = torch.empty(N, 3, dtype=torch.int32, device="cuda"); counter = torch.zeros(2, dtype=torch.int32, device="cuda")
        m.march_rays_train(o, d, bits, bound, dt_gamma, 1024, N, C, 128, M, nears, fars, xyzs, dirs, deltas, rt, counter, noises)
        res.append((xyzs, dirs, deltas, rt, counter))
    lib.pn_set_train_block_skip(was_skip); lib.pn_set_train_write_mode(was_mode)
    (x0, d0, l0, r0, c0), (x1, d1, l1, r1, c1) = res[0], res[1]
    for other in res[2:]:                                         # block skipping / the cooperative write pass change nothing, bit for bit
        for a, b in zip(res[0], other):
            assert torch.equal(a, b)
    assert torch.equal(c0, c1) and int(c0[0]) <= M and int(c0[0]) > 2_000_000, (c0, c1)
    r1s = r1[torch.argsort(r1[:, 0].long())]                      # the reference's rows, by ray
    assert torch.equal(r0[:, 0], r1s[:, 0]) and torch.equal(r0[:, 2], r1s[:, 2])
    num = r0[:, 2].long()
    ray_of = torch.repeat_interleave(torch.arange(N, device="cuda"), num)
    within = torch.arange(int(num.sum()), device="cuda") - torch.repeat_interleave(torch.cumsum(num, 0) - num, num)
    i0 = r0[:, 1].long()[ray_of] + within; i1 = r1s[:, 1].long()[ray_of] + within
    assert torch.equal(x0[i0], x1[i1]) and torch.equal(d0[i0], d1[i1]) and torch.equal(l0[i0], l1[i1])
