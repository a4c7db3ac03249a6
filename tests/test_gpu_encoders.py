"""CUDA hash-grid / SH / fused-field kernels vs the numpy oracle (and vs the reference's own kernels when
oracle/_ref/*.so were built), through the C-ABI.  Tolerances are stated per test."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

from oracle import render_oracle as ro  # noqa: E402
from oracle.build_ref import load_ref  # noqa: E402
from pienerf_b200.synthetic import grid_offsets, make_field  # noqa: E402


def _gpu(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def _inputs(rng, B):
    x = rng.uniform(0, 1, size=(B, 3)).astype(np.float32)
    x[0] = [0, 0, 0]; x[1] = [1, 1, 1]; x[2] = [1.0001, 0.5, 0.5]; x[3] = [0.5, -1e-6, 0.5]; x[4] = [0.5, 0.5, 0.5]
    return x


def test_grid_hot_path_vs_oracle(rng):
    """D=3, C=2, L=16 fp32 (the roofline kernel).  fp32 with FMA contraction vs numpy without: the finest
    level's position x*scale+0.5 can differ by 1 ulp at ~2048 (1.2e-4), so tolerance scales with resolution."""
    from pienerf_b200.gridencoder import grid_encode
    off, s = grid_offsets(desired_resolution=2048)
    emb = rng.uniform(-1, 1, size=(int(off[-1]), 2)).astype(np.float32)
    x = _inputs(rng, 4096)
    got = grid_encode(_gpu(x), _gpu(emb), _gpu(off), s, 16, level_major=True).cpu().numpy()
    want = ro.grid_encode(x, emb, off, np.log2(s), 16)
    assert got.shape == (16, 4096, 2)
    assert (got[:, 2] == 0).all() and (got[:, 3] == 0).all()               # out-of-range rows are exactly zero
    for l in range(16):
        scale = 16 * s ** l
        tol = 2e-6 + 6e-7 * scale                                          # ~ a few ulp(pos) * |d feature / d pos|
        assert np.abs(got[l] - want[l]).max() < tol, (l, np.abs(got[l] - want[l]).max())
    # [B, L*C] layout of grid.py:57
    flat = grid_encode(_gpu(x), _gpu(emb), _gpu(off), s, 16).cpu().numpy()
    assert np.array_equal(flat, got.transpose(1, 0, 2).reshape(4096, 32))


@pytest.mark.parametrize("D,C,gridtype,align,interp,half", [
    (3, 2, 0, False, 0, True), (3, 4, 0, False, 0, False), (3, 8, 1, False, 0, False), (2, 2, 0, True, 0, False),
    (3, 1, 0, False, 1, False), (2, 4, 1, True, 1, False)])
def test_grid_generic_variants(rng, D, C, gridtype, align, interp, half):
    import pienerf_b200._gridencoder as ge
    off, s = grid_offsets(input_dim=D, num_levels=8, desired_resolution=512, align_corners=align, log2_hashmap_size=15)
    emb = rng.uniform(-1, 1, size=(int(off[-1]), C)).astype(np.float32)
    x = rng.uniform(0, 1, size=(1000, D)).astype(np.float32)
    e = _gpu(emb).half() if half else _gpu(emb)
    out = torch.empty(8, 1000, C, device="cuda", dtype=e.dtype)
    ge.grid_encode_forward(_gpu(x), e, _gpu(off), out, 1000, D, C, 8, float(np.log2(s)), 16, None, gridtype, align, interp)
    want = ro.grid_encode(x, e.float().cpu().numpy(), off, np.log2(s), 16, gridtype, align, interp)
    tol = 5e-3 if half else 3e-4
    assert np.abs(out.float().cpu().numpy() - want).max() < tol


def test_grid_dy_dx_matches_finite_difference(rng):
    import pienerf_b200._gridencoder as ge
    off, s = grid_offsets(num_levels=4, desired_resolution=64)
    emb = rng.uniform(-1, 1, size=(int(off[-1]), 2)).astype(np.float32)
    x = rng.uniform(0.1, 0.9, size=(200, 3)).astype(np.float32)
    out = torch.empty(4, 200, 2, device="cuda"); dy = torch.empty(200, 4 * 3 * 2, device="cuda")
    ge.grid_encode_forward(_gpu(x), _gpu(emb), _gpu(off), out, 200, 3, 2, 4, float(np.log2(s)), 16, dy, 0, False, 0)
    dy = dy.cpu().numpy().reshape(200, 4, 3, 2)
    h = 1e-3
    for d in range(3):
        xp = x.copy(); xm = x.copy(); xp[:, d] += h; xm[:, d] -= h
        fd = (ro.grid_encode(xp, emb, off, np.log2(s), 16).astype(np.float64) - ro.grid_encode(xm, emb, off, np.log2(s), 16)) / (2 * h)
        # piecewise-linear: exact except where the +-h stencil crosses a cell face
        err = np.abs(dy[:, :, d, :].transpose(1, 0, 2) - fd)
        assert np.median(err) < 5e-3


def test_grid_error_behaviour():
    import pienerf_b200._gridencoder as ge
    x = torch.zeros(4, 3); e = torch.zeros(8, 2, device="cuda"); o = torch.zeros(2, dtype=torch.int32, device="cuda")
    out = torch.zeros(1, 4, 2, device="cuda")
    with pytest.raises(RuntimeError, match="must be a CUDA tensor"):
        ge.grid_encode_forward(x, e, o, out, 4, 3, 2, 1, 1.0, 16, None, 0, False, 0)
    with pytest.raises(RuntimeError, match="must be an int tensor"):
        ge.grid_encode_forward(x.cuda(), e, o.float(), out, 4, 3, 2, 1, 1.0, 16, None, 0, False, 0)
    with pytest.raises(RuntimeError, match="C must be 1, 2, 4, or 8"):
        ge.grid_encode_forward(x.cuda(), torch.zeros(8, 3, device="cuda"), o, torch.zeros(1, 4, 3, device="cuda"), 4, 3, 3, 1, 1.0, 16, None, 0, False, 0)
    with pytest.raises(TypeError):
        ge.grid_encode_backward()                       # a real entry point (tests/test_gpu_training.py): positional signature


def test_sh_vs_oracle_and_addition_theorem(rng):
    from pienerf_b200.shencoder import SHEncoder, sh_encode
    d = rng.normal(size=(5000, 3)); d /= np.linalg.norm(d, axis=1, keepdims=True)
    d = d.astype(np.float32)
    got = sh_encode(_gpu(d), 4).cpu().numpy()
    assert np.abs(got - ro.sh_encode(d, 4)).max() < 2e-6                     # fp32, FMA contraction only
    y8 = sh_encode(_gpu(d), 8).cpu().numpy().astype(np.float64)
    lo = 0
    for l in range(8):                                                        # sum_m Y_lm^2 = (2l+1)/4pi on the unit sphere
        hi = (l + 1) ** 2
        assert np.abs((y8[:, lo:hi] ** 2).sum(1) - (2 * l + 1) / (4 * np.pi)).max() < 5e-5, l
        lo = hi
    assert np.array_equal(y8[:, :16].astype(np.float32), got)
    enc = SHEncoder(3, 4)
    assert enc(_gpu(d).view(50, 100, 3)).shape == (50, 100, 16)
    with pytest.raises(AssertionError):
        SHEncoder(3, 9)


def test_fused_field_vs_per_op_and_oracle(rng):
    """pn_field_forward (one kernel) vs grid kernel + torch fp32 Linear + SH kernel vs numpy (fp64 accumulate).
    Tolerance: 2e-5 relative on sigma (exp amplifies), 2e-5 absolute on rgb — fp32 summation-order noise."""
    from pienerf_b200.network import NeRFNetwork
    field = make_field(bound=1.0, seed=3)
    model = NeRFNetwork(bound=1).cuda().load_field(field)
    x = rng.uniform(-1, 1, size=(3000, 3)).astype(np.float32); x[0] = [1.5, 0, 0]
    d = rng.normal(size=(3000, 3)); d = (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)
    s_ref, c_ref = ro.OracleField(field, accumulate=np.float64)(x, d)
    s_op, c_op = model(_gpu(x), _gpu(d))
    s_fu, c_fu = model.forward_fused(_gpu(x), _gpu(d))
    for s_, c_ in ((s_op, c_op), (s_fu, c_fu)):
        s_ = s_.cpu().numpy(); c_ = c_.cpu().numpy()
        assert np.abs(s_ / s_ref - 1).max() < 5e-4
        assert np.abs(c_ - c_ref).max() < 1e-4
    assert np.abs(s_fu.cpu().numpy() / s_op.cpu().numpy() - 1).max() < 2e-5
    assert np.abs(c_fu.cpu().numpy() - c_op.cpu().numpy()).max() < 2e-5


def test_against_reference_kernels(rng):
    """Bit-level check against the reference's own gridencoder / shencoder kernels compiled for sm_100a."""
    rg = load_ref("_ref_gridencoder"); rs = load_ref("_ref_shencoder")
    if rg is None or rs is None:
        pytest.skip("oracle/_ref not built (needs /root/reference at build time)")
    import pienerf_b200._gridencoder as ge
    import pienerf_b200._shencoder as se
    off, s = grid_offsets(desired_resolution=2048)
    emb = _gpu(rng.uniform(-1, 1, size=(int(off[-1]), 2)).astype(np.float32)); x = _gpu(_inputs(rng, 100000)); o = _gpu(off)
    a = torch.empty(16, 100000, 2, device="cuda"); b = torch.empty_like(a)
    S = float(np.log2(s))
    rg.grid_encode_forward(x, emb, o, a, 100000, 3, 2, 16, S, 16, None, 0, False, 0)
    ge.grid_encode_forward(x, emb, o, b, 100000, 3, 2, 16, S, 16, None, 0, False, 0)
    torch.cuda.synchronize()
    assert torch.equal(a, b), float((a - b).abs().max())
    dn = rng.normal(size=(10000, 3)); dn /= np.linalg.norm(dn, axis=1, keepdims=True)
    d = _gpu(dn.astype(np.float32))
    for deg in (1, 4, 8):
        ya = torch.empty(10000, deg * deg, device="cuda"); yb = torch.empty_like(ya)
        rs.sh_encode_forward(d, ya, 10000, 3, deg, None); se.sh_encode_forward(d, yb, 10000, 3, deg, None)
        torch.cuda.synchronize()
        if deg <= 4:                                                          # the hot-path degree: bit-exact
            assert torch.equal(ya, yb), (deg, float((ya - yb).abs().max()))
        else:                                                                 # bands 5-8: nvcc contracts a few FMAs differently (1 ulp)
            assert torch.allclose(ya, yb, rtol=0, atol=2e-6), (deg, float((ya - yb).abs().max()))


def test_field_tensor_core_mode(rng):
    """pn_field_forward mode 1 (tcgen05, bf16x3 split GEMMs, fp32 TMEM accumulators) vs mode 0 (fp32 SIMT) and the
    fp64-accumulate numpy oracle.  Tolerance: 2e-4 relative on sigma, 1e-4 absolute on rgb (the bar is 1e-3 RGB)."""
    from pienerf_b200.network import NeRFNetwork
    field = make_field(bound=1.0, seed=5)
    model = NeRFNetwork(bound=1).cuda().load_field(field)
    for M in (1000, 128 * 7, 50001):
        x = rng.uniform(-1, 1, size=(M, 3)).astype(np.float32); x[0] = [1.5, 0, 0]
        d = rng.normal(size=(M, 3)); d = (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)
        s0, c0 = model.forward_fused(_gpu(x), _gpu(d), mode=0)
        s1, c1 = model.forward_fused(_gpu(x), _gpu(d), mode=1)
        torch.cuda.synchronize()
        s0 = s0.cpu().numpy(); c0 = c0.cpu().numpy(); s1 = s1.cpu().numpy(); c1 = c1.cpu().numpy()
        assert np.isfinite(s1).all() and np.isfinite(c1).all()
        assert np.abs(s1 / s0 - 1).max() < 2e-4, np.abs(s1 / s0 - 1).max()
        assert np.abs(c1 - c0).max() < 1e-4, np.abs(c1 - c0).max()
    s_ref, c_ref = ro.OracleField(field, accumulate=np.float64)(x[:2000], d[:2000])
    assert np.abs(s1[:2000] / s_ref - 1).max() < 5e-4 and np.abs(c1[:2000] - c_ref).max() < 2e-4
