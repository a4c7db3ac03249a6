"""Pins the numpy render oracle (oracle/render_oracle.py) with closed forms and structural identities
(SURVEY.md 4): Morton/packbits/near-far/composite closed forms, hash-grid vertex/linearity properties,
SH addition theorem, undeformed inverse warp == identity, rund_cuda == run_cuda on an undeformed body."""
import numpy as np

from oracle import render_oracle as ro
from pienerf_b200.synthetic import grid_offsets, morton3d
from tests.util import deformed_ip_state, small_scene

f32 = np.float32


def test_morton_roundtrip_and_packbits(rng):
    c = rng.integers(0, 128, size=(1000, 3))
    m = ro.morton3D(c[:, 0], c[:, 1], c[:, 2])
    assert (m == morton3d(c[:, 0], c[:, 1], c[:, 2])).all()
    assert m.max() < 128 ** 3
    back = np.stack([ro.morton3D_invert(m >> np.uint32(i)) for i in range(3)], 1)
    assert (back == c).all()
    assert ro.morton3D(1, 0, 0) == 1 and ro.morton3D(0, 1, 0) == 2 and ro.morton3D(0, 0, 1) == 4
    g = rng.uniform(0, 1, size=64).astype(np.float32)
    b = ro.packbits(g, 0.5)
    assert ((np.unpackbits(b, bitorder="little").astype(bool)) == (g > 0.5)).all()


def test_near_far_closed_form():
    o = np.array([[0, 0, -3], [0, 0, -3], [5, 5, -3], [0, 0, 0]], f32)
    d = np.array([[0, 0, 1], [0.0, 0.28, 0.96], [0, 0, 1], [1, 0, 0]], f32)
    aabb = np.array([-1, -1, -1, 1, 1, 1], f32)
    n, f = ro.near_far_from_aabb(o, d, aabb, 0.2)
    assert np.allclose(n[0], 2) and np.allclose(f[0], 4)
    assert n[2] == np.finfo(np.float32).max and f[2] == np.finfo(np.float32).max       # miss
    assert n[3] == f32(0.2) and np.allclose(f[3], 1)                                    # inside: near clamps to min_near
    assert n[1] < f[1]


def test_composite_closed_form():
    n_alive, n_step = 3, 4
    sig = np.full(n_alive * n_step, 2.0, f32); rgb = np.full((n_alive * n_step, 3), 0.5, f32)
    deltas = np.zeros((n_alive * n_step, 2), f32); deltas[:, 0] = 0.1; deltas[:, 1] = 0.1
    deltas[1 * n_step + 2:, :] = 0                                       # ray 1 stops after 2 samples; ray 2 has none
    alive = np.arange(3, dtype=np.int32); t = np.zeros(3, f32); ws = np.zeros(3, f32); dep = np.zeros(3, f32); img = np.zeros((3, 3), f32)
    ro.composite_rays(n_alive, n_step, 1e-4, alive, t, sig, rgb, deltas, ws, dep, img)
    a = 1 - np.exp(-0.2)
    assert np.allclose(ws[0], 1 - (1 - a) ** 4, atol=1e-6) and np.allclose(ws[1], 1 - (1 - a) ** 2, atol=1e-6) and ws[2] == 0
    assert list(alive) == [0, -1, -1] and np.allclose(t[0], 0.4) and t[1] == 0
    assert np.allclose(img[0], 0.5 * ws[0], atol=1e-6)
    # T_thresh: the sample that sees T < thresh is still accumulated, then the ray dies
    alive = np.arange(1, dtype=np.int32); t = np.zeros(1, f32); ws = np.zeros(1, f32); dep = np.zeros(1, f32); img = np.zeros((1, 3), f32)
    big = np.full(4, 100.0, f32)
    ro.composite_rays(1, 4, 0.5, alive, t, big, rgb[:4], deltas[:4], ws, dep, img)
    assert alive[0] == -1 and np.allclose(ws[0], 1.0, atol=1e-4)


def test_grid_offsets_match_survey():
    off, s = grid_offsets(desired_resolution=2048)
    assert off[-1] == 6119864 and off[1] == 4920 and off[2] - off[1] == 13824
    assert abs(s - 1.38191) < 1e-5
    assert all((off[l + 1] - off[l]) == 2 ** 19 for l in range(5, 16))
    off2, _ = grid_offsets(desired_resolution=4096)
    assert off2[-1] == 6328848


def test_grid_encode_vertex_and_linearity(rng):
    off, s = grid_offsets(desired_resolution=2048)
    S = np.log2(s)
    emb = rng.uniform(-1, 1, size=(int(off[-1]), 2)).astype(np.float32)
    # a point on a level-0 vertex returns that table entry: scale0 = 15, pos = x*15+0.5 -> x = (k-0.5)/15
    k = np.array([3, 7, 11])
    x = ((k - 0.5) / 15.0).astype(np.float32)[None]
    out = ro.grid_encode(x, emb, off, S, 16)
    idx = k[0] + k[1] * 17 + k[2] * 17 * 17
    assert np.allclose(out[0, 0], emb[idx], atol=2e-5)
    # out-of-range input -> zeros; boundary values are in range
    z = ro.grid_encode(np.array([[1.5, 0.5, 0.5], [1.0, 0.0, 1.0]], f32), emb, off, S, 16)
    assert (z[:, 0] == 0).all() and np.abs(z[:, 1]).max() > 0
    # linear in the table
    emb2 = rng.uniform(-1, 1, size=emb.shape).astype(np.float32)
    pts = rng.uniform(0, 1, size=(64, 3)).astype(np.float32)
    a = ro.grid_encode(pts, emb, off, S, 16); b = ro.grid_encode(pts, emb2, off, S, 16)
    c = ro.grid_encode(pts, (f32(2) * emb - f32(0.5) * emb2).astype(np.float32), off, S, 16)
    assert np.abs(c - (2 * a - 0.5 * b)).max() < 1e-5
    # interpolation weights sum to one: constant table -> constant output
    one = ro.grid_encode(pts, np.ones_like(emb), off, S, 16)
    assert np.abs(one - 1).max() < 1e-6


def test_sh_addition_theorem(rng):
    d = rng.normal(size=(100, 3)); d /= np.linalg.norm(d, axis=1, keepdims=True)
    y = ro.sh_encode(d.astype(np.float32), 4).astype(np.float64)
    for l, (a, b) in enumerate([(0, 1), (1, 4), (4, 9), (9, 16)]):
        assert np.abs((y[:, a:b] ** 2).sum(1) - (2 * l + 1) / (4 * np.pi)).max() < 1e-5


def test_undeformed_inverse_warp_is_identity(rng):
    body, field, bits, pose, intr = small_scene()
    p_ori, _, _, _ = deformed_ip_state(body, amp=0.0)
    n = p_ori.shape[0]
    F = np.tile(np.eye(3, dtype=np.float32).reshape(1, 9), (n, 1)); dF = np.zeros((n, 27), np.float32)
    bbmin = p_ori.min(0) - f32(1e-3); bbmax = p_ori.max(0) + f32(1e-3)
    res = np.ceil((bbmax - bbmin) / f32(0.06)).astype(np.int32)
    cnt, bgn, idx = ro.get_pnts_in_grids(p_ori, bbmin, 0.06, res)
    assert cnt.sum() == n and (np.sort(idx) == np.arange(n)).all()
    x = rng.uniform(bbmin + 0.01, bbmax - 0.01, size=(200, 3)).astype(np.float32)
    g = np.floor((x - bbmin) / f32(0.06)).astype(np.int64)
    for K in (1, 3):
        if K == 1:
            ip = ro._find_closest_IP(x[:, 0], x[:, 1], x[:, 2], p_ori, g, res.astype(np.int64), cnt, bgn, idx)
            ips, nf = ip[:, None], (ip != -1).astype(np.int64)
            d2 = ((p_ori[None] - x[:, None]) ** 2).sum(-1)
            has = ip != -1
            # whenever the own cell is non-empty the pick is the nearest IP *of that cell*
            assert has.any()
        else:
            ips, nf = ro._find_closest_IPs(x[:, 0], x[:, 1], x[:, 2], p_ori, g, res.astype(np.int64), cnt, bgn, idx, K)
            d2 = ((p_ori[None] - x[:, None]) ** 2).sum(-1)
            assert (ips[:, 0] == d2.argmin(1)).all()                                  # 27-cell search finds the true nearest
            assert (nf == 3).all()
        xm, ym, zm, found = ro.inverse_warp(x[:, 0], x[:, 1], x[:, 2], ips, nf, p_ori, p_ori, F, dF, 1, bbmin, bbmax, 0.0525)
        ok = found & (nf > 0)
        err = np.abs(np.stack([xm, ym, zm], 1) - x)[ok]
        # identity wherever every candidate passes the |p - p_ori| <= IP_dx test
        assert np.median(err) < 1e-6


def test_quadratic_warp_newton_converges():
    body, *_ = small_scene()
    p_ori, p_def, F, dF = deformed_ip_state(body, amp=0.05)
    # forward-map rest points near IP 10 through the exact quadratic field, then invert with many iterations
    k = 10
    q = np.array([[0.01, -0.008, 0.005]], np.float32)
    Fm = F[k].reshape(3, 3).T.astype(np.float64)                                  # F[b][a] at a*3+b
    Hk = dF[k].reshape(3, 3, 3).astype(np.float64)                                # [c][r][j]
    xdef = p_def[k].astype(np.float64) + Fm @ q[0] + 0.5 * np.einsum("crj,j,c->r", Hk, q[0].astype(np.float64), q[0].astype(np.float64))
    x = xdef.astype(np.float32)[None]
    ips = np.array([[k]]); nf = np.array([1])
    big = np.array([-10, -10, -10], np.float32)
    xm, ym, zm, found = ro.inverse_warp(x[:, 0], x[:, 1], x[:, 2], ips, nf, p_ori, p_def, F, dF, 100, big, -big, 1.0)
    # symmetric H (as produced by a true quadratic map) => the reference's mixed-index second-order term is exact
    assert found[0] and np.abs(np.array([xm[0], ym[0], zm[0]]) - (p_ori[k] + q[0])).max() < 2e-6
    xm1, ym1, zm1, _ = ro.inverse_warp(x[:, 0], x[:, 1], x[:, 2], ips, nf, p_ori, p_def, F, dF, 1, big, -big, 1.0)
    lin = p_ori[k] + np.linalg.solve(Fm, (xdef - p_def[k]))
    assert np.abs(np.array([xm1[0], ym1[0], zm1[0]]) - lin).max() < 2e-6             # max_iter_num=1 is the linear warp


def test_rund_equals_run_on_undeformed_body():
    body, field, bits, pose, intr = small_scene(W=24, H=24)
    p_ori, _, _, _ = deformed_ip_state(body, amp=0.0)
    n = p_ori.shape[0]
    F = np.tile(np.eye(3, dtype=np.float32).reshape(1, 9), (n, 1)); dF = np.zeros((n, 27), np.float32)
    rays_o, rays_d = ro.get_rays(pose, intr, 24, 24)
    fld = ro.OracleField(field)
    kw = dict(min_near=0.2, density_scale=20.0, dt_gamma=0.0, max_steps=256, T_thresh=1e-2)
    a = ro.rund_cuda(fld, rays_o, rays_d, p_ori, p_ori, F, dF, 0.0525, bits, 1.0, 1, max_iter_num=1, hash_grid_size=0.06,
                     num_seek_IP=3, return_stats=True, **kw)
    bbmin = p_ori.min(0) - f32(1e-3); bbmax = p_ori.max(0) + f32(1e-3)
    b = ro.run_cuda(fld, rays_o, rays_d, bits, 1.0, 1, aabb=np.concatenate([bbmin, bbmax]), **kw)
    assert a["n_samples"] > 200
    hit = a["weights_sum"] > 0
    assert hit.sum() > 20
    # same samples (identity warp up to ~1e-7 in position) => same image up to field sensitivity
    assert np.abs(a["image"] - b["image"]).max() < 5e-3
    assert np.median(np.abs(a["image"] - b["image"])[hit]) < 1e-4
    assert (a["image"][~hit] == 1).all()                                           # white background
