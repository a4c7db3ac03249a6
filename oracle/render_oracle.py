"""oracle/render_oracle.py -- TEST INFRASTRUCTURE ONLY (never imported by pienerf_b200/).

numpy fp32 restatement of the reference's deformed-space render path, vectorised over rays:
ray generation, near/far, occupancy march (plain + quadratic bending), hash-grid encode,
SH encode, the sigma/colour MLP, composite, and the `rund_cuda` / `run_cuda` driver loops.

PARITY UNPINNED by the reference (it ships no tests / golden vectors).  It is
pinned (a) on the GPU box against the reference's own CUDA kernels when oracle/_ref/*.so
exist (tests/test_ref_kernels_gpu.py), and (b) by closed-form identities (tests/test_render_oracle.py).

fp32 semantics: every array is float32; expressions the reference evaluates in double
(because of unsuffixed literals) are evaluated in float64 and rounded where the C++ does.
FMA contraction by nvcc is NOT modelled: CUDA results may differ in the last ulp.
Each function cites the reference file:line it follows.
"""
import math

import numpy as np

f32 = np.float32
SQRT3 = f32(1.7320508075688772)
FLT_MAX = np.finfo(np.float32).max


# ------------------------------------------------------------------ rays
def get_rays(pose, intrinsics, H, W):
    """nerf/utils.py:55-138 with N=-1 (full frame), B=1.  Returns rays_o, rays_d [H*W,3] f32."""
    fx, fy, cx, cy = [f32(v) for v in intrinsics]
    i = np.tile(np.linspace(0, W - 1, W, dtype=np.float32)[None, :], (H, 1)).reshape(-1) + f32(0.5)
    j = np.tile(np.linspace(0, H - 1, H, dtype=np.float32)[:, None], (1, W)).reshape(-1) + f32(0.5)
    zs = np.ones_like(i)
    xs = (i - cx) / fx * zs
    ys = (j - cy) / fy * zs
    d = np.stack((xs, ys, zs), -1)
    d = d / np.sqrt((d * d).sum(-1, keepdims=True, dtype=np.float32))
    pose = np.asarray(pose, dtype=np.float32)
    rays_d = (d @ pose[:3, :3].T).astype(np.float32)
    rays_o = np.broadcast_to(pose[:3, 3], rays_d.shape).astype(np.float32).copy()
    return rays_o, rays_d


# ------------------------------------------------------------------ near / far
def near_far_from_aabb(rays_o, rays_d, aabb, min_near):
    """raymarching/src/raymarching.cu:92-148."""
    o = rays_o.astype(np.float32); d = rays_d.astype(np.float32); aabb = np.asarray(aabb, np.float32)
    with np.errstate(divide="ignore", invalid="ignore"):
        rd = f32(1) / d
        near = (aabb[0] - o[:, 0]) * rd[:, 0]; far = (aabb[3] - o[:, 0]) * rd[:, 0]
        sw = near > far; near, far = np.where(sw, far, near), np.where(sw, near, far)
        ny = (aabb[1] - o[:, 1]) * rd[:, 1]; fy = (aabb[4] - o[:, 1]) * rd[:, 1]
        sw = ny > fy; ny, fy = np.where(sw, fy, ny), np.where(sw, ny, fy)
        miss = (near > fy) | (ny > far)
        near = np.where(ny > near, ny, near); far = np.where(fy < far, fy, far)
        nz = (aabb[2] - o[:, 2]) * rd[:, 2]; fz = (aabb[5] - o[:, 2]) * rd[:, 2]
        sw = nz > fz; nz, fz = np.where(sw, fz, nz), np.where(sw, nz, fz)
        miss |= (near > fz) | (nz > far)
        near = np.where(nz > near, nz, near); far = np.where(fz < far, fz, far)
    near = np.where(near < f32(min_near), f32(min_near), near)
    near = np.where(miss, FLT_MAX, near).astype(np.float32)
    far = np.where(miss, FLT_MAX, far).astype(np.float32)
    return near, far


# ------------------------------------------------------------------ morton / packbits
def _expand_bits(v):
    v = v.astype(np.uint64)
    v = (v * 0x00010001) & 0xFF0000FF
    v = (v * 0x00000101) & 0x0F00F00F
    v = (v * 0x00000011) & 0xC30C30C3
    v = (v * 0x00000005) & 0x49249249
    return v


def morton3D(x, y, z):
    """raymarching.cu:56-71."""
    return (_expand_bits(np.asarray(x)) | (_expand_bits(np.asarray(y)) << np.uint64(1)) |
            (_expand_bits(np.asarray(z)) << np.uint64(2))).astype(np.uint32)


def morton3D_invert(x):
    """raymarching.cu:73-81 (one coordinate; caller shifts by 0/1/2)."""
    x = np.asarray(x).astype(np.uint32) & np.uint32(0x49249249)
    x = (x | (x >> np.uint32(2))) & np.uint32(0xc30c30c3)
    x = (x | (x >> np.uint32(4))) & np.uint32(0x0f00f00f)
    x = (x | (x >> np.uint32(8))) & np.uint32(0xff0000ff)
    x = (x | (x >> np.uint32(16))) & np.uint32(0x0000ffff)
    return x


def packbits(grid, thresh):
    """raymarching.cu:271-292: bit i of byte n <- grid[8n+i] > thresh."""
    g = (np.asarray(grid, np.float32).reshape(-1, 8) > f32(thresh))
    return (g * (1 << np.arange(8))).sum(-1).astype(np.uint8)


# ------------------------------------------------------------------ march helpers
def _clampf(x, lo, hi):
    return np.minimum(f32(hi) if np.isscalar(hi) else hi, np.maximum(f32(lo) if np.isscalar(lo) else lo, x)).astype(np.float32)


def _frexp_exp(mx):
    return np.frexp(mx.astype(np.float32))[1].astype(np.int32)


def _mip_from_pos(x, y, z, C):
    """raymarching.cu:42-47."""
    mx = np.maximum(np.abs(x), np.maximum(np.abs(y), np.abs(z)))
    return np.minimum(C - 1, np.maximum(0, _frexp_exp(mx))).astype(np.int32)


def _mip_from_dt(dt, H, C):
    """raymarching.cu:49-54 (dt*H in float, *0.5 in double, frexpf on the float cast)."""
    mx = ((dt * f32(H)).astype(np.float64) * 0.5).astype(np.float32)
    return np.minimum(C - 1, np.maximum(0, _frexp_exp(mx))).astype(np.int32)


def _occupancy_and_skip(x, y, z, t, d, rd, dt_gamma, dt_min, dt_max, bound, C, H, grid):
    """Shared tail of raymarching.cu:760-807 / 1385-1432: returns (dt, occ, tt)."""
    dt = _clampf(t * f32(dt_gamma), dt_min, dt_max)
    level = np.maximum(_mip_from_pos(x, y, z, C), _mip_from_dt(dt, H, C))
    mip_bound = np.minimum(np.ldexp(f32(1), level).astype(np.float32), f32(bound))
    mip_rbound = (f32(1) / mip_bound).astype(np.float32)
    Hf = f32(H - 1)

    def cell(v):
        w = (v * mip_rbound + f32(1)).astype(np.float32)
        dbl = 0.5 * w.astype(np.float64) * H
        return _clampf(dbl.astype(np.float32), 0.0, Hf).astype(np.int32)
    nx, ny, nz = cell(x), cell(y), cell(z)
    index = (level.astype(np.int64) * (H ** 3) + morton3D(nx, ny, nz).astype(np.int64))
    occ = (grid[index // 8] & (1 << (index % 8)).astype(np.uint8)) != 0
    rH = f32(1) / f32(H)
    sgn = np.copysign(f32(1), d).astype(np.float32)

    def tnext(n, s, v, r):
        return ((((n.astype(np.float32) + f32(0.5) + f32(0.5) * s) * rH * f32(2) - f32(1)) * mip_bound - v) * r).astype(np.float32)
    with np.errstate(invalid="ignore", over="ignore"):
        tx = tnext(nx, sgn[:, 0], x, rd[:, 0]); ty = tnext(ny, sgn[:, 1], y, rd[:, 1]); tz = tnext(nz, sgn[:, 2], z, rd[:, 2])
        tt = (t + np.maximum(f32(0), np.fmin(tx, np.fmin(ty, tz)))).astype(np.float32)
    return dt, occ, tt


def _skip_to(t, tt, active, dt_gamma, dt_min, dt_max):
    """do { t += clamp(t*dt_gamma) } while (t < tt)  (raymarching.cu:804-806)."""
    t = t.copy()
    todo = active.copy()
    while todo.any():
        t[todo] = (t[todo] + _clampf(t[todo] * f32(dt_gamma), dt_min, dt_max)).astype(np.float32)
        todo &= (t < tt)
    return t


def _march_common(n_alive, n_step, rays_alive, rays_t, rays_o, rays_d, bound, dt_gamma, max_steps, C, H, grid,
                  fars, noises, warp_fn, align):
    M = n_alive * n_step
    if align > 0:
        M += align - (M % align)                              # raymarching.py:335-338 (always adds 1..align)
    xyzs = np.zeros((M, 3), np.float32); dirs = np.zeros((M, 3), np.float32); deltas = np.zeros((M, 2), np.float32)
    idx = rays_alive[:n_alive].astype(np.int64)
    o = rays_o[idx].astype(np.float32); d = rays_d[idx].astype(np.float32)
    with np.errstate(divide="ignore"):
        rd = (f32(1) / d).astype(np.float32)
    t = rays_t[idx].astype(np.float32).copy()
    far = fars[idx].astype(np.float32)
    dt_min = f32(f32(2) * SQRT3 / f32(max_steps))
    dt_max = f32(f32(2) * SQRT3 * f32(1 << (C - 1)) / f32(H))
    t = (t + _clampf(t * f32(dt_gamma), dt_min, dt_max) * noises[:n_alive].astype(np.float32)).astype(np.float32)
    last_t = t.copy()
    step = np.zeros(n_alive, np.int64)
    rows = np.arange(n_alive) * n_step
    while True:
        act = (t < far) & (step < n_step)
        if not act.any():
            break
        a = np.nonzero(act)[0]
        ta = t[a]
        x, y, z, found = warp_fn(o[a], d[a], ta)
        dt, occ, tt = _occupancy_and_skip(x, y, z, ta, d[a], rd[a], dt_gamma, dt_min, dt_max, bound, C, H, grid)
        emit = occ & found
        e = a[emit]
        r = rows[e] + step[e]
        xyzs[r, 0] = x[emit]; xyzs[r, 1] = y[emit]; xyzs[r, 2] = z[emit]
        dirs[r] = d[e]
        t[e] = (ta[emit] + dt[emit]).astype(np.float32)
        deltas[r, 0] = dt[emit]
        deltas[r, 1] = (t[e] - last_t[e]).astype(np.float32)
        last_t[e] = t[e]
        step[e] += 1
        s = a[~emit]
        if s.size:
            t[s] = _skip_to(ta[~emit], tt[~emit], np.ones(s.size, bool), dt_gamma, dt_min, dt_max)
    return xyzs, dirs, deltas


def march_rays(n_alive, n_step, rays_alive, rays_t, rays_o, rays_d, bound, dt_gamma, max_steps, C, H, grid,
               nears, fars, noises=None, align=-1):
    """raymarching.cu:704-809 + raymarching.py:299-357."""
    if noises is None:
        noises = np.zeros(n_alive, np.float32)

    def warp(o, d, t):
        x = _clampf(o[:, 0] + t * d[:, 0], -bound, bound)
        y = _clampf(o[:, 1] + t * d[:, 1], -bound, bound)
        z = _clampf(o[:, 2] + t * d[:, 2], -bound, bound)
        return x, y, z, np.ones(x.shape, bool)
    return _march_common(n_alive, n_step, rays_alive, rays_t, rays_o, rays_d, bound, dt_gamma, max_steps, C, H,
                         grid, fars, noises, warp, align)


# ------------------------------------------------------------------ IP uniform grid
def get_pnts_in_grids(pnts, bbmin, hgs, resolution):
    """nerf/utils.py:355-443: counting sort of IPs into the hgs grid.  The reference's within-cell
    order is atomic-race dependent; this oracle (and the CUDA path) use ascending IP index."""
    p = pnts.astype(np.float32); bbmin = np.asarray(bbmin, np.float32); res = np.asarray(resolution, np.int64)
    g = np.floor((p - bbmin) / f32(hgs)).astype(np.int64)
    gid = g[:, 2] * res[1] * res[0] + g[:, 1] * res[0] + g[:, 0]
    n_grid = int(res[0] * res[1] * res[2])
    cnt = np.bincount(gid, minlength=n_grid).astype(np.int32)
    bgn = (np.cumsum(cnt, dtype=np.int32) - cnt).astype(np.int32)
    idx = np.argsort(gid, kind="stable").astype(np.int32)
    return cnt, bgn, idx


_NEIGH = np.array([
    -1, 0, 0, 0, -1, 0, 0, 0, -1, 1, 0, 0, 0, 1, 0, 0, 0, 1,
    -1, -1, 0, -1, 0, -1, 0, -1, -1, 1, 1, 0, 1, 0, 1, 0, 1, 1,
    -1, 1, 0, -1, 0, 1, 0, -1, 1, 1, -1, 0, 1, 0, -1, 0, 1, -1,
    -1, -1, 1, -1, 1, -1, 1, -1, -1, 1, 1, -1, 1, -1, 1, -1, 1, 1,
    -1, -1, -1, 1, 1, 1], dtype=np.int64).reshape(26, 3)          # raymarching.cu:1011-1021 / 1089-1099


def _find_closest_IP(x, y, z, p_def, g, res, cnt, bgn, pidx):
    """raymarching.cu:986-1045 (num_seek_IP == 1): neighbours only if the own cell is empty.
    NOTE the table is applied as (f,g,h) -> (g2+f, g1+g, g0+h)."""
    n = x.shape[0]
    best = np.full(n, f32(9999.9), np.float32); ip = np.full(n, -1, np.int64)

    def scan(gid, sel):
        c = cnt[gid]; b = bgn[gid]
        for i in range(int(c.max()) if c.size else 0):
            m = sel & (i < c)
            if not m.any():
                continue
            k = pidx[np.where(m, b + i, 0)].astype(np.int64)
            pk = p_def[k]
            d2 = ((pk[:, 0] - x) * (pk[:, 0] - x) + (pk[:, 1] - y) * (pk[:, 1] - y) + (pk[:, 2] - z) * (pk[:, 2] - z)).astype(np.float32)
            u = m & (d2 < best)
            best[u] = d2[u]; ip[u] = k[u]
    gid = g[:, 2] * res[1] * res[0] + g[:, 1] * res[0] + g[:, 0]
    scan(gid, np.ones(n, bool))
    empty = ip == -1
    if empty.any():
        for k in range(26):
            f, gg, h = _NEIGH[k]
            a2 = g[:, 2] + f; a1 = g[:, 1] + gg; a0 = g[:, 0] + h
            ok = empty & ~((a2 >= res[2]) | (a2 < 0) | (a1 >= res[1]) | (a1 < 0) | (a0 >= res[0]) | (a0 < 0))
            gid2 = np.where(ok, a2 * res[1] * res[0] + a1 * res[0] + a0, 0)
            scan(gid2, ok)
    return ip


def _find_closest_IPs(x, y, z, p_def, g, res, cnt, bgn, pidx, K):
    """raymarching.cu:1047-1118: own cell then all 26 neighbours, insertion-sorted best-K.
    Here the table is applied as (dx,dy,dz) -> (g0+dx, g1+dy, g2+dz)."""
    n = x.shape[0]
    dists = np.full((n, K), FLT_MAX, np.float32); ips = np.full((n, K), -1, np.int64)

    def scan(gid, sel):
        c = np.where(sel, cnt[gid], 0); b = bgn[gid]
        for i in range(int(c.max()) if c.size else 0):
            m = sel & (i < c)
            if not m.any():
                continue
            k = pidx[np.where(m, b + i, 0)].astype(np.int64)
            pk = p_def[k]
            d2 = ((pk[:, 0] - x) * (pk[:, 0] - x) + (pk[:, 1] - y) * (pk[:, 1] - y) + (pk[:, 2] - z) * (pk[:, 2] - z)).astype(np.float32)
            placed = ~m
            for j in range(K):
                ins = (~placed) & (d2 < dists[:, j])
                if ins.any():
                    for q in range(K - 1, j, -1):
                        dists[ins, q] = dists[ins, q - 1]; ips[ins, q] = ips[ins, q - 1]
                    dists[ins, j] = d2[ins]; ips[ins, j] = k[ins]
                    placed |= ins
    gid = g[:, 2] * res[1] * res[0] + g[:, 1] * res[0] + g[:, 0]
    scan(gid, np.ones(n, bool))
    for k in range(26):
        ddx, ddy, ddz = _NEIGH[k]
        a0 = g[:, 0] + ddx; a1 = g[:, 1] + ddy; a2 = g[:, 2] + ddz
        ok = (a0 >= 0) & (a0 < res[0]) & (a1 >= 0) & (a1 < res[1]) & (a2 >= 0) & (a2 < res[2])
        gid2 = np.where(ok, a2 * res[1] * res[0] + a1 * res[0] + a0, 0)
        scan(gid2, ok)
    return ips, (ips != -1).sum(1)


def _inv3x3(A):
    """raymarching.cu:960-984 on flat [n,9]; det==0 leaves A_inv = 0."""
    det = (A[:, 0] * (A[:, 4] * A[:, 8] - A[:, 5] * A[:, 7]) - A[:, 1] * (A[:, 3] * A[:, 8] - A[:, 5] * A[:, 6])
           + A[:, 2] * (A[:, 3] * A[:, 7] - A[:, 4] * A[:, 6])).astype(np.float32)
    ok = det != 0
    with np.errstate(divide="ignore", invalid="ignore"):
        inv = (f32(1) / det).astype(np.float32)
    R = np.zeros_like(A)
    R[:, 0] = inv * (A[:, 4] * A[:, 8] - A[:, 5] * A[:, 7]); R[:, 1] = inv * (A[:, 2] * A[:, 7] - A[:, 1] * A[:, 8])
    R[:, 2] = inv * (A[:, 1] * A[:, 5] - A[:, 2] * A[:, 4]); R[:, 3] = inv * (A[:, 5] * A[:, 6] - A[:, 3] * A[:, 8])
    R[:, 4] = inv * (A[:, 0] * A[:, 8] - A[:, 2] * A[:, 6]); R[:, 5] = inv * (A[:, 2] * A[:, 3] - A[:, 0] * A[:, 5])
    R[:, 6] = inv * (A[:, 3] * A[:, 7] - A[:, 4] * A[:, 6]); R[:, 7] = inv * (A[:, 1] * A[:, 6] - A[:, 0] * A[:, 7])
    R[:, 8] = inv * (A[:, 0] * A[:, 4] - A[:, 1] * A[:, 3])
    R[~ok] = 0
    return R.astype(np.float32)


def _mul31(M, V):
    """raymarching.cu:953-958."""
    return np.stack([M[:, 0] * V[:, 0] + M[:, 3] * V[:, 1] + M[:, 6] * V[:, 2],
                     M[:, 1] * V[:, 0] + M[:, 4] * V[:, 1] + M[:, 7] * V[:, 2],
                     M[:, 2] * V[:, 0] + M[:, 5] * V[:, 1] + M[:, 8] * V[:, 2]], -1).astype(np.float32)


def _dot31(T, V):
    """raymarching.cu:940-951."""
    return (T[:, 0:9] * V[:, 0:1] + T[:, 9:18] * V[:, 1:2] + T[:, 18:27] * V[:, 2:3]).astype(np.float32)


def inverse_warp(x, y, z, ips, n_found, p_ori, p_def, F_IP, dF_IP, max_iter_num, bbmin, bbmax, IP_dx):
    """raymarching.cu:1240-1375 for samples that entered the bending branch.
    Returns x_map,y_map,z_map,found.  `ips` [n,K] (-1 padded), n_found [n]."""
    n, K = ips.shape
    bbmin = np.asarray(bbmin, np.float32); bbmax = np.asarray(bbmax, np.float32)
    n_IP = n_found.astype(np.int64).copy()
    found = n_IP > 0
    # boundary filter: loop bound shrinks while iterating (raymarching.cu:1246-1251)
    for k in range(K):
        m = found & (k < n_IP)
        if not m.any():
            continue
        pk_ = p_def[np.where(m, ips[:, k], 0)]
        out = m & ((pk_[:, 0] <= bbmin[0]) | (pk_[:, 1] <= bbmin[1]) | (pk_[:, 2] < bbmin[2]) |
                   (pk_[:, 0] >= bbmax[0]) | (pk_[:, 1] >= bbmax[1]) | (pk_[:, 2] >= bbmax[2]))
        n_IP[out] -= 1
    found &= n_IP > 0
    ps = np.zeros((n, K, 3), np.float32)
    P_ = np.stack([x, y, z], -1).astype(np.float32)
    for k in range(K):
        m = found & (k < n_IP)
        if not m.any():
            continue
        sel = np.nonzero(m)[0]
        ipk = ips[sel, k]
        pk = p_ori[ipk].astype(np.float32); pk_ = p_def[ipk].astype(np.float32)
        Fk = F_IP[ipk].astype(np.float32); dFk = dF_IP[ipk].astype(np.float32)
        p = pk.copy()
        q_ = (P_[sel] - pk_).astype(np.float32)
        run = np.ones(sel.size, bool); num_itr = np.zeros(sel.size, np.int64)
        while True:
            run &= num_itr < max_iter_num
            if not run.any():
                break
            r = np.nonzero(run)[0]
            q = (p[r] - pk[r]).astype(np.float32)
            dFq = _dot31(dFk[r], q)
            A = (Fk[r] + dFq).astype(np.float32)
            Ainv = _inv3x3(A)
            Fq = _mul31(Fk[r], q); dFqq = _mul31(dFq, q)
            b = (Fq.astype(np.float64) + 0.5 * dFqq.astype(np.float64) - q_[r].astype(np.float64)).astype(np.float32)
            dq = _mul31(Ainv, b)
            p[r] = (p[r] - dq).astype(np.float32)
            conv = (dq[:, 0] * dq[:, 0] + dq[:, 1] * dq[:, 1] + dq[:, 2] * dq[:, 2]).astype(np.float32) < f32(1e-12)
            run[r[conv]] = False
            num_itr[r[~conv]] += 1
        ppk = (p - pk).astype(np.float32)
        rej = (np.abs(ppk[:, 0]) > f32(IP_dx)) | (np.abs(ppk[:, 1]) > f32(IP_dx)) | (np.abs(ppk[:, 2]) > f32(IP_dx))
        n_IP[sel[rej]] -= 1                                    # raymarching.cu:1314-1319 (inside the k loop)
        ps[sel, k] = p
    xm = np.zeros(n, np.float32); ym = np.zeros(n, np.float32); zm = np.zeros(n, np.float32)
    m1 = found & (n_IP == 1)
    xm[m1] = ps[m1, 0, 0]; ym[m1] = ps[m1, 0, 1]; zm[m1] = ps[m1, 0, 2]
    if K >= 2:
        m2 = found & (n_IP == 2)
        if m2.any():
            s = np.nonzero(m2)[0]
            dist = np.zeros((s.size, 2), np.float32)
            for k in range(2):
                pk = p_ori[ips[s, k]].astype(np.float32)
                dist[:, k] = np.sqrt(((pk[:, 0] - x[s]) * (pk[:, 0] - x[s]) + (pk[:, 1] - y[s]) * (pk[:, 1] - y[s]) +
                                      (pk[:, 2] - z[s]) * (pk[:, 2] - z[s])).astype(np.float32))
            ds = (dist[:, 0] + dist[:, 1]).astype(np.float32)
            w0 = (dist[:, 1] / ds).astype(np.float32); w1 = (dist[:, 0] / ds).astype(np.float32)
            xm[s] = w0 * ps[s, 0, 0] + w1 * ps[s, 1, 0]; ym[s] = w0 * ps[s, 0, 1] + w1 * ps[s, 1, 1]; zm[s] = w0 * ps[s, 0, 2] + w1 * ps[s, 1, 2]
    if K >= 3:
        m3 = found & (n_IP == 3)
        if m3.any():
            s = np.nonzero(m3)[0]
            dist = np.zeros((s.size, 3), np.float32)
            for k in range(3):
                pk = p_ori[ips[s, k]].astype(np.float32)
                dist[:, k] = np.sqrt(((pk[:, 0] - x[s]) * (pk[:, 0] - x[s]) + (pk[:, 1] - y[s]) * (pk[:, 1] - y[s]) +
                                      (pk[:, 2] - z[s]) * (pk[:, 2] - z[s])).astype(np.float32))
            ds = (dist[:, 0] * dist[:, 1] + dist[:, 1] * dist[:, 2] + dist[:, 2] * dist[:, 0]).astype(np.float32)
            w0 = (dist[:, 1] * dist[:, 2] / ds).astype(np.float32); w1 = (dist[:, 0] * dist[:, 2] / ds).astype(np.float32)
            w2 = (dist[:, 0] * dist[:, 1] / ds).astype(np.float32)
            for c, out in enumerate((xm, ym, zm)):
                out[s] = w0 * ps[s, 0, c] + w1 * ps[s, 1, c] + w2 * ps[s, 2, c]
    # found stays true even when n_IP dropped to 0 inside the loop: sample maps to (0,0,0)
    xo = np.where(found, xm, x).astype(np.float32); yo = np.where(found, ym, y).astype(np.float32); zo = np.where(found, zm, z).astype(np.float32)
    return xo, yo, zo, found


def march_rays_quadratic_bending(pig_cnt, pig_bgn, pig_idx, n_vtx, n_grid, p_def, p_ori, F_IP, dF_IP, max_iter_num,
                                 bbmin, bbmax, hgs, resolution, num_seek_IP, IP_dx, cut, cut_bounds,
                                 n_alive, n_step, rays_alive, rays_t, rays_o, rays_d, bound, dt_gamma, max_steps,
                                 C, H, grid, nears, fars, noises=None, align=-1):
    """raymarching.cu:1122-1434 + raymarching.py:387-438."""
    if noises is None:
        noises = np.zeros(n_alive, np.float32)
    bbmin = np.asarray(bbmin, np.float32); bbmax = np.asarray(bbmax, np.float32)
    res = np.asarray(resolution, np.int64); cb = np.asarray(cut_bounds, np.float32)
    hi = (bbmax.astype(np.float64) - 1e-6).astype(np.float32)            # bbmax[i]-1e-6 evaluated in double
    p_def = np.asarray(p_def, np.float32); p_ori = np.asarray(p_ori, np.float32)
    F_IP = np.asarray(F_IP, np.float32); dF_IP = np.asarray(dF_IP, np.float32)

    def warp(o, d, t):
        if cut:
            x = _clampf(o[:, 0] + t * d[:, 0], -bound, bound); y = _clampf(o[:, 1] + t * d[:, 1], -bound, bound)
            z = _clampf(o[:, 2] + t * d[:, 2], -bound, bound)
            # cut test with the reference's x-for-y typo (raymarching.cu:1210)
            inside = (x > cb[0]) & (x < cb[1]) & (y > cb[2]) & (x < cb[3]) & (z > cb[4]) & (z < cb[5])
        else:
            x = _clampf(o[:, 0] + t * d[:, 0], bbmin[0], hi[0]); y = _clampf(o[:, 1] + t * d[:, 1], bbmin[1], hi[1])
            z = _clampf(o[:, 2] + t * d[:, 2], bbmin[2], hi[2])
            inside = np.ones(x.shape, bool)
        found = ~inside                                                   # outside the cut box: static, found=true
        if inside.any():
            s = np.nonzero(inside)[0]
            xs, ys, zs = x[s], y[s], z[s]
            g = np.stack([np.floor((xs - bbmin[0]) / f32(hgs)), np.floor((ys - bbmin[1]) / f32(hgs)),
                          np.floor((zs - bbmin[2]) / f32(hgs))], -1).astype(np.int64)
            if num_seek_IP == 1:
                ip = _find_closest_IP(xs, ys, zs, p_def, g, res, pig_cnt, pig_bgn, pig_idx)
                ips = ip[:, None]; nf = (ip != -1).astype(np.int64)
            else:
                ips, nf = _find_closest_IPs(xs, ys, zs, p_def, g, res, pig_cnt, pig_bgn, pig_idx, num_seek_IP)
            xm, ym, zm, fnd = inverse_warp(xs, ys, zs, ips, nf, p_ori, p_def, F_IP, dF_IP, max_iter_num, bbmin, bbmax, IP_dx)
            x = x.copy(); y = y.copy(); z = z.copy()
            x[s] = xm; y[s] = ym; z[s] = zm; found[s] = fnd
        return x, y, z, found
    return _march_common(n_alive, n_step, rays_alive, rays_t, rays_o, rays_d, bound, dt_gamma, max_steps, C, H,
                         grid, fars, noises, warp, align)


# ------------------------------------------------------------------ composite
def composite_rays(n_alive, n_step, T_thresh, rays_alive, rays_t, sigmas, rgbs, deltas, weights_sum, depth, image):
    """raymarching.cu:828-914; mutates rays_alive, rays_t, weights_sum, depth, image in place."""
    idx = rays_alive[:n_alive].astype(np.int64)
    t = rays_t[idx].astype(np.float32); ws = weights_sum[idx].astype(np.float32); d = depth[idx].astype(np.float32)
    rgb = image[idx].astype(np.float32)
    run = np.ones(n_alive, bool); step = np.zeros(n_alive, np.int64)
    sig = sigmas[:n_alive * n_step].reshape(n_alive, n_step); col = rgbs[:n_alive * n_step].reshape(n_alive, n_step, 3)
    dl = deltas[:n_alive * n_step].reshape(n_alive, n_step, 2)
    for s in range(n_step):
        run &= dl[:, s, 0] != 0
        if not run.any():
            break
        alpha = (f32(1) - np.exp(-(sig[:, s] * dl[:, s, 0]).astype(np.float32))).astype(np.float32)
        T = (f32(1) - ws).astype(np.float32)
        w = (alpha * T).astype(np.float32)
        ws = np.where(run, ws + w, ws).astype(np.float32)
        t = np.where(run, t + dl[:, s, 1], t).astype(np.float32)
        d = np.where(run, d + w * t, d).astype(np.float32)
        rgb = np.where(run[:, None], rgb + w[:, None] * col[:, s], rgb).astype(np.float32)
        stop = run & (T < f32(T_thresh))
        step = np.where(run & ~stop, step + 1, step)
        run &= ~stop
    dead = step < n_step
    rays_alive[:n_alive] = np.where(dead, -1, rays_alive[:n_alive])
    rays_t[idx[~dead]] = t[~dead]
    weights_sum[idx] = ws; depth[idx] = d; image[idx] = rgb


# ------------------------------------------------------------------ encoders + MLP
def grid_encode(inputs, embeddings, offsets, S, H, gridtype=0, align_corners=False, interp=0):
    """gridencoder/src/gridencoder.cu:87-197.  inputs [B,D] in [0,1]; returns [L,B,C] f32."""
    x = np.asarray(inputs, np.float32); emb = np.asarray(embeddings, np.float32)
    B, D = x.shape; L = len(offsets) - 1; Cc = emb.shape[1]
    out = np.zeros((L, B, Cc), np.float32)
    oob = ((x < 0) | (x > 1)).any(1)
    primes = np.array([1, 2654435761, 805459861, 3674653429, 2097192037, 1434869437, 2165219737], dtype=np.uint64)
    S32 = f32(S)
    for l in range(L):
        T = int(offsets[l + 1] - offsets[l])
        scale = f32(np.exp2(np.float64(f32(l) * S32))) * f32(H) - f32(1)       # exp2f(level*S)*H - 1
        resolution = int(math.ceil(float(scale))) + 1
        pos = (x * scale + (f32(0) if align_corners else f32(0.5))).astype(np.float32)
        pg = np.floor(pos)
        pos = (pos - pg).astype(np.float32); pg = pg.astype(np.uint64)
        if interp == 1:
            pos = (pos * pos * (f32(3) - f32(2) * pos)).astype(np.float32)
        acc = np.zeros((B, Cc), np.float32)
        for corner in range(1 << D):
            w = np.ones(B, np.float32); pl = np.empty((B, D), np.uint64)
            for dd in range(D):
                if corner & (1 << dd):
                    w = (w * pos[:, dd]).astype(np.float32); pl[:, dd] = pg[:, dd] + 1
                else:
                    w = (w * (f32(1) - pos[:, dd])).astype(np.float32); pl[:, dd] = pg[:, dd]
            stride = 1; index = np.zeros(B, np.uint64); dd = 0
            while dd < D and stride <= T:
                index = (index + pl[:, dd] * np.uint64(stride)) & np.uint64(0xFFFFFFFF)
                stride = (stride * (resolution if align_corners else resolution + 1)) & 0xFFFFFFFF
                dd += 1
            if gridtype == 0 and stride > T:
                index = np.zeros(B, np.uint64)
                for dd in range(D):
                    index ^= (pl[:, dd] * primes[dd]) & np.uint64(0xFFFFFFFF)
            index = (index % np.uint64(T)).astype(np.int64) + int(offsets[l])
            acc = (acc + w[:, None] * emb[index]).astype(np.float32)
        acc[oob] = 0
        out[l] = acc
    return out


def sh_encode(dirs, degree=4):
    """shencoder/src/shencoder.cu:27-125.  Bands 0..3 (the hot path uses degree 4) are spelled out here; higher degrees go to
    train_oracle.sh_polynomials, which carries the whole table."""
    if degree > 4:
        from .train_oracle import sh_polynomials
        return sh_polynomials(dirs, degree, np.float32)
    d = np.asarray(dirs, np.float32)
    x, y, z = d[:, 0], d[:, 1], d[:, 2]
    xy = x * y; xz = x * z; yz = y * z; x2 = x * x; y2 = y * y; z2 = z * z
    o = np.zeros((d.shape[0], degree * degree), np.float32)
    o[:, 0] = f32(0.28209479177387814)
    if degree > 1:
        o[:, 1] = f32(-0.48860251190291987) * y; o[:, 2] = f32(0.48860251190291987) * z; o[:, 3] = f32(-0.48860251190291987) * x
    if degree > 2:
        o[:, 4] = f32(1.0925484305920792) * xy; o[:, 5] = f32(-1.0925484305920792) * yz
        o[:, 6] = f32(0.94617469575755997) * z2 - f32(0.31539156525251999)
        o[:, 7] = f32(-1.0925484305920792) * xz; o[:, 8] = f32(0.54627421529603959) * x2 - f32(0.54627421529603959) * y2
    if degree > 3:
        o[:, 9] = f32(0.59004358992664352) * y * (f32(-3) * x2 + y2); o[:, 10] = f32(2.8906114426405538) * xy * z
        o[:, 11] = f32(0.45704579946446572) * y * (f32(1) - f32(5) * z2); o[:, 12] = f32(0.3731763325901154) * z * (f32(5) * z2 - f32(3))
        o[:, 13] = f32(0.45704579946446572) * x * (f32(1) - f32(5) * z2); o[:, 14] = f32(1.4453057213202769) * z * (x2 - y2)
        o[:, 15] = f32(0.59004358992664352) * x * (-x2 + f32(3) * y2)
    return o.astype(np.float32)


def mlp_forward(enc, sh, sigma_net, color_net, accumulate=np.float32):
    """nerf/network.py:98-127: sigma-net (ReLU between, trunc_exp on ch 0), colour-net on cat(SH, geo)."""
    h = enc.astype(accumulate)
    for l, W in enumerate(sigma_net):
        h = h @ W.T.astype(accumulate)
        if l != len(sigma_net) - 1:
            h = np.maximum(h, 0)
    h = h.astype(np.float32)
    sigma = np.exp(h[:, 0]).astype(np.float32)
    geo = h[:, 1:]
    h = np.concatenate([sh.astype(np.float32), geo], -1).astype(accumulate)
    for l, W in enumerate(color_net):
        h = h @ W.T.astype(accumulate)
        if l != len(color_net) - 1:
            h = np.maximum(h, 0)
    color = (1.0 / (1.0 + np.exp(-h.astype(np.float64)))).astype(np.float32)
    return sigma, color


class OracleField:
    """NeRFNetwork.forward (nerf/network.py:98-127) over a synthetic field dict (pienerf_b200.synthetic.make_field)."""

    def __init__(self, field, accumulate=np.float32):
        self.f = field; self.acc = accumulate

    def __call__(self, xyzs, dirs):
        f = self.f; b = f32(f["bound"])
        # grid.py:149; torch-CUDA divides by a python scalar as a multiply by its fp32 reciprocal
        x01 = ((xyzs.astype(np.float32) + b) * (f32(1) / (f32(2) * b))).astype(np.float32)
        enc = grid_encode(x01, f["embeddings"], f["offsets"], np.log2(f["per_level_scale"]), f["base_resolution"])
        enc = enc.transpose(1, 0, 2).reshape(x01.shape[0], -1)                             # grid.py:57
        sh = sh_encode(dirs, 4)
        return mlp_forward(enc, sh, f["sigma_net"], f["color_net"], self.acc)


# ------------------------------------------------------------------ driver loops
def rund_cuda(field_fn, rays_o, rays_d, p_def, p_ori, F_IP, dF_IP, IP_dx, density_bitfield, bound, cascade,
              min_near=0.2, density_scale=1.0, dt_gamma=0.0, max_steps=1024, T_thresh=1e-2, max_iter_num=1,
              hash_grid_size=0.06, cut=False, cut_bounds=(0, 0, 0, 0, 0, 0), num_seek_IP=1, grid_size=128,
              bg_color=1.0, return_stats=False):
    """nerf/renderer.py:755-907."""
    rays_o = np.ascontiguousarray(rays_o, np.float32).reshape(-1, 3); rays_d = np.ascontiguousarray(rays_d, np.float32).reshape(-1, 3)
    N = rays_o.shape[0]
    p_def = np.asarray(p_def, np.float32); p_ori = np.asarray(p_ori, np.float32)
    bmin = p_def.min(0); bmax = p_def.max(0)
    if cut:
        bmin = -f32(bound) * np.ones(3, np.float32); bmax = f32(bound) * np.ones(3, np.float32)
    bbmin = (bmin - f32(1e-3)).astype(np.float32); bbmax = (bmax + f32(1e-3)).astype(np.float32)
    # renderer.py:791; tensor / python-scalar on CUDA is a multiply by the fp32 reciprocal
    resolution = np.ceil((bbmax - bbmin) * (f32(1) / f32(hash_grid_size))).astype(np.int32)
    aabb = np.concatenate([bbmin, bbmax])
    nears, fars = near_far_from_aabb(rays_o, rays_d, aabb, min_near)
    n_vtx = p_ori.shape[0]; n_grid = int(resolution[0]) * int(resolution[1]) * int(resolution[2])
    pig_cnt, pig_bgn, pig_idx = get_pnts_in_grids(p_def, bbmin, hash_grid_size, resolution)
    weights_sum = np.zeros(N, np.float32); depth = np.zeros(N, np.float32); image = np.zeros((N, 3), np.float32)
    rays_alive = np.arange(N, dtype=np.int32); rays_t = nears.copy()
    step = 0; total = 0; iters = 0
    while step < max_steps:
        n_alive = rays_alive.shape[0]
        if n_alive <= 0:
            break
        n_step = max(min(N // n_alive, 8), 1)
        xyzs, dirs, deltas = march_rays_quadratic_bending(
            pig_cnt, pig_bgn, pig_idx, n_vtx, n_grid, p_def, p_ori, F_IP, dF_IP, max_iter_num, bbmin, bbmax,
            hash_grid_size, resolution, num_seek_IP, IP_dx, cut, cut_bounds, n_alive, n_step, rays_alive, rays_t,
            rays_o, rays_d, bound, dt_gamma, max_steps, cascade, grid_size, density_bitfield, nears, fars, None, 128)
        sigmas, rgbs = field_fn(xyzs, dirs)
        sigmas = (f32(density_scale) * sigmas).astype(np.float32)
        total += int((deltas[:, 0] != 0).sum()); iters += 1
        composite_rays(n_alive, n_step, T_thresh, rays_alive, rays_t, sigmas, rgbs, deltas, weights_sum, depth, image)
        rays_alive = rays_alive[rays_alive >= 0]
        step += n_step
    depth_0 = depth.copy()
    image = (image + (f32(1) - weights_sum)[:, None] * f32(bg_color)).astype(np.float32)
    with np.errstate(invalid="ignore", over="ignore"):
        depth = (np.maximum(depth - nears, 0) / (fars - nears)).astype(np.float32)
    out = {"image": image, "depth": depth, "depth_0": depth_0, "weights_sum": weights_sum}
    if return_stats:
        out["n_samples"] = total; out["iters"] = iters
    return out


def run_cuda(field_fn, rays_o, rays_d, density_bitfield, bound, cascade, min_near=0.2, density_scale=1.0,
             dt_gamma=0.0, max_steps=1024, T_thresh=1e-2, grid_size=128, bg_color=1.0, aabb=None):
    """nerf/renderer.py:332-388 (eval branch of run_cuda; the undeformed A/B)."""
    rays_o = np.ascontiguousarray(rays_o, np.float32).reshape(-1, 3); rays_d = np.ascontiguousarray(rays_d, np.float32).reshape(-1, 3)
    N = rays_o.shape[0]
    if aabb is None:
        aabb = np.array([-bound, -bound, -bound, bound, bound, bound], np.float32)
    nears, fars = near_far_from_aabb(rays_o, rays_d, aabb, min_near)
    weights_sum = np.zeros(N, np.float32); depth = np.zeros(N, np.float32); image = np.zeros((N, 3), np.float32)
    rays_alive = np.arange(N, dtype=np.int32); rays_t = nears.copy()
    step = 0
    while step < max_steps:
        n_alive = rays_alive.shape[0]
        if n_alive <= 0:
            break
        n_step = max(min(N // n_alive, 8), 1)
        xyzs, dirs, deltas = march_rays(n_alive, n_step, rays_alive, rays_t, rays_o, rays_d, bound, dt_gamma, max_steps,
                                        cascade, grid_size, density_bitfield, nears, fars, None, 128)
        sigmas, rgbs = field_fn(xyzs, dirs)
        sigmas = (f32(density_scale) * sigmas).astype(np.float32)
        composite_rays(n_alive, n_step, T_thresh, rays_alive, rays_t, sigmas, rgbs, deltas, weights_sum, depth, image)
        rays_alive = rays_alive[rays_alive >= 0]
        step += n_step
    image = (image + (f32(1) - weights_sum)[:, None] * f32(bg_color)).astype(np.float32)
    with np.errstate(invalid="ignore", over="ignore"):
        depth = (np.maximum(depth - nears, 0) / (fars - nears)).astype(np.float32)
    return {"image": image, "depth": depth, "weights_sum": weights_sum}
