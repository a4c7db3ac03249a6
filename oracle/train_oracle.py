"""oracle/train_oracle.py -- TEST INFRASTRUCTURE ONLY (never imported by pienerf_b200/).

numpy restatement of the reference's TRAINING-side kernels, each function citing the reference lines it
follows.  PARITY PINNED: tests/golden/ref_train.npz holds outputs of the reference's own kernels (oracle/_ref/*.so, i.e.
/root/reference/{raymarching,gridencoder,shencoder}/src/*.cu compiled unmodified for sm_100a) on seeded inputs, produced
on a B200 by tests/golden/make_golden_train.py; tests/test_train_oracle.py checks every function below against it.

Where the reference sums through atomics (table gradients, total variation) its fp32 result depends on retirement order;
the oracle sums in float64 and the comparisons carry a tolerance.  Where the reference packs rays in atomic order
(march_rays_train) the oracle, like the CUDA library, packs in ray order; comparisons against the reference are per ray.
"""
import math

import numpy as np

from . import render_oracle as ro

f32 = np.float32


# ------------------------------------------------------------------ raymarching (training)
def march_rays_train(rays_o, rays_d, grid, bound, dt_gamma, max_steps, C, H, M, nears, fars, noises, counter=(0, 0)):
    """raymarching/src/raymarching.cu:305-493.  Returns (xyzs [M,3], dirs [M,3], deltas [M,2], rays [N,3] i32, counter [2])
    with ray n in row n and sample offsets ascending with n, starting at counter[0]."""
    o = np.asarray(rays_o, np.float32).reshape(-1, 3); d = np.asarray(rays_d, np.float32).reshape(-1, 3)
    N = o.shape[0]
    with np.errstate(divide="ignore"):
        rd = (f32(1) / d).astype(np.float32)
    near = np.asarray(nears, np.float32); far = np.asarray(fars, np.float32)
    dt_min = f32(f32(2) * ro.SQRT3 / f32(max_steps))
    dt_max = f32(f32(2) * ro.SQRT3 * f32(1 << (C - 1)) / f32(H))
    with np.errstate(over="ignore", invalid="ignore"):
        t0 = (near + ro._clampf(near * f32(dt_gamma), dt_min, dt_max) * np.asarray(noises, np.float32)).astype(np.float32)   # :352-355
    t = t0.copy(); last_t = t0.copy()
    step = np.zeros(N, np.int64)
    samples = [[] for _ in range(N)]                                 # per ray: (x, y, z, dt, t - last_t)
    while True:
        act = (t < far) & (step < max_steps)                         # :363 / :430 (both passes walk the same sequence)
        if not act.any():
            break
        a = np.nonzero(act)[0]
        ta = t[a]
        x = ro._clampf(o[a, 0] + ta * d[a, 0], -bound, bound); y = ro._clampf(o[a, 1] + ta * d[a, 1], -bound, bound)
        z = ro._clampf(o[a, 2] + ta * d[a, 2], -bound, bound)
        dt, occ, tt = ro._occupancy_and_skip(x, y, z, ta, d[a], rd[a], dt_gamma, dt_min, dt_max, bound, C, H, grid)
        e = a[occ]
        t[e] = (ta[occ] + dt[occ]).astype(np.float32)
        dl = (t[e] - last_t[e]).astype(np.float32)
        last_t[e] = t[e]
        step[e] += 1
        for k, ray in enumerate(e):
            samples[ray].append((x[occ][k], y[occ][k], z[occ][k], dt[occ][k], dl[k]))
        s = a[~occ]
        if s.size:
            t[s] = ro._skip_to(ta[~occ], tt[~occ], np.ones(s.size, bool), dt_gamma, dt_min, dt_max)
    xyzs = np.zeros((M, 3), np.float32); dirs = np.zeros((M, 3), np.float32); deltas = np.zeros((M, 2), np.float32)
    rays = np.zeros((N, 3), np.int32)
    off = int(counter[0])
    for n in range(N):
        k = len(samples[n])
        rays[n] = (n, off, k)
        if k and off + k <= M:                                       # :418-419: a ray that does not fit is dropped whole
            sm = np.array(samples[n], np.float32)
            xyzs[off:off + k] = sm[:, :3]; dirs[off:off + k] = d[n]; deltas[off:off + k] = sm[:, 3:]
        off += k
    return xyzs, dirs, deltas, rays, np.array([off, int(counter[1]) + N], np.int32)


def _expf_fast(x):
    """__expf(x) = ex2.approx(x * log2(e)) of the reference; restated in float32 (the tests allow the approx-unit ulps)."""
    return np.exp2((np.asarray(x, np.float32) * f32(1.4426950408889634)).astype(np.float32)).astype(np.float32)


def composite_rays_train_forward(sigmas, rgbs, deltas, rays, T_thresh=1e-4):
    """raymarching.cu:503-591.  Returns (weights_sum [N], depth [N], image [N,3]) indexed by rays[:,0]."""
    sigmas = np.asarray(sigmas, np.float32); rgbs = np.asarray(rgbs, np.float32); deltas = np.asarray(deltas, np.float32)
    M = sigmas.shape[0]; N = rays.shape[0]
    ws_out = np.zeros(N, np.float32); depth = np.zeros(N, np.float32); image = np.zeros((N, 3), np.float32)
    for n in range(N):
        index, offset, num = (int(v) for v in rays[n])
        ws = f32(0); t = f32(0); dd = f32(0); col = np.zeros(3, np.float32); T = f32(1)
        if num != 0 and offset + num <= M:
            for s in range(offset, offset + num):
                alpha = f32(1) - _expf_fast(-sigmas[s] * deltas[s, 0])
                w = f32(alpha * T)
                col = (col + w * rgbs[s]).astype(np.float32)
                t = f32(t + deltas[s, 1]); dd = f32(dd + w * t); ws = f32(ws + w)
                T = f32(T * (f32(1) - alpha))
                if T < T_thresh:
                    break
        ws_out[index] = ws; depth[index] = dd; image[index] = col
    return ws_out, depth, image


def composite_rays_train_backward(grad_weights_sum, grad_image, sigmas, rgbs, deltas, rays, weights_sum, image, T_thresh=1e-4):
    """raymarching.cu:604-696.  Returns (grad_sigmas [M], grad_rgbs [M,3]); the depth gradient is not propagated."""
    sigmas = np.asarray(sigmas, np.float32); rgbs = np.asarray(rgbs, np.float32); deltas = np.asarray(deltas, np.float32)
    gws = np.asarray(grad_weights_sum, np.float32); gim = np.asarray(grad_image, np.float32)
    M = sigmas.shape[0]; N = rays.shape[0]
    gs = np.zeros(M, np.float32); gc = np.zeros((M, 3), np.float32)
    for n in range(N):
        index, offset, num = (int(v) for v in rays[n])
        if num == 0 or offset + num > M:
            continue
        final = image[index].astype(np.float32); ws_final = f32(weights_sum[index])
        col = np.zeros(3, np.float32); T = f32(1)
        for s in range(offset, offset + num):
            alpha = f32(1) - _expf_fast(-sigmas[s] * deltas[s, 0])
            w = f32(alpha * T)
            col = (col + w * rgbs[s]).astype(np.float32)
            T = f32(T * (f32(1) - alpha))
            gc[s] = gim[index] * w
            inner = (gim[index] * (T * rgbs[s] - (final - col))).astype(np.float32)
            gs[s] = deltas[s, 0] * (inner[0] + inner[1] + inner[2] + gws[index] * (f32(1) - ws_final))
            if T < T_thresh:
                break
    return gs, gc


# ------------------------------------------------------------------ gridencoder (training)
def _level_cells(x, level, offsets, S, H, align_corners, interp):
    """Per-level geometry of gridencoder.cu:100-160 / 262-296: (T, resolution, stride1, cell [B,D] u64, frac [B,D] f32, dfrac)."""
    T = int(offsets[level + 1] - offsets[level])
    scale = f32(np.exp2(np.float64(f32(level) * f32(S)))) * f32(H) - f32(1)
    resolution = int(math.ceil(float(scale))) + 1
    pos = (x * scale + (f32(0) if align_corners else f32(0.5))).astype(np.float32)
    pg = np.floor(pos)
    pos = (pos - pg).astype(np.float32)
    dfrac = np.ones_like(pos)
    if interp == 1:
        dfrac = (f32(6) * pos * (f32(1) - pos)).astype(np.float32)
        pos = (pos * pos * (f32(3) - f32(2) * pos)).astype(np.float32)
    return T, scale, resolution, (resolution if align_corners else resolution + 1), pg.astype(np.uint64), pos, dfrac


_PRIMES = np.array([1, 2654435761, 805459861, 3674653429, 2097192037, 1434869437, 2165219737], dtype=np.uint64)


def _vertex_index(pl, T, stride1, gridtype):
    """gridencoder.cu:50-84 get_grid_index (entry index inside the level)."""
    B, D = pl.shape
    stride = 1; index = np.zeros(B, np.uint64); dd = 0
    while dd < D and stride <= T:
        index = (index + pl[:, dd] * np.uint64(stride)) & np.uint64(0xFFFFFFFF)
        stride = (stride * stride1) & 0xFFFFFFFF
        dd += 1
    if gridtype == 0 and stride > T:
        index = np.zeros(B, np.uint64)
        for dd in range(D):
            index ^= (pl[:, dd] * _PRIMES[dd]) & np.uint64(0xFFFFFFFF)
    return (index % np.uint64(T)).astype(np.int64)


def _corners(cell, frac, D):
    for corner in range(1 << D):
        w = np.ones(cell.shape[0], np.float32); pl = cell.copy()
        for dd in range(D):
            if corner & (1 << dd):
                w = (w * frac[:, dd]).astype(np.float32); pl[:, dd] = cell[:, dd] + np.uint64(1)
            else:
                w = (w * (f32(1) - frac[:, dd])).astype(np.float32)
        yield w, pl


def grid_forward_via_corners(inputs, embeddings, offsets, S, H, gridtype=0, align_corners=False, interp=0):
    """The forward pass through the helpers above; tests pin it to render_oracle.grid_encode (golden-pinned) bit-for-bit."""
    x = np.asarray(inputs, np.float32); emb = np.asarray(embeddings, np.float32)
    B, D = x.shape; L = len(offsets) - 1
    out = np.zeros((L, B, emb.shape[1]), np.float32)
    oob = ((x < 0) | (x > 1)).any(1)
    for l in range(L):
        T, _, _, stride1, cell, frac, _ = _level_cells(x, l, offsets, S, H, align_corners, interp)
        acc = np.zeros((B, emb.shape[1]), np.float32)
        for w, pl in _corners(cell, frac, D):
            acc = (acc + w[:, None] * emb[_vertex_index(pl, T, stride1, gridtype) + int(offsets[l])]).astype(np.float32)
        acc[oob] = 0
        out[l] = acc
    return out


def grid_encode_backward(grad, inputs, n_entries, offsets, S, H, gridtype=0, align_corners=False, interp=0):
    """gridencoder.cu:248-343: grad [L,B,C] -> grad_embeddings [n_entries,C]; float64 sums, returned as float64."""
    x = np.asarray(inputs, np.float32); g = np.asarray(grad, np.float32).astype(np.float64)
    L, B, C = g.shape; D = x.shape[1]
    out = np.zeros((n_entries, C), np.float64)
    ok = ~((x < 0) | (x > 1)).any(1)
    for l in range(L):
        T, _, _, stride1, cell, frac, _ = _level_cells(x, l, offsets, S, H, align_corners, interp)
        for w, pl in _corners(cell, frac, D):
            idx = _vertex_index(pl, T, stride1, gridtype) + int(offsets[l])
            np.add.at(out, idx[ok], (w[:, None].astype(np.float64) * g[l])[ok])
    return out


def grid_dy_dx(inputs, embeddings, offsets, S, H, gridtype=0, align_corners=False, interp=0):
    """gridencoder.cu:201-244: dy_dx [B, L, D, C] (float32 arithmetic, corner order of the reference)."""
    x = np.asarray(inputs, np.float32); emb = np.asarray(embeddings, np.float32)
    B, D = x.shape; L = len(offsets) - 1; C = emb.shape[1]
    out = np.zeros((B, L, D, C), np.float32)
    oob = ((x < 0) | (x > 1)).any(1)
    for l in range(L):
        T, scale, _, stride1, cell, frac, dfrac = _level_cells(x, l, offsets, S, H, align_corners, interp)
        for gd in range(D):
            acc = np.zeros((B, C), np.float32)
            others = [d for d in range(D) if d != gd]
            for corner in range(1 << (D - 1)):
                w = np.full(B, scale, np.float32); pl = cell.copy()
                for nd, d in enumerate(others):
                    if corner & (1 << nd):
                        w = (w * frac[:, d]).astype(np.float32); pl[:, d] = cell[:, d] + np.uint64(1)
                    else:
                        w = (w * (f32(1) - frac[:, d])).astype(np.float32)
                left = emb[_vertex_index(pl, T, stride1, gridtype) + int(offsets[l])]
                pl[:, gd] = cell[:, gd] + np.uint64(1)
                right = emb[_vertex_index(pl, T, stride1, gridtype) + int(offsets[l])]
                acc = (acc + (w[:, None] * (right - left)).astype(np.float32) * dfrac[:, gd:gd + 1]).astype(np.float32)
            acc[oob] = 0
            out[:, l, gd] = acc
    return out


def grid_input_backward(grad, dy_dx):
    """gridencoder.cu:346-370: grad [L,B,C], dy_dx [B,L,D,C] -> grad_inputs [B,D], summed in (l, c) order in float32."""
    g = np.asarray(grad, np.float32); j = np.asarray(dy_dx, np.float32)
    L, B, C = g.shape; D = j.shape[2]
    out = np.zeros((B, D), np.float32)
    for l in range(L):
        for c in range(C):
            out = (out + g[l, :, c:c + 1] * j[:, l, :, c]).astype(np.float32)
    return out


def grad_total_variation(inputs, embeddings, offsets, weight, S, H, gridtype=0, align_corners=False):
    """gridencoder.cu:505-610: returns the TV gradient to ADD to embeddings.grad ([n_entries,C], float64 sums)."""
    x = np.asarray(inputs, np.float32); emb = np.asarray(embeddings, np.float32)
    B, D = x.shape; L = len(offsets) - 1; C = emb.shape[1]
    out = np.zeros(emb.shape, np.float64)
    ok = ~((x < 0) | (x > 1)).any(1)
    w = f32(f32(weight) / f32(2 * D))
    for l in range(L):
        T, _, resolution, stride1, cell, _, _ = _level_cells(x, l, offsets, S, H, align_corners, 0)
        base = int(offsets[l])
        idx = _vertex_index(cell, T, stride1, gridtype) + base
        here = emb[idx]
        sm = np.zeros((B, C), np.float32); sq = np.zeros((B, C), np.float32)
        for d in range(D):
            for side, cond in ((1, cell[:, d] < np.uint64(resolution)), (-1, cell[:, d] > 0)):
                pl = cell.copy()
                pl[:, d] = (cell[:, d].astype(np.int64) + side).astype(np.uint64) & np.uint64(0xFFFFFFFF)
                diff = (here - emb[_vertex_index(pl, T, stride1, gridtype) + base]).astype(np.float32)
                diff[~cond] = 0
                sm = (sm + diff).astype(np.float32); sq = (sq + diff * diff).astype(np.float32)
        val = (w * sm * (f32(1) / np.sqrt((sq + f32(1e-9)).astype(np.float32)))).astype(np.float32)
        np.add.at(out, idx[ok], val[ok].astype(np.float64))
    return out


# ------------------------------------------------------------------ shencoder (training)
def sh_polynomials(dirs, degree, dtype=np.float32):
    """shencoder/src/shencoder.cu:49-123: the 64 Cartesian polynomials, evaluated in `dtype` term for term (bands 4..7 here;
    bands 0..3 through the same expressions as render_oracle.sh_encode).  Used in float32 for values and in float64 for the
    central differences behind sh_jacobian at degree > 4."""
    d = np.asarray(dirs, dtype)
    c = dtype
    x, y, z = d[:, 0], d[:, 1], d[:, 2]
    xy = x * y; xz = x * z; yz = y * z; x2 = x * x; y2 = y * y; z2 = z * z
    x4 = x2 * x2; y4 = y2 * y2; z4 = z2 * z2
    x6 = x4 * x2; y6 = y4 * y2; z6 = z4 * z2
    o = np.zeros((d.shape[0], degree * degree), dtype)
    o[:, 0] = c(0.28209479177387814)
    if degree > 1:
        o[:, 1] = c(-0.48860251190291987) * y; o[:, 2] = c(0.48860251190291987) * z; o[:, 3] = c(-0.48860251190291987) * x
    if degree > 2:
        o[:, 4] = c(1.0925484305920792) * xy; o[:, 5] = c(-1.0925484305920792) * yz
        o[:, 6] = c(0.94617469575755997) * z2 - c(0.31539156525251999)
        o[:, 7] = c(-1.0925484305920792) * xz; o[:, 8] = c(0.54627421529603959) * x2 - c(0.54627421529603959) * y2
    if degree > 3:
        o[:, 9] = c(0.59004358992664352) * y * (c(-3) * x2 + y2); o[:, 10] = c(2.8906114426405538) * xy * z
        o[:, 11] = c(0.45704579946446572) * y * (c(1) - c(5) * z2); o[:, 12] = c(0.3731763325901154) * z * (c(5) * z2 - c(3))
        o[:, 13] = c(0.45704579946446572) * x * (c(1) - c(5) * z2); o[:, 14] = c(1.4453057213202769) * z * (x2 - y2)
        o[:, 15] = c(0.59004358992664352) * x * (-x2 + c(3) * y2)
    if degree > 4:
        o[:, 16] = c(2.5033429417967046) * xy * (x2 - y2)
        o[:, 17] = c(1.7701307697799304) * yz * (-c(3.0) * x2 + y2)
        o[:, 18] = c(0.94617469575756008) * xy * (c(7.0) * z2 - c(1.0))
        o[:, 19] = c(0.66904654355728921) * yz * (c(3.0) - c(7.0) * z2)
        o[:, 20] = -c(3.1735664074561294) * z2 + c(3.7024941420321507) * z4 + c(0.31735664074561293)
        o[:, 21] = c(0.66904654355728921) * xz * (c(3.0) - c(7.0) * z2)
        o[:, 22] = c(0.47308734787878004) * (x2 - y2) * (c(7.0) * z2 - c(1.0))
        o[:, 23] = c(1.7701307697799304) * xz * (-x2 + c(3.0) * y2)
        o[:, 24] = -c(3.7550144126950569) * x2 * y2 + c(0.62583573544917614) * x4 + c(0.62583573544917614) * y4
    if degree > 5:
        o[:, 25] = c(0.65638205684017015) * y * (c(10.0) * x2 * y2 - c(5.0) * x4 - y4)
        o[:, 26] = c(8.3026492595241645) * xy * z * (x2 - y2)
        o[:, 27] = -c(0.48923829943525038) * y * (c(3.0) * x2 - y2) * (c(9.0) * z2 - c(1.0))
        o[:, 28] = c(4.7935367849733241) * xy * z * (c(3.0) * z2 - c(1.0))
        o[:, 29] = c(0.45294665119569694) * y * (c(14.0) * z2 - c(21.0) * z4 - c(1.0))
        o[:, 30] = c(0.1169503224534236) * z * (-c(70.0) * z2 + c(63.0) * z4 + c(15.0))
        o[:, 31] = c(0.45294665119569694) * x * (c(14.0) * z2 - c(21.0) * z4 - c(1.0))
        o[:, 32] = c(2.3967683924866621) * z * (x2 - y2) * (c(3.0) * z2 - c(1.0))
        o[:, 33] = -c(0.48923829943525038) * x * (x2 - c(3.0) * y2) * (c(9.0) * z2 - c(1.0))
        o[:, 34] = c(2.0756623148810411) * z * (-c(6.0) * x2 * y2 + x4 + y4)
        o[:, 35] = c(0.65638205684017015) * x * (c(10.0) * x2 * y2 - x4 - c(5.0) * y4)
    if degree > 6:
        o[:, 36] = c(1.3663682103838286) * xy * (-c(10.0) * x2 * y2 + c(3.0) * x4 + c(3.0) * y4)
        o[:, 37] = c(2.3666191622317521) * yz * (c(10.0) * x2 * y2 - c(5.0) * x4 - y4)
        o[:, 38] = c(2.0182596029148963) * xy * (x2 - y2) * (c(11.0) * z2 - c(1.0))
        o[:, 39] = -c(0.92120525951492349) * yz * (c(3.0) * x2 - y2) * (c(11.0) * z2 - c(3.0))
        o[:, 40] = c(0.92120525951492349) * xy * (-c(18.0) * z2 + c(33.0) * z4 + c(1.0))
        o[:, 41] = c(0.58262136251873131) * yz * (c(30.0) * z2 - c(33.0) * z4 - c(5.0))
        o[:, 42] = c(6.6747662381009842) * z2 - c(20.024298714302954) * z4 + c(14.684485723822165) * z6 - c(0.31784601133814211)
        o[:, 43] = c(0.58262136251873131) * xz * (c(30.0) * z2 - c(33.0) * z4 - c(5.0))
        o[:, 44] = c(0.46060262975746175) * (x2 - y2) * (c(11.0) * z2 * (c(3.0) * z2 - c(1.0)) - c(7.0) * z2 + c(1.0))
        o[:, 45] = -c(0.92120525951492349) * xz * (x2 - c(3.0) * y2) * (c(11.0) * z2 - c(3.0))
        o[:, 46] = c(0.50456490072872406) * (c(11.0) * z2 - c(1.0)) * (-c(6.0) * x2 * y2 + x4 + y4)
        o[:, 47] = c(2.3666191622317521) * xz * (c(10.0) * x2 * y2 - x4 - c(5.0) * y4)
        o[:, 48] = c(10.247761577878714) * x2 * y4 - c(10.247761577878714) * x4 * y2 + c(0.6831841051919143) * x6 - c(0.6831841051919143) * y6
    if degree > 7:
        o[:, 49] = c(0.70716273252459627) * y * (-c(21.0) * x2 * y4 + c(35.0) * x4 * y2 - c(7.0) * x6 + y6)
        o[:, 50] = c(5.2919213236038001) * xy * z * (-c(10.0) * x2 * y2 + c(3.0) * x4 + c(3.0) * y4)
        o[:, 51] = -c(0.51891557872026028) * y * (c(13.0) * z2 - c(1.0)) * (-c(10.0) * x2 * y2 + c(5.0) * x4 + y4)
        o[:, 52] = c(4.1513246297620823) * xy * z * (x2 - y2) * (c(13.0) * z2 - c(3.0))
        o[:, 53] = -c(0.15645893386229404) * y * (c(3.0) * x2 - y2) * (c(13.0) * z2 * (c(11.0) * z2 - c(3.0)) - c(27.0) * z2 + c(3.0))
        o[:, 54] = c(0.44253269244498261) * xy * z * (-c(110.0) * z2 + c(143.0) * z4 + c(15.0))
        o[:, 55] = c(0.090331607582517306) * y * (-c(135.0) * z2 + c(495.0) * z4 - c(429.0) * z6 + c(5.0))
        o[:, 56] = c(0.068284276912004949) * z * (c(315.0) * z2 - c(693.0) * z4 + c(429.0) * z6 - c(35.0))
        o[:, 57] = c(0.090331607582517306) * x * (-c(135.0) * z2 + c(495.0) * z4 - c(429.0) * z6 + c(5.0))
        o[:, 58] = c(0.07375544874083044) * z * (x2 - y2) * (c(143.0) * z2 * (c(3.0) * z2 - c(1.0)) - c(187.0) * z2 + c(45.0))
        o[:, 59] = -c(0.15645893386229404) * x * (x2 - c(3.0) * y2) * (c(13.0) * z2 * (c(11.0) * z2 - c(3.0)) - c(27.0) * z2 + c(3.0))
        o[:, 60] = c(1.0378311574405206) * z * (c(13.0) * z2 - c(3.0)) * (-c(6.0) * x2 * y2 + x4 + y4)
        o[:, 61] = -c(0.51891557872026028) * x * (c(13.0) * z2 - c(1.0)) * (-c(10.0) * x2 * y2 + x4 + c(5.0) * y4)
        o[:, 62] = c(2.6459606618019) * z * (c(15.0) * x2 * y4 - c(15.0) * x4 * y2 + x6 - y6)
        o[:, 63] = c(0.70716273252459627) * x * (-c(35.0) * x2 * y4 + c(21.0) * x4 * y2 - x6 + c(7.0) * y6)
    return o


def sh_jacobian(dirs, degree=4):
    """d Y / d (x,y,z) for bands below `degree`: [B, 3, degree^2].  Bands 0..3: derivatives of the polynomials of
    shencoder.cu:49-77 taken by hand; bands 4..7: float64 central differences of sh_polynomials.  The reference tabulates
    the same derivatives at shencoder.cu:128-344."""
    v = np.asarray(dirs, np.float32)
    x, y, z = v[:, 0], v[:, 1], v[:, 2]
    B = v.shape[0]
    J = np.zeros((B, 3, degree * degree), np.float32)
    X, Y, Z = 0, 1, 2
    if degree > 1:
        c1 = f32(0.48860251190291987)
        J[:, Y, 1] = -c1; J[:, Z, 2] = c1; J[:, X, 3] = -c1
    if degree > 2:
        c2 = f32(1.0925484305920792)
        J[:, X, 4] = c2 * y; J[:, Y, 4] = c2 * x
        J[:, Y, 5] = -c2 * z; J[:, Z, 5] = -c2 * y
        J[:, Z, 6] = f32(2 * 0.94617469575755997) * z
        J[:, X, 7] = -c2 * z; J[:, Z, 7] = -c2 * x
        J[:, X, 8] = f32(2 * 0.54627421529603959) * x; J[:, Y, 8] = f32(-2 * 0.54627421529603959) * y
    if degree > 3:
        x2, y2, z2 = x * x, y * y, z * z
        a = f32(0.59004358992664352); b = f32(2.8906114426405538); c = f32(0.45704579946446572)
        d = f32(0.3731763325901154); e = f32(1.4453057213202769)
        # 9: a*y*(y2 - 3x2)
        J[:, X, 9] = a * y * (f32(-6) * x); J[:, Y, 9] = a * (f32(3) * y2 - f32(3) * x2)
        # 10: b*x*y*z
        J[:, X, 10] = b * y * z; J[:, Y, 10] = b * x * z; J[:, Z, 10] = b * x * y
        # 11: c*y*(1 - 5z2)
        J[:, Y, 11] = c * (f32(1) - f32(5) * z2); J[:, Z, 11] = c * y * (f32(-10) * z)
        # 12: d*z*(5z2 - 3)
        J[:, Z, 12] = d * (f32(15) * z2 - f32(3))
        # 13: c*x*(1 - 5z2)
        J[:, X, 13] = c * (f32(1) - f32(5) * z2); J[:, Z, 13] = c * x * (f32(-10) * z)
        # 14: e*z*(x2 - y2)
        J[:, X, 14] = e * z * f32(2) * x; J[:, Y, 14] = e * z * f32(-2) * y; J[:, Z, 14] = e * (x2 - y2)
        # 15: a*x*(3y2 - x2)
        J[:, X, 15] = a * (f32(3) * y2 - f32(3) * x2); J[:, Y, 15] = a * x * f32(6) * y
    if degree > 4:
        # bands 4..7: central differences of the float64 polynomials (step 1e-6: truncation ~1e-11, rounding ~1e-9), exact for
        # the purpose of a float32 comparison; the closed forms above pin the method on bands 0..3
        h = 1e-6
        base = np.asarray(dirs, np.float64)
        for a in range(3):
            e = np.zeros(3); e[a] = h
            J[:, a, 16:] = ((sh_polynomials(base + e, degree, np.float64) - sh_polynomials(base - e, degree, np.float64)) / (2 * h))[:, 16:]
    return J


def sh_encode_backward(grad, jac):
    """shencoder.cu:358-398: grad [B,C2], jac [B,3,C2] -> grad_inputs [B,3], channels summed in order in float32."""
    g = np.asarray(grad, np.float32); j = np.asarray(jac, np.float32)
    out = np.zeros((g.shape[0], 3), np.float32)
    for ch in range(g.shape[1]):
        out = (out + g[:, ch:ch + 1] * j[:, :, ch]).astype(np.float32)
    return out
