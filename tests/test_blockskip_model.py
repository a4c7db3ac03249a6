"""A scalar float32 model of the training walk of pienerf_b200/csrc/train_rays.cu (walk_ray_impl with empty-space block skipping:
empty_block / block_exit), run against the reference-semantics oracle (oracle/train_oracle.py::march_rays_train) and the
reference's own outputs (tests/golden/ref_train.npz).  CPU only.  It checks the ALGORITHM the kernel uses — skipping an aligned
empty 8^3 / 4^3 block of voxels to the first lattice point past the block's exit unless a lattice point lies in the rounding
interval around that exit — reproduces the per-voxel walk sample for sample; the kernel itself is compared with the reference's
kernel on 640 k rays in tests/test_gpu_training.py."""
import os

import numpy as np
import pytest

from oracle import render_oracle as ro

f32 = np.float32
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_train.npz")


def _bit(bits, index):
    return (int(bits[index >> 3]) >> (index & 7)) & 1


def _morton(nx, ny, nz):
    return int(ro.morton3D(np.array([nx]), np.array([ny]), np.array([nz]))[0])


def walk(o, d, near, far, noise, bits, bound, max_steps, H=128, skip=True, stats=None):
    """One ray, one cascade, dt_gamma = 0; returns the list of (x, y, z, dt, t_after - last_t) and counts skips in `stats`."""
    o = o.astype(f32); d = d.astype(f32)
    with np.errstate(divide="ignore"):
        rd = (f32(1) / d).astype(f32)
    dt = f32(f32(2) * ro.SQRT3 / f32(max_steps))                       # the fixed step: clamp(t * 0, dt_min, dt_max)
    bound = f32(bound); rH = f32(1) / f32(H)
    t = f32(near + dt * noise); last_t = t
    out = []
    sgn = np.where(np.signbit(d), f32(-1), f32(1)).astype(f32)

    def exits(n_planes, pos, t):
        with np.errstate(invalid="ignore", over="ignore"):
            return (((n_planes * rH * f32(2) - f32(1)) * min(f32(1), bound) - pos) * rd).astype(f32)
    while t < far and len(out) < max_steps:
        pos = np.minimum(bound, np.maximum(-bound, (o + t * d).astype(f32))).astype(f32)
        mb = min(f32(1), bound)
        n = np.minimum(f32(H - 1), np.maximum(f32(0), (f32(0.5) * (pos * (f32(1) / mb) + f32(1)) * f32(H)).astype(f32))).astype(np.int64)
        index = _morton(*n)
        if _bit(bits, index):
            t_new = f32(t + dt)
            out.append((pos[0], pos[1], pos[2], dt, f32(t_new - last_t)))
            last_t = t = t_new
            continue
        if skip:
            base = index & ~511
            block = bits[base >> 3:(base >> 3) + 64]
            lb = 3 if not block.any() else (2 if not block[((index >> 6) & 7) * 8:((index >> 6) & 7) * 8 + 8].any() else 0)
            if lb:
                size = 1 << lb
                planes = ((n // size) * size + np.where(sgn > 0, size, 0)).astype(f32)
                ta = exits(planes, pos, t)
                et = f32(f32(4e-6) * (f32(1) + t))
                with np.errstate(invalid="ignore", over="ignore"):
                    e = (f32(4e-6) * np.abs(rd) + et).astype(f32)
                    lo = f32(t + max(f32(0), np.fmin.reduce((ta - e).astype(f32))))
                    hi = f32(f32(t + max(f32(0), np.fmin.reduce((ta + e).astype(f32)))) + et)
                stop = min(lo, f32(far))
                u = t
                while True:
                    u = f32(u + dt)
                    if not u < stop:
                        break
                if u >= hi or u >= far:
                    if stats is not None:
                        stats["skips"] = stats.get("skips", 0) + 1
                    t = u
                    continue
                if stats is not None:
                    stats["fallbacks"] = stats.get("fallbacks", 0) + 1
        ta = exits((n.astype(f32) + f32(0.5) + f32(0.5) * sgn).astype(f32), pos, t)
        with np.errstate(invalid="ignore"):
            tt = f32(t + max(f32(0), np.fmin.reduce(ta)))
        while True:
            t = f32(t + dt)
            if not t < tt:
                break
    return out


def test_block_skipping_reproduces_the_per_voxel_walk():
    if not os.path.exists(GOLD):
        pytest.skip("tests/golden/ref_train.npz not generated yet")
    G = np.load(GOLD)
    bound, dt_gamma, max_steps, C, H = G["mA_par"]
    assert dt_gamma == 0 and C == 1
    o, d, nears, fars, noises, bits = (G[f"mA_{k}"] for k in ("o", "d", "nears", "fars", "noises", "bits"))
    R = G["mA_rays"]
    hit = np.nonzero(R[:, 2] > 0)[0][::4]; miss = np.nonzero(R[:, 2] == 0)[0][::8]     # a quarter of the hits, an eighth of the misses
    stats = {}
    agree = 0
    for n in np.concatenate([hit, miss]):
        a = walk(o[n], d[n], nears[n], fars[n], noises[n], bits, bound, int(max_steps), int(H), skip=True, stats=stats)
        b = walk(o[n], d[n], nears[n], fars[n], noises[n], bits, bound, int(max_steps), int(H), skip=False)
        assert len(a) == len(b) and all(np.array_equal(np.array(p, f32), np.array(q, f32)) for p, q in zip(a, b)), int(n)
        # ... and the per-voxel model is the reference: same count as the reference's kernel produced, same samples to fp32 rounding
        if len(b) == R[n, 2]:
            agree += 1
            if len(b):
                ref = np.concatenate([G["mA_xyzs"][R[n, 1]:R[n, 1] + R[n, 2]], G["mA_deltas"][R[n, 1]:R[n, 1] + R[n, 2]]], axis=1)
                assert np.abs(np.array(b, f32) - ref).max() <= 2e-6
    assert agree >= 0.99 * (len(hit) + len(miss))                        # knife-edge occupancy under FMA contraction only
    assert stats.get("skips", 0) > 20 * stats.get("fallbacks", 0) and stats["skips"] > 500, stats


def test_block_skipping_on_grazing_and_axis_parallel_rays(rng):
    """Rays the interval logic must be careful with: parallel to an axis (an infinite 1/d: that axis is ignored), nearly parallel
    (huge 1/d: wide intervals, per-voxel fallback), starting on voxel planes.  Random sparse occupancy."""
    H = 128
    bits = np.zeros(H ** 3 // 8, np.uint8)
    occ = rng.integers(0, H ** 3, size=3000)
    np.bitwise_or.at(bits, occ >> 3, (1 << (occ & 7)).astype(np.uint8))
    rays = []
    for k in range(24):
        o = np.array([-1.0, rng.uniform(-0.9, 0.9), rng.uniform(-0.9, 0.9)], f32)
        d = np.array([1.0, 0.0, 0.0], f32)
        if k % 3 == 1:
            d = np.array([1.0, rng.uniform(-1e-7, 1e-7), rng.uniform(-1e-4, 1e-4)], f32)
        if k % 3 == 2:
            o[1] = f32(np.round(o[1] * 64) / 64); o[2] = f32(np.round(o[2] * 64) / 64)      # exactly on voxel planes
            d = np.array([1.0, 1e-3, -2e-3], f32)
        d = (d / np.linalg.norm(d)).astype(f32)
        rays.append((o, d))
    stats = {}
    for o, d in rays:
        a = walk(o, d, f32(0.0), f32(2.0), f32(0.37), bits, 1.0, 256, H, skip=True, stats=stats)
        b = walk(o, d, f32(0.0), f32(2.0), f32(0.37), bits, 1.0, 256, H, skip=False)
        assert len(a) == len(b) and all(np.array_equal(np.array(p, f32), np.array(q, f32)) for p, q in zip(a, b))
    assert stats.get("skips", 0) > 100
