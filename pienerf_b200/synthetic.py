"""Synthetic stand-ins for the reference's Google-Drive assets.

Bodies follow the reference ply vertex schema x,y,z,vp,pin,lam,mu,mass
(simulator/solver.py:115-135; README.md:106-108 for mu=lam=1e6, mass=1e3*vp).
Everything here is numpy-on-host and deterministic; no GPU needed.
"""
import math

import numpy as np


def sim_lattice(bound=1.0, dx=0.05):
    """(base[3] f64, res) of the simulator lattice exactly as main_gui.py:39-46 + solver.py:24-32
    produce them: float32 tensors scaled in place by 1.02 / 1.01, then widened to fp64."""
    bbox = np.float32(2.0 * bound) * np.float32(1.02)
    base = np.float32(-bound) * np.float32(1.01)
    res = int(np.floor(np.float64(bbox) / dx))
    return np.full(3, np.float64(base)), res


def make_body(kind="block512", dx=0.05, bound=1.0, seed=0, mu=1e6, lam=1e6, density=1e3, jitter=0.3):
    """Return dict(pos[n,3] f64, mass, mu, lam, pin, ...) with ONE sample point per occupied
    simulator cell (centre + U(-jitter,jitter)*dx), so n_IP == n_pts.

    kinds: block64 (4^3), block512 (8^3), chair2k (13x13x12), block4k (16^3), block8k (20^3),
    chairlike (seat slab 28x2x28 + four 2x12x2 legs + 28x10x1 back, ~2k IPs, spans +-0.7).
    The top y-layer is pinned; mu=lam=1e6, mass=1e3*vp (README.md:106-108)."""
    rng = np.random.default_rng(seed)
    base, res = sim_lattice(bound, dx)
    if kind == "chairlike":
        cells = set()
        for i in range(-14, 14):
            for k in range(-14, 14):
                for j in (0, 1):
                    cells.add((i, j, k))
        for (ci, ck) in ((-14, -14), (12, -14), (-14, 12), (12, 12)):
            for i in range(ci, ci + 2):
                for k in range(ck, ck + 2):
                    for j in range(-12, 0):
                        cells.add((i, j, k))
        for i in range(-14, 14):
            for j in range(2, 12):
                cells.add((i, j, -14))
        cells = np.array(sorted(cells), dtype=np.int64)
    else:
        dims = {"block64": (4, 4, 4), "block512": (8, 8, 8), "chair2k": (13, 13, 12), "block4k": (16, 16, 16),
                "block8k": (20, 20, 20)}[kind]
        cells = np.stack(np.meshgrid(*[np.arange(d) - d // 2 for d in dims], indexing="ij"), -1).reshape(-1, 3)
    cells = cells + res // 2
    pos = (cells + 0.5 + rng.uniform(-jitter, jitter, size=cells.shape)) * dx + base
    n = pos.shape[0]
    vp = np.full(n, dx ** 3)
    pin = cells[:, 1] == cells[:, 1].max()
    return {
        "pos": pos, "vp": vp, "mass": density * vp, "mu": np.full(n, float(mu)), "lam": np.full(n, float(lam)),
        "pin": pin, "cells": cells, "dx": dx, "bound": bound,
    }


# ----------------------------------------------------------------------------------------------
# field: hash grid + MLP weights + occupancy bitfield
# ----------------------------------------------------------------------------------------------

def grid_offsets(input_dim=3, num_levels=16, base_resolution=16, log2_hashmap_size=19, desired_resolution=2048,
                 align_corners=False):
    """Level offsets table exactly as gridencoder/grid.py:100-128 builds it."""
    per_level_scale = np.exp2(np.log2(desired_resolution / base_resolution) / (num_levels - 1))
    max_params = 2 ** log2_hashmap_size
    offsets, offset = [], 0
    for i in range(num_levels):
        resolution = int(np.ceil(base_resolution * per_level_scale ** i))
        params = min(max_params, (resolution if align_corners else resolution + 1) ** input_dim)
        params = int(np.ceil(params / 8) * 8)
        offsets.append(offset)
        offset += params
    offsets.append(offset)
    return np.array(offsets, dtype=np.int32), float(per_level_scale)


def make_field(bound=1.0, seed=0, emb_scale=1.0, w_scale=2.0, num_levels=16, level_dim=2):
    """Random-init 16-level hash grid + the sigma/colour MLP of nerf/network.py:36-71 (bias-free).

    Embeddings U(-emb_scale, emb_scale) (the reference default 1e-4 renders a constant image,
    SURVEY.md 8d); weights = torch-default kaiming-uniform bound 1/sqrt(fan_in) times w_scale."""
    rng = np.random.default_rng(seed)
    offsets, pls = grid_offsets(num_levels=num_levels, desired_resolution=2048 * bound)
    emb = rng.uniform(-emb_scale, emb_scale, size=(int(offsets[-1]), level_dim)).astype(np.float32)

    def lin(out_d, in_d):
        b = w_scale / math.sqrt(in_d)
        return rng.uniform(-b, b, size=(out_d, in_d)).astype(np.float32)

    in_dim = num_levels * level_dim
    return {
        "bound": float(bound), "offsets": offsets, "per_level_scale": pls, "base_resolution": 16,
        "embeddings": emb, "num_levels": num_levels, "level_dim": level_dim,
        "sigma_net": [lin(64, in_dim), lin(16, 64)],
        "color_net": [lin(64, 31), lin(64, 64), lin(3, 64)],
    }


def _expand_bits(v):
    v = (v * 0x00010001) & 0xFF0000FF
    v = (v * 0x00000101) & 0x0F00F00F
    v = (v * 0x00000011) & 0xC30C30C3
    v = (v * 0x00000005) & 0x49249249
    return v


def morton3d(x, y, z):
    x = np.asarray(x, dtype=np.uint64); y = np.asarray(y, dtype=np.uint64); z = np.asarray(z, dtype=np.uint64)
    return (_expand_bits(x) | (_expand_bits(y) << np.uint64(1)) | (_expand_bits(z) << np.uint64(2))).astype(np.uint32)


def occupancy_bitfield(points, half_extent, bound=1.0, H=128):
    """Density bitfield (layout: raymarching.cu:381-382,773-774; bit = level*H^3 + morton(nx,ny,nz))
    with every cascade cell overlapping a cube of +-half_extent around any point marked occupied."""
    cascade = 1 + math.ceil(math.log2(bound)) if bound > 1 else 1
    bits = np.zeros(cascade * H ** 3 // 8, dtype=np.uint8)
    pts = np.asarray(points, dtype=np.float64)
    for lvl in range(cascade):
        mb = min(2.0 ** lvl, bound)
        lo = np.clip(np.floor(0.5 * ((pts - half_extent) / mb + 1) * H), 0, H - 1).astype(np.int64)
        hi = np.clip(np.floor(0.5 * ((pts + half_extent) / mb + 1) * H), 0, H - 1).astype(np.int64)
        occ = np.zeros((H, H, H), dtype=bool)
        span = int((hi - lo).max()) + 1
        for a in range(span):
            for b in range(span):
                for c in range(span):
                    ix = np.minimum(lo[:, 0] + a, hi[:, 0]); iy = np.minimum(lo[:, 1] + b, hi[:, 1]); iz = np.minimum(lo[:, 2] + c, hi[:, 2])
                    occ[ix, iy, iz] = True
        ix, iy, iz = np.nonzero(occ)
        idx = morton3d(ix, iy, iz).astype(np.int64) + lvl * H ** 3
        np.bitwise_or.at(bits, idx // 8, (1 << (idx % 8)).astype(np.uint8))
    return bits


# ----------------------------------------------------------------------------------------------
# camera: nerf/gui.py:14-44 OrbitCamera pose / intrinsics
# ----------------------------------------------------------------------------------------------

def orbit_pose(radius=2.5, yaw_deg=30.0, pitch_deg=-20.0, center=(0.0, 0.0, 0.0)):
    """4x4 cam2world like OrbitCamera.pose (gui.py:28-38): initial R = diag(1,-1,-1) (quat [1,0,0,0]),
    then orbit about world-up / camera-side."""
    def rot(axis, ang):
        axis = np.asarray(axis, dtype=np.float64); axis = axis / np.linalg.norm(axis)
        K = np.array([[0, -axis[2], axis[1]], [axis[2], 0, -axis[0]], [-axis[1], axis[0], 0]])
        return np.eye(3) + math.sin(ang) * K + (1 - math.cos(ang)) * (K @ K)
    R0 = np.diag([1.0, -1.0, -1.0])
    Ry = rot([0, 1, 0], math.radians(yaw_deg))
    side = (Ry @ R0)[:, 0]
    Rx = rot(side, math.radians(pitch_deg))
    Rm = Rx @ Ry @ R0
    res = np.eye(4, dtype=np.float32)
    res[2, 3] -= radius
    rt = np.eye(4, dtype=np.float32)
    rt[:3, :3] = Rm.astype(np.float32)
    res = rt @ res
    res[:3, 3] -= np.asarray(center, dtype=np.float32)
    return res.astype(np.float32)


def orbit_intrinsics(W, H, fovy_deg=50.0):
    focal = H / (2 * np.tan(np.radians(fovy_deg) / 2))
    return np.array([focal, focal, W // 2, H // 2], dtype=np.float64)


CONFIGS = {
    # BASELINE.json configs -> synthetic stand-ins
    "step512": dict(body="block512", bound=1.0, W=0, H=0),
    "chair": dict(body="chair2k", bound=1.0, W=800, H=800, max_steps=1024, T_thresh=1e-2, dt_gamma=0.0,
                  min_near=0.2, max_iter_num=1, num_seek_IP=3, sim_dx=0.05, radius=2.5, fovy=50.0, cut=False),
    # one GPU's share of the chair frame at 8 GPUs (1/8 of the rays, same body / field): single-GPU stand-in for tuning the
    # coexistence of the simulator and a small render on the simulating rank
    "chair8": dict(body="chair2k", bound=1.0, W=283, H=283, max_steps=1024, T_thresh=1e-2, dt_gamma=0.0,
                   min_near=0.2, max_iter_num=1, num_seek_IP=3, sim_dx=0.05, radius=2.5, fovy=50.0, cut=False),
    "chairlike": dict(body="chairlike", bound=1.0, W=800, H=800, max_steps=1024, T_thresh=1e-2, dt_gamma=0.0,
                      min_near=0.2, max_iter_num=1, num_seek_IP=3, sim_dx=0.05, radius=2.5, fovy=50.0, cut=False),
    "trex": dict(body="block8k", bound=2.0, W=1008, H=756, max_steps=300, T_thresh=5e-2, dt_gamma=1.0 / 128,
                 min_near=0.2, max_iter_num=1, num_seek_IP=1, sim_dx=0.05, radius=5.0, fovy=50.0, cut=True,
                 cut_bounds=[-0.6, 0.6, -0.6, 0.6, -0.6, 0.6]),
    "synth1080": dict(body="block4k", bound=1.0, W=1920, H=1080, max_steps=1024, T_thresh=1e-2, dt_gamma=0.0,
                      min_near=0.2, max_iter_num=1, num_seek_IP=3, sim_dx=0.05, radius=2.5, fovy=50.0, cut=False),
}
