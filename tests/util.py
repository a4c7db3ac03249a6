"""Shared helpers for the parity tests (CPU oracle side + CUDA side on the same seeded inputs)."""
import math
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from pienerf_b200.synthetic import (make_body, make_field, occupancy_bitfield, orbit_intrinsics, orbit_pose,  # noqa: E402
                                    sim_lattice)


def small_scene(kind="block64", bound=1.0, W=48, H=48, seed=0, radius=2.5, emb_scale=1.0):
    """Body + field + bitfield + camera, small enough for the numpy oracle to finish in seconds."""
    body = make_body(kind, dx=0.05, bound=bound, seed=seed)
    field = make_field(bound=bound, seed=seed, emb_scale=emb_scale)
    bits = occupancy_bitfield(body["pos"], 0.03, bound=bound)
    pose = orbit_pose(radius=radius)
    intr = orbit_intrinsics(W, H, 50.0)
    # zoom so the small body fills the frame
    extent = float(np.abs(body["pos"]).max()) + 0.1
    focal = 0.5 * H * radius / extent * 0.8
    intr = np.array([focal, focal, W // 2, H // 2], dtype=np.float64)
    return body, field, bits, pose, intr


def deformed_ip_state(body, seed=0, amp=0.02):
    """A smooth synthetic deformation of the IP set: (p_ori, p_def, F [n,9], dF [n,27]) in renderer layouts
    (F[b][a] at a*3+b; dF[c*9+r*3+j] = d2 phi_r / dX_j dX_c) for phi(X) = X + amp * (A X + 1/2 X^T H X)."""
    rng = np.random.default_rng(seed)
    base, res = sim_lattice(body["bound"], body["dx"])
    p_ori = ((body["cells"] + 0.5) * body["dx"] + base).astype(np.float64)
    A = rng.normal(size=(3, 3)) * 0.3
    Hs = rng.normal(size=(3, 3, 3)) * 0.5
    Hs = 0.5 * (Hs + Hs.transpose(0, 2, 1))                     # symmetric in the two derivative indices
    X = p_ori
    quad = 0.5 * np.einsum("rjc,nj,nc->nr", Hs, X, X)
    p_def = X + amp * (X @ A.T + quad)
    Fm = np.eye(3)[None] + amp * (A[None] + np.einsum("rjc,nc->nrj", Hs, X))      # F[r][j]
    n = X.shape[0]
    F = np.zeros((n, 9)); dF = np.zeros((n, 27))
    for a in range(3):
        for b in range(3):
            F[:, a * 3 + b] = Fm[:, b, a]
    for c in range(3):
        for r in range(3):
            for j in range(3):
                dF[:, c * 9 + r * 3 + j] = amp * Hs[r, j, c]
    return p_ori.astype(np.float32), p_def.astype(np.float32), F.astype(np.float32), dF.astype(np.float32)


def psnr(a, b):
    mse = float(np.mean((a.astype(np.float64) - b.astype(np.float64)) ** 2))
    return 99.0 if mse == 0 else -10 * math.log10(mse)


def repack_by_ray(xyzs, dirs, deltas, rays):
    """march_rays_train outputs re-packed in ray order: rows of `rays` sorted by ray index, sample runs concatenated in that
    order and re-based at 0.  The reference packs in the order its atomics retire (raymarching.cu:407-414); the oracle and the
    CUDA library pack in ray order already, for them this is the identity (up to the trim to the samples produced)."""
    order = np.argsort(rays[:, 0], kind="stable")
    R = np.zeros_like(rays)
    X, D, L = [], [], []
    off = 0
    for row, k in enumerate(order):
        n, o, num = (int(v) for v in rays[k])
        fits = num > 0 and o + num <= xyzs.shape[0]
        R[row] = (n, off, num)
        if fits:
            X.append(xyzs[o:o + num]); D.append(dirs[o:o + num]); L.append(deltas[o:o + num])
        elif num > 0:
            X.append(np.zeros((num, 3), xyzs.dtype)); D.append(np.zeros((num, 3), dirs.dtype)); L.append(np.zeros((num, 2), deltas.dtype))
        off += num
    cat = lambda parts, w: np.concatenate(parts) if parts else np.zeros((0, w), np.float32)  # noqa: E731
    return cat(X, 3), cat(D, 3), cat(L, 2), R
