"""Times the training-side kernels against the reference's own kernels (oracle/_ref/*.so) on the same
inputs, CUDA events on the current stream, after warm-up.  Not part of bench.py (the headline is the per-frame loop);
the output goes to profiles/.  Usage: python scripts/time_training.py [out.json]"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.build_ref import load_ref  # noqa: E402  (timing harness: the reference kernels are the baseline arm)
from pienerf_b200 import _gridencoder, _raymarching, _shencoder  # noqa: E402
from pienerf_b200.synthetic import grid_offsets, make_body, occupancy_bitfield, orbit_intrinsics, orbit_pose  # noqa: E402
from pienerf_b200 import raymarching as rm  # noqa: E402


def timed(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


def main(out_path):
    rg = load_ref("_ref_gridencoder"); rr = load_ref("_ref_raymarching"); rs = load_ref("_ref_shencoder")
    res = {}
    g = torch.Generator(device="cuda").manual_seed(0)
    if "PN_TRAIN_WRITE" in os.environ:                   # A/B: 0 = stream stores, 1 = warp-cooperative flush
        from pienerf_b200._lib import lib
        lib.pn_set_train_write_mode(int(os.environ["PN_TRAIN_WRITE"]))
    if "PN_TRAIN_SKIP" in os.environ:                    # A/B: empty-space block skipping of march_rays_train
        from pienerf_b200._lib import lib
        lib.pn_set_train_block_skip(int(os.environ["PN_TRAIN_SKIP"]))
    only_rays = os.environ.get("PN_TIME_ONLY", "") in ("rays", "rays_timed")
    if only_rays:
        rg = rs = None
    if os.environ.get("PN_TIME_ONLY") == "rays":        # for an ncu launch list of the ray kernels alone: a few untimed calls
        global timed
        timed = lambda fn, iters=2, warm=1: [fn() for _ in range(iters + warm)] and 0.0  # noqa: E731
    offsets, pls = grid_offsets()
    emb = torch.rand(int(offsets[-1]), 2, device="cuda", generator=g) * 2 - 1
    off = torch.from_numpy(offsets).cuda(); S = float(np.log2(pls))
    B = 1 << 20
    x = torch.rand(B, 3, device="cuda", generator=g); grad = torch.randn(16, B, 2, device="cuda", generator=g)
    gemb = torch.zeros_like(emb)
    for name, m in (("ours", None if only_rays else _gridencoder), ("reference", rg)):
        if m is None:
            continue
        for dt, tag in ((torch.float32, "f32"), (torch.float16, "f16")):
            e = emb.to(dt); gr = grad.to(dt) * (0.01 if dt == torch.float16 else 1.0); ge = gemb.to(dt)
            res[f"grid_backward_{tag}_B2^20_L16_ms/{name}"] = timed(lambda: m.grid_encode_backward(gr, x, e, off, ge, B, 3, 2, 16, S, 16, None, None, 0, False, 0))
        tv = torch.zeros_like(emb)
        res[f"grad_tv_B2^20_L16_ms/{name}"] = timed(lambda: m.grad_total_variation(x, emb, tv, off, 1e-7, B, 3, 2, 16, S, 16, 0, False))
    # algorithmic traffic of the table gradient: per (sample, level) 12 B input share + 8 B grad + 8 vertices x 8 B reductions
    for name in ("ours", "reference"):
        k = f"grid_backward_f32_B2^20_L16_ms/{name}"
        if k in res:
            res[k.replace("_ms/", "_GB_per_s/")] = B * 16 * (8 + 8 * 8 + 12 / 16) / (res[k] * 1e-3) / 1e9

    # march_rays_train + composite on a full 800x800 frame of the chair-like body
    body = make_body("chair2k", dx=0.05, bound=1.0, seed=0)
    bits = torch.from_numpy(occupancy_bitfield(body["pos"], 0.03, bound=1.0)).cuda()
    W = H = 800
    rays = rm.get_rays(torch.from_numpy(orbit_pose(radius=2.5).astype(np.float32))[None], orbit_intrinsics(W, H, 50.0), H, W)
    o = rays["rays_o"][0].contiguous(); d = rays["rays_d"][0].contiguous(); N = o.shape[0]
    nears, fars = rm.near_far_from_aabb(o, d, torch.tensor([-1.0, -1, -1, 1, 1, 1], device="cuda"), 0.2)
    noises = torch.rand(N, device="cuda", generator=g)
    M = N * 16
    xyzs = torch.zeros(M, 3, device="cuda"); dirs = torch.zeros(M, 3, device="cuda"); deltas = torch.zeros(M, 2, device="cuda")
    rt = torch.empty(N, 3, dtype=torch.int32, device="cuda"); counter = torch.zeros(2, dtype=torch.int32, device="cuda")
    for name, m in (("ours", _raymarching), ("reference", rr)):
        if m is None:
            continue
        def march():
            counter.zero_()
            m.march_rays_train(o, d, bits, 1.0, 0.0, 1024, N, 1, 128, M, nears, fars, xyzs, dirs, deltas, rt, counter, noises)
        res[f"march_rays_train_800x800_ms/{name}"] = timed(march)
        res[f"march_rays_train_samples/{name}"] = int(counter[0])
        sig = xyzs.abs().sum(-1) * 10; rgb = torch.sigmoid(xyzs)
        ws = torch.empty(N, device="cuda"); dep = torch.empty(N, device="cuda"); img = torch.empty(N, 3, device="cuda")
        res[f"composite_train_fwd_ms/{name}"] = timed(lambda: m.composite_rays_train_forward(sig, rgb, deltas, rt, M, N, 1e-4, ws, dep, img))
        gs = torch.zeros(M, device="cuda"); gc = torch.zeros(M, 3, device="cuda"); one = torch.ones(N, device="cuda"); one3 = torch.ones(N, 3, device="cuda")
        res[f"composite_train_bwd_ms/{name}"] = timed(lambda: m.composite_rays_train_backward(one, one3, sig, rgb, deltas, rt, ws, img, M, N, 1e-4, gs, gc))
    dn = torch.nn.functional.normalize(torch.randn(B, 3, device="cuda", generator=g), dim=-1)
    y = torch.empty(B, 16, device="cuda"); j = torch.empty(B, 48, device="cuda")
    for name, m in (("ours", None if only_rays else _shencoder), ("reference", rs)):
        if m is None:
            continue
        res[f"sh_forward_with_jacobian_deg4_B2^20_ms/{name}"] = timed(lambda: m.sh_encode_forward(dn, y, B, 3, 4, j))
    print(json.dumps(res, indent=1))
    if out_path:
        os.makedirs(os.path.dirname(os.path.abspath(out_path)), exist_ok=True)
        json.dump(res, open(out_path, "w"), indent=1)


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else None)
