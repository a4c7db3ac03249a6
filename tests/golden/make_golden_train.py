"""Generates tests/golden/ref_train.npz: outputs of the REFERENCE's own training-side kernels (raymarching.cu:305-696,
gridencoder.cu:248-645, shencoder.cu:128-438) on seeded inputs.

The reference's CUDA sources are compiled unmodified for sm_100a into oracle/_ref/*.so (oracle/build_ref.py, where
/root/reference exists); this script needs a GPU, so it runs on the B200 box:

    gpurun -- 'python tests/golden/make_golden_train.py gpurun_out/ref_train.npz'

and the file it writes is committed as tests/golden/ref_train.npz.  Inputs are stored beside the outputs, so the
consumers (tests/test_train_oracle.py on CPU, tests/test_gpu_training.py on the GPU) need neither the reference nor this
script.  march_rays_train packs rays in the order its atomics retire; the fixture stores them re-packed in ray order
(repack_by_ray), which is what the oracle and the CUDA library produce.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import render_oracle as ro  # noqa: E402
from oracle.build_ref import load_ref  # noqa: E402
from pienerf_b200.synthetic import grid_offsets  # noqa: E402
from tests.util import repack_by_ray, small_scene  # noqa: E402

f32 = np.float32


def march_case(bound, dt_gamma, max_steps, seed, W=24, H=24):
    body, field, bits, pose, intr = small_scene(W=W, H=H, bound=bound, seed=seed)
    o, d = ro.get_rays(pose, intr, H, W)
    aabb = np.array([-bound] * 3 + [bound] * 3, f32)
    nears, fars = ro.near_far_from_aabb(o, d, aabb, 0.2)
    noises = np.random.default_rng(seed + 100).uniform(0, 1, o.shape[0]).astype(f32)
    C = 1 if bound <= 1 else 2
    return dict(o=o, d=d, nears=nears, fars=fars, noises=noises, bits=bits,
                par=np.array([bound, dt_gamma, max_steps, C, 128], np.float64))


def grid_case(seed, D, C, gridtype, align, interp, L, B=400):
    rng = np.random.default_rng(seed)
    off, s = grid_offsets(input_dim=D, num_levels=L, base_resolution=4, log2_hashmap_size=10, desired_resolution=64, align_corners=align)
    emb = rng.uniform(-1, 1, size=(int(off[-1]), C)).astype(f32)
    x = rng.uniform(-0.02, 1.02, size=(B, D)).astype(f32)
    x[0] = 0; x[1] = 1; x[2] = 0.5
    grad = rng.normal(size=(L, B, C)).astype(f32)
    return dict(x=x, emb=emb, off=off, grad=grad, par=np.array([np.log2(s), 4, D, C, gridtype, int(align), interp, L], np.float64))


def main(out_path):
    import torch
    rm = load_ref("_ref_raymarching"); ge = load_ref("_ref_gridencoder"); se = load_ref("_ref_shencoder")
    assert rm is not None and ge is not None and se is not None, "oracle/_ref/*.so missing (python oracle/build_ref.py where /root/reference exists)"
    g = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()  # noqa: E731
    out = {}

    # ---- march_rays_train + composite_rays_train, two configurations
    for tag, case in (("mA", march_case(1.0, 0.0, 256, 0)), ("mB", march_case(2.0, 1.0 / 128, 128, 1))):
        bound, dt_gamma, max_steps, C, H = case["par"]
        N = case["o"].shape[0]; M = N * 64
        xyzs = torch.zeros(M, 3, device="cuda"); dirs = torch.zeros(M, 3, device="cuda"); deltas = torch.zeros(M, 2, device="cuda")
        rays = torch.empty(N, 3, dtype=torch.int32, device="cuda"); counter = torch.zeros(2, dtype=torch.int32, device="cuda")
        rm.march_rays_train(g(case["o"]), g(case["d"]), g(case["bits"]), float(bound), float(dt_gamma), int(max_steps), N, int(C), int(H), M,
                            g(case["nears"]), g(case["fars"]), xyzs, dirs, deltas, rays, counter, g(case["noises"]))
        torch.cuda.synchronize()
        assert int(counter[0]) <= M and int(counter[1]) == N
        X, Dd, Dl, R = repack_by_ray(xyzs.cpu().numpy(), dirs.cpu().numpy(), deltas.cpu().numpy(), rays.cpu().numpy())
        for k, v in case.items():
            out[f"{tag}_{k}"] = v
        out[f"{tag}_xyzs"] = X; out[f"{tag}_dirs"] = Dd; out[f"{tag}_deltas"] = Dl; out[f"{tag}_rays"] = R
        m = X.shape[0]
        rng = np.random.default_rng(7)
        sig = rng.uniform(0, 40, m).astype(f32); rgb = rng.uniform(0, 1, (m, 3)).astype(f32)
        ws = torch.empty(N, device="cuda"); dep = torch.empty(N, device="cuda"); img = torch.empty(N, 3, device="cuda")
        rm.composite_rays_train_forward(g(sig), g(rgb), g(Dl), g(R), m, N, 1e-2, ws, dep, img)
        gws = rng.normal(size=N).astype(f32); gim = rng.normal(size=(N, 3)).astype(f32)
        gs = torch.zeros(m, device="cuda"); gc = torch.zeros(m, 3, device="cuda")
        rm.composite_rays_train_backward(g(gws), g(gim), g(sig), g(rgb), g(Dl), g(R), ws, img, m, N, 1e-2, gs, gc)
        torch.cuda.synchronize()
        out.update({f"{tag}_sig": sig, f"{tag}_rgb": rgb, f"{tag}_ws": ws.cpu().numpy(), f"{tag}_depth": dep.cpu().numpy(),
                    f"{tag}_image": img.cpu().numpy(), f"{tag}_gws": gws, f"{tag}_gim": gim, f"{tag}_gs": gs.cpu().numpy(),
                    f"{tag}_gc": gc.cpu().numpy()})

    # ---- grid backward / input backward / total variation, two configurations
    for tag, case in (("gA", grid_case(0, 3, 2, 0, False, 0, 5)), ("gB", grid_case(1, 2, 4, 1, True, 1, 4))):
        S, H, D, C, gridtype, align, interp, L = case["par"]
        D, C, gridtype, interp, L, H = int(D), int(C), int(gridtype), int(interp), int(L), int(H); align = bool(align)
        B = case["x"].shape[0]
        x = g(case["x"]); emb = g(case["emb"]); off = g(case["off"]); grad = g(case["grad"])
        outputs = torch.empty(L, B, C, device="cuda"); dy_dx = torch.empty(B, L * D * C, device="cuda")
        ge.grid_encode_forward(x, emb, off, outputs, B, D, C, L, float(S), H, dy_dx, gridtype, align, interp)
        gemb = torch.zeros_like(emb); gin = torch.zeros(B, D, device="cuda")
        ge.grid_encode_backward(grad, x, emb, off, gemb, B, D, C, L, float(S), H, dy_dx, gin, gridtype, align, interp)
        tv = torch.zeros_like(emb)
        ge.grad_total_variation(x, emb, tv, off, 1e-2, B, D, C, L, float(S), H, gridtype, align)
        torch.cuda.synchronize()
        for k, v in case.items():
            out[f"{tag}_{k}"] = v
        out.update({f"{tag}_out": outputs.cpu().numpy(), f"{tag}_dy_dx": dy_dx.cpu().numpy(), f"{tag}_gemb": gemb.cpu().numpy(),
                    f"{tag}_gin": gin.cpu().numpy(), f"{tag}_tv": tv.cpu().numpy()})

    # ---- SH Jacobian + backward, degrees 4 (oracle + CUDA) and 8 (CUDA)
    rng = np.random.default_rng(3)
    dirs = rng.normal(size=(200, 3)); dirs /= np.linalg.norm(dirs, axis=1, keepdims=True); dirs = dirs.astype(f32)
    out["sh_dirs"] = dirs
    for deg in (4, 8):
        y = torch.empty(200, deg * deg, device="cuda"); j = torch.empty(200, 3 * deg * deg, device="cuda")
        se.sh_encode_forward(g(dirs), y, 200, 3, deg, j)
        grad = rng.normal(size=(200, deg * deg)).astype(f32)
        gin = torch.zeros(200, 3, device="cuda")
        se.sh_encode_backward(g(grad), g(dirs), 200, 3, deg, j, gin)
        torch.cuda.synchronize()
        out.update({f"sh{deg}_y": y.cpu().numpy(), f"sh{deg}_dy_dx": j.cpu().numpy(), f"sh{deg}_grad": grad, f"sh{deg}_gin": gin.cpu().numpy()})

    os.makedirs(os.path.dirname(os.path.abspath(out_path)), exist_ok=True)
    np.savez_compressed(out_path, **out)
    print("wrote", out_path, os.path.getsize(out_path), "bytes,", len(out), "arrays")


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "ref_train.npz"))
