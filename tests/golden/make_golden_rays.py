"""Golden pixel selections of the reference's get_rays (nerf/utils.py:55-138) for training batches (N > 0): uniform, patch-based
and error-map driven, on the CPU under fixed torch seeds.  Imports the UNMODIFIED /root/reference/nerf/utils.py with the numpy
`warp` stand-in registered (its kernels are only decorated at import, never launched here) and permissive stubs for the
third-party modules it imports at module scope.  Also stores the reference's rays for those pixels (float32 torch arithmetic).
Output: tests/golden/ref_rays.npz.      python tests/golden/make_golden_rays.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)

import warp_shim  # noqa: E402

warp_shim.install("/root/reference")
warp_shim.stub_missing_modules(["imageio", "tensorboardX", "cv2", "matplotlib", "matplotlib.pyplot", "trimesh", "mcubes", "torch_ema", "lpips",
                                "torchmetrics", "torchmetrics.functional", "tqdm", "pandas", "rich", "rich.console"])
import torch  # noqa: E402
from nerf.utils import get_rays  # noqa: E402  (the reference's function, unmodified)

from pienerf_b200.synthetic import orbit_intrinsics, orbit_pose  # noqa: E402

CASES = {"uniform": dict(H=60, W=80, N=500, seed=1), "patch": dict(H=60, W=80, N=512, patch_size=8, seed=2),
         "errmap": dict(H=300, W=200, N=700, seed=3, error_map=True), "clip": dict(H=12, W=10, N=1000, seed=4)}


def main():
    out = {}
    pose = torch.from_numpy(orbit_pose(radius=2.5).astype(np.float32))[None]
    for tag, c in CASES.items():
        intr = orbit_intrinsics(c["W"], c["H"], 50.0)
        torch.manual_seed(c["seed"])
        em = None
        if c.get("error_map"):
            em = torch.rand(1, 128 * 128)
            out[f"{tag}_error_map"] = em.numpy()
        r = get_rays(pose, intr, c["H"], c["W"], c["N"], em, c.get("patch_size", 1))
        out[f"{tag}_inds"] = r["inds"].numpy()
        if "inds_coarse" in r:
            out[f"{tag}_inds_coarse"] = r["inds_coarse"].numpy()
        out[f"{tag}_rays_o"] = r["rays_o"].numpy(); out[f"{tag}_rays_d"] = r["rays_d"].numpy()
        print(tag, r["inds"].shape, r["rays_d"].shape)
    out["pose"] = pose.numpy()
    np.savez_compressed(os.path.join(HERE, "ref_rays.npz"), **out)


if __name__ == "__main__":
    main()
