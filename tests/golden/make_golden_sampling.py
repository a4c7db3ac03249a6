"""Golden vectors for point sampling, produced by the reference's OWN source.

    python tests/golden/make_golden_sampling.py            (CPU only, ~1 min of Python loops)

Imports the UNMODIFIED /root/reference/main_sample.py (and through it nerf/utils.py's get_pnts_in_grids) on top of the numpy
`warp` stand-in of warp_shim.py, with permissive stubs for the third-party modules nerf/utils.py imports at module scope
but sampling never calls, and runs AdaptiveUniformSampling.sample() on an analytic density field.  Three things are
replaced from outside, none in the reference's files: `write_ply` (plyfile is not installed) records its arguments;
`os.mkdir` of /root/reference/model (read-only here) is skipped; out-of-range array accesses follow warp_shim.OOB_TOLERANT.
Output: tests/golden/ref_sampling.npz (inputs: the options and the seed; outputs: points and volumes of each case).
"""
import argparse
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
                                "torchmetrics", "torchmetrics.functional", "tqdm", "pandas", "rich", "rich.console", "dearpygui",
                                "dearpygui.dearpygui", "scipy", "scipy.spatial", "scipy.spatial.transform"])
warp_shim.OOB_TOLERANT = True
import torch  # noqa: E402

from tests.sampling_cases import CASES, BlobField  # noqa: E402


def main():
    sys.argv = sys.argv[:1]
    import main_sample as ms                         # the reference's script, unmodified
    captured = {}
    ms.write_ply = lambda filename, points, volumes, binary=True: captured.update(points=np.asarray(points), volumes=np.asarray(volumes))
    real_mkdir = os.mkdir
    os.mkdir = lambda p, *a, **k: None if str(p).startswith("/root/reference") else real_mkdir(p, *a, **k)
    out = {}
    for tag, o in CASES.items():
        opt = argparse.Namespace(**o)
        torch.manual_seed(int(o["seed"]))
        ms.AdaptiveUniformSampling(opt, BlobField()).sample()
        out[f"{tag}_points"] = captured["points"].astype(np.float32)
        out[f"{tag}_volumes"] = captured["volumes"].astype(np.float32)
        print(tag, captured["points"].shape, float(captured["volumes"].min()), float(captured["volumes"].max()), flush=True)
    np.savez_compressed(os.path.join(HERE, "ref_sampling.npz"), **out)


if __name__ == "__main__":
    main()
