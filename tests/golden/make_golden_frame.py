"""Generates tests/golden/ref_frame.npz: whole frames rendered by the REFERENCE GPU renderer — the reference's own CUDA
kernels (oracle/_ref/*.so, compiled unmodified for sm_100a) driven by its own rund_cuda loop with the fp32 nn.Linear
MLP (oracle/ref_renderer.py) — on the small seeded scene of tests/util.py.  Run on the GPU box:

    gpurun -- 'python tests/golden/make_golden_frame.py gpurun_out/ref_frame.npz'

and copy the file to tests/golden/.  tests/test_golden.py then pins the numpy oracle's rund_cuda (the checker of every
frame-level parity test) to the reference at frame level, on CPU, without the reference tree."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import render_oracle as ro  # noqa: E402
from oracle.ref_renderer import ReferenceRenderer  # noqa: E402
from tests.util import deformed_ip_state, small_scene  # noqa: E402

out_path = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "ref_frame.npz")
W = H = 40
body, field, bits, pose, intr = small_scene(W=W, H=H, seed=0)
rays_o, rays_d = ro.get_rays(pose, intr, H, W)
g = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()  # noqa: E731
G = {"W": np.int32(W), "H": np.int32(H)}
for tag, amp, K, ds in (("undeformed_K3", 0.0, 3, 20.0), ("deformed_K3", 0.03, 3, 20.0), ("deformed_K1", 0.03, 1, 5.0)):
    p_ori, p_def, F, dF = deformed_ip_state(body, seed=0, amp=amp)
    ref = ReferenceRenderer(field, bits, bound=1.0, density_scale=ds, min_near=0.2)
    out = ref.rund_cuda(g(rays_o), g(rays_d), g(p_def), g(p_ori), g(F), g(dF), 0.0525, dt_gamma=0.0, max_steps=256, T_thresh=1e-2,
                        max_iter_num=1, hash_grid_size=0.06, num_seek_IP=K, return_stats=True)
    G[f"{tag}_image"] = out["image"].cpu().numpy(); G[f"{tag}_weights_sum"] = out["weights_sum"].cpu().numpy()
    G[f"{tag}_depth_0"] = out["depth_0"].cpu().numpy(); G[f"{tag}_n_samples"] = np.int64(out["n_samples"])
    G[f"{tag}_cfg"] = np.array([amp, K, ds], dtype=np.float64)
    print(tag, "samples", out["n_samples"], "hit pixels", int((out["weights_sum"] > 0).sum()))
np.savez_compressed(out_path, **G)
print("wrote", out_path, os.path.getsize(out_path), "bytes")
