"""Tiny driver for ncu captures: builds the chair scene and runs a few simulated+rendered frames.
usage: python scripts/profile_frame.py [frames] [density_scale] [mode]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from pienerf_b200.frame import FrameDriver, build_scene  # noqa: E402

frames = int(sys.argv[1]) if len(sys.argv) > 1 else 3
ds = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
mode = int(sys.argv[3]) if len(sys.argv) > 3 else 0
model, sim, opt, pose, intr, body, field = build_scene("chair", density_scale=ds)
drv = FrameDriver(model, sim, opt, fused=True)
opt["mode"] = mode
for i in range(frames):
    out = drv.test_gui(pose, intr, opt.W, opt.H, to_host=False)
torch.cuda.synchronize()
print("stats", model._stats.tolist())
