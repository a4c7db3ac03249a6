"""Render-only timing of the chair frame per render mode (CUDA events, L2 flushed between frames).
usage: python scripts/mode_compare.py [modes, e.g. 0,3] [density_scale] [config]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from pienerf_b200.frame import FrameDriver, build_scene  # noqa: E402

modes = [int(m) for m in (sys.argv[1] if len(sys.argv) > 1 else "0,3").split(",")]
ds = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
config = sys.argv[3] if len(sys.argv) > 3 else "chair"
model, sim, opt, pose, intr, body, field = build_scene(config, density_scale=ds)
drv = FrameDriver(model, sim, opt, fused=True)
for _ in range(3):
    sim.stepforward()
pos, F, dF = sim.get_IP_info()
model.p_def, model.IP_F, model.IP_dF = pos, F, dF
rays, rH, rW = drv.rays(pose, intr, opt.W, opt.H)
flush = torch.empty(64 * 1024 * 1024, device="cuda")
ref = None
for mode in modes:
    ts = []
    for i in range(8):
        if i == 5:
            torch.cuda.profiler.start()                          # ncu --profile-from-start off: frames 5..7 only
        flush.fill_(1.0)
        a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        a.record()
        out = model.render_deformed(rays["rays_o"], rays["rays_d"], mode=mode, **opt)
        b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    torch.cuda.profiler.stop()
    img = out["image"].clone()
    if ref is None:
        ref = img
    print(f"mode {mode} ds {ds} {config}: {np.mean(ts[3:]):.3f} ms/frame (min {min(ts):.3f})  stats {model._stats.tolist()}  max|img - mode{modes[0]}| {float((img - ref).abs().max()):.2e}  finite {bool(torch.isfinite(img).all())}")
