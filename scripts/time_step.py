"""Simulator step timing: one-kernel cluster step vs the multi-kernel CUDA graph, per body.  usage: python scripts/time_step.py [kinds...]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

from pienerf_b200 import _qgmls
from pienerf_b200.simulator import Simulator
from pienerf_b200.synthetic import make_body

for kind in (sys.argv[1:] or ["chair2k", "block4k", "chairlike"]):
    b = make_body(kind)
    for multi in (True, False):
        _qgmls.step_mode(multi)
        s = Simulator(dt=1e-2, iters=10, bbox=torch.tensor([2.0, 2.0, 2.0]), dx=0.05, stiff=1e5, base=torch.tensor([-1.0, -1.0, -1.0]))
        s.set_points(b["pos"], b["mass"], b["mu"], b["lam"], b["pin"]).initialize()
        for _ in range(5):
            s.stepforward()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(50):
            s.stepforward()
        e1.record(); torch.cuda.synchronize()
        print(f"{kind}: n_ip {s.n_ip} n_k {s.n_k} n {s.n}  {'multi-kernel graph' if multi else 'cluster kernel    '} launches {s.step_launches:3d}  {e0.elapsed_time(e1) / 50 * 1e3:8.1f} us/step", flush=True)
_qgmls.step_mode(True)
