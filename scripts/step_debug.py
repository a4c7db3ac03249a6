"""cluster step vs multi-kernel step: per-step max |diff| and the cluster kernel's per-phase cycle counters."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from pienerf_b200 import _qgmls
from pienerf_b200.simulator import Simulator
from pienerf_b200.synthetic import make_body
kind = sys.argv[1] if len(sys.argv) > 1 else "chair2k"
b = make_body(kind)
def mk(multi):
    _qgmls.step_mode(multi)
    s = Simulator(dt=1e-2, iters=10, bbox=torch.tensor([2.0, 2.0, 2.0]), dx=0.05, stiff=1e5, base=torch.tensor([-1.0, -1.0, -1.0]), use_graph=False)
    s.set_points(b["pos"], b["mass"], b["mu"], b["lam"], b["pin"]).initialize()
    return s
a = mk(True)
c = mk(False)
for i in range(4):
    _qgmls.step_mode(True); a.stepforward()
    _qgmls.step_mode(False); c.stepforward()
    torch.cuda.synchronize()
    d = (a.dof - c.dof).abs().max().item(); dv = (a.dof_vel - c.dof_vel).abs().max().item()
    print(f"step {i}: max|ddof| {d:.3e} (scale {a.dof.abs().max().item():.2f}) max|dvel| {dv:.3e} (scale {a.dof_vel.abs().max().item():.3e})")
tail = 16 * 30 * c.n_k
prof = c._scratch[-8 - tail:-tail].view(torch.int64).tolist()
names = ["init+CSR+momentum", "F (8 lanes/IP)", "SVD+stress", "cluster barriers", "rhs gather", "final+matvec", "dof reload", "-"]
tot = sum(prof)
for n, v in zip(names, prof):
    print(f"  {n:18s} {v:10d} cycles  {100.0 * v / tot:5.1f} %")
print("total cycles", tot, "~", tot / 1.9e3, "us at 1.9 GHz")
