"""Golden vectors for the simulator, produced by the reference's OWN source.

    python tests/golden/make_golden_sim.py [block64 block512]        (CPU only; block512 takes ~15 min of Python loops)

Imports the UNMODIFIED `/root/reference/simulator/{func_utils,cpu_utils,cuda_utils,solver}.py` on top of the numpy
stand-ins in warp_shim.py (Warp / kornia / plyfile are not installable here), constructs `Simulator` exactly as
main_gui.py:39-46 does (float32 bbox / base tensors), feeds it the synthetic body of pienerf_b200.synthetic.make_body, and
records: topology, shape functions, IP parameters, the inverted system matrix, rhs vectors, and a 10-step sequence with a
drag force switched on before step 3 and cleared before step 7 (dof, dof_vel after every step; get_IP_info at steps 0, 3, 10).
Output: tests/golden/ref_sim_<body>.npz (fp64; the big shape-function arrays are stored for every `stride`-th IP).
"""
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)

import warp_shim  # noqa: E402

warp_shim.install("/root/reference")
import torch  # noqa: E402
from simulator.solver import Simulator  # noqa: E402  (the reference's class, unmodified)

from pienerf_b200.synthetic import make_body  # noqa: E402

FORCE_IP_FRACTION = 0.37          # the dragged IP: index int(0.37 * n_ip)
FORCE = (3.0e4, -1.0e4, 2.0e4)


def run(kind, steps=10, iters=10):
    body = make_body(kind, dx=0.05, bound=1.0, seed=0)
    t0 = time.time()
    sim = Simulator(dt=1e-2, iters=iters, bbox=torch.tensor([2.0, 2.0, 2.0]), dx=0.05, stiff=1e5, base=torch.tensor([-1.0, -1.0, -1.0]))
    sim.pos = torch.from_numpy(body["pos"].astype(np.float64))
    sim.mass = torch.from_numpy(body["mass"].astype(np.float64))
    sim.mu = torch.from_numpy(body["mu"].astype(np.float64))
    sim.lam = torch.from_numpy(body["lam"].astype(np.float64))
    sim.is_pin = torch.from_numpy(body["pin"].astype(bool))
    sim.initialize()
    print(f"[{kind}] initialize: {time.time() - t0:.1f} s, n_ip {sim.IP_pos.shape[0]}, n_k {sim.kernel_pos.shape[0]}", flush=True)
    n_ip, n_k = sim.IP_pos.shape[0], sim.kernel_pos.shape[0]
    stride = max(1, n_ip // 64)
    out = {
        "kind": kind, "steps": steps, "iters": iters, "stride": stride, "kdx": float(sim.kdx), "res": sim.res.numpy(),
        "IP_grid": sim.IP_grid.numpy(), "IP_pos": sim.IP_pos.numpy(), "IP_kernel": sim.IP_kernel.numpy(), "pts_kernel": sim.pts_kernel.numpy(),
        "pts_IP": sim.pts_IP.numpy(), "kernel_pos": sim.kernel_pos.numpy(),
        "IP_mu": sim.IP_mu.numpy(), "IP_lam": sim.IP_lam.numpy(), "IP_rho": sim.IP_rho.numpy(),
        "IP_Nx": sim.IP_Nx.numpy()[::stride], "IP_dNx": sim.IP_dNx.numpy()[::stride], "IP_ddNx": sim.IP_ddNx.numpy()[::stride],
        "pts_Nx": sim.pts_Nx.numpy()[::stride],
        # the reference stores Mat (x) I3; the compact [n,n] block is rows/cols 0::3
        "global_matrix": sim.global_matrix.numpy()[0::3, 0::3].copy(), "mass_matrix_invt2": sim.mass_matrix_invt2.numpy()[0::3, 0::3].copy(),
        "global_matrix_offdiag_max": float(max(sim.global_matrix.numpy()[0::3, 1::3].__abs__().max(), sim.global_matrix.numpy()[1::3, 2::3].__abs__().max())),
        "rhs_rest": sim.rhs_rest.numpy(), "rhs_gravity": sim.rhs_gravity.numpy(), "dof_rest": sim.dof_rest.numpy(),
    }
    info = {}
    info[0] = [t.numpy().copy() for t in sim.get_IP_info()]
    vid = int(FORCE_IP_FRACTION * n_ip)
    dofs, vels = [], []
    for s in range(steps):
        if s == 3:
            sim.update_force(vid, torch.tensor(FORCE, dtype=torch.float64))
            out["dof_f"] = sim.dof_f.numpy().copy()
        if s == 7:
            sim.clear_force()
        t1 = time.time()
        sim.stepforward()
        dofs.append(sim.dof.numpy().copy()); vels.append(sim.dof_vel.numpy().copy())
        if s + 1 in (3, steps):
            info[s + 1] = [t.numpy().copy() for t in sim.get_IP_info()]
        print(f"[{kind}] step {s}: {time.time() - t1:.1f} s", flush=True)
    out["force_ip"] = vid; out["force"] = np.asarray(FORCE)
    out["dof"] = np.stack(dofs); out["dof_vel"] = np.stack(vels)
    for k, (p, F, dF) in info.items():
        out[f"info{k}_pos"] = p; out[f"info{k}_F"] = F; out[f"info{k}_dF"] = dF
    sim.update_pos()
    out["pos_final"] = sim.pos.numpy().copy()
    path = os.path.join(HERE, f"ref_sim_{kind}.npz")
    np.savez_compressed(path, **out)
    print(f"[{kind}] wrote {path} ({os.path.getsize(path) / 1e6:.2f} MB) in {time.time() - t0:.0f} s", flush=True)


if __name__ == "__main__":
    for k in (sys.argv[1:] or ["block64", "block512"]):
        run(k)
