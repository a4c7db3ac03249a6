"""Pins the simulator oracle (oracle/sim_oracle.c) — and, on the GPU, the CUDA simulator itself — to fixtures produced by
the reference's OWN source: tests/golden/ref_sim_*.npz were written by tests/golden/make_golden_sim.py, which runs the
unmodified /root/reference/simulator/{func_utils,cpu_utils,cuda_utils,solver}.py on numpy stand-ins for warp / kornia
(tests/golden/warp_shim.py).  Sequence: init, 10 x stepforward (iters 10), drag force on before step 3 and off before
step 7, get_IP_info at steps 0 / 3 / 10.  Only wp.svd3's own iteration is not in the fixtures (exact SVD in Warp's convention).

Tolerances (fp64 on both sides, different summation orders): topology bit-equal; shape functions / matrices / rhs 1e-9
relative; DOFs 1e-9; DOF velocities 1e-7 (a small difference divided by dt); fp32 IP state 1e-6."""
import os

import numpy as np
import pytest

from oracle.sim_oracle import OracleSimulator
from pienerf_b200.synthetic import make_body

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
KINDS = [k for k in ("block64", "block512") if os.path.exists(os.path.join(GOLDEN, f"ref_sim_{k}.npz"))]


def _rel(a, b):
    return float(np.abs(np.asarray(a, dtype=np.float64) - b).max() / max(np.abs(b).max(), 1e-300))


def _load(kind):
    return np.load(os.path.join(GOLDEN, f"ref_sim_{kind}.npz"))


def test_fixtures_present():
    assert "block64" in KINDS and "block512" in KINDS, "run tests/golden/make_golden_sim.py"


@pytest.mark.parametrize("kind", KINDS)
def test_oracle_init_matches_reference_source(kind):
    g = _load(kind)
    b = make_body(kind)
    s = OracleSimulator(dt=1e-2, iters=int(g["iters"]), bbox=[2, 2, 2], dx=0.05, stiff=1e5, base=[-1, -1, -1])
    s.initialize(b["pos"], b["mass"], b["mu"], b["lam"], b["pin"])
    st = int(g["stride"])
    assert s.kdx == float(g["kdx"])                                     # fp32 evaluation of res.max()*dx/(kres-1), solver.py:184
    for name, ref in (("ip_grid", "IP_grid"), ("ip_kernel", "IP_kernel"), ("pts_kernel", "pts_kernel"), ("pts_ip", "pts_IP")):
        assert np.array_equal(s.array(name), g[ref]), name               # incl. the meshgrid [1,2] swap and kernel_idx = 0 aliasing
    for name, ref in (("ip_pos", "IP_pos"), ("kernel_pos", "kernel_pos")):
        assert np.array_equal(s.array(name), g[ref]), name
    for name, ref in (("ip_mu", "IP_mu"), ("ip_lam", "IP_lam"), ("ip_rho", "IP_rho")):
        assert _rel(s.array(name), g[ref]) < 1e-12, name
    for name, ref in (("ip_Nx", "IP_Nx"), ("ip_dNx", "IP_dNx"), ("ip_ddNx", "IP_ddNx"), ("pts_Nx", "pts_Nx")):
        assert _rel(s.array(name)[::st], g[ref]) < 1e-9, (name, _rel(s.array(name)[::st], g[ref]))
    assert float(g["global_matrix_offdiag_max"]) == 0.0                  # Mat (x) I3: the compact [n,n] form loses nothing
    assert _rel(s.array("Ainv"), g["global_matrix"]) < 1e-9
    assert _rel(s.array("M"), g["mass_matrix_invt2"]) < 1e-12
    assert _rel(s.array("rhs_rest"), g["rhs_rest"]) < 1e-9
    assert _rel(s.array("rhs_gravity"), g["rhs_gravity"]) < 1e-12
    assert np.array_equal(s.array("dof_rest"), g["dof_rest"])
    p, F, dF = s.get_IP_info()
    assert _rel(p, g["info0_pos"]) < 1e-6 and np.abs(F - g["info0_F"]).max() < 1e-6 and np.abs(dF - g["info0_dF"]).max() < 1e-6


@pytest.mark.parametrize("kind", KINDS)
def test_oracle_step_sequence_matches_reference_source(kind):
    g = _load(kind)
    b = make_body(kind)
    s = OracleSimulator(dt=1e-2, iters=int(g["iters"]), bbox=[2, 2, 2], dx=0.05, stiff=1e5, base=[-1, -1, -1])
    s.initialize(b["pos"], b["mass"], b["mu"], b["lam"], b["pin"])
    for k in range(int(g["steps"])):
        if k == 3:
            s.update_force(int(g["force_ip"]), g["force"])
            assert _rel(s.array("dof_f"), g["dof_f"]) < 1e-12
        if k == 7:
            s.clear_force()
        s.stepforward()
        assert _rel(s.array("dof"), g["dof"][k]) < 1e-9, (k, _rel(s.array("dof"), g["dof"][k]))
        assert _rel(s.array("dof_vel"), g["dof_vel"][k]) < 1e-7, (k, _rel(s.array("dof_vel"), g["dof_vel"][k]))
        if k + 1 in (3, int(g["steps"])):
            p, F, dF = s.get_IP_info()
            assert _rel(p, g[f"info{k + 1}_pos"]) < 1e-6                 # fp32 outputs in the renderer layouts (solver.py:421-424)
            assert np.abs(F - g[f"info{k + 1}_F"]).max() < 1e-6 and np.abs(dF - g[f"info{k + 1}_dF"]).max() < 1e-5
    assert _rel(s.update_pos(), g["pos_final"]) < 1e-9


def test_shim_svd3_convention():
    """The stand-in for wp.svd3 and the oracle's own SVD agree on the convention (rotations, sign on the smallest sigma)."""
    import sys
    sys.path.insert(0, GOLDEN)
    import warp_shim as ws
    from oracle.sim_oracle import svd3
    rng = np.random.default_rng(5)
    for i in range(20):
        A = rng.normal(size=(3, 3))
        if i % 3 == 0:
            A[:, 1] *= -1
        U = np.zeros((3, 3)); sg = np.zeros(3); V = np.zeros((3, 3))
        ws.svd3(A, U, sg, V)
        Uo, so, Vo = svd3(A)
        assert abs(np.linalg.det(U) - 1) < 1e-12 and abs(np.linalg.det(V) - 1) < 1e-12
        assert np.abs(U @ np.diag(sg) @ V.T - A).max() < 1e-12
        assert np.allclose(sg, so, atol=1e-12)
        assert np.abs(U @ V.T - Uo @ Vo.T).max() < 1e-9                  # the quantities the solver uses: R = U V^T ...
        proj = np.array([1.1, 0.95, 0.9])
        assert np.abs(U @ np.diag(proj) @ V.T - Uo @ np.diag(proj) @ Vo.T).max() < 1e-9   # ... and U diag(.) V^T


# ------------------------------------------------------------------------------------------------ CUDA vs the fixtures
@pytest.mark.gpu
@pytest.mark.parametrize("kind", KINDS)
def test_cuda_simulator_matches_reference_source(kind):
    """BASELINE.json bar: IP positions / velocities within 1e-4 relative of the reference on the same step sequence."""
    torch = pytest.importorskip("torch")
    from pienerf_b200.simulator import Simulator
    g = _load(kind)
    b = make_body(kind)
    s = Simulator(dt=1e-2, iters=int(g["iters"]), bbox=torch.tensor([2.0, 2.0, 2.0]), dx=0.05, stiff=1e5, base=torch.tensor([-1.0, -1.0, -1.0]))
    s.set_points(b["pos"], b["mass"], b["mu"], b["lam"], b["pin"]).initialize()
    st = int(g["stride"])
    assert np.array_equal(s.IP_kernel.cpu().numpy(), g["IP_kernel"]) and np.array_equal(s.pts_kernel.cpu().numpy(), g["pts_kernel"])
    assert np.array_equal(s.IP_pos.cpu().numpy(), g["IP_pos"]) and np.array_equal(s.kernel_pos.cpu().numpy(), g["kernel_pos"])
    for t, ref in ((s.IP_Nx, "IP_Nx"), (s.IP_dNx, "IP_dNx"), (s.IP_ddNx, "IP_ddNx")):
        assert _rel(t.cpu().numpy()[::st], g[ref]) < 1e-8, ref
    assert _rel(s.global_matrix.cpu().numpy(), g["global_matrix"]) < 1e-6
    assert _rel(s.mass_matrix_invt2.cpu().numpy(), g["mass_matrix_invt2"]) < 1e-8
    worst_p = worst_v = 0.0
    for k in range(int(g["steps"])):
        if k == 3:
            s.update_force(int(g["force_ip"]), torch.tensor(g["force"]))
            assert _rel(s.dof_f.cpu().numpy().reshape(-1), g["dof_f"]) < 1e-12
        if k == 7:
            s.clear_force()
        s.stepforward()
        worst_p = max(worst_p, _rel(s.dof.cpu().numpy().reshape(-1), g["dof"][k]))
        worst_v = max(worst_v, _rel(s.dof_vel.cpu().numpy().reshape(-1), g["dof_vel"][k]))
        if k + 1 in (3, int(g["steps"])):
            p, F, dF = s.get_IP_info()
            assert _rel(p.cpu().numpy(), g[f"info{k + 1}_pos"]) < 1e-6
            assert np.abs(F.cpu().numpy() - g[f"info{k + 1}_F"]).max() < 1e-5 and np.abs(dF.cpu().numpy() - g[f"info{k + 1}_dF"]).max() < 1e-3
    assert worst_p < 1e-7 and worst_v < 1e-4, (worst_p, worst_v)
    assert _rel(s.update_pos().cpu().numpy(), g["pos_final"]) < 1e-8
