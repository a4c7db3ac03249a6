"""Pins the fp64 C oracle of the simulator (oracle/sim_oracle.c) with the GMLS identities the reference's math
implies and with numpy cross-checks of the third-party arithmetic it restates."""
import numpy as np
import pytest

from oracle.sim_oracle import OracleSimulator, invert, svd3, volume_project
from pienerf_b200.synthetic import make_body


def make(kind="block64", gravity=(0.0, -9.8, 0.0), iters=10):
    b = make_body(kind)
    s = OracleSimulator(dt=1e-2, iters=iters, bbox=[2, 2, 2], dx=0.05, stiff=1e5, base=[-1, -1, -1], gravity=gravity)
    s.initialize(b["pos"], b["mass"], b["mu"], b["lam"], b["pin"])
    return s, b


def test_topology_counts():
    s, b = make("block512")
    assert s.n_ip == 512 and s.n_pts == 512 and s.n_k == 27
    assert abs(s.kdx - float(np.float32(np.float32(40) * np.float32(0.05)) / np.float32(6))) == 0
    ipk = s.array("ip_kernel")
    assert ipk.min() >= 0 and ipk.max() < s.n_k
    # corner S -> (x,y,z) = (S>>2&1, S>>1&1, S&1): corner 7 is the +++ neighbour of corner 0
    kp = s.array("kernel_pos")
    d = kp[ipk[:, 7]] - kp[ipk[:, 0]]
    assert np.allclose(d, s.kdx, atol=1e-6)


def test_partition_of_unity_and_rest_state():
    s, _ = make("block512")
    N = s.array("ip_Nx")
    assert np.abs(N[:, :, 0].sum(1) - 1).max() < 1e-12                # sum_i N_i0 = 1
    pos, F, dF, p64 = s.get_IP_info(with_pos64=True)
    assert np.abs(p64 - s.IP_pos).max() < 1e-12                       # quadratic reproduction at rest
    assert np.abs(F.reshape(-1, 3, 3) - np.eye(3)).max() < 1e-6
    assert np.abs(dF).max() < 1e-6
    dN = s.array("ip_dNx")
    assert np.abs(dN[:, :, :, 0].sum(1)).max() < 1e-10                # gradients of a partition of unity sum to 0


def test_quadratic_reproduction():
    """An arbitrary quadratic map is reproduced exactly when slot idx(x,y) holds d2phi/dXx dXy (single count)."""
    s, _ = make("block64")
    rng = np.random.default_rng(0)
    A = rng.normal(size=(3, 3)); Hs = rng.normal(size=(3, 3, 3)); Hs = 0.5 * (Hs + Hs.transpose(0, 2, 1)); c = rng.normal(size=3)
    kp = s.array("kernel_pos")

    def phi(X):
        return c + X @ A.T + 0.5 * np.einsum("rjk,nj,nk->nr", Hs, X, X)

    def grad(X):
        return A[None] + np.einsum("rjk,nk->nrj", Hs, X)
    dof = np.zeros((s.n_k, 10, 3))
    dof[:, 0] = phi(kp)
    G = grad(kp)
    for x in range(3):
        dof[:, 1 + x] = G[:, :, x]
    slot = {(0, 0): 4, (0, 1): 5, (0, 2): 6, (1, 1): 7, (1, 2): 8, (2, 2): 9}
    for (x, y), sl in slot.items():
        dof[:, sl] = Hs[:, x, y][None]
    s.set_state(dof=dof.reshape(-1))
    pos, F, dF, p64 = s.get_IP_info(with_pos64=True)
    X = s.IP_pos
    assert np.abs(p64 - phi(X)).max() < 1e-10
    Fm = F.reshape(-1, 3, 3).transpose(0, 2, 1)                       # F[a*3+b] = F[b][a]
    assert np.abs(Fm - grad(X)).max() < 1e-5
    dFm = dF.reshape(-1, 3, 3, 3)                                     # [c][r][j]
    want = np.broadcast_to(Hs.transpose(2, 0, 1)[None], dFm.shape)    # d2phi_r/dXj dXc at [c][r][j]
    assert np.abs(dFm - want).max() < 1e-4


def test_rest_is_fixed_point_without_loads():
    s, _ = make("block64", gravity=(0, 0, 0))
    for _ in range(3):
        s.stepforward()
    assert np.abs(s.array("dof") - s.array("dof_rest")).max() < 1e-12
    assert np.abs(s.array("dof_vel")).max() < 1e-10


def test_gravity_sags_and_pins_hold():
    s, b = make("block512")
    p0 = s.get_IP_info(with_pos64=True)[3]
    for _ in range(20):
        s.stepforward()
    p1 = s.get_IP_info(with_pos64=True)[3]
    dy = p1[:, 1] - p0[:, 1]
    assert dy.min() < -1e-4                                           # free part moves down
    top = p0[:, 1] > p0[:, 1].max() - 1e-9
    assert np.abs(p1[top] - p0[top]).max() < 0.5 * np.abs(dy).max()   # penalty pins (stiff=1e5): top layer lags
    assert np.isfinite(s.array("dof")).all()


def test_system_matrices():
    s, _ = make("block64")
    A = s.array("A"); M = s.array("M"); Ai = s.array("Ainv"); act = s.array("active").astype(bool)
    assert np.allclose(A, A.T, rtol=1e-12, atol=1e-9) and np.allclose(M, M.T, rtol=1e-12, atol=1e-9)
    lst = (np.nonzero(act)[0][:, None] * 10 + np.arange(10)[None]).reshape(-1)
    sub = A[np.ix_(lst, lst)] + 1e-3 * np.eye(lst.size)
    assert np.abs(Ai[np.ix_(lst, lst)] @ sub - np.eye(lst.size)).max() < 1e-8
    assert np.linalg.eigvalsh(sub).min() > 0                          # SPD
    # total mass: sum_ij M[i0, j0] * dt^2 = sum rho dx^3 (partition of unity)
    total = M[0::10, 0::10].sum() * s.dt ** 2
    assert abs(total - (s.array("ip_rho") * s.dx ** 3).sum()) < 1e-9 * total


def test_update_force_and_clear():
    s, _ = make("block64")
    s.update_force(5, [1.0, 2.0, 3.0])
    f = s.array("dof_f").reshape(-1, 3)
    m = s.array("ip_rho")[5] * s.dx ** 3
    assert np.allclose(f[0::10].sum(0), m * np.array([1.0, 2.0, 3.0]), rtol=1e-10)   # slot-0 weights sum to 1
    s.clear_force()
    assert np.abs(s.array("dof_f")).max() == 0


@pytest.mark.parametrize("seed", range(5))
def test_svd3_against_numpy(seed):
    rng = np.random.default_rng(seed)
    F = rng.normal(size=(3, 3))
    if seed == 4:
        F[:, 0] *= -1                                                 # force a reflection sometimes
    U, sg, V = svd3(F)
    assert np.abs(U @ np.diag(sg) @ V.T - F).max() < 1e-12
    assert abs(np.linalg.det(U) - 1) < 1e-12 and abs(np.linalg.det(V) - 1) < 1e-12
    assert np.allclose(np.sort(np.abs(sg))[::-1], np.linalg.svd(F)[1], atol=1e-12)
    assert sg[0] >= sg[1] >= abs(sg[2])
    assert np.sign(sg[2]) == np.sign(np.linalg.det(F))


def test_volume_project_and_inverse():
    out = volume_project([1.3, 0.9, 0.7])
    assert abs(np.prod(out) - 1) < 1e-3                               # three Newton-like iterations, not exact
    assert np.allclose(volume_project([2.0, 0.5, 1.0]), [2.0, 0.5, 1.0], atol=1e-14)
    A = np.random.default_rng(3).normal(size=(40, 40))
    assert np.abs(invert(A) - np.linalg.inv(A)).max() < 1e-9
