"""`_qgmls` — the native operator set of the Q-GMLS simulator (one op per Warp kernel of the reference's
simulator/cpu_utils.py + cuda_utils.py, plus the fused step), backed by the C-ABI.  All tensors are CUDA,
contiguous; float64 unless noted."""
import ctypes as C

import torch

from ._lib import check, dptr, lib, stream_ptr

f64, i32 = torch.float64, torch.int32


def shape_functions(r, pos, topo, kernel_pos, want_derivatives=True):
    """calc_G/calc_Gp/calc_weight (cpu_utils.py:3-152): returns Nx [n,8,10], dNx [n,8,3,10], ddNx [n,8,3,3,10]."""
    n = pos.shape[0]
    dev = pos.device
    Nx = torch.empty(n, 8, 10, dtype=f64, device=dev)
    dNx = torch.empty(n, 8, 3, 10, dtype=f64, device=dev) if want_derivatives else None
    ddNx = torch.empty(n, 8, 3, 3, 10, dtype=f64, device=dev) if want_derivatives else None
    status = torch.zeros(1, dtype=i32, device=dev)
    check(lib.pn_qgmls_shape_functions(float(r), dptr(pos, "pos", f64), dptr(topo, "topo", i32), dptr(kernel_pos, "kernel_pos", f64),
                                       n, dptr(Nx), dptr(dNx), dptr(ddNx), dptr(status), stream_ptr()))
    if int(status.item()) != 0:
        raise RuntimeError("Q-GMLS moment matrix is singular for at least one point (too few kernels in reach)")
    return Nx, dNx, ddNx


def collect_param(pts_ip, mu, lam, mass, n_ip, dx):
    dev = mu.device
    ip_mu = torch.empty(n_ip, dtype=f64, device=dev); ip_lam = torch.empty_like(ip_mu); ip_rho = torch.empty_like(ip_mu)
    check(lib.pn_qgmls_collect_param(dptr(pts_ip, "pts_ip", i32), dptr(mu, "mu", f64), dptr(lam, "lam", f64), dptr(mass, "mass", f64),
                                     pts_ip.shape[0], n_ip, float(dx), dptr(ip_mu), dptr(ip_lam), dptr(ip_rho), stream_ptr()))
    return ip_mu, ip_lam, ip_rho


def build_ip_global(dx, dt, topo, mu, lam, rho, Nx, dNx, ddNx, mat):
    check(lib.pn_qgmls_build_ip_global(float(dx), float(dt), dptr(topo, "topo", i32), dptr(mu, "mu", f64), dptr(lam, "lam", f64),
                                       dptr(rho, "rho", f64), dptr(Nx, "Nx", f64), dptr(dNx, "dNx", f64), dptr(ddNx, "ddNx", f64),
                                       topo.shape[0], mat.shape[0], dptr(mat, "mat", f64), stream_ptr()))


def build_pin_global(stiff, vidx, topo, Nx, mat):
    check(lib.pn_qgmls_build_pin_global(float(stiff), dptr(vidx, "vidx", i32), vidx.shape[0], dptr(topo, "topo", i32),
                                        dptr(Nx, "Nx", f64), mat.shape[0], dptr(mat, "mat", f64), stream_ptr()))


def collect_gravity(dx, topo, Nx, gravity, rho, rhs):
    g = (C.c_double * 3)(*[float(v) for v in gravity])
    check(lib.pn_qgmls_collect_gravity(float(dx), dptr(topo, "topo", i32), dptr(Nx, "Nx", f64), C.cast(g, C.c_void_p),
                                       dptr(rho, "rho", f64), topo.shape[0], dptr(rhs, "rhs", f64), stream_ptr()))


def build_rhs(dx, topo, mu, lam, dNx, dof, n_k, adj_bgn, adj, adj_slices, ip_stress, partial, rhs):
    check(lib.pn_qgmls_build_rhs(float(dx), dptr(topo, "topo", i32), dptr(mu, "mu", f64), dptr(lam, "lam", f64), dptr(dNx, "dNx", f64),
                                 dptr(dof, "dof", f64), topo.shape[0], int(n_k), dptr(adj_bgn, "adj_bgn", i32), dptr(adj, "adj", i32),
                                 int(adj_slices), dptr(ip_stress, "ip_stress", f64), dptr(partial, "partial", f64), dptr(rhs, "rhs", f64),
                                 stream_ptr()))


def matvec3(mat, x, y):
    check(lib.pn_qgmls_matvec3(dptr(mat, "mat", f64), dptr(x, "x", f64), mat.shape[0], dptr(y, "y", f64), stream_ptr()))


def step_scratch_doubles(n_ip, n_k, adj_slices):
    return int(lib.pn_qgmls_step_scratch_doubles(int(n_ip), int(n_k), int(adj_slices)))


def step(desc, solver=0):
    """desc: a filled _lib.QgmlsStepT (the Simulator keeps one alive across frames)."""
    check(lib.pn_qgmls_step(C.byref(desc), int(solver), stream_ptr()))


_step_mode = [True]


def step_mode(force_multi_kernel):
    """True (default): 3 + 4*iters launches replayed as one CUDA graph; False: the whole step as ONE thread-block-cluster kernel
    when n <= 1280 (experimental: measured slower on B200)."""
    check(lib.pn_qgmls_step_mode(1 if force_multi_kernel else 0))
    _step_mode[0] = bool(force_multi_kernel)


def step_mode_value():
    return _step_mode[0]


def ip_info(topo, dof, Nx, dNx, ddNx, pos, F, dF):
    check(lib.pn_qgmls_ip_info(dptr(topo, "topo", i32), dptr(dof, "dof", f64), dptr(Nx, "Nx", f64), dptr(dNx, "dNx", f64),
                               dptr(ddNx, "ddNx", f64), topo.shape[0], dptr(pos, "pos", torch.float32), dptr(F, "F", torch.float32),
                               dptr(dF, "dF", torch.float32), stream_ptr()))


def update_pos(topo, dof, Nx, pos):
    check(lib.pn_qgmls_update_pos(dptr(topo, "topo", i32), dptr(dof, "dof", f64), dptr(Nx, "Nx", f64), topo.shape[0],
                                  dptr(pos, "pos", f64), stream_ptr()))


def update_force(vid, f, topo, Nx, rho, dx, dof_f):
    fv = (C.c_double * 3)(*[float(v) for v in f]) if f is not None else None
    check(lib.pn_qgmls_update_force(int(vid), C.cast(fv, C.c_void_p) if fv is not None else C.c_void_p(0), dptr(topo, "topo", i32),
                                    dptr(Nx, "Nx", f64), dptr(rho, "rho", f64), float(dx), dof_f.numel() // 3,
                                    dptr(dof_f, "dof_f", f64), stream_ptr()))
