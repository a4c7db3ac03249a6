"""ctypes front-end for oracle/sim_oracle.c -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference
legs may import this module.  Parity pinned by fixtures generated from the reference's own source
(tests/golden/make_golden_sim.py, tests/test_sim_golden.py; see the sim_oracle.c header).

Mirrors the reference `Simulator` API (simulator/solver.py:12-617) closely
enough that parity tests read like reference usage.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SRC = os.path.join(_HERE, "sim_oracle.c")
_SO = os.path.join(_HERE, "_build", "libsim_oracle.so")


def build(force=False):
    """gcc -O2 -fopenmp the C restatement (also called from __graft_entry__.build)."""
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(_SRC):
        os.makedirs(os.path.dirname(_SO), exist_ok=True)
        subprocess.check_call(["gcc", "-O2", "-fopenmp", "-shared", "-fPIC", "-o", _SO, _SRC, "-lm"])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        _lib.qo_create.restype = C.c_void_p
        _lib.qo_ptr.restype = C.c_void_p
        _lib.qo_kdx.restype = C.c_double
    return _lib


def _dp(a):
    return a.ctypes.data_as(C.c_void_p)


class OracleSimulator:
    """fp64 CPU oracle with the reference Simulator's constructor/method names."""

    def __init__(self, dt=1e-2, iters=20, bbox=(1.0, 1.0, 1.0), kres=7, dx=1.0,
                 gravity=(0.0, -9.8, 0.0), stiff=1e5, base=(-0.5, -0.5, -0.5), scale_dtype=np.float32):
        # solver.py:24-25 scales in the caller's dtype (main_gui.py passes float32 tensors)
        bbox = (np.asarray(bbox, dtype=scale_dtype) * scale_dtype(1.02)).astype(np.float64)
        base = (np.asarray(base, dtype=scale_dtype) * scale_dtype(1.01)).astype(np.float64)
        gravity = np.asarray(gravity, dtype=np.float64)
        self.dt, self.iters, self.dx, self.kres, self.stiff = dt, iters, dx, kres, stiff
        self.base, self.bbox, self.gravity = base, bbox, gravity
        L = lib()
        self._h = C.c_void_p(L.qo_create(C.c_double(dt), C.c_int(iters), _dp(bbox), C.c_int(kres), C.c_double(dx),
                                         _dp(gravity), C.c_double(stiff), _dp(base)))

    def __del__(self):
        try:
            lib().qo_destroy(self._h)
        except Exception:
            pass

    def initialize(self, pos, mass, mu, lam, is_pin):
        pos = np.ascontiguousarray(pos, dtype=np.float64)
        mass = np.ascontiguousarray(mass, dtype=np.float64)
        mu = np.ascontiguousarray(mu, dtype=np.float64)
        lam = np.ascontiguousarray(lam, dtype=np.float64)
        pin = np.ascontiguousarray(is_pin, dtype=np.uint8)
        rc = lib().qo_initialize(self._h, C.c_int(pos.shape[0]), _dp(pos), _dp(mass), _dp(mu), _dp(lam), _dp(pin))
        if rc != 0:
            raise RuntimeError(f"qo_initialize failed rc={rc}")
        self.n_ip = lib().qo_n_ip(self._h)
        self.n_k = lib().qo_n_k(self._h)
        self.n_pts = pos.shape[0]
        self.n = 10 * self.n_k
        self.kdx = lib().qo_kdx(self._h)
        return self

    def array(self, name):
        """Copy of an internal array, shaped as in the reference."""
        n_ip, n_k, n_pts, n = self.n_ip, self.n_k, self.n_pts, self.n
        shapes = {
            "pts_ip": ((n_pts,), np.int32), "pts_kernel": ((n_pts, 8), np.int32), "ip_kernel": ((n_ip, 8), np.int32),
            "ip_grid": ((n_ip, 3), np.int32), "ip_pos": ((n_ip, 3), np.float64), "kernel_pos": ((n_k, 3), np.float64),
            "pts_Nx": ((n_pts, 8, 10), np.float64), "ip_Nx": ((n_ip, 8, 10), np.float64),
            "ip_dNx": ((n_ip, 8, 3, 10), np.float64), "ip_ddNx": ((n_ip, 8, 3, 3, 10), np.float64),
            "ip_mu": ((n_ip,), np.float64), "ip_lam": ((n_ip,), np.float64), "ip_rho": ((n_ip,), np.float64),
            "A": ((n, n), np.float64), "Ainv": ((n, n), np.float64), "M": ((n, n), np.float64),
            "active": ((n_k,), np.uint8),
            "dof": ((3 * n,), np.float64), "dof_rest": ((3 * n,), np.float64), "dof_vel": ((3 * n,), np.float64),
            "dof_f": ((3 * n,), np.float64), "rhs_rest": ((3 * n,), np.float64), "rhs_gravity": ((3 * n,), np.float64),
        }
        shape, dt = shapes[name]
        ptr = lib().qo_ptr(self._h, name.encode())
        cnt = int(np.prod(shape))
        buf = (C.c_char * (cnt * np.dtype(dt).itemsize)).from_address(ptr)
        return np.frombuffer(buf, dtype=dt).reshape(shape).copy()

    @property
    def IP_pos(self):
        return self.array("ip_pos")

    def stepforward(self):
        lib().qo_step(self._h)

    def get_IP_info(self, with_pos64=False):
        n = self.n_ip
        pos = np.empty((n, 3), np.float32); F = np.empty((n, 9), np.float32); dF = np.empty((n, 27), np.float32)
        p64 = np.empty((n, 3), np.float64)
        lib().qo_ip_info(self._h, _dp(pos), _dp(F), _dp(dF), _dp(p64))
        return (pos, F, dF, p64) if with_pos64 else (pos, F, dF)

    def update_force(self, vid, f):
        f = np.ascontiguousarray(f, dtype=np.float64)
        lib().qo_update_force(self._h, C.c_int(int(vid)), _dp(f))

    def clear_force(self):
        f = np.zeros(3)
        lib().qo_update_force(self._h, C.c_int(-1), _dp(f))

    def update_pos(self):
        out = np.empty((self.n_pts, 3), np.float64)
        lib().qo_update_pos(self._h, _dp(out))
        return out

    def set_state(self, dof=None, vel=None):
        d = None if dof is None else np.ascontiguousarray(dof, dtype=np.float64)
        v = None if vel is None else np.ascontiguousarray(vel, dtype=np.float64)
        lib().qo_set_dof(self._h, None if d is None else _dp(d), None if v is None else _dp(v))

    def build_rhs(self):
        out = np.empty(3 * self.n, np.float64)
        lib().qo_build_rhs(self._h, _dp(out))
        return out

    @staticmethod
    def threads():
        return lib().qo_threads()

    @staticmethod
    def set_threads(n):
        lib().qo_set_threads(C.c_int(int(n)))


def svd3(F):
    F = np.ascontiguousarray(F, dtype=np.float64)
    U = np.empty(9); s = np.empty(3); V = np.empty(9)
    lib().qo_svd3(_dp(F), _dp(U), _dp(s), _dp(V))
    return U.reshape(3, 3), s, V.reshape(3, 3)


def volume_project(s):
    s = np.ascontiguousarray(s, dtype=np.float64)
    o = np.empty(3)
    lib().qo_volume_project(_dp(s), _dp(o))
    return o


def invert(A):
    A = np.array(A, dtype=np.float64, order="C")
    Ai = np.empty_like(A)
    rc = lib().qo_invert(_dp(A), _dp(Ai), C.c_int(A.shape[0]))
    if rc:
        raise RuntimeError("singular")
    return Ai
