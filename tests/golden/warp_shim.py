"""numpy stand-ins for the third-party names the reference simulator imports, so that the UNMODIFIED modules
`/root/reference/simulator/{func_utils,cpu_utils,cuda_utils,solver}.py` run on a CPU-only host.

TEST INFRASTRUCTURE ONLY (used by tests/golden/make_golden_sim.py to emit fixtures).  Nothing under
pienerf_b200/ imports this.

What is shimmed, and what that means for the fixtures:
  * `warp` (warp-lang 0.13.0, README.md:38): `@wp.kernel` bodies are executed as plain Python, one call per thread id;
    `wp.vec / wp.mat` values are ndarray subclasses whose `*` follows Warp (mat*mat and mat*vec are matrix products,
    anything with a scalar is a scaling); `wp.array` element access returns views so `Nx[vid, i][0] = v` writes through,
    as a Warp array reference does; `wp.atomic_add` is a serial read-modify-write (thread order = ascending tid, so the
    fp64 summation order differs from the GPU's nondeterministic one by round-off only).
  * `wp.svd3` is NOT Warp's code (its source is not in the reference tree): an exact LAPACK SVD normalised to the
    convention of Warp's implementation (McAdams et al., "Computing the SVD of 3x3 matrices with minimal branching"):
    A = U diag(s) V^T with U and V proper rotations (det +1), |s0| >= |s1| >= |s2|, s0, s1 >= 0 and the sign of det(A)
    carried by s2.  Warp's 4-sweep approximate-Givens Jacobi converges to this decomposition; its residual is the only
    third-party arithmetic the fixtures do not contain.
  * `kornia.utils.grid.create_meshgrid3d` (version unpinned by the reference): restated from kornia's published
    implementation: `stack(meshgrid([zs, xs, ys], indexing="ij"), -1).permute(0, 2, 1, 3)[None]`, i.e. entry [0,d,h,w] =
    (d, w, h).
  * `plyfile`: names only (the fixtures set `pos / mass / mu / lam / is_pin` directly, as InitializeFromPly does after parsing).
  * torch: `torch.set_default_device("cuda")` (func_utils.py:6) and `Tensor.cuda()` become no-ops, so every tensor the
    reference creates lives on the CPU; dtypes, type promotion and the order of operations are torch's own.
"""
import sys
import types

import numpy as np
import torch

_tid = 0


# ----------------------------------------------------------------------------------------------- value types
class _Val(np.ndarray):
    __array_priority__ = 1000.0

    def __mul__(self, o):
        if isinstance(o, _Val):
            if isinstance(self, Mat):                                   # mat * mat, mat * vec: matrix products
                r = np.matmul(np.asarray(self), np.asarray(o))
                return r.view(Mat if r.ndim == 2 else Vec)
            raise TypeError("vec * vec / vec * mat are not used by the reference kernels")
        return np.multiply(np.asarray(self), o).view(type(self))

    def __rmul__(self, o):
        return np.multiply(o, np.asarray(self)).view(type(self))

    def __imul__(self, o):
        np.multiply(np.asarray(self), o, out=np.asarray(self))
        return self


class Vec(_Val):
    pass


class Mat(_Val):
    pass


def _scalar(x):
    if isinstance(x, torch.Tensor):
        return x.item()
    return x


def vec(length, dtype):
    def new(cls, *args):
        a = np.zeros(length, dtype=dtype).view(cls)
        if len(args) == 1:
            a[:] = _scalar(args[0])
        elif len(args) == length:
            a[:] = [_scalar(v) for v in args]
        elif len(args) != 0:
            raise TypeError(f"vec{length}: {len(args)} arguments")
        return a
    return type(f"vec{length}", (Vec,), {"__new__": new, "_shape": (length,), "_dtype": dtype})


def mat(shape, dtype):
    n = shape[0] * shape[1]

    def new(cls, *args):
        a = np.zeros(shape, dtype=dtype).view(cls)
        if len(args) == 1:
            a[:] = _scalar(args[0])
        elif len(args) == n:
            a[:] = np.asarray([_scalar(v) for v in args], dtype=dtype).reshape(shape)     # row-major, as wp.mat(...)
        elif len(args) != 0:
            raise TypeError(f"mat{shape}: {len(args)} arguments")
        return a
    return type(f"mat{shape[0]}{shape[1]}", (Mat,), {"__new__": new, "_shape": tuple(shape), "_dtype": dtype})


float64 = np.float64
float32 = np.float32
int32 = np.int32
vec2i = vec(2, np.int32)
vec3i = vec(3, np.int32)
vec3f = vec(3, np.float32)

# main_sample.py indexes one past the end of its arrays for the last lattice point of every row (see
# pienerf_b200/sampling.py): with OOB_TOLERANT set, such writes are dropped and such reads return 0, the model
# "the memory behind the array is zero and nobody else's" that the reference silently relies on.
OOB_TOLERANT = False


# ----------------------------------------------------------------------------------------------- arrays
class WpArray:
    """wp.array: `data` holds array dims followed by the element dims of a vec / mat dtype."""

    def __init__(self, data, dtype):
        self.data = data
        self.dtype = dtype
        self.elem = getattr(dtype, "_shape", ())
        self.kind = Vec if len(self.elem) == 1 else Mat if len(self.elem) == 2 else None
        self.shape = data.shape[:data.ndim - len(self.elem)]

    def to(self, device):
        return self

    def __getitem__(self, idx):
        if OOB_TOLERANT and np.ndim(idx) == 0 and not (0 <= int(idx) < self.data.shape[0]):
            return np.zeros(self.elem, self.data.dtype).view(self.kind) if self.kind is not None else self.data.dtype.type(0)
        v = self.data[idx]
        if self.kind is not None and isinstance(v, np.ndarray) and v.shape == self.elem:
            return v.view(self.kind)
        return v

    def __setitem__(self, idx, val):
        if OOB_TOLERANT and np.ndim(idx) == 0 and not (0 <= int(idx) < self.data.shape[0]):
            return
        self.data[idx] = np.asarray(val)


def _shape_tuple(shape):
    if isinstance(shape, (tuple, list)):
        return tuple(int(_scalar(s)) for s in shape)
    return (int(_scalar(shape)),)


def _np_dtype(dtype):
    return getattr(dtype, "_dtype", dtype)


def zeros(shape=0, dtype=np.float64, **kw):
    return WpArray(np.zeros(_shape_tuple(shape) + getattr(dtype, "_shape", ()), dtype=_np_dtype(dtype)), dtype)


def array(shape=0, dtype=np.float64, **kw):
    return zeros(shape=shape, dtype=dtype)                               # also evaluated inside kernel annotations


def from_torch(t, dtype=None):
    a = t.detach().numpy()                                               # shares memory with the (CPU) tensor, like Warp
    if dtype is None:
        dtype = {np.dtype(np.float64): np.float64, np.dtype(np.float32): np.float32, np.dtype(np.int32): np.int32}[a.dtype]
    es = getattr(dtype, "_shape", ())
    assert a.shape[a.ndim - len(es):] == es, (a.shape, es)
    return WpArray(a, dtype)


def to_torch(a):
    return torch.from_numpy(a.data)


# ----------------------------------------------------------------------------------------------- builtins
def tid():
    return _tid


def floor(x):
    return np.floor(x)


def printf(*a):
    pass


def dot(a, b):
    return np.float64(np.dot(np.asarray(a), np.asarray(b)))


def length(a):
    a = np.asarray(a)
    return np.sqrt(np.dot(a, a))


def outer(a, b):
    return np.outer(np.asarray(a), np.asarray(b)).view(Mat)


def identity(n, dtype):
    return np.eye(n, dtype=dtype).view(Mat)


def transpose(m):
    return np.ascontiguousarray(np.asarray(m).T).view(Mat)


def atomic_add(arr, *args):
    *idx, val = args
    idx = tuple(int(i) for i in idx)
    idx = idx[0] if len(idx) == 1 else idx
    old = arr.data[idx]
    old = old.copy() if isinstance(old, np.ndarray) else old
    arr.data[idx] += np.asarray(val)
    return old


def svd3(A, U, sig, V):
    """A = U diag(sig) V^T, U and V rotations, the sign of det(A) on sig[2] (see the module docstring)."""
    u, s, vt = np.linalg.svd(np.asarray(A, dtype=np.float64))
    v = vt.T.copy()
    s = s.copy()
    if np.linalg.det(v) < 0:
        v[:, 2] = -v[:, 2]
        s[2] = -s[2]
    if np.linalg.det(u) < 0:
        u[:, 2] = -u[:, 2]
        s[2] = -s[2]
    U[...] = u
    sig[...] = s
    V[...] = v


def func(f):
    return f


class _Kernel:
    def __init__(self, f):
        self.f = f
        self.ann = [f.__annotations__.get(n) for n in f.__code__.co_varnames[:f.__code__.co_argcount]]


def kernel(f):
    return _Kernel(f)


def launch(kernel, dim, inputs, outputs=(), device=None, **kw):
    global _tid
    args = []
    for a, ann in zip(list(inputs) + list(outputs), kernel.ann):
        if ann in (np.float64, np.float32):
            a = ann(_scalar(a))                                          # scalars are converted to the declared type at launch
        elif ann is np.int32:
            a = np.int32(_scalar(a))
        elif isinstance(ann, type) and issubclass(ann, Vec) and isinstance(a, torch.Tensor):
            a = ann(*a.tolist())                                         # a torch tensor passed for a wp.vec3f parameter
        args.append(a)
    n = int(np.prod(_shape_tuple(dim)))
    f = kernel.f
    for t in range(n):
        _tid = t
        f(*args)


def synchronize():
    pass


def set_device(d):
    pass


def init():
    pass


# ----------------------------------------------------------------------------------------------- kornia / plyfile
def create_meshgrid3d(depth, height, width, normalized_coordinates=True, device=None, dtype=None):
    depth, height, width = int(_scalar(depth)), int(_scalar(height)), int(_scalar(width))
    assert not normalized_coordinates
    xs = torch.linspace(0, width - 1, width, dtype=dtype)
    ys = torch.linspace(0, height - 1, height, dtype=dtype)
    zs = torch.linspace(0, depth - 1, depth, dtype=dtype)
    base = torch.stack(torch.meshgrid([zs, xs, ys], indexing="ij"), dim=-1)      # D x W x H x 3
    return base.permute(0, 2, 1, 3).unsqueeze(0)                                  # 1 x D x H x W x 3


def _is_cuda(d):
    return d is not None and str(d).startswith("cuda")


def _cpu_only_factories():
    """device="cuda" / .to("cuda:0") mean "this machine's device" in the reference; here that is the CPU."""
    if getattr(torch, "_pn_cpu_only", False):
        return
    torch._pn_cpu_only = True
    for name in ("zeros", "ones", "empty", "full", "rand", "randn", "tensor", "arange", "linspace", "zeros_like", "ones_like", "eye"):
        orig = getattr(torch, name)

        def wrapped(*a, _orig=orig, **k):
            if _is_cuda(k.get("device")):
                k["device"] = "cpu"
            return _orig(*a, **k)
        setattr(torch, name, wrapped)
    orig_to = torch.Tensor.to

    def to(self, *a, **k):
        a = tuple("cpu" if isinstance(x, (str, torch.device)) and _is_cuda(x) else x for x in a)
        if _is_cuda(k.get("device")):
            k["device"] = "cpu"
        return orig_to(self, *a, **k)
    torch.Tensor.to = to


def stub_missing_modules(names):
    """Permissive stand-ins for third-party modules the reference imports at module scope but the exercised code never calls
    (imageio, cv2, trimesh, lpips, ... in nerf/utils.py): attribute access yields a dummy class."""
    import importlib

    class _Stub(types.ModuleType):
        def __getattr__(self, item):
            if item.startswith("__"):
                raise AttributeError(item)
            return type(item, (), {"__init__": lambda self, *a, **k: None})
    for n in names:
        try:
            importlib.import_module(n)
        except Exception:
            parts = n.split(".")
            for i in range(1, len(parts) + 1):
                sub = ".".join(parts[:i])
                if sub not in sys.modules:
                    sys.modules[sub] = _Stub(sub)
                    sys.modules[sub].__path__ = []


def install(reference_root="/root/reference"):
    """Register the stand-in modules and make torch CPU-only.  Call BEFORE importing `simulator.*`."""
    wp = types.ModuleType("warp")
    for name in ("vec", "mat", "float64", "float32", "int32", "vec2i", "vec3i", "vec3f", "floor", "printf", "zeros", "array", "from_torch", "to_torch", "tid", "dot",
                 "length", "outer", "identity", "transpose", "atomic_add", "svd3", "func", "kernel", "launch", "synchronize",
                 "set_device", "init"):
        setattr(wp, name, globals()[name])
    sys.modules["warp"] = wp
    kornia = types.ModuleType("kornia"); ku = types.ModuleType("kornia.utils"); kg = types.ModuleType("kornia.utils.grid")
    kg.create_meshgrid3d = create_meshgrid3d
    kornia.utils = ku; ku.grid = kg
    sys.modules.update({"kornia": kornia, "kornia.utils": ku, "kornia.utils.grid": kg})
    ply = types.ModuleType("plyfile")
    ply.PlyData = ply.PlyElement = type("Unavailable", (), {})
    sys.modules["plyfile"] = ply
    torch.set_default_device = lambda *a, **k: None
    torch.Tensor.cuda = lambda self, *a, **k: self
    _cpu_only_factories()
    if reference_root not in sys.path:
        sys.path.insert(0, reference_root)
    return wp
