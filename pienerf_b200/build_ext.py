"""Builds the four pybind11 torch-extension modules (`_gridencoder`, `_shencoder`, `_raymarching`, `_qgmls`) from
csrc/bindings.cpp into pienerf_b200/ext/*.so, linked against lib/libpienerf_b200.so (rpath $ORIGIN/../lib).

    python -m pienerf_b200.build_ext            (also called by __graft_entry__.build())

They carry the reference's module names and positional signatures (gridencoder/src/bindings.cpp:5-9, shencoder/src/
bindings.cpp:5-8, raymarching/src/bindings.cpp:5-19), so the reference's wrappers `import _gridencoder as _backend` etc. pick
them up (pienerf_b200.dropin).  The ctypes modules of the same names stay as the loader of last resort."""
import importlib.util
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "ext")
SRC = os.path.join(HERE, "csrc", "bindings.cpp")
LIBDIR = os.path.join(HERE, "lib")
NAMES = {"_gridencoder": "PN_EXT_GRIDENCODER", "_shencoder": "PN_EXT_SHENCODER", "_raymarching": "PN_EXT_RAYMARCHING", "_qgmls": "PN_EXT_QGMLS"}


def build(names=None, verbose=False, force=False):
    os.environ.setdefault("MAX_JOBS", "4")
    from torch.utils.cpp_extension import load
    os.makedirs(OUT, exist_ok=True)
    built = {}
    for name, macro in NAMES.items():
        if names and name not in names:
            continue
        so = os.path.join(OUT, name + ".so")
        deps = [SRC, os.path.join(HERE, "..", "include", "pienerf_b200.h")]
        if not force and os.path.exists(so) and all(os.path.getmtime(so) >= os.path.getmtime(d) for d in deps):
            built[name] = so
            continue
        bdir = os.path.join(OUT, "build_" + name)
        os.makedirs(bdir, exist_ok=True)
        load(name=name, sources=[SRC], build_directory=bdir, extra_cflags=["-O2", "-std=c++17", "-D" + macro], with_cuda=True,
             extra_ldflags=["-L" + LIBDIR, "-lpienerf_b200", "-Wl,-rpath,'$$ORIGIN/../lib'", "-Wl,-rpath," + LIBDIR], verbose=verbose, is_python_module=False)
        os.replace(os.path.join(bdir, name + ".so"), so)
        shutil.rmtree(bdir, ignore_errors=True)
        built[name] = so
    return built


def load_ext(name):
    """The compiled module `name` from pienerf_b200/ext (None if it was not built)."""
    so = os.path.join(OUT, name + ".so")
    if not os.path.exists(so):
        return None
    import torch  # noqa: F401  (libtorch first)
    spec = importlib.util.spec_from_file_location(name, so)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    print(build(sys.argv[1:] or None, verbose=True))
