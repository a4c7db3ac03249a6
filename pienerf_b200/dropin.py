"""Make the B200 operators importable under the reference's extension-module names.

    import pienerf_b200.dropin as d; d.install()
    import gridencoder        # the reference's own gridencoder/grid.py now finds `_gridencoder` = ours

The reference wrappers do `import _gridencoder as _backend` (gridencoder/grid.py:9-12, shencoder/
sphere_harmonics.py:9-12, raymarching/raymarching.py:9-14) before falling back to a JIT build; installing
these four modules in sys.modules is all a maintainer needs.

Two interchangeable implementations of every module, both thin layers over the same C-ABI library:
  * compiled pybind11 torch extensions (csrc/bindings.cpp -> pienerf_b200/ext/*.so, built by build_ext.py / __graft_entry__.build())
    — what the reference ships (gridencoder/src/bindings.cpp etc.); preferred when present;
  * ctypes modules (pienerf_b200/_gridencoder.py ...) — the loader of last resort, no compiler needed beyond nvcc for the library."""
import sys

NAMES = ("_gridencoder", "_shencoder", "_raymarching", "_qgmls")


def install(compiled=None):
    """compiled=None: use the pybind11 modules when they are built, else ctypes; True: require them; False: ctypes only.
    Returns the module names; `installed` maps name -> "pybind11" | "ctypes"."""
    from . import _gridencoder, _qgmls, _raymarching, _shencoder
    from .build_ext import load_ext
    fallback = {"_gridencoder": _gridencoder, "_shencoder": _shencoder, "_raymarching": _raymarching, "_qgmls": _qgmls}
    installed.clear()
    for name in NAMES:
        mod = load_ext(name) if compiled is not False else None
        if mod is None:
            if compiled:
                raise ImportError(f"compiled module {name} not built: run `python -m pienerf_b200.build_ext`")
            mod = fallback[name]
        installed[name] = "pybind11" if mod is not fallback[name] else "ctypes"
        sys.modules[name] = mod
    return NAMES


installed = {}
