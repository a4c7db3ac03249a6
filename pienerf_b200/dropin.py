"""Make the B200 operators importable under the reference's extension-module names.

    import pienerf_b200.dropin as d; d.install()
    import gridencoder        # the reference's own gridencoder/grid.py now finds `_gridencoder` = ours

The reference wrappers do `import _gridencoder as _backend` (gridencoder/grid.py:9-12, shencoder/
sphere_harmonics.py:9-12, raymarching/raymarching.py:9-14) before falling back to a JIT build; installing
these four modules in sys.modules is all a maintainer needs."""
import sys


def install():
    from . import _gridencoder, _qgmls, _raymarching, _shencoder
    sys.modules["_gridencoder"] = _gridencoder
    sys.modules["_shencoder"] = _shencoder
    sys.modules["_raymarching"] = _raymarching
    sys.modules["_qgmls"] = _qgmls
    return ("_gridencoder", "_shencoder", "_raymarching", "_qgmls")
