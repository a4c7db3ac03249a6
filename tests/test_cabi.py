"""No-GPU checks of the drop-in boundary: the C-ABI library loads and exports exactly what include/*.h declares;
training entry points validate their arguments without a GPU; the drop-in module names resolve for the reference's wrappers."""
import ctypes
import os
import re
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "pienerf_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(pn_\w+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from pienerf_b200 import _lib
    names = _declared()
    assert len(names) >= 35
    raw = ctypes.CDLL(_lib.LIB_PATH)
    for n in names:
        assert hasattr(raw, n), f"{n} declared in include/pienerf_b200.h but not exported"
    assert sorted(_lib.EXPORTS) == names, set(names) ^ set(_lib.EXPORTS)
    assert _lib.lib.pn_version() >= 100


def test_header_cites_reference_for_every_entry_point():
    src = open(os.path.join(ROOT, "include", "pienerf_b200.h")).read()
    for f in ("gridencoder.cu", "shencoder.cu", "raymarching.cu", "cuda_utils.py", "solver.py", "renderer.py", "nerf/utils.py"):
        assert f in src, f


def test_training_entry_points_validate_without_gpu():
    """The training entry points are real kernels now: without a GPU they must refuse CPU tensors the way the
    reference's CHECK_CUDA does, and the raw C-ABI must refuse null pointers instead of launching."""
    import torch
    from pienerf_b200 import _gridencoder, _lib, _raymarching, _shencoder
    x = torch.zeros(4, 3); e = torch.zeros(8, 2); o = torch.zeros(2, dtype=torch.int32); g = torch.zeros(1, 4, 2)
    with pytest.raises(RuntimeError, match="must be a CUDA tensor"):
        _gridencoder.grid_encode_backward(g, x, e, o, e.clone(), 4, 3, 2, 1, 1.0, 16, None, None, 0, False, 0)
    with pytest.raises(RuntimeError, match="must be a CUDA tensor"):
        _gridencoder.grad_total_variation(x, e, e.clone(), o, 1.0, 4, 3, 2, 1, 1.0, 16, 0, False)
    with pytest.raises(RuntimeError, match="must be a CUDA tensor"):
        _shencoder.sh_encode_backward(torch.zeros(4, 16), x, 4, 3, 4, torch.zeros(4, 48), torch.zeros(4, 3))
    r = torch.zeros(4, 3, dtype=torch.int32)
    with pytest.raises(RuntimeError, match="must be a CUDA tensor"):
        _raymarching.march_rays_train(x, x, torch.zeros(8, dtype=torch.uint8), 1.0, 0.0, 16, 4, 1, 128, 64, x[:, 0], x[:, 0], x, x, x[:, :2], r,
                                      torch.zeros(2, dtype=torch.int32), x[:, 0])
    with pytest.raises(RuntimeError, match="must be a CUDA tensor"):
        _raymarching.composite_rays_train_forward(x[:, 0], x, x[:, :2], r, 4, 4, 1e-4, x[:, 0], x[:, 0], x)
    with pytest.raises(RuntimeError, match="must be a CUDA tensor"):
        _raymarching.composite_rays_train_backward(x[:, 0], x, x[:, 0], x, x[:, :2], r, x[:, 0], x, 4, 4, 1e-4, x[:, 0], x)
    vp = lambda: _lib.vp(0)  # noqa: E731
    assert _lib.lib.pn_grid_encode_backward(vp(), vp(), vp(), vp(), vp(), 4, 3, 2, 1, 1.0, 16, vp(), vp(), 0, 0, 0, 0, vp()) == -1   # PN_EINVAL
    assert "null pointer" in _lib.last_error()
    assert _lib.lib.pn_march_rays_train(vp(), vp(), vp(), 1.0, 0.0, 16, 4, 1, 128, 64, vp(), vp(), vp(), vp(), vp(), vp(), vp(), vp(), vp()) == -1
    assert _lib.lib.pn_sh_encode_backward(vp(), vp(), 4, 3, 4, vp(), vp(), vp()) == -1


def test_argument_validation_without_gpu():
    import torch
    from pienerf_b200 import _gridencoder, _shencoder
    x = torch.zeros(4, 3)
    with pytest.raises(RuntimeError, match="must be a CUDA tensor"):
        _gridencoder.grid_encode_forward(x, x, x.int(), x, 4, 3, 2, 1, 1.0, 16, None, 0, False, 0)
    with pytest.raises(RuntimeError, match="must be a CUDA tensor"):
        _shencoder.sh_encode_forward(x, x, 4, 3, 4, None)


def test_compiled_extension_modules_mirror_the_reference_bindings():
    """The pybind11 modules (csrc/bindings.cpp) export the entry points of the reference's bindings.cpp files and raise the
    same exception types; no compute without a GPU."""
    import torch
    from pienerf_b200.build_ext import load_ext
    want = {"_gridencoder": ["grid_encode_forward", "grid_encode_backward", "grad_total_variation"],          # gridencoder/src/bindings.cpp:5-9
            "_shencoder": ["sh_encode_forward", "sh_encode_backward"],                                         # shencoder/src/bindings.cpp:5-8
            "_raymarching": ["near_far_from_aabb", "sph_from_ray", "morton3D", "morton3D_invert", "packbits", "march_rays_train",
                             "composite_rays_train_forward", "composite_rays_train_backward", "march_rays", "march_rays_quadratic_bending",
                             "composite_rays"],                                                                # raymarching/src/bindings.cpp:5-19
            "_qgmls": ["shape_functions", "collect_param", "build_ip_global", "build_pin_global", "collect_gravity", "build_rhs", "matvec3", "step",
                       "ip_info", "update_pos", "update_force"]}
    mods = {n: load_ext(n) for n in want}
    if any(m is None for m in mods.values()):
        pytest.skip("compiled modules not built (python -m pienerf_b200.build_ext)")
    for n, fns in want.items():
        for f in fns:
            assert callable(getattr(mods[n], f)), (n, f)
    x = torch.zeros(4, 3)
    with pytest.raises(TypeError):
        mods["_raymarching"].march_rays_train()                                                                # real entry point: positional signature enforced
    with pytest.raises(RuntimeError, match="must be a CUDA tensor"):
        mods["_gridencoder"].grid_encode_backward(torch.zeros(1, 4, 2), x, torch.zeros(8, 2), torch.zeros(2, dtype=torch.int32), torch.zeros(8, 2),
                                                  4, 3, 2, 1, 1.0, 16, None, None, 0, False, 0)
    with pytest.raises(RuntimeError, match="must be a CUDA tensor"):
        mods["_shencoder"].sh_encode_backward(torch.zeros(4, 16), x, 4, 3, 4, torch.zeros(4, 48), torch.zeros(4, 3))
    x = torch.zeros(4, 3)
    with pytest.raises(RuntimeError, match="must be a CUDA tensor"):
        mods["_gridencoder"].grid_encode_forward(x, x, x.int(), x, 4, 3, 2, 1, 1.0, 16, None, 0, False, 0)
    with pytest.raises(RuntimeError, match="must be a CUDA tensor"):
        mods["_shencoder"].sh_encode_forward(x, x, 4, 3, 4, None)
    with pytest.raises(TypeError):
        mods["_raymarching"].composite_rays(1, 2)                                                              # positional signature enforced by pybind11


@pytest.mark.parametrize("compiled", [False, None])
def test_dropin_names_and_reference_wrappers_import(compiled):
    """compiled=None picks the pybind11 modules when pienerf_b200/ext is built (otherwise ctypes); False forces ctypes."""
    from pienerf_b200 import dropin
    names = dropin.install(compiled=compiled)
    if compiled is False:
        assert set(dropin.installed.values()) == {"ctypes"}
    assert names == ("_gridencoder", "_shencoder", "_raymarching", "_qgmls")
    import _gridencoder
    import _raymarching
    import _shencoder
    # every function the reference's bindings export (bindings.cpp of the three extensions)
    for n in ("grid_encode_forward", "grid_encode_backward", "grad_total_variation"):
        assert callable(getattr(_gridencoder, n))
    for n in ("sh_encode_forward", "sh_encode_backward"):
        assert callable(getattr(_shencoder, n))
    for n in ("near_far_from_aabb", "sph_from_ray", "morton3D", "morton3D_invert", "packbits", "march_rays_train",
              "composite_rays_train_forward", "composite_rays_train_backward", "march_rays", "march_rays_quadratic_bending",
              "composite_rays"):
        assert callable(getattr(_raymarching, n))
    ref = "/root/reference"
    if not os.path.isdir(ref):
        pytest.skip("reference tree not present on this box")
    sys.path.insert(0, ref)
    try:
        for m in ("gridencoder", "shencoder", "raymarching"):
            sys.modules.pop(m, None)
        import gridencoder.grid as g        # the reference's own wrapper, unmodified
        import raymarching.raymarching as r
        import shencoder.sphere_harmonics as s
        assert g._backend is _gridencoder and s._backend is _shencoder and r._backend is _raymarching
    finally:
        sys.path.remove(ref)
        for m in list(sys.modules):
            if m.split(".")[0] in ("gridencoder", "shencoder", "raymarching") and "pienerf_b200" not in m:
                sys.modules.pop(m, None)


def test_product_never_imports_the_oracle():
    """The oracle is test infrastructure: nothing under pienerf_b200/ may import, link or execute it."""
    pkg = os.path.join(ROOT, "pienerf_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dp, f)).read()
                assert not re.search(r"^\s*(import|from)\s+oracle\b", txt, flags=re.M), f
                assert "sim_oracle" not in txt and "render_oracle" not in txt and "_ref_" not in txt, f


def test_pass_count_and_workspace_are_host_only_queries():
    """pn_render_pass_count / pn_render_workspace_bytes need no GPU: pass caps 32, 64, ... cover max_steps, plus one spare pass."""
    from pienerf_b200._lib import lib
    want = {48: 3, 256: 5, 300: 5, 1024: 7, 4096: 8}      # 32+64 | 32..256 = 480 | 32..1024 = 2016 | capped at kMaxPass
    for max_steps, n in want.items():
        assert lib.pn_render_pass_count(max_steps) == n, (max_steps, lib.pn_render_pass_count(max_steps))
    small = lib.pn_render_workspace_bytes(1000, 100, 1.0, 0.06)
    big = lib.pn_render_workspace_bytes(640000, 2028, 1.0, 0.06)
    assert 0 < small < big < (2 << 30)


def test_header_is_plain_c_and_the_library_links_from_c(tmp_path):
    """include/pienerf_b200.h compiles as C99 (no torch, no C++), and a C program linked against the library reaches the entry
    points: version, host-only queries, and argument refusal (PN_EINVAL + pn_last_error) of compute calls given null pointers."""
    import shutil
    import subprocess
    from pienerf_b200 import _lib
    if shutil.which("gcc") is None:
        pytest.skip("no gcc")
    src = tmp_path / "abi.c"
    src.write_text(r'''
#include <stdio.h>
#include <string.h>
#include "pienerf_b200.h"
int main(void) {
    if (pn_version() < 100) return 1;
    if (pn_render_pass_count(1024) != 7) return 2;
    if (pn_grid_encode_backward(0, 0, 0, 0, 0, 4, 3, 2, 1, 1.0f, 16, 0, 0, 0, 0, 0, 0, 0) != PN_EINVAL) return 3;
    if (!strstr(pn_last_error(), "null pointer")) return 4;
    if (pn_march_rays_train(0, 0, 0, 1.0f, 0.0f, 16, 4, 1, 128, 64, 0, 0, 0, 0, 0, 0, 0, 0, 0) != PN_EINVAL) return 5;
    if (pn_march_rays_train(0, 0, 0, 1.0f, 0.0f, 16, 0, 1, 128, 64, 0, 0, 0, 0, 0, 0, 0, 0, 0) != PN_OK) return 6;   /* N = 0: nothing to do */
    if (pn_composite_rays_train_forward(0, 0, 0, 0, 0, 0, 1e-4f, 0, 0, 0, 0) != PN_OK) return 7;
    if (pn_set_train_block_skip(1) != 1 || pn_set_train_write_mode(1) != 1) return 8;
    printf("ok %d\n", pn_version());
    return 0;
}
''')
    exe = tmp_path / "abi"
    libdir = os.path.dirname(_lib.LIB_PATH)
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe),
                    "-L", libdir, "-lpienerf_b200", f"-Wl,-rpath,{libdir}"], check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True)
    assert out.returncode == 0 and out.stdout.startswith("ok"), (out.returncode, out.stdout, out.stderr)
