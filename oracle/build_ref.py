"""oracle/build_ref.py -- TEST INFRASTRUCTURE ONLY.

Compiles the REFERENCE's own CUDA extensions (gridencoder, shencoder, raymarching) for sm_100a straight from
the sources where they lie under /root/reference (nothing is copied into this repo) into oracle/_ref/*.so
(git-ignored, but shipped to the GPU box by gpurun).  They are the GPU oracle that pins oracle/render_oracle.py
and the CUDA path, and the `--impl reference` arm of bench.py.

The reference's shipped flags (-std=c++14) are rejected by torch >= 2.1 headers; both host and nvcc get
-std=c++17.  Run here (no GPU needed: nvcc cross-compiles); /root/reference does not exist on
the GPU box, where only the prebuilt .so files are used.
"""
import os
import sys

REF = os.environ.get("PIENERF_REFERENCE", "/root/reference")
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")

EXTS = {
    "_ref_gridencoder": ["gridencoder/src/gridencoder.cu", "gridencoder/src/bindings.cpp"],
    "_ref_shencoder": ["shencoder/src/shencoder.cu", "shencoder/src/bindings.cpp"],
    "_ref_raymarching": ["raymarching/src/raymarching.cu", "raymarching/src/bindings.cpp"],
}


def build(names=None, verbose=False):
    if not os.path.isdir(REF):
        return {}
    os.environ.setdefault("TORCH_CUDA_ARCH_LIST", "10.0")
    os.environ.setdefault("MAX_JOBS", "4")
    from torch.utils.cpp_extension import load
    built = {}
    for name, srcs in EXTS.items():
        if names and name not in names:
            continue
        so = os.path.join(OUT, name + ".so")
        if os.path.exists(so):
            built[name] = so
            continue
        bdir = os.path.join(OUT, "build_" + name)
        os.makedirs(bdir, exist_ok=True)
        load(name=name, sources=[os.path.join(REF, s) for s in srcs], build_directory=bdir,
             extra_cflags=["-O3", "-std=c++17"],
             extra_cuda_cflags=["-O3", "-std=c++17", "-U__CUDA_NO_HALF_OPERATORS__", "-U__CUDA_NO_HALF_CONVERSIONS__",
                                "-U__CUDA_NO_HALF2_OPERATORS__", "-gencode", "arch=compute_100a,code=sm_100a"],
             verbose=verbose, is_python_module=False)
        os.replace(os.path.join(bdir, name + ".so"), so)
        import shutil
        shutil.rmtree(bdir, ignore_errors=True)           # keep only the .so: gpurun ships oracle/_ref to the GPU box
        built[name] = so
    return built


def load_ref(name):
    """Import a prebuilt reference extension from oracle/_ref (None if it was never built)."""
    so = os.path.join(OUT, name + ".so")
    if not os.path.exists(so):
        return None
    import importlib.util
    import torch  # noqa: F401  (libtorch must be loaded first)
    spec = importlib.util.spec_from_file_location(name, so)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    print(build(sys.argv[1:] or None, verbose=True))
