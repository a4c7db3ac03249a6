"""`_shencoder` — same entry points as the reference pybind module (shencoder/src/bindings.cpp:5-8)."""
import torch

from . import _lib
from ._lib import check, dptr, lib, stream_ptr


def sh_encode_forward(inputs, outputs, B, D, C, dy_dx):
    _lib._chk(inputs, "inputs", (torch.float32, torch.float16, torch.float64))
    _lib._chk(outputs, "outputs", (torch.float32, torch.float16, torch.float64))
    if inputs.dtype != torch.float32 or outputs.dtype != torch.float32:
        raise NotImplementedError("the B200 SH encoder is fp32 (the hot path casts inputs to float32, sphere_harmonics.py:15)")
    check(lib.pn_sh_encode_forward(dptr(inputs), dptr(outputs), int(B), int(D), int(C), dptr(dy_dx), stream_ptr()))


def sh_encode_backward(*args):
    check(lib.pn_sh_encode_backward())
