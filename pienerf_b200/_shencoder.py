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


def sh_encode_backward(grad, inputs, B, D, C, dy_dx, grad_inputs):
    f32 = torch.float32
    for t, n in ((grad, "grad"), (inputs, "inputs"), (dy_dx, "dy_dx"), (grad_inputs, "grad_inputs")):   # shencoder.cu:420-433
        _lib._chk(t, n, (torch.float32, torch.float16, torch.float64))
        if t.dtype != f32:
            raise NotImplementedError("the B200 SH encoder is fp32 (sphere_harmonics.py:15 casts inputs to float32)")
    check(lib.pn_sh_encode_backward(dptr(grad), dptr(inputs), int(B), int(D), int(C), dptr(dy_dx), dptr(grad_inputs), stream_ptr()))
