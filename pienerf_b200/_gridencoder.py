"""`_gridencoder` — same entry points as the reference pybind module (gridencoder/src/bindings.cpp:5-9),
backed by the C-ABI.  Positional signatures match gridencoder/src/gridencoder.h:12-15, so the reference's
own gridencoder/grid.py runs unmodified against this module (see pienerf_b200.dropin)."""
import torch

from . import _lib
from ._lib import check, dptr, lib, stream_ptr


def grid_encode_forward(inputs, embeddings, offsets, outputs, B, D, C, L, S, H, dy_dx, gridtype, align_corners, interp):
    # reference checks: CUDA + contiguous + dtype (gridencoder.cu:449-465)
    _lib._chk(inputs, "inputs", (torch.float32, torch.float16, torch.float64))
    _lib._chk(embeddings, "embeddings", (torch.float32, torch.float16, torch.float64))
    _lib._chk(offsets, "offsets", torch.int32)
    _lib._chk(outputs, "outputs", (torch.float32, torch.float16, torch.float64))
    if inputs.dtype != torch.float32:
        raise RuntimeError("inputs must be float32 (the reference reads inputs.data_ptr<float>())")
    if embeddings.dtype == torch.float64:
        raise NotImplementedError("double embeddings are not part of the B200 hot path")
    if outputs.dtype != embeddings.dtype:
        raise RuntimeError("outputs must have the dtype of embeddings")
    if int(C) not in (1, 2, 4, 8):
        raise RuntimeError("GridEncoding: C must be 1, 2, 4, or 8.")
    check(lib.pn_grid_encode_forward(dptr(inputs), dptr(embeddings), dptr(offsets), dptr(outputs), int(B), int(D), int(C),
                                     int(L), float(S), int(H), dptr(dy_dx), int(gridtype), int(bool(align_corners)),
                                     int(interp), int(embeddings.dtype == torch.float16), stream_ptr()))


def _table_dtype(embeddings, what):
    if embeddings.dtype == torch.float64:
        raise NotImplementedError(f"{what}: double embeddings are not provided by the B200 library")
    return embeddings.dtype


def grid_encode_backward(grad, inputs, embeddings, offsets, grad_embeddings, B, D, C, L, S, H, dy_dx, grad_inputs, gridtype,
                         align_corners, interp):
    # reference checks: CUDA + contiguous + dtype of the five mandatory tensors (gridencoder.cu:474-494)
    fl = (torch.float32, torch.float16, torch.float64)
    _lib._chk(grad, "grad", fl)
    _lib._chk(inputs, "inputs", fl)
    _lib._chk(embeddings, "embeddings", fl)
    _lib._chk(offsets, "offsets", torch.int32)
    _lib._chk(grad_embeddings, "grad_embeddings", fl)
    dt = _table_dtype(embeddings, "grid_encode_backward")
    if inputs.dtype != torch.float32:
        raise RuntimeError("inputs must be float32 (the reference reads inputs.data_ptr<float>())")
    for t, n in ((grad, "grad"), (grad_embeddings, "grad_embeddings"), (dy_dx, "dy_dx"), (grad_inputs, "grad_inputs")):
        if t is not None and t.dtype != dt:
            raise RuntimeError(f"{n} must have the dtype of embeddings")
    if (dy_dx is None) != (grad_inputs is None):
        raise RuntimeError("dy_dx and grad_inputs go together")
    if int(C) not in (1, 2, 4, 8):
        raise RuntimeError("GridEncoding: C must be 1, 2, 4, or 8.")
    check(lib.pn_grid_encode_backward(dptr(grad), dptr(inputs), dptr(embeddings), dptr(offsets), dptr(grad_embeddings), int(B),
                                      int(D), int(C), int(L), float(S), int(H), dptr(dy_dx), dptr(grad_inputs), int(gridtype),
                                      int(bool(align_corners)), int(interp), int(dt == torch.float16), stream_ptr()))


def grad_total_variation(inputs, embeddings, grad, offsets, weight, B, D, C, L, S, H, gridtype, align_corners):
    fl = (torch.float32, torch.float16, torch.float64)
    _lib._chk(inputs, "inputs", fl)
    _lib._chk(embeddings, "embeddings", fl)
    _lib._chk(grad, "grad", fl)
    _lib._chk(offsets, "offsets", torch.int32)
    dt = _table_dtype(embeddings, "grad_total_variation")
    if inputs.dtype != dt or grad.dtype != dt:       # gridencoder.cu:641-644 reads all three through embeddings' scalar_t
        raise RuntimeError("inputs and grad must have the dtype of embeddings")
    if int(C) not in (1, 2, 4, 8):
        raise RuntimeError("GridEncoding: C must be 1, 2, 4, or 8.")
    check(lib.pn_grad_total_variation(dptr(inputs), dptr(embeddings), dptr(grad), dptr(offsets), float(weight), int(B), int(D),
                                      int(C), int(L), float(S), int(H), int(gridtype), int(bool(align_corners)),
                                      int(dt == torch.float16), stream_ptr()))
