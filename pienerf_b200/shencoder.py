"""Direction encoder on top of the `_shencoder` entry points: the `SHEncoder` module and the `sh_encode` function with the
interface of the reference's shencoder/sphere_harmonics.py (constructor arguments, attributes, `forward(inputs, size)`,
gradients with respect to the directions when they require grad)."""
import torch
from torch import nn
from torch.amp import custom_bwd, custom_fwd

from . import _shencoder as _backend

_MAX_DEGREE = 8


def _run_forward(directions, degree, with_jacobian):
    """One kernel call: [B,3] float32 directions -> ([B, degree^2] harmonics, [B, 3*degree^2] Jacobian or None)."""
    n, dim = directions.shape
    width = degree * degree
    new = dict(dtype=torch.float32, device=directions.device)
    harmonics = torch.empty((n, width), **new)
    jacobian = torch.empty((n, dim * width), **new) if with_jacobian else None
    _backend.sh_encode_forward(directions, harmonics, n, dim, degree, jacobian)
    return harmonics, jacobian


class _sh_encoder(torch.autograd.Function):
    """Autograd node of sphere_harmonics.py:14-56.  The Jacobian is produced by the forward kernel only when a gradient with
    respect to the directions can be asked for; without it backward has nothing to return."""

    @staticmethod
    @custom_fwd(device_type="cuda", cast_inputs=torch.float32)        # fp32 even under autocast (sphere_harmonics.py:15)
    def forward(ctx, inputs, degree, calc_grad_inputs=False):
        directions = inputs.to(torch.float32).contiguous()
        harmonics, jacobian = _run_forward(directions, int(degree), bool(calc_grad_inputs))
        ctx.degree = int(degree)
        ctx.save_for_backward(directions, jacobian)
        return harmonics

    @staticmethod
    @custom_bwd(device_type="cuda")
    def backward(ctx, grad_harmonics):
        directions, jacobian = ctx.saved_tensors
        grad_directions = None
        if jacobian is not None:
            n, dim = directions.shape
            grad_directions = torch.zeros_like(directions)              # the kernel accumulates into it
            _backend.sh_encode_backward(grad_harmonics.to(torch.float32).contiguous(), directions, n, dim, ctx.degree, jacobian,
                                        grad_directions)
        return grad_directions, None, None


def sh_encode(inputs, degree, calc_grad_inputs=False):
    """The reference exports `_sh_encoder.apply` under this name (sphere_harmonics.py:58); keyword arguments work here."""
    return _sh_encoder.apply(inputs, degree, calc_grad_inputs)


class SHEncoder(nn.Module):
    """Real spherical harmonics of unit directions, bands 0..degree-1 (shencoder/sphere_harmonics.py:61-87)."""

    def __init__(self, input_dim=3, degree=4):
        super().__init__()
        assert input_dim == 3, "SH encoder only support input dim == 3"
        assert 0 < degree <= _MAX_DEGREE, "SH encoder only supports degree in [1, 8]"
        self.input_dim, self.degree, self.output_dim = input_dim, degree, degree * degree

    def __repr__(self):
        return f"SHEncoder: input_dim={self.input_dim} degree={self.degree}"

    def forward(self, inputs, size=1):
        scaled = inputs / size                                           # positions in [-size, size] -> [-1, 1]
        flat = scaled.reshape(-1, self.input_dim)
        encoded = sh_encode(flat, self.degree, flat.requires_grad)
        return encoded.reshape(*scaled.shape[:-1], self.output_dim)
