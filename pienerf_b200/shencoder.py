"""Host-side mirror of shencoder/sphere_harmonics.py (SHEncoder) for the inference hot path."""
import torch
import torch.nn as nn

from . import _shencoder as _backend


@torch.no_grad()
def sh_encode(inputs, degree, calc_grad_inputs=False):
    """sphere_harmonics.py:14-39 forward: inputs [B,3] -> [B, degree**2] float32."""
    inputs = inputs.to(torch.float32).contiguous()
    B, input_dim = inputs.shape
    outputs = torch.empty(B, degree ** 2, dtype=inputs.dtype, device=inputs.device)
    dy_dx = torch.empty(B, input_dim * degree ** 2, dtype=inputs.dtype, device=inputs.device) if calc_grad_inputs else None
    _backend.sh_encode_forward(inputs, outputs, B, input_dim, degree, dy_dx)
    return outputs


class SHEncoder(nn.Module):
    """shencoder/sphere_harmonics.py:61-87."""

    def __init__(self, input_dim=3, degree=4):
        super().__init__()
        self.input_dim = input_dim
        self.degree = degree
        self.output_dim = degree ** 2
        assert self.input_dim == 3, "SH encoder only support input dim == 3"
        assert self.degree > 0 and self.degree <= 8, "SH encoder only supports degree in [1, 8]"

    def __repr__(self):
        return f"SHEncoder: input_dim={self.input_dim} degree={self.degree}"

    def forward(self, inputs, size=1):
        inputs = inputs / size
        prefix_shape = list(inputs.shape[:-1])
        inputs = inputs.reshape(-1, self.input_dim)
        outputs = sh_encode(inputs, self.degree, False)
        return outputs.reshape(prefix_shape + [self.output_dim])
