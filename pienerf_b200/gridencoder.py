"""Host-side mirror of gridencoder/grid.py (GridEncoder, grid_encode) for the inference hot path.

Same constructor arguments, attributes (`offsets`, `embeddings`, `per_level_scale`, ...) and call
semantics as the reference module; forward only (backward/TV are training-side)."""
import numpy as np
import torch
import torch.nn as nn

from . import _gridencoder as _backend

_gridtype_to_id = {"hash": 0, "tiled": 1}
_interp_to_id = {"linear": 0, "smoothstep": 1}


@torch.no_grad()
def grid_encode(inputs, embeddings, offsets, per_level_scale, base_resolution, calc_grad_inputs=False, gridtype=0,
                align_corners=False, interpolation=0, level_major=False):
    """grid.py:24-63 forward: inputs [B,D] in [0,1] -> [B, L*C] (or the kernel's native [L,B,C] if level_major)."""
    inputs = inputs.to(torch.float32).contiguous()
    B, D = inputs.shape
    L = offsets.shape[0] - 1
    C = embeddings.shape[1]
    S = np.log2(per_level_scale)
    if torch.is_autocast_enabled() and C % 2 == 0:           # grid.py:43-44
        embeddings = embeddings.to(torch.half)
    outputs = torch.empty(L, B, C, device=inputs.device, dtype=embeddings.dtype)
    dy_dx = torch.empty(B, L * D * C, device=inputs.device, dtype=embeddings.dtype) if calc_grad_inputs else None
    _backend.grid_encode_forward(inputs, embeddings.contiguous(), offsets, outputs, B, D, C, L, S, base_resolution, dy_dx,
                                 gridtype, align_corners, interpolation)
    if level_major:
        return outputs
    return outputs.permute(1, 0, 2).reshape(B, L * C)          # grid.py:57


class GridEncoder(nn.Module):
    """gridencoder/grid.py:96-161."""

    def __init__(self, input_dim=3, num_levels=16, level_dim=2, per_level_scale=2, base_resolution=16,
                 log2_hashmap_size=19, desired_resolution=None, gridtype="hash", align_corners=False,
                 interpolation="linear"):
        super().__init__()
        if desired_resolution is not None:
            per_level_scale = np.exp2(np.log2(desired_resolution / base_resolution) / (num_levels - 1))
        self.input_dim = input_dim
        self.num_levels = num_levels
        self.level_dim = level_dim
        self.per_level_scale = per_level_scale
        self.log2_hashmap_size = log2_hashmap_size
        self.base_resolution = base_resolution
        self.output_dim = num_levels * level_dim
        self.gridtype = gridtype
        self.gridtype_id = _gridtype_to_id[gridtype]
        self.interpolation = interpolation
        self.interp_id = _interp_to_id[interpolation]
        self.align_corners = align_corners

        offsets, offset = [], 0
        self.max_params = 2 ** log2_hashmap_size
        for i in range(num_levels):
            resolution = int(np.ceil(base_resolution * per_level_scale ** i))
            params_in_level = min(self.max_params, (resolution if align_corners else resolution + 1) ** input_dim)
            params_in_level = int(np.ceil(params_in_level / 8) * 8)
            offsets.append(offset)
            offset += params_in_level
        offsets.append(offset)
        self.register_buffer("offsets", torch.from_numpy(np.array(offsets, dtype=np.int32)))
        self.n_params = offsets[-1] * level_dim
        self.embeddings = nn.Parameter(torch.empty(offset, level_dim))
        self.reset_parameters()

    def reset_parameters(self):
        self.embeddings.data.uniform_(-1e-4, 1e-4)

    def __repr__(self):
        return (f"GridEncoder: input_dim={self.input_dim} num_levels={self.num_levels} level_dim={self.level_dim} "
                f"resolution={self.base_resolution} -> {int(round(self.base_resolution * self.per_level_scale ** (self.num_levels - 1)))} "
                f"per_level_scale={self.per_level_scale:.4f} params={tuple(self.embeddings.shape)} gridtype={self.gridtype} "
                f"align_corners={self.align_corners} interpolation={self.interpolation}")

    def forward(self, inputs, bound=1):
        inputs = (inputs + bound) / (2 * bound)
        prefix_shape = list(inputs.shape[:-1])
        inputs = inputs.view(-1, self.input_dim)
        outputs = grid_encode(inputs, self.embeddings, self.offsets, self.per_level_scale, self.base_resolution,
                              False, self.gridtype_id, self.align_corners, self.interp_id)
        return outputs.view(prefix_shape + [self.output_dim])

    def grad_total_variation(self, *args, **kwargs):
        _backend.grad_total_variation()
