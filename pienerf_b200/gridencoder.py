"""Multiresolution hash-grid encoder on top of the `_gridencoder` entry points.

`GridEncoder` and `grid_encode` keep the interface of the reference's gridencoder/grid.py — constructor arguments, the
attributes other code reads (`offsets`, `embeddings`, `per_level_scale`, `output_dim`, `gridtype_id`, ...), `forward(inputs,
bound)`, gradients for the table and (when the inputs require grad) the inputs, `grad_total_variation` — so a checkpoint or a
caller written against the reference works unchanged.  Layout conventions of the kernels: inputs [B, D] in [0, 1], table
[entries, C], outputs level-major [L, B, C] (returned as [B, L*C] unless `level_major`), Jacobian [B, L*D*C]."""
from dataclasses import dataclass

import numpy as np
import torch
from torch import nn
from torch.amp import custom_bwd, custom_fwd

from . import _gridencoder as _backend

_gridtype_to_id = {"hash": 0, "tiled": 1}
_interp_to_id = {"linear": 0, "smoothstep": 1}


@dataclass
class _Geometry:
    """What every kernel call needs besides the tensors (argument order of gridencoder.h:12-15)."""
    B: int
    D: int
    C: int
    L: int
    S: float          # log2 of the per-level scale (the kernels evaluate exp2f(level * S))
    H: int            # base resolution
    gridtype: int
    align_corners: bool
    interpolation: int

    def head(self):
        return (self.B, self.D, self.C, self.L, self.S, self.H)


def level_offsets(input_dim, num_levels, per_level_scale, base_resolution, log2_hashmap_size, align_corners):
    """First entry of each level in the table (+ the total), grid.py:110-122: a level holds min(2^log2_hashmap_size, its dense
    vertex count) entries, rounded up to a multiple of 8."""
    cap = 2 ** log2_hashmap_size
    starts = [0]
    for level in range(num_levels):
        res = int(np.ceil(base_resolution * per_level_scale ** level))
        verts = (res if align_corners else res + 1) ** input_dim
        starts.append(starts[-1] + int(np.ceil(min(cap, verts) / 8) * 8))
    return np.asarray(starts, dtype=np.int32)


class _grid_encode(torch.autograd.Function):
    """Autograd node of grid.py:24-91."""

    @staticmethod
    @custom_fwd(device_type="cuda")
    def forward(ctx, inputs, embeddings, offsets, per_level_scale, base_resolution, calc_grad_inputs=False, gridtype=0,
                align_corners=False, interpolation=0, level_major=False):
        x = inputs.to(torch.float32).contiguous()                   # positions stay fp32 whatever the table is
        table = embeddings
        if torch.is_autocast_enabled() and table.shape[1] % 2 == 0:  # fp16 table under autocast, pairs only (grid.py:43-44)
            table = table.to(torch.half)
        table = table.contiguous()
        g = _Geometry(B=x.shape[0], D=x.shape[1], C=table.shape[1], L=offsets.shape[0] - 1, S=float(np.log2(per_level_scale)),
                      H=base_resolution, gridtype=gridtype, align_corners=align_corners, interpolation=interpolation)
        like = dict(device=x.device, dtype=table.dtype)
        features = torch.empty((g.L, g.B, g.C), **like)
        jacobian = torch.empty((g.B, g.L * g.D * g.C), **like) if calc_grad_inputs else None
        _backend.grid_encode_forward(x, table, offsets, features, *g.head(), jacobian, g.gridtype, g.align_corners, g.interpolation)
        ctx.geometry, ctx.level_major = g, bool(level_major)
        ctx.save_for_backward(x, table, offsets, jacobian)
        if level_major:                                              # the kernels' native layout (the fused field reads it)
            return features
        return features.permute(1, 0, 2).reshape(g.B, g.L * g.C)

    @staticmethod
    @custom_bwd(device_type="cuda")
    def backward(ctx, grad_features):
        x, table, offsets, jacobian = ctx.saved_tensors
        g = ctx.geometry
        if not ctx.level_major:                                      # [B, L*C] -> [L, B, C]
            grad_features = grad_features.reshape(g.B, g.L, g.C).permute(1, 0, 2)
        grad_features = grad_features.to(table.dtype).contiguous()
        grad_table = torch.zeros_like(table)                         # the kernel reduces into it
        grad_x = torch.empty_like(x, dtype=table.dtype) if jacobian is not None else None
        _backend.grid_encode_backward(grad_features, x, table, offsets, grad_table, *g.head(), jacobian, grad_x, g.gridtype,
                                      g.align_corners, g.interpolation)
        if grad_x is not None:
            grad_x = grad_x.to(x.dtype)
        return (grad_x, grad_table) + (None,) * 8


def grid_encode(inputs, embeddings, offsets, per_level_scale, base_resolution, calc_grad_inputs=False, gridtype=0,
                align_corners=False, interpolation=0, level_major=False):
    """The reference exports `_grid_encode.apply` under this name (grid.py:93); keyword arguments work here."""
    return _grid_encode.apply(inputs, embeddings, offsets, per_level_scale, base_resolution, calc_grad_inputs, gridtype,
                              align_corners, interpolation, level_major)


class GridEncoder(nn.Module):
    """gridencoder/grid.py:96-190."""

    def __init__(self, input_dim=3, num_levels=16, level_dim=2, per_level_scale=2, base_resolution=16,
                 log2_hashmap_size=19, desired_resolution=None, gridtype="hash", align_corners=False,
                 interpolation="linear"):
        super().__init__()
        if desired_resolution is not None:                           # the finest level's resolution fixes the growth factor
            per_level_scale = np.exp2(np.log2(desired_resolution / base_resolution) / (num_levels - 1))
        self.input_dim, self.num_levels, self.level_dim = input_dim, num_levels, level_dim
        self.per_level_scale, self.base_resolution = per_level_scale, base_resolution
        self.log2_hashmap_size = log2_hashmap_size
        self.max_params = 2 ** log2_hashmap_size
        self.output_dim = num_levels * level_dim
        self.gridtype, self.gridtype_id = gridtype, _gridtype_to_id[gridtype]
        self.interpolation, self.interp_id = interpolation, _interp_to_id[interpolation]
        self.align_corners = align_corners
        starts = level_offsets(input_dim, num_levels, per_level_scale, base_resolution, log2_hashmap_size, align_corners)
        self.register_buffer("offsets", torch.from_numpy(starts))
        self.n_params = int(starts[-1]) * level_dim
        self.embeddings = nn.Parameter(torch.empty(int(starts[-1]), level_dim))
        self.reset_parameters()

    def reset_parameters(self):
        self.embeddings.data.uniform_(-1e-4, 1e-4)

    def __repr__(self):
        finest = int(round(self.base_resolution * self.per_level_scale ** (self.num_levels - 1)))
        return (f"GridEncoder: input_dim={self.input_dim} num_levels={self.num_levels} level_dim={self.level_dim} "
                f"resolution={self.base_resolution} -> {finest} per_level_scale={self.per_level_scale:.4f} "
                f"params={tuple(self.embeddings.shape)} gridtype={self.gridtype} align_corners={self.align_corners} "
                f"interpolation={self.interpolation}")

    def _unit_cube(self, positions, bound):
        return (positions + bound) / (2 * bound)                     # [-bound, bound] -> [0, 1]

    def forward(self, inputs, bound=1):
        unit = self._unit_cube(inputs, bound)
        flat = unit.view(-1, self.input_dim)
        features = grid_encode(flat, self.embeddings, self.offsets, self.per_level_scale, self.base_resolution,
                               flat.requires_grad, self.gridtype_id, self.align_corners, self.interp_id)
        return features.view(*unit.shape[:-1], self.output_dim)

    @torch.autocast("cuda", enabled=False)                           # always in the table's own precision
    def grad_total_variation(self, weight=1e-7, inputs=None, bound=1, B=1000000):
        """Adds the total-variation gradient at `inputs` (default: B uniform random points) to `embeddings.grad`; call it
        between loss.backward() and optimizer.step() (grid.py:163-190)."""
        if self.embeddings.grad is None:
            raise ValueError('grad is None, should be called after loss.backward() and before optimizer.step()!')
        if inputs is None:
            where = torch.rand(B, self.input_dim, device=self.embeddings.device)
        else:
            where = self._unit_cube(inputs, bound).view(-1, self.input_dim)
        where = where.to(self.embeddings.dtype).contiguous()
        _backend.grad_total_variation(where, self.embeddings, self.embeddings.grad, self.offsets, weight, where.shape[0],
                                      self.input_dim, self.embeddings.shape[1], self.offsets.shape[0] - 1,
                                      float(np.log2(self.per_level_scale)), self.base_resolution, self.gridtype_id, self.align_corners)
