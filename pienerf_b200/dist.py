"""Multi-GPU plumbing of the frame loop: one process per GPU, rays sharded by interleaved
image tiles, ONE broadcast of the packed IP state per step (the simulator stays on rank 0) and ONE gather of
the framebuffer tiles per frame.  Backend-agnostic (`nccl` on GPUs, `gloo` in the CPU tests)."""
import numpy as np
import torch
import torch.distributed as dist


def tile_partition(H, W, world_size, tile=16):
    """Round-robin assignment of tile x tile pixel blocks to ranks (contiguous bands would not balance: the
    object covers a minority of the pixels).  Returns a list of int64 pixel-index arrays (row-major), one per
    rank, every pixel exactly once."""
    ty, tx = (H + tile - 1) // tile, (W + tile - 1) // tile
    tid = (np.arange(H)[:, None] // tile) * tx + (np.arange(W)[None, :] // tile)
    # diagonal interleave so neighbouring tiles in x AND y land on different ranks
    owner = ((np.arange(H)[:, None] // tile) + (np.arange(W)[None, :] // tile)) % world_size if world_size > 1 else np.zeros((H, W), np.int64)
    del tid, ty
    flat = owner.reshape(-1)
    return [np.nonzero(flat == r)[0].astype(np.int64) for r in range(world_size)]


def pack_ip_state(pos, F, dF, out=None):
    """[n,3] | [n,9] | [n,27] fp32 -> one [n,39] buffer (156 B per IP) for a single broadcast."""
    if out is None:
        out = torch.empty(pos.shape[0], 39, dtype=torch.float32, device=pos.device)
    out[:, 0:3] = pos; out[:, 3:12] = F; out[:, 12:39] = dF
    return out


def unpack_ip_state(buf):
    return buf[:, 0:3].contiguous(), buf[:, 3:12].contiguous(), buf[:, 12:39].contiguous()


def broadcast_ip_state(buf, src=0):
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.broadcast(buf, src=src)
    return buf


class FrameGather:
    """Gathers per-rank [n_r, C] pixel rows into the full [H*W, C] framebuffer on rank 0."""

    def __init__(self, parts, channels, device, dtype=torch.float32):
        self.parts = parts
        self.world = len(parts)
        self.rank = dist.get_rank() if dist.is_initialized() else 0
        self.n_max = max(len(p) for p in parts)
        self.send = torch.zeros(self.n_max, channels, dtype=dtype, device=device)
        self.n_pix = sum(len(p) for p in parts)
        if self.rank == 0:
            self.recv = [torch.zeros(self.n_max, channels, dtype=dtype, device=device) for _ in range(self.world)]
            self.index = [torch.from_numpy(p).to(device) for p in parts]
            self.frame = torch.zeros(self.n_pix, channels, dtype=dtype, device=device)
        else:
            self.recv = None

    def __call__(self, local_rows):
        n = local_rows.shape[0]
        self.send[:n] = local_rows
        if self.world > 1:
            dist.gather(self.send, self.recv if self.rank == 0 else None, dst=0)
        elif self.rank == 0:
            self.recv[0].copy_(self.send)
        if self.rank != 0:
            return None
        for r in range(self.world):
            self.frame.index_copy_(0, self.index[r], self.recv[r][:len(self.parts[r])])
        return self.frame
