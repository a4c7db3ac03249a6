"""Multi-GPU plumbing of the frame loop: one process per GPU, rays sharded by interleaved
image tiles, ONE broadcast of the packed IP state per step (the simulator stays on rank 0) and ONE gather of
the framebuffer tiles per frame.  Backend-agnostic (`nccl` on GPUs, `gloo` in the CPU tests)."""
import numpy as np
import torch
import torch.distributed as dist


def tile_partition(H, W, world_size, tile=16, weights=None):
    """Assignment of tile x tile pixel blocks to ranks, interleaved along diagonals so neighbouring tiles in x AND
    y land on different ranks (contiguous bands would not balance: the object covers a minority of the pixels).
    `weights` (len world_size, sum 1) gives each rank's share of the tiles — rank 0 also runs the simulator, so it
    gets a smaller share (see DistFrameDriver.calibrate).  Deterministic: every rank computes the same partition.
    Returns a list of int64 pixel-index arrays (row-major), one per rank, every pixel exactly once."""
    ty, tx = (H + tile - 1) // tile, (W + tile - 1) // tile
    if world_size == 1:
        return [np.arange(H * W, dtype=np.int64)]
    ys, xs = np.divmod(np.arange(ty * tx), tx)
    order = np.lexsort((ys, (ys + xs)))                      # walk tiles diagonal by diagonal
    if weights is None:
        owner_of_seq = np.arange(ty * tx) % world_size
    else:
        w = np.asarray(weights, dtype=np.float64)
        w = w / w.sum()
        owner_of_seq = np.empty(ty * tx, dtype=np.int64)
        if world_size > 2 and np.allclose(w[1:], w[1]):
            # rank 0 takes evenly spaced tiles of the diagonal walk, the others keep a plain round robin over the rest
            # (preserves the diagonal interleave, which balances much better than a generic weighted scheme)
            n0 = int(round(w[0] * ty * tx))
            take0 = np.zeros(ty * tx, dtype=bool)
            if n0 > 0:
                take0[np.unique(np.floor((np.arange(n0) + 0.5) * (ty * tx) / n0).astype(np.int64))] = True
            owner_of_seq[take0] = 0
            owner_of_seq[~take0] = 1 + np.arange(int((~take0).sum())) % (world_size - 1)
        else:
            credit = np.zeros(world_size)
            for i in range(ty * tx):                         # largest-remaining-credit weighted round robin
                credit += w
                r = int(np.argmax(credit))
                owner_of_seq[i] = r
                credit[r] -= 1.0
    tile_owner = np.empty(ty * tx, dtype=np.int64)
    tile_owner[order] = owner_of_seq
    pix_tile = (np.arange(H)[:, None] // tile) * tx + (np.arange(W)[None, :] // tile)
    flat = tile_owner[pix_tile].reshape(-1)
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


def ip_state_views(buf, n_ip):
    """One flat fp32 buffer of 39 n floats = [pos n*3 | F n*9 | dF n*27]: the simulator writes its IP state straight into
    these contiguous views and ONE broadcast of the flat buffer moves it — no pack / unpack copies."""
    pos = buf[:3 * n_ip].view(n_ip, 3)
    F = buf[3 * n_ip:12 * n_ip].view(n_ip, 9)
    dF = buf[12 * n_ip:39 * n_ip].view(n_ip, 27)
    return pos, F, dF


class PlanarFrameGather:
    """Copy-free variant of FrameGather.  Every rank renders straight into its send segment
    [image n_max*3 | depth n_max | depth_0 n_max] (views handed to the renderer as `out=`), ONE gather moves the segments,
    and rank 0 scatters them into the [H*W,5] frame with three index_copy_ calls over the concatenated pixel indices
    (padding rows of the shorter segments land in a dummy row)."""

    def __init__(self, parts, device):
        self.parts, self.world = parts, len(parts)
        self.rank = dist.get_rank() if dist.is_initialized() else 0
        self.n = len(parts[self.rank])
        self.n_max = max(len(p) for p in parts)
        self.n_pix = sum(len(p) for p in parts)
        nm = self.n_max
        self.send = torch.zeros(5 * nm, dtype=torch.float32, device=device)
        self.out = {"image": self.send[:3 * nm].view(nm, 3)[:self.n], "depth": self.send[3 * nm:4 * nm][:self.n],
                    "depth_0": self.send[4 * nm:5 * nm][:self.n], "weights_sum": torch.empty(self.n, dtype=torch.float32, device=device)}
        if self.rank == 0 and self.world > 1:
            self.recv = torch.zeros(self.world, 5 * nm, dtype=torch.float32, device=device)
            self.recv_list = list(self.recv.unbind(0))
            idx = np.full((self.world, nm), self.n_pix, dtype=np.int64)          # padding -> dummy row n_pix
            for r, p in enumerate(parts):
                idx[r, :len(p)] = p
            self.index = torch.from_numpy(idx.reshape(-1)).to(device)
            self.frame = {"image": torch.zeros(self.n_pix + 1, 3, dtype=torch.float32, device=device),
                          "depth": torch.zeros(self.n_pix + 1, dtype=torch.float32, device=device),
                          "depth_0": torch.zeros(self.n_pix + 1, dtype=torch.float32, device=device)}

    def __call__(self):
        """Returns {"image" [H*W,3], "depth" [H*W], "depth_0" [H*W]} on rank 0 (row-major pixels), None elsewhere."""
        if self.world == 1:
            return {k: self.out[k] for k in ("image", "depth", "depth_0")}      # the renderer's own outputs, no copy at all
        dist.gather(self.send, self.recv_list if self.rank == 0 else None, dst=0)
        if self.rank != 0:
            return None
        nm = self.n_max
        self.frame["image"].index_copy_(0, self.index, self.recv[:, :3 * nm].reshape(-1, 3))
        self.frame["depth"].index_copy_(0, self.index, self.recv[:, 3 * nm:4 * nm].reshape(-1))
        self.frame["depth_0"].index_copy_(0, self.index, self.recv[:, 4 * nm:5 * nm].reshape(-1))
        return {k: v[:self.n_pix] for k, v in self.frame.items()}


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
