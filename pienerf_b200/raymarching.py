"""Host-side mirror of raymarching/raymarching.py (inference functions and the training pair march_rays_train /
composite_rays_train) plus the two per-frame helpers the
renderer takes from nerf/utils.py (get_rays, get_pnts_in_grids).  Same names, argument order, return values and
caller-allocates-zeroed-outputs convention as the reference wrappers."""
import torch
from torch.amp import custom_bwd, custom_fwd
from torch.autograd import Function

from . import _raymarching as _backend
from ._lib import check, dptr, lib, stream_ptr


def _f32c(t):
    return t.to(torch.float32).contiguous()


@torch.no_grad()
def near_far_from_aabb(rays_o, rays_d, aabb, min_near=0.2):
    """raymarching.py:21-51."""
    if not rays_o.is_cuda: rays_o = rays_o.cuda()
    if not rays_d.is_cuda: rays_d = rays_d.cuda()
    rays_o = _f32c(rays_o).view(-1, 3)
    rays_d = _f32c(rays_d).view(-1, 3)
    N = rays_o.shape[0]
    nears = torch.empty(N, dtype=rays_o.dtype, device=rays_o.device)
    fars = torch.empty(N, dtype=rays_o.dtype, device=rays_o.device)
    _backend.near_far_from_aabb(rays_o, rays_d, _f32c(aabb), N, min_near, nears, fars)
    return nears, fars


@torch.no_grad()
def sph_from_ray(rays_o, rays_d, radius):
    """raymarching.py:54-82."""
    rays_o = _f32c(rays_o.cuda()).view(-1, 3)
    rays_d = _f32c(rays_d.cuda()).view(-1, 3)
    N = rays_o.shape[0]
    coords = torch.empty(N, 2, dtype=rays_o.dtype, device=rays_o.device)
    _backend.sph_from_ray(rays_o, rays_d, radius, N, coords)
    return coords


@torch.no_grad()
def morton3D(coords):
    """raymarching.py:85-106."""
    if not coords.is_cuda: coords = coords.cuda()
    N = coords.shape[0]
    indices = torch.empty(N, dtype=torch.int32, device=coords.device)
    _backend.morton3D(coords.int().contiguous(), N, indices)
    return indices


@torch.no_grad()
def morton3D_invert(indices):
    """raymarching.py:108-128."""
    if not indices.is_cuda: indices = indices.cuda()
    N = indices.shape[0]
    coords = torch.empty(N, 3, dtype=torch.int32, device=indices.device)
    _backend.morton3D_invert(indices.int().contiguous(), N, coords)
    return coords


@torch.no_grad()
def packbits(grid, thresh, bitfield=None):
    """raymarching.py:131-157."""
    if not grid.is_cuda: grid = grid.cuda()
    grid = _f32c(grid)
    C, H3 = grid.shape[0], grid.shape[1]
    N = C * H3 // 8
    if bitfield is None:
        bitfield = torch.empty(N, dtype=torch.uint8, device=grid.device)
    _backend.packbits(grid, N, thresh, bitfield)
    return bitfield


def _alloc_samples(n_alive, n_step, align, like):
    M = n_alive * n_step
    if align > 0:
        M += align - (M % align)                               # raymarching.py:337-338: always pads 1..align rows
    xyzs = torch.zeros(M, 3, dtype=like.dtype, device=like.device)
    dirs = torch.zeros(M, 3, dtype=like.dtype, device=like.device)
    deltas = torch.zeros(M, 2, dtype=like.dtype, device=like.device)
    return xyzs, dirs, deltas


@torch.no_grad()
def march_rays(n_alive, n_step, rays_alive, rays_t, rays_o, rays_d, bound, density_bitfield, C, H, near, far, align=-1,
               perturb=False, dt_gamma=0, max_steps=1024):
    """raymarching.py:299-359."""
    if not rays_o.is_cuda: rays_o = rays_o.cuda()
    if not rays_d.is_cuda: rays_d = rays_d.cuda()
    rays_o = _f32c(rays_o).view(-1, 3)
    rays_d = _f32c(rays_d).view(-1, 3)
    xyzs, dirs, deltas = _alloc_samples(n_alive, n_step, align, rays_o)
    if perturb:
        noises = torch.rand(n_alive, dtype=rays_o.dtype, device=rays_o.device)
    else:
        noises = torch.zeros(n_alive, dtype=rays_o.dtype, device=rays_o.device)
    _backend.march_rays(n_alive, n_step, rays_alive, rays_t, rays_o, rays_d, bound, dt_gamma, max_steps, C, H,
                        density_bitfield, near, far, xyzs, dirs, deltas, noises)
    return xyzs, dirs, deltas


@torch.no_grad()
def composite_rays(n_alive, n_step, rays_alive, rays_t, sigmas, rgbs, deltas, weights_sum, depth, image, T_thresh=1e-2):
    """raymarching.py:362-384 (in-place on rays_alive, rays_t, weights_sum, depth, image)."""
    _backend.composite_rays(n_alive, n_step, T_thresh, rays_alive, rays_t, _f32c(sigmas), _f32c(rgbs), deltas, weights_sum,
                            depth, image)
    return tuple()


@torch.no_grad()
def march_rays_quadratic_bending(pig_cnt, pig_bgn, pig_idx, n_vtx, n_grid, p_def, p_ori, F_IP, dF_IP, max_iter_num, bbmin,
                                 bbmax, hgs, res, num_seek_IP, IP_dx, cut, cut_bounds, n_alive, n_step, rays_alive, rays_t,
                                 rays_o, rays_d, bound, density_bitfield, C, H, near, far, align=-1, perturb=False,
                                 dt_gamma=0, max_steps=1024):
    """raymarching.py:387-441."""
    if not rays_o.is_cuda: rays_o = rays_o.cuda()
    if not rays_d.is_cuda: rays_d = rays_d.cuda()
    rays_o = _f32c(rays_o).view(-1, 3)
    rays_d = _f32c(rays_d).view(-1, 3)
    xyzs, dirs, deltas = _alloc_samples(n_alive, n_step, align, rays_o)
    if perturb:
        noises = torch.rand(n_alive, dtype=rays_o.dtype, device=rays_o.device)
    else:
        noises = torch.zeros(n_alive, dtype=rays_o.dtype, device=rays_o.device)
    _backend.march_rays_quadratic_bending(
        pig_cnt, pig_bgn, pig_idx, int(n_vtx), int(n_grid), _f32c(p_def), _f32c(p_ori), _f32c(F_IP), _f32c(dF_IP),
        max_iter_num, _f32c(bbmin), _f32c(bbmax), hgs, res, num_seek_IP, IP_dx, cut, _f32c(cut_bounds),
        n_alive, n_step, rays_alive, rays_t, rays_o, rays_d, bound, dt_gamma, max_steps, C, H, density_bitfield, near, far,
        xyzs, dirs, deltas, noises)
    return xyzs, dirs, deltas


# ---- training side ---------------------------------------------------------------------

class _march_rays_train(Function):
    """raymarching.py:163-235: forward only; returns (xyzs [M,3], dirs [M,3], deltas [M,2], rays [N,3])."""

    @staticmethod
    @custom_fwd(device_type="cuda", cast_inputs=torch.float32)
    def forward(ctx, rays_o, rays_d, bound, density_bitfield, C, H, nears, fars, step_counter=None, mean_count=-1, perturb=False,
                align=-1, force_all_rays=False, dt_gamma=0, max_steps=1024):
        if not rays_o.is_cuda: rays_o = rays_o.cuda()
        if not rays_d.is_cuda: rays_d = rays_d.cuda()
        if not density_bitfield.is_cuda: density_bitfield = density_bitfield.cuda()
        rays_o = rays_o.contiguous().view(-1, 3)
        rays_d = rays_d.contiguous().view(-1, 3)
        density_bitfield = density_bitfield.contiguous()
        N = rays_o.shape[0]
        M = N * max_steps
        if not force_all_rays and mean_count > 0:            # running estimate of the sample count (raymarching.py:202-205)
            if align > 0:
                mean_count += align - mean_count % align
            M = mean_count
        xyzs = torch.zeros(M, 3, dtype=rays_o.dtype, device=rays_o.device)
        dirs = torch.zeros(M, 3, dtype=rays_o.dtype, device=rays_o.device)
        deltas = torch.zeros(M, 2, dtype=rays_o.dtype, device=rays_o.device)
        rays = torch.empty(N, 3, dtype=torch.int32, device=rays_o.device)
        if step_counter is None:
            step_counter = torch.zeros(2, dtype=torch.int32, device=rays_o.device)
        if perturb:
            noises = torch.rand(N, dtype=rays_o.dtype, device=rays_o.device)
        else:
            noises = torch.zeros(N, dtype=rays_o.dtype, device=rays_o.device)
        _backend.march_rays_train(rays_o, rays_d, density_bitfield, bound, dt_gamma, max_steps, N, C, H, M, nears.contiguous(),
                                  fars.contiguous(), xyzs, dirs, deltas, rays, step_counter, noises)
        if force_all_rays or mean_count <= 0:                # first epochs: trim to the samples actually produced
            m = step_counter[0].item()
            if align > 0:
                m += align - m % align
            xyzs, dirs, deltas = xyzs[:m], dirs[:m], deltas[:m]
        return xyzs, dirs, deltas, rays


march_rays_train = _march_rays_train.apply


class _composite_rays_train(Function):
    """raymarching.py:240-290: (sigmas [M], rgbs [M,3], deltas [M,2], rays [N,3]) -> weights_sum [N], depth [N], image [N,3]."""

    @staticmethod
    @custom_fwd(device_type="cuda", cast_inputs=torch.float32)
    def forward(ctx, sigmas, rgbs, deltas, rays, T_thresh=1e-4):
        sigmas = sigmas.contiguous()
        rgbs = rgbs.contiguous()
        deltas = deltas.contiguous()
        M = sigmas.shape[0]
        N = rays.shape[0]
        weights_sum = torch.empty(N, dtype=sigmas.dtype, device=sigmas.device)
        depth = torch.empty(N, dtype=sigmas.dtype, device=sigmas.device)
        image = torch.empty(N, 3, dtype=sigmas.dtype, device=sigmas.device)
        _backend.composite_rays_train_forward(sigmas, rgbs, deltas, rays, M, N, T_thresh, weights_sum, depth, image)
        ctx.save_for_backward(sigmas, rgbs, deltas, rays, weights_sum, depth, image)
        ctx.dims = [M, N, T_thresh]
        return weights_sum, depth, image

    @staticmethod
    @custom_bwd(device_type="cuda")
    def backward(ctx, grad_weights_sum, grad_depth, grad_image):
        # the depth gradient is not propagated, as in the reference (raymarching.py:277)
        grad_weights_sum = grad_weights_sum.contiguous()
        grad_image = grad_image.contiguous()
        sigmas, rgbs, deltas, rays, weights_sum, depth, image = ctx.saved_tensors
        M, N, T_thresh = ctx.dims
        grad_sigmas = torch.zeros_like(sigmas)
        grad_rgbs = torch.zeros_like(rgbs)
        _backend.composite_rays_train_backward(grad_weights_sum, grad_image, sigmas, rgbs, deltas, rays, weights_sum, image, M, N,
                                               T_thresh, grad_sigmas, grad_rgbs)
        return grad_sigmas, grad_rgbs, None, None, None


composite_rays_train = _composite_rays_train.apply


# ---- nerf/utils.py pieces of the hot path --------------------------------------------------------------

def select_ray_indices(H, W, N, B=1, error_map=None, patch_size=1, device="cuda"):
    """Which pixels a training batch looks at (nerf/utils.py:76-114): N uniform random pixels (duplicates allowed), or N // p^2 random
    p x p patches, or N pixels drawn from the 128 x 128 error map and jittered inside their coarse cell.  Same random draws in the
    same order as the reference, so a seeded generator on the same device selects the same pixels.  Returns (inds [B, N] int64,
    inds_coarse [B, N] or None)."""
    N = min(int(N), H * W)
    if patch_size > 1:                                   # patches: the error map is ignored
        num_patch = N // (patch_size ** 2)
        top = torch.randint(0, H - patch_size, size=[num_patch], device=device)
        left = torch.randint(0, W - patch_size, size=[num_patch], device=device)
        di, dj = torch.meshgrid(torch.arange(patch_size, device=device), torch.arange(patch_size, device=device), indexing="ij")
        rows = (top[:, None] + di.reshape(1, -1)).reshape(-1)
        cols = (left[:, None] + dj.reshape(1, -1)).reshape(-1)
        flat = rows * W + cols
        return flat.expand([B, flat.shape[0]]), None       # the reference expands to [B, N]: N must be a multiple of p^2 there
    if error_map is None:
        return torch.randint(0, H * W, size=[N], device=device).expand([B, N]), None
    coarse = torch.multinomial(error_map.to(device), N, replacement=False)
    cx, cy = coarse // 128, coarse % 128
    sx, sy = H / 128, W / 128
    rows = (cx * sx + torch.rand(B, N, device=device) * sx).long().clamp(max=H - 1)
    cols = (cy * sy + torch.rand(B, N, device=device) * sy).long().clamp(max=W - 1)
    return rows * W + cols, coarse


@torch.no_grad()
def get_rays(poses, intrinsics, H, W, N=-1, error_map=None, patch_size=1):
    """nerf/utils.py:55-138 for one camera (B = 1): the full frame the GUI path renders (N = -1) or a training batch of N pixels
    (uniform, patch-based or error-map driven; `inds` / `inds_coarse` returned as the reference returns them)."""
    poses = torch.as_tensor(poses, dtype=torch.float32)
    if poses.dim() == 3:
        if poses.shape[0] != 1:
            raise NotImplementedError("B must be 1")
        poses = poses[0]
    pose_host = poses.detach().cpu().contiguous()
    fx, fy, cx, cy = [float(torch.tensor(float(v), dtype=torch.float32)) for v in intrinsics]
    dev = torch.device("cuda")
    if N > 0:
        inds, inds_coarse = select_ray_indices(H, W, N, 1, error_map, patch_size, dev)
        n = inds.shape[1]
        cam = torch.cat([pose_host.reshape(-1)[:16], torch.tensor([fx, fy, cx, cy], dtype=torch.float32)]).to(dev)
        pix = inds[0].to(torch.int32).contiguous()
        rays_o = torch.empty(1, n, 3, dtype=torch.float32, device=dev)
        rays_d = torch.empty(1, n, 3, dtype=torch.float32, device=dev)
        check(lib.pn_get_rays_pix(dptr(cam), int(H), int(W), dptr(pix), int(n), dptr(rays_o), dptr(rays_d), stream_ptr()))
        out = {"rays_o": rays_o, "rays_d": rays_d, "inds": inds}
        if inds_coarse is not None:
            out["inds_coarse"] = inds_coarse
        return out
    rays_o = torch.empty(1, H * W, 3, dtype=torch.float32, device=dev)
    rays_d = torch.empty(1, H * W, 3, dtype=torch.float32, device=dev)
    import ctypes
    check(lib.pn_get_rays(ctypes.c_void_p(pose_host.data_ptr()), fx, fy, cx, cy, int(H), int(W), dptr(rays_o), dptr(rays_d),
                          stream_ptr()))
    return {"rays_o": rays_o, "rays_d": rays_d, "inds": None}


@torch.no_grad()
def ip_bbox(p_def, hgs, cut=False, bound=1.0):
    """nerf/renderer.py:782-791 as one kernel: returns bbmin[3], bbmax[3] (f32) and resolution[3] (i32) on device."""
    p_def = _f32c(p_def)
    bbmin = torch.empty(3, dtype=torch.float32, device=p_def.device)
    bbmax = torch.empty(3, dtype=torch.float32, device=p_def.device)
    res = torch.empty(3, dtype=torch.int32, device=p_def.device)
    check(lib.pn_ip_bbox(dptr(p_def), p_def.shape[0], float(hgs), int(bool(cut)), float(bound), dptr(bbmin), dptr(bbmax),
                         dptr(res), stream_ptr()))
    return bbmin, bbmax, res


@torch.no_grad()
def get_pnts_in_grids(n_vtx, n_grid, pnts, bbmin, bbmax, hgs, resolution):
    """nerf/utils.py:355-386: (pig_cnt, pig_bgn, pig_idx).  Within-cell order is ascending IP index."""
    n_grid = int(n_grid)
    pnts = _f32c(pnts)
    pig_idx = torch.zeros((n_vtx,), dtype=torch.int32, device=pnts.device)
    pig_cnt = torch.zeros((n_grid,), dtype=torch.int32, device=pnts.device)
    pig_bgn = torch.zeros((n_grid,), dtype=torch.int32, device=pnts.device)
    check(lib.pn_build_ip_grid(dptr(pnts), int(n_vtx), dptr(_f32c(bbmin)), float(hgs), dptr(resolution.to(torch.int32).contiguous()),
                               n_grid, dptr(pig_cnt), dptr(pig_bgn), dptr(pig_idx), stream_ptr()))
    return pig_cnt, pig_bgn, pig_idx
