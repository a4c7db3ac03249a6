"""`_raymarching` — same entry points as the reference pybind module (raymarching/src/bindings.cpp:5-19),
positional signatures of raymarching/src/raymarching.h:7-36.  Like the reference, these do not validate
devices/contiguity beyond what is needed to take a pointer; the Python wrappers in raymarching.py do."""
import torch

from ._lib import check, dptr, lib, stream_ptr

f32 = torch.float32


def near_far_from_aabb(rays_o, rays_d, aabb, N, min_near, nears, fars):
    check(lib.pn_near_far_from_aabb(dptr(rays_o, "rays_o", f32), dptr(rays_d, "rays_d", f32), dptr(aabb, "aabb", f32), int(N),
                                    float(min_near), dptr(nears, "nears", f32), dptr(fars, "fars", f32), stream_ptr()))


def sph_from_ray(rays_o, rays_d, radius, N, coords):
    check(lib.pn_sph_from_ray(dptr(rays_o, "rays_o", f32), dptr(rays_d, "rays_d", f32), float(radius), int(N),
                              dptr(coords, "coords", f32), stream_ptr()))


def morton3D(coords, N, indices):
    check(lib.pn_morton3D(dptr(coords, "coords", torch.int32), int(N), dptr(indices, "indices", torch.int32), stream_ptr()))


def morton3D_invert(indices, N, coords):
    check(lib.pn_morton3D_invert(dptr(indices, "indices", torch.int32), int(N), dptr(coords, "coords", torch.int32), stream_ptr()))


def packbits(grid, N, density_thresh, bitfield):
    check(lib.pn_packbits(dptr(grid, "grid", f32), int(N), float(density_thresh), dptr(bitfield, "bitfield", torch.uint8), stream_ptr()))


def march_rays(n_alive, n_step, rays_alive, rays_t, rays_o, rays_d, bound, dt_gamma, max_steps, C, H, grid, near, far,
               xyzs, dirs, deltas, noises):
    check(lib.pn_march_rays(int(n_alive), int(n_step), dptr(rays_alive, "rays_alive", torch.int32), dptr(rays_t, "rays_t", f32),
                            dptr(rays_o, "rays_o", f32), dptr(rays_d, "rays_d", f32), float(bound), float(dt_gamma),
                            int(max_steps), int(C), int(H), dptr(grid, "grid", torch.uint8), dptr(near, "near", f32),
                            dptr(far, "far", f32), dptr(xyzs, "xyzs", f32), dptr(dirs, "dirs", f32), dptr(deltas, "deltas", f32),
                            dptr(noises, "noises", f32), stream_ptr()))


def composite_rays(n_alive, n_step, T_thresh, rays_alive, rays_t, sigmas, rgbs, deltas, weights, depth, image):
    check(lib.pn_composite_rays(int(n_alive), int(n_step), float(T_thresh), dptr(rays_alive, "rays_alive", torch.int32),
                                dptr(rays_t, "rays_t", f32), dptr(sigmas, "sigmas", f32), dptr(rgbs, "rgbs", f32),
                                dptr(deltas, "deltas", f32), dptr(weights, "weights", f32), dptr(depth, "depth", f32),
                                dptr(image, "image", f32), stream_ptr()))


def march_rays_quadratic_bending(pig_cnt, pig_bgn, pig_idx, n_vtx, n_grid, p_def, p_ori, F_IP, dF_IP, max_iter_num, bbmin,
                                 bbmax, hgs, resolution, num_seek_IP, IP_dx, cut, cut_bounds, n_alive, n_step, rays_alive,
                                 rays_t, rays_o, rays_d, bound, dt_gamma, max_steps, C, H, grid, near, far, xyzs, dirs,
                                 deltas, noises):
    i32 = torch.int32
    check(lib.pn_march_rays_quadratic_bending(
        dptr(pig_cnt, "pig_cnt", i32), dptr(pig_bgn, "pig_bgn", i32), dptr(pig_idx, "pig_idx", i32), int(n_vtx), int(n_grid),
        dptr(p_def, "p_def", f32), dptr(p_ori, "p_ori", f32), dptr(F_IP, "F_IP", f32), dptr(dF_IP, "dF_IP", f32),
        int(max_iter_num), dptr(bbmin, "bbmin", f32), dptr(bbmax, "bbmax", f32), float(hgs), dptr(resolution, "resolution", i32),
        int(num_seek_IP), float(IP_dx), int(bool(cut)), dptr(cut_bounds, "cut_bounds", f32), int(n_alive), int(n_step),
        dptr(rays_alive, "rays_alive", i32), dptr(rays_t, "rays_t", f32), dptr(rays_o, "rays_o", f32), dptr(rays_d, "rays_d", f32),
        float(bound), float(dt_gamma), int(max_steps), int(C), int(H), dptr(grid, "grid", torch.uint8), dptr(near, "near", f32),
        dptr(far, "far", f32), dptr(xyzs, "xyzs", f32), dptr(dirs, "dirs", f32), dptr(deltas, "deltas", f32),
        dptr(noises, "noises", f32), stream_ptr()))


def march_rays_train(rays_o, rays_d, grid, bound, dt_gamma, max_steps, N, C, H, M, nears, fars, xyzs, dirs, deltas, rays,
                     counter, noises):
    """raymarching.cu:485-493; fp32 only (the reference's wrapper casts to float32, raymarching.py:165)."""
    i32 = torch.int32
    check(lib.pn_march_rays_train(dptr(rays_o, "rays_o", f32), dptr(rays_d, "rays_d", f32), dptr(grid, "grid", torch.uint8),
                                  float(bound), float(dt_gamma), int(max_steps), int(N), int(C), int(H), int(M),
                                  dptr(nears, "nears", f32), dptr(fars, "fars", f32), dptr(xyzs, "xyzs", f32),
                                  dptr(dirs, "dirs", f32), dptr(deltas, "deltas", f32), dptr(rays, "rays", i32),
                                  dptr(counter, "counter", i32), dptr(noises, "noises", f32), stream_ptr()))


def composite_rays_train_forward(sigmas, rgbs, deltas, rays, M, N, T_thresh, weights_sum, depth, image):
    check(lib.pn_composite_rays_train_forward(dptr(sigmas, "sigmas", f32), dptr(rgbs, "rgbs", f32), dptr(deltas, "deltas", f32),
                                              dptr(rays, "rays", torch.int32), int(M), int(N), float(T_thresh),
                                              dptr(weights_sum, "weights_sum", f32), dptr(depth, "depth", f32),
                                              dptr(image, "image", f32), stream_ptr()))


def composite_rays_train_backward(grad_weights_sum, grad_image, sigmas, rgbs, deltas, rays, weights_sum, image, M, N, T_thresh,
                                  grad_sigmas, grad_rgbs):
    check(lib.pn_composite_rays_train_backward(
        dptr(grad_weights_sum, "grad_weights_sum", f32), dptr(grad_image, "grad_image", f32), dptr(sigmas, "sigmas", f32),
        dptr(rgbs, "rgbs", f32), dptr(deltas, "deltas", f32), dptr(rays, "rays", torch.int32),
        dptr(weights_sum, "weights_sum", f32), dptr(image, "image", f32), int(M), int(N), float(T_thresh),
        dptr(grad_sigmas, "grad_sigmas", f32), dptr(grad_rgbs, "grad_rgbs", f32), stream_ptr()))
