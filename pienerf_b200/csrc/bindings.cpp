// pybind11 torch-extension modules with the reference's names and positional signatures, over the C-ABI:
//   _gridencoder  gridencoder/src/bindings.cpp:5-9   (gridencoder.h:12-15)
//   _shencoder    shencoder/src/bindings.cpp:5-8     (shencoder.h:7-10)
//   _raymarching  raymarching/src/bindings.cpp:5-19  (raymarching.h:7-36)
//   _qgmls        no reference counterpart (the reference simulator is Python + Warp kernels): one op per Warp kernel of
//                 simulator/cpu_utils.py + cuda_utils.py plus the fused step, the boundary SURVEY.md 8(b) prescribes
// One translation unit, compiled four times with -DPN_EXT_<NAME> (pienerf_b200/build_ext.py) and linked against
// libpienerf_b200.so; no kernels live here.  Tensors are checked the way the reference checks them (CUDA + contiguous + dtype,
// gridencoder.cu:15-18), every call enqueues on torch's current stream, a non-zero C-ABI return code becomes the exception the
// reference's bindings raise (RuntimeError; NotImplementedError for dtypes the library does not carry, e.g. fp64 tables).
#include <torch/extension.h>
#include <ATen/cuda/CUDAContext.h>
#include <c10/util/Optional.h>
#include "../../include/pienerf_b200.h"

namespace {

// (argument evaluation order is unspecified: this may run before the tensor checks, so it must not throw on a GPU-less host)
void *cur_stream() { return at::cuda::is_available() ? (void *)at::cuda::getCurrentCUDAStream().stream() : nullptr; }

void pn_check(int rc) {
    if (rc == PN_OK) return;
    const char *msg = pn_last_error();
    TORCH_CHECK_NOT_IMPLEMENTED(rc != PN_ENOTIMPL, msg && msg[0] ? msg : "not provided by the B200 library");   // -> NotImplementedError
    TORCH_CHECK(false, msg && msg[0] ? msg : "pienerf_b200 error");
}

#define PN_CHECK_CUDA(x) TORCH_CHECK((x).device().is_cuda(), #x " must be a CUDA tensor")
#define PN_CHECK_CONTIGUOUS(x) TORCH_CHECK((x).is_contiguous(), #x " must be a contiguous tensor")
#define PN_CHECK_F32(x) TORCH_CHECK((x).scalar_type() == at::ScalarType::Float, #x " must be a float32 tensor")
#define PN_CHECK_F64(x) TORCH_CHECK((x).scalar_type() == at::ScalarType::Double, #x " must be a float64 tensor")
#define PN_CHECK_I32(x) TORCH_CHECK((x).scalar_type() == at::ScalarType::Int, #x " must be an int tensor")
inline void chk_cuda_contig(const at::Tensor &x, const char *name) {
    TORCH_CHECK(x.device().is_cuda(), name, " must be a CUDA tensor");
    TORCH_CHECK(x.is_contiguous(), name, " must be a contiguous tensor");
}
inline float *f32_ptr(const at::Tensor &x, const char *name) {
    chk_cuda_contig(x, name);
    TORCH_CHECK(x.scalar_type() == at::ScalarType::Float, name, " must be a float32 tensor");
    return x.data_ptr<float>();
}
inline double *f64_ptr(const at::Tensor &x, const char *name) {
    chk_cuda_contig(x, name);
    TORCH_CHECK(x.scalar_type() == at::ScalarType::Double, name, " must be a float64 tensor");
    return x.data_ptr<double>();
}
inline int *i32_ptr(const at::Tensor &x, const char *name) {
    chk_cuda_contig(x, name);
    TORCH_CHECK(x.scalar_type() == at::ScalarType::Int, name, " must be an int tensor");
    return x.data_ptr<int>();
}
#define PN_F32(x) f32_ptr(x, #x)
#define PN_F64(x) f64_ptr(x, #x)
#define PN_I32(x) i32_ptr(x, #x)

template <typename T>
T *opt_ptr(const c10::optional<at::Tensor> &t) {
    if (!t.has_value() || !t->defined()) return nullptr;
    TORCH_CHECK(t->device().is_cuda() && t->is_contiguous(), "optional tensor must be a contiguous CUDA tensor");
    return (T *)t->data_ptr();
}

}  // namespace

#if defined(PN_EXT_GRIDENCODER)
// gridencoder.cu:448-471
void grid_encode_forward(const at::Tensor inputs, const at::Tensor embeddings, const at::Tensor offsets, at::Tensor outputs, const uint32_t B,
                         const uint32_t D, const uint32_t C, const uint32_t L, const float S, const uint32_t H, const c10::optional<at::Tensor> dy_dx,
                         const uint32_t gridtype, const bool align_corners, const uint32_t interp) {
    PN_CHECK_CUDA(inputs); PN_CHECK_CUDA(embeddings); PN_CHECK_CUDA(offsets); PN_CHECK_CUDA(outputs);
    PN_CHECK_CONTIGUOUS(inputs); PN_CHECK_CONTIGUOUS(embeddings); PN_CHECK_CONTIGUOUS(offsets); PN_CHECK_CONTIGUOUS(outputs);
    PN_CHECK_F32(inputs); PN_CHECK_I32(offsets);
    const bool half = embeddings.scalar_type() == at::ScalarType::Half;
    TORCH_CHECK(half || embeddings.scalar_type() == at::ScalarType::Float, "embeddings must be a float32 or float16 tensor");
    TORCH_CHECK(outputs.scalar_type() == embeddings.scalar_type(), "outputs must have the dtype of embeddings");
    TORCH_CHECK(C == 1 || C == 2 || C == 4 || C == 8, "GridEncoding: C must be 1, 2, 4, or 8.");
    pn_check(pn_grid_encode_forward(inputs.data_ptr<float>(), embeddings.data_ptr(), offsets.data_ptr<int>(), outputs.data_ptr(), B, D, C, L, S, H,
                                    opt_ptr<void>(dy_dx), gridtype, align_corners ? 1 : 0, interp, half ? 1 : 0, cur_stream()));
}
inline bool table_is_half(const at::Tensor &embeddings) {
    const bool half = embeddings.scalar_type() == at::ScalarType::Half;
    TORCH_CHECK_NOT_IMPLEMENTED(half || embeddings.scalar_type() == at::ScalarType::Float, "embeddings must be a float32 or float16 tensor (no fp64 tables)");
    return half;
}
inline void *same_dtype_ptr(const at::Tensor &x, const at::Tensor &like, const char *name) {
    chk_cuda_contig(x, name);
    TORCH_CHECK(x.scalar_type() == like.scalar_type(), name, " must have the dtype of embeddings");
    return x.data_ptr();
}
// gridencoder.cu:473-503
void grid_encode_backward(const at::Tensor grad, const at::Tensor inputs, const at::Tensor embeddings, const at::Tensor offsets, at::Tensor grad_embeddings,
                          const uint32_t B, const uint32_t D, const uint32_t C, const uint32_t L, const float S, const uint32_t H,
                          const c10::optional<at::Tensor> dy_dx, c10::optional<at::Tensor> grad_inputs, const uint32_t gridtype, const bool align_corners,
                          const uint32_t interp) {
    chk_cuda_contig(embeddings, "embeddings");
    const bool half = table_is_half(embeddings);
    const bool has_j = dy_dx.has_value() && dy_dx->defined(), has_gi = grad_inputs.has_value() && grad_inputs->defined();
    TORCH_CHECK(has_j == has_gi, "dy_dx and grad_inputs go together");
    TORCH_CHECK(C == 1 || C == 2 || C == 4 || C == 8, "GridEncoding: C must be 1, 2, 4, or 8.");
    pn_check(pn_grid_encode_backward(same_dtype_ptr(grad, embeddings, "grad"), PN_F32(inputs), embeddings.data_ptr(), PN_I32(offsets),
                                     same_dtype_ptr(grad_embeddings, embeddings, "grad_embeddings"), B, D, C, L, S, H,
                                     has_j ? same_dtype_ptr(*dy_dx, embeddings, "dy_dx") : nullptr,
                                     has_gi ? same_dtype_ptr(*grad_inputs, embeddings, "grad_inputs") : nullptr, gridtype, align_corners ? 1 : 0, interp,
                                     half ? 1 : 0, cur_stream()));
}
// gridencoder.cu:639-645
void grad_total_variation(const at::Tensor inputs, const at::Tensor embeddings, at::Tensor grad, const at::Tensor offsets, const float weight,
                          const uint32_t B, const uint32_t D, const uint32_t C, const uint32_t L, const float S, const uint32_t H, const uint32_t gridtype,
                          const bool align_corners) {
    chk_cuda_contig(embeddings, "embeddings");
    const bool half = table_is_half(embeddings);
    TORCH_CHECK(C == 1 || C == 2 || C == 4 || C == 8, "GridEncoding: C must be 1, 2, 4, or 8.");
    pn_check(pn_grad_total_variation(same_dtype_ptr(inputs, embeddings, "inputs"), embeddings.data_ptr(), same_dtype_ptr(grad, embeddings, "grad"),
                                     PN_I32(offsets), weight, B, D, C, L, S, H, gridtype, align_corners ? 1 : 0, half ? 1 : 0, cur_stream()));
}

PYBIND11_MODULE(TORCH_EXTENSION_NAME, m) {
    m.def("grid_encode_forward", &grid_encode_forward, "grid_encode_forward (CUDA)");
    m.def("grid_encode_backward", &grid_encode_backward, "grid_encode_backward (CUDA)");
    m.def("grad_total_variation", &grad_total_variation, "grad_total_variation (CUDA)");
}
#endif

#if defined(PN_EXT_SHENCODER)
// shencoder.cu:400-417
void sh_encode_forward(at::Tensor inputs, at::Tensor outputs, const uint32_t B, const uint32_t D, const uint32_t C, c10::optional<at::Tensor> dy_dx) {
    pn_check(pn_sh_encode_forward(PN_F32(inputs), PN_F32(outputs), B, D, C, opt_ptr<float>(dy_dx), cur_stream()));
}
// shencoder.cu:419-438
void sh_encode_backward(at::Tensor grad, at::Tensor inputs, const uint32_t B, const uint32_t D, const uint32_t C, at::Tensor dy_dx, at::Tensor grad_inputs) {
    pn_check(pn_sh_encode_backward(PN_F32(grad), PN_F32(inputs), B, D, C, PN_F32(dy_dx), PN_F32(grad_inputs), cur_stream()));
}

PYBIND11_MODULE(TORCH_EXTENSION_NAME, m) {
    m.def("sh_encode_forward", &sh_encode_forward, "SH encode forward (CUDA)");
    m.def("sh_encode_backward", &sh_encode_backward, "SH encode backward (CUDA)");
}
#endif

#if defined(PN_EXT_RAYMARCHING)
void near_far_from_aabb(const at::Tensor rays_o, const at::Tensor rays_d, const at::Tensor aabb, const uint32_t N, const float min_near, at::Tensor nears,
                        at::Tensor fars) {
    pn_check(pn_near_far_from_aabb(PN_F32(rays_o), PN_F32(rays_d), PN_F32(aabb), N, min_near, PN_F32(nears), PN_F32(fars), cur_stream()));
}
void sph_from_ray(const at::Tensor rays_o, const at::Tensor rays_d, const float radius, const uint32_t N, at::Tensor coords) {
    pn_check(pn_sph_from_ray(PN_F32(rays_o), PN_F32(rays_d), radius, N, PN_F32(coords), cur_stream()));
}
void morton3D(const at::Tensor coords, const uint32_t N, at::Tensor indices) { pn_check(pn_morton3D(PN_I32(coords), N, PN_I32(indices), cur_stream())); }
void morton3D_invert(const at::Tensor indices, const uint32_t N, at::Tensor coords) {
    pn_check(pn_morton3D_invert(PN_I32(indices), N, PN_I32(coords), cur_stream()));
}
void packbits(const at::Tensor grid, const uint32_t N, const float density_thresh, at::Tensor bitfield) {
    PN_CHECK_CUDA(bitfield); PN_CHECK_CONTIGUOUS(bitfield);
    TORCH_CHECK(bitfield.scalar_type() == at::ScalarType::Byte, "bitfield must be a uint8 tensor");
    pn_check(pn_packbits(PN_F32(grid), N, density_thresh, bitfield.data_ptr<uint8_t>(), cur_stream()));
}
void march_rays(const uint32_t n_alive, const uint32_t n_step, const at::Tensor rays_alive, const at::Tensor rays_t, const at::Tensor rays_o,
                const at::Tensor rays_d, const float bound, const float dt_gamma, const uint32_t max_steps, const uint32_t C, const uint32_t H,
                const at::Tensor grid, const at::Tensor near, const at::Tensor far, at::Tensor xyzs, at::Tensor dirs, at::Tensor deltas, at::Tensor noises) {
    PN_CHECK_CUDA(grid); PN_CHECK_CONTIGUOUS(grid);
    pn_check(pn_march_rays(n_alive, n_step, PN_I32(rays_alive), PN_F32(rays_t), PN_F32(rays_o), PN_F32(rays_d), bound, dt_gamma, max_steps, C, H,
                           (const uint8_t *)grid.data_ptr(), PN_F32(near), PN_F32(far), PN_F32(xyzs), PN_F32(dirs), PN_F32(deltas), PN_F32(noises),
                           cur_stream()));
}
void composite_rays(const uint32_t n_alive, const uint32_t n_step, const float T_thresh, at::Tensor rays_alive, at::Tensor rays_t, at::Tensor sigmas,
                    at::Tensor rgbs, at::Tensor deltas, at::Tensor weights, at::Tensor depth, at::Tensor image) {
    pn_check(pn_composite_rays(n_alive, n_step, T_thresh, PN_I32(rays_alive), PN_F32(rays_t), PN_F32(sigmas), PN_F32(rgbs), PN_F32(deltas),
                               PN_F32(weights), PN_F32(depth), PN_F32(image), cur_stream()));
}
// raymarching.h:20-36, same 36 positional arguments
void march_rays_quadratic_bending(const at::Tensor pig_cnt, const at::Tensor pig_bgn, const at::Tensor pig_idx, const int n_vtx, const int n_grid,
                                  const at::Tensor p_def, const at::Tensor p_ori, const at::Tensor F_IP, const at::Tensor dF_IP, const int max_iter_num,
                                  const at::Tensor bbmin, const at::Tensor bbmax, const float hgs, const at::Tensor resolution, const int num_seek_IP,
                                  const float IP_dx, const bool cut, const at::Tensor cut_bounds, const uint32_t n_alive, const uint32_t n_step,
                                  const at::Tensor rays_alive, const at::Tensor rays_t, const at::Tensor rays_o, const at::Tensor rays_d, const float bound,
                                  const float dt_gamma, const uint32_t max_steps, const uint32_t C, const uint32_t H, const at::Tensor grid,
                                  const at::Tensor near, const at::Tensor far, at::Tensor xyzs, at::Tensor dirs, at::Tensor deltas, at::Tensor noises) {
    PN_CHECK_CUDA(grid); PN_CHECK_CONTIGUOUS(grid);
    pn_check(pn_march_rays_quadratic_bending(PN_I32(pig_cnt), PN_I32(pig_bgn), PN_I32(pig_idx), n_vtx, n_grid, PN_F32(p_def), PN_F32(p_ori), PN_F32(F_IP),
                                             PN_F32(dF_IP), max_iter_num, PN_F32(bbmin), PN_F32(bbmax), hgs, PN_I32(resolution), num_seek_IP, IP_dx,
                                             cut ? 1 : 0, PN_F32(cut_bounds), n_alive, n_step, PN_I32(rays_alive), PN_F32(rays_t), PN_F32(rays_o),
                                             PN_F32(rays_d), bound, dt_gamma, max_steps, C, H, (const uint8_t *)grid.data_ptr(), PN_F32(near), PN_F32(far),
                                             PN_F32(xyzs), PN_F32(dirs), PN_F32(deltas), PN_F32(noises), cur_stream()));
}
inline uint8_t *u8_ptr(const at::Tensor &x, const char *name) {
    chk_cuda_contig(x, name);
    TORCH_CHECK(x.scalar_type() == at::ScalarType::Byte, name, " must be a uint8 tensor");
    return x.data_ptr<uint8_t>();
}
// raymarching.cu:485-493
void march_rays_train(const at::Tensor rays_o, const at::Tensor rays_d, const at::Tensor grid, const float bound, const float dt_gamma,
                      const uint32_t max_steps, const uint32_t N, const uint32_t C, const uint32_t H, const uint32_t M, const at::Tensor nears,
                      const at::Tensor fars, at::Tensor xyzs, at::Tensor dirs, at::Tensor deltas, at::Tensor rays, at::Tensor counter, at::Tensor noises) {
    pn_check(pn_march_rays_train(PN_F32(rays_o), PN_F32(rays_d), u8_ptr(grid, "grid"), bound, dt_gamma, max_steps, N, C, H, M, PN_F32(nears), PN_F32(fars),
                                 PN_F32(xyzs), PN_F32(dirs), PN_F32(deltas), PN_I32(rays), PN_I32(counter), PN_F32(noises), cur_stream()));
}
// raymarching.cu:583-591
void composite_rays_train_forward(const at::Tensor sigmas, const at::Tensor rgbs, const at::Tensor deltas, const at::Tensor rays, const uint32_t M,
                                  const uint32_t N, const float T_thresh, at::Tensor weights_sum, at::Tensor depth, at::Tensor image) {
    pn_check(pn_composite_rays_train_forward(PN_F32(sigmas), PN_F32(rgbs), PN_F32(deltas), PN_I32(rays), M, N, T_thresh, PN_F32(weights_sum),
                                             PN_F32(depth), PN_F32(image), cur_stream()));
}
// raymarching.cu:688-696
void composite_rays_train_backward(const at::Tensor grad_weights_sum, const at::Tensor grad_image, const at::Tensor sigmas, const at::Tensor rgbs,
                                   const at::Tensor deltas, const at::Tensor rays, const at::Tensor weights_sum, const at::Tensor image, const uint32_t M,
                                   const uint32_t N, const float T_thresh, at::Tensor grad_sigmas, at::Tensor grad_rgbs) {
    pn_check(pn_composite_rays_train_backward(PN_F32(grad_weights_sum), PN_F32(grad_image), PN_F32(sigmas), PN_F32(rgbs), PN_F32(deltas), PN_I32(rays),
                                              PN_F32(weights_sum), PN_F32(image), M, N, T_thresh, PN_F32(grad_sigmas), PN_F32(grad_rgbs), cur_stream()));
}
// extras over the reference module: the per-frame preparation the reference does with torch / Warp glue (nerf/utils.py:55-138,355-443)
void get_rays(const at::Tensor pose_host, const float fx, const float fy, const float cx, const float cy, const uint32_t H, const uint32_t W,
              at::Tensor rays_o, at::Tensor rays_d) {
    TORCH_CHECK(!pose_host.device().is_cuda() && pose_host.is_contiguous() && pose_host.scalar_type() == at::ScalarType::Float && pose_host.numel() >= 12,
                "pose_host must be a contiguous float32 CPU tensor [3..4, 4]");
    pn_check(pn_get_rays(pose_host.data_ptr<float>(), fx, fy, cx, cy, H, W, PN_F32(rays_o), PN_F32(rays_d), cur_stream()));
}
void ip_bbox(const at::Tensor p_def, const float hgs, const bool cut, const float bound, at::Tensor bbmin, at::Tensor bbmax, at::Tensor resolution) {
    pn_check(pn_ip_bbox(PN_F32(p_def), (int)p_def.size(0), hgs, cut ? 1 : 0, bound, PN_F32(bbmin), PN_F32(bbmax), PN_I32(resolution), cur_stream()));
}
void build_ip_grid(const at::Tensor p_def, const at::Tensor bbmin, const float hgs, const at::Tensor resolution, const int n_grid, at::Tensor pig_cnt,
                   at::Tensor pig_bgn, at::Tensor pig_idx) {
    pn_check(pn_build_ip_grid(PN_F32(p_def), (int)p_def.size(0), PN_F32(bbmin), hgs, PN_I32(resolution), n_grid, PN_I32(pig_cnt), PN_I32(pig_bgn),
                              PN_I32(pig_idx), cur_stream()));
}

PYBIND11_MODULE(TORCH_EXTENSION_NAME, m) {
    m.def("near_far_from_aabb", &near_far_from_aabb, "near_far_from_aabb (CUDA)");
    m.def("sph_from_ray", &sph_from_ray, "sph_from_ray (CUDA)");
    m.def("morton3D", &morton3D, "morton3D (CUDA)");
    m.def("morton3D_invert", &morton3D_invert, "morton3D_invert (CUDA)");
    m.def("packbits", &packbits, "packbits (CUDA)");
    m.def("march_rays_train", &march_rays_train, "march_rays_train (CUDA)");
    m.def("composite_rays_train_forward", &composite_rays_train_forward, "composite_rays_train_forward (CUDA)");
    m.def("composite_rays_train_backward", &composite_rays_train_backward, "composite_rays_train_backward (CUDA)");
    m.def("march_rays", &march_rays, "march rays (CUDA)");
    m.def("march_rays_quadratic_bending", &march_rays_quadratic_bending, "march rays through the quadratic GMLS warp (CUDA)");
    m.def("composite_rays", &composite_rays, "composite rays (CUDA)");
    m.def("get_rays", &get_rays, "pinhole rays of a whole frame (CUDA)");
    m.def("ip_bbox", &ip_bbox, "bounding box + grid resolution of the deformed IPs (CUDA)");
    m.def("build_ip_grid", &build_ip_grid, "counting sort of the IPs into the hash grid (CUDA)");
}
#endif

#if defined(PN_EXT_QGMLS)
// simulator/cpu_utils.py:3-152 (calc_G / calc_Gp / calc_weight)
void shape_functions(const double r, const at::Tensor pos, const at::Tensor topo, const at::Tensor kernel_pos, at::Tensor Nx, c10::optional<at::Tensor> dNx,
                     c10::optional<at::Tensor> ddNx, at::Tensor status) {
    pn_check(pn_qgmls_shape_functions(r, PN_F64(pos), PN_I32(topo), PN_F64(kernel_pos), (int)pos.size(0), PN_F64(Nx), opt_ptr<double>(dNx),
                                      opt_ptr<double>(ddNx), PN_I32(status), cur_stream()));
}
// simulator/cuda_utils.py:3-19
void collect_param(const at::Tensor pts_ip, const at::Tensor mu, const at::Tensor lam, const at::Tensor mass, const int n_ip, const double dx, at::Tensor ip_mu,
                   at::Tensor ip_lam, at::Tensor ip_rho) {
    pn_check(pn_qgmls_collect_param(PN_I32(pts_ip), PN_F64(mu), PN_F64(lam), PN_F64(mass), (int)pts_ip.size(0), n_ip, dx, PN_F64(ip_mu), PN_F64(ip_lam),
                                    PN_F64(ip_rho), cur_stream()));
}
// simulator/cuda_utils.py:22-55 (mu / lam may be None: the mass matrix, solver.py:517-531)
void build_ip_global(const double dx, const double dt, const at::Tensor topo, c10::optional<at::Tensor> mu, c10::optional<at::Tensor> lam, const at::Tensor rho,
                     const at::Tensor Nx, const at::Tensor dNx, const at::Tensor ddNx, at::Tensor mat) {
    pn_check(pn_qgmls_build_ip_global(dx, dt, PN_I32(topo), opt_ptr<double>(mu), opt_ptr<double>(lam), PN_F64(rho), PN_F64(Nx), PN_F64(dNx), PN_F64(ddNx),
                                      (int)topo.size(0), (int)mat.size(0), PN_F64(mat), cur_stream()));
}
// simulator/cuda_utils.py:58-81
void build_pin_global(const double stiff, const at::Tensor vidx, const at::Tensor topo, const at::Tensor Nx, at::Tensor mat) {
    pn_check(pn_qgmls_build_pin_global(stiff, PN_I32(vidx), (int)vidx.size(0), PN_I32(topo), PN_F64(Nx), (int)mat.size(0), PN_F64(mat), cur_stream()));
}
// simulator/cuda_utils.py:262-279
void collect_gravity(const double dx, const at::Tensor topo, const at::Tensor Nx, const std::vector<double> gravity, const at::Tensor rho, at::Tensor rhs) {
    TORCH_CHECK(gravity.size() == 3, "gravity must have 3 components");
    pn_check(pn_qgmls_collect_gravity(dx, PN_I32(topo), PN_F64(Nx), gravity.data(), PN_F64(rho), (int)topo.size(0), PN_F64(rhs), cur_stream()));
}
// simulator/cuda_utils.py:83-151 (calc_elastic + collect_rhs_IP)
void build_rhs(const double dx, const at::Tensor topo, const at::Tensor mu, const at::Tensor lam, const at::Tensor dNx, const at::Tensor dof, const int n_k,
               const at::Tensor adj_bgn, const at::Tensor adj, const int adj_slices, at::Tensor ip_stress, at::Tensor partial, at::Tensor rhs) {
    pn_check(pn_qgmls_build_rhs(dx, PN_I32(topo), PN_F64(mu), PN_F64(lam), PN_F64(dNx), PN_F64(dof), (int)topo.size(0), n_k, PN_I32(adj_bgn), PN_I32(adj),
                                adj_slices, PN_F64(ip_stress), PN_F64(partial), PN_F64(rhs), cur_stream()));
}
void matvec3(const at::Tensor mat, const at::Tensor x, at::Tensor y) { pn_check(pn_qgmls_matvec3(PN_F64(mat), PN_F64(x), (int)mat.size(0), PN_F64(y), cur_stream())); }
int64_t step_scratch_doubles(const int n_ip, const int n_k, const int adj_slices) { return (int64_t)pn_qgmls_step_scratch_doubles(n_ip, n_k, adj_slices); }
// simulator/solver.py:574-602 as one call (dense pre-inverted matrix; A / active only for the PCG variant)
void step(const int iters, const double dt, const double dx, const at::Tensor topo, const at::Tensor mu, const at::Tensor lam, const at::Tensor dNx,
          const at::Tensor adj_bgn, const at::Tensor adj, const int adj_slices, const at::Tensor Ainv, const at::Tensor M, c10::optional<at::Tensor> A,
          c10::optional<at::Tensor> active, const int pcg_iters, const at::Tensor dof_rest, const at::Tensor dof_f, const at::Tensor rhs_rest,
          const at::Tensor rhs_gravity, at::Tensor dof, at::Tensor dof_vel, at::Tensor scratch, const int solver) {
    pn_qgmls_step_t s{};
    s.n_ip = (int)topo.size(0); s.n_k = (int)(M.size(0) / 10); s.iters = iters; s.dt = dt; s.dx = dx;
    s.topo = PN_I32(topo); s.mu = PN_F64(mu); s.lam = PN_F64(lam); s.dNx = PN_F64(dNx);
    s.adj_bgn = PN_I32(adj_bgn); s.adj = PN_I32(adj); s.adj_slices = adj_slices;
    s.Ainv = PN_F64(Ainv); s.M = PN_F64(M); s.A = opt_ptr<double>(A); s.active = opt_ptr<unsigned char>(active); s.pcg_iters = pcg_iters;
    s.dof_rest = PN_F64(dof_rest); s.dof_f = PN_F64(dof_f); s.rhs_rest = PN_F64(rhs_rest); s.rhs_gravity = PN_F64(rhs_gravity);
    s.dof = PN_F64(dof); s.dof_vel = PN_F64(dof_vel); s.scratch = PN_F64(scratch);
    TORCH_CHECK((uint64_t)scratch.numel() >= pn_qgmls_step_scratch_doubles(s.n_ip, s.n_k, adj_slices), "scratch too small (step_scratch_doubles)");
    pn_check(pn_qgmls_step(&s, solver, cur_stream()));
}
// simulator/solver.py:402-424 + cuda_utils.py:206-233
void ip_info(const at::Tensor topo, const at::Tensor dof, const at::Tensor Nx, const at::Tensor dNx, const at::Tensor ddNx, at::Tensor pos, at::Tensor F,
             at::Tensor dF) {
    pn_check(pn_qgmls_ip_info(PN_I32(topo), PN_F64(dof), PN_F64(Nx), PN_F64(dNx), PN_F64(ddNx), (int)topo.size(0), PN_F32(pos), PN_F32(F), PN_F32(dF),
                              cur_stream()));
}
// simulator/cuda_utils.py:191-203
void update_pos(const at::Tensor topo, const at::Tensor dof, const at::Tensor Nx, at::Tensor pos) {
    pn_check(pn_qgmls_update_pos(PN_I32(topo), PN_F64(dof), PN_F64(Nx), (int)topo.size(0), PN_F64(pos), cur_stream()));
}
// simulator/solver.py:578-593 (vid < 0 clears)
void update_force(const int vid, const std::vector<double> f, const at::Tensor topo, const at::Tensor Nx, const at::Tensor rho, const double dx, at::Tensor dof_f) {
    TORCH_CHECK(vid < 0 || f.size() == 3, "force must have 3 components");
    pn_check(pn_qgmls_update_force(vid, vid < 0 ? nullptr : f.data(), PN_I32(topo), PN_F64(Nx), PN_F64(rho), dx, (int)(dof_f.numel() / 3), PN_F64(dof_f),
                                   cur_stream()));
}

PYBIND11_MODULE(TORCH_EXTENSION_NAME, m) {
    m.def("shape_functions", &shape_functions); m.def("collect_param", &collect_param); m.def("build_ip_global", &build_ip_global);
    m.def("build_pin_global", &build_pin_global); m.def("collect_gravity", &collect_gravity); m.def("build_rhs", &build_rhs);
    m.def("matvec3", &matvec3); m.def("step_scratch_doubles", &step_scratch_doubles); m.def("step", &step); m.def("ip_info", &ip_info);
    m.def("update_pos", &update_pos); m.def("update_force", &update_force);
}
#endif
