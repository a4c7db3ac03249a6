/*
 * pienerf_b200 C-ABI — the drop-in boundary of the B200-native PIE-NeRF hot path.
 *
 * One shared library (pienerf_b200/lib/libpienerf_b200.so), plain pointers and sizes only.
 * Every pointer is a DEVICE pointer unless the name ends in `_host`.  Every function enqueues
 * on `stream` (a cudaStream_t passed as void*; NULL = legacy default stream) and returns
 * 0 on success or a negative PN_E* code; pn_last_error() gives the message.  Nothing here
 * synchronises the host unless stated.
 *
 * Each entry point cites the reference interface it replaces (paths relative to the reference
 * repo).  Group A are 1:1 replacements for the reference's pybind entry points (same argument
 * order minus at::Tensor wrappers); group B replaces the Warp kernels / torch glue of the
 * simulator and of rund_cuda's per-frame preparation; group C are the fused B200 fast paths.
 */
#ifndef PIENERF_B200_H
#define PIENERF_B200_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define PN_OK 0
#define PN_EINVAL (-1)   /* bad argument (unsupported D/C/degree, null pointer, ...) */
#define PN_ECUDA (-2)    /* CUDA runtime error; see pn_last_error() */
#define PN_ENOTIMPL (-3) /* combination the B200 library does not provide (e.g. fp64 tables) */

const char *pn_last_error(void);
int pn_version(void);
/* device properties the host side sizes persistent grids with */
int pn_device_sm_count(int *sm_count);

/* ---------------------------------------------------------------- A. _gridencoder */
/* gridencoder/src/gridencoder.h:12 + gridencoder.cu:448-471 grid_encode_forward.
 * inputs [B,D] f32 in [0,1]; embeddings [sO,C] f32 (emb_half=0) or f16 (emb_half=1);
 * offsets [L+1] i32; outputs [L,B,C] same dtype as embeddings; dy_dx [B,L*D*C] or NULL. */
int pn_grid_encode_forward(const float *inputs, const void *embeddings, const int *offsets, void *outputs,
                           uint32_t B, uint32_t D, uint32_t C, uint32_t L, float S, uint32_t H, void *dy_dx,
                           uint32_t gridtype, int align_corners, uint32_t interp, int emb_half, void *stream);
/* Training side.
 * gridencoder/src/gridencoder.h:13 + gridencoder.cu:473-503 grid_encode_backward.  grad [L,B,C], grad_embeddings [sO,C]
 * (caller-zeroed, accumulated into), dy_dx [B,L*D*C] / grad_inputs [B,D] (both or neither; grad_inputs is overwritten):
 * all of the embeddings' dtype (emb_half).  inputs stay f32.  Sums land through atomics: run-to-run rounding differs. */
int pn_grid_encode_backward(const void *grad, const float *inputs, const void *embeddings, const int *offsets,
                            void *grad_embeddings, uint32_t B, uint32_t D, uint32_t C, uint32_t L, float S, uint32_t H,
                            const void *dy_dx, void *grad_inputs, uint32_t gridtype, int align_corners, uint32_t interp,
                            int emb_half, void *stream);
/* gridencoder.h:15 + gridencoder.cu:639-645 grad_total_variation.  inputs [B,D] in [0,1] OF THE EMBEDDINGS' DTYPE (as the
 * reference reads them); grad [sO,C] is accumulated into. */
int pn_grad_total_variation(const void *inputs, const void *embeddings, void *grad, const int *offsets, float weight,
                            uint32_t B, uint32_t D, uint32_t C, uint32_t L, float S, uint32_t H, uint32_t gridtype,
                            int align_corners, int emb_half, void *stream);

/* ---------------------------------------------------------------- A. _shencoder */
/* shencoder/src/shencoder.h:7 + shencoder.cu:400-417 sh_encode_forward.
 * inputs [B,D>=3] f32; outputs [B,C*C] f32; C = degree in 1..8; dy_dx [B,D*C*C] or NULL. */
int pn_sh_encode_forward(const float *inputs, float *outputs, uint32_t B, uint32_t D, uint32_t C, float *dy_dx,
                         void *stream);
/* Training side: shencoder.h:10 + shencoder.cu:419-438 sh_encode_backward.  grad [B,C*C], dy_dx [B,D*C*C] as written by
 * pn_sh_encode_forward (D must be 3 there); grad_inputs [B,D] is ACCUMULATED into (caller zeroes it), as in the reference. */
int pn_sh_encode_backward(const float *grad, const float *inputs, uint32_t B, uint32_t D, uint32_t C, const float *dy_dx,
                          float *grad_inputs, void *stream);

/* ---------------------------------------------------------------- A. _raymarching */
/* raymarching/src/raymarching.h:7 + raymarching.cu:151-159 */
int pn_near_far_from_aabb(const float *rays_o, const float *rays_d, const float *aabb, uint32_t N, float min_near,
                          float *nears, float *fars, void *stream);
/* raymarching.cu:204-212 */
int pn_sph_from_ray(const float *rays_o, const float *rays_d, float radius, uint32_t N, float *coords, void *stream);
/* raymarching.cu:232-235 / 260-263 */
int pn_morton3D(const int *coords, uint32_t N, int *indices, void *stream);
int pn_morton3D_invert(const int *indices, uint32_t N, int *coords, void *stream);
/* raymarching.cu:295-303; grid [N*8] f32, bitfield [N] u8 */
int pn_packbits(const float *grid, uint32_t N, float density_thresh, uint8_t *bitfield, void *stream);
/* raymarching.cu:812-824 */
int pn_march_rays(uint32_t n_alive, uint32_t n_step, const int *rays_alive, const float *rays_t, const float *rays_o,
                  const float *rays_d, float bound, float dt_gamma, uint32_t max_steps, uint32_t C, uint32_t H,
                  const uint8_t *grid, const float *nears, const float *fars, float *xyzs, float *dirs, float *deltas,
                  const float *noises, void *stream);
/* raymarching.cu:917-923 */
int pn_composite_rays(uint32_t n_alive, uint32_t n_step, float T_thresh, int *rays_alive, float *rays_t,
                      const float *sigmas, const float *rgbs, const float *deltas, float *weights_sum, float *depth,
                      float *image, void *stream);
/* raymarching.h:20-36 + raymarching.cu:1436-1489, same argument order (p_def before p_ori).
 * Unlike the reference wrapper this does NOT cudaEventSynchronize. */
int pn_march_rays_quadratic_bending(const int *pig_cnt, const int *pig_bgn, const int *pig_idx, int n_vtx, int n_grid,
                                    const float *p_def, const float *p_ori, const float *F_IP, const float *dF_IP,
                                    int max_iter_num, const float *bbmin, const float *bbmax, float hgs,
                                    const int *resolution, int num_seek_IP, float IP_dx, int cut,
                                    const float *cut_bounds, uint32_t n_alive, uint32_t n_step, const int *rays_alive,
                                    const float *rays_t, const float *rays_o, const float *rays_d, float bound,
                                    float dt_gamma, uint32_t max_steps, uint32_t C, uint32_t H, const uint8_t *grid,
                                    const float *nears, const float *fars, float *xyzs, float *dirs, float *deltas,
                                    const float *noises, void *stream);
/* Training side.
 * raymarching.h:13 + raymarching.cu:485-493 march_rays_train.  rays_o/rays_d [N,3], nears/fars/noises [N], grid = density
 * bitfield; xyzs/dirs [M,3], deltas [M,2] (caller-zeroed), rays [N,3] i32 = (ray, offset, num_steps), counter [2] i32 =
 * (points, rays), advanced as the reference's atomics advance it.  Packing is deterministic here: row n of `rays` is ray n
 * and offsets ascend with n (the reference's order is the order its atomics retire in); a ray whose samples do not fit in M
 * is dropped whole, as in the reference.  Three launches, no host synchronisation. */
int pn_march_rays_train(const float *rays_o, const float *rays_d, const uint8_t *grid, float bound, float dt_gamma,
                        uint32_t max_steps, uint32_t N, uint32_t C, uint32_t H, uint32_t M, const float *nears,
                        const float *fars, float *xyzs, float *dirs, float *deltas, int *rays, int *counter,
                        const float *noises, void *stream);
/* A/B switch for march_rays_train's empty-space skipping over 8^3 / 4^3 voxel blocks (single cascade; on by default).  The samples
 * are the reference's either way (tests assert bit-equality both ways); returns the previous setting. */
int pn_set_train_block_skip(int on);
/* Write pass of march_rays_train: 1 = warp-cooperative flush through shared memory (default), 0 = per-lane stream stores.  Same
 * output; returns the previous setting. */
int pn_set_train_write_mode(int mode);
/* raymarching.h:14 + raymarching.cu:583-591: sigmas [M], rgbs [M,3], deltas [M,2], rays [N,3] -> weights_sum/depth [N],
 * image [N,3], indexed by rays[:,0]. */
int pn_composite_rays_train_forward(const float *sigmas, const float *rgbs, const float *deltas, const int *rays, uint32_t M,
                                    uint32_t N, float T_thresh, float *weights_sum, float *depth, float *image,
                                    void *stream);
/* raymarching.h:15 + raymarching.cu:688-696: grad_sigmas [M] / grad_rgbs [M,3] caller-zeroed; no depth gradient. */
int pn_composite_rays_train_backward(const float *grad_weights_sum, const float *grad_image, const float *sigmas,
                                     const float *rgbs, const float *deltas, const int *rays, const float *weights_sum,
                                     const float *image, uint32_t M, uint32_t N, float T_thresh, float *grad_sigmas,
                                     float *grad_rgbs, void *stream);

/* ---------------------------------------------------------------- B. per-frame preparation */
/* nerf/utils.py:55-138 get_rays (N=-1, B=1): pose_host = 16 floats row-major cam2world (HOST). */
int pn_get_rays(const float *pose_host, float fx, float fy, float cx, float cy, uint32_t H, uint32_t W, float *rays_o,
                float *rays_d, void *stream);
/* Same rays for a subset of the pixels with the camera in DEVICE memory (so the launch can live in a CUDA graph):
 * cam [20] f32 = pose (16, row-major cam2world) | fx fy cx cy; pix [n_rays] i32 row-major pixel indices (NULL: all H*W). */
int pn_get_rays_pix(const float *cam, uint32_t H, uint32_t W, const int *pix, uint32_t n_rays, float *rays_o,
                    float *rays_d, void *stream);
/* nerf/utils.py:355-443 get_pnts_in_grids: deterministic counting sort (ascending IP index inside a
 * cell; the reference's order is atomic-race dependent).  bbmin [3] f32 and resolution [3] i32 are DEVICE
 * arrays as in the reference call; n_grid = res0*res1*res2 sizes pig_cnt/pig_bgn [n_grid]; pig_idx [n_vtx]. */
int pn_build_ip_grid(const float *p_def, int n_vtx, const float *bbmin, float hgs, const int *resolution, int n_grid,
                     int *pig_cnt, int *pig_bgn, int *pig_idx, void *stream);
/* nerf/renderer.py:782-791: bbox of p_def +-1e-3 (or +-bound if cut) and resolution=ceil((max-min)/hgs).
 * Writes bbmin[3], bbmax[3] (f32, device) and resolution[3] (i32, device). */
int pn_ip_bbox(const float *p_def, int n_vtx, float hgs, int cut, float bound, float *bbmin, float *bbmax,
               int *resolution, void *stream);

/* ---------------------------------------------------------------- C. fused render */
/* nerf/network.py:98-127 NeRFNetwork.forward as ONE kernel over M samples:
 * hash-grid (L=16,C=2,D=3) -> 32->64->16 -> trunc_exp | SH(4) || geo(15) -> 31->64->64->3 -> sigmoid.
 * weights: w_sigma0 [64,32], w_sigma1 [16,64], w_color0 [64,31], w_color1 [64,64], w_color2 [3,64] (torch
 * nn.Linear row-major [out,in], f32).  xyzs in [-bound,bound].
 * mode 0 = fp32 SIMT; mode 1 = tcgen05 tensor cores, each GEMM as 3 bf16 MMAs (hi/lo split) with fp32 TMEM accumulators. */
typedef struct {
    const float *embeddings; const int *offsets; float S; uint32_t H; uint32_t L; float bound;
    const float *w_sigma0, *w_sigma1, *w_color0, *w_color1, *w_color2;
} pn_field_t;
int pn_field_forward(const pn_field_t *field_host, const float *xyzs, const float *dirs, uint32_t M, float *sigmas,
                     float *rgbs, int mode, void *stream);

/* nerf/network.py:105-127 alone (sigma_net 32-64-16, trunc_exp, SH(4) of dirs || geo, color_net 31-64-64-3, sigmoid) on the
 * tcgen05 pipeline of the frame renderer, hash-grid gather replaced by a load of pre-encoded features enc [M,32] f32 (the
 * [B, L*C] layout grid.py:57 hands to the MLP).  For measuring the MLP pass by itself; pn_set_profile_event_list events [0], [1]
 * bracket the tensor-core kernel. */
uint64_t pn_mlp_workspace_bytes(uint32_t M);
int pn_mlp_forward(const pn_field_t *field_host, const float *enc, const float *dirs, uint32_t M, float *sigmas, float *rgbs,
                   void *workspace, uint64_t workspace_bytes, void *stream);

/* nerf/renderer.py:755-907 rund_cuda as a device-resident frame (enqueue-only, no host sync): near/far, IP bbox +
 * grid, then the render itself.  mode 3 (product path) = wavefront: per pass a march kernel (lattice march +
 * inverse warp, samples appended to a compact list), the field kernel (hash-grid gather + tcgen05 MLP over 128-row
 * tiles) and a per-ray compositor; modes 0-2 = single fused persistent kernels (0: tcgen05 MLP, 1: fp32 SIMT MLP,
 * 2: one lane per ray).  Outputs image [N,3], depth [N], depth_0 [N], weights_sum [N] as rund_cuda returns them.
 * stats (optional, device int64[8]): [0] composited samples, [1] rays that hit the aabb, [2] field evaluations,
 * [3] rows the field kernel processed (mode 3; samples + slab padding), [4] error bits (1: the deformed IP bbox needed more
 * grid cells than the workspace holds and was clamped — diverged body; 2: rays were cut short because the sample list and
 * the passes ran out), [5] number of such rays, [6] chunks a full sample list deferred to a later pass, [7] passes that had rays. */
typedef struct {
    const float *p_def, *p_ori, *F_IP, *dF_IP; int n_vtx; float IP_dx;
    const uint8_t *density_bitfield; float bound; uint32_t cascade; uint32_t grid_size;
    float min_near, density_scale, dt_gamma; uint32_t max_steps; float T_thresh; int max_iter_num;
    float hgs; int cut; float cut_bounds[6]; int num_seek_IP; float bg_color;
} pn_deform_t;
int pn_render_deformed(const pn_field_t *field_host, const pn_deform_t *deform_host, const float *rays_o,
                       const float *rays_d, uint32_t N, float *image, float *depth, float *depth_0, float *weights_sum,
                       void *workspace, uint64_t workspace_bytes, long long *stats, int mode, void *stream);
/* The same frame as one rank's share of a multi-GPU frame (trainer.py:284-329 ordering is the caller's):
 * io->pix [N] maps ray -> pixel of the frame-sized image / depth / depth_0 (which may be PEER memory: another GPU's frame
 * opened with pn_peer_open — the compositor's stores then cross NVLink and no gather is needed); weights_sum stays [N].
 * io->epoch: device counter of this frame slot, bumped by the first kernel of the call; the call then waits until every
 * *wait_flag[i] >= epoch (flags in THIS GPU's memory, raised by a peer: "this frame's IP state has landed") and, after
 * the last output is written, stores epoch into every signal_flag[i] (usually one word in the frame owner's memory).
 * A wait that exceeds timeout_ms stores a nonzero code in *status and lets the stream continue (no GPU hang).
 * wait_flag / signal_flag are DEVICE arrays of pointers.  io == NULL: plain pn_render_deformed. */
typedef struct {
    const int *pix;
    uint32_t *epoch;
    const uint32_t *const *wait_flag; int n_wait;
    uint32_t *const *signal_flag; int n_signal;
    int *status; uint32_t timeout_ms;
    const float *noises;   /* [N] or NULL: perturb (renderer.py:863): ray n starts at near + noises[n] * dt for its first sample */
    int max_passes;        /* > 0: at most this many wavefront passes, the last one unbounded (stats[7] of earlier frames tells how many are used) */
    uint32_t flags;        /* PN_IO_* */
} pn_frame_io_t;
#define PN_IO_WEIGHTS_READY 1u   /* the workspace already holds the weight image of these weights (an earlier call with the same workspace) */
int pn_render_deformed_ex(const pn_field_t *field_host, const pn_deform_t *deform_host, const float *rays_o,
                          const float *rays_d, uint32_t N, float *image, float *depth, float *depth_0,
                          float *weights_sum, void *workspace, uint64_t workspace_bytes, long long *stats, int mode,
                          const pn_frame_io_t *io_host, void *stream);

/* ---------------------------------------------------------------- D. peer memory (multi-GPU frame)
 * The reference is single-GPU; these replace the NCCL broadcast / gather a straightforward port would use for the
 * frame's two exchanges (IP state out, pixels back) with stores into peer memory + epoch flags.
 * pn_peer_alloc: zero-filled cudaMalloc memory that can be exported; pn_peer_export / pn_peer_open: CUDA IPC handle
 * (64 bytes) out / in — the opened pointer is valid in kernels of the opening process (NVLink peer access enabled
 * lazily).  All three synchronise the device; they are set-up calls. */
int pn_peer_alloc(uint64_t bytes, void **ptr);
int pn_peer_free(void *ptr);
int pn_peer_export(void *ptr, unsigned char *handle64_host);
int pn_peer_open(const unsigned char *handle64_host, void **ptr);
int pn_peer_close(void *ptr);
/* One launch: copy `bytes` (multiple of 16) from src into each of the n_dst destinations (device array of pointers,
 * local or peer) with 128-bit stores. */
int pn_peer_put(const void *src, uint64_t bytes, void *const *dsts_dev, int n_dst, void *stream);
/* Epoch flags.  wait: (bump ? ++*epoch : *epoch) = e, then spin until *flags[i] + lag >= e for all i (timeout -> *status).
 * signal: system-scope fence, then *flags[i] = *epoch for all i.  flags_dev: device array of (local or peer) pointers. */
int pn_epoch_wait(uint32_t *epoch, int bump, const uint32_t *const *flags_dev, int n_flags, uint32_t lag, int *status,
                  uint32_t timeout_ms, void *stream);
int pn_epoch_signal(const uint32_t *epoch, uint32_t *const *flags_dev, int n_flags, void *stream);

/* Rows of the per-pass sample list of mode 3 (0 = automatic: 24 per ray, 1 Mi .. 32 Mi).  Call before sizing the workspace with
 * pn_render_workspace_bytes.  A full list defers the remaining rays to the next pass (stats[6]); tests use a small list. */
int pn_set_wave_capacity(int rows);
/* Per-frame IP preparation of pn_render_deformed: 0 (default) = one single-CTA kernel when the IP-grid capacity is <= 64 Ki cells
 * (every configuration without --cut), 1 = always the multi-kernel chain (bbox, counting sort, pack, neighbourhood lists). */
int pn_set_prep_mode(int force_multi_kernel);
/* Size the persistent march / field grids of mode 3 for (SM count - n_sm) SMs, leaving n_sm SMs' worth of CTA slots to kernels of
 * other streams (the simulator's launches on the GPU that also renders).  0 = use every SM (default).  Process-wide; a CUDA
 * graph captured afterwards keeps the grid sizes it was captured with. */
int pn_set_render_sm_reserve(int n_sm);
/* Optional profiling hook: two cudaEvent_t (as void*) that pn_render_deformed records on its stream right
 * around the persistent render kernel (NULL, NULL disables).  Used by bench.py for the roofline line.
 * Process-wide and unsynchronised, like the other pn_set_* switches: a measurement aid for one host thread driving one
 * renderer; set it before the calls it should bracket and clear it before another thread renders. */
int pn_set_profile_events(void *start_event, void *stop_event);
/* Same for mode 3: events[2k], events[2k+1] (cudaEvent_t) bracket the k-th field-kernel launch of a frame, k < n/2;
 * the pair of pn_set_profile_events then brackets all passes.  (NULL, 0) disables.  The array must stay alive. */
int pn_set_profile_event_list(void **events, int n);
/* number of (march, field, composite) passes mode 3 enqueues for a per-ray sample cap of max_steps (pass caps 32, 64, ...
 * plus one spare pass for the < 32-sample overshoot of a pass) */
int pn_render_pass_count(uint32_t max_steps);
/* bytes of scratch pn_render_deformed needs for N rays, n_vtx IPs, scene bound and IP-grid cell size hgs */
uint64_t pn_render_workspace_bytes(uint32_t N, int n_vtx, float bound, float hgs);

/* ---------------------------------------------------------------- B. _qgmls (simulator, fp64) */
/* simulator/cpu_utils.py:3-152 calc_G/calc_Gp/calc_weight + solver.py:334-399 init_GMLS.
 * pos [n,3], topo [n,8], kernel_pos [nk,3]; Nx [n,8,10], dNx [n,8,3,10], ddNx [n,8,3,3,10]
 * (zero-filled here; dNx/ddNx may be NULL).  status: device int, set nonzero on a singular G. */
int pn_qgmls_shape_functions(double r, const double *pos, const int *topo, const double *kernel_pos, int n, double *Nx,
                             double *dNx, double *ddNx, int *status, void *stream);
/* simulator/cuda_utils.py:3-19 collect_param (then the /rho, /dx^3 of solver.py:450). */
int pn_qgmls_collect_param(const int *pts_ip, const double *mu, const double *lam, const double *mass, int n_pts,
                           int n_ip, double dx, double *ip_mu, double *ip_lam, double *ip_rho, void *stream);
/* simulator/cuda_utils.py:22-55 build_IP_global into mat [n,n] (n = 10*n_k), += semantics. */
int pn_qgmls_build_ip_global(double dx, double dt, const int *topo, const double *mu, const double *lam,
                             const double *rho, const double *Nx, const double *dNx, const double *ddNx, int n_ip,
                             int n, double *mat, void *stream);
/* simulator/cuda_utils.py:58-81 build_pin_global */
int pn_qgmls_build_pin_global(double stiff, const int *vidx, int n_pin, const int *topo, const double *Nx, int n,
                              double *mat, void *stream);
/* simulator/cuda_utils.py:262-279 collect_gravity into rhs [n,3] (+=) */
int pn_qgmls_collect_gravity(double dx, const int *topo, const double *Nx, const double *gravity_host,
                             const double *rho, int n_ip, double *rhs, void *stream);
/* simulator/cuda_utils.py:83-151 calc_elastic + collect_rhs_IP fused; rhs [n,3] is overwritten.
 * adjacency = the kernel->(ip,corner) CSR of solver.py:279-313 (adj_bgn [n_k+1], adj [tot] = ip*8+corner);
 * deterministic gather instead of fp64 atomics.  adj_slices = ceil(max entries per kernel / 128);
 * ip_stress [n_ip,9] and partial [n_k, adj_slices, 30] are scratch. */
int pn_qgmls_build_rhs(double dx, const int *topo, const double *mu, const double *lam, const double *dNx,
                       const double *dof, int n_ip, int n_k, const int *adj_bgn, const int *adj, int adj_slices,
                       double *ip_stress, double *partial, double *rhs, void *stream);
/* y[n,3] = Mat[n,n] x[n,3]: the compact form of `(Mat (x) I3) @ v` (solver.py:576,600). */
int pn_qgmls_matvec3(const double *mat, const double *x, int n, double *y, void *stream);
/* simulator/solver.py:574-576,595-602 stepforward (iters local-global iterations) as one call. */
typedef struct {
    int n_ip, n_k, iters; double dt, dx;
    const int *topo; const double *mu, *lam, *dNx;   /* IP_kernel [n_ip,8], IP_mu/lam [n_ip], IP_dNx [n_ip,8,3,10] */
    const int *adj_bgn, *adj; int adj_slices;        /* kernel -> (ip*8+corner) CSR; ceil(max entries per kernel / 128) */
    const double *Ainv, *M;                          /* [n,n] compact: global_matrix / mass_matrix_invt2 without (x)I3 */
    const double *A; const unsigned char *active;    /* PCG only: assembled system [n,n] (no +1e-3) and active kernels [n_k] */
    int pcg_iters;
    const double *dof_rest, *dof_f, *rhs_rest, *rhs_gravity;   /* [n,3] */
    double *dof, *dof_vel;                           /* in/out [n,3] */
    double *scratch;                                 /* >= pn_qgmls_step_scratch_doubles(n_ip, n_k, adj_slices) doubles */
} pn_qgmls_step_t;
uint64_t pn_qgmls_step_scratch_doubles(int n_ip, int n_k, int adj_slices);
int pn_qgmls_step(const pn_qgmls_step_t *step_host, int solver /*0 dense inverse, 1 PCG on A*/, void *stream);
/* How pn_qgmls_step (solver 0) runs: force_multi_kernel = 1 (default) = 3 + 4*iters launches; 0 = ONE thread-block-cluster
 * kernel for the whole step when the system is small enough (n <= 1280; experimental, measured slower on B200).  The two
 * differ by fp64 round-off (different fixed summation order in the rhs gather); each is bit-reproducible. */
int pn_qgmls_step_mode(int force_multi_kernel);
/* kernels one pn_qgmls_step enqueues for this problem size (1 when the cluster kernel applies) */
int pn_qgmls_step_launches(int n_ip, int n_k, int iters, int solver, int pcg_iters);
/* simulator/solver.py:402-424 get_IP_info + cuda_utils.py:206-233 update_F_kernel: emits the fp32
 * renderer layouts directly: pos [n,3], F [n,9] (F[b][a] at a*3+b), dF [n,27] (c*9+r*3+j). */
int pn_qgmls_ip_info(const int *topo, const double *dof, const double *Nx, const double *dNx, const double *ddNx,
                     int n_ip, float *pos, float *F, float *dF, void *stream);
/* simulator/cuda_utils.py:191-203 update_pos_kernel: pos [n_pts,3] f64 */
int pn_qgmls_update_pos(const int *topo, const double *dof, const double *Nx, int n_pts, double *pos, void *stream);
/* simulator/solver.py:578-588 update_force for one IP (vid<0 clears): dof_f [n,3] overwritten. */
int pn_qgmls_update_force(int vid, const double *f_host, const int *topo, const double *Nx, const double *rho,
                          double dx, int n, double *dof_f, void *stream);

#ifdef __cplusplus
}
#endif
#endif
