// Device-resident deformed-space frame renderer + its per-frame preparation kernels.
//
// Replaces the host-driven wavefront loop of nerf/renderer.py:755-907 (rund_cuda) and the torch / Warp glue
// around it (nerf/utils.py:55-138 get_rays, :355-443 get_pnts_in_grids, renderer.py:782-797 bbox + near/far)
// with one enqueue-only call and no host synchronisation:
//   1. ip_bbox_kernel        bbox of deformed IP centres -> bbmin/bbmax/resolution (device)
//   2. ip_grid_*             deterministic counting sort of IPs into the hgs grid
//   3. frame_setup_kernel    near/far, output init, compaction of the rays that hit the IP box
//   4. ip_pack / nb_*        cell-sorted IP records (with F^-1) and the per-cell neighbourhood lists (render_warp.cuh)
//   5. the render itself     mode 3 (product path): wavefront passes of march / field / composite kernels (render_wave.cuh)
//                            mode 0/1: one fused persistent kernel, warp = ray (render_warp.cuh)
//                            mode 2: render_persistent below, one lane = one ray with the reference's own search order
//                            (slow; kept as the in-tree cross-check of the march decisions)
// The per-ray sample sequence is exactly the reference's in every mode (the reference's loop only batches it
// differently), so images agree to MLP rounding; see DESIGN.md for the documented deviations (per-ray cap = max_steps).
#include <cstdlib>
#include "field_device.cuh"
#include "march_device.cuh"

namespace {

// ------------------------------------------------------------------------------------------- get_rays
__global__ void __launch_bounds__(256) get_rays_kernel(float r00, float r01, float r02, float r10, float r11, float r12,
                                                       float r20, float r21, float r22, float tx, float ty, float tz,
                                                       float fx, float fy, float cx, float cy, uint32_t H, uint32_t W,
                                                       float *__restrict__ rays_o, float *__restrict__ rays_d) {
    const uint32_t n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= H * W) return;
    const float i = (float)(n % W) + 0.5f, j = (float)(n / W) + 0.5f;
    const float xs = (i - cx) / fx * 1.0f, ys = (j - cy) / fy * 1.0f, zs = 1.0f;
    const float nrm = sqrtf(xs * xs + ys * ys + zs * zs);
    const float ux = xs / nrm, uy = ys / nrm, uz = zs / nrm;
    rays_d[3 * n] = ux * r00 + uy * r01 + uz * r02;
    rays_d[3 * n + 1] = ux * r10 + uy * r11 + uz * r12;
    rays_d[3 * n + 2] = ux * r20 + uy * r21 + uz * r22;
    rays_o[3 * n] = tx; rays_o[3 * n + 1] = ty; rays_o[3 * n + 2] = tz;
}

// Same arithmetic for a subset of the pixels (one rank's image tiles), camera read from DEVICE memory so that the launch
// can sit in a CUDA graph: cam = [pose 16 floats row-major | fx fy cx cy].  pix == nullptr: all H*W pixels in order.
__global__ void __launch_bounds__(256) get_rays_pix_kernel(const float *__restrict__ cam, uint32_t W, const int *__restrict__ pix,
                                                           uint32_t n_rays, float *__restrict__ rays_o, float *__restrict__ rays_d) {
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n_rays) return;
    const uint32_t n = pix ? (uint32_t)pix[k] : k;
    const float fx = cam[16], fy = cam[17], cx = cam[18], cy = cam[19];
    const float i = (float)(n % W) + 0.5f, j = (float)(n / W) + 0.5f;
    const float xs = (i - cx) / fx * 1.0f, ys = (j - cy) / fy * 1.0f, zs = 1.0f;
    const float nrm = sqrtf(xs * xs + ys * ys + zs * zs);
    const float ux = xs / nrm, uy = ys / nrm, uz = zs / nrm;
    rays_d[3 * k] = ux * cam[0] + uy * cam[1] + uz * cam[2];
    rays_d[3 * k + 1] = ux * cam[4] + uy * cam[5] + uz * cam[6];
    rays_d[3 * k + 2] = ux * cam[8] + uy * cam[9] + uz * cam[10];
    rays_o[3 * k] = cam[3]; rays_o[3 * k + 1] = cam[7]; rays_o[3 * k + 2] = cam[11];
}

// ------------------------------------------------------------------------------------------- IP bbox
struct FrameGeom {  // lives in the workspace, written by ip_bbox_kernel
    float bbmin[3], bbmax[3], hi[3];
    int res[3];
    int n_grid;
    int overflow;   // 1: the IP bbox needed more cells than the workspace holds (diverged / out-of-scene body): resolution was clamped
};

// res_max > 0 (frame path): every resolution component is clamped to [1, res_max] so that res0*res1*res2 never exceeds the
// workspace's cell capacity (res_max^3, max_cells_for) — a diverged simulation (huge / NaN / inf positions) then renders
// garbage instead of indexing out of bounds, and FrameGeom.overflow reports it (stats[4] bit 0).
__global__ void __launch_bounds__(1024) ip_bbox_kernel(const float *__restrict__ p, int n, float hgs, int cut, float bound, int res_max,
                                                       FrameGeom *__restrict__ g, float *__restrict__ bbmin_out,
                                                       float *__restrict__ bbmax_out, int *__restrict__ res_out) {
    __shared__ float smin[3][32], smax[3][32];
    __shared__ int s_over[3];
    float lo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, hi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
#pragma unroll
        for (int c = 0; c < 3; c++) { const float v = p[3 * i + c]; lo[c] = fminf(lo[c], v); hi[c] = fmaxf(hi[c], v); }
    }
#pragma unroll
    for (int c = 0; c < 3; c++) {
        for (int o = 16; o > 0; o >>= 1) {
            lo[c] = fminf(lo[c], __shfl_xor_sync(0xffffffffu, lo[c], o));
            hi[c] = fmaxf(hi[c], __shfl_xor_sync(0xffffffffu, hi[c], o));
        }
        if ((threadIdx.x & 31) == 0) { smin[c][threadIdx.x >> 5] = lo[c]; smax[c][threadIdx.x >> 5] = hi[c]; }
    }
    __syncthreads();
    if (threadIdx.x < 3) {
        const int c = threadIdx.x;
        float a = FLT_MAX, b = -FLT_MAX;
        for (int w = 0; w < (int)(blockDim.x >> 5); w++) { a = fminf(a, smin[c][w]); b = fmaxf(b, smax[c][w]); }
        if (cut) { a = -bound; b = bound; }                       // renderer.py:784-786
        const float mn = a - 1e-3f, mx = b + 1e-3f;               // renderer.py:787-789
        // renderer.py:791: tensor / python-scalar on CUDA multiplies by the fp32 reciprocal
        const float rf = ceilf((mx - mn) * (1.0f / hgs));
        int r = (int)rf;
        s_over[c] = 0;
        if (res_max > 0 && !(rf >= 1.0f && rf <= (float)res_max)) { r = rf >= 1.0f ? res_max : 1; s_over[c] = 1; }   // also catches NaN
        if (g) { g->bbmin[c] = mn; g->bbmax[c] = mx; g->hi[c] = (float)((double)mx - 1e-6); g->res[c] = r; }
        if (bbmin_out) { bbmin_out[c] = mn; bbmax_out[c] = mx; res_out[c] = r; }
    }
    __syncthreads();
    if (threadIdx.x == 0 && g) { g->n_grid = g->res[0] * g->res[1] * g->res[2]; g->overflow = s_over[0] | s_over[1] | s_over[2]; }
}

// ------------------------------------------------------------------------------------------- IP grid
__device__ __forceinline__ int ip_cell(const float *__restrict__ p, int i, const float *bbmin, float hgs, const int *res) {
    // nerf/utils.py:388-408 p2g (true fp32 division)
    // the clamps only act when ip_bbox_kernel had to clamp the resolution (FrameGeom.overflow) or a coordinate is NaN
    const int g0 = min(max((int)floorf((p[3 * i] - bbmin[0]) / hgs), 0), res[0] - 1);
    const int g1 = min(max((int)floorf((p[3 * i + 1] - bbmin[1]) / hgs), 0), res[1] - 1);
    const int g2 = min(max((int)floorf((p[3 * i + 2] - bbmin[2]) / hgs), 0), res[2] - 1);
    return (g2 * res[1] + g1) * res[0] + g0;
}

__global__ void __launch_bounds__(256) ip_grid_count(const float *__restrict__ p, int n, const float *__restrict__ bbmin,
                                                     float hgs, const int *__restrict__ res, int n_grid_cap,
                                                     int *__restrict__ cnt) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int gid = ip_cell(p, i, bbmin, hgs, res);
    if (gid >= 0 && gid < n_grid_cap) atomicAdd(cnt + gid, 1);
}

// single-block exclusive scan over n_grid cells (n_grid <= a few 1e5): bgn = cumsum(cnt) - cnt
__global__ void __launch_bounds__(1024) ip_grid_scan(const int *__restrict__ cnt, const int *__restrict__ res,
                                                     int n_grid_cap, int *__restrict__ bgn, int *__restrict__ fill) {
    __shared__ int warp_tot[32];
    __shared__ int carry;
    const int n_grid = min(res[0] * res[1] * res[2], n_grid_cap);
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < n_grid; base += blockDim.x) {
        const int i = base + threadIdx.x;
        const int v = i < n_grid ? cnt[i] : 0;
        int incl = v;
        for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, o); if ((threadIdx.x & 31) >= o) incl += t; }
        if ((threadIdx.x & 31) == 31) warp_tot[threadIdx.x >> 5] = incl;
        __syncthreads();
        if (threadIdx.x < 32) {
            int w = warp_tot[threadIdx.x];
            for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, w, o); if (threadIdx.x >= o) w += t; }
            warp_tot[threadIdx.x] = w;
        }
        __syncthreads();
        const int before = carry + ((threadIdx.x >> 5) ? warp_tot[(threadIdx.x >> 5) - 1] : 0) + incl - v;
        if (i < n_grid) { bgn[i] = before; fill[i] = 0; }
        __syncthreads();
        if (threadIdx.x == blockDim.x - 1) carry = before + v;
        __syncthreads();
    }
}

__global__ void __launch_bounds__(256) ip_grid_fill(const float *__restrict__ p, int n, const float *__restrict__ bbmin,
                                                    float hgs, const int *__restrict__ res, int n_grid_cap,
                                                    const int *__restrict__ bgn, int *__restrict__ fill,
                                                    int *__restrict__ idx) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int gid = ip_cell(p, i, bbmin, hgs, res);
    if (gid >= 0 && gid < n_grid_cap) idx[bgn[gid] + atomicAdd(fill + gid, 1)] = i;
}

// make the within-cell order deterministic (ascending IP index); cells hold a handful of IPs
__global__ void __launch_bounds__(256) ip_grid_sort(const int *__restrict__ cnt, const int *__restrict__ bgn,
                                                    const int *__restrict__ res, int n_grid_cap, int *__restrict__ idx) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    const int n_grid = min(res[0] * res[1] * res[2], n_grid_cap);
    if (c >= n_grid) return;
    const int n = cnt[c];
    int *a = idx + bgn[c];
    for (int i = 1; i < n; i++) {
        const int v = a[i];
        int j = i - 1;
        while (j >= 0 && a[j] > v) { a[j + 1] = a[j]; j--; }
        a[j + 1] = v;
    }
}

// ------------------------------------------------------------------------------------------- frame setup
struct FrameQueue { int n_active; int next; long long samples; long long pad; };

__global__ void __launch_bounds__(256) frame_setup_kernel(const float *__restrict__ rays_o, const float *__restrict__ rays_d,
                                                          uint32_t N, const FrameGeom *__restrict__ g, float min_near,
                                                          float bg, float *__restrict__ nears, float *__restrict__ fars,
                                                          int *__restrict__ active, FrameQueue *__restrict__ q,
                                                          const int *__restrict__ pix, float *__restrict__ image,
                                                          float *__restrict__ depth, float *__restrict__ depth0,
                                                          float *__restrict__ wsum) {
    const uint32_t n = blockIdx.x * blockDim.x + threadIdx.x;
    bool hit = false;
    if (n < N) {
        const float ox = rays_o[3 * n], oy = rays_o[3 * n + 1], oz = rays_o[3 * n + 2];
        const float rdx = 1 / rays_d[3 * n], rdy = 1 / rays_d[3 * n + 1], rdz = 1 / rays_d[3 * n + 2];
        float near = (g->bbmin[0] - ox) * rdx, far = (g->bbmax[0] - ox) * rdx;
        if (near > far) { const float s = near; near = far; far = s; }
        float ny = (g->bbmin[1] - oy) * rdy, fy = (g->bbmax[1] - oy) * rdy;
        if (ny > fy) { const float s = ny; ny = fy; fy = s; }
        bool miss = near > fy || ny > far;
        if (!miss) {
            if (ny > near) near = ny;
            if (fy < far) far = fy;
            float nz = (g->bbmin[2] - oz) * rdz, fz = (g->bbmax[2] - oz) * rdz;
            if (nz > fz) { const float s = nz; nz = fz; fz = s; }
            miss = near > fz || nz > far;
            if (!miss) {
                if (nz > near) near = nz;
                if (fz < far) far = fz;
                if (near < min_near) near = min_near;
            }
        }
        if (miss) near = far = FLT_MAX;
        nears[n] = near; fars[n] = far;
        hit = near < far;  // a ray with t >= far never emits a sample
        if (!hit) {
            // what rund_cuda leaves for a ray that never composites anything (renderer.py:896-901)
            const size_t o = pix ? (size_t)pix[n] : (size_t)n;   // pixel of the (possibly remote) frame this ray belongs to
            image[3 * o] = bg; image[3 * o + 1] = bg; image[3 * o + 2] = bg;
            depth0[o] = 0.f; wsum[n] = 0.f;
            depth[o] = fmaxf(0.f - near, 0.f) / (far - near);  // NaN for a miss, as in the reference
        }
    }
    const uint32_t m = __ballot_sync(0xffffffffu, hit);
    if (m) {
        const int lane = threadIdx.x & 31;
        int base = 0;
        if (lane == 0) base = atomicAdd(&q->n_active, __popc(m));
        base = __shfl_sync(0xffffffffu, base, 0);
        if (hit) active[base + __popc(m & ((1u << lane) - 1))] = (int)n;
    }
}

// ------------------------------------------------------------------------------------------- persistent renderer
struct RenderArgs {
    pn_field_t field;
    pn::MarchCfg march;
    pn::BendCfg bend;  // bbmin/bbmax/hi/res filled on device from FrameGeom
    const FrameGeom *geom;
    const float *rays_o, *rays_d, *nears, *fars;
    const int *active;
    FrameQueue *queue;
    float *image, *depth, *depth0, *wsum;   // image / depth / depth0 are indexed by pix[ray] when pix != nullptr (frame-sized, maybe peer memory)
    const int *pix;
    const float *noises;                    // [N] or nullptr: start offset of each ray in units of its first step (perturb, raymarching.cu:1187)
    float density_scale, T_thresh, bg;
    uint32_t max_samples;
};

template <int KMAX>
__global__ void __launch_bounds__(128, 3) render_persistent(const RenderArgs A) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    pn::FieldBlockSmem &fbs = *reinterpret_cast<pn::FieldBlockSmem *>(smem_raw);
    pn::FieldSmem &fs = fbs.w;
    float *scratch = fbs.scratch + threadIdx.x;
    pn::field_smem_fill(fs, A.field);
    pn::BendCfg bc = A.bend;
#pragma unroll
    for (int i = 0; i < 3; i++) {
        bc.bbmin[i] = A.geom->bbmin[i]; bc.bbmax[i] = A.geom->bbmax[i]; bc.hi[i] = A.geom->hi[i]; bc.res[i] = A.geom->res[i];
    }
    __syncthreads();
    const pn::MarchCfg m = A.march;
    const float2 *table = reinterpret_cast<const float2 *>(A.field.embeddings);
    const int lane = threadIdx.x & 31;
    const int n_active = A.queue->n_active;

    // per-lane ray state
    bool alive = false, exhausted = false;
    int ray = -1;
    float ox = 0, oy = 0, oz = 0, dx = 0, dy = 0, dz = 0, rdx = 0, rdy = 0, rdz = 0;
    float t = 0, last_t = 0, far = 0, near = 0, tdepth = 0;
    float ws = 0, dep = 0, cr = 0, cg = 0, cb = 0;
    float sh[16];
    uint32_t nsamp = 0;
    long long my_samples = 0;

    while (true) {
        // ---- refill dead lanes from the global queue (one atomic per warp)
        const uint32_t need = __ballot_sync(0xffffffffu, !alive && !exhausted);
        if (need) {
            int base = 0;
            if (lane == __ffs(need) - 1) base = atomicAdd(&A.queue->next, __popc(need));
            base = __shfl_sync(0xffffffffu, base, __ffs(need) - 1);
            if (!alive && !exhausted) {
                const int slot = base + __popc(need & ((1u << lane) - 1));
                if (slot < n_active) {
                    ray = A.active[slot];
                    ox = A.rays_o[3 * ray]; oy = A.rays_o[3 * ray + 1]; oz = A.rays_o[3 * ray + 2];
                    dx = A.rays_d[3 * ray]; dy = A.rays_d[3 * ray + 1]; dz = A.rays_d[3 * ray + 2];
                    rdx = 1 / dx; rdy = 1 / dy; rdz = 1 / dz;
                    near = A.nears[ray]; far = A.fars[ray];
                    t = near;                     // rays_t = nears.clone(); noise = 0 (perturb off in the sim GUI)
                    last_t = t; tdepth = near;
                    ws = dep = cr = cg = cb = 0.f;
                    nsamp = 0;
                    pn::sh_eval<4>(dx, dy, dz, sh);
                    alive = true;
                } else {
                    exhausted = true;
                }
            }
        }
        if (!__any_sync(0xffffffffu, alive)) break;

        // ---- march this lane's ray to its next kept sample
        bool have = false;
        float x = 0, y = 0, z = 0, dt = 0, dreal = 0;
        if (alive) {
            while (t < far) {
                pn::deformed_sample(bc, ox, oy, oz, dx, dy, dz, t, x, y, z);
                const bool found = pn::bend_sample<KMAX>(bc, x, y, z);
                dt = pn::step_size(m, t);
                float tt;
                const bool occ = pn::occupancy_and_exit(m, x, y, z, t, dt, dx, dy, dz, rdx, rdy, rdz, tt);
                if (occ && found) {
                    t += dt;
                    dreal = t - last_t;
                    last_t = t;
                    have = true;
                    break;
                }
                do { t += pn::step_size(m, t); } while (t < tt);
            }
        }
        bool finish = alive && !have;

        // ---- field + composite for lanes that produced a sample
        if (have) {
            float sigma, r, g, b;
            pn::field_eval(fs, table, A.field.bound, x, y, z, sh, scratch, pn::kFieldThreads, sigma, r, g, b);
            sigma = A.density_scale * sigma;
            const float alpha = 1.0f - __expf(-sigma * dt);
            const float T = 1 - ws;
            const float w = alpha * T;
            ws += w;
            tdepth += dreal;
            dep += w * tdepth;
            cr += w * r; cg += w * g; cb += w * b;
            nsamp++;
            my_samples++;
            if (T < A.T_thresh || nsamp >= A.max_samples) finish = true;
        }
        if (finish) {
            // renderer.py:896-901
            A.image[3 * ray] = cr + (1 - ws) * A.bg;
            A.image[3 * ray + 1] = cg + (1 - ws) * A.bg;
            A.image[3 * ray + 2] = cb + (1 - ws) * A.bg;
            A.depth0[ray] = dep;
            A.depth[ray] = fmaxf(dep - near, 0.f) / (far - near);
            A.wsum[ray] = ws;
            alive = false;
        }
    }
    for (int o = 16; o > 0; o >>= 1) my_samples += __shfl_xor_sync(0xffffffffu, my_samples, o);
    if (lane == 0 && my_samples) atomicAdd((unsigned long long *)&A.queue->samples, (unsigned long long)my_samples);
}

__global__ void copy_stats_kernel(const FrameQueue *q, const FrameGeom *g, long long *stats) {
    stats[0] = q->samples;   // composited samples
    stats[1] = q->n_active;  // rays that hit the IP box
    stats[2] = q->pad;       // field evaluations (>= stats[0]: samples marched past an early termination)
    stats[3] = 0;
    stats[4] = g->overflow ? 1 : 0;
    stats[5] = stats[6] = stats[7] = 0;
}

// stand-alone field pass over M samples (the body of NeRFNetwork.forward as one kernel)
__global__ void __launch_bounds__(128, 3) field_forward_kernel(const pn_field_t f, const float *__restrict__ xyzs,
                                                               const float *__restrict__ dirs, uint32_t M,
                                                               float *__restrict__ sigmas, float *__restrict__ rgbs) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    pn::FieldBlockSmem &fbs = *reinterpret_cast<pn::FieldBlockSmem *>(smem_raw);
    pn::FieldSmem &fs = fbs.w;
    float *scratch = fbs.scratch + threadIdx.x;
    pn::field_smem_fill(fs, f);
    __syncthreads();
    const float2 *table = reinterpret_cast<const float2 *>(f.embeddings);
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < M; i += gridDim.x * blockDim.x) {
        float sh[16];
        pn::sh_eval<4>(dirs[3 * i], dirs[3 * i + 1], dirs[3 * i + 2], sh);
        float sigma, r, g, b;
        pn::field_eval(fs, table, f.bound, xyzs[3 * i], xyzs[3 * i + 1], xyzs[3 * i + 2], sh, scratch, pn::kFieldThreads, sigma, r, g, b);
        sigmas[i] = sigma;
        rgbs[3 * i] = r; rgbs[3 * i + 1] = g; rgbs[3 * i + 2] = b;
    }
}

constexpr size_t kWeightsImageBytes = 48 * 1024;   // >= sizeof(pn::tc::Weights), checked in render_wave.cuh
struct WorkspaceLayout {
    size_t geom, queue, nears, fars, active, pig_cnt, pig_bgn, pig_fill, pig_idx, ip_pos, ip_rec, nb_cnt, nb_start, nb_fill, nb_list;
    size_t ctl, alive0, alive1, rs_march, rs_comp, link, xyzdt, meta, out, slab_next, counters, weights_img;   // wavefront mode
    int cap;
    size_t total;
};

// rows of the per-pass sample list: 24 per ray (a chair frame keeps ~11 per ray over all passes), whole slabs
int g_wave_cap_override = 0;   // pn_set_wave_capacity: tests shrink the list to exercise the "list full -> resume next pass" path
int wave_capacity(uint32_t N) {
    if (g_wave_cap_override > 0) return g_wave_cap_override / 256 * 256;
    long long c = 24ll * (long long)N;
    if (c < (1ll << 20)) c = 1ll << 20;
    if (c > (32ll << 20)) c = 32ll << 20;
    return (int)(c / 256 * 256);   // a multiple of every slab size and of the 128-row field tile
}

WorkspaceLayout layout(uint32_t N, int n_vtx, int max_cells) {
    WorkspaceLayout w;
    size_t o = 0;
    auto take = [&](size_t bytes) { size_t r = o; o += (bytes + 255) & ~size_t(255); return r; };
    w.geom = take(sizeof(FrameGeom));
    w.queue = take(sizeof(FrameQueue));
    w.nears = take(sizeof(float) * N);
    w.fars = take(sizeof(float) * N);
    w.active = take(sizeof(int) * N);
    w.pig_cnt = take(sizeof(int) * (size_t)max_cells);
    w.pig_bgn = take(sizeof(int) * ((size_t)max_cells + 1));
    w.pig_fill = take(sizeof(int) * (size_t)max_cells);
    w.pig_idx = take(sizeof(int) * (size_t)n_vtx);
    w.ip_pos = take(sizeof(float4) * (size_t)n_vtx);
    w.ip_rec = take(sizeof(float) * 16 * (size_t)n_vtx);
    w.nb_cnt = take(sizeof(int) * (size_t)max_cells); w.nb_start = take(sizeof(int) * ((size_t)max_cells + 1));
    w.nb_fill = take(sizeof(int) * (size_t)max_cells); w.nb_list = take(sizeof(float4) * 27 * (size_t)n_vtx);
    w.cap = wave_capacity(N);
    w.ctl = take(16 * 16);
    w.counters = take(64);
    w.weights_img = take(kWeightsImageBytes);
    w.alive0 = take(sizeof(int) * N); w.alive1 = take(sizeof(int) * N);
    w.rs_march = take(16 * (size_t)N); w.rs_comp = take(32 * (size_t)N); w.link = take(8 * (size_t)N);
    w.xyzdt = take(16 * (size_t)w.cap); w.meta = take(8 * (size_t)w.cap); w.out = take(16 * (size_t)w.cap);
    w.slab_next = take(sizeof(int) * (size_t)(w.cap / 32));
    w.total = o;
    return w;
}

// IP-grid capacity of the workspace: a body anywhere inside the scene box [-bound, bound]^3 fits (res_max cells per axis)
int res_max_for(float bound, float hgs) { return (int)ceilf((2 * bound + 2e-3f) / hgs) + 1; }
int max_cells_for(float bound, float hgs) {
    const int r = res_max_for(bound, hgs);
    return r * r * r;
}

template <typename K>
int set_smem(K kernel, size_t bytes) {
    if (bytes > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
        if (e != cudaSuccess) { pn_set_error("cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return PN_ECUDA; }
    }
    return PN_OK;
}

}  // namespace

#include "render_wave.cuh"

namespace {

// block-wide exclusive scan of v[0..n) into out[0..n) (+ total in out[n] when `with_total`), 1024 threads, any n
__device__ void block_scan_exclusive(const int *v, int n, int *out, int *zero_fill, bool with_total) {
    __shared__ int warp_tot[32];
    __shared__ int carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < n; base += blockDim.x) {
        const int i = base + threadIdx.x;
        const int x = i < n ? v[i] : 0;
        int incl = x;
        for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, o); if ((threadIdx.x & 31) >= o) incl += t; }
        if ((threadIdx.x & 31) == 31) warp_tot[threadIdx.x >> 5] = incl;
        __syncthreads();
        if (threadIdx.x < 32) {
            int w = warp_tot[threadIdx.x];
            for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, w, o); if (threadIdx.x >= o) w += t; }
            warp_tot[threadIdx.x] = w;
        }
        __syncthreads();
        const int before = carry + ((threadIdx.x >> 5) ? warp_tot[(threadIdx.x >> 5) - 1] : 0) + incl - x;
        if (i < n) { out[i] = before; if (zero_fill) zero_fill[i] = 0; }
        __syncthreads();
        if (threadIdx.x == blockDim.x - 1) carry = before + x;
        __syncthreads();
    }
    if (with_total && threadIdx.x == 0) out[n] = carry;
    __syncthreads();
}

// The whole per-frame IP preparation of the wavefront renderer as ONE single-CTA kernel: bbox + resolution (renderer.py:782-791),
// the deterministic counting sort of get_pnts_in_grids (nerf/utils.py:355-443), the cell-sorted IP records and the per-cell
// neighbourhood lists — ten dependent launches of a few microseconds each otherwise, which every rank of a multi-GPU frame repeats.
// Same arithmetic and the same results as the stand-alone kernels (ip_bbox / ip_grid_* / ip_pack / nb_*), phases separated by
// __syncthreads; the work is tiny (a few thousand IPs, ~10^3..10^4 grid cells), so one SM is enough and 147 stay with the renderer.
__global__ void __launch_bounds__(1024) ip_prep_fused_kernel(const float *__restrict__ p_def, const float *__restrict__ p_ori, const float *__restrict__ F,
                                                             int n, float hgs, int cut, float bound, int res_max, int n_grid_cap, int zyx_order,
                                                             FrameGeom *__restrict__ g, int *cnt, int *bgn, int *fill, int *idx, float4 *ip_pos,
                                                             float *ip_rec, int *nb_cnt, int *nb_start, float4 *nb_list) {
    __shared__ float smin[3][32], smax[3][32];
    __shared__ int s_over[3], s_res[3];
    __shared__ float s_bbmin[3];
    const int tid = threadIdx.x, nt = blockDim.x;
    // ---- bbox (ip_bbox_kernel)
    float lo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, hi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    for (int i = tid; i < n; i += nt) {
#pragma unroll
        for (int c = 0; c < 3; c++) { const float v = p_def[3 * i + c]; lo[c] = fminf(lo[c], v); hi[c] = fmaxf(hi[c], v); }
    }
#pragma unroll
    for (int c = 0; c < 3; c++) {
        for (int o = 16; o > 0; o >>= 1) {
            lo[c] = fminf(lo[c], __shfl_xor_sync(0xffffffffu, lo[c], o));
            hi[c] = fmaxf(hi[c], __shfl_xor_sync(0xffffffffu, hi[c], o));
        }
        if ((tid & 31) == 0) { smin[c][tid >> 5] = lo[c]; smax[c][tid >> 5] = hi[c]; }
    }
    __syncthreads();
    if (tid < 3) {
        const int c = tid;
        float a = FLT_MAX, b = -FLT_MAX;
        for (int w = 0; w < (nt >> 5); w++) { a = fminf(a, smin[c][w]); b = fmaxf(b, smax[c][w]); }
        if (cut) { a = -bound; b = bound; }
        const float mn = a - 1e-3f, mx = b + 1e-3f;
        const float rf = ceilf((mx - mn) * (1.0f / hgs));
        int r = (int)rf;
        s_over[c] = 0;
        if (!(rf >= 1.0f && rf <= (float)res_max)) { r = rf >= 1.0f ? res_max : 1; s_over[c] = 1; }
        g->bbmin[c] = mn; g->bbmax[c] = mx; g->hi[c] = (float)((double)mx - 1e-6); g->res[c] = r;
        s_res[c] = r; s_bbmin[c] = mn;
    }
    __syncthreads();
    const int r0 = s_res[0], r1 = s_res[1], r2 = s_res[2];
    const int n_grid = min(r0 * r1 * r2, n_grid_cap);
    if (tid == 0) { g->n_grid = r0 * r1 * r2; g->overflow = s_over[0] | s_over[1] | s_over[2]; }
    // ---- counting sort of the IPs into the hgs grid (ip_grid_count / scan / fill / sort)
    for (int c = tid; c < n_grid; c += nt) cnt[c] = 0;
    __syncthreads();
    for (int i = tid; i < n; i += nt) {
        const int gid = ip_cell(p_def, i, s_bbmin, hgs, s_res);
        if (gid >= 0 && gid < n_grid_cap) atomicAdd(cnt + gid, 1);
    }
    __syncthreads();
    block_scan_exclusive(cnt, n_grid, bgn, fill, false);
    for (int i = tid; i < n; i += nt) {
        const int gid = ip_cell(p_def, i, s_bbmin, hgs, s_res);
        if (gid >= 0 && gid < n_grid_cap) idx[bgn[gid] + atomicAdd(fill + gid, 1)] = i;
    }
    __syncthreads();
    for (int c = tid; c < n_grid; c += nt) {                  // ascending IP index inside a cell (deterministic)
        const int m = cnt[c];
        int *a = idx + bgn[c];
        for (int i = 1; i < m; i++) {
            const int v = a[i];
            int j = i - 1;
            while (j >= 0 && a[j] > v) { a[j + 1] = a[j]; j--; }
            a[j + 1] = v;
        }
    }
    __syncthreads();
    // ---- cell-sorted IP records with F^-1 (ip_pack_kernel)
    if (tid == 0) bgn[n_grid] = n;
    for (int k = tid; k < n; k += nt) {
        const int ip = idx[k];
        const float px = p_def[3 * ip], py = p_def[3 * ip + 1], pz = p_def[3 * ip + 2];
        ip_pos[k] = make_float4(px, py, pz, __int_as_float(ip));
        float A[9], Ai[9];
#pragma unroll
        for (int i = 0; i < 9; i++) A[i] = F[9 * ip + i] + 0.0f;
        pn::adjugate_inverse(A, Ai);
        float *r = ip_rec + 16 * (size_t)k;
        r[0] = p_ori[3 * ip]; r[1] = p_ori[3 * ip + 1]; r[2] = p_ori[3 * ip + 2];
        r[3] = px; r[4] = py; r[5] = pz;
#pragma unroll
        for (int i = 0; i < 9; i++) r[6 + i] = Ai[i];
        r[15] = 0.f;
    }
    // ---- per-cell neighbourhood lists (nb_count_kernel / scan / nb_fill_kernel)
    for (int c = tid; c < n_grid; c += nt) {
        const int g0 = c % r0, g1 = (c / r0) % r1, g2 = c / (r0 * r1);
        int m = 0;
        for (int dz = -1; dz <= 1; dz++)
            for (int dy = -1; dy <= 1; dy++)
                for (int dx = -1; dx <= 1; dx++) {
                    const int a0 = g0 + dx, a1 = g1 + dy, a2 = g2 + dz;
                    if (a0 >= 0 && a0 < r0 && a1 >= 0 && a1 < r1 && a2 >= 0 && a2 < r2) m += cnt[(a2 * r1 + a1) * r0 + a0];
                }
        nb_cnt[c] = m;
    }
    __syncthreads();
    block_scan_exclusive(nb_cnt, n_grid, nb_start, nullptr, true);
    for (int c = tid; c < n_grid; c += nt) {
        const int g0 = c % r0, g1 = (c / r0) % r1, g2 = c / (r0 * r1);
        int o = nb_start[c];
        if (nb_start[c + 1] == o) continue;
        for (int q = -1; q < 26; q++) {
            int d0 = 0, d1 = 0, d2 = 0;
            if (q >= 0) {
                if (zyx_order) { d2 = pn::kNeigh[q][0]; d1 = pn::kNeigh[q][1]; d0 = pn::kNeigh[q][2]; }
                else { d0 = pn::kNeigh[q][0]; d1 = pn::kNeigh[q][1]; d2 = pn::kNeigh[q][2]; }
            }
            const int a0 = g0 + d0, a1 = g1 + d1, a2 = g2 + d2;
            if (a0 < 0 || a0 >= r0 || a1 < 0 || a1 >= r1 || a2 < 0 || a2 >= r2) continue;
            const int gid = (a2 * r1 + a1) * r0 + a0;
            for (int k = bgn[gid]; k < bgn[gid + 1]; k++) {
                const float4 p = ip_pos[k];
                nb_list[o++] = make_float4(p.x, p.y, p.z, __int_as_float(k));
            }
        }
    }
}

constexpr int kFusedPrepMaxCells = 64 * 1024;   // beyond this (trex with --cut: 67^3 cells) the multi-kernel path has the parallelism

}  // namespace

int pn_field_forward_tc(const pn_field_t *f, const float *xyzs, const float *dirs, uint32_t M, float *sigmas, float *rgbs,
                        cudaStream_t st);

extern "C" int pn_get_rays(const float *pose, float fx, float fy, float cx, float cy, uint32_t H, uint32_t W,
                           float *rays_o, float *rays_d, void *stream) {
    PN_REQUIRE(pose && rays_o && rays_d, "null pointer");
    if (H * W == 0) return PN_OK;
    get_rays_kernel<<<div_up(H * W, 256u), 256, 0, PN_STREAM(stream)>>>(pose[0], pose[1], pose[2], pose[4], pose[5],
                                                                        pose[6], pose[8], pose[9], pose[10], pose[3],
                                                                        pose[7], pose[11], fx, fy, cx, cy, H, W, rays_o,
                                                                        rays_d);
    PN_LAUNCH_CHECK("get_rays_kernel");
    return PN_OK;
}

extern "C" int pn_get_rays_pix(const float *cam, uint32_t H, uint32_t W, const int *pix, uint32_t n_rays, float *rays_o,
                               float *rays_d, void *stream) {
    PN_REQUIRE(cam && rays_o && rays_d, "null pointer");
    PN_REQUIRE(pix || n_rays == H * W, "without a pixel list n_rays must be H*W");
    if (n_rays == 0) return PN_OK;
    get_rays_pix_kernel<<<div_up(n_rays, 256u), 256, 0, PN_STREAM(stream)>>>(cam, W, pix, n_rays, rays_o, rays_d);
    PN_LAUNCH_CHECK("get_rays_pix_kernel");
    return PN_OK;
}

extern "C" int pn_ip_bbox(const float *p_def, int n_vtx, float hgs, int cut, float bound, float *bbmin, float *bbmax,
                          int *resolution, void *stream) {
    PN_REQUIRE(p_def && bbmin && bbmax && resolution && n_vtx > 0, "null pointer / empty IP set");
    ip_bbox_kernel<<<1, 1024, 0, PN_STREAM(stream)>>>(p_def, n_vtx, hgs, cut, bound, 0, nullptr, bbmin, bbmax, resolution);
    PN_LAUNCH_CHECK("ip_bbox_kernel");
    return PN_OK;
}

static int build_ip_grid_impl(const float *p_def, int n_vtx, const float *bbmin, float hgs, const int *res, int n_grid_cap,
                              int *cnt, int *bgn, int *fill, int *idx, cudaStream_t st) {
    PN_CUDA(cudaMemsetAsync(cnt, 0, sizeof(int) * (size_t)n_grid_cap, st));
    ip_grid_count<<<div_up(n_vtx, 256), 256, 0, st>>>(p_def, n_vtx, bbmin, hgs, res, n_grid_cap, cnt);
    ip_grid_scan<<<1, 1024, 0, st>>>(cnt, res, n_grid_cap, bgn, fill);
    ip_grid_fill<<<div_up(n_vtx, 256), 256, 0, st>>>(p_def, n_vtx, bbmin, hgs, res, n_grid_cap, bgn, fill, idx);
    ip_grid_sort<<<div_up(n_grid_cap, 256), 256, 0, st>>>(cnt, bgn, res, n_grid_cap, idx);
    PN_LAUNCH_CHECK("ip_grid");
    return PN_OK;
}

extern "C" int pn_build_ip_grid(const float *p_def, int n_vtx, const float *bbmin, float hgs, const int *resolution,
                                int n_grid, int *pig_cnt, int *pig_bgn, int *pig_idx, void *stream) {
    PN_REQUIRE(p_def && bbmin && resolution && pig_cnt && pig_bgn && pig_idx, "null pointer");
    PN_REQUIRE(n_vtx > 0 && n_grid > 0, "empty IP set / grid");
    int *fill = nullptr;
    cudaStream_t st = PN_STREAM(stream);
    PN_CUDA(cudaMallocAsync(&fill, sizeof(int) * (size_t)n_grid, st));
    const int rc = build_ip_grid_impl(p_def, n_vtx, bbmin, hgs, resolution, n_grid, pig_cnt, pig_bgn, fill, pig_idx, st);
    cudaFreeAsync(fill, st);
    return rc;
}

extern "C" int pn_field_forward(const pn_field_t *f, const float *xyzs, const float *dirs, uint32_t M, float *sigmas,
                                float *rgbs, int mode, void *stream) {
    PN_REQUIRE(f && xyzs && dirs && sigmas && rgbs, "null pointer");
    PN_REQUIRE(f->L == pn::kLevels, "fused field expects the 16-level C=2 D=3 grid of nerf/network.py");
    PN_REQUIRE(mode == 0 || mode == 1, "field mode: 0 = fp32 SIMT, 1 = tcgen05 (bf16x3 split)");
    if (M == 0) return PN_OK;
    if (mode == 1) return pn_field_forward_tc(f, xyzs, dirs, M, sigmas, rgbs, PN_STREAM(stream));
    const size_t smem = sizeof(pn::FieldBlockSmem);
    if (int rc = set_smem(field_forward_kernel, smem)) return rc;
    const uint32_t blocks = min(div_up(M, 128u), (uint32_t)pn_sm_count_cached() * 3u);
    field_forward_kernel<<<blocks, 128, smem, PN_STREAM(stream)>>>(*f, xyzs, dirs, M, sigmas, rgbs);
    PN_LAUNCH_CHECK("field_forward_kernel");
    return PN_OK;
}

extern "C" int pn_set_wave_capacity(int rows) {
    PN_REQUIRE(rows == 0 || rows >= 1024, "rows: 0 (automatic) or >= 1024");
    g_wave_cap_override = rows;
    return PN_OK;
}
static int g_prep_force_multi = 0;   // tests: compare the fused preparation kernel with the multi-kernel chain
extern "C" int pn_set_prep_mode(int force_multi_kernel) { g_prep_force_multi = force_multi_kernel; return PN_OK; }
static int g_sm_reserve = 0;
extern "C" int pn_set_render_sm_reserve(int n_sm) {
    PN_REQUIRE(n_sm >= 0 && n_sm < pn_sm_count_cached(), "reserve must leave SMs for the renderer");
    g_sm_reserve = n_sm;
    return PN_OK;
}
static cudaEvent_t g_prof_start = nullptr, g_prof_stop = nullptr;
static cudaEvent_t *g_prof_list = nullptr;   // wavefront mode: events [2k], [2k+1] bracket the k-th field-kernel launch of a frame
static int g_prof_n = 0;
extern "C" int pn_set_profile_event_list(void **events, int n) {
    g_prof_list = (cudaEvent_t *)events;
    g_prof_n = events ? n : 0;
    return PN_OK;
}
extern "C" int pn_set_profile_events(void *start, void *stop) {
    g_prof_start = (cudaEvent_t)start;
    g_prof_stop = (cudaEvent_t)stop;
    return PN_OK;
}

extern "C" uint64_t pn_mlp_workspace_bytes(uint32_t M) {
    const uint64_t rows = ((uint64_t)M + 127) / 128 * 128;
    return 4096 + kWeightsImageBytes + rows * (8 + 16 + 16);
}

// nerf/network.py:105-127 (sigma_net + color_net) alone, on the product's tensor-core pipeline: the wavefront field kernel with
// its hash-grid gather replaced by a coalesced load of pre-encoded features.  Exists to measure what the tcgen05 / TMEM MLP
// sustains by itself (bench.py "mlp_pass"); the profiling event list of pn_set_profile_event_list brackets the field kernel.
extern "C" int pn_mlp_forward(const pn_field_t *f, const float *enc, const float *dirs, uint32_t M, float *sigmas, float *rgbs, void *workspace,
                              uint64_t workspace_bytes, void *stream) {
    PN_REQUIRE(f && enc && dirs && sigmas && rgbs && workspace, "null pointer");
    PN_REQUIRE(workspace_bytes >= pn_mlp_workspace_bytes(M), "workspace too small (see pn_mlp_workspace_bytes)");
    if (M == 0) return PN_OK;
    cudaStream_t st = PN_STREAM(stream);
    const uint64_t rows = ((uint64_t)M + 127) / 128 * 128;
    unsigned char *base = (unsigned char *)workspace;
    WaveArgs Wv{};
    Wv.ctl = (PassCtl *)base;
    pn::tc::Weights *wimg = (pn::tc::Weights *)(base + 4096);
    Wv.weights_img = wimg;
    Wv.meta = (int2 *)(base + 4096 + kWeightsImageBytes);
    Wv.xyzdt = (float4 *)((unsigned char *)Wv.meta + rows * 8);
    Wv.out = (float4 *)((unsigned char *)Wv.xyzdt + rows * 16);
    Wv.enc = enc; Wv.cap = (int)rows;
    RenderArgs A{};
    A.field = *f; A.rays_d = dirs; A.density_scale = 1.0f;
    PN_CUDA(cudaMemsetAsync(base, 0, 4096, st));
    mlp_rows_kernel<<<div_up(M, 256u), 256, 0, st>>>(M, Wv.meta, Wv.xyzdt, Wv.ctl);
    field_weights_kernel<<<1, 256, 0, st>>>(A.field, wimg);
    const size_t smem = sizeof(WaveWsSmem) + 128;
    if (int rc = set_smem(wave_field_ws_kernel, smem)) return rc;
    if (g_prof_list && g_prof_n >= 2) PN_CUDA(cudaEventRecord(g_prof_list[0], st));
    wave_field_ws_kernel<<<(uint32_t)pn_sm_count_cached(), (kWsProd + kWsCons) * 128, smem, st>>>(A, Wv, 0);
    if (g_prof_list && g_prof_n >= 2) PN_CUDA(cudaEventRecord(g_prof_list[1], st));
    mlp_unpack_kernel<<<div_up(M, 256u), 256, 0, st>>>(M, Wv.out, sigmas, rgbs);
    PN_LAUNCH_CHECK("pn_mlp_forward");
    return PN_OK;
}

#ifndef PN_WAVE_FIRST_CAP
#define PN_WAVE_FIRST_CAP 32   // samples per ray in the first pass (doubles every pass): saturating rays waste <= one 32-sample chunk
#endif
extern "C" int pn_render_pass_count(uint32_t max_steps) {
    int n_pass = 0, covered = 0;
    for (int cap_p = PN_WAVE_FIRST_CAP; covered < (int)max_steps && n_pass < kMaxPass - 1; cap_p *= 2) { covered += cap_p; n_pass++; }
    return n_pass + 1;
}

extern "C" uint64_t pn_render_workspace_bytes(uint32_t N, int n_vtx, float bound, float hgs) {
    return layout(N, n_vtx, max_cells_for(bound, hgs)).total;
}

extern "C" int pn_render_deformed(const pn_field_t *f, const pn_deform_t *d, const float *rays_o, const float *rays_d,
                                  uint32_t N, float *image, float *depth, float *depth_0, float *weights_sum,
                                  void *workspace, uint64_t workspace_bytes, long long *stats, int mode, void *stream) {
    return pn_render_deformed_ex(f, d, rays_o, rays_d, N, image, depth, depth_0, weights_sum, workspace, workspace_bytes, stats,
                                 mode, nullptr, stream);
}

int pn_flag_wait_launch(const pn_frame_io_t *io, cudaStream_t st);      // peer.cu
int pn_flag_signal_launch(const pn_frame_io_t *io, cudaStream_t st);

extern "C" int pn_render_deformed_ex(const pn_field_t *f, const pn_deform_t *d, const float *rays_o, const float *rays_d,
                                     uint32_t N, float *image, float *depth, float *depth_0, float *weights_sum,
                                     void *workspace, uint64_t workspace_bytes, long long *stats, int mode,
                                     const pn_frame_io_t *io, void *stream) {
    PN_REQUIRE(f && d && rays_o && rays_d && image && depth && depth_0 && weights_sum && workspace, "null pointer");
    PN_REQUIRE(!io || !(io->pix || io->noises) || mode == 3, "a pixel map (scattered / peer frame output) or perturbed ray starts need the wavefront renderer (mode 3)");
    PN_REQUIRE(f->L == pn::kLevels, "fused field expects the 16-level C=2 D=3 grid of nerf/network.py");
    PN_REQUIRE(mode >= 0 && mode <= 3, "render mode: 0 = fused warp-cooperative + tcgen05 MLP, 1 = same with the fp32 SIMT MLP, 2 = one lane per ray, 3 = wavefront (march / field / composite kernels)");
    PN_REQUIRE(d->n_vtx > 0 && d->num_seek_IP >= 1 && d->num_seek_IP <= 3, "need IPs and num_seek_IP in 1..3");
    PN_REQUIRE(d->cascade >= 1 && d->cascade <= 8 && d->grid_size == 128, "cascade/grid_size out of range");
    if (N == 0) return PN_OK;
    cudaStream_t st = PN_STREAM(stream);
    const int max_cells = max_cells_for(d->bound, d->hgs);
    const WorkspaceLayout w = layout(N, d->n_vtx, max_cells);
    PN_REQUIRE(workspace_bytes >= w.total, "workspace too small (see pn_render_workspace_bytes)");
    unsigned char *base = (unsigned char *)workspace;
    FrameGeom *geom = (FrameGeom *)(base + w.geom);
    FrameQueue *queue = (FrameQueue *)(base + w.queue);
    float *nears = (float *)(base + w.nears), *fars = (float *)(base + w.fars);
    int *active = (int *)(base + w.active);
    int *cnt = (int *)(base + w.pig_cnt), *bgn = (int *)(base + w.pig_bgn), *fill = (int *)(base + w.pig_fill), *idx = (int *)(base + w.pig_idx);

    if (io && io->epoch)
        if (int rc = pn_flag_wait_launch(io, st)) return rc;       // bump the slot's epoch; wait e.g. for this frame's IP state to land in this GPU's memory
    PN_CUDA(cudaMemsetAsync(queue, 0, sizeof(FrameQueue), st));
    // IP preparation: one single-CTA kernel when the grid is small (every config without --cut), else the parallel multi-kernel chain
    const bool fused_prep = mode != 2 && max_cells <= kFusedPrepMaxCells && !g_prep_force_multi;
    if (fused_prep) {
        ip_prep_fused_kernel<<<1, 1024, 0, st>>>(d->p_def, d->p_ori, d->F_IP, d->n_vtx, d->hgs, d->cut, d->bound, res_max_for(d->bound, d->hgs), max_cells,
                                                 d->num_seek_IP == 1, geom, cnt, bgn, fill, idx, (float4 *)(base + w.ip_pos), (float *)(base + w.ip_rec),
                                                 (int *)(base + w.nb_cnt), (int *)(base + w.nb_start), (float4 *)(base + w.nb_list));
        PN_LAUNCH_CHECK("ip_prep_fused_kernel");
    } else {
        ip_bbox_kernel<<<1, 1024, 0, st>>>(d->p_def, d->n_vtx, d->hgs, d->cut, d->bound, res_max_for(d->bound, d->hgs), geom, nullptr, nullptr, nullptr);
        if (int rc = build_ip_grid_impl(d->p_def, d->n_vtx, geom->bbmin, d->hgs, geom->res, max_cells, cnt, bgn, fill, idx, st)) return rc;
    }
    frame_setup_kernel<<<div_up(N, 256u), 256, 0, st>>>(rays_o, rays_d, N, geom, d->min_near, d->bg_color, nears, fars,
                                                         active, queue, io ? io->pix : nullptr, image, depth, depth_0, weights_sum);
    PN_LAUNCH_CHECK("frame_setup_kernel");

    RenderArgs A{};
    A.field = *f;
    A.march.bound = d->bound; A.march.dt_gamma = d->dt_gamma;
    A.march.dt_min = 2 * 1.7320508075688772f / d->max_steps;
    A.march.dt_max = 2 * 1.7320508075688772f * (1 << (d->cascade - 1)) / d->grid_size;
    A.march.cascade = (int)d->cascade; A.march.H = (int)d->grid_size; A.march.bits = d->density_bitfield;
    A.bend.pig_cnt = cnt; A.bend.pig_bgn = bgn; A.bend.pig_idx = idx;
    A.bend.p_ori = d->p_ori; A.bend.p_def = d->p_def; A.bend.F = d->F_IP; A.bend.dF = d->dF_IP;
    A.bend.n_grid = max_cells; A.bend.max_iter = d->max_iter_num; A.bend.K = d->num_seek_IP;
    A.bend.hgs = d->hgs; A.bend.IP_dx = d->IP_dx; A.bend.bound = d->bound; A.bend.cut = d->cut != 0;
    for (int i = 0; i < 6; i++) A.bend.cb[i] = d->cut_bounds[i];
    A.geom = geom; A.rays_o = rays_o; A.rays_d = rays_d; A.nears = nears; A.fars = fars; A.active = active; A.queue = queue;
    A.image = image; A.depth = depth; A.depth0 = depth_0; A.wsum = weights_sum; A.pix = io ? io->pix : nullptr; A.noises = io ? io->noises : nullptr;
    A.density_scale = d->density_scale; A.T_thresh = d->T_thresh; A.bg = d->bg_color; A.max_samples = d->max_steps;

    const uint32_t blocks = (uint32_t)pn_sm_count_cached() * 3u;
    if (mode == 3) {
        float4 *ip_pos = (float4 *)(base + w.ip_pos);
        float *ip_rec = (float *)(base + w.ip_rec);
        int *nb_cnt = (int *)(base + w.nb_cnt), *nb_start = (int *)(base + w.nb_start), *nb_fill = (int *)(base + w.nb_fill);
        float4 *nb_list = (float4 *)(base + w.nb_list);
        if (!fused_prep) {
            ip_pack_kernel<<<div_up(d->n_vtx, 256), 256, 0, st>>>(d->p_def, d->p_ori, d->F_IP, idx, d->n_vtx, geom->res, max_cells, bgn, ip_pos, ip_rec);
            nb_count_kernel<<<div_up(max_cells, 256), 256, 0, st>>>(cnt, geom->res, max_cells, nb_cnt);
            ip_grid_scan<<<1, 1024, 0, st>>>(nb_cnt, geom->res, max_cells, nb_start, nb_fill);
            nb_fill_kernel<<<div_up(max_cells, 256), 256, 0, st>>>(bgn, ip_pos, geom->res, max_cells, d->num_seek_IP == 1, nb_start, nb_list);
            PN_LAUNCH_CHECK("ip_pack / neighbourhood lists");
        }
        IpPack P{ip_pos, ip_rec, bgn, nb_start, nb_list};
        WaveArgs Wv{};
        Wv.ctl = (PassCtl *)(base + w.ctl); Wv.counters = (long long *)(base + w.counters);
        Wv.alive[0] = (int *)(base + w.alive0); Wv.alive[1] = (int *)(base + w.alive1);
        Wv.rs_march = (float4 *)(base + w.rs_march); Wv.rs_comp = (float4 *)(base + w.rs_comp); Wv.link = (int2 *)(base + w.link);
        Wv.xyzdt = (float4 *)(base + w.xyzdt); Wv.meta = (int2 *)(base + w.meta); Wv.out = (float4 *)(base + w.out);
        Wv.slab_next = (int *)(base + w.slab_next); Wv.cap = w.cap;
        Wv.weights_img = (const pn::tc::Weights *)(base + w.weights_img);
        // the bf16 hi / lo weight image only changes with the weights: a caller that renders frame after frame builds it once
        if (!(io && (io->flags & PN_IO_WEIGHTS_READY))) field_weights_kernel<<<1, 256, 0, st>>>(A.field, (pn::tc::Weights *)(base + w.weights_img));
        PN_CUDA(cudaMemsetAsync(base + w.ctl, 0, 16 * 16 + 256, st));       // ctl + counters (adjacent, 256-byte aligned blocks)
        const size_t smem = sizeof(WaveWsSmem) + 128;
        if (int rc = set_smem(wave_field_ws_kernel, smem)) return rc;
        // SMs the persistent march / field grids are sized for: all of them, minus the ones the caller keeps free for kernels of
        // another stream (pn_set_render_sm_reserve: on the GPU that also runs the simulator, its many tiny launches otherwise queue
        // behind render CTAs that hold every SM until their kernel ends)
        const uint32_t sms = (uint32_t)max(8, pn_sm_count_cached() - g_sm_reserve);
        // pass caps 32, 64, ... until the per-ray cap is covered; one spare pass absorbs the <32-sample overshoot per pass.
        // io->max_passes < that: fewer launches, the LAST pass marches every remaining sample (cap = max_steps) — rays are then
        // cut short only if the sample list itself overflows, which stats[4..5] report.
        int n_pass = pn_render_pass_count(d->max_steps);
        const bool limited = io && io->max_passes > 0 && io->max_passes < n_pass;
        if (limited) n_pass = io->max_passes;
        int cap_p = PN_WAVE_FIRST_CAP, fk = 0;
        if (g_prof_start) PN_CUDA(cudaEventRecord(g_prof_start, st));
        for (int p = 0; p < n_pass; p++, cap_p *= 2) {
            const int cap_now = (limited && p == n_pass - 1) ? (int)d->max_steps : cap_p;
            switch (d->num_seek_IP) {
                case 1: wave_march_kernel<1><<<sms * PN_MARCH_MINB, PN_MARCH_THREADS, 0, st>>>(A, P, Wv, p, cap_now); break;
                case 2: wave_march_kernel<2><<<sms * PN_MARCH_MINB, PN_MARCH_THREADS, 0, st>>>(A, P, Wv, p, cap_now); break;
                default: wave_march_kernel<3><<<sms * PN_MARCH_MINB, PN_MARCH_THREADS, 0, st>>>(A, P, Wv, p, cap_now); break;
            }
            if (g_prof_list && 2 * fk + 1 < g_prof_n) PN_CUDA(cudaEventRecord(g_prof_list[2 * fk], st));
            wave_field_ws_kernel<<<sms, (kWsProd + kWsCons) * 128, smem, st>>>(A, Wv, p);
            if (g_prof_list && 2 * fk + 1 < g_prof_n) PN_CUDA(cudaEventRecord(g_prof_list[2 * fk + 1], st));
            fk++;
            wave_composite_kernel<<<min(div_up(N, 256u), sms * 8u), 256, 0, st>>>(A, Wv, p, p == n_pass - 1);
        }
        PN_LAUNCH_CHECK("wavefront passes");
        if (g_prof_stop) PN_CUDA(cudaEventRecord(g_prof_stop, st));
        if (stats) {
            wave_stats_kernel<<<1, 1, 0, st>>>(queue, geom, Wv, n_pass, stats);
            PN_LAUNCH_CHECK("wave_stats_kernel");
        }
        if (io && io->epoch && io->n_signal > 0)
            if (int rc = pn_flag_signal_launch(io, st)) return rc;   // every pixel of this rank is in the frame: tell its owner
        return PN_OK;
    }
    if (mode == 0 || mode == 1) {
        float4 *ip_pos = (float4 *)(base + w.ip_pos);
        float *ip_rec = (float *)(base + w.ip_rec);
        int *nb_cnt = (int *)(base + w.nb_cnt), *nb_start = (int *)(base + w.nb_start), *nb_fill = (int *)(base + w.nb_fill);
        float4 *nb_list = (float4 *)(base + w.nb_list);
        if (!fused_prep) {
            ip_pack_kernel<<<div_up(d->n_vtx, 256), 256, 0, st>>>(d->p_def, d->p_ori, d->F_IP, idx, d->n_vtx, geom->res, max_cells, bgn, ip_pos, ip_rec);
            nb_count_kernel<<<div_up(max_cells, 256), 256, 0, st>>>(cnt, geom->res, max_cells, nb_cnt);
            ip_grid_scan<<<1, 1024, 0, st>>>(nb_cnt, geom->res, max_cells, nb_start, nb_fill);
            nb_fill_kernel<<<div_up(max_cells, 256), 256, 0, st>>>(bgn, ip_pos, geom->res, max_cells, d->num_seek_IP == 1, nb_start, nb_list);
            PN_LAUNCH_CHECK("ip_pack / neighbourhood lists");
        }
        IpPack P{ip_pos, ip_rec, bgn, nb_start, nb_list};
        const bool tc = mode == 0;
        const size_t smem = tc ? (((sizeof(RenderTcSmem) + 127) & ~size_t(127)) + kTcGroups * 4 * sizeof(WarpShared) + 128)
                               : (((sizeof(pn::FieldBlockSmem) + 127) & ~size_t(127)) + 4 * sizeof(WarpShared) + 128);
        const uint32_t nblk = tc ? (uint32_t)pn_sm_count_cached() : blocks;
        const uint32_t nthr = tc ? kTcGroups * 128 : 128;
        if (g_prof_start) PN_CUDA(cudaEventRecord(g_prof_start, st));
#define PN_LAUNCH_WARP(K)                                                                       \
        if (tc) {                                                                               \
            if (int rc = set_smem(render_warp_kernel<K, true>, smem)) return rc;                \
            render_warp_kernel<K, true><<<nblk, nthr, smem, st>>>(A, P);                        \
        } else {                                                                                \
            if (int rc = set_smem(render_warp_kernel<K, false>, smem)) return rc;               \
            render_warp_kernel<K, false><<<nblk, nthr, smem, st>>>(A, P);                       \
        }
        switch (d->num_seek_IP) {
            case 1: PN_LAUNCH_WARP(1) break;
            case 2: PN_LAUNCH_WARP(2) break;
            default: PN_LAUNCH_WARP(3) break;
        }
#undef PN_LAUNCH_WARP
    } else {
    const size_t smem = sizeof(pn::FieldBlockSmem);
    if (g_prof_start) PN_CUDA(cudaEventRecord(g_prof_start, st));
    switch (d->num_seek_IP) {
        case 1:
            if (int rc = set_smem(render_persistent<1>, smem)) return rc;
            render_persistent<1><<<blocks, 128, smem, st>>>(A);
            break;
        case 2:
            if (int rc = set_smem(render_persistent<2>, smem)) return rc;
            render_persistent<2><<<blocks, 128, smem, st>>>(A);
            break;
        default:
            if (int rc = set_smem(render_persistent<3>, smem)) return rc;
            render_persistent<3><<<blocks, 128, smem, st>>>(A);
            break;
    }
    }
    PN_LAUNCH_CHECK("render_persistent");
    if (g_prof_stop) PN_CUDA(cudaEventRecord(g_prof_stop, st));
    if (stats) {
        copy_stats_kernel<<<1, 1, 0, st>>>(queue, geom, stats);
        PN_LAUNCH_CHECK("copy_stats_kernel");
    }
    if (io && io->epoch && io->n_signal > 0)
        if (int rc = pn_flag_signal_launch(io, st)) return rc;
    return PN_OK;
}
