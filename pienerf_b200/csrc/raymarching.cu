// Drop-in ray-marching entry points (inference side) for sm_100a.
// Replaces raymarching/src/raymarching.cu:92-303,704-923,1122-1489 of the reference, one kernel per entry
// point, all launched on the caller's stream with no host synchronisation.  Output layouts are the
// reference's ([M,3] xyzs/dirs, [M,2] deltas, caller-zeroed) so the reference's Python wrappers run unchanged.
#include "march_device.cuh"

namespace {

constexpr int kRayBlock = 128;

__global__ void __launch_bounds__(256) near_far_kernel(const float *__restrict__ rays_o, const float *__restrict__ rays_d,
                                                       const float *__restrict__ aabb, uint32_t N, float min_near,
                                                       float *__restrict__ nears, float *__restrict__ fars) {
    const uint32_t n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    const float ox = rays_o[3 * n], oy = rays_o[3 * n + 1], oz = rays_o[3 * n + 2];
    const float rdx = 1 / rays_d[3 * n], rdy = 1 / rays_d[3 * n + 1], rdz = 1 / rays_d[3 * n + 2];
    // slab test, x then y then z, early-out on a miss exactly where the reference does (raymarching.cu:113-139)
    float near = (aabb[0] - ox) * rdx, far = (aabb[3] - ox) * rdx;
    if (near > far) { const float s = near; near = far; far = s; }
    float ny = (aabb[1] - oy) * rdy, fy = (aabb[4] - oy) * rdy;
    if (ny > fy) { const float s = ny; ny = fy; fy = s; }
    bool miss = near > fy || ny > far;
    if (!miss) {
        if (ny > near) near = ny;
        if (fy < far) far = fy;
        float nz = (aabb[2] - oz) * rdz, fz = (aabb[5] - oz) * rdz;
        if (nz > fz) { const float s = nz; nz = fz; fz = s; }
        miss = near > fz || nz > far;
        if (!miss) {
            if (nz > near) near = nz;
            if (fz < far) far = fz;
            if (near < min_near) near = min_near;
        }
    }
    nears[n] = miss ? FLT_MAX : near;
    fars[n] = miss ? FLT_MAX : far;
}

__global__ void __launch_bounds__(256) sph_from_ray_kernel(const float *__restrict__ rays_o, const float *__restrict__ rays_d,
                                                           float radius, uint32_t N, float *__restrict__ coords) {
    const uint32_t n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    const float ox = rays_o[3 * n], oy = rays_o[3 * n + 1], oz = rays_o[3 * n + 2];
    const float dx = rays_d[3 * n], dy = rays_d[3 * n + 1], dz = rays_d[3 * n + 2];
    // far intersection with the sphere |o + t d| = radius (raymarching.cu:186-200)
    const float A = dx * dx + dy * dy + dz * dz;
    const float Bh = ox * dx + oy * dy + oz * dz;
    const float Cc = ox * ox + oy * oy + oz * oz - radius * radius;
    const float t = (-Bh + sqrtf(Bh * Bh - A * Cc)) / A;
    const float x = ox + t * dx, y = oy + t * dy, z = oz + t * dz;
    const float theta = atan2f(sqrtf(x * x + z * z), y);
    const float phi = atan2f(z, x);
    coords[2 * n] = 2 * theta * 0.3183098861837907f - 1;
    coords[2 * n + 1] = phi * 0.3183098861837907f;
}

__global__ void __launch_bounds__(256) morton_kernel(const int *__restrict__ coords, uint32_t N, int *__restrict__ out) {
    const uint32_t n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n < N) out[n] = (int)pn::morton3(coords[3 * n], coords[3 * n + 1], coords[3 * n + 2]);
}

__global__ void __launch_bounds__(256) morton_invert_kernel(const int *__restrict__ idx, uint32_t N, int *__restrict__ out) {
    const uint32_t n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    const int v = idx[n];  // arithmetic shifts of the signed index, as the reference does
    out[3 * n] = (int)pn::compact3(v >> 0);
    out[3 * n + 1] = (int)pn::compact3(v >> 1);
    out[3 * n + 2] = (int)pn::compact3(v >> 2);
}

// one thread packs 8 floats (two 128-bit loads) into one byte
__global__ void __launch_bounds__(256) packbits_kernel(const float4 *__restrict__ grid, uint32_t N, float thresh,
                                                       uint8_t *__restrict__ bits) {
    const uint32_t n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    const float4 a = __ldg(grid + 2 * n), b = __ldg(grid + 2 * n + 1);
    uint32_t v = 0;
    v |= (a.x > thresh) << 0; v |= (a.y > thresh) << 1; v |= (a.z > thresh) << 2; v |= (a.w > thresh) << 3;
    v |= (b.x > thresh) << 4; v |= (b.y > thresh) << 5; v |= (b.z > thresh) << 6; v |= (b.w > thresh) << 7;
    bits[n] = (uint8_t)v;
}

// ---- inference march (plain and bending share the loop) -----------------------------------------------
struct MarchIO {
    uint32_t n_alive, n_step;
    const int *rays_alive;
    const float *rays_t, *rays_o, *rays_d, *fars, *noises;
    float *xyzs, *dirs, *deltas;
};

struct BendPtrs {  // device-resident small arrays of the reference API
    const float *bbmin, *bbmax, *cut_bounds;
    const int *resolution;
};

template <int KMAX>  // KMAX == 0: no bending (march_rays)
__global__ void __launch_bounds__(kRayBlock) march_kernel(MarchIO io, pn::MarchCfg m, pn::BendCfg bc, BendPtrs bp) {
    const uint32_t n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= io.n_alive) return;
    if constexpr (KMAX > 0) {
#pragma unroll
        for (int i = 0; i < 3; i++) {
            bc.bbmin[i] = bp.bbmin[i];
            bc.bbmax[i] = bp.bbmax[i];
            bc.hi[i] = (float)((double)bp.bbmax[i] - 1e-6);
            bc.res[i] = bp.resolution[i];
        }
#pragma unroll
        for (int i = 0; i < 6; i++) bc.cb[i] = bc.cut ? bp.cut_bounds[i] : 0.f;
    }
    const int ray = io.rays_alive[n];
    const float noise = io.noises[n];
    const float ox = io.rays_o[3 * ray], oy = io.rays_o[3 * ray + 1], oz = io.rays_o[3 * ray + 2];
    const float dx = io.rays_d[3 * ray], dy = io.rays_d[3 * ray + 1], dz = io.rays_d[3 * ray + 2];
    const float rdx = 1 / dx, rdy = 1 / dy, rdz = 1 / dz;
    float t = io.rays_t[ray];
    const float far = io.fars[ray];
    float *xyz = io.xyzs + (size_t)n * io.n_step * 3;
    float *dir = io.dirs + (size_t)n * io.n_step * 3;
    float *del = io.deltas + (size_t)n * io.n_step * 2;

    t += pn::step_size(m, t) * noise;
    float last_t = t;
    uint32_t step = 0;
    while (t < far && step < io.n_step) {
        float x, y, z;
        bool found = true;
        if constexpr (KMAX > 0) {
            pn::deformed_sample(bc, ox, oy, oz, dx, dy, dz, t, x, y, z);
            found = pn::bend_sample<KMAX>(bc, x, y, z);
        } else {
            x = pn::clampf(ox + t * dx, -m.bound, m.bound);
            y = pn::clampf(oy + t * dy, -m.bound, m.bound);
            z = pn::clampf(oz + t * dz, -m.bound, m.bound);
        }
        const float dt = pn::step_size(m, t);
        float tt;
        const bool occ = pn::occupancy_and_exit(m, x, y, z, t, dt, dx, dy, dz, rdx, rdy, rdz, tt);
        if (occ && found) {
            xyz[0] = x; xyz[1] = y; xyz[2] = z;
            dir[0] = dx; dir[1] = dy; dir[2] = dz;
            t += dt;
            del[0] = dt;
            del[1] = t - last_t;
            last_t = t;
            xyz += 3; dir += 3; del += 2;
            step++;
        } else {
            do { t += pn::step_size(m, t); } while (t < tt);
        }
    }
}

__global__ void __launch_bounds__(kRayBlock) composite_kernel(uint32_t n_alive, uint32_t n_step, float T_thresh,
                                                              int *__restrict__ rays_alive, float *__restrict__ rays_t,
                                                              const float *__restrict__ sigmas,
                                                              const float *__restrict__ rgbs,
                                                              const float *__restrict__ deltas,
                                                              float *__restrict__ weights_sum, float *__restrict__ depth,
                                                              float *__restrict__ image) {
    const uint32_t n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= n_alive) return;
    const int ray = rays_alive[n];
    const float *sg = sigmas + (size_t)n * n_step;
    const float *col = rgbs + (size_t)n * n_step * 3;
    const float *del = deltas + (size_t)n * n_step * 2;
    float t = rays_t[ray], ws = weights_sum[ray], d = depth[ray];
    float r = image[3 * ray], g = image[3 * ray + 1], b = image[3 * ray + 2];
    uint32_t step = 0;
    for (; step < n_step; step++) {
        const float d0 = del[2 * step];
        if (d0 == 0) break;  // no sample: the ray left the volume (raymarching.cu:866-867)
        const float alpha = 1.0f - __expf(-sg[step] * d0);
        const float T = 1 - ws;
        const float w = alpha * T;
        ws += w;
        t += del[2 * step + 1];
        d += w * t;
        r += w * col[3 * step]; g += w * col[3 * step + 1]; b += w * col[3 * step + 2];
        if (T < T_thresh) break;  // saturated: accumulate this sample, then stop
    }
    if (step < n_step) rays_alive[n] = -1; else rays_t[ray] = t;
    weights_sum[ray] = ws; depth[ray] = d;
    image[3 * ray] = r; image[3 * ray + 1] = g; image[3 * ray + 2] = b;
}

}  // namespace

extern "C" int pn_near_far_from_aabb(const float *rays_o, const float *rays_d, const float *aabb, uint32_t N,
                                     float min_near, float *nears, float *fars, void *stream) {
    PN_REQUIRE(rays_o && rays_d && aabb && nears && fars, "null pointer");
    if (N == 0) return PN_OK;
    near_far_kernel<<<div_up(N, 256u), 256, 0, PN_STREAM(stream)>>>(rays_o, rays_d, aabb, N, min_near, nears, fars);
    PN_LAUNCH_CHECK("near_far_kernel");
    return PN_OK;
}

extern "C" int pn_sph_from_ray(const float *rays_o, const float *rays_d, float radius, uint32_t N, float *coords,
                               void *stream) {
    PN_REQUIRE(rays_o && rays_d && coords, "null pointer");
    if (N == 0) return PN_OK;
    sph_from_ray_kernel<<<div_up(N, 256u), 256, 0, PN_STREAM(stream)>>>(rays_o, rays_d, radius, N, coords);
    PN_LAUNCH_CHECK("sph_from_ray_kernel");
    return PN_OK;
}

extern "C" int pn_morton3D(const int *coords, uint32_t N, int *indices, void *stream) {
    PN_REQUIRE(coords && indices, "null pointer");
    if (N == 0) return PN_OK;
    morton_kernel<<<div_up(N, 256u), 256, 0, PN_STREAM(stream)>>>(coords, N, indices);
    PN_LAUNCH_CHECK("morton_kernel");
    return PN_OK;
}

extern "C" int pn_morton3D_invert(const int *indices, uint32_t N, int *coords, void *stream) {
    PN_REQUIRE(coords && indices, "null pointer");
    if (N == 0) return PN_OK;
    morton_invert_kernel<<<div_up(N, 256u), 256, 0, PN_STREAM(stream)>>>(indices, N, coords);
    PN_LAUNCH_CHECK("morton_invert_kernel");
    return PN_OK;
}

extern "C" int pn_packbits(const float *grid, uint32_t N, float density_thresh, uint8_t *bitfield, void *stream) {
    PN_REQUIRE(grid && bitfield, "null pointer");
    PN_REQUIRE((reinterpret_cast<uintptr_t>(grid) & 15) == 0, "grid must be 16-byte aligned");
    if (N == 0) return PN_OK;
    packbits_kernel<<<div_up(N, 256u), 256, 0, PN_STREAM(stream)>>>((const float4 *)grid, N, density_thresh, bitfield);
    PN_LAUNCH_CHECK("packbits_kernel");
    return PN_OK;
}

extern "C" int pn_march_rays(uint32_t n_alive, uint32_t n_step, const int *rays_alive, const float *rays_t,
                             const float *rays_o, const float *rays_d, float bound, float dt_gamma, uint32_t max_steps,
                             uint32_t C, uint32_t H, const uint8_t *grid, const float *nears, const float *fars,
                             float *xyzs, float *dirs, float *deltas, const float *noises, void *stream) {
    (void)nears;
    PN_REQUIRE(rays_alive && rays_t && rays_o && rays_d && grid && fars && xyzs && dirs && deltas && noises, "null pointer");
    PN_REQUIRE(C >= 1 && C <= 8 && H >= 1 && H <= 256 && max_steps > 0, "bad C/H/max_steps");
    if (n_alive == 0 || n_step == 0) return PN_OK;
    MarchIO io{n_alive, n_step, rays_alive, rays_t, rays_o, rays_d, fars, noises, xyzs, dirs, deltas};
    pn::MarchCfg m;
    m.bound = bound; m.dt_gamma = dt_gamma;
    m.dt_min = 2 * 1.7320508075688772f / max_steps;
    m.dt_max = 2 * 1.7320508075688772f * (1 << (C - 1)) / H;
    m.cascade = (int)C; m.H = (int)H; m.bits = grid;
    pn::BendCfg bc{};
    BendPtrs bp{};
    march_kernel<0><<<div_up(n_alive, (uint32_t)kRayBlock), kRayBlock, 0, PN_STREAM(stream)>>>(io, m, bc, bp);
    PN_LAUNCH_CHECK("march_kernel<0>");
    return PN_OK;
}

extern "C" int pn_composite_rays(uint32_t n_alive, uint32_t n_step, float T_thresh, int *rays_alive, float *rays_t,
                                 const float *sigmas, const float *rgbs, const float *deltas, float *weights_sum,
                                 float *depth, float *image, void *stream) {
    PN_REQUIRE(rays_alive && rays_t && sigmas && rgbs && deltas && weights_sum && depth && image, "null pointer");
    if (n_alive == 0) return PN_OK;
    composite_kernel<<<div_up(n_alive, (uint32_t)kRayBlock), kRayBlock, 0, PN_STREAM(stream)>>>(
        n_alive, n_step, T_thresh, rays_alive, rays_t, sigmas, rgbs, deltas, weights_sum, depth, image);
    PN_LAUNCH_CHECK("composite_kernel");
    return PN_OK;
}

extern "C" int pn_march_rays_quadratic_bending(
    const int *pig_cnt, const int *pig_bgn, const int *pig_idx, int n_vtx, int n_grid, const float *p_def,
    const float *p_ori, const float *F_IP, const float *dF_IP, int max_iter_num, const float *bbmin, const float *bbmax,
    float hgs, const int *resolution, int num_seek_IP, float IP_dx, int cut, const float *cut_bounds, uint32_t n_alive,
    uint32_t n_step, const int *rays_alive, const float *rays_t, const float *rays_o, const float *rays_d, float bound,
    float dt_gamma, uint32_t max_steps, uint32_t C, uint32_t H, const uint8_t *grid, const float *nears,
    const float *fars, float *xyzs, float *dirs, float *deltas, const float *noises, void *stream) {
    (void)nears; (void)n_vtx;
    PN_REQUIRE(pig_cnt && pig_bgn && pig_idx && p_def && p_ori && F_IP && dF_IP && bbmin && bbmax && resolution, "null pointer");
    PN_REQUIRE(rays_alive && rays_t && rays_o && rays_d && grid && fars && xyzs && dirs && deltas && noises, "null pointer");
    PN_REQUIRE(!cut || cut_bounds, "cut needs cut_bounds");
    PN_REQUIRE(num_seek_IP >= 1 && num_seek_IP <= 10, "num_seek_IP must be in 1..10");
    PN_REQUIRE(C >= 1 && C <= 8 && H >= 1 && H <= 256 && max_steps > 0, "bad C/H/max_steps");
    if (n_alive == 0 || n_step == 0) return PN_OK;
    MarchIO io{n_alive, n_step, rays_alive, rays_t, rays_o, rays_d, fars, noises, xyzs, dirs, deltas};
    pn::MarchCfg m;
    m.bound = bound; m.dt_gamma = dt_gamma;
    m.dt_min = 2 * 1.7320508075688772f / max_steps;
    m.dt_max = 2 * 1.7320508075688772f * (1 << (C - 1)) / H;
    m.cascade = (int)C; m.H = (int)H; m.bits = grid;
    pn::BendCfg bc{};
    bc.pig_cnt = pig_cnt; bc.pig_bgn = pig_bgn; bc.pig_idx = pig_idx;
    bc.p_ori = p_ori; bc.p_def = p_def; bc.F = F_IP; bc.dF = dF_IP;
    bc.n_grid = n_grid; bc.max_iter = max_iter_num; bc.K = num_seek_IP;
    bc.hgs = hgs; bc.IP_dx = IP_dx; bc.bound = bound; bc.cut = cut != 0;
    BendPtrs bp{bbmin, bbmax, cut_bounds, resolution};
    const uint32_t grid_dim = div_up(n_alive, (uint32_t)kRayBlock);
    cudaStream_t st = PN_STREAM(stream);
    switch (num_seek_IP) {
        case 1: march_kernel<1><<<grid_dim, kRayBlock, 0, st>>>(io, m, bc, bp); break;
        case 2: march_kernel<2><<<grid_dim, kRayBlock, 0, st>>>(io, m, bc, bp); break;
        case 3: march_kernel<3><<<grid_dim, kRayBlock, 0, st>>>(io, m, bc, bp); break;
        default: march_kernel<10><<<grid_dim, kRayBlock, 0, st>>>(io, m, bc, bp); break;
    }
    PN_LAUNCH_CHECK("march_kernel<bend>");
    return PN_OK;
}
