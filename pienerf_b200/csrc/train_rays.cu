// Training-side ray marching and compositing for sm_100a.
// Replaces raymarching/src/raymarching.cu:305-493 (march_rays_train), :495-591 (composite_rays_train_forward) and
// :593-696 (composite_rays_train_backward) of the reference.
//
// march_rays_train is three launches on the caller's stream instead of the reference's one kernel with two global
// atomics per ray:
//   1. train_count_kernel  — one thread per ray walks the occupancy bitfield and counts its samples into rays[n] =
//                            (n, -, num_steps);
//   2. train_scan_kernel   — one 1024-thread CTA turns the counts into exclusive offsets (rays[n].offset) and advances
//                            the (points, rays) counter the way the reference's atomics do;
//   3. train_write_kernel  — one thread per ray repeats the walk and writes xyzs / dirs / deltas at its offset.
// The sample stream per ray is identical to the reference's (same float arithmetic through march_device.cuh); what
// changes is the packing: ray n is row n of `rays` and the samples of ray n precede those of ray n+1, so the output
// is the same on every run (the reference's order is whatever order its atomics retire in).
#include "march_device.cuh"

namespace {

constexpr int kTrainBlock = 128;
constexpr int kScanThreads = 1024;

struct TrainRay {
    float ox, oy, oz, dx, dy, dz, rdx, rdy, rdz, t0, far;
};

__device__ __forceinline__ TrainRay load_train_ray(const pn::MarchCfg &m, const float *__restrict__ rays_o,
                                                   const float *__restrict__ rays_d, const float *__restrict__ nears,
                                                   const float *__restrict__ fars, const float *__restrict__ noises,
                                                   uint32_t n) {
    TrainRay r;
    r.ox = rays_o[3 * n]; r.oy = rays_o[3 * n + 1]; r.oz = rays_o[3 * n + 2];
    r.dx = rays_d[3 * n]; r.dy = rays_d[3 * n + 1]; r.dz = rays_d[3 * n + 2];
    r.rdx = 1 / r.dx; r.rdy = 1 / r.dy; r.rdz = 1 / r.dz;
    r.far = fars[n];
    const float near = nears[n];
    r.t0 = near + pn::step_size(m, near) * noises[n];  // perturbation of the first sample (raymarching.cu:352-355)
    return r;
}

// Walks ray r.  WRITE=false: returns the number of occupied samples (capped at `limit`).  WRITE=true: emits the first
// `limit` samples.  One loop body for both passes so the two walks cannot drift apart.
template <bool WRITE>
__device__ __forceinline__ uint32_t walk_ray(const pn::MarchCfg &m, const TrainRay &r, uint32_t limit, float *xyzs,
                                             float *dirs, float *deltas) {
    float t = r.t0, last_t = r.t0;
    uint32_t step = 0;
    while (t < r.far && step < limit) {
        const float x = pn::clampf(r.ox + t * r.dx, -m.bound, m.bound);
        const float y = pn::clampf(r.oy + t * r.dy, -m.bound, m.bound);
        const float z = pn::clampf(r.oz + t * r.dz, -m.bound, m.bound);
        const float dt = pn::step_size(m, t);
        float tt;
        if (pn::occupancy_and_exit(m, x, y, z, t, dt, r.dx, r.dy, r.dz, r.rdx, r.rdy, r.rdz, tt)) {
            t += dt;
            if (WRITE) {
                xyzs[0] = x; xyzs[1] = y; xyzs[2] = z;
                dirs[0] = r.dx; dirs[1] = r.dy; dirs[2] = r.dz;
                deltas[0] = dt;
                deltas[1] = t - last_t;  // distance from the previous sample: what depth integrates
                last_t = t;
                xyzs += 3; dirs += 3; deltas += 2;
            }
            step++;
        } else {
            do { t += pn::step_size(m, t); } while (t < tt);  // leave the empty voxel
        }
    }
    return step;
}

__global__ void __launch_bounds__(kTrainBlock) train_count_kernel(pn::MarchCfg m, uint32_t max_steps, uint32_t N,
                                                                  const float *__restrict__ rays_o,
                                                                  const float *__restrict__ rays_d,
                                                                  const float *__restrict__ nears,
                                                                  const float *__restrict__ fars,
                                                                  const float *__restrict__ noises,
                                                                  int *__restrict__ rays) {
    const uint32_t n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    const TrainRay r = load_train_ray(m, rays_o, rays_d, nears, fars, noises, n);
    rays[3 * n] = (int)n;
    rays[3 * n + 2] = (int)walk_ray<false>(m, r, max_steps, nullptr, nullptr, nullptr);
}

// Exclusive scan of rays[:,2] into rays[:,1], starting at the incoming point counter; one CTA, each thread owns a
// contiguous run of rays so the packing is ray-ordered.
__global__ void __launch_bounds__(kScanThreads) train_scan_kernel(uint32_t N, int *__restrict__ rays,
                                                                  int *__restrict__ counter) {
    __shared__ uint32_t warp_tot[kScanThreads / 32];
    __shared__ uint32_t base_s;
    const uint32_t per = div_up(N, (uint32_t)kScanThreads);
    const uint32_t lo = min(N, threadIdx.x * per), hi = min(N, lo + per);
    uint32_t mine = 0;
    for (uint32_t n = lo; n < hi; n++) mine += (uint32_t)rays[3 * n + 2];
    uint32_t incl = mine;
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= (uint32_t)o) incl += v;
    }
    if (lane == 31) warp_tot[warp] = incl;
    if (threadIdx.x == 0) base_s = (uint32_t)counter[0];
    __syncthreads();
    if (warp == 0) {
        uint32_t w = warp_tot[lane];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t v = __shfl_up_sync(0xffffffffu, w, o);
            if (lane >= (uint32_t)o) w += v;
        }
        warp_tot[lane] = w;  // inclusive over warps
    }
    __syncthreads();
    uint32_t off = base_s + (warp ? warp_tot[warp - 1] : 0u) + incl - mine;
    for (uint32_t n = lo; n < hi; n++) {
        rays[3 * n + 1] = (int)off;
        off += (uint32_t)rays[3 * n + 2];
    }
    if (threadIdx.x == kScanThreads - 1) {  // what the reference's atomicAdd(counter, num_steps) / (counter+1, 1) leave
        counter[0] = (int)(base_s + warp_tot[kScanThreads / 32 - 1]);
        counter[1] += (int)N;
    }
}

__global__ void __launch_bounds__(kTrainBlock) train_write_kernel(pn::MarchCfg m, uint32_t N, uint32_t M,
                                                                  const float *__restrict__ rays_o,
                                                                  const float *__restrict__ rays_d,
                                                                  const float *__restrict__ nears,
                                                                  const float *__restrict__ fars,
                                                                  const float *__restrict__ noises,
                                                                  const int *__restrict__ rays, float *__restrict__ xyzs,
                                                                  float *__restrict__ dirs, float *__restrict__ deltas) {
    const uint32_t n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    const uint32_t off = (uint32_t)rays[3 * n + 1], num = (uint32_t)rays[3 * n + 2];
    if (num == 0 || off + num > M) return;  // a ray that does not fit is dropped whole (raymarching.cu:418-419)
    const TrainRay r = load_train_ray(m, rays_o, rays_d, nears, fars, noises, n);
    walk_ray<true>(m, r, num, xyzs + (size_t)off * 3, dirs + (size_t)off * 3, deltas + (size_t)off * 2);
}

// One thread per ray: front-to-back accumulation; the early exit keeps the sample that crossed T_thresh.
__global__ void __launch_bounds__(kTrainBlock) train_composite_fwd_kernel(const float *__restrict__ sigmas,
                                                                          const float *__restrict__ rgbs,
                                                                          const float *__restrict__ deltas,
                                                                          const int *__restrict__ rays, uint32_t M,
                                                                          uint32_t N, float T_thresh,
                                                                          float *__restrict__ weights_sum,
                                                                          float *__restrict__ depth,
                                                                          float *__restrict__ image) {
    const uint32_t n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    const uint32_t index = (uint32_t)rays[3 * n], offset = (uint32_t)rays[3 * n + 1], num = (uint32_t)rays[3 * n + 2];
    float r = 0, g = 0, b = 0, ws = 0, t = 0, d = 0;
    if (num != 0 && offset + num <= M) {
        const float *sg = sigmas + offset, *col = rgbs + (size_t)offset * 3, *del = deltas + (size_t)offset * 2;
        float T = 1.0f;
        for (uint32_t s = 0; s < num; s++) {
            const float alpha = 1.0f - __expf(-sg[s] * del[2 * s]);
            const float w = alpha * T;
            r += w * col[3 * s]; g += w * col[3 * s + 1]; b += w * col[3 * s + 2];
            t += del[2 * s + 1];
            d += w * t;
            ws += w;
            T *= 1.0f - alpha;
            if (T < T_thresh) break;
        }
    }
    weights_sum[index] = ws;
    depth[index] = d;
    image[3 * index] = r; image[3 * index + 1] = g; image[3 * index + 2] = b;
}

// d(image, weights_sum)/d(sigma, rgb) with the forward recurrences replayed (no depth gradient, as in the reference).
__global__ void __launch_bounds__(kTrainBlock) train_composite_bwd_kernel(
    const float *__restrict__ grad_weights_sum, const float *__restrict__ grad_image, const float *__restrict__ sigmas,
    const float *__restrict__ rgbs, const float *__restrict__ deltas, const int *__restrict__ rays,
    const float *__restrict__ weights_sum, const float *__restrict__ image, uint32_t M, uint32_t N, float T_thresh,
    float *__restrict__ grad_sigmas, float *__restrict__ grad_rgbs) {
    const uint32_t n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    const uint32_t index = (uint32_t)rays[3 * n], offset = (uint32_t)rays[3 * n + 1], num = (uint32_t)rays[3 * n + 2];
    if (num == 0 || offset + num > M) return;
    const float gws = grad_weights_sum[index];
    const float gr = grad_image[3 * index], gg = grad_image[3 * index + 1], gb = grad_image[3 * index + 2];
    const float r_final = image[3 * index], g_final = image[3 * index + 1], b_final = image[3 * index + 2];
    const float ws_final = weights_sum[index];
    const float *sg = sigmas + offset, *col = rgbs + (size_t)offset * 3, *del = deltas + (size_t)offset * 2;
    float *gs = grad_sigmas + offset, *gc = grad_rgbs + (size_t)offset * 3;
    float T = 1.0f, r = 0, g = 0, b = 0;
    for (uint32_t s = 0; s < num; s++) {
        const float d0 = del[2 * s];
        const float alpha = 1.0f - __expf(-sg[s] * d0);
        const float w = alpha * T;
        const float c0 = col[3 * s], c1 = col[3 * s + 1], c2 = col[3 * s + 2];
        r += w * c0; g += w * c1; b += w * c2;
        T *= 1.0f - alpha;
        gc[3 * s] = gr * w; gc[3 * s + 1] = gg * w; gc[3 * s + 2] = gb * w;
        gs[s] = d0 * (gr * (T * c0 - (r_final - r)) + gg * (T * c1 - (g_final - g)) + gb * (T * c2 - (b_final - b)) +
                      gws * (1 - ws_final));
        if (T < T_thresh) break;
    }
}

}  // namespace

extern "C" int pn_march_rays_train(const float *rays_o, const float *rays_d, const uint8_t *grid, float bound,
                                   float dt_gamma, uint32_t max_steps, uint32_t N, uint32_t C, uint32_t H, uint32_t M,
                                   const float *nears, const float *fars, float *xyzs, float *dirs, float *deltas,
                                   int *rays, int *counter, const float *noises, void *stream) {
    if (N == 0) return PN_OK;
    PN_REQUIRE(rays_o && rays_d && grid && nears && fars && rays && counter && noises, "null pointer");
    PN_REQUIRE(M == 0 || (xyzs && dirs && deltas), "null output pointer");
    PN_REQUIRE(C >= 1 && C <= 8 && H >= 1 && H <= 256 && max_steps > 0, "bad C/H/max_steps");
    cudaStream_t st = PN_STREAM(stream);
    pn::MarchCfg m;
    m.bound = bound; m.dt_gamma = dt_gamma;
    m.dt_min = 2 * 1.7320508075688772f / max_steps;
    m.dt_max = 2 * 1.7320508075688772f * (1 << (C - 1)) / H;
    m.cascade = (int)C; m.H = (int)H; m.bits = grid;
    const uint32_t blocks = div_up(N, (uint32_t)kTrainBlock);
    train_count_kernel<<<blocks, kTrainBlock, 0, st>>>(m, max_steps, N, rays_o, rays_d, nears, fars, noises, rays);
    PN_LAUNCH_CHECK("train_count_kernel");
    train_scan_kernel<<<1, kScanThreads, 0, st>>>(N, rays, counter);
    PN_LAUNCH_CHECK("train_scan_kernel");
    train_write_kernel<<<blocks, kTrainBlock, 0, st>>>(m, N, M, rays_o, rays_d, nears, fars, noises, rays, xyzs, dirs, deltas);
    PN_LAUNCH_CHECK("train_write_kernel");
    return PN_OK;
}

extern "C" int pn_composite_rays_train_forward(const float *sigmas, const float *rgbs, const float *deltas,
                                               const int *rays, uint32_t M, uint32_t N, float T_thresh,
                                               float *weights_sum, float *depth, float *image, void *stream) {
    if (N == 0) return PN_OK;
    PN_REQUIRE(rays && weights_sum && depth && image, "null pointer");
    PN_REQUIRE(M == 0 || (sigmas && rgbs && deltas), "null sample pointer");
    train_composite_fwd_kernel<<<div_up(N, (uint32_t)kTrainBlock), kTrainBlock, 0, PN_STREAM(stream)>>>(
        sigmas, rgbs, deltas, rays, M, N, T_thresh, weights_sum, depth, image);
    PN_LAUNCH_CHECK("train_composite_fwd_kernel");
    return PN_OK;
}

extern "C" int pn_composite_rays_train_backward(const float *grad_weights_sum, const float *grad_image,
                                                const float *sigmas, const float *rgbs, const float *deltas,
                                                const int *rays, const float *weights_sum, const float *image,
                                                uint32_t M, uint32_t N, float T_thresh, float *grad_sigmas,
                                                float *grad_rgbs, void *stream) {
    if (N == 0 || M == 0) return PN_OK;
    PN_REQUIRE(grad_weights_sum && grad_image && rays && weights_sum && image, "null pointer");
    PN_REQUIRE(sigmas && rgbs && deltas && grad_sigmas && grad_rgbs, "null sample pointer");
    train_composite_bwd_kernel<<<div_up(N, (uint32_t)kTrainBlock), kTrainBlock, 0, PN_STREAM(stream)>>>(
        grad_weights_sum, grad_image, sigmas, rgbs, deltas, rays, weights_sum, image, M, N, T_thresh, grad_sigmas, grad_rgbs);
    PN_LAUNCH_CHECK("train_composite_bwd_kernel");
    return PN_OK;
}
