// Training-side ray marching and compositing for sm_100a.
// Replaces raymarching/src/raymarching.cu:305-493 (march_rays_train), :495-591 (composite_rays_train_forward) and
// :593-696 (composite_rays_train_backward) of the reference.
//
// march_rays_train is three launches on the caller's stream instead of the reference's one kernel with two global
// atomics per ray:
//   1. train_count_kernel  — one thread per ray walks the occupancy bitfield and counts its samples into rays[n] =
//                            (-, t_first, num_steps), t_first = ray parameter of its first sample; each 128-ray CTA leaves
//                            its total in the index slot of its first ray;
//   2. train_scan_kernel   — one 1024-thread CTA turns the per-CTA totals into per-CTA base offsets (N/128 values, in
//                            place) and advances the (points, rays) counter the way the reference's atomics do;
//   3. train_write_[coop_]kernel — each CTA scans its 128 counts from its base, completes rays[n] = (n, offset, num_steps),
//                            and every thread resumes its walk AT t_first: the empty space in front of the object (most of
//                            a walk) is crossed once, not twice as in the reference.  Samples are staged in shared memory
//                            eight at a time per lane and written by the whole warp as contiguous runs.
// No scratch memory: the index and offset columns of `rays` carry the partial sums and t_first between the launches.
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

// The voxel test of pn::occupancy_and_exit (march_device.cuh) split in two, same expressions: the exit distance is only
// computed for an empty voxel, and with a single cascade (SINGLE) the mip level is 0 whatever the position and the step
// (min(C-1, .) of raymarching.cu:42-54), which removes the two frexpf of the level selection from every step.
struct Voxel { int nx, ny, nz; float mip_bound; };
template <bool SINGLE>
__device__ __forceinline__ bool voxel_occupied(const pn::MarchCfg &m, float x, float y, float z, float dt, Voxel &v) {
    const int level = SINGLE ? 0 : pn::mip_level(m, x, y, z, dt);
    v.mip_bound = fminf(scalbnf(1, level), m.bound);
    const float mip_rbound = 1 / v.mip_bound;
    const float Hm1 = (float)(m.H - 1);
    v.nx = (int)pn::clampf(0.5f * (x * mip_rbound + 1) * m.H, 0.0f, Hm1);
    v.ny = (int)pn::clampf(0.5f * (y * mip_rbound + 1) * m.H, 0.0f, Hm1);
    v.nz = (int)pn::clampf(0.5f * (z * mip_rbound + 1) * m.H, 0.0f, Hm1);
    const uint32_t index = (uint32_t)level * (uint32_t)(m.H * m.H * m.H) + pn::morton3(v.nx, v.ny, v.nz);
    return m.bits[index >> 3] & (1 << (index & 7));
}
__device__ __forceinline__ float voxel_exit(const pn::MarchCfg &m, const Voxel &v, float x, float y, float z, float t,
                                            const TrainRay &r) {
    const float rH = 1 / (float)m.H;
    const float tx = (((v.nx + 0.5f + 0.5f * copysignf(1.0f, r.dx)) * rH * 2 - 1) * v.mip_bound - x) * r.rdx;
    const float ty = (((v.ny + 0.5f + 0.5f * copysignf(1.0f, r.dy)) * rH * 2 - 1) * v.mip_bound - y) * r.rdy;
    const float tz = (((v.nz + 0.5f + 0.5f * copysignf(1.0f, r.dz)) * rH * 2 - 1) * v.mip_bound - z) * r.rdz;
    return t + fmaxf(0.0f, fminf(tx, fminf(ty, tz)));
}

// Empty-space skipping over aligned 8^3 / 4^3 blocks of voxels, with the reference's result (single cascade only).
//
// The reference leaves an empty voxel by advancing t along its lattice (t += dt) until t >= tt, tt = the voxel's exit
// (voxel_exit), and repeats per voxel.  The lattice does not depend on the voxels, only the landing points do; through a
// run of empty voxels the only landing that matters is the last one, the first lattice point past the exit E of the run.
// In Morton order an aligned 8^3 block is 64 contiguous bytes of the bitfield (an aligned 4^3 block 8 bytes), so "the
// whole block is empty" is one 64-byte read, and the block's exit plane is the exit plane of its last voxel — the same
// float value the per-voxel chain would use.  What differs is rounding: the chain computes its last tt from a later
// landing point, the float position crosses the plane within an ulp or two of where the real one does.  block_exit
// therefore returns an interval [lo, hi] that contains every value the chain's last tt can take (and the parameters at
// which the float position can cross any of the three planes the ray is heading for); the walk advances to the first
// lattice point >= lo, and only accepts it when it is also >= hi.  A lattice point inside [lo, hi) — a knife edge, a few
// 1e-4 of the skips — sends the walk back to the block's entry to take the reference's per-voxel step instead.
// Margins: position 4e-6 (>= 10x the rounding of o + t*d and of the voxel index expression at |x| <= 8), time 4e-6*(1+t).
struct Span { float lo, hi; };
__device__ __forceinline__ Span block_exit(const pn::MarchCfg &m, const Voxel &v, int log2b, float x, float y, float z, float t,
                                           const TrainRay &r) {
    const float rH = 1 / (float)m.H;
    const int keep = ~((1 << log2b) - 1), size = 1 << log2b;
    const float px = (float)((v.nx & keep) + (copysignf(1.0f, r.dx) > 0 ? size : 0));
    const float py = (float)((v.ny & keep) + (copysignf(1.0f, r.dy) > 0 ? size : 0));
    const float pz = (float)((v.nz & keep) + (copysignf(1.0f, r.dz) > 0 ? size : 0));
    const float tx = ((px * rH * 2 - 1) * v.mip_bound - x) * r.rdx;
    const float ty = ((py * rH * 2 - 1) * v.mip_bound - y) * r.rdy;
    const float tz = ((pz * rH * 2 - 1) * v.mip_bound - z) * r.rdz;
    const float et = 4e-6f * (1.0f + t);
    const float ex = 4e-6f * fabsf(r.rdx) + et, ey = 4e-6f * fabsf(r.rdy) + et, ez = 4e-6f * fabsf(r.rdz) + et;
    // an axis the ray is parallel to gives inf - inf = NaN here; fminf drops it, as it drops the reference's NaN exits
    Span s;
    s.lo = t + fmaxf(0.0f, fminf(tx - ex, fminf(ty - ey, tz - ez)));
    s.hi = t + fmaxf(0.0f, fminf(tx + ex, fminf(ty + ey, tz + ez))) + et;
    return s;
}
// 0: no skip; 3: the 8^3 block around voxel v is empty; 2: its 4^3 sub-block is
__device__ __forceinline__ int empty_block(const pn::MarchCfg &m, const Voxel &v) {
    const uint32_t index = pn::morton3(v.nx, v.ny, v.nz);
    const uint4 *blk = reinterpret_cast<const uint4 *>(m.bits + ((index & ~511u) >> 3));
    const uint4 a = __ldg(blk), b = __ldg(blk + 1), c = __ldg(blk + 2), d = __ldg(blk + 3);
    const uint32_t any = (a.x | a.y | a.z | a.w) | (b.x | b.y | b.z | b.w) | (c.x | c.y | c.z | c.w) | (d.x | d.y | d.z | d.w);
    if (any == 0) return 3;
    const uint32_t sub = (index >> 6) & 7;  // which 64-bit word pair of the 16 words
    const uint4 q = sub < 2 ? a : sub < 4 ? b : sub < 6 ? c : d;
    const uint32_t w = (sub & 1) ? (q.z | q.w) : (q.x | q.y);
    return w == 0 ? 2 : 0;
}

// A thread's samples are one contiguous run of floats per output array.  Stream4 turns that run into aligned 128-bit
// stores (scalar stores only before the first and after the last 16-byte boundary): a sample's 3+3+2 floats would
// otherwise be 8 separate 4-byte requests, each to a sector no other lane of the warp touches.
struct Stream4 {
    float *p;            // next float of the run
    float b0, b1, b2, b3;  // the four floats pushed last (b3 newest)
    uint32_t held;       // how many of them are not stored yet
    __device__ __forceinline__ explicit Stream4(float *start) : p(start), b0(0), b1(0), b2(0), b3(0), held(0) {}
    __device__ __forceinline__ void flush_scalar() {
        if (held >= 3) p[-3] = b1;
        if (held >= 2) p[-2] = b2;
        if (held >= 1) p[-1] = b3;
        held = 0;
    }
    __device__ __forceinline__ void push(float v) {
        b0 = b1; b1 = b2; b2 = b3; b3 = v;
        held++; p++;
        if ((reinterpret_cast<uintptr_t>(p) & 15) == 0) {  // a 16-byte window just completed
            if (held == 4) { *reinterpret_cast<float4 *>(p - 4) = make_float4(b0, b1, b2, b3); held = 0; }
            else flush_scalar();  // the run started inside this window
        }
    }
};

// Walks ray r.  WRITE=false: returns the number of occupied samples (capped at `limit`) and the ray parameter of the
// first one in t_first.  WRITE=true: starts AT t_first (the empty space in front of the object, most of the walk, was
// crossed by the counting pass and need not be crossed again) and emits the first `limit` samples.  One loop body for
// both passes so the two walks cannot drift apart.
// FIXED: dt_gamma == 0, the step is the same at every t (clamp(t * 0, dt_min, dt_max)), so the lattice loops are one add and
// one compare per point instead of recomputing the clamp.
template <bool WRITE, bool SINGLE, bool FIXED>
__device__ __forceinline__ uint32_t walk_ray_impl(const pn::MarchCfg &m, const TrainRay &r, uint32_t limit, float &t_first,
                                                  float *xyzs, float *dirs, float *deltas, bool blocks) {
    float t = WRITE ? t_first : r.t0, last_t = r.t0;
    uint32_t step = 0;
    Stream4 sx(xyzs), sd(dirs);
    const bool pair = (reinterpret_cast<uintptr_t>(deltas) & 7) == 0;
    const float dt_fixed = pn::step_size(m, 0.0f);
    while (t < r.far && step < limit) {
        const float x = pn::clampf(r.ox + t * r.dx, -m.bound, m.bound);
        const float y = pn::clampf(r.oy + t * r.dy, -m.bound, m.bound);
        const float z = pn::clampf(r.oz + t * r.dz, -m.bound, m.bound);
        const float dt = FIXED ? dt_fixed : pn::step_size(m, t);
        Voxel v;
        if (voxel_occupied<SINGLE>(m, x, y, z, dt, v)) {
            if (!WRITE && step == 0) t_first = t;
            t += dt;
            if (WRITE) {
                sx.push(x); sx.push(y); sx.push(z);
                sd.push(r.dx); sd.push(r.dy); sd.push(r.dz);
                const float dl = t - last_t;  // distance from the previous sample: what depth integrates
                if (pair) *reinterpret_cast<float2 *>(deltas) = make_float2(dt, dl);
                else { deltas[0] = dt; deltas[1] = dl; }
                last_t = t;
                deltas += 2;
            }
            step++;
        } else {
            if (SINGLE && blocks) {
                const int lb = empty_block(m, v);
                if (lb) {
                    const Span e = block_exit(m, v, lb, x, y, z, t, r);
                    const float stop = fminf(e.lo, r.far);
                    float u = t;
                    if (FIXED) { do { u += dt_fixed; } while (u < stop); }
                    else { do { u += pn::step_size(m, u); } while (u < stop); }
                    if (u >= e.hi || u >= r.far) { t = u; continue; }  // the landing the per-voxel chain ends on (or the end of the ray)
                }
            }
            const float tt = voxel_exit(m, v, x, y, z, t, r);
            if (FIXED) { do { t += dt_fixed; } while (t < tt); }        // leave the empty voxel
            else { do { t += pn::step_size(m, t); } while (t < tt); }
        }
    }
    if (WRITE) { sx.flush_scalar(); sd.flush_scalar(); }
    return step;
}
template <bool WRITE, bool SINGLE>
__device__ __forceinline__ uint32_t walk_ray(const pn::MarchCfg &m, const TrainRay &r, uint32_t limit, float &t_first,
                                             float *xyzs, float *dirs, float *deltas, bool blocks) {
    if (m.dt_gamma == 0.0f) return walk_ray_impl<WRITE, SINGLE, true>(m, r, limit, t_first, xyzs, dirs, deltas, blocks);
    return walk_ray_impl<WRITE, SINGLE, false>(m, r, limit, t_first, xyzs, dirs, deltas, blocks);
}

// inclusive scan over the kTrainBlock threads of a CTA; returns this thread's inclusive value, *total = CTA sum
__device__ __forceinline__ uint32_t block_scan_inclusive(uint32_t v, uint32_t *warp_tot, uint32_t *total) {
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t u = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= (uint32_t)o) v += u;
    }
    if (lane == 31) warp_tot[warp] = v;
    __syncthreads();
    uint32_t before = 0, all = 0;
#pragma unroll
    for (int w = 0; w < kTrainBlock / 32; w++) {
        const uint32_t t = warp_tot[w];
        if ((uint32_t)w < warp) before += t;
        all += t;
    }
    *total = all;
    return v + before;
}

template <bool SINGLE>
__global__ void __launch_bounds__(kTrainBlock) train_count_kernel(pn::MarchCfg m, uint32_t max_steps, uint32_t N,
                                                                  const float *__restrict__ rays_o,
                                                                  const float *__restrict__ rays_d,
                                                                  const float *__restrict__ nears,
                                                                  const float *__restrict__ fars,
                                                                  const float *__restrict__ noises,
                                                                  int *__restrict__ rays, bool blocks) {
    __shared__ uint32_t warp_tot[kTrainBlock / 32];
    const uint32_t n = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t num = 0;
    if (n < N) {
        const TrainRay r = load_train_ray(m, rays_o, rays_d, nears, fars, noises, n);
        float t_first = r.t0;
        num = walk_ray<false, SINGLE>(m, r, max_steps, t_first, nullptr, nullptr, nullptr, blocks);
        rays[3 * n + 1] = __float_as_int(t_first);  // parked in the offset slot until the write pass replaces it
        rays[3 * n + 2] = (int)num;
    }
    uint32_t total;
    block_scan_inclusive(num, warp_tot, &total);
    if (threadIdx.x == 0) rays[3 * n] = (int)total;  // CTA total, parked in the index slot of the CTA's first ray
}

// Exclusive scan of the per-CTA totals (one per kTrainBlock rays) into per-CTA base offsets, in place, starting at the
// incoming point counter; one CTA, each thread owns a contiguous run of totals.
__global__ void __launch_bounds__(kScanThreads) train_scan_kernel(uint32_t N, int *__restrict__ rays,
                                                                  int *__restrict__ counter) {
    __shared__ uint32_t warp_tot[kScanThreads / 32];
    __shared__ uint32_t base_s;
    const uint32_t nb = div_up(N, (uint32_t)kTrainBlock);
    const uint32_t per = div_up(nb, (uint32_t)kScanThreads);
    const uint32_t lo = min(nb, threadIdx.x * per), hi = min(nb, lo + per);
    uint32_t mine = 0;
    for (uint32_t b = lo; b < hi; b++) mine += (uint32_t)rays[3 * (size_t)b * kTrainBlock];
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
    for (uint32_t b = lo; b < hi; b++) {
        const size_t slot = 3 * (size_t)b * kTrainBlock;
        const uint32_t tot = (uint32_t)rays[slot];
        rays[slot] = (int)off;
        off += tot;
    }
    if (threadIdx.x == kScanThreads - 1) {  // what the reference's atomicAdd(counter, num_steps) / (counter+1, 1) leave
        counter[0] = (int)(base_s + warp_tot[kScanThreads / 32 - 1]);
        counter[1] += (int)N;
    }
}

template <bool SINGLE>
__global__ void __launch_bounds__(kTrainBlock) train_write_kernel(pn::MarchCfg m, uint32_t N, uint32_t M,
                                                                  const float *__restrict__ rays_o,
                                                                  const float *__restrict__ rays_d,
                                                                  const float *__restrict__ nears,
                                                                  const float *__restrict__ fars,
                                                                  const float *__restrict__ noises,
                                                                  int *__restrict__ rays, float *__restrict__ xyzs,
                                                                  float *__restrict__ dirs, float *__restrict__ deltas,
                                                                  bool blocks) {
    __shared__ uint32_t warp_tot[kTrainBlock / 32];
    const uint32_t n = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t base = (uint32_t)rays[3 * (size_t)blockIdx.x * kTrainBlock];  // this CTA's first sample
    const uint32_t num = n < N ? (uint32_t)rays[3 * n + 2] : 0u;
    uint32_t total;
    const uint32_t off = base + block_scan_inclusive(num, warp_tot, &total) - num;  // the barrier inside orders the read of `base`
    if (n >= N) return;
    float t_first = __int_as_float(rays[3 * n + 1]);
    rays[3 * n] = (int)n;                   // the row is final from here: (ray, offset, num_steps)
    rays[3 * n + 1] = (int)off;
    if (num == 0 || off + num > M) return;  // a ray that does not fit is dropped whole (raymarching.cu:418-419)
    const TrainRay r = load_train_ray(m, rays_o, rays_d, nears, fars, noises, n);
    walk_ray<true, SINGLE>(m, r, num, t_first, xyzs + (size_t)off * 3, dirs + (size_t)off * 3, deltas + (size_t)off * 2, blocks);
}

// Cooperative write pass (the default; pn_set_train_write_mode(0) selects train_write_kernel above).  Same thread-per-ray walk, but a lane stages up
// to kStage samples in shared memory and the warp then writes every lane's staged run together: lane i stores float i of the run,
// so a run of 8 samples leaves as one 96-byte (xyzs, dirs) / 64-byte (deltas) contiguous store per array instead of twenty 4..16-byte
// pieces spread over the walk.  The walk body is walk_ray_impl's (same helpers, same expressions); only where the sample goes differs.
constexpr int kStage = 8;
struct WalkState { float t, last_t; uint32_t step; };

template <bool SINGLE, bool FIXED>
__device__ __forceinline__ uint32_t walk_stage(const pn::MarchCfg &m, const TrainRay &r, uint32_t limit, WalkState &w, float *sx,
                                               float *sl, bool blocks) {
    uint32_t cnt = 0;
    const float dt_fixed = pn::step_size(m, 0.0f);
    float t = w.t;
    while (t < r.far && w.step < limit && cnt < (uint32_t)kStage) {
        const float x = pn::clampf(r.ox + t * r.dx, -m.bound, m.bound);
        const float y = pn::clampf(r.oy + t * r.dy, -m.bound, m.bound);
        const float z = pn::clampf(r.oz + t * r.dz, -m.bound, m.bound);
        const float dt = FIXED ? dt_fixed : pn::step_size(m, t);
        Voxel v;
        if (voxel_occupied<SINGLE>(m, x, y, z, dt, v)) {
            t += dt;
            sx[3 * cnt] = x; sx[3 * cnt + 1] = y; sx[3 * cnt + 2] = z;
            sl[2 * cnt] = dt; sl[2 * cnt + 1] = t - w.last_t;
            w.last_t = t;
            cnt++;
            w.step++;
        } else {
            if (SINGLE && blocks) {
                const int lb = empty_block(m, v);
                if (lb) {
                    const Span e = block_exit(m, v, lb, x, y, z, t, r);
                    const float stop = fminf(e.lo, r.far);
                    float u = t;
                    if (FIXED) { do { u += dt_fixed; } while (u < stop); }
                    else { do { u += pn::step_size(m, u); } while (u < stop); }
                    if (u >= e.hi || u >= r.far) { t = u; continue; }
                }
            }
            const float tt = voxel_exit(m, v, x, y, z, t, r);
            if (FIXED) { do { t += dt_fixed; } while (t < tt); }
            else { do { t += pn::step_size(m, t); } while (t < tt); }
        }
    }
    w.t = t;
    return cnt;
}

template <bool SINGLE>
__global__ void __launch_bounds__(kTrainBlock) train_write_coop_kernel(pn::MarchCfg m, uint32_t N, uint32_t M,
                                                                       const float *__restrict__ rays_o,
                                                                       const float *__restrict__ rays_d,
                                                                       const float *__restrict__ nears,
                                                                       const float *__restrict__ fars,
                                                                       const float *__restrict__ noises,
                                                                       int *__restrict__ rays, float *__restrict__ xyzs,
                                                                       float *__restrict__ dirs, float *__restrict__ deltas,
                                                                       bool blocks) {
    __shared__ uint32_t warp_tot[kTrainBlock / 32];
    __shared__ float stage_x[kTrainBlock][3 * kStage + 1];  // +1: lanes start in different banks
    __shared__ float stage_l[kTrainBlock][2 * kStage + 1];
    const uint32_t n = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t base = (uint32_t)rays[3 * (size_t)blockIdx.x * kTrainBlock];
    const uint32_t num = n < N ? (uint32_t)rays[3 * n + 2] : 0u;
    uint32_t total;
    const uint32_t off = base + block_scan_inclusive(num, warp_tot, &total) - num;
    TrainRay r{};
    WalkState w{0.f, 0.f, 0u};
    uint32_t todo = 0;  // samples this lane still has to produce; lanes without work stay for the warp's flushes
    if (n < N) {
        w.t = __int_as_float(rays[3 * n + 1]);
        rays[3 * n] = (int)n;
        rays[3 * n + 1] = (int)off;
        if (num != 0 && off + num <= M) {  // a ray that does not fit is dropped whole (raymarching.cu:418-419)
            r = load_train_ray(m, rays_o, rays_d, nears, fars, noises, n);
            w.last_t = r.t0;
            todo = num;
        }
    }
    const bool fixed = m.dt_gamma == 0.0f;
    uint32_t done = 0;  // samples of this lane already in global memory
    float *sx = stage_x[threadIdx.x], *sl = stage_l[threadIdx.x];
    const float(*wx)[3 * kStage + 1] = stage_x + (threadIdx.x & ~31u);  // this warp's rows
    const float(*wl)[2 * kStage + 1] = stage_l + (threadIdx.x & ~31u);
    for (;;) {
        uint32_t cnt = 0;
        if (todo) cnt = fixed ? walk_stage<SINGLE, true>(m, r, num, w, sx, sl, blocks) : walk_stage<SINGLE, false>(m, r, num, w, sx, sl, blocks);
        todo -= min(todo, cnt);
        if (cnt == 0) todo = 0;  // (a walk that ends early — it cannot, the count pass walked the same ray — must not spin)
        __syncwarp();
        uint32_t pending = __ballot_sync(0xffffffffu, cnt > 0);
        if (pending == 0) break;
        while (pending) {
            const int src = __ffs(pending) - 1;
            pending &= pending - 1;
            const uint32_t c = __shfl_sync(0xffffffffu, cnt, src);
            const size_t o = (size_t)__shfl_sync(0xffffffffu, off + done, src);
            const float ddx = __shfl_sync(0xffffffffu, r.dx, src), ddy = __shfl_sync(0xffffffffu, r.dy, src),
                        ddz = __shfl_sync(0xffffffffu, r.dz, src);
            if (lane < 3 * c) {
                xyzs[3 * o + lane] = wx[src][lane];
                const uint32_t k = lane % 3;
                dirs[3 * o + lane] = k == 0 ? ddx : k == 1 ? ddy : ddz;
            }
            if (lane < 2 * c) deltas[2 * o + lane] = wl[src][lane];
        }
        done += cnt;
        __syncwarp();
    }
}

// Compositing: kRayLanes lanes per ray.  The lanes of a group load kRayLanes consecutive samples at a time (the packing is
// ray-ordered, so a warp reads one contiguous stretch of sigmas / rgbs / deltas), then every lane replays the reference's
// front-to-back recurrence over the group's samples through width-kRayLanes shuffles, in sample order: the arithmetic is the
// reference's, operation for operation (T *= 1 - alpha, r += w * c, ... and the T_thresh exit after the sample that crossed
// it), only the loads are cooperative.  A thread-per-ray loop reads 24 B per sample at a ~400 B stride between lanes.
constexpr int kRayLanes = 8;
constexpr int kRaysPerBlock = kTrainBlock / kRayLanes;

struct RayRow { uint32_t index, offset, num; bool ok; };
__device__ __forceinline__ RayRow load_ray_row(const int *__restrict__ rays, uint32_t n, uint32_t N, uint32_t M) {
    RayRow r{0, 0, 0, false};
    if (n < N) {
        r.index = (uint32_t)rays[3 * n]; r.offset = (uint32_t)rays[3 * n + 1]; r.num = (uint32_t)rays[3 * n + 2];
        r.ok = r.num != 0 && r.offset + r.num <= M;  // empty ray, or one that did not fit (raymarching.cu:524-531)
    }
    return r;
}

__global__ void __launch_bounds__(kTrainBlock) train_composite_fwd_kernel(const float *__restrict__ sigmas,
                                                                          const float *__restrict__ rgbs,
                                                                          const float *__restrict__ deltas,
                                                                          const int *__restrict__ rays, uint32_t M,
                                                                          uint32_t N, float T_thresh,
                                                                          float *__restrict__ weights_sum,
                                                                          float *__restrict__ depth,
                                                                          float *__restrict__ image) {
    const uint32_t n = blockIdx.x * kRaysPerBlock + threadIdx.x / kRayLanes;
    const uint32_t gl = threadIdx.x % kRayLanes;
    const uint32_t gmask = ((1u << kRayLanes) - 1u) << ((threadIdx.x & 31) / kRayLanes * kRayLanes);
    const RayRow row = load_ray_row(rays, n, N, M);
    if (n >= N) return;  // whole groups leave together
    float r = 0, g = 0, b = 0, ws = 0, t = 0, d = 0, T = 1.0f;
    bool done = !row.ok;
    for (uint32_t c = 0; c < row.num && !done; c += kRayLanes) {
        const uint32_t i = c + gl;
        float alpha = 0, c0 = 0, c1 = 0, c2 = 0, d1 = 0;
        if (i < row.num) {
            const size_t s = (size_t)row.offset + i;
            alpha = 1.0f - __expf(-sigmas[s] * deltas[2 * s]);
            d1 = deltas[2 * s + 1];
            c0 = rgbs[3 * s]; c1 = rgbs[3 * s + 1]; c2 = rgbs[3 * s + 2];
        }
        const uint32_t cnt = min((uint32_t)kRayLanes, row.num - c);
        for (uint32_t k = 0; k < cnt; k++) {
            const float a = __shfl_sync(gmask, alpha, k, kRayLanes);
            const float w = a * T;
            r += w * __shfl_sync(gmask, c0, k, kRayLanes);
            g += w * __shfl_sync(gmask, c1, k, kRayLanes);
            b += w * __shfl_sync(gmask, c2, k, kRayLanes);
            t += __shfl_sync(gmask, d1, k, kRayLanes);
            d += w * t;
            ws += w;
            T *= 1.0f - a;
            if (T < T_thresh) { done = true; break; }
        }
    }
    if (gl == 0) {
        weights_sum[row.index] = ws;
        depth[row.index] = d;
        image[3 * row.index] = r; image[3 * row.index + 1] = g; image[3 * row.index + 2] = b;
    }
}

// d(image, weights_sum)/d(sigma, rgb) with the forward recurrences replayed (no depth gradient, as in the reference).
// Lane k of a group keeps the gradients of the group's k-th sample and stores them with its neighbours' (coalesced);
// samples behind the T_thresh exit are not written (the caller zero-fills, raymarching.py:285-286).
__global__ void __launch_bounds__(kTrainBlock) train_composite_bwd_kernel(
    const float *__restrict__ grad_weights_sum, const float *__restrict__ grad_image, const float *__restrict__ sigmas,
    const float *__restrict__ rgbs, const float *__restrict__ deltas, const int *__restrict__ rays,
    const float *__restrict__ weights_sum, const float *__restrict__ image, uint32_t M, uint32_t N, float T_thresh,
    float *__restrict__ grad_sigmas, float *__restrict__ grad_rgbs) {
    const uint32_t n = blockIdx.x * kRaysPerBlock + threadIdx.x / kRayLanes;
    const uint32_t gl = threadIdx.x % kRayLanes;
    const uint32_t gmask = ((1u << kRayLanes) - 1u) << ((threadIdx.x & 31) / kRayLanes * kRayLanes);
    const RayRow row = load_ray_row(rays, n, N, M);
    if (n >= N || !row.ok) return;
    const float gws = grad_weights_sum[row.index];
    const float gr = grad_image[3 * row.index], gg = grad_image[3 * row.index + 1], gb = grad_image[3 * row.index + 2];
    const float r_final = image[3 * row.index], g_final = image[3 * row.index + 1], b_final = image[3 * row.index + 2];
    const float ws_final = weights_sum[row.index];
    float T = 1.0f, r = 0, g = 0, b = 0;
    bool done = false;
    for (uint32_t c = 0; c < row.num && !done; c += kRayLanes) {
        const uint32_t i = c + gl;
        const size_t s = (size_t)row.offset + i;
        float alpha = 0, c0 = 0, c1 = 0, c2 = 0, d0 = 0;
        if (i < row.num) {
            d0 = deltas[2 * s];
            alpha = 1.0f - __expf(-sigmas[s] * d0);
            c0 = rgbs[3 * s]; c1 = rgbs[3 * s + 1]; c2 = rgbs[3 * s + 2];
        }
        const uint32_t cnt = min((uint32_t)kRayLanes, row.num - c);
        bool mine = false;
        float w_mine = 0, gs_mine = 0;
        for (uint32_t k = 0; k < cnt; k++) {
            const float a = __shfl_sync(gmask, alpha, k, kRayLanes);
            const float w = a * T;
            r += w * __shfl_sync(gmask, c0, k, kRayLanes);
            g += w * __shfl_sync(gmask, c1, k, kRayLanes);
            b += w * __shfl_sync(gmask, c2, k, kRayLanes);
            T *= 1.0f - a;
            if (gl == k) {
                mine = true;
                w_mine = w;
                gs_mine = d0 * (gr * (T * c0 - (r_final - r)) + gg * (T * c1 - (g_final - g)) + gb * (T * c2 - (b_final - b)) +
                                gws * (1 - ws_final));
            }
            if (T < T_thresh) { done = true; break; }
        }
        if (mine) {
            grad_rgbs[3 * s] = gr * w_mine; grad_rgbs[3 * s + 1] = gg * w_mine; grad_rgbs[3 * s + 2] = gb * w_mine;
            grad_sigmas[s] = gs_mine;
        }
    }
}

}  // namespace

static int g_train_block_skip = 1;
static int g_train_write_mode = 1;
// Write pass of march_rays_train: 1 = warp-cooperative flush through shared memory (train_write_coop_kernel, default; 0.74 vs
// 0.90 ms per 800x800 frame), 0 = every lane streams its own samples (train_write_kernel).  Same output either way; returns the previous setting.
extern "C" int pn_set_train_write_mode(int mode) {
    const int was = g_train_write_mode;
    g_train_write_mode = mode != 0;
    return was;
}
// A/B switch for the empty-space block skipping of march_rays_train (1 = on, the default; results are identical either way).
extern "C" int pn_set_train_block_skip(int on) {
    const int was = g_train_block_skip;
    g_train_block_skip = on != 0;
    return was;
}

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
    // block skipping needs 8^3 blocks that tile the grid and 128-bit loads of the bitfield; one cascade (see block_exit)
    const bool skip = g_train_block_skip && C == 1 && H % 8 == 0 && (reinterpret_cast<uintptr_t>(grid) & 15) == 0;
    if (C == 1) train_count_kernel<true><<<blocks, kTrainBlock, 0, st>>>(m, max_steps, N, rays_o, rays_d, nears, fars, noises, rays, skip);
    else train_count_kernel<false><<<blocks, kTrainBlock, 0, st>>>(m, max_steps, N, rays_o, rays_d, nears, fars, noises, rays, false);
    PN_LAUNCH_CHECK("train_count_kernel");
    train_scan_kernel<<<1, kScanThreads, 0, st>>>(N, rays, counter);
    PN_LAUNCH_CHECK("train_scan_kernel");
    if (g_train_write_mode) {
        if (C == 1) train_write_coop_kernel<true><<<blocks, kTrainBlock, 0, st>>>(m, N, M, rays_o, rays_d, nears, fars, noises, rays, xyzs, dirs, deltas, skip);
        else train_write_coop_kernel<false><<<blocks, kTrainBlock, 0, st>>>(m, N, M, rays_o, rays_d, nears, fars, noises, rays, xyzs, dirs, deltas, false);
    } else {
        if (C == 1) train_write_kernel<true><<<blocks, kTrainBlock, 0, st>>>(m, N, M, rays_o, rays_d, nears, fars, noises, rays, xyzs, dirs, deltas, skip);
        else train_write_kernel<false><<<blocks, kTrainBlock, 0, st>>>(m, N, M, rays_o, rays_d, nears, fars, noises, rays, xyzs, dirs, deltas, false);
    }
    PN_LAUNCH_CHECK("train_write_kernel");
    return PN_OK;
}

extern "C" int pn_composite_rays_train_forward(const float *sigmas, const float *rgbs, const float *deltas,
                                               const int *rays, uint32_t M, uint32_t N, float T_thresh,
                                               float *weights_sum, float *depth, float *image, void *stream) {
    if (N == 0) return PN_OK;
    PN_REQUIRE(rays && weights_sum && depth && image, "null pointer");
    PN_REQUIRE(M == 0 || (sigmas && rgbs && deltas), "null sample pointer");
    train_composite_fwd_kernel<<<div_up(N, (uint32_t)kRaysPerBlock), kTrainBlock, 0, PN_STREAM(stream)>>>(
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
    train_composite_bwd_kernel<<<div_up(N, (uint32_t)kRaysPerBlock), kTrainBlock, 0, PN_STREAM(stream)>>>(
        grad_weights_sum, grad_image, sigmas, rgbs, deltas, rays, weights_sum, image, M, N, T_thresh, grad_sigmas, grad_rgbs);
    PN_LAUNCH_CHECK("train_composite_bwd_kernel");
    return PN_OK;
}
