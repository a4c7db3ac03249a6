// Wavefront deformed-space renderer (render mode 3).  Included by render_fused.cu.
//
// The frame is rendered in a few device-resident passes; every pass is three launches on one stream, no host sync:
//   wave_march_kernel      one warp = one ray at a time, 32 lanes = 32 consecutive lattice points (same exact
//                          lattice/ballot replay as render_warp.cuh), inverse warp through the quadratic GMLS field,
//                          occupancy test; kept samples are appended to a compact global sample list (slabs of 32
//                          rows handed out by one atomic each) — up to `pass_cap` samples per ray per pass;
//   wave_field_ws_kernel   THE hash-lookup + MLP pass: 128-row tiles of the sample list; producer warps gather the 16
//                          hash-grid levels into a shared-memory ring, consumer warps run the 5-layer MLP on tcgen05
//                          with activations and accumulators in TMEM (field_tc.cuh); writes (alpha, r, g, b) per sample;
//   wave_composite_kernel  one thread = one ray: the reference's sequential front-to-back recurrence over the ray's
//                          samples of this pass (raymarching.cu:862-913), early termination, survivors are compacted
//                          into the next pass's ray list.
// pass_cap doubles every pass (32, 64, ...), so a saturating ray wastes at most ~32-63 field evaluations (the reference's
// wavefront loop has the same property with n_step <= 8 per iteration) while fog rays need only a handful of passes.
// Why three kernels instead of one fused one: marching (divergent neighbour search, ~100 registers), the field
// (gather-bound, 16-20 warps per SM of tensor-core tiles) and compositing (a scalar recurrence) want different
// occupancies; fused they ran at 12 warps / SM and 34 % issue utilisation.  The extra global traffic is 84 B / sample
// (~0.6 GB per chair frame, < 0.1 ms of HBM time).
#pragma once
#include "render_warp.cuh"

namespace {

#ifndef PN_WAVE_SLAB
#define PN_WAVE_SLAB 32
#endif
constexpr int kSlab = PN_WAVE_SLAB;   // sample rows per slab: the allocation unit of a march warp (a field tile = 128 rows of whatever slabs)
constexpr int kMaxPass = 8;

struct PassCtl { int n_alive; int next; int n_reserved; int pad; };

struct WaveArgs {
    PassCtl *ctl;          // [kMaxPass + 1], zeroed per frame
    int *alive[2];         // ray lists written by the compositor for the next pass
    float4 *rs_march;      // [N] t_next, skip_until, bitcast emitted, finished flag
    float4 *rs_comp;       // [2N] ws dep cr cg | cb tdepth last_t -
    int2 *link;            // [N] first sample row of this pass, number of samples
    float4 *xyzdt;         // [cap] rest-space position + dt
    int2 *meta;            // [cap] (ray or -1 for a padding row, bitcast t_after)
    float4 *out;           // [cap] alpha r g b
    int *slab_next;        // [cap / kSlab] row at which a warp's sample stream continues after this slab
    long long *counters;   // [0] composited samples, [1] emitted samples, [2] rays cut short by the last pass, [3] chunks deferred by a full sample list
    const pn::tc::Weights *weights_img;   // bf16 hi/lo weight images + level geometry, built once per frame (field_weights_kernel)
    const float *enc;      // MLP-only measurement (pn_mlp_forward): pre-encoded [rows,32] features replace the hash-grid gather
    int cap;               // rows available per pass
};

// --------------------------------------------------------------------------------------------- march
#ifndef PN_MARCH_MINB
#define PN_MARCH_MINB 4   // 32 warps / SM at 64 registers (some spills) beat 24 at 80 and 16 at 128: the list scan is latency-bound
#endif
#ifndef PN_MARCH_THREADS
#define PN_MARCH_THREADS 256   // A/B: 1024 with PN_MARCH_MINB 1 = one CTA per SM, so a grid of (SMs - reserve) CTAs leaves whole SMs free
#endif
template <int KMAX>
__global__ void __launch_bounds__(PN_MARCH_THREADS, PN_MARCH_MINB) wave_march_kernel(const RenderArgs A, const IpPack P, const WaveArgs Wv, int pass, int pass_cap) {
    pn::BendCfg bc = A.bend;
#pragma unroll
    for (int i = 0; i < 3; i++) { bc.bbmin[i] = A.geom->bbmin[i]; bc.bbmax[i] = A.geom->bbmax[i]; bc.hi[i] = A.geom->hi[i]; bc.res[i] = A.geom->res[i]; }
    __syncthreads();
    const pn::MarchCfg m = A.march;
    PassCtl *ctl = Wv.ctl + pass;
    const int n_alive = pass == 0 ? A.queue->n_active : ctl->n_alive;
    const int *alive = pass == 0 ? A.active : Wv.alive[(pass - 1) & 1];
    const int lane = threadIdx.x & 31;
    const uint32_t lt_mask = (1u << lane) - 1;
    int cur = 0, end = 0;                                              // this warp's write cursor inside its current slab
    long long emitted_total = 0;

    while (true) {
        int slot = 0;
        if (lane == 0) slot = atomicAdd(&ctl->next, 1);
        slot = __shfl_sync(0xffffffffu, slot, 0);
        if (slot >= n_alive) break;
        const int ray = alive[slot];
        const float ox = A.rays_o[3 * ray], oy = A.rays_o[3 * ray + 1], oz = A.rays_o[3 * ray + 2];
        const float dx = A.rays_d[3 * ray], dy = A.rays_d[3 * ray + 1], dz = A.rays_d[3 * ray + 2];
        const float rdx = 1 / dx, rdy = 1 / dy, rdz = 1 / dz;
        const float near = A.nears[ray], far = A.fars[ray];
        float t_next = near, skip_until = 0.f;
        int emitted = 0;
        // perturb (renderer.py:863, raymarching.cu:1187): the reference offsets t by noise * dt for the FIRST march call only, which
        // emits one sample per ray (n_step = 1 while every ray is alive); composite_rays then stores rays_t = near + (t - t_start),
        // i.e. the offset is dropped again, and all later calls march the lattice that starts at that rays_t
        float perturb_off = 0.f;
        if (pass > 0) {
            const float4 s = Wv.rs_march[ray];
            t_next = s.x; skip_until = s.y; emitted = __float_as_int(s.z);
        } else if (A.noises) {
            perturb_off = pn::step_size(m, near) * A.noises[ray];
            t_next = near + perturb_off;
        }
        int pass_emitted = 0, first_idx = -1;
        bool finished = false;
        while (true) {
            // lane i evaluates lattice point t_i = f^i(t_next)
            float t = t_next;
            for (int i = 0; i < lane; i++) t += pn::step_size(m, t);
            const float dt = pn::step_size(m, t);
            const float t_after = t + dt;
            const bool valid = t < far;
            const bool need = valid && t >= skip_until;
            float x = 0, y = 0, z = 0, tt = 0;
            bool emit = false;
            if (need) {
                pn::deformed_sample(bc, ox, oy, oz, dx, dy, dz, t, x, y, z);
                const bool found = bend_sample_packed<KMAX>(P, bc, x, y, z);
                const bool occ = pn::occupancy_and_exit(m, x, y, z, t, dt, dx, dy, dz, rdx, rdy, rdz, tt);
                emit = occ && found;
            }
            const uint32_t valid_m = __ballot_sync(0xffffffffu, valid);
            const uint32_t need_m = __ballot_sync(0xffffffffu, need);
            const uint32_t emit_m = __ballot_sync(0xffffffffu, emit);
            // replay the reference's visit order over this chunk (raymarching.cu:1385-1432)
            uint32_t take = 0;
            float carry = skip_until;
            int i = need_m ? __ffs(need_m) - 1 : (valid_m == 0xffffffffu ? 32 : __popc(valid_m));
            while (i < 32 && ((valid_m >> i) & 1u)) {
                if ((emit_m >> i) & 1u) {
                    const uint32_t inv = ~(emit_m >> i);
                    const int run = inv ? __ffs(inv) - 1 : 32;
                    take |= ((run >= 32 ? 0xffffffffu : ((1u << run) - 1u)) << i);
                    i += run;
                    carry = 0.f;
                } else {
                    const float tti = __shfl_sync(0xffffffffu, tt, i);
                    const uint32_t ge = __ballot_sync(0xffffffffu, valid && t >= tti) & ~((2u << i) - 1u);
                    const uint32_t inval = ~valid_m & ~((2u << i) - 1u);
                    if (ge) { i = __ffs(ge) - 1; carry = 0.f; }
                    else if (inval) { i = __ffs(inval) - 1; }
                    else { i = 32; carry = tti; }
                }
            }
            bool ray_left = i < 32;
            float t_restart = 0.f;
            const bool first_perturbed = perturb_off != 0.f && take != 0u;  // warp-uniform
            if (first_perturbed) {                                         // keep only the first emitted sample of the perturbed lattice
                const int l0 = __ffs(take) - 1;
                take = 1u << l0;
                t_restart = __shfl_sync(0xffffffffu, t_after, l0) - perturb_off;
                ray_left = false;
            }
            int ntake = __popc(take);
            if (emitted + ntake > (int)A.max_samples) {                   // per-ray sample cap
                int keep = (int)A.max_samples - emitted;
                uint32_t tk = take, out = 0;
                while (keep-- > 0 && tk) { const uint32_t lowest = tk & (0u - tk); out |= lowest; tk ^= lowest; }
                take = out; ntake = __popc(take);
            }
            // rows for the kept samples of this chunk
            const int room = end - cur;
            int nb = 0;
            if (ntake > room) {
                if (lane == 0) nb = atomicAdd(&ctl->n_reserved, kSlab);
                nb = __shfl_sync(0xffffffffu, nb, 0);
                if (nb + kSlab > Wv.cap) {                                  // sample list full: this chunk is redone next pass
                    if (lane == 0) atomicAdd((unsigned long long *)&Wv.counters[3], 1ull);
                    break;
                }
                if (lane == 0 && end > 0) Wv.slab_next[(end - 1) / kSlab] = nb;
            }
            const int rank = __popc(take & lt_mask);
            const int idx = rank < room ? cur + rank : nb + (rank - room);
            if ((take >> lane) & 1u) {
                Wv.xyzdt[idx] = make_float4(x, y, z, dt);
                Wv.meta[idx] = make_int2(ray, __float_as_int(first_perturbed ? t_restart : t_after));
            }
            if (ntake > 0 && first_idx < 0) first_idx = room > 0 ? cur : nb;
            if (ntake > room) { cur = nb + (ntake - room); end = nb + kSlab; }
            else cur += ntake;
            emitted += ntake; pass_emitted += ntake;
            if (ray_left || emitted >= (int)A.max_samples) { finished = true; break; }
            t_next = __shfl_sync(0xffffffffu, t_after, 31);
            skip_until = carry;
            if (first_perturbed) { t_next = t_restart; skip_until = 0.f; perturb_off = 0.f; }
            if (pass_emitted >= pass_cap) break;
        }
        if (lane == 0) {
            Wv.rs_march[ray] = make_float4(t_next, skip_until, __int_as_float(emitted), finished ? 1.f : 0.f);
            Wv.link[ray] = make_int2(first_idx, pass_emitted);
        }
        emitted_total += pass_emitted;
    }
    for (int i = cur + lane; i < end; i += 32) Wv.meta[i] = make_int2(-1, 0);   // padding rows of the last slab
    if (lane == 0 && emitted_total) atomicAdd((unsigned long long *)&Wv.counters[1], (unsigned long long)emitted_total);
}

// --------------------------------------------------------------------------------------------- field
// the weight image every field CTA fetches with one TMA bulk copy (instead of converting the fp32 weights per CTA per launch)
static_assert(sizeof(pn::tc::Weights) <= kWeightsImageBytes && sizeof(pn::tc::Weights) % 16 == 0, "weight image: workspace slot / TMA size");
__global__ void __launch_bounds__(256) field_weights_kernel(const pn_field_t f, pn::tc::Weights *out) { pn::tc::weights_fill(*out, f); }

// One persistent CTA per SM, warps split by role:
//   producer groups (4 warps = 128 rows each): hash-grid gather only — 32 gathers in flight per thread
//                   (encode_rows_ilp) — writing the [128,32] bf16 hi/lo input of sigma_net[0] into a ring of stages;
//   consumer groups (4 warps each): the five tensor-core layers; layer 1 reads its A operand straight from the ring
//                   stage and releases it with tcgen05.commit on the stage's `empty` mbarrier.
// The gathers (L2 latency bound) no longer stop while a tile is in its MLP phases, and vice versa.
#ifndef PN_WS_PROD
#define PN_WS_PROD 3
#endif
#ifndef PN_WS_CONS
#define PN_WS_CONS 2
#endif
#ifndef PN_WS_STAGES
#define PN_WS_STAGES 4
#endif
constexpr int kWsProd = PN_WS_PROD, kWsCons = PN_WS_CONS, kWsStages = PN_WS_STAGES;

struct __align__(128) WsStage { __nv_bfloat16 a[2][128 * 32]; };   // hi | lo, chunk-major: (k/8)*2048 + row*16 + (k%8)*2
#ifdef PN_WS_SS
using WsTile = pn::tc::TileSmem;                                    // A/B build: activations through shared memory
#else
struct WsTile { uint64_t bar; uint32_t tmem; uint32_t pad; };       // activations live in TMEM: only the barrier + TMEM base
#endif
#ifndef PN_WS_LVL0
#define PN_WS_LVL0 0            // 1: also stage hash level 0 (<= 4920 entries, 39 KB) in shared memory by TMA
#endif
constexpr int kLvl0Entries = 4920;
struct __align__(128) WaveWsSmem {
    pn::tc::Weights w;
#if PN_WS_LVL0
    float2 lvl0[kLvl0Entries];
#endif
    uint64_t wbar;
    WsTile tile[kWsCons];
    WsStage stage[kWsStages];
    uint64_t full[kWsStages], empty[kWsStages];
    int released[kWsStages];   // uses of a stage whose tile is completely done (software count next to the parity-only mbarriers, see below)
    uint32_t tmem_base;
};

__global__ void __launch_bounds__((kWsProd + kWsCons) * 128, 1) wave_field_ws_kernel(const RenderArgs A, const WaveArgs Wv, int pass) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    WaveWsSmem &S = *reinterpret_cast<WaveWsSmem *>(smem_raw);
    const int wg = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 7), 0), row = threadIdx.x & 127, lane = threadIdx.x & 31;
    const int n_rows = min(Wv.ctl[pass].n_reserved, Wv.cap);
    if (n_rows == 0) return;
    const float2 *table = reinterpret_cast<const float2 *>(A.field.embeddings);
    if (threadIdx.x == 0) {
        for (int s = 0; s < kWsStages; s++) { pn::tc::mbar_init(&S.full[s], 4); pn::tc::mbar_init(&S.empty[s], 1); S.released[s] = 0; }
        for (int g = 0; g < kWsCons; g++) pn::tc::mbar_init(&S.tile[g].bar, 1);
        pn::tc::mbar_init(&S.wbar, 1);
        pn::tc::fence_barrier_init();
        // TMA: the 41 KB weight image (bf16 hi/lo of the five layers + level geometry) — and the coarsest table level —
        // arrive as bulk copies while the other threads allocate TMEM; one mbarrier counts the bytes
        uint32_t bytes = (uint32_t)sizeof(pn::tc::Weights);
#if PN_WS_LVL0
        const uint32_t l0 = (uint32_t)(A.field.offsets[1] - A.field.offsets[0]);
        const bool stage0 = l0 <= (uint32_t)kLvl0Entries && (l0 & 1u) == 0;
        if (stage0) bytes += l0 * 8u;
#endif
        pn::tc::mbar_expect_tx(&S.wbar, bytes);
        pn::tc::tma_load_1d(&S.w, Wv.weights_img, (uint32_t)sizeof(pn::tc::Weights), &S.wbar);
#if PN_WS_LVL0
        if (stage0) pn::tc::tma_load_1d(S.lvl0, table, l0 * 8u, &S.wbar);
#endif
    }
    if (threadIdx.x < 32) pn::tc::tmem_alloc(&S.tmem_base, kWsCons <= 2 ? 256 : 512);
    pn::tc::tc_fence_before();
    __syncthreads();
    pn::tc::tc_fence_after();
    pn::tc::mbar_wait(&S.wbar, 0);
#if PN_WS_LVL0
    const float2 *lvl0 = ((uint32_t)(A.field.offsets[1] - A.field.offsets[0]) <= (uint32_t)kLvl0Entries && S.w.geo[0].dense3) ? S.lvl0 : nullptr;
#else
    const float2 *lvl0 = nullptr;
#endif
    const int n_tiles = (n_rows + 127) / 128;                           // rows past n_rows in the last tile were never written
    // the j-th tile of this CTA is tile blockIdx.x + j * gridDim.x and lives in stage j % kWsStages
    if (wg >= kWsCons) {
        // ---------------------------------------------------------------- producer
        const int pg = wg - kWsCons;
        for (int j = pg; blockIdx.x + j * (int)gridDim.x < n_tiles; j += kWsProd) {
            const int tile = blockIdx.x + j * gridDim.x, st = j % kWsStages, use = j / kWsStages;
            const int i = tile * 128 + row;
            const int2 mt = i < n_rows ? Wv.meta[i] : make_int2(-1, 0);
            const bool valid = mt.x >= 0;
            float4 sm = make_float4(0.f, 0.f, 0.f, 0.f);
            if (valid) sm = Wv.xyzdt[i];
            // A stage is shared by the producer groups (4 stages, 3 groups), and an mbarrier wait only sees the PARITY of the phase:
            // a group that got two uses ahead of a slow neighbour would read the parity of `use - 2` as "free" and overwrite a
            // stage that was never filled.  The consumers' software count keeps everybody within one use (never seen with the
            // gather, whose groups run at the same pace; reproducible with the fast feature-load producers of pn_mlp_forward).
            if (use >= 2) { while (*reinterpret_cast<volatile int *>(&S.released[st]) < use - 1) __nanosleep(20); }
            pn::tc::mbar_wait(&S.empty[st], (use & 1) ^ 1);               // stage free (a fresh barrier passes at once)
            char *hi = reinterpret_cast<char *>(S.stage[st].a[0]) + row * 16, *lo = reinterpret_cast<char *>(S.stage[st].a[1]) + row * 16;
            if (Wv.enc) {
                // features in, sigma / rgb out: the same ring, barriers and tensor-core pipeline fed by a plain 128-byte row load
                const float4 *e4 = reinterpret_cast<const float4 *>(Wv.enc + 32 * (size_t)i);
#pragma unroll
                for (int c = 0; c < 4; c++) {
                    float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
                    if (valid) { a = __ldg(e4 + 2 * c); b = __ldg(e4 + 2 * c + 1); }
                    uint4 h, l;
                    pn::tc::split_pair(a.x, a.y, h.x, l.x); pn::tc::split_pair(a.z, a.w, h.y, l.y);
                    pn::tc::split_pair(b.x, b.y, h.z, l.z); pn::tc::split_pair(b.z, b.w, h.w, l.w);
                    *reinterpret_cast<uint4 *>(hi + c * 2048) = h;
                    *reinterpret_cast<uint4 *>(lo + c * 2048) = l;
                }
            } else if (S.w.fast) {
                pn::tc::encode_rows_ilp(hi, lo, S.w, table, A.field.bound, valid, sm.x, sm.y, sm.z, lvl0);
            } else {
                // generic table shapes: per-level path (tiled grids, non power-of-two hash sizes)
                const float inv = 1.0f / (2 * A.field.bound);
                const float u = (sm.x + A.field.bound) * inv, vv = (sm.y + A.field.bound) * inv, ww = (sm.z + A.field.bound) * inv;
                const bool in = valid && !(u < 0 || u > 1 || vv < 0 || vv > 1 || ww < 0 || ww > 1);
#pragma unroll 1
                for (int l = 0; l < 16; l++) {
                    float2 e = make_float2(0.f, 0.f);
                    if (in) e = pn::lookup3_c2(table + S.w.level_off[l], S.w.geo[l], u, vv, ww, 0);
                    uint32_t hw, lw;
                    pn::tc::split_pair(e.x, e.y, hw, lw);
                    const int off = (l >> 2) * 2048 + (l & 3) * 4;
                    *reinterpret_cast<uint32_t *>(hi + off) = hw;
                    *reinterpret_cast<uint32_t *>(lo + off) = lw;
                }
            }
            pn::tc::fence_async_smem();                                   // generic-proxy stores -> visible to the tensor core
            __syncwarp();
            if (lane == 0) pn::tc::mbar_arrive(&S.full[st]);
        }
    } else {
        // ---------------------------------------------------------------- consumer
        const int cg = wg;
        WsTile &T = S.tile[cg];
        if (row == 0) T.tmem = S.tmem_base + cg * pn::tc::kTmemCols;
        pn::tc::group_sync(cg);
        uint32_t phase = 0;
        for (int j = cg; blockIdx.x + j * (int)gridDim.x < n_tiles; j += kWsCons) {
            const int tile = blockIdx.x + j * gridDim.x, st = j % kWsStages, use = j / kWsStages;
            const int i = tile * 128 + row;
            const int2 mt = i < n_rows ? Wv.meta[i] : make_int2(-1, 0);
            const bool valid = mt.x >= 0;
            float dt = 0.f, dx = 0, dy = 0, dz = 1;
            if (valid) {
                dt = Wv.xyzdt[i].w;
                dx = __ldg(A.rays_d + 3 * mt.x); dy = __ldg(A.rays_d + 3 * mt.x + 1); dz = __ldg(A.rays_d + 3 * mt.x + 2);
            }
            float sh[16];
            pn::sh_eval<4>(dx, dy, dz, sh);
            pn::tc::mbar_wait(&S.full[st], use & 1);
            float sigma, r, g, b;
#ifdef PN_WS_SS
            pn::tc::mlp_tile(T, S.w, cg, row, sh, phase, sigma, r, g, b, S.stage[st].a[0], S.stage[st].a[1], &S.empty[st]);
#else
            pn::tc::mlp_tile_ts(T, S.w, cg, row, sh, phase, sigma, r, g, b, S.stage[st].a[0], S.stage[st].a[1], &S.empty[st]);
#endif
            if (row == 0) *reinterpret_cast<volatile int *>(&S.released[st]) = use + 1;
            if (valid) {
                sigma = A.density_scale * sigma;
                Wv.out[i] = make_float4(Wv.enc ? sigma : 1.0f - __expf(-sigma * dt), r, g, b);
            }
        }
    }
    pn::tc::tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) pn::tc::tmem_dealloc(S.tmem_base, kWsCons <= 2 ? 256 : 512);
}

// --------------------------------------------------------------------------------------------- composite
#ifndef PN_COMP_BATCH
#define PN_COMP_BATCH 16     // sample rows in flight per ray: the walk is a chain of L2 round trips, 16 halves their number vs 8
#endif
constexpr int kCompBatch = PN_COMP_BATCH;
__global__ void __launch_bounds__(256) wave_composite_kernel(const RenderArgs A, const WaveArgs Wv, int pass, int last_pass) {
    const int n_alive = pass == 0 ? A.queue->n_active : Wv.ctl[pass].n_alive;
    const int *alive = pass == 0 ? A.active : Wv.alive[(pass - 1) & 1];
    int *alive_out = Wv.alive[pass & 1];
    const int lane = threadIdx.x & 31;
    long long kept = 0;
    // grid-stride over the pass's rays, one warp-uniform trip count per warp (the survivor compaction votes)
    for (int i0 = blockIdx.x * blockDim.x + (threadIdx.x & ~31); i0 < n_alive; i0 += gridDim.x * blockDim.x) {
        const int i = i0 + lane;
        bool survive = false;
        int ray = -1;
        if (i < n_alive) {
            ray = alive[i];
            const int2 lk = Wv.link[ray];
            const bool finished = Wv.rs_march[ray].w != 0.f;
            const float near = A.nears[ray], far = A.fars[ray];
            float ws = 0, dep = 0, cr = 0, cg = 0, cb = 0, tdepth = near, last_t = near;
            if (pass > 0) {
                const float4 a = Wv.rs_comp[2 * ray], b = Wv.rs_comp[2 * ray + 1];
                ws = a.x; dep = a.y; cr = a.z; cg = a.w; cb = b.x; tdepth = b.y; last_t = b.z;
            }
            bool terminated = false;
            int idx = lk.x, rem = lk.y;
            while (rem > 0 && !terminated) {
                // rows of a ray are contiguous up to the end of a slab: kCompBatch independent loads per step, then the recurrence;
                // the row at which the ray continues after this slab is fetched first (the only dependent load of the walk)
                const int seg = min(rem, kSlab - (idx & (kSlab - 1)));
                int next_idx = 0;
                if (rem > seg) next_idx = Wv.slab_next[(idx + seg) / kSlab - 1];
                for (int s = 0; s < seg && !terminated; s += kCompBatch) {
                    float4 o[kCompBatch];
                    float ta[kCompBatch];
#pragma unroll
                    for (int j = 0; j < kCompBatch; j++)
                        if (s + j < seg) { o[j] = Wv.out[idx + s + j]; ta[j] = __int_as_float(Wv.meta[idx + s + j].y); }
#pragma unroll
                    for (int j = 0; j < kCompBatch; j++) {
                        if (s + j < seg && !terminated) {
                            const float T = 1 - ws;                             // raymarching.cu:890-906
                            const float w = o[j].x * T;
                            ws += w;
                            tdepth += ta[j] - last_t;
                            last_t = ta[j];
                            dep += w * tdepth;
                            cr += w * o[j].y; cg += w * o[j].z; cb += w * o[j].w;
                            kept++;
                            if (T < A.T_thresh) terminated = true;
                        }
                    }
                }
                rem -= seg; idx = next_idx;
            }
            if (terminated || finished || last_pass) {
                const size_t o = A.pix ? (size_t)A.pix[ray] : (size_t)ray;   // this ray's pixel in the (possibly remote) frame
                A.image[3 * o] = cr + (1 - ws) * A.bg; A.image[3 * o + 1] = cg + (1 - ws) * A.bg; A.image[3 * o + 2] = cb + (1 - ws) * A.bg;
                A.depth0[o] = dep;
                A.depth[o] = fmaxf(dep - near, 0.f) / (far - near);
                A.wsum[ray] = ws;
                if (!terminated && !finished) atomicAdd((unsigned long long *)&Wv.counters[2], 1ull);   // cut short: out of passes
            } else {
                survive = true;
                Wv.rs_comp[2 * ray] = make_float4(ws, dep, cr, cg);
                Wv.rs_comp[2 * ray + 1] = make_float4(cb, tdepth, last_t, 0.f);
            }
        }
        const uint32_t sm = __ballot_sync(0xffffffffu, survive);
        if (sm) {
            int base = 0;
            if (lane == 0) base = atomicAdd(&Wv.ctl[pass + 1].n_alive, __popc(sm));
            base = __shfl_sync(0xffffffffu, base, 0);
            if (survive) alive_out[base + __popc(sm & ((1u << lane) - 1))] = ray;
        }
    }
    for (int o = 16; o > 0; o >>= 1) kept += __shfl_xor_sync(0xffffffffu, kept, o);
    if (lane == 0 && kept) atomicAdd((unsigned long long *)&Wv.counters[0], (unsigned long long)kept);
}

// pn_mlp_forward: a one-pass "sample list" whose row i is (ray i, dt 1), and the unpacking of its results
__global__ void __launch_bounds__(256) mlp_rows_kernel(uint32_t M, int2 *__restrict__ meta, float4 *__restrict__ xyzdt, PassCtl *__restrict__ ctl) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0) ctl[0].n_reserved = (int)M;
    if (i < M) { meta[i] = make_int2((int)i, 0); xyzdt[i] = make_float4(0.f, 0.f, 0.f, 1.f); }
}
__global__ void __launch_bounds__(256) mlp_unpack_kernel(uint32_t M, const float4 *__restrict__ out, float *__restrict__ sigmas, float *__restrict__ rgbs) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= M) return;
    const float4 o = out[i];
    sigmas[i] = o.x; rgbs[3 * i] = o.y; rgbs[3 * i + 1] = o.z; rgbs[3 * i + 2] = o.w;
}

__global__ void wave_stats_kernel(const FrameQueue *q, const FrameGeom *g, const WaveArgs Wv, int n_pass, long long *stats) {
    stats[0] = Wv.counters[0];   // composited samples
    stats[1] = q->n_active;      // rays that hit the IP box
    stats[2] = Wv.counters[1];   // field evaluations (>= stats[0]: samples marched past an early termination)
    long long rows = 0;
    for (int p = 0; p < n_pass; p++) rows += min(Wv.ctl[p].n_reserved, Wv.cap);
    stats[3] = rows;             // rows the field kernel processed (samples + slab padding)
    stats[4] = (g->overflow ? 1 : 0) | (Wv.counters[2] ? 2 : 0);   // bit 0: IP grid clamped; bit 1: rays cut short (sample list + passes exhausted)
    stats[5] = Wv.counters[2];   // rays finalised by the last pass before they finished or terminated
    stats[6] = Wv.counters[3];   // 32-lattice-point chunks a full sample list deferred to the next pass (harmless unless [5] > 0)
    int used = 0;
    for (int p = 0; p < n_pass; p++) used += (p == 0 ? q->n_active : Wv.ctl[p].n_alive) > 0;
    stats[7] = used;             // passes that had rays to march (a frame pipeline sizes its graphs with this, io->max_passes)
}

}  // namespace
