// Warp-cooperative deformed-space renderer (render mode 0).  Included by render_fused.cu.
//
// Observation that makes this exact: every t the reference's march loop ever visits on a ray lies on ONE
// deterministic lattice t_{k+1} = t_k + clamp(t_k * dt_gamma, dt_min, dt_max), t_0 = near, whether the step is an
// emitted sample (`t += dt`) or part of the empty-space skip loop (`do t += dt while (t < tt)`),
// raymarching.cu:1385-1432.  So the 32 lanes of a warp evaluate 32 consecutive lattice points of the SAME ray in
// parallel (inverse warp + occupancy), and a ballot-driven resolve replays the reference's visit order
// (emit -> next point, skip -> first point with t >= voxel exit).  Consequences:
//   * neighbouring lanes sit ~dt apart: their IP-grid cells, IP records and hash-grid cells mostly coincide,
//     so gathers coalesce into a few cache lines instead of 32;
//   * the work unit is a 32-sample chunk, not a ray: load balance no longer depends on rays per lane;
//   * emitted samples go through a per-warp FIFO (tagged by ray) so the field kernel always runs on full
//     32-sample tiles even when rays are short; compositing replays the reference's sequential recurrence in
//     order (bit-identical accumulation order), with early termination fed back to the marcher.
#pragma once
#include "field_tc.cuh"

namespace {

struct IpPack {
    const float4 *pos;      // [n] (p_def.xyz, bitcast original ip) in cell order
    const float *rec;       // [n,16]: p_ori(0..2) p_def(3..5) Finv(6..14) for the max_iter_num == 1 fast path
    const int *cell_start;  // [n_grid+1]
    const int *nb_start;    // [n_grid+1] CSR of the per-cell neighbourhood lists
    const float4 *nb_list;  // [<= 27 n] (p_def.xyz, bitcast position k in the cell-sorted arrays)
};

// Per-frame neighbourhood lists: for every IP-grid cell, the IPs of its 27-cell neighbourhood in exactly the order in
// which the reference's search visits them (own cell, then the 26 offsets of kNeigh — applied as (x,y,z) offsets by
// find_closest_IPs and as (z,y,x) offsets by find_closest_IP, raymarching.cu:986-1118 — IPs of a cell in ascending
// index).  A lattice point then scans ONE contiguous list with a strict-less insertion: the reference's tie-breaking
// falls out of the order, no per-row cell_start lookups, no rank keys, no divergent row loop, loads are independent of
// the loop state (unrolled, prefetchable).  Every IP appears in at most 27 lists: <= 27 n entries (0.9 MB for 2k IPs).
__global__ void __launch_bounds__(256) nb_count_kernel(const int *__restrict__ cnt, const int *__restrict__ res, int n_grid_cap,
                                                       int *__restrict__ nb_cnt) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    const int r0 = res[0], r1 = res[1], r2 = res[2];
    if (c >= min(r0 * r1 * r2, n_grid_cap)) return;
    const int g0 = c % r0, g1 = (c / r0) % r1, g2 = c / (r0 * r1);
    int n = 0;
    for (int dz = -1; dz <= 1; dz++)
        for (int dy = -1; dy <= 1; dy++)
            for (int dx = -1; dx <= 1; dx++) {
                const int a0 = g0 + dx, a1 = g1 + dy, a2 = g2 + dz;
                if (a0 >= 0 && a0 < r0 && a1 >= 0 && a1 < r1 && a2 >= 0 && a2 < r2) n += cnt[(a2 * r1 + a1) * r0 + a0];
            }
    nb_cnt[c] = n;
}
__global__ void __launch_bounds__(256) nb_fill_kernel(const int *__restrict__ cell_start, const float4 *__restrict__ pos,
                                                      const int *__restrict__ res, int n_grid_cap, int zyx_order,
                                                      int *__restrict__ nb_start, float4 *__restrict__ nb_list) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    const int r0 = res[0], r1 = res[1], r2 = res[2];
    const int n_grid = min(r0 * r1 * r2, n_grid_cap);
    if (c >= n_grid) return;
    const int g0 = c % r0, g1 = (c / r0) % r1, g2 = c / (r0 * r1);
    int o = nb_start[c];
    for (int q = -1; q < 26; q++) {
        int d0 = 0, d1 = 0, d2 = 0;
        if (q >= 0) {
            if (zyx_order) { d2 = pn::kNeigh[q][0]; d1 = pn::kNeigh[q][1]; d0 = pn::kNeigh[q][2]; }
            else { d0 = pn::kNeigh[q][0]; d1 = pn::kNeigh[q][1]; d2 = pn::kNeigh[q][2]; }
        }
        const int a0 = g0 + d0, a1 = g1 + d1, a2 = g2 + d2;
        if (a0 < 0 || a0 >= r0 || a1 < 0 || a1 >= r1 || a2 < 0 || a2 >= r2) continue;
        const int gid = (a2 * r1 + a1) * r0 + a0;
        for (int k = cell_start[gid]; k < cell_start[gid + 1]; k++) {
            const float4 p = pos[k];
            nb_list[o++] = make_float4(p.x, p.y, p.z, __int_as_float(k));
        }
    }
    if (c == n_grid - 1) nb_start[n_grid] = o;
}

// K nearest IPs of (x,y,z) among the 27 cells around cell (g0,g1,g2): one pass over the cell's neighbourhood list.
// ks[] = positions in the cell-sorted arrays, nearest first; ties keep the earlier-visited IP (strict <), as the
// reference's insertion does.  KMAX == 1 follows find_closest_IP: the own cell alone decides unless it is empty.
template <int KMAX>
__device__ __forceinline__ int nearest_list(const IpPack &P, const pn::BendCfg &c, float x, float y, float z, int g0, int g1, int g2,
                                            int (&ks)[KMAX]) {
    const int cell = (g2 * c.res[1] + g1) * c.res[0] + g0;
    int s = __ldg(P.nb_start + cell);
    const int e = __ldg(P.nb_start + cell + 1);
    float bd[KMAX];
#pragma unroll
    for (int i = 0; i < KMAX; i++) { bd[i] = KMAX == 1 ? 9999.9f : FLT_MAX; ks[i] = -1; }
    if (KMAX == 1) {
        const int e_own = s + (__ldg(P.cell_start + cell + 1) - __ldg(P.cell_start + cell));
        for (; s < e_own; s++) {
            const float4 q = __ldg(P.nb_list + s);
            const float d = (q.x - x) * (q.x - x) + (q.y - y) * (q.y - y) + (q.z - z) * (q.z - z);
            if (d < bd[0]) { bd[0] = d; ks[0] = __float_as_int(q.w); }
        }
        if (ks[0] != -1) return 1;
    }
#ifndef PN_LIST_BATCH
#define PN_LIST_BATCH 8
#endif
    // PN_LIST_BATCH independent loads are issued before the first distance is needed (the list is L2-resident, the scan is
    // latency-bound); indices past the end re-read the last entry and are masked out
    for (int k = s; k < e; k += PN_LIST_BATCH) {
        float4 qs[PN_LIST_BATCH];
#pragma unroll
        for (int j = 0; j < PN_LIST_BATCH; j++) qs[j] = __ldg(P.nb_list + min(k + j, e - 1));
#pragma unroll
        for (int j = 0; j < PN_LIST_BATCH; j++) {
            const float4 q = qs[j];
            const float d = (q.x - x) * (q.x - x) + (q.y - y) * (q.y - y) + (q.z - z) * (q.z - z);
            if (k + j < e && d < bd[KMAX - 1]) {
                const int id = __float_as_int(q.w);
                if (KMAX == 1) {
                    bd[0] = d; ks[0] = id;
                } else if (KMAX == 2) {
                    if (d < bd[0]) { bd[KMAX - 1] = bd[0]; ks[KMAX - 1] = ks[0]; bd[0] = d; ks[0] = id; }
                    else { bd[KMAX - 1] = d; ks[KMAX - 1] = id; }
                } else {
                    if (d < bd[1 % KMAX]) {
                        bd[KMAX - 1] = bd[1 % KMAX]; ks[KMAX - 1] = ks[1 % KMAX];
                        if (d < bd[0]) { bd[1 % KMAX] = bd[0]; ks[1 % KMAX] = ks[0]; bd[0] = d; ks[0] = id; }
                        else { bd[1 % KMAX] = d; ks[1 % KMAX] = id; }
                    } else { bd[KMAX - 1] = d; ks[KMAX - 1] = id; }
                }
            }
        }
    }
    int found = 0;
#pragma unroll
    for (int i = 0; i < KMAX; i++) found += ks[i] != -1;
    return found;
}

// Per-frame packing of the IP state in IP-grid cell order.  Finv uses the very arithmetic of the per-sample
// inverse (pn::adjugate_inverse), so hoisting it out of the march loop changes no bit when max_iter_num == 1
// (first Newton iterate: q = 0 => A = F exactly, b = -q_ exactly; raymarching.cu:1268-1304).
__global__ void __launch_bounds__(256) ip_pack_kernel(const float *__restrict__ p_def, const float *__restrict__ p_ori,
                                                      const float *__restrict__ F, const int *__restrict__ idx, int n,
                                                      const int *__restrict__ res, int n_grid_cap, int *__restrict__ bgn,
                                                      float4 *__restrict__ pos_out, float *__restrict__ rec_out) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k == 0) bgn[min(res[0] * res[1] * res[2], n_grid_cap)] = n;  // terminal entry of the CSR
    if (k >= n) return;
    const int ip = idx[k];
    const float px = p_def[3 * ip], py = p_def[3 * ip + 1], pz = p_def[3 * ip + 2];
    pos_out[k] = make_float4(px, py, pz, __int_as_float(ip));
    float A[9], Ai[9];
#pragma unroll
    for (int i = 0; i < 9; i++) A[i] = F[9 * ip + i] + 0.0f;  // Fk[m] + dFk_q[m] with dFk_q == +0
    pn::adjugate_inverse(A, Ai);
    float *r = rec_out + 16 * (size_t)k;
    r[0] = p_ori[3 * ip]; r[1] = p_ori[3 * ip + 1]; r[2] = p_ori[3 * ip + 2];
    r[3] = px; r[4] = py; r[5] = pz;
#pragma unroll
    for (int i = 0; i < 9; i++) r[6 + i] = Ai[i];
    r[15] = 0.f;
}

// bend_sample (march_device.cuh) over the packed, cell-sorted IP state.  Identical decisions and arithmetic.
template <int KMAX>
__device__ __forceinline__ bool bend_sample_packed(const IpPack &P, const pn::BendCfg &c, float &x, float &y, float &z) {
    if (c.cut && !(x > c.cb[0] && x < c.cb[1] && y > c.cb[2] && x < c.cb[3] && z > c.cb[4] && z < c.cb[5])) return true;
    int g0 = (int)floorf((x - c.bbmin[0]) / c.hgs);
    int g1 = (int)floorf((y - c.bbmin[1]) / c.hgs);
    int g2 = (int)floorf((z - c.bbmin[2]) / c.hgs);
    g0 = min(max(g0, 0), c.res[0] - 1); g1 = min(max(g1, 0), c.res[1] - 1); g2 = min(max(g2, 0), c.res[2] - 1);
    int ks[KMAX];
    int n_ip = nearest_list<KMAX>(P, c, x, y, z, g0, g1, g2, ks);
    if (n_ip <= 0) return false;
    for (int k = 0; k < n_ip; k++) {  // boundary filter with the shrinking loop bound (raymarching.cu:1246-1251)
        const float4 q = __ldg(P.pos + ks[k < KMAX ? k : 0]);
        if (q.x <= c.bbmin[0] || q.y <= c.bbmin[1] || q.z < c.bbmin[2] || q.x >= c.bbmax[0] || q.y >= c.bbmax[1] || q.z >= c.bbmax[2]) n_ip--;
    }
    if (n_ip <= 0) return false;
    float ps[KMAX][3], po[KMAX][3];
#pragma unroll
    for (int k = 0; k < KMAX; k++) { ps[k][0] = ps[k][1] = ps[k][2] = 0.f; po[k][0] = po[k][1] = po[k][2] = 0.f; }
#pragma unroll
    for (int k = 0; k < KMAX; k++) {
        if (k < n_ip) {
            float p[3];
            if (c.max_iter == 1) {
                const float4 *r4 = reinterpret_cast<const float4 *>(P.rec + 16 * (size_t)ks[k]);
                const float4 r0 = __ldg(r4), r1 = __ldg(r4 + 1), r2 = __ldg(r4 + 2), r3 = __ldg(r4 + 3);
                po[k][0] = r0.x; po[k][1] = r0.y; po[k][2] = r0.z;
                const float Ai[9] = {r1.z, r1.w, r2.x, r2.y, r2.z, r2.w, r3.x, r3.y, r3.z};
                const float q_[3] = {x - r0.w, y - r1.x, z - r1.y};
                float b[3], dq[3];
#pragma unroll
                for (int i = 0; i < 3; i++) b[i] = (float)((double)0.0f + 0.5 * (double)0.0f - (double)q_[i]);
                pn::matvec_cm(Ai, b, dq);
                p[0] = po[k][0] - dq[0]; p[1] = po[k][1] - dq[1]; p[2] = po[k][2] - dq[2];
            } else {
                const int ip = __float_as_int(__ldg(P.pos + ks[k]).w);
                pn::newton_rest_point(c, ip, x, y, z, p);
                po[k][0] = c.p_ori[3 * ip]; po[k][1] = c.p_ori[3 * ip + 1]; po[k][2] = c.p_ori[3 * ip + 2];
            }
            if (fabsf(p[0] - po[k][0]) > c.IP_dx || fabsf(p[1] - po[k][1]) > c.IP_dx || fabsf(p[2] - po[k][2]) > c.IP_dx) n_ip--;
            ps[k][0] = p[0]; ps[k][1] = p[1]; ps[k][2] = p[2];
        }
    }
    float xm = 0.f, ym = 0.f, zm = 0.f;
    if (n_ip == 1) {
        xm = ps[0][0]; ym = ps[0][1]; zm = ps[0][2];
    } else if (KMAX >= 2 && n_ip == 2) {
        float d[2];
#pragma unroll
        for (int k = 0; k < 2; k++) d[k] = sqrtf((po[k % KMAX][0] - x) * (po[k % KMAX][0] - x) + (po[k % KMAX][1] - y) * (po[k % KMAX][1] - y) + (po[k % KMAX][2] - z) * (po[k % KMAX][2] - z));
        const float s = d[0] + d[1], w0 = d[1] / s, w1 = d[0] / s;
        xm = w0 * ps[0][0] + w1 * ps[1 % KMAX][0];
        ym = w0 * ps[0][1] + w1 * ps[1 % KMAX][1];
        zm = w0 * ps[0][2] + w1 * ps[1 % KMAX][2];
    } else if (KMAX >= 3 && n_ip == 3) {
        float d[3];
#pragma unroll
        for (int k = 0; k < 3; k++) d[k] = sqrtf((po[k % KMAX][0] - x) * (po[k % KMAX][0] - x) + (po[k % KMAX][1] - y) * (po[k % KMAX][1] - y) + (po[k % KMAX][2] - z) * (po[k % KMAX][2] - z));
        const float s = d[0] * d[1] + d[1] * d[2] + d[2] * d[0];
        const float w0 = d[1] * d[2] / s, w1 = d[0] * d[2] / s, w2 = d[0] * d[1] / s;
        xm = w0 * ps[0][0] + w1 * ps[1 % KMAX][0] + w2 * ps[2 % KMAX][0];
        ym = w0 * ps[0][1] + w1 * ps[1 % KMAX][1] + w2 * ps[2 % KMAX][1];
        zm = w0 * ps[0][2] + w1 * ps[1 % KMAX][2] + w2 * ps[2 % KMAX][2];
    }
    x = xm; y = ym; z = zm;
    return true;
}

constexpr int kQueue = 64;   // per-warp sample FIFO (power of two)
constexpr int kRing = 128;   // in-flight ray descriptors per warp (> kQueue + 2)

struct __align__(16) WarpShared {
    float q[kQueue][8];     // x y z dirx diry dirz dt -
    int2 qmeta[kQueue];     // (ray tag, bitcast t_after): what the compositor needs, one 64-bit load
    float4 st[32];          // (alpha, r, g, b) of the current tile
    int ring_ray[kRing];    // per in-flight ray (tag % kRing): pixel id, near, far
    float ring_near[kRing], ring_far[kRing];
};

constexpr int kTcGroups = 3;   // 128-sample tile groups per CTA in tensor-core mode (384 threads, 1 CTA / SM)
struct __align__(128) RenderTcSmem {
    pn::tc::Weights w;
    pn::tc::TileSmem tile[kTcGroups];
    uint32_t tmem_base;
    int nq[kTcGroups][4];       // samples each warp of a group contributes to the current tile
};

// what rund_cuda leaves for a finished ray (renderer.py:896-901); one shared copy of the code (instruction-fetch budget)
static __device__ __noinline__ void finalize_ray(const RenderArgs &A, int ray, float near, float far, float ws, float dep, float cr,
                                                 float cg, float cb) {
    A.image[3 * ray] = cr + (1 - ws) * A.bg; A.image[3 * ray + 1] = cg + (1 - ws) * A.bg; A.image[3 * ray + 2] = cb + (1 - ws) * A.bg;
    A.depth0[ray] = dep;
    A.depth[ray] = fmaxf(dep - near, 0.f) / (far - near);
    A.wsum[ray] = ws;
}

// TC = false: fp32 SIMT field (128 threads, 3 CTAs / SM).  TC = true: MLP on tcgen05 (field_tc.cuh), 384 threads.
template <int KMAX, bool TC>
__global__ void __launch_bounds__(TC ? kTcGroups * 128 : 128, TC ? 1 : 3) render_warp_kernel(const RenderArgs A, const IpPack P) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    pn::FieldBlockSmem &fbs = *reinterpret_cast<pn::FieldBlockSmem *>(smem_raw);
    pn::FieldSmem &fs = fbs.w;
    float *scratch = fbs.scratch + threadIdx.x;
    RenderTcSmem &TS = *reinterpret_cast<RenderTcSmem *>(smem_raw);
    const int group = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 7), 0), row = threadIdx.x & 127;  // provably warp-uniform
    WarpShared *wsh_all = reinterpret_cast<WarpShared *>(smem_raw + (((TC ? sizeof(RenderTcSmem) : sizeof(pn::FieldBlockSmem)) + 127) & ~size_t(127)));
    uint32_t phase = 0;
    if constexpr (TC) {
        pn::tc::weights_fill(TS.w, A.field);
        if (row == 0) pn::tc::mbar_init(&TS.tile[group].bar, 1);
        pn::tc::fence_barrier_init();
        if (threadIdx.x < 32) pn::tc::tmem_alloc(&TS.tmem_base, 512);
        pn::tc::fence_async_smem();
        pn::tc::tc_fence_before();
    } else {
        pn::field_smem_fill(fs, A.field);
    }
    pn::BendCfg bc = A.bend;
#pragma unroll
    for (int i = 0; i < 3; i++) { bc.bbmin[i] = A.geom->bbmin[i]; bc.bbmax[i] = A.geom->bbmax[i]; bc.hi[i] = A.geom->hi[i]; bc.res[i] = A.geom->res[i]; }
    __syncthreads();
    if constexpr (TC) {
        pn::tc::tc_fence_after();
        if (row == 0) TS.tile[group].tmem = TS.tmem_base + group * pn::tc::kTmemCols;
        pn::tc::group_sync(group);
    }
    const pn::MarchCfg m = A.march;
    const float2 *table = reinterpret_cast<const float2 *>(A.field.embeddings);
    const int lane = threadIdx.x & 31;
    const uint32_t lt_mask = (1u << lane) - 1;
    WarpShared &W = wsh_all[threadIdx.x >> 5];
    const int n_active = A.queue->n_active;

    // marcher state (warp-uniform)
    bool m_have = false, exhausted = false;
    int m_tag = -1, next_tag = 0, m_emitted = 0;
    float ox = 0, oy = 0, oz = 0, dx = 0, dy = 0, dz = 0, rdx = 0, rdy = 0, rdz = 0, m_far = 0, m_near = 0;
    float t_next = 0, skip_until = 0;
    int m_ray = -1;
    // FIFO
    int qhead = 0, qcount = 0;
    // compositor state (warp-uniform, every lane runs the same recurrence)
    int c_tag = -1;
    bool c_done = true;
    float ws = 0, dep = 0, cr = 0, cg = 0, cb = 0, tdepth = 0, last_t = 0;
    long long kept = 0, evaluated = 0;

    auto finalize = [&](int tag) {
        if (lane == 0)
            finalize_ray(A, W.ring_ray[tag & (kRing - 1)], W.ring_near[tag & (kRing - 1)], W.ring_far[tag & (kRing - 1)], ws, dep, cr, cg, cb);
    };

    while (true) {
        // ------------------------------------------------------------------ march until a full tile is queued
        while (qcount < 32 && !(exhausted && !m_have)) {
            if (!m_have) {
                int slot = 0;
                if (lane == 0) slot = atomicAdd(&A.queue->next, 1);
                slot = __shfl_sync(0xffffffffu, slot, 0);
                if (slot >= n_active) { exhausted = true; break; }
                m_ray = A.active[slot];
                ox = A.rays_o[3 * m_ray]; oy = A.rays_o[3 * m_ray + 1]; oz = A.rays_o[3 * m_ray + 2];
                dx = A.rays_d[3 * m_ray]; dy = A.rays_d[3 * m_ray + 1]; dz = A.rays_d[3 * m_ray + 2];
                rdx = 1 / dx; rdy = 1 / dy; rdz = 1 / dz;
                m_near = A.nears[m_ray]; m_far = A.fars[m_ray];
                t_next = m_near; skip_until = 0.f;
                m_tag = next_tag++; m_emitted = 0; m_have = true;
                __syncwarp();
                if (lane == 0) { W.ring_ray[m_tag & (kRing - 1)] = m_ray; W.ring_near[m_tag & (kRing - 1)] = m_near; W.ring_far[m_tag & (kRing - 1)] = m_far; }
            }
            // lane i evaluates lattice point t_i = f^i(t_next)
            float t = t_next;
            for (int i = 0; i < lane; i++) t += pn::step_size(m, t);
            const float dt = pn::step_size(m, t);
            const float t_after = t + dt;
            const bool valid = t < m_far;
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
            // replay the reference's visit order over this chunk
            uint32_t take = 0;
            float carry = skip_until;
            int i = need_m ? __ffs(need_m) - 1 : (valid_m == 0xffffffffu ? 32 : __popc(valid_m));
            while (i < 32 && ((valid_m >> i) & 1u)) {
                if ((emit_m >> i) & 1u) {
                    const uint32_t inv = ~(emit_m >> i);
                    const int run = inv ? __ffs(inv) - 1 : 32;           // consecutive emits: each advances one point
                    take |= ((run >= 32 ? 0xffffffffu : ((1u << run) - 1u)) << i);
                    i += run;
                    carry = 0.f;
                } else {
                    const float tti = __shfl_sync(0xffffffffu, tt, i);
                    const uint32_t ge = __ballot_sync(0xffffffffu, valid && t >= tti) & ~((2u << i) - 1u);
                    const uint32_t inval = ~valid_m & ~((2u << i) - 1u);
                    if (ge) { i = __ffs(ge) - 1; carry = 0.f; }
                    else if (inval) { i = __ffs(inval) - 1; }            // skipped past `far` inside this chunk
                    else { i = 32; carry = tti; }
                }
            }
            const bool ray_left = i < 32;                                // reached a lattice point with t >= far
            // per-ray sample cap (the reference's loop stops issuing steps after max_steps)
            int ntake = __popc(take);
            if (m_emitted + ntake > (int)A.max_samples) {
                int keep = (int)A.max_samples - m_emitted;
                uint32_t tk = take, out = 0;
                while (keep-- > 0 && tk) { const uint32_t lowest = tk & (0u - tk); out |= lowest; tk ^= lowest; }
                take = out; ntake = __popc(take);
            }
            if ((take >> lane) & 1u) {
                const int s = (qhead + qcount + __popc(take & lt_mask)) & (kQueue - 1);
                float *e = W.q[s];
                *reinterpret_cast<float4 *>(e) = make_float4(x, y, z, dx);
                *reinterpret_cast<float4 *>(e + 4) = make_float4(dy, dz, dt, 0.f);
                W.qmeta[s] = make_int2(m_tag, __float_as_int(t_after));
            }
            qcount += ntake; m_emitted += ntake;
            const bool capped = m_emitted >= (int)A.max_samples;
            if (ray_left || capped) {
                if (m_emitted == 0) {
                    // crossed the IP box without a single kept sample: what rund_cuda leaves for such a ray
                    if (lane == 0) {
                        A.image[3 * m_ray] = A.bg; A.image[3 * m_ray + 1] = A.bg; A.image[3 * m_ray + 2] = A.bg;
                        A.depth0[m_ray] = 0.f; A.wsum[m_ray] = 0.f;
                        A.depth[m_ray] = fmaxf(0.f - m_near, 0.f) / (m_far - m_near);
                    }
                    next_tag--;                                            // nothing queued: the tag (ring slot) is free again
                }
                m_have = false;
            } else {
                t_next = __shfl_sync(0xffffffffu, t_after, 31);
                skip_until = carry;
            }
            __syncwarp();
        }
        // ------------------------------------------------------------------ field on a 32-sample tile
        const int n = min(32, qcount);
        float alpha = 0.f, r = 0.f, g = 0.f, b = 0.f;
        if constexpr (TC) {
            // the 4 warps of a group pool their tiles into one 128-row tensor-core tile; the group leaves together
            if (lane == 0) TS.nq[group][(threadIdx.x >> 5) & 3] = n;
            pn::tc::group_sync(group);
            const int total = TS.nq[group][0] + TS.nq[group][1] + TS.nq[group][2] + TS.nq[group][3];
            pn::tc::group_sync(group);
            if (total == 0) break;
            const bool valid = lane < n;
            const float *e = W.q[(qhead + (valid ? lane : 0)) & (kQueue - 1)];
            float sh[16];
            pn::sh_eval<4>(valid ? e[3] : 0.f, valid ? e[4] : 0.f, valid ? e[5] : 1.f, sh);
            pn::tc::encode_to_tile(TS.tile[group], TS.w, table, A.field.bound, row, valid, e[0], e[1], e[2]);
            float sigma;
            pn::tc::mlp_tile(TS.tile[group], TS.w, group, row, sh, phase, sigma, r, g, b);
            if (valid) {
                sigma = A.density_scale * sigma;
                alpha = 1.0f - __expf(-sigma * e[6]);
            }
        } else {
            if (qcount == 0) break;
            if (lane < n) {
                const float *e = W.q[(qhead + lane) & (kQueue - 1)];
                float sh[16];
                pn::sh_eval<4>(e[3], e[4], e[5], sh);
                float sigma;
                pn::field_eval(fs, table, A.field.bound, e[0], e[1], e[2], sh, scratch, pn::kFieldThreads, sigma, r, g, b);
                sigma = A.density_scale * sigma;
                alpha = 1.0f - __expf(-sigma * e[6]);
            }
        }
        W.st[lane] = make_float4(alpha, r, g, b);
        __syncwarp();
        evaluated += n;
        // ------------------------------------------------------------------ composite, in FIFO order (raymarching.cu:862-913)
        for (int j = 0; j < n; j++) {
            const int2 meta = W.qmeta[(qhead + j) & (kQueue - 1)];
            const int tag = meta.x;
            if (tag != c_tag) {
                if (c_tag >= 0 && !c_done) finalize(c_tag);
                c_tag = tag; c_done = false;
                ws = dep = cr = cg = cb = 0.f;
                tdepth = W.ring_near[tag & (kRing - 1)]; last_t = tdepth;
            }
            if (c_done) continue;                                          // samples marched past an early termination
            const float4 sa = W.st[j];
            const float T = 1 - ws;
            const float w = sa.x * T;
            ws += w;
            const float ta = __int_as_float(meta.y);
            tdepth += ta - last_t;
            last_t = ta;
            dep += w * tdepth;
            cr += w * sa.y; cg += w * sa.z; cb += w * sa.w;
            kept++;
            if (T < A.T_thresh) {
                finalize(c_tag);
                c_done = true;
                if (m_have && m_tag == c_tag) m_have = false;              // stop marching a saturated ray
            }
        }
        qhead = (qhead + n) & (kQueue - 1);
        qcount -= n;
        __syncwarp();
    }
    if (c_tag >= 0 && !c_done) finalize(c_tag);
    if (lane == 0 && (kept || evaluated)) {
        atomicAdd((unsigned long long *)&A.queue->samples, (unsigned long long)kept);
        atomicAdd((unsigned long long *)&A.queue->pad, (unsigned long long)evaluated);
    }
    if constexpr (TC) {
        pn::tc::tc_fence_before();
        __syncthreads();
        if (threadIdx.x < 32) pn::tc::tmem_dealloc(TS.tmem_base, 512);
    }
}

}  // namespace
