// Device-side ray-marching primitives shared by the drop-in kernels (raymarching.cu) and the fused
// frame renderer (render_fused.cu).  Semantics follow raymarching/src/raymarching.cu:42-81,704-809,
// 930-1434 of the reference including its accidental behaviours; rounding points are
// kept (expressions the reference evaluates in double because of unsuffixed literals stay in double
// unless the float form is provably identical, noted inline).
#pragma once
#include <cfloat>
#include "common.cuh"

namespace pn {

__device__ __forceinline__ float clampf(float x, float lo, float hi) { return fminf(hi, fmaxf(lo, x)); }

__device__ __forceinline__ uint32_t spread3(uint32_t v) {  // 10-bit -> every third bit
    v = (v * 0x00010001u) & 0xFF0000FFu;
    v = (v * 0x00000101u) & 0x0F00F00Fu;
    v = (v * 0x00000011u) & 0xC30C30C3u;
    v = (v * 0x00000005u) & 0x49249249u;
    return v;
}
__device__ __forceinline__ uint32_t morton3(uint32_t x, uint32_t y, uint32_t z) {
    return spread3(x) | (spread3(y) << 1) | (spread3(z) << 2);
}
__device__ __forceinline__ uint32_t compact3(uint32_t x) {
    x &= 0x49249249u;
    x = (x | (x >> 2)) & 0xc30c30c3u;
    x = (x | (x >> 4)) & 0x0f00f00fu;
    x = (x | (x >> 8)) & 0xff0000ffu;
    x = (x | (x >> 16)) & 0x0000ffffu;
    return x;
}

struct MarchCfg {
    float bound, dt_gamma, dt_min, dt_max;
    int cascade;          // C
    int H;                // density grid resolution (128)
    const uint8_t *bits;  // density bitfield
};

__device__ __forceinline__ MarchCfg make_march_cfg(float bound, float dt_gamma, uint32_t max_steps, uint32_t C,
                                                   uint32_t H, const uint8_t *bits) {
    MarchCfg m;
    m.bound = bound; m.dt_gamma = dt_gamma;
    m.dt_min = 2 * 1.7320508075688772f / max_steps;
    m.dt_max = 2 * 1.7320508075688772f * (1 << (C - 1)) / H;
    m.cascade = (int)C; m.H = (int)H; m.bits = bits;
    return m;
}

__device__ __forceinline__ float step_size(const MarchCfg &m, float t) { return clampf(t * m.dt_gamma, m.dt_min, m.dt_max); }

__device__ __forceinline__ int mip_level(const MarchCfg &m, float x, float y, float z, float dt) {
    const float Cm1 = (float)m.cascade - 1;
    int e1, e2;
    frexpf(fmaxf(fabsf(x), fmaxf(fabsf(y), fabsf(z))), &e1);
    frexpf(dt * m.H * 0.5f, &e2);  // (dt*H)*0.5 in double == float: scaling by 0.5 is exact
    const int l1 = (int)fminf(Cm1, fmaxf(0.f, (float)e1));
    const int l2 = (int)fminf(Cm1, fmaxf(0.f, (float)e2));
    return max(l1, l2);
}

// Occupancy lookup at (x,y,z) and, when the caller has to skip, the exit distance of the voxel along the ray.
// Returns occ; writes tt = t + max(0, min(tx,ty,tz)).
__device__ __forceinline__ bool occupancy_and_exit(const MarchCfg &m, float x, float y, float z, float t, float dt,
                                                   float dx, float dy, float dz, float rdx, float rdy, float rdz,
                                                   float &tt) {
    const int level = mip_level(m, x, y, z, dt);
    const float mip_bound = fminf(scalbnf(1, level), m.bound);
    const float mip_rbound = 1 / mip_bound;
    const float Hm1 = (float)(m.H - 1);
    // 0.5 * (v*rb + 1) * H in double then rounded == the float product: one rounding of an exact value
    const int nx = (int)clampf(0.5f * (x * mip_rbound + 1) * m.H, 0.0f, Hm1);
    const int ny = (int)clampf(0.5f * (y * mip_rbound + 1) * m.H, 0.0f, Hm1);
    const int nz = (int)clampf(0.5f * (z * mip_rbound + 1) * m.H, 0.0f, Hm1);
    // level*H^3 + morton is evaluated in float by the reference; exact below 2^24 (H=128, C<=8)
    const uint32_t index = (uint32_t)level * (uint32_t)(m.H * m.H * m.H) + morton3(nx, ny, nz);
    const bool occ = m.bits[index >> 3] & (1 << (index & 7));
    const float rH = 1 / (float)m.H;
    const float tx = (((nx + 0.5f + 0.5f * copysignf(1.0f, dx)) * rH * 2 - 1) * mip_bound - x) * rdx;
    const float ty = (((ny + 0.5f + 0.5f * copysignf(1.0f, dy)) * rH * 2 - 1) * mip_bound - y) * rdy;
    const float tz = (((nz + 0.5f + 0.5f * copysignf(1.0f, dz)) * rH * 2 - 1) * mip_bound - z) * rdz;
    tt = t + fmaxf(0.0f, fminf(tx, fminf(ty, tz)));
    return occ;
}

// ---------------------------------------------------------------------------------------------------
// quadratic GMLS inverse warp
// ---------------------------------------------------------------------------------------------------

struct BendCfg {
    const int *pig_cnt, *pig_bgn, *pig_idx;
    const float *p_ori, *p_def, *F, *dF;
    int n_grid, max_iter, K;
    float bbmin[3], bbmax[3], hi[3];  // hi = (float)((double)bbmax - 1e-6)
    int res[3];
    float hgs, IP_dx, bound;
    bool cut;
    float cb[6];
};

// (f,g,h) triplets of the reference's fixed 26-neighbour visiting order (raymarching.cu:1011-1021)
__constant__ signed char kNeigh[26][3] = {
    {-1, 0, 0}, {0, -1, 0}, {0, 0, -1}, {1, 0, 0}, {0, 1, 0}, {0, 0, 1},
    {-1, -1, 0}, {-1, 0, -1}, {0, -1, -1}, {1, 1, 0}, {1, 0, 1}, {0, 1, 1},
    {-1, 1, 0}, {-1, 0, 1}, {0, -1, 1}, {1, -1, 0}, {1, 0, -1}, {0, 1, -1},
    {-1, -1, 1}, {-1, 1, -1}, {1, -1, -1}, {1, 1, -1}, {1, -1, 1}, {-1, 1, 1},
    {-1, -1, -1}, {1, 1, 1}};

__device__ __forceinline__ float dist2_to(const float *__restrict__ p, float x, float y, float z) {
    return (p[0] - x) * (p[0] - x) + (p[1] - y) * (p[1] - y) + (p[2] - z) * (p[2] - z);
}

// num_seek_IP == 1: own cell, the 26 neighbours only when the own cell is empty (raymarching.cu:986-1045).
// The neighbour triplet is applied as (z,y,x) offsets here, as the reference does.
__device__ __forceinline__ int nearest_ip_single(const BendCfg &c, float x, float y, float z, int g0, int g1, int g2) {
    float best = 9999.9f;
    int ip = -1;
    {
        const int gid = (g2 * c.res[1] + g1) * c.res[0] + g0;
        const int n = c.pig_cnt[gid], b = c.pig_bgn[gid];
        for (int i = 0; i < n; i++) {
            const int k = c.pig_idx[b + i];
            const float d = dist2_to(c.p_def + 3 * k, x, y, z);
            if (d < best) { best = d; ip = k; }
        }
    }
    if (ip == -1) {
        for (int q = 0; q < 26; q++) {
            const int a2 = g2 + kNeigh[q][0], a1 = g1 + kNeigh[q][1], a0 = g0 + kNeigh[q][2];
            if (a2 >= c.res[2] || a2 < 0 || a1 >= c.res[1] || a1 < 0 || a0 >= c.res[0] || a0 < 0) continue;
            const int gid = (a2 * c.res[1] + a1) * c.res[0] + a0;
            const int n = c.pig_cnt[gid], b = c.pig_bgn[gid];
            for (int i = 0; i < n; i++) {
                const int k = c.pig_idx[b + i];
                const float d = dist2_to(c.p_def + 3 * k, x, y, z);
                if (d < best) { best = d; ip = k; }
            }
        }
    }
    return ip;
}

// num_seek_IP > 1: own cell then all 26 neighbours (triplet applied as (x,y,z) offsets), insertion-sorted
// best-K, first found wins ties (raymarching.cu:1047-1118).  Returns how many slots are filled.
template <int KMAX>
__device__ __forceinline__ int nearest_ips(const BendCfg &c, float x, float y, float z, int g0, int g1, int g2,
                                           int (&ips)[KMAX], int K) {
    float dist[KMAX];
#pragma unroll
    for (int i = 0; i < KMAX; i++) { dist[i] = FLT_MAX; ips[i] = -1; }
    auto scan = [&](int gid) {
        const int n = c.pig_cnt[gid], b = c.pig_bgn[gid];
        for (int i = 0; i < n; i++) {
            const int k = c.pig_idx[b + i];
            const float d = dist2_to(c.p_def + 3 * k, x, y, z);
#pragma unroll
            for (int j = 0; j < KMAX; j++) {
                if (j < K && d < dist[j]) {
#pragma unroll
                    for (int s = KMAX - 1; s > 0; s--)
                        if (s > j && s < K) { dist[s] = dist[s - 1]; ips[s] = ips[s - 1]; }
                    dist[j] = d; ips[j] = k;
                    break;
                }
            }
        }
    };
    scan((g2 * c.res[1] + g1) * c.res[0] + g0);
    for (int q = 0; q < 26; q++) {
        const int a0 = g0 + kNeigh[q][0], a1 = g1 + kNeigh[q][1], a2 = g2 + kNeigh[q][2];
        if (a0 >= 0 && a0 < c.res[0] && a1 >= 0 && a1 < c.res[1] && a2 >= 0 && a2 < c.res[2])
            scan((a2 * c.res[1] + a1) * c.res[0] + a0);
    }
    int found = 0;
#pragma unroll
    for (int i = 0; i < KMAX; i++) found += (i < K && ips[i] != -1);
    return found;
}

// flat 3x3 helpers with the reference's index conventions (raymarching.cu:940-984)
__device__ __forceinline__ void matvec_cm(const float *M, const float *v, float *r) {
    r[0] = M[0] * v[0] + M[3] * v[1] + M[6] * v[2];
    r[1] = M[1] * v[0] + M[4] * v[1] + M[7] * v[2];
    r[2] = M[2] * v[0] + M[5] * v[1] + M[8] * v[2];
}
__device__ __forceinline__ void adjugate_inverse(const float *A, float *R) {
    const float det = A[0] * (A[4] * A[8] - A[5] * A[7]) - A[1] * (A[3] * A[8] - A[5] * A[6]) +
                      A[2] * (A[3] * A[7] - A[4] * A[6]);
    if (det == 0) {  // the reference cannot detect this (bool vs -1 compare): its A_inv stays zero
#pragma unroll
        for (int i = 0; i < 9; i++) R[i] = 0.f;
        return;
    }
    const float id = 1.0f / det;
    R[0] = id * (A[4] * A[8] - A[5] * A[7]);
    R[1] = id * (A[2] * A[7] - A[1] * A[8]);
    R[2] = id * (A[1] * A[5] - A[2] * A[4]);
    R[3] = id * (A[5] * A[6] - A[3] * A[8]);
    R[4] = id * (A[0] * A[8] - A[2] * A[6]);
    R[5] = id * (A[2] * A[3] - A[0] * A[5]);
    R[6] = id * (A[3] * A[7] - A[4] * A[6]);
    R[7] = id * (A[1] * A[6] - A[0] * A[7]);
    R[8] = id * (A[0] * A[4] - A[1] * A[3]);
}

// Newton inverse of x_def ~ p_def + F q + 1/2 (dF.q) q around one IP (raymarching.cu:1264-1312).
__device__ __forceinline__ void newton_rest_point(const BendCfg &c, int ip, float x, float y, float z, float *p) {
    const float *pk = c.p_ori + 3 * ip, *pk_ = c.p_def + 3 * ip;
    float Fk[9];
#pragma unroll
    for (int i = 0; i < 9; i++) Fk[i] = __ldg(c.F + 9 * ip + i);
    const float *dFk = c.dF + 27 * ip;
    const float k0 = pk[0], k1 = pk[1], k2 = pk[2];
    p[0] = k0; p[1] = k1; p[2] = k2;
    const float q_[3] = {x - pk_[0], y - pk_[1], z - pk_[2]};
    for (int it = 0; it < c.max_iter;) {
        const float q[3] = {p[0] - k0, p[1] - k1, p[2] - k2};
        float dFq[9], A[9], Ai[9];
#pragma unroll
        for (int m = 0; m < 9; m++) {
            dFq[m] = __ldg(dFk + m) * q[0] + __ldg(dFk + 9 + m) * q[1] + __ldg(dFk + 18 + m) * q[2];
            A[m] = Fk[m] + dFq[m];
        }
        adjugate_inverse(A, Ai);
        float Fq[3], dFqq[3], b[3], dq[3];
        matvec_cm(Fk, q, Fq);
        matvec_cm(dFq, q, dFqq);
#pragma unroll
        for (int i = 0; i < 3; i++) b[i] = (float)((double)Fq[i] + 0.5 * (double)dFqq[i] - (double)q_[i]);
        matvec_cm(Ai, b, dq);
        p[0] -= dq[0]; p[1] -= dq[1]; p[2] -= dq[2];
        if (dq[0] * dq[0] + dq[1] * dq[1] + dq[2] * dq[2] < 1e-12) break;  // float sum compared against a double literal
        it++;
    }
}

// Maps a deformed-space sample to rest space (raymarching.cu:1210-1383).  x,y,z in/out.  Returns `found`.
template <int KMAX>
__device__ __forceinline__ bool bend_sample(const BendCfg &c, float &x, float &y, float &z) {
    if (c.cut && !(x > c.cb[0] && x < c.cb[1] && y > c.cb[2] && x < c.cb[3] && z > c.cb[4] && z < c.cb[5]))
        return true;  // outside the cut box (test keeps the reference's x-for-y typo): static, no bending
    int g0 = (int)floorf((x - c.bbmin[0]) / c.hgs);
    int g1 = (int)floorf((y - c.bbmin[1]) / c.hgs);
    int g2 = (int)floorf((z - c.bbmin[2]) / c.hgs);
    // the reference only printf()s when these leave the grid; stay in range instead of reading out of bounds
    g0 = min(max(g0, 0), c.res[0] - 1); g1 = min(max(g1, 0), c.res[1] - 1); g2 = min(max(g2, 0), c.res[2] - 1);

    int ips[KMAX];
    int n_ip;
    if (c.K == 1) {
        ips[0] = nearest_ip_single(c, x, y, z, g0, g1, g2);
        n_ip = ips[0] != -1;
    } else {
        n_ip = nearest_ips<KMAX>(c, x, y, z, g0, g1, g2, ips, c.K);
    }
    if (n_ip <= 0) return false;
    // boundary filter: the loop bound shrinks while iterating, so the farthest kept IP is dropped, not
    // the offending one (raymarching.cu:1246-1251)
    for (int k = 0; k < n_ip; k++) {
        const float *q = c.p_def + 3 * ips[k < KMAX ? k : 0];
        if (q[0] <= c.bbmin[0] || q[1] <= c.bbmin[1] || q[2] < c.bbmin[2] || q[0] >= c.bbmax[0] || q[1] >= c.bbmax[1] ||
            q[2] >= c.bbmax[2])
            n_ip--;
    }
    if (n_ip <= 0) return false;

    float ps[KMAX][3];
#pragma unroll
    for (int k = 0; k < KMAX; k++) { ps[k][0] = 0.f; ps[k][1] = 0.f; ps[k][2] = 0.f; }
#pragma unroll
    for (int k = 0; k < KMAX; k++) {
        if (k < n_ip) {  // n_ip may shrink inside the loop (raymarching.cu:1314-1319)
            float p[3];
            newton_rest_point(c, ips[k], x, y, z, p);
            const float *pk = c.p_ori + 3 * ips[k];
            if (fabsf(p[0] - pk[0]) > c.IP_dx || fabsf(p[1] - pk[1]) > c.IP_dx || fabsf(p[2] - pk[2]) > c.IP_dx) n_ip--;
            ps[k][0] = p[0]; ps[k][1] = p[1]; ps[k][2] = p[2];
        }
    }
    float xm = 0.f, ym = 0.f, zm = 0.f;  // n_ip == 0 (or > 3): the sample maps to the origin and still counts as found
    if (n_ip == 1) {
        xm = ps[0][0]; ym = ps[0][1]; zm = ps[0][2];
    } else if (KMAX >= 2 && n_ip == 2) {
        float d[2];
#pragma unroll
        for (int k = 0; k < 2; k++) d[k] = sqrtf(dist2_to(c.p_ori + 3 * ips[k], x, y, z));
        const float s = d[0] + d[1], w0 = d[1] / s, w1 = d[0] / s;
        xm = w0 * ps[0][0] + w1 * ps[1 % KMAX][0];
        ym = w0 * ps[0][1] + w1 * ps[1 % KMAX][1];
        zm = w0 * ps[0][2] + w1 * ps[1 % KMAX][2];
    } else if (KMAX >= 3 && n_ip == 3) {
        float d[3];
#pragma unroll
        for (int k = 0; k < 3; k++) d[k] = sqrtf(dist2_to(c.p_ori + 3 * ips[k % KMAX], x, y, z));
        const float s = d[0] * d[1] + d[1] * d[2] + d[2] * d[0];
        const float w0 = d[1] * d[2] / s, w1 = d[0] * d[2] / s, w2 = d[0] * d[1] / s;
        xm = w0 * ps[0][0] + w1 * ps[1 % KMAX][0] + w2 * ps[2 % KMAX][0];
        ym = w0 * ps[0][1] + w1 * ps[1 % KMAX][1] + w2 * ps[2 % KMAX][1];
        zm = w0 * ps[0][2] + w1 * ps[1 % KMAX][2] + w2 * ps[2 % KMAX][2];
    }
    x = xm; y = ym; z = zm;
    return true;
}

// The sample position before bending (raymarching.cu:1197-1208)
__device__ __forceinline__ void deformed_sample(const BendCfg &c, float ox, float oy, float oz, float dx, float dy,
                                                float dz, float t, float &x, float &y, float &z) {
    if (c.cut) {
        x = clampf(ox + t * dx, -c.bound, c.bound);
        y = clampf(oy + t * dy, -c.bound, c.bound);
        z = clampf(oz + t * dz, -c.bound, c.bound);
    } else {
        x = clampf(ox + t * dx, c.bbmin[0], c.hi[0]);
        y = clampf(oy + t * dy, c.bbmin[1], c.hi[1]);
        z = clampf(oz + t * dz, c.bbmin[2], c.hi[2]);
    }
}

}  // namespace pn
