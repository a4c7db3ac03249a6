// Device-side multiresolution hash-grid lookup shared by the stand-alone encoder kernel and the
// fused field / render kernels.  Arithmetic follows gridencoder/src/gridencoder.cu:50-197 of the
// reference operation-for-operation (same rounding points) so results are bit-comparable; the
// scheduling around it is ours.
#pragma once
#include "common.cuh"

namespace pn {

__device__ __forceinline__ uint32_t grid_primes(uint32_t d) {
    // spatial-hash primes of the reference (gridencoder.cu:54)
    switch (d) {
        case 0: return 1u;
        case 1: return 2654435761u;
        case 2: return 805459861u;
        case 3: return 3674653429u;
        case 4: return 2097192037u;
        case 5: return 1434869437u;
        default: return 2165219737u;
    }
}

struct LevelGeom {
    float scale;          // exp2f(level*S)*H - 1
    uint32_t resolution;  // ceil(scale)+1
    uint32_t size;        // entries in this level (offsets[l+1]-offsets[l])
    uint32_t stride1;     // resolution (+1 unless align_corners)
    bool dense3;          // D=3: all three axes index linearly (no hashing, no early stop)
    uint32_t mask;        // size-1 when size is a power of two (every hashed level of the reference config), else 0
};

__device__ __forceinline__ LevelGeom level_geom(uint32_t level, float S, uint32_t H, const int *__restrict__ offsets,
                                                bool align_corners) {
    LevelGeom g;
    g.scale = exp2f(level * S) * H - 1.0f;
    g.resolution = (uint32_t)ceilf(g.scale) + 1;
    g.size = (uint32_t)(offsets[level + 1] - offsets[level]);
    g.stride1 = align_corners ? g.resolution : g.resolution + 1;
    g.dense3 = (uint64_t)g.stride1 * g.stride1 * g.stride1 <= (uint64_t)g.size;
    g.mask = (g.size & (g.size - 1)) == 0 ? g.size - 1 : 0;
    return g;
}

// entry index (not yet multiplied by C) of one grid vertex; gridencoder.cu:66-84
template <uint32_t D>
__device__ __forceinline__ uint32_t vertex_index(const uint32_t (&v)[D], const LevelGeom &g, uint32_t gridtype) {
    uint32_t stride = 1, index = 0;
#pragma unroll
    for (uint32_t d = 0; d < D; d++) {
        if (stride <= g.size) {
            index += v[d] * stride;
            stride *= g.stride1;
        }
    }
    if (gridtype == 0 && stride > g.size) {
        uint32_t h = 0;
#pragma unroll
        for (uint32_t d = 0; d < D; d++) h ^= v[d] * grid_primes(d);
        index = h;
    }
    return g.mask ? (index & g.mask) : (index % g.size);  // same value; avoids a runtime-divisor modulo
}

// Generic-index fallback of lookup3_c2 (tiled grids, hashed levels whose size is not a power of two): rare, so it is
// kept out of line (scalar arguments only: nothing of the hot paths is forced through local memory) to spare the
// instruction cache of the fused kernels.  Same weights and corner order as the inline paths.
static __device__ __noinline__ float2 lookup3_c2_generic(const float2 *__restrict__ tab, uint32_t size, uint32_t stride1,
                                                         uint32_t gridtype, uint32_t gx, uint32_t gy, uint32_t gz, float px,
                                                         float py, float pz) {
    LevelGeom g;
    g.size = size; g.stride1 = stride1; g.mask = 0; g.scale = 0.f; g.resolution = 0; g.dense3 = false;
    float2 r = make_float2(0.f, 0.f);
#pragma unroll 1
    for (int c = 0; c < 8; c++) {
        const uint32_t vv[3] = {gx + (c & 1), gy + ((c >> 1) & 1), gz + ((c >> 2) & 1)};
        const float2 v = __ldg(tab + vertex_index<3>(vv, g, gridtype));
        float w = (c & 1) ? px : 1.0f - px;
        w *= (c & 2) ? py : 1.0f - py;
        w *= (c & 4) ? pz : 1.0f - pz;
        r.x += w * v.x;
        r.y += w * v.y;
    }
    return r;
}

// Trilinear lookup of one level for D=3, C=2, fp32 table, linear interpolation: the hot instantiation.
// x01 in [0,1] (caller handles the out-of-range -> zeros rule).  `tab` points at the level's first entry.
// Three warp-uniform index paths (the level decides): dense (linear index, provably < size), hashed with a
// power-of-two table (every hashed level of the reference config: y*P1 and z*P2 are hoisted out of the corner
// loop, (v+1)*P == v*P + P mod 2^32, and `& mask` replaces the runtime-divisor modulo), and the generic fallback.
// Same vertices, same weights ((wx*wy)*wz) and the same corner order of the sum as gridencoder.cu:137-197.
__device__ __forceinline__ float2 lookup3_c2(const float2 *__restrict__ tab, const LevelGeom &g, float x, float y,
                                             float z, uint32_t gridtype = 0) {
    float px = x * g.scale + 0.5f, py = y * g.scale + 0.5f, pz = z * g.scale + 0.5f;
    const float fx = floorf(px), fy = floorf(py), fz = floorf(pz);
    px -= fx; py -= fy; pz -= fz;
    const uint32_t gx = (uint32_t)fx, gy = (uint32_t)fy, gz = (uint32_t)fz;
    float2 v[8];
    // (128-bit loads serving both x-neighbours of a corner pair when they share an aligned slot were measured on B200:
    //  fewer L1 wavefronts, no gain in time — the lookup is latency-, not wavefront-bound; plain 64-bit gathers kept.)
    if (g.dense3) {
        const uint32_t s1 = g.stride1, s2 = g.stride1 * g.stride1;
        const float2 *b00 = tab + (gx + gy * s1 + gz * s2);  // idx < stride1^3 <= size
        const float2 *b10 = b00 + s1, *b01 = b00 + s2, *b11 = b01 + s1;
        v[0] = __ldg(b00); v[1] = __ldg(b00 + 1); v[2] = __ldg(b10); v[3] = __ldg(b10 + 1);
        v[4] = __ldg(b01); v[5] = __ldg(b01 + 1); v[6] = __ldg(b11); v[7] = __ldg(b11 + 1);
    } else if (gridtype == 0 && g.mask) {
        const uint32_t hy0 = gy * 2654435761u, hy1 = hy0 + 2654435761u;
        const uint32_t hz0 = gz * 805459861u, hz1 = hz0 + 805459861u;
        const uint32_t x1 = gx + 1, m = g.mask;
        v[0] = __ldg(tab + ((gx ^ hy0 ^ hz0) & m)); v[1] = __ldg(tab + ((x1 ^ hy0 ^ hz0) & m));
        v[2] = __ldg(tab + ((gx ^ hy1 ^ hz0) & m)); v[3] = __ldg(tab + ((x1 ^ hy1 ^ hz0) & m));
        v[4] = __ldg(tab + ((gx ^ hy0 ^ hz1) & m)); v[5] = __ldg(tab + ((x1 ^ hy0 ^ hz1) & m));
        v[6] = __ldg(tab + ((gx ^ hy1 ^ hz1) & m)); v[7] = __ldg(tab + ((x1 ^ hy1 ^ hz1) & m));
    } else {
        return lookup3_c2_generic(tab, g.size, g.stride1, gridtype, gx, gy, gz, px, py, pz);
    }
    const float qx = 1.0f - px, qy = 1.0f - py, qz = 1.0f - pz;
    const float w00 = qx * qy, w10 = px * qy, w01 = qx * py, w11 = px * py;  // (1*wx)*wy, 1*wx exact
    const float w[8] = {w00 * qz, w10 * qz, w01 * qz, w11 * qz, w00 * pz, w10 * pz, w01 * pz, w11 * pz};
    float2 r = make_float2(0.f, 0.f);
#pragma unroll
    for (int c = 0; c < 8; c++) {
        r.x += w[c] * v[c].x;
        r.y += w[c] * v[c].y;
    }
    return r;
}

}  // namespace pn
