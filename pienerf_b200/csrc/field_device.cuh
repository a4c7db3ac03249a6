// Fused NeRF field evaluation for one sample per thread (fp32 SIMT mode):
//   hash-grid encode (16 x trilinear) -> 32->64->16 -> trunc_exp | SH(4) || geo(15) -> 31->64->64->3 -> sigmoid
// Restates nerf/network.py:98-127 (+ grid.py:145-161, sphere_harmonics.py:75-87) of the reference as one
// device function so activations never leave registers.  Weights live in shared memory in layouts chosen
// for 128-bit broadcast reads; layers are fused pairwise (1+2, 4+5) so at most ~100 activations are live.
#pragma once
#include "grid_device.cuh"
#include "sh_device.cuh"

namespace pn {

constexpr int kLevels = 16;

struct FieldSmem {
    // Layouts chosen so that every layer runs as a ROLLED loop over one index with the other index static in
    // registers (small code: the fully unrolled form was ~9k SASS instructions and stalled on instruction fetch):
    float2 W1P[kLevels * 64];  // sigma_net[0] as [level][j] = (W[j][2l], W[j][2l+1])      -> accumulate over levels
    float W2[16 * 64];         // sigma_net[1] row-major [o][j]                            -> dot per output o
    float W3T[32 * 64];        // color_net[0]^T [k][j], k = 31 inputs (+1 zero row)        -> accumulate over inputs
    float W4[64 * 64];         // color_net[1] row-major [i][j]                            -> dot per output i
    float W5T[64 * 4];         // color_net[2]^T [i][c] padded to 4                        -> accumulate over i
    LevelGeom geo[kLevels];
    uint32_t level_off[kLevels];
};

constexpr int kFieldThreads = 128;  // block size of every kernel that evaluates the field
struct FieldBlockSmem {
    FieldSmem w;
    float scratch[32 * kFieldThreads];  // [k][thread]: colour-net input row of each thread (conflict-free columns)
};

// Cooperative fill by the whole block; caller __syncthreads() afterwards.
__device__ __forceinline__ void field_smem_fill(FieldSmem &s, const pn_field_t &f) {
    const int tid = threadIdx.x, nt = blockDim.x;
    for (int i = tid; i < kLevels * 64; i += nt) { const int l = i / 64, j = i % 64; s.W1P[i] = make_float2(__ldg(f.w_sigma0 + j * 32 + 2 * l), __ldg(f.w_sigma0 + j * 32 + 2 * l + 1)); }
    for (int i = tid; i < 16 * 64; i += nt) s.W2[i] = __ldg(f.w_sigma1 + i);
    for (int i = tid; i < 32 * 64; i += nt) { const int k = i / 64, j = i % 64; s.W3T[i] = k < 31 ? __ldg(f.w_color0 + j * 31 + k) : 0.f; }
    for (int i = tid; i < 64 * 64; i += nt) s.W4[i] = __ldg(f.w_color1 + i);
    for (int i = tid; i < 64 * 4; i += nt) { const int j = i / 4, o = i % 4; s.W5T[i] = o < 3 ? __ldg(f.w_color2 + o * 64 + j) : 0.f; }
    if (tid < kLevels) {
        s.geo[tid] = level_geom(tid, f.S, f.H, f.offsets, false);
        s.level_off[tid] = (uint32_t)f.offsets[tid];
    }
}

__device__ __forceinline__ float dot64(const float *__restrict__ wrow, const float (&a)[64]) {
    float acc = 0.f;
    const float4 *w4 = reinterpret_cast<const float4 *>(wrow);
#pragma unroll
    for (int k = 0; k < 16; k++) {
        const float4 w = w4[k];
        acc += w.x * a[4 * k]; acc += w.y * a[4 * k + 1]; acc += w.z * a[4 * k + 2]; acc += w.w * a[4 * k + 3];
    }
    return acc;
}

// v[j] += w[j] * x for the 64 static accumulators, w = one 64-float row in shared memory (16 x LDS.128 broadcast)
__device__ __forceinline__ void axpy64(const float *__restrict__ wrow, float x, float (&v)[64]) {
    const float4 *w4 = reinterpret_cast<const float4 *>(wrow);
#pragma unroll
    for (int q = 0; q < 16; q++) {
        const float4 w = w4[q];
        v[4 * q] += w.x * x; v[4 * q + 1] += w.y * x; v[4 * q + 2] += w.z * x; v[4 * q + 3] += w.w * x;
    }
}

// Whole field for one sample: 16-level encode (grid.py:149 maps [-bound,bound] -> [0,1]) feeding sigma_net layer 0
// level by level, then the colour net.  sh[16] = SH basis of the (unbent) ray direction.  `scratch` is this
// thread's 32-float column of a [32][blockDim.x] shared array (the colour-net input row: SH then geo feature).
// Every sum runs in the index order of a plain row-major GEMV, so results do not depend on this scheduling.
__device__ __forceinline__ void field_eval(const FieldSmem &s, const float2 *__restrict__ table, float bound, float x,
                                           float y, float z, const float (&sh)[16], float *__restrict__ scratch,
                                           int scratch_stride, float &sigma, float &r, float &g, float &b) {
    // torch evaluates `(inputs + bound) / (2 * bound)` on CUDA as a multiply by the fp32 reciprocal of the scalar
    const float inv = 1.0f / (2 * bound);
    const float u = (x + bound) * inv, v = (y + bound) * inv, w = (z + bound) * inv;
    const bool oob = (u < 0 || u > 1 || v < 0 || v > 1 || w < 0 || w > 1);
    float h[64];
#pragma unroll
    for (int j = 0; j < 64; j++) h[j] = 0.f;
#pragma unroll 1
    for (int l = 0; l < kLevels; l++) {                       // sigma_net[0]: h[j] += W[j][2l] e0 + W[j][2l+1] e1
        float2 e = make_float2(0.f, 0.f);
        if (!oob) e = lookup3_c2(table + s.level_off[l], s.geo[l], u, v, w, 0);
        const float4 *w4 = reinterpret_cast<const float4 *>(s.W1P + l * 64);
#pragma unroll
        for (int q = 0; q < 32; q++) {
            const float4 ww = w4[q];                          // (W[2q][2l], W[2q][2l+1], W[2q+1][2l], W[2q+1][2l+1])
            h[2 * q] += ww.x * e.x; h[2 * q] += ww.y * e.y;
            h[2 * q + 1] += ww.z * e.x; h[2 * q + 1] += ww.w * e.y;
        }
    }
#pragma unroll
    for (int j = 0; j < 64; j++) h[j] = fmaxf(h[j], 0.f);
#pragma unroll
    for (int k = 0; k < 16; k++) scratch[k * scratch_stride] = sh[k];
    float h0 = 0.f;
#pragma unroll 1
    for (int o = 0; o < 16; o++) {                            // sigma_net[1]: one output per iteration
        const float t = dot64(s.W2 + o * 64, h);
        if (o == 0) h0 = t; else scratch[(15 + o) * scratch_stride] = t;
    }
    sigma = expf(h0);                                         // trunc_exp forward (activation.py:9-11)
#pragma unroll
    for (int j = 0; j < 64; j++) h[j] = 0.f;
#pragma unroll 1
    for (int k = 0; k < 31; k++) axpy64(s.W3T + k * 64, scratch[k * scratch_stride], h);   // color_net[0]
#pragma unroll
    for (int j = 0; j < 64; j++) h[j] = fmaxf(h[j], 0.f);
    float o0 = 0.f, o1 = 0.f, o2 = 0.f;
#pragma unroll 1
    for (int i = 0; i < 64; i++) {                            // color_net[1] row i, consumed at once by color_net[2]
        const float c = fmaxf(dot64(s.W4 + i * 64, h), 0.f);
        const float4 ww = *reinterpret_cast<const float4 *>(s.W5T + i * 4);
        o0 += ww.x * c; o1 += ww.y * c; o2 += ww.z * c;
    }
    r = 1.0f / (1.0f + expf(-o0));
    g = 1.0f / (1.0f + expf(-o1));
    b = 1.0f / (1.0f + expf(-o2));
}

}  // namespace pn
