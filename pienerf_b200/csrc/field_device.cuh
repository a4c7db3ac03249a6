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
    // row-major [out][in] exactly as nn.Linear stores them, except the *T ones which are [in][out(+pad)]
    float W1[64 * 32];   // sigma_net[0]   [64,32]
    float W2T[64 * 16];  // sigma_net[1]^T [64,16]
    float W3[64 * 32];   // color_net[0]   [64,31] padded to 32 (pad column = 0)
    float W4[64 * 64];   // color_net[1]   [64,64]
    float W5T[64 * 4];   // color_net[2]^T [64,3] padded to 4
    LevelGeom geo[kLevels];
    uint32_t level_off[kLevels];
};

// Cooperative fill by the whole block; caller __syncthreads() afterwards.
__device__ __forceinline__ void field_smem_fill(FieldSmem &s, const pn_field_t &f) {
    const int tid = threadIdx.x, nt = blockDim.x;
    for (int i = tid; i < 64 * 32; i += nt) s.W1[i] = __ldg(f.w_sigma0 + i);
    for (int i = tid; i < 64 * 16; i += nt) { const int j = i / 16, o = i % 16; s.W2T[i] = __ldg(f.w_sigma1 + o * 64 + j); }
    for (int i = tid; i < 64 * 32; i += nt) { const int j = i / 32, k = i % 32; s.W3[i] = k < 31 ? __ldg(f.w_color0 + j * 31 + k) : 0.f; }
    for (int i = tid; i < 64 * 64; i += nt) s.W4[i] = __ldg(f.w_color1 + i);
    for (int i = tid; i < 64 * 4; i += nt) { const int j = i / 4, o = i % 4; s.W5T[i] = o < 3 ? __ldg(f.w_color2 + o * 64 + j) : 0.f; }
    if (tid < kLevels) {
        s.geo[tid] = level_geom(tid, f.S, f.H, f.offsets, false);
        s.level_off[tid] = (uint32_t)f.offsets[tid];
    }
}

// 16-level encode of a world-space point (grid.py:149 maps [-bound,bound] -> [0,1]); enc[32] level-major
// pairs, i.e. the [B, L*C] row the reference feeds to sigma_net.
__device__ __forceinline__ void encode_point(const FieldSmem &s, const float2 *__restrict__ table, float bound, float x,
                                             float y, float z, float (&enc)[32]) {
    // torch evaluates `(inputs + bound) / (2 * bound)` on CUDA as a multiply by the fp32 reciprocal of the scalar
    const float inv = 1.0f / (2 * bound);
    const float u = (x + bound) * inv, v = (y + bound) * inv, w = (z + bound) * inv;
    const bool oob = (u < 0 || u > 1 || v < 0 || v > 1 || w < 0 || w > 1);
#pragma unroll
    for (int l = 0; l < kLevels; l++) {
        float2 r = make_float2(0.f, 0.f);
        if (!oob) r = lookup3_c2(table + s.level_off[l], s.geo[l], u, v, w, 0);
        enc[2 * l] = r.x;
        enc[2 * l + 1] = r.y;
    }
}

__device__ __forceinline__ float dot32(const float *__restrict__ wrow, const float (&a)[32]) {
    float acc = 0.f;
    const float4 *w4 = reinterpret_cast<const float4 *>(wrow);
#pragma unroll
    for (int k = 0; k < 8; k++) {
        const float4 w = w4[k];
        acc += w.x * a[4 * k]; acc += w.y * a[4 * k + 1]; acc += w.z * a[4 * k + 2]; acc += w.w * a[4 * k + 3];
    }
    return acc;
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

// sigma_net on the 32 encoded features; h[0] is the pre-activation density, h[1..15] the geometry feature.
__device__ __forceinline__ void sigma_net(const FieldSmem &s, const float (&enc)[32], float (&h)[16]) {
#pragma unroll
    for (int o = 0; o < 16; o++) h[o] = 0.f;
#pragma unroll 4
    for (int j = 0; j < 64; j++) {
        const float a = fmaxf(dot32(s.W1 + j * 32, enc), 0.f);
        const float4 *w = reinterpret_cast<const float4 *>(s.W2T + j * 16);
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const float4 ww = w[q];
            h[4 * q] += ww.x * a; h[4 * q + 1] += ww.y * a; h[4 * q + 2] += ww.z * a; h[4 * q + 3] += ww.w * a;
        }
    }
}

// color_net on cat(SH16(dir), geo15) -> rgb (after sigmoid)
__device__ __forceinline__ void color_net(const FieldSmem &s, const float (&in)[32], float &r, float &g, float &b) {
    float a[64];
#pragma unroll
    for (int j = 0; j < 64; j++) a[j] = fmaxf(dot32(s.W3 + j * 32, in), 0.f);
    float o0 = 0.f, o1 = 0.f, o2 = 0.f;
#pragma unroll 4
    for (int j = 0; j < 64; j++) {
        const float c = fmaxf(dot64(s.W4 + j * 64, a), 0.f);
        const float4 w = *reinterpret_cast<const float4 *>(s.W5T + j * 4);
        o0 += w.x * c; o1 += w.y * c; o2 += w.z * c;
    }
    r = 1.0f / (1.0f + expf(-o0));
    g = 1.0f / (1.0f + expf(-o1));
    b = 1.0f / (1.0f + expf(-o2));
}

// Whole field for one sample.  sh[16] = SH basis of the (unbent) ray direction.
__device__ __forceinline__ void field_eval(const FieldSmem &s, const float2 *__restrict__ table, float bound, float x,
                                           float y, float z, const float (&sh)[16], float &sigma, float &r, float &g,
                                           float &b) {
    float h[16];
    {
        float enc[32];
        encode_point(s, table, bound, x, y, z, enc);
        sigma_net(s, enc, h);
    }
    sigma = expf(h[0]);
    float in[32];
#pragma unroll
    for (int i = 0; i < 16; i++) in[i] = sh[i];
#pragma unroll
    for (int i = 0; i < 15; i++) in[16 + i] = h[1 + i];
    in[31] = 0.f;
    color_net(s, in, r, g, b);
}

}  // namespace pn
