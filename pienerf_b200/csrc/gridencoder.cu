// Multiresolution hash-grid encoder, forward only (the hot path).
// Replaces gridencoder/src/gridencoder.cu:87-245,372-400,448-471 of the reference.
//
// Two kernels:
//  * grid_forward_generic<T,D,C>  — every (D,C,dtype,gridtype,align,interp,dy_dx) combination the
//    reference's entry point accepts; one thread per (sample, level).
//  * grid_forward_d3c2            — the hot instantiation (fp32 table, D=3, C=2, no dy_dx):
//    one thread per (sample, level-pair) with 64-bit vertex loads, level-major block order so that
//    co-resident CTAs gather from the same 4 MiB level slice (L2-resident), streaming stores for
//    the [L,B,2] output.
#include "grid_device.cuh"

namespace {

template <typename T> struct Acc;
template <> struct Acc<float> {
    static __device__ __forceinline__ float load(const float *p) { return __ldg(p); }
    static __device__ __forceinline__ void fma_into(float &acc, float w, float v) { acc += w * v; }
    static __device__ __forceinline__ float zero() { return 0.f; }
    static __device__ __forceinline__ float from_float(float v) { return v; }
    static __device__ __forceinline__ float to_float(float v) { return v; }
};
template <> struct Acc<__half> {
    // the reference accumulates in at::Half: every += rounds back to fp16
    static __device__ __forceinline__ __half load(const __half *p) { return __ldg(p); }
    static __device__ __forceinline__ void fma_into(__half &acc, float w, __half v) {
        acc = __float2half(__half2float(acc) + w * __half2float(v));
    }
    static __device__ __forceinline__ __half zero() { return __float2half(0.f); }
    static __device__ __forceinline__ __half from_float(float v) { return __float2half(v); }
    static __device__ __forceinline__ float to_float(__half v) { return __half2float(v); }
};

template <typename T, uint32_t D, uint32_t C>
__global__ void __launch_bounds__(256) grid_forward_generic(const float *__restrict__ inputs, const T *__restrict__ table,
                                                            const int *__restrict__ offsets, T *__restrict__ outputs,
                                                            uint32_t B, uint32_t L, float S, uint32_t H,
                                                            T *__restrict__ dy_dx, uint32_t gridtype,
                                                            bool align_corners, uint32_t interp) {
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const uint32_t level = blockIdx.y;
    const float *in = inputs + (size_t)b * D;
    T *out = outputs + ((size_t)level * B + b) * C;

    float x[D];
    bool oob = false;
#pragma unroll
    for (uint32_t d = 0; d < D; d++) {
        x[d] = in[d];
        oob |= (x[d] < 0 || x[d] > 1);
    }
    if (oob) {  // gridencoder.cu:110-135
#pragma unroll
        for (uint32_t c = 0; c < C; c++) out[c] = Acc<T>::zero();
        if (dy_dx) {
            T *g = dy_dx + (size_t)b * D * L * C + (size_t)level * D * C;
#pragma unroll
            for (uint32_t i = 0; i < D * C; i++) g[i] = Acc<T>::zero();
        }
        return;
    }
    const pn::LevelGeom geo = pn::level_geom(level, S, H, offsets, align_corners);
    const T *tab = table + (size_t)(uint32_t)offsets[level] * C;

    float frac[D], dfrac[D];
    uint32_t cell[D];
#pragma unroll
    for (uint32_t d = 0; d < D; d++) {
        float p = x[d] * geo.scale + (align_corners ? 0.0f : 0.5f);
        const float fl = floorf(p);
        cell[d] = (uint32_t)fl;
        p -= (float)cell[d];
        if (interp == 1) {  // smoothstep (gridencoder.cu:39-47)
            dfrac[d] = 6 * p * (1.0f - p);
            p = p * p * (3.0f - 2.0f * p);
        } else {
            dfrac[d] = 1.0f;
        }
        frac[d] = p;
    }

    T acc[C];
#pragma unroll
    for (uint32_t c = 0; c < C; c++) acc[c] = Acc<T>::zero();
#pragma unroll
    for (uint32_t corner = 0; corner < (1u << D); corner++) {
        float w = 1;
        uint32_t v[D];
#pragma unroll
        for (uint32_t d = 0; d < D; d++) {
            if (corner & (1u << d)) { w *= frac[d]; v[d] = cell[d] + 1; }
            else { w *= 1 - frac[d]; v[d] = cell[d]; }
        }
        const uint32_t e = pn::vertex_index<D>(v, geo, gridtype) * C;
#pragma unroll
        for (uint32_t c = 0; c < C; c++) Acc<T>::fma_into(acc[c], w, Acc<T>::load(tab + e + c));
    }
#pragma unroll
    for (uint32_t c = 0; c < C; c++) out[c] = acc[c];

    if (dy_dx) {  // gridencoder.cu:201-244
        T *g = dy_dx + (size_t)b * D * L * C + (size_t)level * D * C;
#pragma unroll
        for (uint32_t gd = 0; gd < D; gd++) {
            T gacc[C];
#pragma unroll
            for (uint32_t c = 0; c < C; c++) gacc[c] = Acc<T>::zero();
#pragma unroll
            for (uint32_t corner = 0; corner < (1u << (D - 1)); corner++) {
                float w = geo.scale;
                uint32_t v[D];
#pragma unroll
                for (uint32_t nd = 0; nd < D - 1; nd++) {
                    const uint32_t d = (nd >= gd) ? (nd + 1) : nd;
                    if (corner & (1u << nd)) { w *= frac[d]; v[d] = cell[d] + 1; }
                    else { w *= 1 - frac[d]; v[d] = cell[d]; }
                }
                v[gd] = cell[gd];
                const uint32_t el = pn::vertex_index<D>(v, geo, gridtype) * C;
                v[gd] = cell[gd] + 1;
                const uint32_t er = pn::vertex_index<D>(v, geo, gridtype) * C;
#pragma unroll
                for (uint32_t c = 0; c < C; c++) {
                    const float diff = Acc<T>::to_float(Acc<T>::load(tab + er + c)) - Acc<T>::to_float(Acc<T>::load(tab + el + c));
                    gacc[c] = Acc<T>::from_float(Acc<T>::to_float(gacc[c]) + w * diff * dfrac[gd]);
                }
            }
#pragma unroll
            for (uint32_t c = 0; c < C; c++) g[gd * C + c] = gacc[c];
        }
    }
}

// Hot instantiation.  blockIdx.y = level (level-major order: the block scheduler drains x first).
__global__ void __launch_bounds__(256) grid_forward_d3c2(const float *__restrict__ inputs,
                                                         const float2 *__restrict__ table,
                                                         const int *__restrict__ offsets, float2 *__restrict__ outputs,
                                                         uint32_t B, float S, uint32_t H, uint32_t gridtype) {
    const uint32_t level = blockIdx.y;
    const pn::LevelGeom geo = pn::level_geom(level, S, H, offsets, false);
    const float2 *tab = table + (uint32_t)offsets[level];
    float2 *out = outputs + (size_t)level * B;
    for (uint32_t b = blockIdx.x * blockDim.x + threadIdx.x; b < B; b += gridDim.x * blockDim.x) {
        const float x = __ldg(inputs + (size_t)b * 3), y = __ldg(inputs + (size_t)b * 3 + 1), z = __ldg(inputs + (size_t)b * 3 + 2);
        float2 r = make_float2(0.f, 0.f);
        if (!(x < 0 || x > 1 || y < 0 || y > 1 || z < 0 || z > 1)) r = pn::lookup3_c2(tab, geo, x, y, z, gridtype);
        __stcs(out + b, r);
    }
}

template <typename T, uint32_t D>
int launch_generic(const float *inputs, const T *table, const int *offsets, T *outputs, uint32_t B, uint32_t C,
                   uint32_t L, float S, uint32_t H, T *dy_dx, uint32_t gridtype, bool align, uint32_t interp,
                   cudaStream_t st) {
    const dim3 grid(div_up(B, 256u), L, 1);
#define PN_GRID_CASE(CC)                                                                                        \
    case CC:                                                                                                    \
        grid_forward_generic<T, D, CC><<<grid, 256, 0, st>>>(inputs, table, offsets, outputs, B, L, S, H, dy_dx, \
                                                             gridtype, align, interp);                          \
        break;
    switch (C) {
        PN_GRID_CASE(1) PN_GRID_CASE(2) PN_GRID_CASE(4) PN_GRID_CASE(8)
        default:
            pn_set_error("GridEncoding: C must be 1, 2, 4, or 8.");
            return PN_EINVAL;
    }
#undef PN_GRID_CASE
    PN_LAUNCH_CHECK("grid_forward_generic");
    return PN_OK;
}

template <typename T>
int dispatch_D(const float *inputs, const T *table, const int *offsets, T *outputs, uint32_t B, uint32_t D, uint32_t C,
               uint32_t L, float S, uint32_t H, T *dy_dx, uint32_t gridtype, bool align, uint32_t interp,
               cudaStream_t st) {
    switch (D) {
        case 2: return launch_generic<T, 2>(inputs, table, offsets, outputs, B, C, L, S, H, dy_dx, gridtype, align, interp, st);
        case 3: return launch_generic<T, 3>(inputs, table, offsets, outputs, B, C, L, S, H, dy_dx, gridtype, align, interp, st);
        case 4: return launch_generic<T, 4>(inputs, table, offsets, outputs, B, C, L, S, H, dy_dx, gridtype, align, interp, st);
        case 5: return launch_generic<T, 5>(inputs, table, offsets, outputs, B, C, L, S, H, dy_dx, gridtype, align, interp, st);
        default:
            pn_set_error("GridEncoding: D must be 2, 3, 4 or 5.");
            return PN_EINVAL;
    }
}

}  // namespace

extern "C" int pn_grid_encode_forward(const float *inputs, const void *embeddings, const int *offsets, void *outputs,
                                      uint32_t B, uint32_t D, uint32_t C, uint32_t L, float S, uint32_t H, void *dy_dx,
                                      uint32_t gridtype, int align_corners, uint32_t interp, int emb_half,
                                      void *stream) {
    PN_REQUIRE(inputs && embeddings && offsets && outputs, "null pointer");
    PN_REQUIRE(gridtype <= 1 && interp <= 1, "gridtype/interp out of range");
    if (B == 0 || L == 0) return PN_OK;
    cudaStream_t st = PN_STREAM(stream);
    if (!emb_half && D == 3 && C == 2 && !dy_dx && !align_corners && interp == 0) {
        // persistent-ish grid: enough CTAs per level for >= 2 waves of 148 SMs x 8 resident CTAs
        const uint32_t per_level = min(div_up(B, 256u), 148u * 8u * 2u);
        grid_forward_d3c2<<<dim3(per_level, L, 1), 256, 0, st>>>(inputs, (const float2 *)embeddings, offsets,
                                                                 (float2 *)outputs, B, S, H, gridtype);
        PN_LAUNCH_CHECK("grid_forward_d3c2");
        return PN_OK;
    }
    if (emb_half)
        return dispatch_D<__half>(inputs, (const __half *)embeddings, offsets, (__half *)outputs, B, D, C, L, S, H,
                                  (__half *)dy_dx, gridtype, align_corners != 0, interp, st);
    return dispatch_D<float>(inputs, (const float *)embeddings, offsets, (float *)outputs, B, D, C, L, S, H,
                             (float *)dy_dx, gridtype, align_corners != 0, interp, st);
}
