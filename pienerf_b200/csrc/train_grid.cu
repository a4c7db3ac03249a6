// Training-side hash-grid kernels for sm_100a.
// Replaces gridencoder/src/gridencoder.cu:248-343 (embedding gradients), :346-370 (input gradients), :473-503 (entry
// point), :505-645 (total-variation gradient) of the reference.
//
//  * grid_backward_kernel<T,D,C> — one thread per (sample, level) scatters w * dL/dy into the 2^D vertices of its
//    cell.  All C channels of a vertex leave in vector reductions where the hardware has them (REDG.E.ADD.F32x2 /
//    .F32x4 for fp32 pairs and quads, red.global.add.noftz.f16x2 for fp16 pairs) instead of the reference's one atomic per channel pair per thread with
//    C/2 threads recomputing the same vertex indices.  blockIdx.y = level, so co-resident CTAs reduce into the same
//    level slice of the table gradient and the reductions resolve in L2.
//  * grid_input_backward_kernel<T,D,C> — dL/dx[b,d] = sum_{l,c} dL/dy[l,b,c] * dy_dx[b,l,d,c], summed in (l,c) order.
//  * grid_tv_kernel<T,D,C> — gradient of the normalised total variation around the vertex each sample falls on.
// Sums land through atomics in both implementations, so results agree to rounding, not bit-for-bit.
#include "grid_device.cuh"

namespace {

template <typename T> __device__ __forceinline__ float to_f(T v);
template <> __device__ __forceinline__ float to_f<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f<__half>(__half v) { return __half2float(v); }
template <typename T> __device__ __forceinline__ T from_f(float v);
template <> __device__ __forceinline__ float from_f<float>(float v) { return v; }
template <> __device__ __forceinline__ __half from_f<__half>(float v) { return __float2half(v); }

// C contiguous channels of one vertex: vector reductions where available.
template <uint32_t C>
__device__ __forceinline__ void reduce_vertex(float *dst, const float (&g)[C], float w) {
    if constexpr (C % 4 == 0) {
#pragma unroll
        for (uint32_t c = 0; c < C; c += 4)
            atomicAdd(reinterpret_cast<float4 *>(dst + c), make_float4(w * g[c], w * g[c + 1], w * g[c + 2], w * g[c + 3]));
    } else if constexpr (C % 2 == 0) {
#pragma unroll
        for (uint32_t c = 0; c < C; c += 2) atomicAdd(reinterpret_cast<float2 *>(dst + c), make_float2(w * g[c], w * g[c + 1]));
    } else {
#pragma unroll
        for (uint32_t c = 0; c < C; c++) atomicAdd(dst + c, w * g[c]);
    }
}
// fp16 tables: the table gradient is global memory, so say so (the generic atomicAdd overloads compile to an
// address-space check plus a shared-memory CAS loop beside every reduction)
__device__ __forceinline__ void red_f16x2(__half *dst, float a, float b) {
    const __half2 v = __halves2half2(__float2half(a), __float2half(b));
    asm volatile("red.global.add.noftz.f16x2 [%0], %1;" ::"l"(dst), "r"(*reinterpret_cast<const uint32_t *>(&v)) : "memory");
}
__device__ __forceinline__ void red_f16(__half *dst, float a) {
    const __half v = __float2half(a);
    asm volatile("red.global.add.noftz.f16 [%0], %1;" ::"l"(dst), "h"(*reinterpret_cast<const uint16_t *>(&v)) : "memory");
}
template <uint32_t C>
__device__ __forceinline__ void reduce_vertex(__half *dst, const float (&g)[C], float w) {
    if constexpr (C % 2 == 0) {
#pragma unroll
        for (uint32_t c = 0; c < C; c += 2) red_f16x2(dst + c, w * g[c], w * g[c + 1]);
    } else {
#pragma unroll
        for (uint32_t c = 0; c < C; c++) red_f16(dst + c, w * g[c]);
    }
}
__device__ __forceinline__ void tv_add(float *dst, float v) { atomicAdd(dst, v); }
__device__ __forceinline__ void tv_add(__half *dst, float v) { red_f16(dst, v); }

template <typename T, uint32_t D, uint32_t C>
__global__ void __launch_bounds__(256) grid_backward_kernel(const T *__restrict__ grad, const float *__restrict__ inputs,
                                                            const int *__restrict__ offsets, T *__restrict__ grad_grid,
                                                            uint32_t B, float S, uint32_t H, uint32_t gridtype,
                                                            bool align_corners, uint32_t interp) {
    const uint32_t level = blockIdx.y;
    const pn::LevelGeom geo = pn::level_geom(level, S, H, offsets, align_corners);
    T *tab = grad_grid + (size_t)(uint32_t)offsets[level] * C;
    for (uint32_t b = blockIdx.x * blockDim.x + threadIdx.x; b < B; b += gridDim.x * blockDim.x) {
        const float *in = inputs + (size_t)b * D;
        float frac[D];
        uint32_t cell[D];
        bool oob = false;
#pragma unroll
        for (uint32_t d = 0; d < D; d++) {
            const float x = in[d];
            oob |= (x < 0 || x > 1);
            float p = x * geo.scale + (align_corners ? 0.0f : 0.5f);
            const float fl = floorf(p);
            cell[d] = (uint32_t)fl;
            p -= (float)cell[d];
            frac[d] = interp == 1 ? p * p * (3.0f - 2.0f * p) : p;
        }
        if (oob) continue;  // out-of-range samples produced zeros in the forward pass: no gradient
        float g[C];
        const T *gp = grad + ((size_t)level * B + b) * C;
#pragma unroll
        for (uint32_t c = 0; c < C; c++) g[c] = to_f<T>(gp[c]);
#pragma unroll
        for (uint32_t corner = 0; corner < (1u << D); corner++) {
            float w = 1;
            uint32_t v[D];
#pragma unroll
            for (uint32_t d = 0; d < D; d++) {
                if (corner & (1u << d)) { w *= frac[d]; v[d] = cell[d] + 1; }
                else { w *= 1 - frac[d]; v[d] = cell[d]; }
            }
            reduce_vertex<C>(tab + (size_t)pn::vertex_index<D>(v, geo, gridtype) * C, g, w);
        }
    }
}

// at::Half arithmetic of the reference rounds the product and the running sum to fp16 separately
template <typename T> __device__ __forceinline__ void mac(T &acc, T a, T b);
template <> __device__ __forceinline__ void mac<float>(float &acc, float a, float b) { acc += a * b; }
template <> __device__ __forceinline__ void mac<__half>(__half &acc, __half a, __half b) {
    acc = __float2half(__half2float(acc) + __half2float(__float2half(__half2float(a) * __half2float(b))));
}

template <typename T, uint32_t D, uint32_t C>
__global__ void __launch_bounds__(256) grid_input_backward_kernel(const T *__restrict__ grad, const T *__restrict__ dy_dx,
                                                                  T *__restrict__ grad_inputs, uint32_t B, uint32_t L) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= B * D) return;
    const uint32_t b = t / D, d = t - b * D;
    const T *j = dy_dx + (size_t)b * L * D * C + d * C;
    T acc = from_f<T>(0.f);
    for (uint32_t l = 0; l < L; l++) {
        const T *g = grad + ((size_t)l * B + b) * C;
#pragma unroll
        for (uint32_t c = 0; c < C; c++) mac<T>(acc, g[c], j[(size_t)l * D * C + c]);
    }
    grad_inputs[t] = acc;
}

template <typename T, uint32_t D, uint32_t C>
__global__ void __launch_bounds__(256) grid_tv_kernel(const T *__restrict__ inputs, const T *__restrict__ grid,
                                                      T *__restrict__ grad, const int *__restrict__ offsets, float weight,
                                                      uint32_t B, float S, uint32_t H, uint32_t gridtype,
                                                      bool align_corners) {
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const uint32_t level = blockIdx.y;
    const pn::LevelGeom geo = pn::level_geom(level, S, H, offsets, align_corners);
    const T *tab = grid + (size_t)(uint32_t)offsets[level] * C;
    T *gtab = grad + (size_t)(uint32_t)offsets[level] * C;
    uint32_t v[D];
    bool oob = false;
#pragma unroll
    for (uint32_t d = 0; d < D; d++) {
        const float x = to_f<T>(inputs[(size_t)b * D + d]);
        oob |= (x < 0 || x > 1);
        v[d] = (uint32_t)floorf(x * geo.scale + (align_corners ? 0.0f : 0.5f));
    }
    if (oob) return;
    const uint32_t e = pn::vertex_index<D>(v, geo, gridtype) * C;
    float here[C], sum[C], sq[C];
#pragma unroll
    for (uint32_t c = 0; c < C; c++) { here[c] = to_f<T>(tab[e + c]); sum[c] = 0; sq[c] = 0; }
#pragma unroll
    for (uint32_t d = 0; d < D; d++) {
        const uint32_t cur = v[d];
        if (cur < geo.resolution) {  // neighbour on the + side
            v[d] = cur + 1;
            const uint32_t er = pn::vertex_index<D>(v, geo, gridtype) * C;
#pragma unroll
            for (uint32_t c = 0; c < C; c++) {
                const float diff = here[c] - to_f<T>(tab[er + c]);
                sum[c] += diff;
                sq[c] += diff * diff;
            }
        }
        if (cur > 0) {  // neighbour on the - side
            v[d] = cur - 1;
            const uint32_t el = pn::vertex_index<D>(v, geo, gridtype) * C;
#pragma unroll
            for (uint32_t c = 0; c < C; c++) {
                const float diff = here[c] - to_f<T>(tab[el + c]);
                sum[c] += diff;
                sq[c] += diff * diff;
            }
        }
        v[d] = cur;
    }
    const float w = weight / (2 * D);
#pragma unroll
    for (uint32_t c = 0; c < C; c++) tv_add(gtab + e + c, w * sum[c] * rsqrtf(sq[c] + 1e-9f));
}

template <typename T, uint32_t D, uint32_t C>
int launch_backward(const T *grad, const float *inputs, const int *offsets, T *grad_emb, uint32_t B, uint32_t L, float S,
                    uint32_t H, const T *dy_dx, T *grad_inputs, uint32_t gridtype, bool align, uint32_t interp,
                    cudaStream_t st) {
    const uint32_t per_level = min(div_up(B, 256u), 148u * 8u * 2u);
    grid_backward_kernel<T, D, C><<<dim3(per_level, L, 1), 256, 0, st>>>(grad, inputs, offsets, grad_emb, B, S, H, gridtype,
                                                                        align, interp);
    PN_LAUNCH_CHECK("grid_backward_kernel");
    if (dy_dx) {
        grid_input_backward_kernel<T, D, C><<<div_up(B * D, 256u), 256, 0, st>>>(grad, dy_dx, grad_inputs, B, L);
        PN_LAUNCH_CHECK("grid_input_backward_kernel");
    }
    return PN_OK;
}

template <typename T, uint32_t D, uint32_t C>
int launch_tv(const T *inputs, const T *emb, T *grad, const int *offsets, float weight, uint32_t B, uint32_t L, float S,
              uint32_t H, uint32_t gridtype, bool align, cudaStream_t st) {
    grid_tv_kernel<T, D, C><<<dim3(div_up(B, 256u), L, 1), 256, 0, st>>>(inputs, emb, grad, offsets, weight, B, S, H, gridtype,
                                                                        align);
    PN_LAUNCH_CHECK("grid_tv_kernel");
    return PN_OK;
}

#define PN_DC_SWITCH(CALL)                                                                     \
    switch (D * 16 + C) {                                                                      \
        case 2 * 16 + 1: return CALL(2, 1); case 2 * 16 + 2: return CALL(2, 2);                \
        case 2 * 16 + 4: return CALL(2, 4); case 2 * 16 + 8: return CALL(2, 8);                \
        case 3 * 16 + 1: return CALL(3, 1); case 3 * 16 + 2: return CALL(3, 2);                \
        case 3 * 16 + 4: return CALL(3, 4); case 3 * 16 + 8: return CALL(3, 8);                \
        case 4 * 16 + 1: return CALL(4, 1); case 4 * 16 + 2: return CALL(4, 2);                \
        case 4 * 16 + 4: return CALL(4, 4); case 4 * 16 + 8: return CALL(4, 8);                \
        case 5 * 16 + 1: return CALL(5, 1); case 5 * 16 + 2: return CALL(5, 2);                \
        case 5 * 16 + 4: return CALL(5, 4); case 5 * 16 + 8: return CALL(5, 8);                \
        default:                                                                               \
            pn_set_error("GridEncoding: D must be 2..5 and C must be 1, 2, 4, or 8.");         \
            return PN_EINVAL;                                                                  \
    }

template <typename T>
int backward_dispatch(const void *grad, const float *inputs, const int *offsets, void *grad_emb, uint32_t B, uint32_t D,
                      uint32_t C, uint32_t L, float S, uint32_t H, const void *dy_dx, void *grad_inputs, uint32_t gridtype,
                      bool align, uint32_t interp, cudaStream_t st) {
#define PN_BW(DD, CC) launch_backward<T, DD, CC>((const T *)grad, inputs, offsets, (T *)grad_emb, B, L, S, H, (const T *)dy_dx, \
                                                 (T *)grad_inputs, gridtype, align, interp, st)
    PN_DC_SWITCH(PN_BW)
#undef PN_BW
}

template <typename T>
int tv_dispatch(const void *inputs, const void *emb, void *grad, const int *offsets, float weight, uint32_t B, uint32_t D,
                uint32_t C, uint32_t L, float S, uint32_t H, uint32_t gridtype, bool align, cudaStream_t st) {
#define PN_TV(DD, CC) launch_tv<T, DD, CC>((const T *)inputs, (const T *)emb, (T *)grad, offsets, weight, B, L, S, H, gridtype, align, st)
    PN_DC_SWITCH(PN_TV)
#undef PN_TV
}

}  // namespace

extern "C" int pn_grid_encode_backward(const void *grad, const float *inputs, const void *embeddings, const int *offsets,
                                       void *grad_embeddings, uint32_t B, uint32_t D, uint32_t C, uint32_t L, float S,
                                       uint32_t H, const void *dy_dx, void *grad_inputs, uint32_t gridtype,
                                       int align_corners, uint32_t interp, int emb_half, void *stream) {
    (void)embeddings;  // the table values do not enter its own gradient; kept for the reference's signature
    PN_REQUIRE(grad && inputs && offsets && grad_embeddings, "null pointer");
    PN_REQUIRE(!dy_dx || grad_inputs, "dy_dx given without grad_inputs");
    PN_REQUIRE(gridtype <= 1 && interp <= 1, "gridtype/interp out of range");
    if (B == 0 || L == 0) return PN_OK;
    cudaStream_t st = PN_STREAM(stream);
    if (emb_half)
        return backward_dispatch<__half>(grad, inputs, offsets, grad_embeddings, B, D, C, L, S, H, dy_dx, grad_inputs, gridtype,
                                         align_corners != 0, interp, st);
    return backward_dispatch<float>(grad, inputs, offsets, grad_embeddings, B, D, C, L, S, H, dy_dx, grad_inputs, gridtype,
                                    align_corners != 0, interp, st);
}

extern "C" int pn_grad_total_variation(const void *inputs, const void *embeddings, void *grad, const int *offsets,
                                       float weight, uint32_t B, uint32_t D, uint32_t C, uint32_t L, float S, uint32_t H,
                                       uint32_t gridtype, int align_corners, int emb_half, void *stream) {
    PN_REQUIRE(inputs && embeddings && grad && offsets, "null pointer");
    PN_REQUIRE(gridtype <= 1, "gridtype out of range");
    if (B == 0 || L == 0) return PN_OK;
    cudaStream_t st = PN_STREAM(stream);
    if (emb_half)
        return tv_dispatch<__half>(inputs, embeddings, grad, offsets, weight, B, D, C, L, S, H, gridtype, align_corners != 0, st);
    return tv_dispatch<float>(inputs, embeddings, grad, offsets, weight, B, D, C, L, S, H, gridtype, align_corners != 0, st);
}
