// Real spherical-harmonics direction encoder (degree 1..8; the hot path uses 4 -> 16 outputs, forward only).
// Replaces shencoder/src/shencoder.cu:27-125,400-417 of the reference, and for training its
// input-gradient table :128-344 (sh_forward_jacobian: the same polynomials evaluated on dual numbers, one seeded
// direction per pass) and backward kernel :358-398,419-438 (sh_backward).  The Cartesian polynomial forms and
// their evaluation order are kept term-for-term so outputs are bit-comparable; one thread per direction,
// 128-bit stores when the row is 16-byte aligned.
#include "common.cuh"
#include "sh_device.cuh"

namespace {

template <uint32_t DEG>
__global__ void __launch_bounds__(256) sh_forward(const float *__restrict__ in, float *__restrict__ out, uint32_t B,
                                                  uint32_t D) {
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const float *p = in + (size_t)b * D;
    float y[DEG * DEG];
    pn::sh_eval<DEG>(p[0], p[1], p[2], y);
    float *o = out + (size_t)b * DEG * DEG;
    if constexpr ((DEG * DEG) % 4 == 0) {
        float4 *o4 = reinterpret_cast<float4 *>(o);
#pragma unroll
        for (uint32_t i = 0; i < DEG * DEG / 4; i++) o4[i] = make_float4(y[4 * i], y[4 * i + 1], y[4 * i + 2], y[4 * i + 3]);
    } else {
#pragma unroll
        for (uint32_t i = 0; i < DEG * DEG; i++) o[i] = y[i];
    }
}

// dy_dx [B, 3, DEG*DEG]: row a holds d Y / d (x,y,z)[a].  The outputs themselves come from sh_forward (launched beside this
// kernel), so they are bit-identical with and without the Jacobian.
template <uint32_t DEG>
__global__ void __launch_bounds__(128) sh_forward_jacobian(const float *__restrict__ in, float *__restrict__ dy_dx, uint32_t B,
                                                           uint32_t D, bool aligned16) {
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    constexpr uint32_t C2 = DEG * DEG;
    const float *p = in + (size_t)b * D;
    const float x = p[0], y = p[1], z = p[2];
    float *j = dy_dx + (size_t)b * D * C2;
#pragma unroll
    for (uint32_t a = 0; a < 3; a++) {
        pn::SHDual v[C2];
        pn::sh_eval<DEG, pn::SHDual>(pn::SHDual(x, a == 0 ? 1.f : 0.f), pn::SHDual(y, a == 1 ? 1.f : 0.f),
                                     pn::SHDual(z, a == 2 ? 1.f : 0.f), v);
        if (C2 % 4 == 0 && aligned16) {  // rows of the Jacobian are 16-byte aligned: 128-bit stores
            float4 *j4 = reinterpret_cast<float4 *>(j + a * C2);
#pragma unroll
            for (uint32_t i = 0; i < C2 / 4; i++) j4[i] = make_float4(v[4 * i].d, v[4 * i + 1].d, v[4 * i + 2].d, v[4 * i + 3].d);
        } else {
#pragma unroll
            for (uint32_t i = 0; i < C2; i++) j[a * C2 + i] = v[i].d;
        }
    }
}

// grad_inputs[b,d] += sum_ch grad[b,ch] * dy_dx[b,d,ch]  (accumulates into the caller's buffer, as the reference does)
__global__ void __launch_bounds__(256) sh_backward(const float *__restrict__ grad, const float *__restrict__ dy_dx,
                                                   float *__restrict__ grad_inputs, uint32_t B, uint32_t D, uint32_t C2) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= B * D) return;
    const uint32_t b = t / D;
    const float *g = grad + (size_t)b * C2, *j = dy_dx + (size_t)t * C2;
    float acc = grad_inputs[t];
    for (uint32_t ch = 0; ch < C2; ch++) acc += g[ch] * j[ch];
    grad_inputs[t] = acc;
}

}  // namespace

extern "C" int pn_sh_encode_forward(const float *inputs, float *outputs, uint32_t B, uint32_t D, uint32_t C,
                                    float *dy_dx, void *stream) {
    PN_REQUIRE(inputs && outputs, "null pointer");
    PN_REQUIRE(D >= 3, "SH encoder expects 3-component directions");
    if (B == 0) return PN_OK;
    cudaStream_t st = PN_STREAM(stream);
    if (dy_dx) {
        PN_REQUIRE(D == 3, "SH input gradients expect D == 3 (dy_dx is [B, 3, C*C])");
        const uint32_t gj = div_up(B, 128u);
        const bool al = (reinterpret_cast<uintptr_t>(dy_dx) & 15) == 0;
        switch (C) {
            case 1: sh_forward_jacobian<1><<<gj, 128, 0, st>>>(inputs, dy_dx, B, D, al); break;
            case 2: sh_forward_jacobian<2><<<gj, 128, 0, st>>>(inputs, dy_dx, B, D, al); break;
            case 3: sh_forward_jacobian<3><<<gj, 128, 0, st>>>(inputs, dy_dx, B, D, al); break;
            case 4: sh_forward_jacobian<4><<<gj, 128, 0, st>>>(inputs, dy_dx, B, D, al); break;
            case 5: sh_forward_jacobian<5><<<gj, 128, 0, st>>>(inputs, dy_dx, B, D, al); break;
            case 6: sh_forward_jacobian<6><<<gj, 128, 0, st>>>(inputs, dy_dx, B, D, al); break;
            case 7: sh_forward_jacobian<7><<<gj, 128, 0, st>>>(inputs, dy_dx, B, D, al); break;
            case 8: sh_forward_jacobian<8><<<gj, 128, 0, st>>>(inputs, dy_dx, B, D, al); break;
            default:
                pn_set_error("SH encoder: degree must be in 1..8");
                return PN_EINVAL;
        }
        PN_LAUNCH_CHECK("sh_forward_jacobian");
    }
    const uint32_t grid = div_up(B, 256u);
    switch (C) {
        case 1: sh_forward<1><<<grid, 256, 0, st>>>(inputs, outputs, B, D); break;
        case 2: sh_forward<2><<<grid, 256, 0, st>>>(inputs, outputs, B, D); break;
        case 3: sh_forward<3><<<grid, 256, 0, st>>>(inputs, outputs, B, D); break;
        case 4: sh_forward<4><<<grid, 256, 0, st>>>(inputs, outputs, B, D); break;
        case 5: sh_forward<5><<<grid, 256, 0, st>>>(inputs, outputs, B, D); break;
        case 6: sh_forward<6><<<grid, 256, 0, st>>>(inputs, outputs, B, D); break;
        case 7: sh_forward<7><<<grid, 256, 0, st>>>(inputs, outputs, B, D); break;
        case 8: sh_forward<8><<<grid, 256, 0, st>>>(inputs, outputs, B, D); break;
        default:
            pn_set_error("SH encoder: degree must be in 1..8");
            return PN_EINVAL;
    }
    PN_LAUNCH_CHECK("sh_forward");
    return PN_OK;
}

extern "C" int pn_sh_encode_backward(const float *grad, const float *inputs, uint32_t B, uint32_t D, uint32_t C,
                                     const float *dy_dx, float *grad_inputs, void *stream) {
    (void)inputs;  // the Jacobian already carries everything the inputs would; kept for the reference's signature
    PN_REQUIRE(grad && dy_dx && grad_inputs, "null pointer");
    PN_REQUIRE(C >= 1 && C <= 8, "degree must be in 1..8");
    if (B == 0 || D == 0) return PN_OK;
    sh_backward<<<div_up(B * D, 256u), 256, 0, PN_STREAM(stream)>>>(grad, dy_dx, grad_inputs, B, D, C * C);
    PN_LAUNCH_CHECK("sh_backward");
    return PN_OK;
}
