// Real spherical-harmonics direction encoder, forward (degree 1..8; the hot path uses 4 -> 16 outputs).
// Replaces shencoder/src/shencoder.cu:27-125,400-417 of the reference.  The Cartesian polynomial forms and
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

}  // namespace

extern "C" int pn_sh_encode_forward(const float *inputs, float *outputs, uint32_t B, uint32_t D, uint32_t C,
                                    float *dy_dx, void *stream) {
    PN_REQUIRE(inputs && outputs, "null pointer");
    PN_REQUIRE(D >= 3, "SH encoder expects 3-component directions");
    if (dy_dx) {
        pn_set_error("sh_encode_forward: dy_dx (input gradients) is training-only, outside the hot path");
        return PN_ENOTIMPL;
    }
    if (B == 0) return PN_OK;
    cudaStream_t st = PN_STREAM(stream);
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

extern "C" int pn_sh_encode_backward(void) {
    pn_set_error("sh_encode_backward is training-only and outside the B200 hot path (SURVEY.md 8f.4)");
    return PN_ENOTIMPL;
}
