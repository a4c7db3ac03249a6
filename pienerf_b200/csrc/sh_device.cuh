// Real spherical harmonics Y_l^m(x,y,z) in Cartesian polynomial form, bands l = 0..DEG-1, written into
// y[DEG*DEG] in the reference's output order (shencoder/src/shencoder.cu:49-123).  Term order inside each
// polynomial matches the reference so fp32 results agree bit-for-bit under the same FMA contraction.
//
// sh_eval is a template over the scalar type: with T = float it is the forward encoder; with T = SHDual (a value and
// one directional derivative, forward-mode differentiation) the same polynomials yield d Y / d x, d Y / d y, d Y / d z
// (shencoder.cu:128-344 of the reference spells those 3 x 64 derivatives out by hand).
#pragma once

namespace pn {

struct SHDual {
    float v, d;  // value, derivative along the seeded direction
    __device__ __forceinline__ SHDual() {}
    __device__ __forceinline__ SHDual(float c) : v(c), d(0.f) {}
    __device__ __forceinline__ SHDual(float vv, float dd) : v(vv), d(dd) {}
};
__device__ __forceinline__ SHDual operator+(SHDual a, SHDual b) { return SHDual(a.v + b.v, a.d + b.d); }
__device__ __forceinline__ SHDual operator-(SHDual a, SHDual b) { return SHDual(a.v - b.v, a.d - b.d); }
__device__ __forceinline__ SHDual operator-(SHDual a) { return SHDual(-a.v, -a.d); }
__device__ __forceinline__ SHDual operator*(SHDual a, SHDual b) { return SHDual(a.v * b.v, a.d * b.v + a.v * b.d); }
__device__ __forceinline__ SHDual operator+(float a, SHDual b) { return SHDual(a + b.v, b.d); }
__device__ __forceinline__ SHDual operator+(SHDual a, float b) { return SHDual(a.v + b, a.d); }
__device__ __forceinline__ SHDual operator-(float a, SHDual b) { return SHDual(a - b.v, -b.d); }
__device__ __forceinline__ SHDual operator-(SHDual a, float b) { return SHDual(a.v - b, a.d); }
__device__ __forceinline__ SHDual operator*(float a, SHDual b) { return SHDual(a * b.v, a * b.d); }
__device__ __forceinline__ SHDual operator*(SHDual a, float b) { return SHDual(a.v * b, a.d * b); }

template <unsigned DEG, typename T = float>
__device__ __forceinline__ void sh_eval(T x, T y, T z, T *__restrict__ o) {
    const T xy = x * y, xz = x * z, yz = y * z, x2 = x * x, y2 = y * y, z2 = z * z;
    o[0] = 0.28209479177387814f;
    if constexpr (DEG > 1) {
        o[1] = -0.48860251190291987f * y;
        o[2] = 0.48860251190291987f * z;
        o[3] = -0.48860251190291987f * x;
    }
    if constexpr (DEG > 2) {
        o[4] = 1.0925484305920792f * xy;
        o[5] = -1.0925484305920792f * yz;
        o[6] = 0.94617469575755997f * z2 - 0.31539156525251999f;
        o[7] = -1.0925484305920792f * xz;
        o[8] = 0.54627421529603959f * x2 - 0.54627421529603959f * y2;
    }
    if constexpr (DEG > 3) {
        o[9] = 0.59004358992664352f * y * (-3.0f * x2 + y2);
        o[10] = 2.8906114426405538f * xy * z;
        o[11] = 0.45704579946446572f * y * (1.0f - 5.0f * z2);
        o[12] = 0.3731763325901154f * z * (5.0f * z2 - 3.0f);
        o[13] = 0.45704579946446572f * x * (1.0f - 5.0f * z2);
        o[14] = 1.4453057213202769f * z * (x2 - y2);
        o[15] = 0.59004358992664352f * x * (-x2 + 3.0f * y2);
    }
    if constexpr (DEG > 4) {
        const T x4 = x2 * x2, y4 = y2 * y2, z4 = z2 * z2;
        o[16] = 2.5033429417967046f * xy * (x2 - y2);
        o[17] = 1.7701307697799304f * yz * (-3.0f * x2 + y2);
        o[18] = 0.94617469575756008f * xy * (7.0f * z2 - 1.0f);
        o[19] = 0.66904654355728921f * yz * (3.0f - 7.0f * z2);
        o[20] = -3.1735664074561294f * z2 + 3.7024941420321507f * z4 + 0.31735664074561293f;
        o[21] = 0.66904654355728921f * xz * (3.0f - 7.0f * z2);
        o[22] = 0.47308734787878004f * (x2 - y2) * (7.0f * z2 - 1.0f);
        o[23] = 1.7701307697799304f * xz * (-x2 + 3.0f * y2);
        o[24] = -3.7550144126950569f * x2 * y2 + 0.62583573544917614f * x4 + 0.62583573544917614f * y4;
        if constexpr (DEG > 5) {
            o[25] = 0.65638205684017015f * y * (10.0f * x2 * y2 - 5.0f * x4 - y4);
            o[26] = 8.3026492595241645f * xy * z * (x2 - y2);
            o[27] = -0.48923829943525038f * y * (3.0f * x2 - y2) * (9.0f * z2 - 1.0f);
            o[28] = 4.7935367849733241f * xy * z * (3.0f * z2 - 1.0f);
            o[29] = 0.45294665119569694f * y * (14.0f * z2 - 21.0f * z4 - 1.0f);
            o[30] = 0.1169503224534236f * z * (-70.0f * z2 + 63.0f * z4 + 15.0f);
            o[31] = 0.45294665119569694f * x * (14.0f * z2 - 21.0f * z4 - 1.0f);
            o[32] = 2.3967683924866621f * z * (x2 - y2) * (3.0f * z2 - 1.0f);
            o[33] = -0.48923829943525038f * x * (x2 - 3.0f * y2) * (9.0f * z2 - 1.0f);
            o[34] = 2.0756623148810411f * z * (-6.0f * x2 * y2 + x4 + y4);
            o[35] = 0.65638205684017015f * x * (10.0f * x2 * y2 - x4 - 5.0f * y4);
        }
        if constexpr (DEG > 6) {
            const T x6 = x4 * x2, y6 = y4 * y2, z6 = z4 * z2;
            o[36] = 1.3663682103838286f * xy * (-10.0f * x2 * y2 + 3.0f * x4 + 3.0f * y4);
            o[37] = 2.3666191622317521f * yz * (10.0f * x2 * y2 - 5.0f * x4 - y4);
            o[38] = 2.0182596029148963f * xy * (x2 - y2) * (11.0f * z2 - 1.0f);
            o[39] = -0.92120525951492349f * yz * (3.0f * x2 - y2) * (11.0f * z2 - 3.0f);
            o[40] = 0.92120525951492349f * xy * (-18.0f * z2 + 33.0f * z4 + 1.0f);
            o[41] = 0.58262136251873131f * yz * (30.0f * z2 - 33.0f * z4 - 5.0f);
            o[42] = 6.6747662381009842f * z2 - 20.024298714302954f * z4 + 14.684485723822165f * z6 - 0.31784601133814211f;
            o[43] = 0.58262136251873131f * xz * (30.0f * z2 - 33.0f * z4 - 5.0f);
            o[44] = 0.46060262975746175f * (x2 - y2) * (11.0f * z2 * (3.0f * z2 - 1.0f) - 7.0f * z2 + 1.0f);
            o[45] = -0.92120525951492349f * xz * (x2 - 3.0f * y2) * (11.0f * z2 - 3.0f);
            o[46] = 0.50456490072872406f * (11.0f * z2 - 1.0f) * (-6.0f * x2 * y2 + x4 + y4);
            o[47] = 2.3666191622317521f * xz * (10.0f * x2 * y2 - x4 - 5.0f * y4);
            o[48] = 10.247761577878714f * x2 * y4 - 10.247761577878714f * x4 * y2 + 0.6831841051919143f * x6 - 0.6831841051919143f * y6;
            if constexpr (DEG > 7) {
                o[49] = 0.70716273252459627f * y * (-21.0f * x2 * y4 + 35.0f * x4 * y2 - 7.0f * x6 + y6);
                o[50] = 5.2919213236038001f * xy * z * (-10.0f * x2 * y2 + 3.0f * x4 + 3.0f * y4);
                o[51] = -0.51891557872026028f * y * (13.0f * z2 - 1.0f) * (-10.0f * x2 * y2 + 5.0f * x4 + y4);
                o[52] = 4.1513246297620823f * xy * z * (x2 - y2) * (13.0f * z2 - 3.0f);
                o[53] = -0.15645893386229404f * y * (3.0f * x2 - y2) * (13.0f * z2 * (11.0f * z2 - 3.0f) - 27.0f * z2 + 3.0f);
                o[54] = 0.44253269244498261f * xy * z * (-110.0f * z2 + 143.0f * z4 + 15.0f);
                o[55] = 0.090331607582517306f * y * (-135.0f * z2 + 495.0f * z4 - 429.0f * z6 + 5.0f);
                o[56] = 0.068284276912004949f * z * (315.0f * z2 - 693.0f * z4 + 429.0f * z6 - 35.0f);
                o[57] = 0.090331607582517306f * x * (-135.0f * z2 + 495.0f * z4 - 429.0f * z6 + 5.0f);
                o[58] = 0.07375544874083044f * z * (x2 - y2) * (143.0f * z2 * (3.0f * z2 - 1.0f) - 187.0f * z2 + 45.0f);
                o[59] = -0.15645893386229404f * x * (x2 - 3.0f * y2) * (13.0f * z2 * (11.0f * z2 - 3.0f) - 27.0f * z2 + 3.0f);
                o[60] = 1.0378311574405206f * z * (13.0f * z2 - 3.0f) * (-6.0f * x2 * y2 + x4 + y4);
                o[61] = -0.51891557872026028f * x * (13.0f * z2 - 1.0f) * (-10.0f * x2 * y2 + x4 + 5.0f * y4);
                o[62] = 2.6459606618019f * z * (15.0f * x2 * y4 - 15.0f * x4 * y2 + x6 - y6);
                o[63] = 0.70716273252459627f * x * (-35.0f * x2 * y4 + 21.0f * x4 * y2 - x6 + 7.0f * y6);
            }
        }
    }
}

}  // namespace pn
