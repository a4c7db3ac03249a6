// Tensor-core (tcgen05 / TMEM) evaluation of the sigma/colour MLP of nerf/network.py:98-127 for 128-sample tiles.
//
// One "tile group" = 4 warps = 128 threads = 128 samples (thread r owns row r of every activation matrix and TMEM
// lane r).  Per layer: every thread writes its activation row into shared memory as bf16 hi/lo parts in the UMMA
// canonical K-major no-swizzle layout (chunk-major: 16-byte k-chunks of all 128 rows are contiguous, so the 128
// threads store 2 KB contiguous — conflict-free), one elected thread issues tcgen05.mma (M=128, N=64 or 16, K=16 per
// instruction) with the accumulator in TMEM, tcgen05.commit signals an mbarrier, and each warp reads its 32 TMEM
// lanes back with tcgen05.ld for the ReLU / exp / sigmoid epilogue.
//
// Precision: the reference MLP is fp32 SGEMM with TF32 off and the parity bar is 1e-3 absolute RGB with
// sigma = exp(h) amplifying errors, so single-pass bf16 is not enough.  Each GEMM is computed as three bf16 MMAs
//   A W^T ~= A_lo W_hi^T + A_hi W_lo^T + A_hi W_hi^T      (fp32 accumulate in TMEM)
// which keeps ~16 mantissa bits per operand (dropped term <= 2^-18 relative).
#pragma once
#include <cuda_bf16.h>
#include "grid_device.cuh"
#include "sh_device.cuh"

namespace pn {
namespace tc {

constexpr int kTile = 128;             // rows per tile group
constexpr uint32_t kTmemCols = 128;    // per tile group: two 64-column fp32 accumulators (ping-pong)

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---- mbarrier
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "LAB_WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra LAB_DONE_%=;\n"
        "bra LAB_WAIT_%=;\n"
        "LAB_DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity)
        : "memory");
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void group_sync(int group) {  // named barrier 1+group over the group's 128 threads
    asm volatile("bar.sync %0, 128;" ::"r"(group + 1) : "memory");
}

// ---- TMEM
__device__ __forceinline__ void tmem_alloc(uint32_t *dst_smem, uint32_t ncols) {  // one full warp
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {    // same warp that allocated
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// one lane of a converged warp (the compiler then keeps MMA descriptors in uniform registers: no per-thread waterfall)
__device__ __forceinline__ bool elect_one() {
    uint32_t p;
    asm volatile("{\n.reg .pred p;\nelect.sync _|p, 0xffffffff;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(p));
    return p != 0;
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// 16 consecutive fp32 columns of this thread's TMEM lane
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; i++) v[i] = __uint_as_float(r[i]);
}

// ---- UMMA descriptors (cute/arch/mma_sm100_desc.hpp bit layouts)
// K-major, no swizzle: 8x(16 B) core matrices; LBO = byte distance between the two 16-byte k-chunks of one MMA,
// SBO = byte distance between 8-row groups.  version = 1 (sm_100), layout_type = 0 (SWIZZLE_NONE / interleave).
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}
// kind::f16, A = B = bf16, D = fp32, both K-major, M = 128
__host__ __device__ constexpr uint32_t instr_desc_bf16(uint32_t N) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((N >> 3) << 17) | ((128u >> 4) << 24);
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// ---- shared-memory images
// Weight matrix [N,K] (nn.Linear [out,in]) as B operand: element (n,k) of part p at  p*N*K*2 + (k/8)*(N*16) + n*16 + (k%8)*2.
struct __align__(128) Weights {
    __nv_bfloat16 w1[2][64 * 32];   // sigma_net[0]   N=64 K=32   [hi|lo]
    __nv_bfloat16 w2[2][16 * 64];   // sigma_net[1]   N=16 K=64
    __nv_bfloat16 w3[2][64 * 32];   // color_net[0]   N=64 K=32 (input 31 zero-padded)
    __nv_bfloat16 w4[2][64 * 64];   // color_net[1]   N=64 K=64
    __nv_bfloat16 w5[2][16 * 64];   // color_net[2]   N=16 (3 used) K=64
    LevelGeom geo[16];
    uint32_t level_off[16];
    uint32_t fast;                  // every level is dense or a power-of-two hash table (encode_rows_ilp applies)
};
// Activations of one tile group: [128 rows, 64 k] bf16, hi and lo images, chunk-major: (r,k) at (k/8)*2048 + r*16 + (k%8)*2
struct __align__(128) TileSmem {
    __nv_bfloat16 a[2][kTile * 64];
    uint64_t bar;
    uint32_t tmem;
    uint32_t pad;
};

__device__ __forceinline__ void split_bf16(float x, __nv_bfloat16 &hi, __nv_bfloat16 &lo) {
    hi = __float2bfloat16_rn(x);
    lo = __float2bfloat16_rn(x - __bfloat162float(hi));
}

// whole block cooperates; W row-major [N_src, K_src] fp32 in global, zero-padded to [N,K]
template <int N, int K>
__device__ __forceinline__ void load_weight(__nv_bfloat16 *hi, __nv_bfloat16 *lo, const float *__restrict__ W, int n_src, int k_src) {
    for (int i = threadIdx.x; i < N * K; i += blockDim.x) {
        const int n = i / K, k = i % K;
        const float v = (n < n_src && k < k_src) ? __ldg(W + n * k_src + k) : 0.f;
        __nv_bfloat16 h, l;
        split_bf16(v, h, l);
        const int o = (k >> 3) * (N * 8) + n * 8 + (k & 7);
        hi[o] = h; lo[o] = l;
    }
}

__device__ __forceinline__ void weights_fill(Weights &w, const pn_field_t &f) {
    load_weight<64, 32>(w.w1[0], w.w1[1], f.w_sigma0, 64, 32);
    load_weight<16, 64>(w.w2[0], w.w2[1], f.w_sigma1, 16, 64);
    load_weight<64, 32>(w.w3[0], w.w3[1], f.w_color0, 64, 31);
    load_weight<64, 64>(w.w4[0], w.w4[1], f.w_color1, 64, 64);
    load_weight<16, 64>(w.w5[0], w.w5[1], f.w_color2, 3, 64);
    if (threadIdx.x < 16) {
        w.geo[threadIdx.x] = level_geom(threadIdx.x, f.S, f.H, f.offsets, false);
        w.level_off[threadIdx.x] = (uint32_t)f.offsets[threadIdx.x];
    }
    if (threadIdx.x == 0) {
        uint32_t fast = 1;
        for (int l = 0; l < 16; l++) {
            const LevelGeom g = level_geom(l, f.S, f.H, f.offsets, false);
            if (!g.dense3 && !g.mask) fast = 0;
        }
        w.fast = fast;
    }
}

// bf16 hi/lo split of two values at once: hi = {rn(a), rn(b)}, lo = {rn(a - hi_a), rn(b - hi_b)} (a in the low half =
// the lower k index).  One packed convert per word instead of one per value; bit-identical to split_bf16.
__device__ __forceinline__ void split_pair(float a, float b, uint32_t &hi, uint32_t &lo) {
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(hi) : "f"(b), "f"(a));
    const float ha = __uint_as_float(hi << 16), hb = __uint_as_float(hi & 0xffff0000u);
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(lo) : "f"(b - hb), "f"(a - ha));
}

// store 8 consecutive k-values (one 16-byte chunk) of this thread's row into the hi and lo activation images
__device__ __forceinline__ void store_chunk(TileSmem &t, int row, int chunk, const float (&v)[8]) {
    uint4 h, l;
    split_pair(v[0], v[1], h.x, l.x); split_pair(v[2], v[3], h.y, l.y);
    split_pair(v[4], v[5], h.z, l.z); split_pair(v[6], v[7], h.w, l.w);
    *reinterpret_cast<uint4 *>(reinterpret_cast<char *>(t.a[0]) + chunk * 2048 + row * 16) = h;
    *reinterpret_cast<uint4 *>(reinterpret_cast<char *>(t.a[1]) + chunk * 2048 + row * 16) = l;
}

// D[128,N] (TMEM columns tmem_d..) = A[128,K] W[N,K]^T with the 3-term bf16 split.  Called by ONE thread of the group.
// Rolled loops with incremental descriptors: this runs on a single thread, code size matters more than issue rate.
template <int N, int K>
__device__ __forceinline__ void issue_layer(const void *a_hi_img, const void *a_lo_img, const __nv_bfloat16 *w_hi, const __nv_bfloat16 *w_lo,
                                            uint32_t tmem_d) {
    constexpr uint32_t idesc = instr_desc_bf16(N);
    const uint64_t a_hi = smem_desc(smem_u32(a_hi_img), 2048, 128), a_lo = smem_desc(smem_u32(a_lo_img), 2048, 128);
    const uint64_t b_hi = smem_desc(smem_u32(w_hi), N * 16, 128), b_lo = smem_desc(smem_u32(w_lo), N * 16, 128);
    uint32_t acc = 0;
#pragma unroll 1
    for (int term = 0; term < 3; term++) {  // small terms first, hi*hi last
        uint64_t ad = term == 0 ? a_lo : a_hi;
        uint64_t bd = term == 1 ? b_lo : b_hi;
#pragma unroll 1
        for (int ks = 0; ks < K / 16; ks++) {
            umma_bf16(tmem_d, ad, bd, idesc, acc);
            acc = 1;
            ad += (2 * 2048) >> 4;          // start-address field counts 16-byte units: next 16 k-values = 2 chunks
            bd += (2 * N * 16) >> 4;
        }
    }
}

// ReLU epilogue of a 64-wide hidden layer: TMEM columns [col, col+64) of this thread's lane -> bf16 hi/lo activation row.
// Deliberately not inlined: it is used three times per tile and the frame kernel is instruction-fetch sensitive.
static __device__ __noinline__ void epilogue_relu64(TileSmem &t, int row, uint32_t taddr) {
#pragma unroll 1
    for (int q = 0; q < 4; q++) {
        float v[16], c8[8];
        tmem_ld16(taddr + q * 16, v);
#pragma unroll
        for (int hh = 0; hh < 2; hh++) {
#pragma unroll
            for (int i = 0; i < 8; i++) c8[i] = fmaxf(v[hh * 8 + i], 0.f);
            store_chunk(t, row, q * 2 + hh, c8);
        }
    }
}

// Full MLP for one tile.  Preconditions: this thread's 32 encoded features are already stored (chunks 0..3 of t.a);
// `sh` = SH(4) of the sample's ray direction.  `phase` is the group's running mbarrier parity (updated).
// All 128 threads of the group must call this together.
// `a0_hi/a0_lo`: the [128,32] bf16 hi/lo images sigma_net[0] reads (the tile's own t.a, or a stage of a producer ring);
// `release` (optional): mbarrier that is arrived on once the layer-1 MMAs have finished reading those images.
__device__ __forceinline__ void mlp_tile(TileSmem &t, const Weights &w, int group, int row, const float (&sh)[16], uint32_t &phase,
                                         float &sigma, float &r, float &g, float &b, const void *a0_hi, const void *a0_lo,
                                         uint64_t *release = nullptr) {
    const uint32_t tm = t.tmem + ((uint32_t)(row & ~31) << 16);  // this warp's 32 TMEM lanes
    const bool warp0 = (row >> 5) == 0;  // warp-uniform: the group's first warp issues the MMAs through one elected lane
    float v[16], c8[8];

    // ---- sigma_net[0]: [128,32] x [64,32]^T -> cols 0..63
    fence_async_smem();
    tc_fence_before();
    group_sync(group);
    if (warp0) {
        tc_fence_after();
        if (elect_one()) { issue_layer<64, 32>(a0_hi, a0_lo, w.w1[0], w.w1[1], t.tmem); if (release) umma_commit(release); umma_commit(&t.bar); }
        __syncwarp();
    }
    mbar_wait(&t.bar, phase); phase ^= 1;
    tc_fence_after();
    epilogue_relu64(t, row, tm);
    // ---- sigma_net[1]: [128,64] x [16,64]^T -> cols 64..79
    fence_async_smem();
    tc_fence_before();
    group_sync(group);
    if (warp0) { tc_fence_after(); if (elect_one()) { issue_layer<16, 64>(t.a[0], t.a[1], w.w2[0], w.w2[1], t.tmem + 64); umma_commit(&t.bar); } __syncwarp(); }
    mbar_wait(&t.bar, phase); phase ^= 1;
    tc_fence_after();
    tmem_ld16(tm + 64, v);
    sigma = expf(v[0]);
    // colour-net input row: SH(16) | geo(15) | 0
#pragma unroll
    for (int i = 0; i < 8; i++) c8[i] = sh[i];
    store_chunk(t, row, 0, c8);
#pragma unroll
    for (int i = 0; i < 8; i++) c8[i] = sh[8 + i];
    store_chunk(t, row, 1, c8);
#pragma unroll
    for (int i = 0; i < 8; i++) c8[i] = v[1 + i];
    store_chunk(t, row, 2, c8);
#pragma unroll
    for (int i = 0; i < 7; i++) c8[i] = v[9 + i];
    c8[7] = 0.f;
    store_chunk(t, row, 3, c8);
    // ---- color_net[0]: [128,32] x [64,32]^T -> cols 0..63
    fence_async_smem();
    tc_fence_before();
    group_sync(group);
    if (warp0) { tc_fence_after(); if (elect_one()) { issue_layer<64, 32>(t.a[0], t.a[1], w.w3[0], w.w3[1], t.tmem); umma_commit(&t.bar); } __syncwarp(); }
    mbar_wait(&t.bar, phase); phase ^= 1;
    tc_fence_after();
    epilogue_relu64(t, row, tm);
    // ---- color_net[1]: [128,64] x [64,64]^T -> cols 64..127
    fence_async_smem();
    tc_fence_before();
    group_sync(group);
    if (warp0) { tc_fence_after(); if (elect_one()) { issue_layer<64, 64>(t.a[0], t.a[1], w.w4[0], w.w4[1], t.tmem + 64); umma_commit(&t.bar); } __syncwarp(); }
    mbar_wait(&t.bar, phase); phase ^= 1;
    tc_fence_after();
    epilogue_relu64(t, row, tm + 64);
    // ---- color_net[2]: [128,64] x [16,64]^T -> cols 0..15 (3 used)
    fence_async_smem();
    tc_fence_before();
    group_sync(group);
    if (warp0) { tc_fence_after(); if (elect_one()) { issue_layer<16, 64>(t.a[0], t.a[1], w.w5[0], w.w5[1], t.tmem); umma_commit(&t.bar); } __syncwarp(); }
    mbar_wait(&t.bar, phase); phase ^= 1;
    tc_fence_after();
    tmem_ld16(tm, v);
    r = 1.0f / (1.0f + expf(-v[0]));
    g = 1.0f / (1.0f + expf(-v[1]));
    b = 1.0f / (1.0f + expf(-v[2]));
    tc_fence_before();  // order these TMEM reads before the next tile's first MMA (which follows a group_sync)
}

__device__ __forceinline__ void mlp_tile(TileSmem &t, const Weights &w, int group, int row, const float (&sh)[16], uint32_t &phase,
                                         float &sigma, float &r, float &g, float &b) {
    mlp_tile(t, w, group, row, sh, phase, sigma, r, g, b, t.a[0], t.a[1], nullptr);
}

// ---- TS variant: activations stay in tensor memory.  A operand of layers 2..5 is read by tcgen05.mma straight from
// TMEM (bf16 pairs packed along K: element (row, k) = lane `row`, 32-bit column k/2, half k%2), written there by the
// epilogue threads with tcgen05.st.  Removes the activation round trip through shared memory (stores + tensor-core
// operand reads), which shares the L1 data pipe with the hash-table gathers.
// TMEM columns of one group: D [0,64) fp32 accumulator | A_hi [64,96) | A_lo [96,128).
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&v)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]),
                 "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
                 : "memory");
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void umma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n"
        "}\n" ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
template <int N, int K>
__device__ __forceinline__ void issue_layer_ts(uint32_t a_hi_t, uint32_t a_lo_t, const __nv_bfloat16 *w_hi, const __nv_bfloat16 *w_lo, uint32_t tmem_d) {
    constexpr uint32_t idesc = instr_desc_bf16(N);
    const uint64_t b_hi = smem_desc(smem_u32(w_hi), N * 16, 128), b_lo = smem_desc(smem_u32(w_lo), N * 16, 128);
    uint32_t acc = 0;
#pragma unroll 1
    for (int term = 0; term < 3; term++) {  // small terms first, hi*hi last (same order as issue_layer)
        uint32_t at = term == 0 ? a_lo_t : a_hi_t;
        uint64_t bd = term == 1 ? b_lo : b_hi;
#pragma unroll 1
        for (int ks = 0; ks < K / 16; ks++) {
            umma_bf16_ts(tmem_d, at, bd, idesc, acc);
            acc = 1;
            at += 8;                        // 16 bf16 k-values = 8 columns
            bd += (2 * N * 16) >> 4;
        }
    }
}
// 8 consecutive k-values of this thread's row -> 4 hi words and 4 lo words
__device__ __forceinline__ void split8(const float (&v)[8], uint32_t *h, uint32_t *l) {
    split_pair(v[0], v[1], h[0], l[0]); split_pair(v[2], v[3], h[1], l[1]);
    split_pair(v[4], v[5], h[2], l[2]); split_pair(v[6], v[7], h[3], l[3]);
}
static __device__ __noinline__ void epilogue_relu64_ts(uint32_t tm) {   // D [0,64) -> relu -> A_hi/A_lo (this warp's lanes)
#pragma unroll 1
    for (int q = 0; q < 4; q++) {
        float v[16], c8[8];
        uint32_t h[8], l[8];
        tmem_ld16(tm + q * 16, v);
#pragma unroll
        for (int hh = 0; hh < 2; hh++) {
#pragma unroll
            for (int i = 0; i < 8; i++) c8[i] = fmaxf(v[hh * 8 + i], 0.f);
            split8(c8, h + 4 * hh, l + 4 * hh);
        }
        tmem_st8(tm + 64 + q * 8, h);
        tmem_st8(tm + 96 + q * 8, l);
    }
}
// Same contract as mlp_tile; layer 1 reads its A operand from shared memory (a0_hi/a0_lo), layers 2..5 from TMEM.
template <class TileT>   // anything with .tmem (TMEM base column of the group) and .bar (the group's MMA-done mbarrier)
__device__ __forceinline__ void mlp_tile_ts(TileT &t, const Weights &w, int group, int row, const float (&sh)[16], uint32_t &phase,
                                            float &sigma, float &r, float &g, float &b, const void *a0_hi, const void *a0_lo,
                                            uint64_t *release) {
    const uint32_t tm = t.tmem + ((uint32_t)(row & ~31) << 16);  // this warp's 32 TMEM lanes
    const bool warp0 = (row >> 5) == 0;
    const uint32_t a_hi_t = t.tmem + 64, a_lo_t = t.tmem + 96;
    float v[16], c8[8];
    uint32_t h[8], l[8];
    // ---- sigma_net[0]: A from the producer stage (smem) -> D
    tc_fence_before();
    group_sync(group);
    if (warp0) {
        tc_fence_after();
        if (elect_one()) { issue_layer<64, 32>(a0_hi, a0_lo, w.w1[0], w.w1[1], t.tmem); if (release) umma_commit(release); umma_commit(&t.bar); }
        __syncwarp();
    }
    mbar_wait(&t.bar, phase); phase ^= 1;
    tc_fence_after();
    epilogue_relu64_ts(tm);
    // ---- sigma_net[1]: [128,64] x [16,64]^T -> D cols 0..15
    tmem_wait_st();
    tc_fence_before();
    group_sync(group);
    if (warp0) { tc_fence_after(); if (elect_one()) { issue_layer_ts<16, 64>(a_hi_t, a_lo_t, w.w2[0], w.w2[1], t.tmem); umma_commit(&t.bar); } __syncwarp(); }
    mbar_wait(&t.bar, phase); phase ^= 1;
    tc_fence_after();
    tmem_ld16(tm, v);
    sigma = expf(v[0]);
    // colour-net input row: SH(16) | geo(15) | 0  (K = 32 -> 16 columns per image)
#pragma unroll
    for (int i = 0; i < 8; i++) c8[i] = sh[i];
    split8(c8, h, l);
#pragma unroll
    for (int i = 0; i < 8; i++) c8[i] = sh[8 + i];
    split8(c8, h + 4, l + 4);
    tmem_st8(tm + 64, h); tmem_st8(tm + 96, l);
#pragma unroll
    for (int i = 0; i < 8; i++) c8[i] = v[1 + i];
    split8(c8, h, l);
#pragma unroll
    for (int i = 0; i < 7; i++) c8[i] = v[9 + i];
    c8[7] = 0.f;
    split8(c8, h + 4, l + 4);
    tmem_st8(tm + 64 + 8, h); tmem_st8(tm + 96 + 8, l);
    // ---- color_net[0]: K = 32
    tmem_wait_st();
    tc_fence_before();
    group_sync(group);
    if (warp0) { tc_fence_after(); if (elect_one()) { issue_layer_ts<64, 32>(a_hi_t, a_lo_t, w.w3[0], w.w3[1], t.tmem); umma_commit(&t.bar); } __syncwarp(); }
    mbar_wait(&t.bar, phase); phase ^= 1;
    tc_fence_after();
    epilogue_relu64_ts(tm);
    // ---- color_net[1]: K = 64
    tmem_wait_st();
    tc_fence_before();
    group_sync(group);
    if (warp0) { tc_fence_after(); if (elect_one()) { issue_layer_ts<64, 64>(a_hi_t, a_lo_t, w.w4[0], w.w4[1], t.tmem); umma_commit(&t.bar); } __syncwarp(); }
    mbar_wait(&t.bar, phase); phase ^= 1;
    tc_fence_after();
    epilogue_relu64_ts(tm);
    // ---- color_net[2]: N = 16 (3 used)
    tmem_wait_st();
    tc_fence_before();
    group_sync(group);
    if (warp0) { tc_fence_after(); if (elect_one()) { issue_layer_ts<16, 64>(a_hi_t, a_lo_t, w.w5[0], w.w5[1], t.tmem); umma_commit(&t.bar); } __syncwarp(); }
    mbar_wait(&t.bar, phase); phase ^= 1;
    tc_fence_after();
    tmem_ld16(tm, v);
    r = 1.0f / (1.0f + expf(-v[0]));
    g = 1.0f / (1.0f + expf(-v[1]));
    b = 1.0f / (1.0f + expf(-v[2]));
    tc_fence_before();
}

// ---- TMA (1-D bulk copy global -> shared, completion on an mbarrier)
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_1d(void *smem_dst, const void *gsrc, uint32_t bytes, uint64_t *bar) {   // 16-byte multiples
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)),
                 "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// High-ILP encoder for a producer warp: all 8 corner indices of FOUR levels are computed first (the dense / hashed
// choice is a select, not a branch), then the 32 gathers are issued back to back, then interpolated, split and stored
// as one 16-byte k-chunk of the hi and of the lo image at `hi_row`/`lo_row` (+ chunk * 2048).  Requires every level to
// be dense or a power-of-two hash table (Weights::fast); same vertices, weights and summation order as lookup3_c2.
// `L0S`: level 0 (chunk 0, j == 0) is read from its shared-memory copy `lvl0` instead of global memory.
template <bool L0S>
__device__ __forceinline__ void encode_chunk_ilp(int c4, char *hi_row, char *lo_row, const Weights &w, const float2 *__restrict__ table,
                                                 const float2 *lvl0, bool in, float u, float vv, float ww) {
    float fr[4][3];
    // (one 128-bit load per x-neighbour pair that shares an aligned slot was measured here too: 3.5 vs 2.5 ms per
    //  frame — 128-bit gathers cost more L1 wavefronts than they save; plain 64-bit gathers.)
    uint32_t idx[4][8];
#pragma unroll
    for (int j = 0; j < 4; j++) {
        const int l = 4 * c4 + j;
        const LevelGeom g = w.geo[l];
        const uint32_t off = w.level_off[l];   // entry offsets stay 32-bit (whole table < 2^32 entries): one IMAD.WIDE per gather
        float px = u * g.scale + 0.5f, py = vv * g.scale + 0.5f, pz = ww * g.scale + 0.5f;
        const float fx = floorf(px), fy = floorf(py), fz = floorf(pz);
        fr[j][0] = px - fx; fr[j][1] = py - fy; fr[j][2] = pz - fz;
        const uint32_t gx = (uint32_t)fx, gy = (uint32_t)fy, gz = (uint32_t)fz;
        const uint32_t s1 = g.stride1, s2 = g.stride1 * g.stride1;
        const uint32_t dy0 = gy * s1, dz0 = gz * s2;
        const uint32_t hy0 = gy * 2654435761u, hz0 = gz * 805459861u;
        // per-axis terms: dense -> added, hashed -> xored; (v+1)*P == v*P + P mod 2^32
        const uint32_t ty0 = g.dense3 ? dy0 : hy0, ty1 = g.dense3 ? dy0 + s1 : hy0 + 2654435761u;
        const uint32_t tz0 = g.dense3 ? dz0 : hz0, tz1 = g.dense3 ? dz0 + s2 : hz0 + 805459861u;
        const uint32_t m = g.dense3 ? 0xffffffffu : g.mask;
#pragma unroll
        for (int c = 0; c < 8; c++) {
            const uint32_t tx = gx + (c & 1), ty = (c & 2) ? ty1 : ty0, tz = (c & 4) ? tz1 : tz0;
            const uint32_t id = g.dense3 ? (tx + ty + tz) : ((tx ^ ty ^ tz) & m);
            idx[j][c] = in ? id + off : 0u;
        }
    }
    float2 v[4][8];
#pragma unroll
    for (int j = 0; j < 4; j++)
#pragma unroll
        for (int c = 0; c < 8; c++) v[j][c] = (L0S && j == 0) ? lvl0[idx[j][c]] : __ldg(table + idx[j][c]);   // level 0 starts at entry 0
    uint32_t hw[4], lw[4];
#pragma unroll
    for (int j = 0; j < 4; j++) {
        const float px = fr[j][0], py = fr[j][1], pz = fr[j][2];
        const float qx = 1.0f - px, qy = 1.0f - py, qz = 1.0f - pz;
        const float w00 = qx * qy, w10 = px * qy, w01 = qx * py, w11 = px * py;
        const float wt[8] = {w00 * qz, w10 * qz, w01 * qz, w11 * qz, w00 * pz, w10 * pz, w01 * pz, w11 * pz};
        float2 e = make_float2(0.f, 0.f);
#pragma unroll
        for (int c = 0; c < 8; c++) { e.x += wt[c] * v[j][c].x; e.y += wt[c] * v[j][c].y; }
        if (!in) e = make_float2(0.f, 0.f);
        split_pair(e.x, e.y, hw[j], lw[j]);
    }
    *reinterpret_cast<uint4 *>(hi_row + c4 * 2048) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
    *reinterpret_cast<uint4 *>(lo_row + c4 * 2048) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
}
__device__ __forceinline__ void encode_rows_ilp(char *hi_row, char *lo_row, const Weights &w, const float2 *__restrict__ table, float bound,
                                                bool valid, float x, float y, float z, const float2 *lvl0 = nullptr) {
    const float inv = 1.0f / (2 * bound);
    const float u = (x + bound) * inv, vv = (y + bound) * inv, ww = (z + bound) * inv;
    const bool in = valid && !(u < 0 || u > 1 || vv < 0 || vv > 1 || ww < 0 || ww > 1);
    if (lvl0) encode_chunk_ilp<true>(0, hi_row, lo_row, w, table, lvl0, in, u, vv, ww);
    else encode_chunk_ilp<false>(0, hi_row, lo_row, w, table, nullptr, in, u, vv, ww);
#pragma unroll 1
    for (int c4 = 1; c4 < 4; c4++) encode_chunk_ilp<false>(c4, hi_row, lo_row, w, table, nullptr, in, u, vv, ww);
}

// Encode one sample and store its 32 features (k = 2*level, 2*level+1; chunks 0..3) for sigma_net[0].  Zero row for
// invalid samples.  Four levels (= one 16-byte k-chunk of the hi and of the lo image) per rolled iteration: 32 gathers
// in flight per thread, and one conflict-free 128-bit store per image instead of four 4-byte stores at a 16-byte
// stride (which serialised 4-way on the shared-memory banks).
__device__ __forceinline__ void encode_to_tile(TileSmem &t, const Weights &w, const float2 *__restrict__ table, float bound, int row,
                                               bool valid, float x, float y, float z) {
    const float inv = 1.0f / (2 * bound);
    const float u = (x + bound) * inv, vv = (y + bound) * inv, ww = (z + bound) * inv;
    const bool in = valid && !(u < 0 || u > 1 || vv < 0 || vv > 1 || ww < 0 || ww > 1);
    char *hi = reinterpret_cast<char *>(t.a[0]) + row * 16, *lo = reinterpret_cast<char *>(t.a[1]) + row * 16;
#pragma unroll 1
    for (int c4 = 0; c4 < 4; c4++) {
        uint32_t hw[4], lw[4];
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const int l = 4 * c4 + j;
            float2 e = make_float2(0.f, 0.f);
            if (in) e = lookup3_c2(table + w.level_off[l], w.geo[l], u, vv, ww, 0);
            split_pair(e.x, e.y, hw[j], lw[j]);
        }
        *reinterpret_cast<uint4 *>(hi + c4 * 2048) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
        *reinterpret_cast<uint4 *>(lo + c4 * 2048) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
    }
}

}  // namespace tc
}  // namespace pn
