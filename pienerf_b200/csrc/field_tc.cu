// Stand-alone tensor-core field pass: NeRFNetwork.forward (nerf/network.py:98-127) over M samples with the MLP on
// tcgen05 (pn_field_forward mode 1).  One persistent CTA per SM, 3 independent 128-sample tile groups per CTA that
// share the bf16 hi/lo weight images in shared memory; see field_tc.cuh.
#include "field_tc.cuh"

namespace {

constexpr int kGroups = 3;

struct __align__(128) FieldTcSmem {
    pn::tc::Weights w;
    pn::tc::TileSmem tile[kGroups];
    uint32_t tmem_base;
};

__global__ void __launch_bounds__(kGroups * 128, 1) field_forward_tc_kernel(const pn_field_t f, const float *__restrict__ xyzs,
                                                                            const float *__restrict__ dirs, uint32_t M,
                                                                            float *__restrict__ sigmas, float *__restrict__ rgbs) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    FieldTcSmem &S = *reinterpret_cast<FieldTcSmem *>(smem_raw);
    const int group = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 7), 0), row = threadIdx.x & 127;  // provably warp-uniform
    pn::tc::TileSmem &T = S.tile[group];
    pn::tc::weights_fill(S.w, f);
    if (row == 0) pn::tc::mbar_init(&T.bar, 1);
    pn::tc::fence_barrier_init();
    // one warp allocates for the whole CTA (tcgen05.relinquish_alloc_permit forbids further allocations by the CTA);
    // 3 groups x 128 columns rounded up to the next power of two
    if (threadIdx.x < 32) pn::tc::tmem_alloc(&S.tmem_base, 512);
    pn::tc::fence_async_smem();
    pn::tc::tc_fence_before();
    __syncthreads();
    pn::tc::tc_fence_after();
    if (row == 0) T.tmem = S.tmem_base + group * pn::tc::kTmemCols;
    pn::tc::group_sync(group);
    const float2 *table = reinterpret_cast<const float2 *>(f.embeddings);
    uint32_t phase = 0;
    const uint32_t n_tiles = (M + 127) / 128;
    for (uint32_t tile = blockIdx.x * kGroups + group; tile < n_tiles; tile += gridDim.x * kGroups) {
        const uint32_t i = tile * 128 + row;
        const bool valid = i < M;
        float x = 0, y = 0, z = 0, dx = 0, dy = 0, dz = 1;
        if (valid) { x = xyzs[3 * i]; y = xyzs[3 * i + 1]; z = xyzs[3 * i + 2]; dx = dirs[3 * i]; dy = dirs[3 * i + 1]; dz = dirs[3 * i + 2]; }
        float sh[16];
        pn::sh_eval<4>(dx, dy, dz, sh);
        pn::tc::encode_to_tile(T, S.w, table, f.bound, row, valid, x, y, z);
        float sigma, r, g, b;
        pn::tc::mlp_tile(T, S.w, group, row, sh, phase, sigma, r, g, b);
        if (valid) { sigmas[i] = sigma; rgbs[3 * i] = r; rgbs[3 * i + 1] = g; rgbs[3 * i + 2] = b; }
    }
    pn::tc::tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) pn::tc::tmem_dealloc(S.tmem_base, 512);
}

}  // namespace

int pn_field_forward_tc(const pn_field_t *f, const float *xyzs, const float *dirs, uint32_t M, float *sigmas, float *rgbs,
                        cudaStream_t st) {
    const size_t smem = sizeof(FieldTcSmem) + 128;
    cudaError_t e = cudaFuncSetAttribute(field_forward_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { pn_set_error("cudaFuncSetAttribute(field_forward_tc_kernel): %s", cudaGetErrorString(e)); return PN_ECUDA; }
    const uint32_t n_tiles = (M + 127) / 128;
    const uint32_t blocks = min((n_tiles + kGroups - 1) / kGroups, (uint32_t)pn_sm_count_cached());
    field_forward_tc_kernel<<<blocks, kGroups * 128, smem, st>>>(*f, xyzs, dirs, M, sigmas, rgbs);
    PN_LAUNCH_CHECK("field_forward_tc_kernel");
    return PN_OK;
}
