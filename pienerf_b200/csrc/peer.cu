// Peer-memory plumbing of the multi-GPU frame (one process per GPU on one NVLink / NVSwitch node).
//
// The reference has no multi-GPU path; BASELINE.json's north_star shards the frame by image tile with the simulator on
// one GPU.  Instead of an NCCL broadcast + gather per frame, the frame's two exchanges are stores into peer memory:
//   * rank 0 pushes the packed IP state (156 B / IP) into every rank's state slot (pn_peer_put) and then raises that
//     rank's `state` flag (pn_epoch_signal);
//   * every rank's compositor writes its finished pixels straight into rank 0's frame (pn_render_deformed_ex with a pixel
//     map and peer pointers) and then raises its `done` flag in rank 0's memory;
//   * consumers wait on flags that live in THEIR OWN memory (pn_epoch_wait / the wait of pn_render_deformed_ex), so
//     polling never crosses NVLink.
// Flags carry epochs: each frame slot owns a device-resident counter that the slot's first kernel bumps, so the same
// launches can be replayed from a CUDA graph frame after frame with no host-side argument patching.
// Memory comes from cudaMalloc and travels between processes as CUDA IPC handles (cudaIpcMemLazyEnablePeerAccess maps it
// over NVLink).  Every wait has a timeout and reports through a status word instead of hanging the GPU.
#include <cstring>
#include "common.cuh"

namespace {

__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t *p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(uint32_t *p, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned long long global_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// One thread per flag.  Thread 0 first bumps the slot's epoch (bump = 1); every thread then spins until
// *flags[t] + lag >= epoch.  A timeout stores (1 + t) in *status and lets the stream continue.
__global__ void epoch_wait_kernel(uint32_t *epoch, int bump, const uint32_t *const *flags, int n_flags, uint32_t lag, int *status,
                                  unsigned long long timeout_ns) {
    __shared__ uint32_t e;
    if (threadIdx.x == 0) {
        uint32_t v = *epoch;
        if (bump) { v += 1; *epoch = v; }
        e = v;
    }
    __syncthreads();
    const int t = threadIdx.x;
    if (t >= n_flags) return;
    const uint32_t *f = flags[t];
    const unsigned long long t0 = global_ns();
    while ((int32_t)(ld_acquire_sys(f) + lag - e) < 0) {
        __nanosleep(100);
        if (global_ns() - t0 > timeout_ns) {
            if (status) atomicExch(status, 1 + t);
            break;
        }
    }
}

__global__ void epoch_signal_kernel(const uint32_t *epoch, uint32_t *const *flags, int n_flags) {
    const int t = threadIdx.x;
    if (t >= n_flags) return;
    __threadfence_system();              // everything this stream wrote before (peer stores included) is visible first
    st_release_sys(flags[t], *epoch);
}

// grid = (blocks_per_dst, n_dst): 128-bit stores of the same source into every destination (NVLink writes)
__global__ void __launch_bounds__(256) peer_put_kernel(const int4 *__restrict__ src, size_t n16, void *const *dsts) {
    int4 *dst = reinterpret_cast<int4 *>(dsts[blockIdx.y]);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (size_t)gridDim.x * blockDim.x) dst[i] = src[i];
}

}  // namespace

extern "C" int pn_epoch_wait(uint32_t *epoch, int bump, const uint32_t *const *flags_dev, int n_flags, uint32_t lag, int *status,
                             uint32_t timeout_ms, void *stream) {
    PN_REQUIRE(epoch && (n_flags == 0 || flags_dev) && n_flags >= 0 && n_flags <= 64, "epoch / flags");
    epoch_wait_kernel<<<1, 64, 0, PN_STREAM(stream)>>>(epoch, bump, flags_dev, n_flags, lag, status, (unsigned long long)timeout_ms * 1000000ull);
    PN_LAUNCH_CHECK("epoch_wait_kernel");
    return PN_OK;
}

extern "C" int pn_epoch_signal(const uint32_t *epoch, uint32_t *const *flags_dev, int n_flags, void *stream) {
    PN_REQUIRE(epoch && flags_dev && n_flags > 0 && n_flags <= 64, "epoch / flags");
    epoch_signal_kernel<<<1, 64, 0, PN_STREAM(stream)>>>(epoch, flags_dev, n_flags);
    PN_LAUNCH_CHECK("epoch_signal_kernel");
    return PN_OK;
}

extern "C" int pn_peer_put(const void *src, uint64_t bytes, void *const *dsts_dev, int n_dst, void *stream) {
    PN_REQUIRE(src && dsts_dev && n_dst > 0 && bytes % 16 == 0 && ((uintptr_t)src & 15) == 0, "pn_peer_put: 16-byte aligned source and size");
    if (bytes == 0) return PN_OK;
    const size_t n16 = bytes / 16;
    const unsigned bx = (unsigned)min((size_t)32, div_up(n16, (size_t)256));
    peer_put_kernel<<<dim3(bx, (unsigned)n_dst), 256, 0, PN_STREAM(stream)>>>((const int4 *)src, n16, dsts_dev);
    PN_LAUNCH_CHECK("peer_put_kernel");
    return PN_OK;
}

// the two flag operations of pn_render_deformed_ex (render_fused.cu)
int pn_flag_wait_launch(const pn_frame_io_t *io, cudaStream_t st) {
    epoch_wait_kernel<<<1, 64, 0, st>>>(io->epoch, 1, io->wait_flag, io->n_wait, 0, io->status, (unsigned long long)io->timeout_ms * 1000000ull);
    PN_LAUNCH_CHECK("epoch_wait_kernel");
    return PN_OK;
}
int pn_flag_signal_launch(const pn_frame_io_t *io, cudaStream_t st) {
    epoch_signal_kernel<<<1, 64, 0, st>>>(io->epoch, io->signal_flag, io->n_signal);
    PN_LAUNCH_CHECK("epoch_signal_kernel");
    return PN_OK;
}

// ------------------------------------------------------------------------------------------------ memory + IPC
extern "C" int pn_peer_alloc(uint64_t bytes, void **ptr) {
    PN_REQUIRE(ptr && bytes > 0, "null pointer / zero size");
    PN_CUDA(cudaMalloc(ptr, bytes));
    PN_CUDA(cudaMemset(*ptr, 0, bytes));
    PN_CUDA(cudaDeviceSynchronize());
    return PN_OK;
}
extern "C" int pn_peer_free(void *ptr) {
    PN_CUDA(cudaFree(ptr));
    return PN_OK;
}
extern "C" int pn_peer_export(void *ptr, unsigned char *handle64) {
    PN_REQUIRE(ptr && handle64, "null pointer");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    cudaIpcMemHandle_t h;
    PN_CUDA(cudaIpcGetMemHandle(&h, ptr));
    memcpy(handle64, &h, 64);
    return PN_OK;
}
extern "C" int pn_peer_open(const unsigned char *handle64, void **ptr) {
    PN_REQUIRE(ptr && handle64, "null pointer");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, 64);
    PN_CUDA(cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return PN_OK;
}
extern "C" int pn_peer_close(void *ptr) {
    PN_CUDA(cudaIpcCloseMemHandle(ptr));
    return PN_OK;
}
