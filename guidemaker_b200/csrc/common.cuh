// common.cuh -- shared host/device helpers of libgm_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/gm_b200.h"

namespace gm {

// ---- error plumbing (thread-local message behind gm_last_error) -----------------------------------
void set_error(const char *fmt, ...);
int cuda_fail(cudaError_t e, const char *what, const char *file, int line);

#define GM_CUDA(call)                                                        \
    do {                                                                     \
        cudaError_t e__ = (call);                                            \
        if (e__ != cudaSuccess) return gm::cuda_fail(e__, #call, __FILE__, __LINE__); \
    } while (0)

#define GM_ARG(cond, ...)                                                    \
    do {                                                                     \
        if (!(cond)) { gm::set_error(__VA_ARGS__); return GM_ERR_ARG; }      \
    } while (0)

// ---- profiling counters (gm_prof_*) ----------------------------------------------------------------
struct Prof {
    bool on = false;
    long long all_launches = 0;
    long long scan_launches = 0;
    double pairs = 0.0;
    double scan_ms_done = 0.0;       // already synchronised and summed
    static const int MAXEV = 4096;
    cudaEvent_t ev[MAXEV][2];
    int n_ev = 0, n_alloc = 0;
};
Prof &prof();
inline void count_launch(int n = 1) { prof().all_launches += n; }

int device_sm_count();
bool initialised();
int ensure_init();

// Device scratch memory comes from a small caching allocator (api.cu): cudaMalloc'd blocks are kept on a free list and
// handed out again, stream-ordered -- a block records an event on the stream it was released on, and a taker on another
// stream waits for that event first.  In the steady state no driver allocation call is made at all.  (The driver's own
// stream-ordered pool, cudaMallocAsync, was used first: on the B200 boxes it stalled sporadically for 50-2000 ms inside
// an allocation when buffers of a few hundred MB were returned and taken again every call -- tools/e2e_stall_probe.py.)
cudaError_t dev_alloc(void **p, size_t bytes, cudaStream_t st);
void dev_free(void *p, cudaStream_t st);

// Touch every page of a host OUTPUT buffer with several threads before a large device->host copy lands in it.  A fresh
// numpy array is untouched virtual memory: the copy's destination pages are then faulted in (and zero-filled by the
// kernel) one at a time by the single thread that drives the copy -- measured on the B200 box: 512 MB device->host takes
// 255 ms into untouched memory, 26 ms into touched memory (tools/probes/alloc_probe.cu).  Faulting scales with threads.
void prefault(void *p, size_t bytes);

// GM_TRACE=1 prints wall-clock per host-side phase of the host-buffer entry points
bool trace_on();
double now_ms();
void trace(const char *what, double t0_ms);

// ---- device helpers ---------------------------------------------------------------------------------
#ifdef __CUDACC__

__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// mbarrier + 1-D bulk async copy (TMA engine, SASS: UBLKCP) -- used to stream the target table
// through shared memory.
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}\n"
        : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
// Wait with a watchdog: a pipeline bug traps (the launch fails with an error) instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    if (mbar_try_wait(bar, parity)) return;
    const long long t0 = clock64();
    uint32_t spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        if ((++spins & 0xFFFu) == 0 && clock64() - t0 > 20000000000LL) {      // ~10 s at 2 GHz
            printf("libgm_b200: mbarrier watchdog fired (block %d,%d thread %d)\n", blockIdx.x, blockIdx.y, threadIdx.x);
            __trap();
        }
    }
}
// global -> shared bulk copy, completion signalled on `bar` as transaction bytes.
// dst/src 16-byte aligned, bytes a multiple of 16.
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// 2-bit interleaved guide -> bit planes: lo bit i = bit 2i, hi bit i = bit 2i+1.
__device__ __forceinline__ uint32_t compress_even_bits(uint64_t x) {
    x &= 0x5555555555555555ULL;
    x = (x | (x >> 1)) & 0x3333333333333333ULL;
    x = (x | (x >> 2)) & 0x0F0F0F0F0F0F0F0FULL;
    x = (x | (x >> 4)) & 0x00FF00FF00FF00FFULL;
    x = (x | (x >> 8)) & 0x0000FFFF0000FFFFULL;
    x = (x | (x >> 16)) & 0x00000000FFFFFFFFULL;
    return static_cast<uint32_t>(x);
}
__device__ __forceinline__ uint2 to_planes(uint64_t g) {
    return make_uint2(compress_even_bits(g), compress_even_bits(g >> 1));
}
__device__ __forceinline__ uint64_t spread_bits(uint32_t v) {
    uint64_t x = v;
    x = (x | (x << 16)) & 0x0000FFFF0000FFFFULL;
    x = (x | (x << 8)) & 0x00FF00FF00FF00FFULL;
    x = (x | (x << 4)) & 0x0F0F0F0F0F0F0F0FULL;
    x = (x | (x << 2)) & 0x3333333333333333ULL;
    x = (x | (x << 1)) & 0x5555555555555555ULL;
    return x;
}
__device__ __forceinline__ uint64_t from_planes(uint32_t lo, uint32_t hi) {
    return spread_bits(lo) | (spread_bits(hi) << 1);
}

#endif  // __CUDACC__
}  // namespace gm
