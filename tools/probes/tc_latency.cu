// tc_latency.cu -- hand-off latencies that bound the K3b pipeline (one CTA, cycle counter of its SM).  Standalone:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/probes/tc_latency tools/probes/tc_latency.cu
// 1. n x tcgen05.mma kind::i8 (128x128x32) + tcgen05.commit, waited for by the issuing thread: issue -> wake, n = 1,3,6,12
//    (slope = pipe throughput, intercept = completion + commit + try_wait latency)
// 2. the same, waited for by a lane of ANOTHER warp (what an epilogue warp sees), stamped with the shared SM clock
// 3. mbarrier.arrive by warp B -> try_wait wake in warp A (what an issuer sees when an epilogue warp frees a buffer)
// 4. tcgen05.ld 32x32b.x32.pack::16b + wait::ld round trip of one warp (64 columns), and two back to back
// 5. try_wait on an already completed phase; volatile shared-memory load (the cheap alternative for polling)
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %d: %s\n", #x, __LINE__, cudaGetErrorString(e_)); return 1; } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void tmem_alloc(uint32_t *dst_smem, uint32_t cols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t cols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
}
__device__ __forceinline__ void fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ bool mbar_wait(uint64_t *bar, uint32_t parity) {
    for (int spin = 0; spin < (1 << 22); spin++)
        if (mbar_try(bar, parity)) return true;
    return false;
}
__device__ __forceinline__ void mma_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mma_i8(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
                 "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
                 : "memory");
}
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((smem_addr & 0x3FFFFu) >> 4) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32) | ((uint64_t)1 << 46);
}
#define TMEM_LD_X32_PACK(r, taddr)                                                                                \
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.pack::16b.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15," \
                 "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"                       \
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), \
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]),       \
                   "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]),     \
                   "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]),     \
                   "=r"(r[29]), "=r"(r[30]), "=r"(r[31])                                                          \
                 : "r"(taddr) : "memory")

// out[0..3]: self-wait latency for n = 1,3,6,12 MMAs; out[4..7]: other-warp wake latency; out[8]: arrive -> wake;
// out[9]: one LDTM round trip; out[10]: two LDTMs + one wait; out[11]: try_wait on a completed phase; out[12]: volatile LDS;
// out[13]: tcgen05.commit with nothing outstanding -> wake in the same thread
__global__ void __launch_bounds__(128) k_latency(unsigned long long *out, uint32_t *sink, int reps) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar, bar2, bar_done;
    __shared__ uint32_t s_base;
    __shared__ volatile long long s_stamp;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int i = tid; i < 64 * 1024 / 4; i += 128) reinterpret_cast<uint32_t *>(smem)[i] = 0x01010101u;
    if (tid == 0) {
        mbar_init(&bar, 1); mbar_init(&bar2, 1); mbar_init(&bar_done, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (warp == 0) tmem_alloc(&s_base, 512);
    fence_before();
    __syncthreads();
    fence_after();
    const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((128u >> 3) << 17) | ((128u >> 4) << 24);
    const uint64_t da0 = make_desc(smem_u32(smem), 128 * 16, 128), db0 = make_desc(smem_u32(smem) + 16384, 128 * 16, 128);
    const int ns[4] = {1, 3, 6, 12};
    uint32_t ph = 0, ph2 = 0;
    // ---- 1 + 2: issue n MMAs + commit in warp 0 lane 0; warp 0 waits itself (mode 0) or warp 1 lane 0 waits (mode 1)
    for (int mode = 0; mode < 2; mode++)
        for (int c = 0; c < 4; c++) {
            long long acc = 0;
            for (int r = 0; r < reps; r++) {
                __syncthreads();
                if (tid == 0) {
                    const long long t0 = clock64();
                    s_stamp = t0;
                    for (int i = 0; i < ns[c]; i++) mma_i8(s_base + (uint32_t)(i & 1) * 128, da0 + (uint64_t)((i % 3) * 256), db0 + (uint64_t)((i % 12) * 256), idesc, i > 1);
                    mma_commit(&bar);
                    if (mode == 0) { mbar_wait(&bar, ph); acc += clock64() - t0; }
                }
                if (mode == 1 && tid == 32) { mbar_wait(&bar, ph); const long long t1 = clock64(); acc += t1 - s_stamp; }
                if (mode == 1 && tid == 0) mbar_wait(&bar, ph);
                ph ^= 1u;
                fence_after();
            }
            if ((mode == 0 && tid == 0) || (mode == 1 && tid == 32)) out[mode * 4 + c] = (unsigned long long)(acc / reps);
        }
    // ---- 3: warp 1 arrives, warp 0 is already waiting
    {
        long long acc = 0;
        for (int r = 0; r < reps; r++) {
            __syncthreads();
            if (tid == 32) { for (int d = 0; d < 200; d++) asm volatile("nanosleep.u32 20;"); s_stamp = clock64(); mbar_arrive(&bar2); }
            if (tid == 0) { mbar_wait(&bar2, ph2); const long long t1 = clock64(); acc += t1 - s_stamp; }
            ph2 ^= 1u;
        }
        if (tid == 0) out[8] = (unsigned long long)(acc / reps);
    }
    // ---- 4: LDTM round trips (warp 0)
    if (warp == 0) {
        uint32_t a[32], b[32], o = 0;
        long long acc1 = 0, acc2 = 0;
        for (int r = 0; r < reps; r++) {
            long long t0 = clock64();
            TMEM_LD_X32_PACK(a, s_base);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            for (int i = 0; i < 32; i++) o |= a[i];
            acc1 += clock64() - t0;
            t0 = clock64();
            TMEM_LD_X32_PACK(a, s_base);
            TMEM_LD_X32_PACK(b, s_base + 64);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            for (int i = 0; i < 32; i++) o |= a[i] | b[i];
            acc2 += clock64() - t0;
        }
        if (o == 0x12345678u) sink[0] = o;
        if (lane == 0) { out[9] = (unsigned long long)(acc1 / reps); out[10] = (unsigned long long)(acc2 / reps); }
    }
    // ---- 5: try_wait on a completed phase, volatile LDS, empty commit
    if (tid == 0) {
        mbar_arrive(&bar_done);                                   // phase 0 complete
        long long acc = 0;
        uint32_t o = 0;
        for (int r = 0; r < reps; r++) { const long long t0 = clock64(); o += mbar_try(&bar_done, 0); acc += clock64() - t0; }
        out[11] = (unsigned long long)(acc / reps);
        acc = 0;
        for (int r = 0; r < reps; r++) { const long long t0 = clock64(); o += *reinterpret_cast<volatile uint32_t *>(&s_base); acc += clock64() - t0; }
        out[12] = (unsigned long long)(acc / reps);
        acc = 0;
        for (int r = 0; r < reps; r++) { const long long t0 = clock64(); mma_commit(&bar); mbar_wait(&bar, ph); ph ^= 1u; acc += clock64() - t0; }
        out[13] = (unsigned long long)(acc / reps);
        if (o == 0x12345678u) sink[1] = o;
    }
    // ---- 6: polling the mbarrier WORD with a volatile 64-bit load instead of try_wait (phase = bit 63 ?)
    __syncthreads();
    {
        __shared__ __align__(8) uint64_t bar3;
        volatile unsigned long long *w = reinterpret_cast<volatile unsigned long long *>(&bar3);
        if (tid == 0) {
            mbar_init(&bar3, 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            out[16] = *w;
            mbar_arrive(&bar3); out[17] = *w;
            mbar_arrive(&bar3); out[18] = *w;
            mbar_arrive(&bar3); out[20] = *w;
            mbar_arrive(&bar3); out[21] = *w;
        }
        __syncthreads();
        unsigned long long base_phase = (*w) >> 63;                 // after four completions
        long long acc = 0, acc_c = 0;
        for (int r = 0; r < reps; r++) {
            __syncthreads();
            const unsigned long long want = base_phase ^ 1ull;
            if (tid == 32) { for (int d = 0; d < 100; d++) asm volatile("nanosleep.u32 20;"); s_stamp = clock64(); mbar_arrive(&bar3); }
            if (tid == 0) { int spin = 0; while (((*w) >> 63) != want && ++spin < (1 << 11)) { } const long long t1 = clock64(); acc += t1 - s_stamp; }
            base_phase ^= 1ull;
        }
        if (tid == 0) out[14] = (unsigned long long)(acc / reps);
        for (int r = 0; r < reps; r++) {                            // 3 MMAs + commit, another warp polls the word
            __syncthreads();
            const unsigned long long want = base_phase ^ 1ull;
            if (tid == 0) {
                s_stamp = clock64();
                for (int i = 0; i < 3; i++) mma_i8(s_base + (uint32_t)(i & 1) * 128, da0 + (uint64_t)((i % 3) * 256), db0 + (uint64_t)((i % 12) * 256), idesc, i > 1);
                mma_commit(&bar3);
            }
            if (tid == 32) { int spin = 0; while (((*w) >> 63) != want && ++spin < (1 << 11)) { } const long long t1 = clock64(); acc_c += t1 - s_stamp; }
            __syncthreads();
            if (tid == 0) { int spin = 0; while (((*w) >> 63) != want && ++spin < (1 << 11)) { } }
            base_phase ^= 1ull;
            fence_after();
        }
        if (tid == 32) out[15] = (unsigned long long)(acc_c / reps);
        // plain shared-memory flag: warp B stores, warp A polls
        __shared__ volatile uint32_t flag;
        if (tid == 0) flag = 0;
        long long acc_f = 0;
        for (int r = 0; r < reps; r++) {
            __syncthreads();
            if (tid == 32) { for (int d = 0; d < 100; d++) asm volatile("nanosleep.u32 20;"); s_stamp = clock64(); __threadfence_block(); flag = (uint32_t)r + 1u; }
            if (tid == 0) { int spin = 0; while (flag != (uint32_t)r + 1u && ++spin < (1 << 11)) { } const long long t1 = clock64(); acc_f += t1 - s_stamp; }
        }
        if (tid == 0) out[19] = (unsigned long long)(acc_f / reps);
    }
    fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(s_base, 512);
}

int main() {
    unsigned long long *d_out, h[24] = {0};
    uint32_t *d_sink;
    CK(cudaMalloc(&d_out, sizeof h));
    CK(cudaMemset(d_out, 0, sizeof h));
    CK(cudaMalloc(&d_sink, 64));
    CK(cudaFuncSetAttribute(k_latency, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
    for (int rep = 0; rep < 2; rep++) {
        k_latency<<<1, 128, 64 * 1024>>>(d_out, d_sink, 200);
        CK(cudaDeviceSynchronize());
    }
    CK(cudaMemcpy(h, d_out, sizeof h, cudaMemcpyDeviceToHost));
    printf("MMA x n + commit -> wake in the issuing thread : n=1 %llu  n=3 %llu  n=6 %llu  n=12 %llu clk\n", h[0], h[1], h[2], h[3]);
    printf("MMA x n + commit -> wake in another warp       : n=1 %llu  n=3 %llu  n=6 %llu  n=12 %llu clk\n", h[4], h[5], h[6], h[7]);
    printf("mbarrier.arrive (warp B) -> try_wait wake (warp A): %llu clk\n", h[8]);
    printf("tcgen05.ld x32 pack + wait: %llu clk; two loads + one wait: %llu clk\n", h[9], h[10]);
    printf("try_wait on a completed phase: %llu clk; volatile LDS: %llu clk; empty commit -> wake: %llu clk\n", h[11], h[12], h[13]);
    printf("mbarrier word: init %016llx, after 1st completion %016llx, after 2nd %016llx, 3rd %016llx, 4th %016llx\n", h[16], h[17], h[18], h[20], h[21]);
    printf("polling the mbarrier word (volatile LDS.64): arrive -> seen %llu clk; 3 MMAs + commit -> seen in another warp %llu clk\n", h[14], h[15]);
    printf("plain shared-memory flag: store (warp B) -> seen by polling warp A: %llu clk\n", h[19]);
    return 0;
}
