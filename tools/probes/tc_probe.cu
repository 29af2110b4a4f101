// tc_probe.cu -- feasibility probes for the tensor-core kNN variant (K3b).  Standalone:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/probes/tc_probe tools/probes/tc_probe.cu
// 1. TMEM read bandwidth (tcgen05.ld 32x32b) per SM with 4 warps.
// 2. tcgen05.mma kind::i8 (M=128, N=256, K=32 per instruction), operands in shared memory in the
//    canonical no-swizzle K-major layout: correctness against the host and instruction throughput.
// Every wait is bounded, so a wrong descriptor makes the probe report failure instead of hanging.
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

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
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

#define TMEM_LD_X32(r, taddr)                                                                                     \
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15," \
                 "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"                       \
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), \
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]),       \
                   "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]),     \
                   "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]),     \
                   "=r"(r[29]), "=r"(r[30]), "=r"(r[31])                                                          \
                 : "r"(taddr))

#define TMEM_LD_X32_PACK(r, taddr)                                                                                \
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.pack::16b.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15," \
                 "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"                       \
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), \
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]),       \
                   "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]),     \
                   "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]),     \
                   "=r"(r[29]), "=r"(r[30]), "=r"(r[31])                                                          \
                 : "r"(taddr))

// packed 16-bit read: 32 registers per load cover 64 columns.  Also reports how the two columns land in a register.
__global__ void __launch_bounds__(256) k_tmem_bw_pack(unsigned long long *cycles, uint32_t *sink, int iters, int nwarps) {
    __shared__ uint32_t s_base;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) tmem_alloc(&s_base, 512);
    fence_before();
    __syncthreads();
    fence_after();
    const uint32_t addr = s_base + ((uint32_t)((warp & 3) * 32) << 16);
    uint32_t acc = 0;
    uint32_t a[32], b[32];
    __syncthreads();
    const long long t0 = clock64();
    if (warp < nwarps) {
        for (int it = 0; it < iters; it++) {
            for (int col = 0; col < 512; col += 128) {
                TMEM_LD_X32_PACK(a, addr + col);
                TMEM_LD_X32_PACK(b, addr + col + 64);
                tmem_wait_ld();
#pragma unroll
                for (int i = 0; i < 32; i += 2) acc |= a[i] | a[i + 1];
#pragma unroll
                for (int i = 0; i < 32; i += 2) acc |= b[i] | b[i + 1];
            }
        }
    }
    const long long t1 = clock64();
    if (acc == 0x12345678u) sink[0] = acc;
    if (threadIdx.x == 0) cycles[blockIdx.x] = (unsigned long long)(t1 - t0);
    fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(s_base, 512);
}

// ---- probe 1: TMEM read bandwidth -------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_tmem_bw(unsigned long long *cycles, uint32_t *sink, int iters, int inflight) {
    __shared__ uint32_t s_base;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) tmem_alloc(&s_base, 512);
    fence_before();
    __syncthreads();
    fence_after();
    const uint32_t addr = s_base + ((uint32_t)(warp * 32) << 16);
    uint32_t acc = 0;
    uint32_t a[32], b[32], c[32], d[32];
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
        for (int col = 0; col < 512; col += 128) {
            TMEM_LD_X32(a, addr + col);
            TMEM_LD_X32(b, addr + col + 32);
            if (inflight >= 4) {
                TMEM_LD_X32(c, addr + col + 64);
                TMEM_LD_X32(d, addr + col + 96);
            }
            tmem_wait_ld();
#pragma unroll
            for (int i = 0; i < 32; i += 2) acc |= a[i] | a[i + 1];
#pragma unroll
            for (int i = 0; i < 32; i += 2) acc |= b[i] | b[i + 1];
            if (inflight >= 4) {
#pragma unroll
                for (int i = 0; i < 32; i += 2) acc |= c[i] | c[i + 1];
#pragma unroll
                for (int i = 0; i < 32; i += 2) acc |= d[i] | d[i + 1];
            } else {
                TMEM_LD_X32(a, addr + col + 64);
                TMEM_LD_X32(b, addr + col + 96);
                tmem_wait_ld();
#pragma unroll
                for (int i = 0; i < 32; i += 2) acc |= a[i] | a[i + 1];
#pragma unroll
                for (int i = 0; i < 32; i += 2) acc |= b[i] | b[i + 1];
            }
        }
    }
    const long long t1 = clock64();
    if (acc == 0x12345678u) sink[0] = acc;
    if (threadIdx.x == 0) cycles[blockIdx.x] = (unsigned long long)(t1 - t0);
    fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(s_base, 512);
}

// ---- probe 2: kind::i8 MMA ------------------------------------------------------------------------------
static constexpr int PM = 128, PN = 256, PK = 32;

__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    // UMMA shared-memory descriptor, SWIZZLE_NONE, K-major (cute::UMMA::SmemDescriptor):
    //   [0,14) start>>4 | [16,30) LBO>>4 (stride between the two 16-byte K chunks) | [32,46) SBO>>4 (stride
    //   between 8-row groups) | [46,48) version = 1 | [61,64) layout = 0
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}

__device__ __forceinline__ void mma_i8(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void mma_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}

__global__ void __launch_bounds__(128) k_mma_i8(const int8_t *A, const int8_t *B, int32_t *D, int n_mma, unsigned long long *cycles,
                                                int *status) {
    __shared__ __align__(128) int8_t sA[PM * PK];      // [k/16][row][16]
    __shared__ __align__(128) int8_t sB[PN * PK];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t s_base;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < PM * PK; i += 128) { const int r = i / PK, k = i % PK; sA[(k / 16) * (PM * 16) + r * 16 + (k % 16)] = A[i]; }
    for (int i = tid; i < PN * PK; i += 128) { const int r = i / PK, k = i % PK; sB[(k / 16) * (PN * 16) + r * 16 + (k % 16)] = B[i]; }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy smem writes -> visible to the tensor core
    if (warp == 0) tmem_alloc(&s_base, 256);
    fence_before();
    __syncthreads();
    fence_after();
    const uint32_t tmem = s_base;
    // instruction descriptor (cute::UMMA::InstrDescriptor): c_format S32=2 @4, a/b format INT8=1 @7/@10,
    // K-major A and B, N>>3 @17, M>>4 @24
    const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(PN >> 3) << 17) | ((uint32_t)(PM >> 4) << 24);
    long long t0 = 0, t1 = 0;
    if (tid == 0) {
        const uint64_t da = make_desc(smem_u32(sA), PM * 16, 128);
        const uint64_t db = make_desc(smem_u32(sB), PN * 16, 128);
        t0 = clock64();
        for (int i = 0; i < n_mma; i++) mma_i8(tmem, da, db, idesc, i > 0 ? 1u : 0u);
        mma_commit(&bar);
    }
    int spins = 0;
    bool done = false;
    while (!(done = mbar_try(&bar, 0)) && spins < (1 << 22)) spins++;
    if (tid == 0) { t1 = clock64(); cycles[0] = (unsigned long long)(t1 - t0); status[0] = done ? 1 : -1; }
    fence_after();
    if (done) {
        uint32_t r[32];
        for (int col = 0; col < PN; col += 32) {
            TMEM_LD_X32(r, tmem + ((uint32_t)(warp * 32) << 16) + col);
            tmem_wait_ld();
#pragma unroll
            for (int i = 0; i < 32; i++) D[(size_t)tid * PN + col + i] = (int32_t)r[i];
        }
    }
    fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 256);
}

// ---- probe 3: issue cost of the K3b pattern: per "tile" n_ks accumulating MMAs (M=128, N=ncols) + n_commit commits ----
__global__ void __launch_bounds__(128) k_issue_pattern(int n_tiles, int ncols, int n_ks, int n_commit, int vary_desc, unsigned long long *out, int poll_mode = 0) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t bars[8];
    __shared__ uint32_t s_base;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < 64 * 1024 / 4; i += 128) reinterpret_cast<uint32_t *>(smem)[i] = 0x01010101u;
    if (tid == 0) {
        for (int b = 0; b < 8; b++) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bars[b])) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&bars[6])) : "memory");   // bars[6]: phase 0 complete
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (warp == 0) tmem_alloc(&s_base, 512);
    fence_before();
    __syncthreads();
    fence_after();
    if (tid == 0) {
        const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(ncols >> 3) << 17) | ((uint32_t)(PM >> 4) << 24);
        const uint64_t da0 = make_desc(smem_u32(smem), PM * 16, 128);
        const uint64_t db0 = make_desc(smem_u32(smem) + 16384, ncols * 16, 128);
        long long t_issue = 0, t_poll = 0;
        const long long t0 = clock64();
        int stage = 0;
        for (int i = 0; i < n_tiles; i++) {
            const uint32_t d = s_base + (uint32_t)(i & 1) * 256;
            const uint64_t db = vary_desc ? db0 + (uint64_t)(stage * 256) : db0;
            const long long ti = clock64();
            for (int ks = 0; ks < n_ks; ks++)
                mma_i8(d, da0 + (uint64_t)(ks * 256), db + (uint64_t)(ks * 64), idesc, ks > 0 ? 1u : 0u);
            for (int c = 0; c < n_commit; c++) mma_commit(&bars[(i * 2 + c) % 6]);
            t_issue += clock64() - ti;
            if (poll_mode == 1) { const long long tp = clock64(); (void)mbar_try(&bars[6], 0); t_poll += clock64() - tp; }
            if (poll_mode == 2) { const long long tp = clock64(); volatile uint32_t *f = (volatile uint32_t *)&s_base; (void)*f; t_poll += clock64() - tp; }
            if (++stage == 4) stage = 0;
        }
        const long long t1 = clock64();
        mma_commit(&bars[7]);
        out[0] = (unsigned long long)(t1 - t0);
        out[1] = (unsigned long long)t_issue;
        out[2] = (unsigned long long)t_poll;
    }
    __syncthreads();
    // drain: wait generously for the tensor pipe before freeing TMEM
    if (tid == 0) { long long t = clock64(); while (clock64() - t < 4000000) { } }
    fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(s_base, 512);
}

int main() {
    int dev = 0, sms = 0, khz = 0;
    CK(cudaSetDevice(dev));
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    CK(cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev));
    printf("SMs %d, clock attr %d kHz\n", sms, khz);

    // ---- probe 1
    unsigned long long *d_cyc; uint32_t *d_sink;
    CK(cudaMalloc(&d_cyc, sizeof(unsigned long long) * 1024));
    CK(cudaMalloc(&d_sink, 64));
    for (int inflight : {2, 4}) {
        for (int grid : {1, sms}) {
            const int iters = 2000;
            k_tmem_bw<<<grid, 128>>>(d_cyc, d_sink, iters, inflight);
            CK(cudaDeviceSynchronize());
            std::vector<unsigned long long> c(grid);
            CK(cudaMemcpy(c.data(), d_cyc, sizeof(unsigned long long) * grid, cudaMemcpyDeviceToHost));
            double avg = 0; for (auto v : c) avg += (double)v; avg /= grid;
            const double bytes = 4.0 * 32 * 512 * 4 * iters;          // 4 warps x 32 lanes x 512 cols x 4 B
            printf("tmem_read: grid %4d, %d x32 loads in flight: %.1f cycles per 256 KB sweep -> %.1f B/clk/SM\n", grid, inflight,
                   avg / iters, bytes / avg);
        }
    }

    for (int nwarps : {4, 8}) {
        const int iters = 2000;
        k_tmem_bw_pack<<<sms, 256>>>(d_cyc, d_sink, iters, nwarps);
        CK(cudaDeviceSynchronize());
        std::vector<unsigned long long> c(sms);
        CK(cudaMemcpy(c.data(), d_cyc, sizeof(unsigned long long) * sms, cudaMemcpyDeviceToHost));
        double avg = 0; for (auto v : c) avg += (double)v; avg /= sms;
        const double cols = (double)nwarps * 32 * 512 * iters;             // lane-columns read
        printf("tmem_read pack::16b: %d warps, 2 loads (64 cols each) in flight: %.1f cycles per sweep -> %.1f lane-columns/clk/SM (= %.1f B/clk/SM of 32-bit cells)\n",
               nwarps, avg / iters, cols / avg, 4.0 * cols / avg);
    }
    for (int inflight : {4}) {
        // 8 warps (two per lane quadrant) with unpacked loads, for comparison
    }

    // ---- probe 3
    CK(cudaFuncSetAttribute(k_issue_pattern, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    struct { int ncols, n_ks, n_commit, vary; } pats[] = {{128, 3, 2, 1}, {128, 3, 2, 0}, {128, 3, 1, 1}, {128, 3, 0, 1}, {256, 3, 2, 1}, {128, 1, 2, 1}, {128, 6, 2, 1}, {64, 3, 2, 1}};
    for (auto &pt : pats) {
        const int n_tiles = 2000;
        k_issue_pattern<<<1, 128, 80 * 1024>>>(n_tiles, pt.ncols, pt.n_ks, pt.n_commit, pt.vary, d_cyc);
        CK(cudaDeviceSynchronize());
        unsigned long long c[2]; CK(cudaMemcpy(c, d_cyc, 16, cudaMemcpyDeviceToHost));
        printf("issue pattern N=%3d, %d MMAs + %d commits per tile, %s descriptors: %.1f clk per tile (issue section %.1f)\n", pt.ncols, pt.n_ks,
               pt.n_commit, pt.vary ? "varying" : "constant", (double)c[0] / n_tiles, (double)c[1] / n_tiles);
    }
    for (int poll = 1; poll <= 2; poll++) for (int ncm = 0; ncm <= 1; ncm++) {
        const int n_tiles = 2000;
        k_issue_pattern<<<1, 128, 80 * 1024>>>(n_tiles, 128, 3, ncm, 1, d_cyc, poll);
        CK(cudaDeviceSynchronize());
        unsigned long long c[3]; CK(cudaMemcpy(c, d_cyc, 24, cudaMemcpyDeviceToHost));
        printf("after 3 MMAs + %d commit: %s on an already-complete barrier/flag costs %.1f clk (tile %.1f, issue %.1f)\n", ncm,
               poll == 1 ? "mbarrier.try_wait" : "ld.volatile.shared", (double)c[2] / n_tiles, (double)c[0] / n_tiles, (double)c[1] / n_tiles);
    }

    // ---- probe 2
    std::vector<int8_t> hA(PM * PK), hB(PN * PK);
    srand(1);
    for (auto &v : hA) v = (int8_t)(rand() % 7 - 3);
    for (auto &v : hB) v = (int8_t)(rand() % 5 - 2);
    int8_t *dA, *dB; int32_t *dD; int *d_status;
    CK(cudaMalloc(&dA, hA.size())); CK(cudaMalloc(&dB, hB.size())); CK(cudaMalloc(&dD, sizeof(int32_t) * PM * PN)); CK(cudaMalloc(&d_status, 4));
    CK(cudaMemcpy(dA, hA.data(), hA.size(), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, hB.data(), hB.size(), cudaMemcpyHostToDevice));
    CK(cudaMemset(dD, 0xEE, sizeof(int32_t) * PM * PN));
    k_mma_i8<<<1, 128>>>(dA, dB, dD, 1, d_cyc, d_status);
    CK(cudaDeviceSynchronize());
    int status = 0; CK(cudaMemcpy(&status, d_status, 4, cudaMemcpyDeviceToHost));
    std::vector<int32_t> hD(PM * PN);
    CK(cudaMemcpy(hD.data(), dD, sizeof(int32_t) * PM * PN, cudaMemcpyDeviceToHost));
    long bad = 0; int first_bad = -1;
    for (int i = 0; i < PM; i++) for (int j = 0; j < PN; j++) {
        int ref = 0; for (int k = 0; k < PK; k++) ref += (int)hA[i * PK + k] * (int)hB[j * PK + k];
        if (hD[i * PN + j] != ref) { if (first_bad < 0) first_bad = i * PN + j; bad++; }
    }
    printf("mma_i8 128x256x32: barrier %s, mismatches %ld of %d", status == 1 ? "completed" : "TIMED OUT", bad, PM * PN);
    if (first_bad >= 0) printf(" (first at row %d col %d: got %d)", first_bad / PN, first_bad % PN, hD[first_bad]);
    printf("\n");
    if (status == 1) {
        for (int n_mma : {64, 1024}) {
            k_mma_i8<<<1, 128>>>(dA, dB, dD, n_mma, d_cyc, d_status);
            CK(cudaDeviceSynchronize());
            unsigned long long cyc = 0; CK(cudaMemcpy(&cyc, d_cyc, 8, cudaMemcpyDeviceToHost));
            printf("mma_i8 throughput: %d MMAs (128x256x32) in %llu cycles -> %.1f cycles/MMA, %.0f MAC/clk/SM\n", n_mma, cyc,
                   (double)cyc / n_mma, (double)PM * PN * PK * n_mma / (double)cyc);
        }
    }
    return 0;
}
