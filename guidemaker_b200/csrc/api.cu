// api.cu -- library lifecycle, error reporting, profiling hooks and the register-resident
// instruction-rate microbenchmarks that serve as roofline denominators for the INT pipes.
#include <stdarg.h>
#include <stdlib.h>
#include <chrono>
#include <mutex>
#include <thread>
#include <unordered_map>
#include <vector>

#include "common.cuh"

namespace gm { int microbench_mma_i8(int variant, double *ops_per_s); }

namespace gm {

static thread_local char g_err[512] = "";
static int g_device = -1;
static int g_sm_count = 0;
static int g_cc_major = 0, g_cc_minor = 0;
static long long g_mem = 0;
static Prof g_prof;

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
}

int cuda_fail(cudaError_t e, const char *what, const char *file, int line) {
    set_error("CUDA error %d (%s) at %s:%d in `%s`", (int)e, cudaGetErrorString(e), file, line, what);
    if (e == cudaErrorMemoryAllocation) return GM_ERR_NOMEM;
    return GM_ERR_CUDA;
}

Prof &prof() { return g_prof; }

bool trace_on() {
    static int on = -1;
    if (on < 0) { const char *e = getenv("GM_TRACE"); on = (e && e[0] == '1') ? 1 : 0; }
    return on == 1;
}
double now_ms() {
    return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}
void trace(const char *what, double t0_ms) {
    if (trace_on()) fprintf(stderr, "[gm_trace] %-28s %9.3f ms\n", what, now_ms() - t0_ms);
}
// ---- caching device allocator -------------------------------------------------------------------------------------------
struct DevBlock { void *p; size_t bytes; cudaEvent_t ev; cudaStream_t st; };
static std::mutex g_mem_mu;
static std::vector<DevBlock> g_mem_free;
static std::unordered_map<void *, DevBlock> g_mem_live;

static size_t round_bytes(size_t b) {
    if (b < 512) b = 512;
    const size_t g = b >= (1u << 20) ? (size_t)2 << 20 : 512;            // 2 MB steps for large blocks, 512 B for small ones
    return (b + g - 1) / g * g;
}

static void mem_trim_locked() {
    for (auto &b : g_mem_free) { cudaEventSynchronize(b.ev); cudaEventDestroy(b.ev); cudaFree(b.p); }
    g_mem_free.clear();
}

cudaError_t dev_alloc(void **p, size_t bytes, cudaStream_t st) {
    bytes = round_bytes(bytes);
    {
        std::lock_guard<std::mutex> lk(g_mem_mu);
        int best = -1;
        const size_t limit = bytes + (bytes >> 1) + ((size_t)4 << 20);   // do not burn a much larger block on a small request
        for (int i = 0; i < (int)g_mem_free.size(); i++)
            if (g_mem_free[i].bytes >= bytes && g_mem_free[i].bytes <= limit && (best < 0 || g_mem_free[i].bytes < g_mem_free[best].bytes)) best = i;
        if (best >= 0) {
            DevBlock b = g_mem_free[best];
            g_mem_free.erase(g_mem_free.begin() + best);
            if (b.st != st) {                                            // released on another stream: order behind that release
                cudaError_t e = cudaStreamWaitEvent(st, b.ev, 0);
                if (e != cudaSuccess) { g_mem_free.push_back(b); return e; }
            }
            b.st = st;
            g_mem_live[b.p] = b;
            *p = b.p;
            return cudaSuccess;
        }
    }
    DevBlock b;
    b.bytes = bytes;
    b.st = st;
    cudaError_t e = cudaMalloc(&b.p, bytes);
    if (e == cudaErrorMemoryAllocation) {                                // give the cached blocks back and try once more
        cudaGetLastError();
        { std::lock_guard<std::mutex> lk(g_mem_mu); mem_trim_locked(); }
        e = cudaMalloc(&b.p, bytes);
    }
    if (e != cudaSuccess) return e;
    e = cudaEventCreateWithFlags(&b.ev, cudaEventDisableTiming);
    if (e != cudaSuccess) { cudaFree(b.p); return e; }
    std::lock_guard<std::mutex> lk(g_mem_mu);
    g_mem_live[b.p] = b;
    *p = b.p;
    return cudaSuccess;
}

void dev_free(void *p, cudaStream_t st) {
    if (!p) return;
    std::lock_guard<std::mutex> lk(g_mem_mu);
    auto it = g_mem_live.find(p);
    if (it == g_mem_live.end()) { cudaFree(p); return; }                 // not ours (should not happen)
    DevBlock b = it->second;
    g_mem_live.erase(it);
    b.st = st;
    cudaEventRecord(b.ev, st);                                           // the block is reusable once the work queued so far is done
    g_mem_free.push_back(b);
}

void prefault(void *p, size_t bytes) {
    if (!p || bytes < (8u << 20)) return;
    unsigned nt = std::thread::hardware_concurrency();
    nt = nt == 0 ? 4 : (nt > 16 ? 16 : nt);
    const size_t page = 4096, slice = ((bytes / nt) + page - 1) / page * page;
    std::vector<std::thread> th;
    for (unsigned t = 0; t < nt; t++) {
        const size_t lo = (size_t)t * slice, hi = lo + slice < bytes ? lo + slice : bytes;
        if (lo >= hi) break;
        th.emplace_back([=]() {
            volatile char *c = (volatile char *)p;
            for (size_t o = lo; o < hi; o += page) c[o] = 0;
            c[hi - 1] = 0;
        });
    }
    for (auto &x : th) x.join();
}
int device_sm_count() { return g_sm_count; }
bool initialised() { return g_device >= 0; }

int ensure_init() {
    if (initialised()) {
        // other threads of the same process must bind the same device
        cudaError_t e = cudaSetDevice(g_device);
        if (e != cudaSuccess) return cuda_fail(e, "cudaSetDevice", __FILE__, __LINE__);
        return GM_OK;
    }
    return gm_init(0);
}

}  // namespace gm

using namespace gm;

extern "C" int gm_version(void) { return 100; }

extern "C" const char *gm_last_error(void) { return g_err; }

extern "C" int gm_init(int device) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n <= 0) {
        set_error("no CUDA device available (%s); libgm_b200 has no CPU fallback",
                  e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
        return GM_ERR_NODEV;
    }
    GM_ARG(device >= 0 && device < n, "gm_init: device %d out of range [0,%d)", device, n);
    cudaDeviceProp p;
    GM_CUDA(cudaGetDeviceProperties(&p, device));
    if (p.major != 10) {
        set_error("device %d (%s) is sm_%d%d; this library is built for sm_100a only", device, p.name, p.major, p.minor);
        return GM_ERR_NODEV;
    }
    GM_CUDA(cudaSetDevice(device));
    GM_CUDA(cudaFree(0));
    g_device = device;
    g_sm_count = p.multiProcessorCount;
    g_cc_major = p.major;
    g_cc_minor = p.minor;
    g_mem = (long long)p.totalGlobalMem;
    return GM_OK;
}

extern "C" int gm_trim(void) {
    std::lock_guard<std::mutex> lk(g_mem_mu);
    mem_trim_locked();
    return GM_OK;
}

extern "C" int gm_device_info(int *sm_count, int *cc_major, int *cc_minor, int64_t *mem_bytes) {
    int rc = ensure_init();
    if (rc) return rc;
    if (sm_count) *sm_count = g_sm_count;
    if (cc_major) *cc_major = g_cc_major;
    if (cc_minor) *cc_minor = g_cc_minor;
    if (mem_bytes) *mem_bytes = g_mem;
    return GM_OK;
}

// ---- profiling -------------------------------------------------------------------------------------

static int prof_drain() {
    Prof &p = g_prof;
    for (int i = 0; i < p.n_ev; i++) {
        float ms = 0.f;
        GM_CUDA(cudaEventSynchronize(p.ev[i][1]));
        GM_CUDA(cudaEventElapsedTime(&ms, p.ev[i][0], p.ev[i][1]));
        p.scan_ms_done += ms;
    }
    p.n_ev = 0;
    return GM_OK;
}

namespace gm {
// called by knn.cu around the pair-scan kernel
int prof_begin(cudaStream_t s) {
    Prof &p = g_prof;
    if (!p.on) return -1;
    if (p.n_ev == Prof::MAXEV) {
        if (prof_drain()) return -1;
    }
    int i = p.n_ev;
    if (i >= p.n_alloc) {
        if (cudaEventCreate(&p.ev[i][0]) != cudaSuccess || cudaEventCreate(&p.ev[i][1]) != cudaSuccess) return -1;
        p.n_alloc = i + 1;
    }
    cudaEventRecord(p.ev[i][0], s);
    return i;
}
void prof_end(int slot, cudaStream_t s, double pairs) {
    Prof &p = g_prof;
    p.scan_launches++;
    p.pairs += pairs;
    if (slot < 0) return;
    cudaEventRecord(p.ev[slot][1], s);
    p.n_ev = slot + 1;
}
}  // namespace gm

extern "C" int gm_prof_enable(int on) {
    g_prof.on = on != 0;
    return GM_OK;
}

extern "C" int gm_prof_reset(void) {
    int rc = prof_drain();
    if (rc) return rc;
    g_prof.all_launches = g_prof.scan_launches = 0;
    g_prof.pairs = 0.0;
    g_prof.scan_ms_done = 0.0;
    return GM_OK;
}

extern "C" int gm_prof_read(double *scan_kernel_ms, int64_t *scan_kernel_launches, double *pairs,
                            int64_t *all_kernel_launches) {
    int rc = prof_drain();
    if (rc) return rc;
    if (scan_kernel_ms) *scan_kernel_ms = g_prof.scan_ms_done;
    if (scan_kernel_launches) *scan_kernel_launches = g_prof.scan_launches;
    if (pairs) *pairs = g_prof.pairs;
    if (all_kernel_launches) *all_kernel_launches = g_prof.all_launches;
    return GM_OK;
}

// ---- instruction-rate microbenchmarks ------------------------------------------------------------
// Eight independent register-resident chains per thread, no memory traffic; the achieved
// lane-ops/s of the whole GPU is the denominator of the INT-pipe roofline.

static constexpr int MB_CHAINS = 8;

template <int WHAT>
__global__ void __launch_bounds__(256) mb_kernel(uint32_t *out, int iters) {
    uint32_t a[MB_CHAINS];
#pragma unroll
    for (int i = 0; i < MB_CHAINS; i++) a[i] = threadIdx.x * 2654435761u + i * 40503u + blockIdx.x;
    const uint32_t c = out[0] | 0x9E3779B1u;   // opaque to the compiler
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < MB_CHAINS; i++) {
            if (WHAT == 0) {
                a[i] = __popc(a[i]) + c;       // one POPC (+ one IADD on the ALU pipe) per step
            } else if (WHAT == 1) {
                asm volatile("lop3.b32 %0, %0, %1, %2, 0xE8;" : "+r"(a[i]) : "r"(a[(i + 1) % MB_CHAINS]), "r"(c));
            } else {
                a[i] = a[i] * c + a[(i + 1) % MB_CHAINS];
            }
        }
    }
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < MB_CHAINS; i++) s ^= a[i];
    if (s == 0x12345u) out[1] = s;             // practically never; keeps the chains alive
}

extern "C" int gm_microbench(int what, double *ops_per_s) {
    int rc = ensure_init();
    if (rc) return rc;
    GM_ARG(what >= 0 && what <= 11 && ops_per_s, "gm_microbench: what must be 0..11");
    if (what >= 3) return microbench_mma_i8(what - 3, ops_per_s);
    uint32_t *d = nullptr;
    GM_CUDA(dev_alloc((void **)&d, 64, 0));
    GM_CUDA(cudaMemsetAsync(d, 0, 64, 0));
    const int grid = g_sm_count * 8, iters = 8192;
    cudaEvent_t e0, e1;
    GM_CUDA(cudaEventCreate(&e0));
    GM_CUDA(cudaEventCreate(&e1));
    float best = 1e30f;
    for (int rep = 0; rep < 4; rep++) {
        GM_CUDA(cudaEventRecord(e0));
        if (what == 0) mb_kernel<0><<<grid, 256>>>(d, iters);
        else if (what == 1) mb_kernel<1><<<grid, 256>>>(d, iters);
        else mb_kernel<2><<<grid, 256>>>(d, iters);
        count_launch();
        GM_CUDA(cudaEventRecord(e1));
        GM_CUDA(cudaEventSynchronize(e1));
        float ms;
        GM_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        if (rep > 0 && ms < best) best = ms;
    }
    GM_CUDA(cudaGetLastError());
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    dev_free(d, 0);
    *ops_per_s = (double)grid * 256.0 * iters * MB_CHAINS / (best * 1e-3);
    return GM_OK;
}
