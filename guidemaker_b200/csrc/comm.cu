// comm.cu -- the multi-GPU step of the kNN query inside the C ABI: one process per GPU, query rows sharded, the guide
// table replicated, per-rank top-k rows all-gathered over NVLink with NCCL.
//
// The reference has no multi-device path; this is the native counterpart of guidemaker_b200/sharding.py for hosts that
// are not Python/torch: the host language only has to carry 128 bytes (the NCCL unique id) from rank 0 to the other
// ranks.  NCCL is bound lazily with dlopen("libnccl.so.2") the first time a communicator is made, so single-GPU use of
// the library has no NCCL dependency (and a process that already loaded torch's NCCL shares that copy).
#include <dlfcn.h>
#include <new>

#include "knn_common.cuh"
#include "scan.cuh"

namespace gm {

// the handful of NCCL entry points used, with NCCL's own (stable) C signatures
typedef struct { char internal[128]; } nccl_id_t;
typedef void *nccl_comm_t;
typedef int (*fn_get_id)(nccl_id_t *);
typedef int (*fn_init_rank)(nccl_comm_t *, int, nccl_id_t, int);
typedef int (*fn_all_gather)(const void *, void *, size_t, int, nccl_comm_t, cudaStream_t);
typedef int (*fn_destroy)(nccl_comm_t);
typedef const char *(*fn_errstr)(int);
static const int NCCL_INT8 = 0, NCCL_INT32 = 2;      // ncclDataType_t

struct Nccl {
    void *lib = nullptr;
    fn_get_id get_id = nullptr;
    fn_init_rank init_rank = nullptr;
    fn_all_gather all_gather = nullptr;
    fn_destroy destroy = nullptr;
    fn_errstr errstr = nullptr;
};
static Nccl g_nccl;

static int nccl_load() {
    if (g_nccl.lib) return GM_OK;
    void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) { set_error("cannot load libnccl.so.2: %s", dlerror()); return GM_ERR_NODEV; }
    g_nccl.get_id = (fn_get_id)dlsym(h, "ncclGetUniqueId");
    g_nccl.init_rank = (fn_init_rank)dlsym(h, "ncclCommInitRank");
    g_nccl.all_gather = (fn_all_gather)dlsym(h, "ncclAllGather");
    g_nccl.destroy = (fn_destroy)dlsym(h, "ncclCommDestroy");
    g_nccl.errstr = (fn_errstr)dlsym(h, "ncclGetErrorString");
    if (!g_nccl.get_id || !g_nccl.init_rank || !g_nccl.all_gather || !g_nccl.destroy) {
        set_error("libnccl.so.2 lacks an expected symbol");
        dlclose(h);
        return GM_ERR_NODEV;
    }
    g_nccl.lib = h;
    return GM_OK;
}

static int nccl_fail(int rc, const char *what) {
    set_error("NCCL error %d (%s) in %s", rc, g_nccl.errstr ? g_nccl.errstr(rc) : "?", what);
    return GM_ERR_CUDA;
}

struct Comm {
    nccl_comm_t comm = nullptr;
    int rank = 0, world = 1;
};

// contiguous, balanced row range of `rank`: sizes differ by at most one (sharding.shard_bounds)
static void shard_bounds(int64_t n, int rank, int world, int64_t *lo, int64_t *hi) {
    const int64_t base = n / world, rem = n % world;
    *lo = rank * base + (rank < rem ? rank : rem);
    *hi = *lo + base + (rank < rem ? 1 : 0);
}

// kNN of this rank's shard (d_q_local = rows [lo, hi) of the q query rows, on the device) and all-gather of the fixed-size
// result rows; out_* (host) receive all q rows.
static int knn_sharded_dev(Index *ix, Comm *c, const uint64_t *d_q_local, int64_t q, int k, int32_t *out_idx, uint8_t *out_dist, cudaStream_t st) {
    int64_t lo, hi;
    shard_bounds(q, c->rank, c->world, &lo, &hi);
    const int64_t rows = (q + c->world - 1) / c->world;              // largest shard: every rank sends this many rows
    int32_t *d_idx = nullptr, *g_idx = nullptr;
    uint8_t *d_dist = nullptr, *g_dist = nullptr;
    cudaError_t e = dev_alloc((void **)&d_idx, (size_t)rows * k * 4, st);
    if (e == cudaSuccess) e = dev_alloc((void **)&d_dist, (size_t)rows * k, st);
    if (e == cudaSuccess) e = dev_alloc((void **)&g_idx, (size_t)rows * k * 4 * c->world, st);
    if (e == cudaSuccess) e = dev_alloc((void **)&g_dist, (size_t)rows * k * c->world, st);
    if (e == cudaSuccess) e = cudaMemsetAsync(d_idx, 0xFF, (size_t)rows * k * 4, st);      // padding rows: idx -1, dist 255
    if (e == cudaSuccess) e = cudaMemsetAsync(d_dist, 0xFF, (size_t)rows * k, st);
    int rc = GM_OK;
    if (e == cudaSuccess && hi > lo) rc = gm_knn_dev(ix, d_q_local, hi - lo, k, d_idx, d_dist, st);
    if (e == cudaSuccess && rc == GM_OK) {
        int n = g_nccl.all_gather(d_idx, g_idx, (size_t)rows * k, NCCL_INT32, c->comm, st);
        if (n == 0) n = g_nccl.all_gather(d_dist, g_dist, (size_t)rows * k, NCCL_INT8, c->comm, st);
        if (n != 0) rc = nccl_fail(n, "ncclAllGather");
    }
    if (e == cudaSuccess && rc == GM_OK) {
        prefault(out_idx, (size_t)q * k * 4);                       // overlaps the kernels and the collective enqueued above
        prefault(out_dist, (size_t)q * k);
        for (int r = 0; r < c->world && e == cudaSuccess; r++) {    // trim the padding of the shorter shards
            int64_t a, b;
            shard_bounds(q, r, c->world, &a, &b);
            if (b == a) continue;
            e = cudaMemcpyAsync(out_idx + a * k, g_idx + (size_t)r * rows * k, (size_t)(b - a) * k * 4, cudaMemcpyDeviceToHost, st);
            if (e == cudaSuccess) e = cudaMemcpyAsync(out_dist + a * k, g_dist + (size_t)r * rows * k, (size_t)(b - a) * k, cudaMemcpyDeviceToHost, st);
        }
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    }
    dev_free(d_idx, st); dev_free(d_dist, st); dev_free(g_idx, st); dev_free(g_dist, st);
    if (rc) return rc;
    if (e != cudaSuccess) return cuda_fail(e, "gm_knn_sharded", __FILE__, __LINE__);
    return GM_OK;
}

}  // namespace gm

using namespace gm;

extern "C" int gm_comm_unique_id(uint8_t *id128) {
    GM_ARG(id128, "gm_comm_unique_id: NULL buffer");
    int rc = nccl_load();
    if (rc) return rc;
    nccl_id_t id;
    int n = g_nccl.get_id(&id);
    if (n != 0) return nccl_fail(n, "ncclGetUniqueId");
    memcpy(id128, &id, sizeof id);
    return GM_OK;
}

extern "C" int gm_comm_create(const uint8_t *id128, int rank, int world, void **comm) {
    int rc = ensure_init();
    if (rc) return rc;
    GM_ARG(id128 && comm && world >= 1 && rank >= 0 && rank < world, "gm_comm_create: bad argument");
    *comm = nullptr;
    rc = nccl_load();
    if (rc) return rc;
    Comm *c = new (std::nothrow) Comm();
    if (!c) { set_error("out of host memory"); return GM_ERR_NOMEM; }
    nccl_id_t id;
    memcpy(&id, id128, sizeof id);
    int n = g_nccl.init_rank(&c->comm, world, id, rank);
    if (n != 0) { delete c; return nccl_fail(n, "ncclCommInitRank"); }
    c->rank = rank;
    c->world = world;
    *comm = c;
    return GM_OK;
}

extern "C" int gm_comm_free(void *comm) {
    Comm *c = (Comm *)comm;
    if (!c) return GM_OK;
    cudaDeviceSynchronize();
    if (c->comm && g_nccl.destroy) g_nccl.destroy(c->comm);
    delete c;
    return GM_OK;
}

extern "C" int gm_knn_sharded(void *index, void *comm, const uint64_t *q2bit, int64_t q, int k, int32_t *out_idx, uint8_t *out_dist) {
    int rc = ensure_init();
    if (rc) return rc;
    Index *ix = (Index *)index;
    Comm *c = (Comm *)comm;
    GM_ARG(ix && c && q >= 0 && k >= 1 && k <= GM_MAX_K, "gm_knn_sharded: bad argument");
    if (q == 0) return GM_OK;
    GM_ARG(q2bit && out_idx && out_dist, "gm_knn_sharded: NULL buffer");
    int64_t lo, hi;
    shard_bounds(q, c->rank, c->world, &lo, &hi);
    uint64_t *d_q = nullptr;                                        // only this rank's rows cross PCIe
    cudaError_t e = dev_alloc((void **)&d_q, (size_t)(hi - lo + 1) * 8, 0);
    if (e == cudaSuccess && hi > lo) e = cudaMemcpyAsync(d_q, q2bit + lo, (size_t)(hi - lo) * 8, cudaMemcpyHostToDevice, 0);
    if (e != cudaSuccess) { dev_free(d_q, 0); return cuda_fail(e, "gm_knn_sharded", __FILE__, __LINE__); }
    rc = knn_sharded_dev(ix, c, d_q, q, k, out_idx, out_dist, 0);
    dev_free(d_q, 0);
    return rc;
}
