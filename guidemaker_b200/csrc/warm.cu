// warm.cu -- neighbourhood warm start of the exact Hamming scan (K3b).
//
// The pair scan is exact whatever bound a query starts with, as long as the bound is valid: the k-th smallest distance
// over ANY k distinct guides of the table is an upper bound of the query's final k-th distance.  The tighter it is, the
// fewer candidates the tensor-core filter flags while the lists are still loose (the start-up phase of K3b: 12 % of a
// CTA's life on the 6.3 Mb table with a bound taken from the first 8192 guides).  Guides that share a long run of bases
// with the query are much better than a random sample, and sorting finds them: the index keeps three copies of the table
// sorted by the guide read as a base-4 number, with its positions rotated by 0, L/3 and 2L/3, so a query's neighbours in
// copy c share the run of bases that ends at the rotation point.  Per query batch and copy: sort the
// queries the same way (radix sort of 2L-bit keys), locate every query in the sorted table (binary search) and scan the
// W guides around that rank, keeping the k best of all copies (without duplicates) in shared memory.  Queries that are adjacent in sorted order have overlapping windows, so a CTA of 256
// queries stages the union of its windows in shared memory once.  The result is an ordinary [q][k] list of
// (distance << 27 | guide index) keys; only the distance of its last entry is used (knn_tc.cu: inclusive bound).
//
// The reference has no counterpart (NMSLib's HNSW graph plays this role approximately: core.py:418-523).
#include <cub/cub.cuh>
#include <stdlib.h>

#include "knn_common.cuh"

namespace gm {

static constexpr int WARM_THREADS = 256;
static constexpr int WARM_CAP = 3072;           // guides a CTA can stage: 36 KB of shared memory
int warm_copies() { const char *e = getenv("GM_WARM_COPIES"); const int c = e ? atoi(e) : 3; return c < 1 ? 1 : c > 4 ? 4 : c; }

// positions rotated right by h within the L-bit planes (a common permutation of the positions keeps Hamming distances)
__device__ __forceinline__ uint2 warm_rot(uint2 p, int h, int L) {
    if (h == 0) return p;
    const uint32_t m = (1u << L) - 1u;
    return make_uint2(((p.x >> h) | (p.x << (L - h))) & m, ((p.y >> h) | (p.y << (L - h))) & m);
}

__global__ void warm_keys_kernel(const uint2 *__restrict__ planes, int64_t n, int h, int L, uint64_t *__restrict__ keys,
                                 uint32_t *__restrict__ ids) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint2 p = warm_rot(planes[i], h, L);
    keys[i] = from_planes(p.x, p.y);
    ids[i] = (uint32_t)i;
}

__global__ void warm_gather_kernel(const uint2 *__restrict__ planes, const uint32_t *__restrict__ ids, int64_t n, int h, int L,
                                   uint2 *__restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = warm_rot(planes[ids[i]], h, L);
}

// Insert `key` into the thread's list (shared memory, stride WARM_THREADS between ranks) unless the list already holds it
// (the sorted copies show a query some guides twice).  Returns the distance bound for the next test: insert candidates
// have distance <= bound.
static __device__ __forceinline__ uint32_t warm_insert(uint32_t *lst, int k, uint32_t key) {
    if (key < lst[(k - 1) * WARM_THREADS]) {
        int pos = k - 1;
        while (pos > 0 && lst[(pos - 1) * WARM_THREADS] > key) pos--;
        if (pos == 0 || lst[(pos - 1) * WARM_THREADS] != key) {
            for (int j = k - 1; j > pos; j--) lst[j * WARM_THREADS] = lst[(j - 1) * WARM_THREADS];
            lst[pos * WARM_THREADS] = key;
        }
    }
    const uint32_t worst = lst[(k - 1) * WARM_THREADS];
    return worst == KEY_EMPTY ? 32u : worst >> IDX_BITS;
}

// dynamic shared memory: WARM_CAP staged guides (planes, indices) + the CTA's lists [k][WARM_THREADS]
static size_t warm_smem_bytes(int k) { return (size_t)WARM_CAP * 12 + (size_t)k * WARM_THREADS * 4; }

__global__ void __launch_bounds__(WARM_THREADS) warm_window_kernel(const uint2 *__restrict__ sp, const uint32_t *__restrict__ si, int n,
                                                                   const uint2 *__restrict__ qplanes, const uint32_t *__restrict__ sq,
                                                                   int64_t q, int k, int W, int h, int L, uint32_t *__restrict__ wlists) {
    extern __shared__ __align__(16) uint8_t warm_smem[];
    uint2 *s_p = reinterpret_cast<uint2 *>(warm_smem);
    uint32_t *s_i = reinterpret_cast<uint32_t *>(s_p + WARM_CAP);
    uint32_t *s_l = s_i + WARM_CAP;
    __shared__ int s_lo, s_hi;
    const int tid = threadIdx.x;
    const int64_t i = (int64_t)blockIdx.x * WARM_THREADS + tid;
    const bool active = i < q;
    if (tid == 0) { s_lo = n; s_hi = 0; }
    __syncthreads();
    uint32_t qi = 0;
    uint2 p = make_uint2(0u, 0u);
    int lo = 0, hi = 0;
    uint32_t *lst = s_l + tid;
    if (active) {
        qi = sq[i];
        p = warm_rot(qplanes[qi], h, L);
        for (int j = 0; j < k; j++) lst[j * WARM_THREADS] = wlists[(size_t)qi * k + j];     // the previous copy's result
        const uint64_t key = from_planes(p.x, p.y);
        int a = 0, b = n;                                         // first guide whose key is >= the query's
        while (a < b) {
            const int mid = (a + b) >> 1;
            const uint2 t = sp[mid];
            if (from_planes(t.x, t.y) < key) a = mid + 1; else b = mid;
        }
        lo = min(max(a - W / 2, 0), max(n - W, 0));
        hi = min(lo + W, n);
        atomicMin(&s_lo, lo);
        atomicMax(&s_hi, hi);
    }
    __syncthreads();
    const int base = s_lo, span = s_hi - s_lo;
    const bool staged = span <= WARM_CAP;                         // block-uniform
    if (staged) {
        for (int j = tid; j < span; j += WARM_THREADS) { s_p[j] = sp[base + j]; s_i[j] = si[base + j]; }
        __syncthreads();
    }
    if (!active) return;
    const uint32_t w0 = lst[(k - 1) * WARM_THREADS];
    uint32_t bound = w0 == KEY_EMPTY ? 32u : w0 >> IDX_BITS;
    if (staged) {
#pragma unroll 4
        for (int j = lo - base; j < hi - base; j++) {
            const uint2 t = s_p[j];
            const uint32_t d = (uint32_t)hamming_planes(p.x, p.y, t.x, t.y);
            if (d <= bound) bound = warm_insert(lst, k, (d << IDX_BITS) | s_i[j]);
        }
    } else {                                                      // sparse queries: every window straight from L2
#pragma unroll 4
        for (int j = lo; j < hi; j++) {
            const uint2 t = sp[j];
            const uint32_t d = (uint32_t)hamming_planes(p.x, p.y, t.x, t.y);
            if (d <= bound) bound = warm_insert(lst, k, (d << IDX_BITS) | si[j]);
        }
    }
    for (int j = 0; j < k; j++) wlists[(size_t)qi * k + j] = lst[j * WARM_THREADS];
}

// keys -> ids in key order (ids_out), by radix sort over the 2L significant bits
static int warm_sort(const uint2 *planes, int64_t n, int h, int L, uint32_t *ids_out, cudaStream_t st) {
    uint64_t *k_in = nullptr, *k_out = nullptr;
    uint32_t *v_in = nullptr;
    void *tmp = nullptr;
    size_t tmp_bytes = 0;
    cudaError_t e = cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, k_in, k_out, v_in, ids_out, (int)n, 0, 2 * L, st);
    if (e == cudaSuccess) e = dev_alloc((void **)&k_in, (size_t)n * 8, st);
    if (e == cudaSuccess) e = dev_alloc((void **)&k_out, (size_t)n * 8, st);
    if (e == cudaSuccess) e = dev_alloc((void **)&v_in, (size_t)n * 4, st);
    if (e == cudaSuccess) e = dev_alloc(&tmp, tmp_bytes ? tmp_bytes : 1, st);
    if (e == cudaSuccess) {
        warm_keys_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(planes, n, h, L, k_in, v_in);
        count_launch();
        e = cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, k_in, k_out, v_in, ids_out, (int)n, 0, 2 * L, st);
        count_launch(2 + (2 * L + 7) / 8);                       // histogram, exclusive sum, one onesweep pass per 8 key bits
    }
    dev_free(k_in, st); dev_free(k_out, st); dev_free(v_in, st); dev_free(tmp, st);     // stream ordered
    if (e != cudaSuccess) return cuda_fail(e, "warm_sort", __FILE__, __LINE__);
    return GM_OK;
}

// the sorted copies of the table, built on the first query that wants them (all of them or none)
static int warm_build_index(Index *ix, cudaStream_t st) {
    if (ix->sorted_p[0]) return GM_OK;
    const int64_t n = ix->n_u;
    const int C = warm_copies();
    int rc = GM_OK;
    for (int c = 0; c < C && rc == GM_OK; c++) {
        const int h = c * ix->L / C;
        cudaError_t e = dev_alloc((void **)&ix->sorted_i[c], (size_t)n * 4, st);
        if (e == cudaSuccess) e = dev_alloc((void **)&ix->sorted_p[c], (size_t)n * 8, st);
        if (e != cudaSuccess) { rc = cuda_fail(e, "warm_build_index", __FILE__, __LINE__); break; }
        rc = warm_sort(ix->planes, n, h, ix->L, ix->sorted_i[c], st);
        if (rc) break;
        warm_gather_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(ix->planes, ix->sorted_i[c], n, h, ix->L, ix->sorted_p[c]);
        count_launch();
        e = cudaGetLastError();
        if (e != cudaSuccess) rc = cuda_fail(e, "warm_gather_kernel", __FILE__, __LINE__);
    }
    if (rc) {
        for (int c = 0; c < 4; c++) {
            dev_free(ix->sorted_p[c], st); dev_free(ix->sorted_i[c], st);
            ix->sorted_p[c] = nullptr; ix->sorted_i[c] = nullptr;
        }
    }
    return rc;
}

void warm_free_index(Index *ix) {
    dev_free(ix->prefix_p, 0); dev_free(ix->prefix_i, 0);
    ix->prefix_p = nullptr; ix->prefix_i = nullptr;
    for (int c = 0; c < 4; c++) {
        dev_free(ix->sorted_p[c], 0); dev_free(ix->sorted_i[c], 0);
        ix->sorted_p[c] = nullptr; ix->sorted_i[c] = nullptr;
    }
}

// ---- K4p: the table sorted by the guides read from their FIRST base (knn.cu: knn_leven_prefix_kernel) ---------------
__global__ void prefix_keys_kernel(const uint2 *__restrict__ planes, int64_t n, int L, uint64_t *__restrict__ keys, uint32_t *__restrict__ ids) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint2 p = planes[i];                                    // bit j = base j: reverse, so that base 0 is the most significant
    keys[i] = from_planes(__brev(p.x) >> (32 - L), __brev(p.y) >> (32 - L));
    ids[i] = (uint32_t)i;
}
__global__ void prefix_gather_kernel(const uint2 *__restrict__ planes, const uint32_t *__restrict__ ids, int64_t n, int64_t n_pad,
                                     uint2 *__restrict__ out, uint32_t *__restrict__ ids_pad) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_pad) return;
    uint64_t code = 0;                                            // the kernel wants the 2-bit code (base j = bits 2j, 2j+1)
    if (i < n) { const uint2 p = planes[ids[i]]; code = from_planes(p.x, p.y); }
    out[i] = make_uint2((uint32_t)code, (uint32_t)(code >> 32));
    if (i >= n) ids_pad[i] = 0u;
}

int sorted_by_prefix(Index *ix, cudaStream_t st) {
    if (ix->prefix_p) return GM_OK;
    const int64_t n = ix->n_u, n_pad = ix->n_pad;
    uint64_t *k_in = nullptr, *k_out = nullptr;
    uint32_t *v_in = nullptr;
    void *tmp = nullptr;
    size_t tmp_bytes = 0;
    cudaError_t e = dev_alloc((void **)&ix->prefix_i, (size_t)n_pad * 4, st);
    if (e == cudaSuccess) e = dev_alloc((void **)&ix->prefix_p, (size_t)n_pad * 8, st);
    if (e == cudaSuccess) e = cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, k_in, k_out, v_in, ix->prefix_i, (int)n, 0, 2 * ix->L, st);
    if (e == cudaSuccess) e = dev_alloc((void **)&k_in, (size_t)n * 8, st);
    if (e == cudaSuccess) e = dev_alloc((void **)&k_out, (size_t)n * 8, st);
    if (e == cudaSuccess) e = dev_alloc((void **)&v_in, (size_t)n * 4, st);
    if (e == cudaSuccess) e = dev_alloc(&tmp, tmp_bytes ? tmp_bytes : 1, st);
    if (e == cudaSuccess) {
        prefix_keys_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(ix->planes, n, ix->L, k_in, v_in);
        count_launch();
        e = cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, k_in, k_out, v_in, ix->prefix_i, (int)n, 0, 2 * ix->L, st);
        count_launch(2 + (2 * ix->L + 7) / 8);
    }
    if (e == cudaSuccess) {
        prefix_gather_kernel<<<(unsigned)((n_pad + 255) / 256), 256, 0, st>>>(ix->planes, ix->prefix_i, n, n_pad, ix->prefix_p, ix->prefix_i);
        count_launch();
        e = cudaGetLastError();
    }
    dev_free(k_in, st); dev_free(k_out, st); dev_free(v_in, st); dev_free(tmp, st);
    if (e != cudaSuccess) {
        dev_free(ix->prefix_p, st); dev_free(ix->prefix_i, st);
        ix->prefix_p = nullptr; ix->prefix_i = nullptr;
        return cuda_fail(e, "sorted_by_prefix", __FILE__, __LINE__);
    }
    return GM_OK;
}

// wlists ([q_pad][k], preset to KEY_EMPTY) <- k best of the copies x W guides around every query's rank in the sorted copies
int warm_window(Index *ix, const uint2 *qplanes, int64_t q, int k, int W, uint32_t *wlists, cudaStream_t st) {
    int rc = warm_build_index(ix, st);
    if (rc) return rc;
    static bool attr_set = false;
    if (!attr_set) {
        GM_CUDA(cudaFuncSetAttribute(warm_window_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)warm_smem_bytes(GM_MAX_K)));
        attr_set = true;
    }
    uint32_t *sq = nullptr;
    GM_CUDA(dev_alloc((void **)&sq, (size_t)q * 4, st));
    const int C = warm_copies();
    for (int c = 0; c < C && rc == GM_OK; c++) {
        const int h = c * ix->L / C;
        rc = warm_sort(qplanes, q, h, ix->L, sq, st);
        if (rc) break;
        warm_window_kernel<<<(unsigned)((q + WARM_THREADS - 1) / WARM_THREADS), WARM_THREADS, warm_smem_bytes(k), st>>>(
            ix->sorted_p[c], ix->sorted_i[c], (int)ix->n_u, qplanes, sq, q, k, W, h, ix->L, wlists);
        count_launch();
    }
    dev_free(sq, st);
    if (rc) return rc;
    GM_CUDA(cudaGetLastError());
    return GM_OK;
}

}  // namespace gm
