// dedup.cu -- K2: keep-first duplicate detection over packed keys on the GPU.
//
// Replaces pandas' khash `Series.duplicated()` (core.py:416) for the seed region and Python's
// `list(set(...))` (core.py:446) for the index.  An open-addressing table of 64-bit keys is filled
// with atomicCAS; every slot also keeps, by atomicMin, the smallest row that carries its key.  A row
// is a duplicate iff that minimum is not itself.  The result is independent of the order in which
// threads win the CAS races, so it is bit-exact and deterministic.
//
// HBM traffic per row: one 8-byte key read, ~1.5 probes of a 12-byte slot in each pass (table load
// factor <= 0.5, random access -> 32-byte sectors), one output write.
#include "scan.cuh"

namespace gm {

static constexpr unsigned long long SLOT_EMPTY = 0xFFFFFFFFFFFFFFFFULL;   // keys use <= 54 bits

__device__ __forceinline__ uint64_t mix64(uint64_t x) {
    x ^= x >> 33; x *= 0xff51afd7ed558ccdULL;
    x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL;
    x ^= x >> 33;
    return x;
}

__device__ __forceinline__ uint64_t seed_key_of(uint64_t g, int L, int lsr, int five_prime) {
    // first lsr bases (5prime) or last lsr bases (3prime); whole guide when lsr == 0 (core.py:402-412)
    if (lsr == 0 || lsr >= L) return g;
    if (five_prime) return g & ((1ULL << (2 * lsr)) - 1ULL);
    return g >> (2 * (L - lsr));
}

__global__ void dedup_insert_kernel(const uint64_t *__restrict__ keys_in, int64_t n, int L, int lsr, int five_prime,
                                    unsigned long long *__restrict__ slots, unsigned int *__restrict__ minrow, uint64_t cap_mask) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const unsigned long long key = seed_key_of(keys_in[i], L, lsr, five_prime);
    uint64_t h = mix64(key) & cap_mask;
    while (true) {
        const unsigned long long prev = atomicCAS(&slots[h], SLOT_EMPTY, key);
        if (prev == SLOT_EMPTY || prev == key) {
            atomicMin(&minrow[h], (unsigned int)i);
            return;
        }
        h = (h + 1) & cap_mask;
    }
}

__global__ void dedup_lookup_kernel(const uint64_t *__restrict__ keys_in, int64_t n, int L, int lsr, int five_prime,
                                    const unsigned long long *__restrict__ slots, const unsigned int *__restrict__ minrow,
                                    uint64_t cap_mask, uint8_t *__restrict__ is_dup, int64_t *__restrict__ first_row,
                                    int32_t *__restrict__ first32) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const unsigned long long key = seed_key_of(keys_in[i], L, lsr, five_prime);
    uint64_t h = mix64(key) & cap_mask;
    while (slots[h] != key) h = (h + 1) & cap_mask;          // the key was inserted by the first pass
    const unsigned int r = minrow[h];
    if (is_dup) is_dup[i] = r != (unsigned int)i;
    if (first_row) first_row[i] = (int64_t)r;
    if (first32) first32[i] = (int32_t)r;
}

// Device-resident form: keys, flags and first rows stay in HBM; everything is enqueued on `st` (no synchronisation).
// d_first32[i] = smallest row with the key of row i (rows < 2^31).
int dedup_dev(const uint64_t *d_keys, int64_t n, int L, int lsr, int five_prime, uint8_t *d_is_dup, int64_t *d_first_row,
              int32_t *d_first32, cudaStream_t st) {
    GM_ARG(n >= 0, "dedup: negative row count");
    if (n == 0) return GM_OK;
    GM_ARG(d_keys && (d_is_dup || d_first_row || d_first32), "dedup: NULL buffer");
    GM_ARG(L >= 1 && L <= GM_MAX_L && lsr >= 0 && lsr <= GM_MAX_L, "dedup: L=%d lsr=%d out of range", L, lsr);
    if (n >= (1LL << 31)) { set_error("dedup: %lld rows exceed 2^31", (long long)n); return GM_ERR_RANGE; }
    uint64_t cap = 1024;
    while (cap < (uint64_t)n * 2) cap <<= 1;
    unsigned long long *d_slots = nullptr;
    unsigned int *d_minrow = nullptr;
    cudaError_t e = dev_alloc((void **)&d_slots, cap * 8, st);
    if (e == cudaSuccess) e = dev_alloc((void **)&d_minrow, cap * 4, st);
    if (e == cudaSuccess) e = cudaMemsetAsync(d_slots, 0xFF, cap * 8, st);
    if (e == cudaSuccess) e = cudaMemsetAsync(d_minrow, 0xFF, cap * 4, st);
    if (e == cudaSuccess) {
        const unsigned grid = (unsigned)((n + 255) / 256);
        dedup_insert_kernel<<<grid, 256, 0, st>>>(d_keys, n, L, lsr, five_prime, d_slots, d_minrow, cap - 1);
        dedup_lookup_kernel<<<grid, 256, 0, st>>>(d_keys, n, L, lsr, five_prime, d_slots, d_minrow, cap - 1, d_is_dup, d_first_row, d_first32);
        count_launch(2);
        e = cudaGetLastError();
    }
    dev_free(d_slots, st);
    dev_free(d_minrow, st);
    if (e != cudaSuccess) return cuda_fail(e, "dedup", __FILE__, __LINE__);
    return GM_OK;
}

static int dedup_run(const uint64_t *h_keys, int64_t n, int L, int lsr, int five_prime, uint8_t *h_is_dup, int64_t *h_first_row) {
    int rc = ensure_init();
    if (rc) return rc;
    GM_ARG(n >= 0, "dedup: negative row count");
    if (n == 0) return GM_OK;
    GM_ARG(h_keys && (h_is_dup || h_first_row), "dedup: NULL buffer");
    uint64_t *d_keys = nullptr;
    uint8_t *d_dup = nullptr;
    int64_t *d_first = nullptr;
    cudaStream_t st = 0;
    cudaError_t e = dev_alloc((void **)&d_keys, (size_t)n * 8, st);
    if (e == cudaSuccess && h_is_dup) e = dev_alloc((void **)&d_dup, (size_t)n, st);
    if (e == cudaSuccess && h_first_row) e = dev_alloc((void **)&d_first, (size_t)n * 8, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_keys, h_keys, (size_t)n * 8, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) rc = dedup_dev(d_keys, n, L, lsr, five_prime, d_dup, d_first, nullptr, st);
    if (e == cudaSuccess && rc == GM_OK && h_is_dup) e = cudaMemcpyAsync(h_is_dup, d_dup, (size_t)n, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess && rc == GM_OK && h_first_row) e = cudaMemcpyAsync(h_first_row, d_first, (size_t)n * 8, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    dev_free(d_keys, st); dev_free(d_dup, st); dev_free(d_first, st);
    if (rc) return rc;
    if (e != cudaSuccess) return cuda_fail(e, "dedup", __FILE__, __LINE__);
    return GM_OK;
}

}  // namespace gm

extern "C" int gm_seed_dedup(const uint64_t *guide2bit, int64_t n, int L, int lsr, int five_prime, uint8_t *is_dup) {
    return gm::dedup_run(guide2bit, n, L, lsr, five_prime, is_dup, nullptr);
}

extern "C" int gm_first_occurrence(const uint64_t *keys, int64_t n, int64_t *first_row) {
    // lsr = 0 -> the whole 64-bit word is the key
    return gm::dedup_run(keys, n, GM_MAX_L, 0, 0, nullptr, first_row);
}

/* device-resident variants: pointers into HBM, enqueued on `stream` (NULL = default stream), no synchronisation */
extern "C" int gm_seed_dedup_dev(const uint64_t *d_guide2bit, int64_t n, int L, int lsr, int five_prime, uint8_t *d_is_dup, void *stream) {
    int rc = gm::ensure_init();
    if (rc) return rc;
    return gm::dedup_dev(d_guide2bit, n, L, lsr, five_prime, d_is_dup, nullptr, nullptr, (cudaStream_t)stream);
}

extern "C" int gm_first_occurrence_dev(const uint64_t *d_keys, int64_t n, int32_t *d_first_row, void *stream) {
    int rc = gm::ensure_init();
    if (rc) return rc;
    return gm::dedup_dev(d_keys, n, GM_MAX_L, 0, 0, nullptr, nullptr, d_first_row, (cudaStream_t)stream);
}
