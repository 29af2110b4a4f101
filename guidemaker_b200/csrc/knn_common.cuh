// knn_common.cuh -- declarations shared by the pair-scan kernels (knn.cu: K3a/K4, knn_tc.cu: K3b).
#pragma once
#include "common.cuh"
#include "distance.cuh"

namespace gm {

int prof_begin(cudaStream_t s);
void prof_end(int slot, cudaStream_t s, double pairs);

static constexpr int CHUNK = 1024;      // targets per shared-memory stage (8 KB)
static constexpr int NSTAGE = 3;
static constexpr int THREADS = 128;
static constexpr int MAX_SPLITS = 64;
static constexpr uint32_t KEY_EMPTY = 0xFFFFFFFFu;
static constexpr int IDX_BITS = 27;

struct Index {
    uint2 *planes = nullptr;
    uint2 *planes_perm = nullptr;   // Hamming only: bit-permuted copy for K3b (knn_tc.cu, "Bit order")
    int64_t n_u = 0, n_pad = 0;
    int L = 0, metric = 0;
    void *ws = nullptr;
    size_t ws_bytes = 0;
    // per-handle tuning (gm_index_tune); a negative value follows the process-wide default (gm_knn_tune / gm_knn_engine)
    int engine = -1, tune_r = -1, tune_splits = -1, tune_warm = -2;
    unsigned long long *dbg = nullptr;      // GM_TC_DEBUG counters of this index
    // warm.cu: copies of the table sorted by guide with its positions rotated by c * L / copies, built lazily
    uint2 *sorted_p[4] = {nullptr, nullptr, nullptr, nullptr};
    uint32_t *sorted_i[4] = {nullptr, nullptr, nullptr, nullptr};
    // warm.cu: Levenshtein only -- the table sorted by the guides read from their first base (K4p), as 2-bit codes
    // (lo, hi word), padded like `planes`
    uint2 *prefix_p = nullptr;
    uint32_t *prefix_i = nullptr;
};


// ---- list maintenance ------------------------------------------------------------------------------

// Insert key into the thread-private ascending list if it beats the current worst entry.
// Returns the distance of the (new) worst entry, 31 while the list is not full.
static __device__ __noinline__ uint32_t list_insert(uint32_t *__restrict__ lst, int k, uint32_t key) {
    uint32_t worst = lst[k - 1];
    if (key < worst) {
        int pos = k - 1;
        while (pos > 0) {
            uint32_t v = lst[pos - 1];
            if (v <= key) break;
            lst[pos] = v;
            pos--;
        }
        lst[pos] = key;
        worst = lst[k - 1];
    }
    return worst >> IDX_BITS;
}

// bias constant of the packed threshold test: byte = 128 + (tau - 1); after subtracting a distance
// p <= 27 the byte keeps bit 7 iff p <= tau - 1, i.e. p < tau.  Bytes stay within [100, 158]: no
// borrow ever crosses a byte boundary.
__device__ __forceinline__ uint32_t bias_of(uint32_t tau) { return 0x7F7F7F7Fu + tau * 0x01010101u; }

struct ScanArgs {
    const uint2 *tplanes;
    const uint2 *tperm;       // K3b: bit-permuted planes of the same table
    const uint32_t *tidx;     // K4p: original index of every row of the prefix-sorted table `tplanes` points to
    int prefix_c0;            // K4p: first prefix length whose state is kept in shared memory
    int first_chunk;          // K3b: the scan starts here (behind the warm sample, whose lists split 0 inherits)
    int tile_offset;          // K3b: blockIdx.x + tile_offset = query tile (the tail launch covers the last tiles)
    int n_chunks;             // chunks to cover (ceil(n_scan / CHUNK))
    int chunks_per_split;
    int64_t n_u;              // targets beyond this index are padding
    const uint2 *qplanes;
    int64_t q, q_pad;
    int k;
    uint32_t *lists;          // [gridDim.y][list_stride][k]; row of query qi = qi - list_q0
    int64_t list_stride;      // queries per split in `lists` (q_pad, or the tail's query count for K3b's tail launch)
    int64_t list_q0;          // first query stored in `lists`
    const uint32_t *warm;     // [q_pad][k] lists of the warm-up launch or nullptr
    int warm_any_subset;      // K3b: `warm` comes from an arbitrary subset of the table (warm.cu), not from its first chunks:
                              // only the k-th distance is used, as an INCLUSIVE bound, and no split inherits the lists
    int L;
    unsigned long long *dbg;  // optional per-role cycle counters of block (0,0) (GM_TC_DEBUG=1), else nullptr
};

__device__ __forceinline__ void issue_chunk(uint2 *dst, const uint2 *src, uint64_t *bar) {
    mbar_expect_tx(bar, CHUNK * (uint32_t)sizeof(uint2));
    bulk_g2s(dst, src, CHUNK * (uint32_t)sizeof(uint2), bar);
}


// K3b launcher (knn_tc.cu): same contract as the K3a scan launch
int launch_hamming_tc(dim3 grid, cudaStream_t st, const ScanArgs &a);
int tc_permute_planes(const uint2 *planes, int64_t n, uint2 *out, cudaStream_t st);
int tc_query_tile();
int tc_k_chunks(int L);
int microbench_mma_i8(int variant, double *ops_per_s);
// neighbourhood warm start (warm.cu)
int warm_window(Index *ix, const uint2 *qplanes, int64_t q, int k, int W, uint32_t *wlists, cudaStream_t st);
void warm_free_index(Index *ix);
int warm_copies();
int sorted_by_prefix(Index *ix, cudaStream_t st);

}  // namespace gm
