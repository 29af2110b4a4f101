// knn.cu -- exact brute-force k-nearest-neighbour search over guides (K3a Hamming, K4 Levenshtein,
// K5 min-distance) for sm_100a.  Replaces nmslib's HNSW index (core.py:418-523, :603-606).
//
// Layout in HBM
//   index  : uint2 planes[n_pad]    (lo, hi) bit planes of every distinct guide, zero padded to a
//            multiple of CHUNK so every bulk copy is full-sized.
//   queries: uint2 qplanes[q_pad]   same layout, padded to a multiple of the query tile.
//   lists  : uint32 keys[split][q_pad][k]  per (split, query) ascending list of
//            key = (distance << 27) | target index, 0xFFFFFFFF = empty.
//
// Pair-scan kernel (the dominant kernel, one launch per gm_knn call plus an optional warm-up launch)
//   grid = (query tiles, target splits).  A CTA owns THREADS*R queries -- R per thread, held in
//   registers as planes -- and streams its split of the target table through shared memory in
//   CHUNK-sized stages filled by the TMA engine (cp.async.bulk, completion on an mbarrier).  Every
//   lane reads the same target (shared-memory broadcast) and evaluates it against its R queries:
//   2 LOP3 + 1 POPC per pair.  Four distances are packed into the bytes of one word with
//   multiply-adds on the FMA pipe and tested against the query's current k-th-best distance with a
//   single biased subtraction: bit 7 of byte j is set iff distance j beats the threshold.  Only then
//   (rare: the threshold tightens after a few hundred targets) does the thread fall into the
//   insertion path, which keeps its private sorted list in global memory.  Targets stream in
//   ascending index and an equal distance never displaces an earlier entry, which yields the
//   deterministic (distance, index) order.
//
//   Warm start: an optional first launch scans only the first `warm` targets; its k-th-best
//   distance per query is a valid upper bound and seeds the thresholds of the full scan, so the
//   flood of insertions at the start of every split disappears.
//
// Merge kernel: k smallest keys over the splits of a query -> (int32 idx, uint8 dist) rows.
#include "knn_common.cuh"
#include <mutex>
#include <new>
#include <stdlib.h>

namespace gm {

static int g_tune_r = 8;
static int g_tune_splits = 0;
static int g_tune_warm = -1;
static int g_tune_engine = 1;          // 1 = K3b tcgen05 one-hot GEMM (tensor pipe, default), 0 = K3a XOR/POPC (INT pipes)

// ---- small kernels -------------------------------------------------------------------------------

__global__ void to_planes_kernel(const uint64_t *__restrict__ g, int64_t n, int64_t n_pad, uint2 *__restrict__ out) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_pad) return;
    out[i] = i < n ? to_planes(g[i]) : make_uint2(0u, 0u);
}

// ---- K3a: Hamming pair scan --------------------------------------------------------------------------
template <int R>
__global__ void __launch_bounds__(THREADS) knn_hamming_scan_kernel(const ScanArgs a) {
    __shared__ __align__(128) uint2 s_t[NSTAGE][CHUNK];
    __shared__ __align__(8) uint64_t s_full[NSTAGE];

    const int tid = threadIdx.x;
    const int c0 = blockIdx.y * a.chunks_per_split;
    const int c1 = min(c0 + a.chunks_per_split, a.n_chunks);
    if (c0 >= c1) return;

    const int64_t qbase = (int64_t)blockIdx.x * (THREADS * R) + tid;
    uint32_t qlo[R], qhi[R], tau[R], C[R];
#pragma unroll
    for (int r = 0; r < R; r++) {
        const int64_t qi = qbase + (int64_t)r * THREADS;      // < q_pad by construction
        const uint2 p = a.qplanes[qi];
        qlo[r] = p.x;
        qhi[r] = p.y;
        uint32_t t = 31u;
        if (a.warm) t = min((a.warm[(size_t)qi * a.k + (a.k - 1)] >> IDX_BITS) + 1u, 31u);
        tau[r] = qi < a.q ? t : 0u;                           // padding queries never insert
        C[r] = bias_of(tau[r]);
    }
    uint32_t *const my_lists = a.lists + ((size_t)blockIdx.y * a.list_stride + (qbase - a.list_q0)) * a.k;

    if (tid == 0) {
        for (int s = 0; s < NSTAGE; s++) mbar_init(&s_full[s], 1);
        mbar_fence_init();
    }
    __syncthreads();
    if (tid == 0) {
        for (int s = 0; s < NSTAGE && c0 + s < c1; s++)
            issue_chunk(s_t[s], a.tplanes + (size_t)(c0 + s) * CHUNK, &s_full[s]);
    }

    int stage = 0;
    uint32_t parity = 0;
    for (int c = c0; c < c1; c++) {
        mbar_wait(&s_full[stage], parity);
        const uint4 *s4 = reinterpret_cast<const uint4 *>(s_t[stage]);
        const uint32_t tbase = (uint32_t)c * CHUNK;

#pragma unroll 2
        for (int g = 0; g < CHUNK / 4; g++) {
            const uint4 t01 = s4[2 * g];          // targets 4g, 4g+1: (lo, hi, lo, hi) -- broadcast reads
            const uint4 t23 = s4[2 * g + 1];      // targets 4g+2, 4g+3
            uint32_t x[R];
            uint32_t any = 0;
#pragma unroll
            for (int r = 0; r < R; r++) {
                const uint32_t p0 = __popc((qlo[r] ^ t01.x) | (qhi[r] ^ t01.y));
                const uint32_t p1 = __popc((qlo[r] ^ t01.z) | (qhi[r] ^ t01.w));
                const uint32_t p2 = __popc((qlo[r] ^ t23.x) | (qhi[r] ^ t23.y));
                const uint32_t p3 = __popc((qlo[r] ^ t23.z) | (qhi[r] ^ t23.w));
                // pack on the FMA pipe: x = C - p0 - p1<<8 - p2<<16 - p3<<24
                uint32_t v = C[r] - p0;
                v = p1 * 0xFFFFFF00u + v;
                v = p2 * 0xFFFF0000u + v;
                v = p3 * 0xFF000000u + v;
                x[r] = v;
                any |= v;
            }
            if (any & 0x80808080u) {
#pragma unroll
                for (int r = 0; r < R; r++) {
                    const uint32_t h = x[r] & 0x80808080u;
                    if (h) {
#pragma unroll
                        for (int j = 0; j < 4; j++) {
                            if (h & (0x80u << (8 * j))) {
                                const uint32_t d = ((C[r] >> (8 * j)) & 0xFFu) - ((x[r] >> (8 * j)) & 0xFFu);
                                const uint32_t t = tbase + 4u * g + j;
                                if (t < a.n_u) {
                                    const uint32_t w = list_insert(my_lists + (size_t)r * THREADS * a.k, a.k, (d << IDX_BITS) | t);
                                    tau[r] = min(tau[r], w);
                                }
                            }
                        }
                        C[r] = bias_of(tau[r]);
                    }
                }
            }
        }

        __syncthreads();                                   // every lane is done with this stage
        if (tid == 0 && c + NSTAGE < c1)
            issue_chunk(s_t[stage], a.tplanes + (size_t)(c + NSTAGE) * CHUNK, &s_full[stage]);
        if (++stage == NSTAGE) { stage = 0; parity ^= 1u; }
    }
}

// ---- K4: Levenshtein pair scan (Myers bit-parallel, one 32-bit word per pair) -------------------------
// Myers' recurrence in Hyyro's global-alignment form (distance.cuh: myers_planes) with the roles arranged for the SIMT
// machine: the query is the pattern (its L rows live in one 32-bit word per thread), the target is the text and is
// WARP-UNIFORM.  The match vector of a step, Eq = "pattern rows equal to text base j", is then one of four per-query
// constants PM[A|C|G|T] chosen by a value every lane agrees on -- so the choice is a uniform branch, not arithmetic:
// the step body exists four times (once per text base) and costs 7 LOP3 + 3 IMAD per pair, against 10 LOP3 + 3 IMAD (+ the
// broadcast of the text bit planes) when Eq is recomputed from the planes.  The ALU pipe (LOP3) is what bounds the kernel.
#define GM_MYERS_STEP(EQ)                                                                                    \
    _Pragma("unroll") for (int r = 0; r < R; r++) {                                                          \
        const uint32_t Eq = (EQ)[r];                                                                         \
        const uint32_t Xv = Eq | Mv[r];                                                                      \
        const uint32_t Xh = (((Eq & Pv[r]) + Pv[r]) ^ Pv[r]) | Eq;                                           \
        const uint32_t Ph = (Mv[r] | ~(Xh | Pv[r])) * two + 1u;      /* (Ph << 1) | 1 as ONE multiply-add */  \
        const uint32_t Mh = (Pv[r] & Xh) * two;                                                              \
        Pv[r] = Mh | ~(Xv | Ph);                                                                             \
        Mv[r] = Ph & Xv;                                                                                     \
    }

template <int R>
__global__ void __launch_bounds__(THREADS) knn_leven_scan_kernel(const ScanArgs a) {
    __shared__ __align__(128) uint2 s_t[NSTAGE][CHUNK];
    __shared__ __align__(8) uint64_t s_full[NSTAGE];

    const int tid = threadIdx.x;
    const int c0 = blockIdx.y * a.chunks_per_split;
    const int c1 = min(c0 + a.chunks_per_split, a.n_chunks);
    if (c0 >= c1) return;

    const int L = a.L;
    const uint32_t lmask = (1u << L) - 1u;
    // The shifts of the recurrence run as multiply-adds on the FMA pipe (x * 2 + 1, x * 2), which has spare issue slots;
    // `two` is opaque to the compiler (L <= 27), otherwise it turns them back into an add plus a LOP3 on the ALU pipe.
    const uint32_t two = 2u + ((uint32_t)L >> 30);
    const int64_t qbase = (int64_t)blockIdx.x * (THREADS * R) + tid;
    uint32_t pmA[R], pmC[R], pmG[R], pmT[R], tau[R];       // bits >= L hold garbage that never flows downwards
#pragma unroll
    for (int r = 0; r < R; r++) {
        const int64_t qi = qbase + (int64_t)r * THREADS;
        const uint2 p = a.qplanes[qi];
        pmA[r] = ~(p.x | p.y);
        pmC[r] = p.x & ~p.y;
        pmG[r] = p.y & ~p.x;
        pmT[r] = p.x & p.y;
        uint32_t t = 31u;
        if (a.warm) t = min((a.warm[(size_t)qi * a.k + (a.k - 1)] >> IDX_BITS) + 1u, 31u);
        tau[r] = qi < a.q ? t : 0u;
    }
    uint32_t *const my_lists = a.lists + ((size_t)blockIdx.y * a.list_stride + (qbase - a.list_q0)) * a.k;

    if (tid == 0) {
        for (int s = 0; s < NSTAGE; s++) mbar_init(&s_full[s], 1);
        mbar_fence_init();
    }
    __syncthreads();
    if (tid == 0) {
        for (int s = 0; s < NSTAGE && c0 + s < c1; s++)
            issue_chunk(s_t[s], a.tplanes + (size_t)(c0 + s) * CHUNK, &s_full[s]);
    }

    int stage = 0;
    uint32_t parity = 0;
    for (int c = c0; c < c1; c++) {
        mbar_wait(&s_full[stage], parity);
        const uint32_t tbase = (uint32_t)c * CHUNK;
        const int n_here = (int)min((int64_t)CHUNK, a.n_u - (int64_t)tbase);   // skip padding targets

        for (int g = 0; g < n_here; g++) {
            const uint2 t = s_t[stage][g];                  // warp-uniform target
            // text bases, 2 bits each, in step order: base j = (code >> 2j) & 3 (A=0 C=1 G=2 T=3)
            uint64_t code = spread_bits(t.x) | (spread_bits(t.y) << 1);
            uint32_t Pv[R], Mv[R];
#pragma unroll
            for (int r = 0; r < R; r++) { Pv[r] = 0xFFFFFFFFu; Mv[r] = 0u; }
#pragma unroll 1
            for (int j = 0; j < L; j++) {
                const uint32_t base = (uint32_t)code & 3u;   // the same in every lane: the switch is a uniform branch
                code >>= 2;
                if (base == 0u) { GM_MYERS_STEP(pmA) }
                else if (base == 1u) { GM_MYERS_STEP(pmC) }
                else if (base == 2u) { GM_MYERS_STEP(pmG) }
                else { GM_MYERS_STEP(pmT) }
            }
#pragma unroll
            for (int r = 0; r < R; r++) {
                const uint32_t d = (uint32_t)(L + __popc(Pv[r] & lmask) - __popc(Mv[r] & lmask));
                if (d < tau[r]) {
                    const uint32_t w = list_insert(my_lists + (size_t)r * THREADS * a.k, a.k, (d << IDX_BITS) | (tbase + g));
                    tau[r] = min(tau[r], w);
                }
            }
        }

        __syncthreads();
        if (tid == 0 && c + NSTAGE < c1)
            issue_chunk(s_t[stage], a.tplanes + (size_t)(c + NSTAGE) * CHUNK, &s_full[stage]);
        if (++stage == NSTAGE) { stage = 0; parity ^= 1u; }
    }
}

// ---- K4p: Levenshtein pair scan over the PREFIX-SORTED table ---------------------------------------------------------
// Myers' recurrence consumes the text (the target) base by base and its state after j bases -- (Pv, Mv) per query --
// depends only on the text's first j bases.  The index keeps a copy of the table sorted by the guides read from their
// first base (warm.cu: sorted_by_prefix); consecutive guides of that copy share their first ~log4(n) bases (10 of 20 for
// 1.4 M guides), and the target is warp- and CTA-uniform, so the states of the previous target after c0 .. c0+NL-1 bases
// are kept in shared memory ([level][2R][thread], 8 KB per level at R = 8) and a target resumes from the deepest kept
// level its common prefix with its predecessor reaches: ~L - log4(n) steps per pair instead of L.  Targets no longer
// arrive in index order, so the lists work on full (distance, original index) keys: a tie at the k-th distance gets in
// iff its index is lower.  Same bits as the plain kernel.
#ifndef GM_PFX_LEVELS
#define GM_PFX_LEVELS 2
#endif
static constexpr int PFX_LEVELS = GM_PFX_LEVELS;
static constexpr int PFX_STAGES = 2;
static size_t pfx_smem_bytes(int R) { return (size_t)PFX_STAGES * CHUNK * 12 + (size_t)PFX_LEVELS * 2 * R * THREADS * 4; }

// insert by full key; returns the key of the (new) worst entry, KEY_EMPTY while the list is not full
static __device__ __noinline__ uint32_t list_insert_key(uint32_t *__restrict__ lst, int k, uint32_t key) {
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
    return worst;
}

template <int R>
__global__ void __launch_bounds__(THREADS) knn_leven_prefix_kernel(const ScanArgs a) {
    extern __shared__ __align__(128) uint8_t pfx_smem[];
    uint2 *s_t = reinterpret_cast<uint2 *>(pfx_smem);                               // [PFX_STAGES][CHUNK] planes
    uint32_t *s_i = reinterpret_cast<uint32_t *>(s_t + PFX_STAGES * CHUNK);         // [PFX_STAGES][CHUNK] original indices
    uint32_t *s_stack = s_i + PFX_STAGES * CHUNK;                                   // [PFX_LEVELS][2R][THREADS]
    __shared__ __align__(8) uint64_t s_full[PFX_STAGES];

    const int tid = threadIdx.x;
    const int c0 = blockIdx.y * a.chunks_per_split;
    const int c1 = min(c0 + a.chunks_per_split, a.n_chunks);
    if (c0 >= c1) return;

    const int L = a.L;
    const uint32_t lmask = (1u << L) - 1u;
    const uint32_t two = 2u + ((uint32_t)L >> 30);          // opaque 2: keeps the shifts on the FMA pipe (see K4)
    const int lv0 = a.prefix_c0, lv_top = a.prefix_c0 + PFX_LEVELS - 1;             // kept levels: lv0 .. lv_top (< L)
    const int64_t qbase = (int64_t)blockIdx.x * (THREADS * R) + tid;
    uint32_t pmA[R], pmC[R], pmG[R], pmT[R], bound[R];
#pragma unroll
    for (int r = 0; r < R; r++) {
        const int64_t qi = qbase + (int64_t)r * THREADS;
        const uint2 p = a.qplanes[qi];
        pmA[r] = ~(p.x | p.y);
        pmC[r] = p.x & ~p.y;
        pmG[r] = p.y & ~p.x;
        pmT[r] = p.x & p.y;
        uint32_t b = KEY_EMPTY;                             // insert iff key < bound
        if (a.warm) {
            const uint32_t w = a.warm[(size_t)qi * a.k + (a.k - 1)];
            if (w != KEY_EMPTY) b = ((w >> IDX_BITS) + 1u) << IDX_BITS;            // everything up to the warm k-th distance
        }
        bound[r] = qi < a.q ? b : 0u;
    }
    uint32_t *const my_lists = a.lists + ((size_t)blockIdx.y * a.list_stride + (qbase - a.list_q0)) * a.k;

    if (tid == 0) {
        for (int s = 0; s < PFX_STAGES; s++) mbar_init(&s_full[s], 1);
        mbar_fence_init();
    }
    __syncthreads();
    auto issue = [&](int stage, int c) {
        mbar_expect_tx(&s_full[stage], CHUNK * 12u);
        bulk_g2s(s_t + stage * CHUNK, a.tplanes + (size_t)c * CHUNK, CHUNK * 8u, &s_full[stage]);
        bulk_g2s(s_i + stage * CHUNK, a.tidx + (size_t)c * CHUNK, CHUNK * 4u, &s_full[stage]);
    };
    if (tid == 0) {
        for (int s = 0; s < PFX_STAGES && c0 + s < c1; s++) issue(s, c0 + s);
    }

    int stage = 0;
    uint32_t parity = 0;
    uint64_t prev = 0;
    bool have_prev = false;
    for (int c = c0; c < c1; c++) {
        mbar_wait(&s_full[stage], parity);
        const uint32_t tbase = (uint32_t)c * CHUNK;
        const int n_here = (int)min((int64_t)CHUNK, a.n_u - (int64_t)tbase);       // skip padding targets

        for (int g = 0; g < n_here; g++) {
            // CTA-uniform target, stored as its 2-bit code (base j = bits 2j, 2j+1): no per-target re-interleaving of planes
            // (-6 %).  (Measured and dropped: fetching the next target and its prefix length one iteration ahead, +10 %;
            // the prefix length precomputed into the spare high bits of the index word, +4 % -- in both the values leave
            // the uniform datapath.)
            const uint2 t = s_t[stage * CHUNK + g];
            const uint32_t oidx = s_i[stage * CHUNK + g];
            const uint64_t tcode = ((uint64_t)t.y << 32) | t.x;
            // common prefix with the previous target, in bases
            int pl = 0;
            if (have_prev) pl = min((__ffsll((long long)((tcode ^ prev) | (1ull << (2 * L)))) - 1) >> 1, lv_top);
            if (pl < lv0) pl = 0;
            prev = tcode;
            have_prev = true;
            uint32_t Pv[R], Mv[R];
            if (pl == 0) {
#pragma unroll
                for (int r = 0; r < R; r++) { Pv[r] = 0xFFFFFFFFu; Mv[r] = 0u; }
            } else {
                const uint32_t *s = s_stack + (size_t)(pl - lv0) * (2 * R * THREADS) + tid;
#pragma unroll
                for (int r = 0; r < R; r++) { Pv[r] = s[(2 * r) * THREADS]; Mv[r] = s[(2 * r + 1) * THREADS]; }
            }
            uint64_t code = tcode >> (2 * pl);
#pragma unroll 1
            for (int j = pl; j < L; j++) {
                const uint32_t base = (uint32_t)code & 3u;   // the same in every lane: the switch is a uniform branch
                code >>= 2;
                if (base == 0u) { GM_MYERS_STEP(pmA) }
                else if (base == 1u) { GM_MYERS_STEP(pmC) }
                else if (base == 2u) { GM_MYERS_STEP(pmG) }
                else { GM_MYERS_STEP(pmT) }
                if (j + 1 >= lv0 && j + 1 <= lv_top) {      // state after j + 1 bases: keep it for the next targets
                    uint32_t *s = s_stack + (size_t)(j + 1 - lv0) * (2 * R * THREADS) + tid;
#pragma unroll
                    for (int r = 0; r < R; r++) { s[(2 * r) * THREADS] = Pv[r]; s[(2 * r + 1) * THREADS] = Mv[r]; }
                }
            }
#pragma unroll
            for (int r = 0; r < R; r++) {
                const uint32_t d = (uint32_t)(L + __popc(Pv[r] & lmask) - __popc(Mv[r] & lmask));
                const uint32_t key = (d << IDX_BITS) | oidx;
                if (key < bound[r]) {
                    const uint32_t w = list_insert_key(my_lists + (size_t)r * THREADS * a.k, a.k, key);
                    bound[r] = min(bound[r], w);
                }
            }
        }

        __syncthreads();
        if (tid == 0 && c + PFX_STAGES < c1) issue(stage, c + PFX_STAGES);
        if (++stage == PFX_STAGES) { stage = 0; parity ^= 1u; }
    }
}
#undef GM_MYERS_STEP

// ---- merge: k smallest keys over the splits of each query ------------------------------------------------
// Queries below `tail_q0` have `splits` lists in `lists` ([split][q_pad][k]); the others (K3b's split tail wave) have
// `tail_splits` lists in `tail_lists` ([split][tail_stride][k], row = query - tail_q0).
__global__ void knn_merge_kernel(const uint32_t *__restrict__ lists, int splits, int64_t q, int64_t q_pad, int k,
                                 const uint32_t *__restrict__ tail_lists, int tail_splits, int64_t tail_q0, int64_t tail_stride,
                                 int32_t *__restrict__ out_idx, uint8_t *__restrict__ out_dist, int dist_only) {
    const int64_t qi = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (qi >= q) return;
    const bool tail = qi >= tail_q0;
    const uint32_t *base = tail ? tail_lists + (size_t)(qi - tail_q0) * k : lists + (size_t)qi * k;
    const size_t stride = (size_t)(tail ? tail_stride : q_pad) * k;
    const int ns = tail ? tail_splits : splits;
    uint8_t ptr[MAX_SPLITS];
    for (int s = 0; s < ns; s++) ptr[s] = 0;
    for (int j = 0; j < k; j++) {
        uint32_t best = KEY_EMPTY;
        int bs = -1;
        for (int s = 0; s < ns; s++) {
            if (ptr[s] < k) {
                const uint32_t v = base[(size_t)s * stride + ptr[s]];
                if (v < best) { best = v; bs = s; }
            }
        }
        if (bs >= 0) ptr[bs]++;
        const uint8_t d = best == KEY_EMPTY ? (uint8_t)255 : (uint8_t)(best >> IDX_BITS);
        if (dist_only) {
            if (j == 0) out_dist[qi] = d;
        } else {
            out_idx[qi * k + j] = best == KEY_EMPTY ? -1 : (int32_t)(best & ((1u << IDX_BITS) - 1u));
            out_dist[qi * k + j] = d;
        }
    }
}

// ---- host side ------------------------------------------------------------------------------------------

// One retired workspace is kept for the next index (callers that rebuild the index for every batch -- bench e2e, the
// control rounds -- would otherwise return ~0.4 GB to the pool and take it back on every call).  gm_index_free parks
// it here after a device synchronise, so whoever adopts it next needs no stream ordering.
static std::mutex g_spare_mu;
static void *g_spare_ws = nullptr;
static size_t g_spare_bytes = 0;

static void park_ws(void *ws, size_t bytes) {
    if (!ws) return;
    std::lock_guard<std::mutex> lk(g_spare_mu);
    if (bytes > g_spare_bytes) { dev_free(g_spare_ws, 0); g_spare_ws = ws; g_spare_bytes = bytes; }
    else dev_free(ws, 0);
}

static int ensure_ws(Index *ix, size_t bytes, cudaStream_t st) {
    if (bytes <= ix->ws_bytes) return GM_OK;
    const double t0 = now_ms();
    dev_free(ix->ws, st);                      // stream ordered: earlier kernels on `st` finish first
    ix->ws = nullptr;
    ix->ws_bytes = 0;
    {
        std::lock_guard<std::mutex> lk(g_spare_mu);
        if (g_spare_ws && g_spare_bytes >= bytes) {
            ix->ws = g_spare_ws; ix->ws_bytes = g_spare_bytes;
            g_spare_ws = nullptr; g_spare_bytes = 0;
            return GM_OK;
        }
    }
    bytes = (bytes + (bytes >> 3) + 4095) & ~(size_t)4095;
    GM_CUDA(dev_alloc(&ix->ws, bytes, st));
    ix->ws_bytes = bytes;
    trace("  knn: workspace alloc", t0);
    return GM_OK;
}

template <int R>
static void launch_scan(int metric, dim3 grid, cudaStream_t st, const ScanArgs &a) {
    if (metric == GM_METRIC_HAMMING) knn_hamming_scan_kernel<R><<<grid, THREADS, 0, st>>>(a);
    else knn_leven_scan_kernel<R><<<grid, THREADS, 0, st>>>(a);
    count_launch();
}

template <int R>
static int launch_leven_prefix(dim3 grid, cudaStream_t st, const ScanArgs &a) {
    static bool attr_set = false;
    if (!attr_set) {
        GM_CUDA(cudaFuncSetAttribute(knn_leven_prefix_kernel<R>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pfx_smem_bytes(R)));
        attr_set = true;
    }
    knn_leven_prefix_kernel<R><<<grid, THREADS, pfx_smem_bytes(R), st>>>(a);
    count_launch();
    return GM_OK;
}

static int knn_run(Index *ix, const uint64_t *d_q, int64_t q, int k, int32_t *d_idx, uint8_t *d_dist, int dist_only,
                   cudaStream_t st) {
    GM_ARG(ix && ix->planes, "gm_knn: invalid index handle");
    GM_ARG(k >= 1 && k <= GM_MAX_K, "gm_knn: k=%d outside [1,%d]", k, GM_MAX_K);
    GM_ARG(q >= 0, "gm_knn: negative query count");
    if (q == 0) return GM_OK;
    GM_ARG(d_q && d_dist && (dist_only || d_idx), "gm_knn: NULL buffer");

    // queries per thread: Levenshtein keeps 2 more state words per pair, so it uses R=4
    const int tune_r = ix->tune_r > 0 ? ix->tune_r : g_tune_r;
    const int tune_splits = ix->tune_splits >= 0 ? ix->tune_splits : g_tune_splits;
    const int tune_warm = ix->tune_warm >= -1 ? ix->tune_warm : g_tune_warm;
    const int R = tune_r == 4 ? 4 : 8;
    const bool use_tc = (ix->engine >= 0 ? ix->engine : g_tune_engine) == 1 && ix->metric == GM_METRIC_HAMMING;
    // Levenshtein: the prefix-sharing scan over the sorted copy (K4p) unless the handle asks for the plain kernel
    // (engine 0) or the table is too small / the guides too short for shared prefixes to exist
    const bool use_prefix = ix->metric == GM_METRIC_LEVEN && (ix->engine >= 0 ? ix->engine : g_tune_engine) == 1 &&
                            ix->n_u >= 4096 && ix->L >= 2 * PFX_LEVELS;
    const int QT = THREADS * R;
    const int64_t tiles = (q + QT - 1) / QT;
    const int64_t q_pad = tiles * QT;
    const int n_chunks = (int)(ix->n_pad / CHUNK);

    // target splits: enough CTAs for >= ~16 per SM so the last wave is a small fraction
    int splits = tune_splits;
    // K3b runs one CTA per SM, so a launch of T query tiles takes ceil(T / SMs) waves and the last one may be nearly
    // empty (T = 335 on each of 8 GPUs for the 6.3 Mb config: 2.26 waves of work in 3).  Two-tier launch: the largest
    // multiple of the SM count runs unsplit, the remaining `tail_tiles` are cut into `tail_splits` target ranges so they
    // fill ONE short wave (a split restarts with loose bounds, which costs ~16 % of an unsplit CTA, measured).
    int tail_tiles = 0, tail_splits = 1;
    if (use_tc && splits <= 0) {
        const int sms = device_sm_count();
        const int64_t grid_x = q_pad / tc_query_tile();
        const int rem = (int)(grid_x % sms);
        if (grid_x > sms && rem > 0 && sms / rem >= 2) {
            tail_tiles = rem;
            tail_splits = sms / rem > 8 ? 8 : sms / rem;
        }
    }
    if (tail_tiles) {
        splits = 1;
    } else if (splits <= 0) {
        // K3a: >= 16 CTAs per SM.  K3b: one resident CTA per SM and every split restarts with the loose warm-start
        // thresholds (measured ~0.6 ms of SM time per extra CTA), so it only splits to reach ~2 CTAs per SM
        const int64_t grid_x = use_tc ? q_pad / tc_query_tile() : tiles;
        const int64_t want = (int64_t)device_sm_count() * (use_tc ? 2 : 16);
        splits = (int)((want + grid_x - 1) / grid_x);
        // (choosing the split count so that the CTAs fill whole waves was measured for K4 / K4p: no gain, 594 vs 582 ms)
    }
    // warm start: worthwhile only when the table is much larger than the sample
    // measured: 8192 is best for K3b, 4096 for K3a; K4 (Levenshtein) runs 1 % faster WITHOUT a warm-up launch (its warm
    // launch has only one CTA per 1024 queries and the insertions it saves are cheap against ~200 ALU ops per pair)
    int warm = tune_warm < 0 ? (ix->metric == GM_METRIC_LEVEN ? 0 : (use_tc ? 8 : 4) * CHUNK) : tune_warm;
    warm = (warm + CHUNK - 1) / CHUNK;                       // in chunks
    if (n_chunks < 16 * warm || ix->n_u < (int64_t)warm * CHUNK) warm = 0;
    // K3b's default is the neighbourhood warm start instead (warm.cu): a bound from the WINDOW guides around the query's
    // rank in each of three sorted copies of the table; an explicit warm_sample > 0 keeps the first-chunks sample.
    // Measured on the 6.3 Mb table (tools/warm_sweep.py, ms per pass; first 8192 guides: 62.9, no warm start: 113.4):
    //   window   256    512    1024   2048
    //   1 copy   57.5   56.7   56.0   55.8
    //   2        54.4   54.0   53.7   54.6
    //   3        52.9   53.0   53.9   55.6
    //   4        52.7   53.4   54.5   56.9
    // GM_WARM_WINDOW / GM_WARM_COPIES override the defaults for such sweeps.
    const char *w_env = getenv("GM_WARM_WINDOW");
    const int WINDOW = w_env && atoi(w_env) > 0 ? atoi(w_env) : 256;
    const bool window_warm = use_tc && tune_warm < 0 && ix->n_u >= 64 * WINDOW && q < (1LL << 31);
    if (window_warm) warm = 0;
    // K3b inherits the warm lists and scans only the chunks behind the sample; K3a rescans from chunk 0
    const int first_chunk = use_tc ? warm : 0;
    const int scan_chunks = n_chunks - first_chunk;

    if (splits > scan_chunks) splits = scan_chunks;
    if (splits > MAX_SPLITS) splits = MAX_SPLITS;
    if (splits < 1) splits = 1;
    int cps = (scan_chunks + splits - 1) / splits;
    splits = (scan_chunks + cps - 1) / cps;
    int tail_cps = 0;
    if (tail_tiles) {
        if (tail_splits > scan_chunks) tail_splits = scan_chunks;
        tail_cps = (scan_chunks + tail_splits - 1) / tail_splits;
        tail_splits = (scan_chunks + tail_cps - 1) / tail_cps;
        if (tail_splits < 2) { tail_tiles = 0; tail_splits = 1; }
    }
    // lists: [splits][q_pad][k] for the main launch; the split tail wave of K3b keeps its own compact
    // [tail_splits][tail_q][k] block behind it (only the tail's queries have more than one list)
    const int64_t tail_q = (int64_t)tail_tiles * (use_tc ? tc_query_tile() : 0);
    const int64_t tail_q0 = tail_tiles ? q_pad - tail_q : q_pad;
    const size_t qp_bytes = (size_t)q_pad * sizeof(uint2);
    const size_t main_bytes = (size_t)splits * q_pad * k * sizeof(uint32_t);
    const size_t tail_bytes = (size_t)(tail_tiles ? tail_splits : 0) * tail_q * k * sizeof(uint32_t);
    const size_t list_bytes = main_bytes + tail_bytes;
    const size_t warm_bytes = (warm || window_warm) ? (size_t)q_pad * k * sizeof(uint32_t) : 0;
    int rc = ensure_ws(ix, qp_bytes + list_bytes + warm_bytes, st);
    if (rc) return rc;
    uint2 *qplanes = reinterpret_cast<uint2 *>(ix->ws);
    uint32_t *lists = reinterpret_cast<uint32_t *>((char *)ix->ws + qp_bytes);
    uint32_t *tlists = reinterpret_cast<uint32_t *>((char *)ix->ws + qp_bytes + main_bytes);
    uint32_t *wlists = warm_bytes ? reinterpret_cast<uint32_t *>((char *)ix->ws + qp_bytes + list_bytes) : nullptr;

    double t_l = now_ms();
    to_planes_kernel<<<(unsigned)((q_pad + 255) / 256), 256, 0, st>>>(d_q, q, q_pad, qplanes);
    count_launch();
    GM_CUDA(cudaMemsetAsync(lists, 0xFF, list_bytes + warm_bytes, st));
    trace("  knn: to_planes + memset enqueue", t_l);

    ScanArgs a;
    a.tplanes = ix->planes;
    a.tperm = ix->planes_perm;
    a.first_chunk = 0;
    a.tile_offset = 0;
    a.n_u = ix->n_u;
    a.qplanes = qplanes;
    a.q = q;
    a.q_pad = q_pad;
    a.k = k;
    a.L = ix->L;
    a.list_stride = q_pad;
    a.list_q0 = 0;
    a.warm_any_subset = window_warm ? 1 : 0;
    a.tidx = nullptr;
    a.prefix_c0 = 0;
    a.dbg = nullptr;
    unsigned long long *&d_dbg = ix->dbg;
    const char *dbg_env = getenv("GM_TC_DEBUG");
    const bool dbg_on = use_tc && dbg_env && dbg_env[0] == '1';
    if (dbg_on) {
        if (!d_dbg) GM_CUDA(cudaMalloc(&d_dbg, 256 * sizeof(unsigned long long)));
        GM_CUDA(cudaMemsetAsync(d_dbg, 0, 256 * sizeof(unsigned long long), st));
    }

    double pairs = 0.0;
    const int slot = prof_begin(st);
    if (warm) {
        a.n_chunks = warm;
        a.chunks_per_split = warm;
        a.lists = wlists;
        a.warm = nullptr;
        if (R == 8) launch_scan<8>(ix->metric, dim3((unsigned)tiles, 1), st, a);
        else launch_scan<4>(ix->metric, dim3((unsigned)tiles, 1), st, a);
        pairs += (double)q * (double)warm * CHUNK;
    }
    if (window_warm) {
        rc = warm_window(ix, qplanes, q, k, WINDOW, wlists, st);
        if (rc) return rc;
        pairs += (double)q * warm_copies() * WINDOW;
    }
    a.n_chunks = n_chunks;
    a.chunks_per_split = cps;
    a.first_chunk = first_chunk;
    a.lists = lists;
    a.warm = wlists;
    if (use_tc) {
        a.dbg = dbg_on ? d_dbg : nullptr;
        const unsigned grid_x = (unsigned)(q_pad / tc_query_tile());
        rc = launch_hamming_tc(dim3(grid_x - (unsigned)tail_tiles, (unsigned)splits), st, a);
        if (rc) return rc;
        if (tail_tiles) {                                           // the last query tiles, split to fill one short wave
            a.tile_offset = (int)grid_x - tail_tiles;
            a.chunks_per_split = tail_cps;
            a.lists = tlists;
            a.list_stride = tail_q;
            a.list_q0 = tail_q0;
            rc = launch_hamming_tc(dim3((unsigned)tail_tiles, (unsigned)tail_splits), st, a);
            if (rc) return rc;
        }
        if (dbg_on) {
            unsigned long long h[96];
            GM_CUDA(cudaStreamSynchronize(st));
            GM_CUDA(cudaMemcpy(h, d_dbg, sizeof h, cudaMemcpyDeviceToHost));
            // (the counters are compiled in with -DGM_TC_STATS, see tools/tc_ablate.py; otherwise they read 0)
            fprintf(stderr, "[tc_dbg] candidate events %llu (%.2f per query), list inserts %llu (%.2f per query), grid %u x %d; "
                    "epilogue warps: %.1f %% of their time behind a full candidate queue (%llu stalls)\n", h[0],
                    (double)h[0] / (double)q, h[1], (double)h[1] / (double)q, (unsigned)(q_pad / tc_query_tile()), tail_tiles ? tail_splits : splits,
                    h[4] ? 100.0 * (double)h[2] / (double)h[4] : 0.0, h[3]);
            {   // per role: share of its warps' lifetime spent inside each wait (GM_TC_STATS)
                const char *role[4] = {"read-out", "producer", "issuer", "candidate"};
                const char *what[4][3] = {{"acc_full", "-", "-"}, {"b_empty", "-", "-"}, {"b_full", "acc_empty", "issue token"}, {"-", "-", "-"}};
                for (int r = 0; r < 3; r++) {
                    const double life = (double)h[64 + r * 4 + 3];
                    fprintf(stderr, "[tc_dbg] %s warps wait:", role[r]);
                    for (int w = 0; w < 3; w++)
                        if (what[r][w][0] != '-') fprintf(stderr, " %s %.1f %%", what[r][w], life > 0 ? 100.0 * (double)h[64 + r * 4 + w] / life : 0.0);
                    fprintf(stderr, "\n");
                }
            }
            fprintf(stderr, "[tc_dbg] cycles per tile over successive 256-tile windows of one CTA:");
            for (int w = 1; w < 48 && h[8 + w]; w++) fprintf(stderr, " %.0f", (double)(h[8 + w] - h[8 + w - 1]) / 256.0);
            fprintf(stderr, "\n");
        }
    } else if (use_prefix) {
        rc = sorted_by_prefix(ix, st);
        if (rc) return rc;
        a.tplanes = ix->prefix_p;
        a.tidx = ix->prefix_i;
        // keep the states after c0 .. c0 + PFX_LEVELS - 1 bases, c0 = floor(log4 n) - 1: a sorted neighbour shares at least
        // c0 bases with probability > 0.98 (n / 4^c0 >= 4 guides per prefix)
        int lg = 0;
        while ((ix->n_u >> (2 * (lg + 1))) > 0) lg++;
        a.prefix_c0 = lg - 1 < 1 ? 1 : lg - 1;
        if (a.prefix_c0 + PFX_LEVELS > ix->L) a.prefix_c0 = ix->L - PFX_LEVELS;
        rc = R == 8 ? launch_leven_prefix<8>(dim3((unsigned)tiles, (unsigned)splits), st, a)
                    : launch_leven_prefix<4>(dim3((unsigned)tiles, (unsigned)splits), st, a);
        if (rc) return rc;
    } else if (R == 8) launch_scan<8>(ix->metric, dim3((unsigned)tiles, (unsigned)splits), st, a);
    else launch_scan<4>(ix->metric, dim3((unsigned)tiles, (unsigned)splits), st, a);
    pairs += (double)q * ((double)ix->n_u - (double)first_chunk * CHUNK);
    prof_end(slot, st, pairs);

    trace("  knn: scan launches enqueue", t_l);
    knn_merge_kernel<<<(unsigned)((q + 127) / 128), 128, 0, st>>>(lists, splits, q, q_pad, k, tlists, tail_splits, tail_q0, tail_q, d_idx, d_dist,
                                                                  dist_only);
    count_launch();
    GM_CUDA(cudaGetLastError());
    return GM_OK;
}

}  // namespace gm

using namespace gm;

extern "C" int gm_knn_engine(int engine) {
    GM_ARG(engine == 0 || engine == 1, "gm_knn_engine: 0 = XOR/POPC (INT pipe), 1 = tcgen05 one-hot GEMM (tensor pipe)");
    g_tune_engine = engine;
    return GM_OK;
}

extern "C" int gm_knn_tune(int queries_per_thread, int splits, int warm_sample) {
    GM_ARG(queries_per_thread == 0 || queries_per_thread == 4 || queries_per_thread == 8, "gm_knn_tune: queries_per_thread must be 4 or 8");
    GM_ARG(splits >= 0 && splits <= MAX_SPLITS, "gm_knn_tune: splits outside [0,%d]", MAX_SPLITS);
    if (queries_per_thread) g_tune_r = queries_per_thread;
    g_tune_splits = splits;
    g_tune_warm = warm_sample;
    return GM_OK;
}

extern "C" int gm_index_tune(void *index, int engine, int queries_per_thread, int splits, int warm_sample) {
    Index *ix = (Index *)index;
    GM_ARG(ix, "gm_index_tune: NULL index");
    GM_ARG(engine >= -1 && engine <= 1, "gm_index_tune: engine must be -1 (default), 0 or 1");
    GM_ARG(queries_per_thread == -1 || queries_per_thread == 4 || queries_per_thread == 8, "gm_index_tune: queries_per_thread must be -1, 4 or 8");
    GM_ARG(splits >= -1 && splits <= MAX_SPLITS, "gm_index_tune: splits outside [-1,%d]", MAX_SPLITS);
    GM_ARG(warm_sample >= -2, "gm_index_tune: warm_sample must be >= -2");
    ix->engine = engine;
    ix->tune_r = queries_per_thread;
    ix->tune_splits = splits;
    ix->tune_warm = warm_sample;
    return GM_OK;
}

extern "C" int gm_index_create_dev(const uint64_t *d_uniq2bit, int64_t n_u, int L, int metric, void **index, void *stream) {
    int rc = ensure_init();
    if (rc) return rc;
    GM_ARG(index, "gm_index_create: NULL handle pointer");
    *index = nullptr;
    GM_ARG(d_uniq2bit && n_u >= 1, "gm_index_create: empty guide table");
    GM_ARG(L >= 1 && L <= GM_MAX_L, "gm_index_create: L=%d outside [1,%d]", L, GM_MAX_L);
    GM_ARG(metric == GM_METRIC_HAMMING || metric == GM_METRIC_LEVEN, "gm_index_create: unknown metric %d", metric);
    if (n_u >= (1LL << IDX_BITS)) {
        set_error("gm_index_create: %lld guides exceed the 2^27 limit of the 32-bit (distance,index) key", (long long)n_u);
        return GM_ERR_RANGE;
    }
    Index *ix = new (std::nothrow) Index();
    if (!ix) { set_error("out of host memory"); return GM_ERR_NOMEM; }
    ix->n_u = n_u;
    ix->n_pad = (n_u + CHUNK - 1) / CHUNK * CHUNK;
    ix->L = L;
    ix->metric = metric;
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t e = dev_alloc((void **)&ix->planes, (size_t)ix->n_pad * sizeof(uint2), st);
    if (e != cudaSuccess) { delete ix; return cuda_fail(e, "dev_alloc(index)", __FILE__, __LINE__); }
    to_planes_kernel<<<(unsigned)((ix->n_pad + 255) / 256), 256, 0, st>>>(d_uniq2bit, n_u, ix->n_pad, ix->planes);
    count_launch();
    e = cudaGetLastError();
    if (e != cudaSuccess) { dev_free(ix->planes, st); delete ix; return cuda_fail(e, "to_planes_kernel", __FILE__, __LINE__); }
    if (metric == GM_METRIC_HAMMING) {               // K3b's bit order (knn_tc.cu)
        e = dev_alloc((void **)&ix->planes_perm, (size_t)ix->n_pad * sizeof(uint2), st);
        if (e == cudaSuccess) { tc_permute_planes(ix->planes, ix->n_pad, ix->planes_perm, st); e = cudaGetLastError(); }
        if (e != cudaSuccess) { dev_free(ix->planes_perm, st); dev_free(ix->planes, st); delete ix; return cuda_fail(e, "tc_permute_planes", __FILE__, __LINE__); }
    }
    *index = ix;
    return GM_OK;
}

extern "C" int gm_index_create(const uint64_t *uniq2bit, int64_t n_u, int L, int metric, void **index) {
    int rc = ensure_init();
    if (rc) return rc;
    GM_ARG(index, "gm_index_create: NULL handle pointer");
    *index = nullptr;
    GM_ARG(uniq2bit && n_u >= 1, "gm_index_create: empty guide table");
    uint64_t *d = nullptr;
    const double t0 = now_ms();
    GM_CUDA(dev_alloc((void **)&d, (size_t)n_u * sizeof(uint64_t), 0));
    cudaError_t e = cudaMemcpyAsync(d, uniq2bit, (size_t)n_u * sizeof(uint64_t), cudaMemcpyHostToDevice, 0);
    if (e != cudaSuccess) { dev_free(d, 0); return cuda_fail(e, "cudaMemcpy(H2D guides)", __FILE__, __LINE__); }
    rc = gm_index_create_dev(d, n_u, L, metric, index, nullptr);
    dev_free(d, 0);
    cudaError_t e2 = cudaStreamSynchronize(0);
    trace("gm_index_create", t0);
    if (rc) return rc;
    if (e2 != cudaSuccess) { gm_index_free(*index); *index = nullptr; return cuda_fail(e2, "index build", __FILE__, __LINE__); }
    return GM_OK;
}

extern "C" int gm_index_info(void *index, int64_t *n_u, int *L, int *metric) {
    Index *ix = (Index *)index;
    GM_ARG(ix, "gm_index_info: NULL index");
    if (n_u) *n_u = ix->n_u;
    if (L) *L = ix->L;
    if (metric) *metric = ix->metric;
    return GM_OK;
}

extern "C" int gm_index_free(void *index) {
    Index *ix = (Index *)index;
    if (!ix) return GM_OK;
    cudaDeviceSynchronize();
    dev_free(ix->planes, 0);
    dev_free(ix->planes_perm, 0);
    warm_free_index(ix);
    park_ws(ix->ws, ix->ws_bytes);
    if (ix->dbg) cudaFree(ix->dbg);
    delete ix;
    return GM_OK;
}

extern "C" int gm_knn_dev(void *index, const uint64_t *d_q2bit, int64_t q, int k, int32_t *d_out_idx, uint8_t *d_out_dist,
                          void *stream) {
    int rc = ensure_init();
    if (rc) return rc;
    return knn_run((Index *)index, d_q2bit, q, k, d_out_idx, d_out_dist, 0, (cudaStream_t)stream);
}

extern "C" int gm_min_dist_dev(void *index, const uint64_t *d_q2bit, int64_t q, uint8_t *d_out_dist, void *stream) {
    int rc = ensure_init();
    if (rc) return rc;
    return knn_run((Index *)index, d_q2bit, q, 1, nullptr, d_out_dist, 1, (cudaStream_t)stream);
}

static int knn_host(void *index, const uint64_t *q2bit, int64_t q, int k, int32_t *out_idx, uint8_t *out_dist, int dist_only) {
    int rc = ensure_init();
    if (rc) return rc;
    GM_ARG(index, "gm_knn: NULL index");
    GM_ARG(q >= 0, "gm_knn: negative query count");
    if (q == 0) return GM_OK;
    GM_ARG(q2bit && out_dist && (dist_only || out_idx), "gm_knn: NULL buffer");
    GM_ARG(k >= 1 && k <= GM_MAX_K, "gm_knn: k=%d outside [1,%d]", k, GM_MAX_K);
    uint64_t *d_q = nullptr;
    int32_t *d_idx = nullptr;
    uint8_t *d_dist = nullptr;
    const size_t nd = dist_only ? (size_t)q : (size_t)q * k;
    cudaStream_t st = 0;
    double t0 = now_ms();
    cudaError_t e = dev_alloc((void **)&d_q, (size_t)q * sizeof(uint64_t), st);
    if (e == cudaSuccess && !dist_only) e = dev_alloc((void **)&d_idx, nd * sizeof(int32_t), st);
    if (e == cudaSuccess) e = dev_alloc((void **)&d_dist, nd, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_q, q2bit, (size_t)q * sizeof(uint64_t), cudaMemcpyHostToDevice, st);
    trace("knn: alloc + H2D enqueue", t0);
    if (e == cudaSuccess) {
        t0 = now_ms();
        rc = knn_run((Index *)index, d_q, q, k, d_idx, d_dist, dist_only, st);
        trace("knn: launch", t0);
        if (rc == GM_OK) {
            t0 = now_ms();
            if (!dist_only) prefault(out_idx, nd * sizeof(int32_t));        // overlaps the kernels enqueued above
            prefault(out_dist, nd);
            if (!dist_only) e = cudaMemcpyAsync(out_idx, d_idx, nd * sizeof(int32_t), cudaMemcpyDeviceToHost, st);
            if (e == cudaSuccess) e = cudaMemcpyAsync(out_dist, d_dist, nd, cudaMemcpyDeviceToHost, st);
            if (e == cudaSuccess) e = cudaStreamSynchronize(st);
            trace("knn: kernels + D2H", t0);
        }
    }
    dev_free(d_q, st);
    dev_free(d_idx, st);
    dev_free(d_dist, st);
    if (rc) return rc;
    if (e != cudaSuccess) return cuda_fail(e, "gm_knn", __FILE__, __LINE__);
    return GM_OK;
}

extern "C" int gm_knn(void *index, const uint64_t *q2bit, int64_t q, int k, int32_t *out_idx, uint8_t *out_dist) {
    return knn_host(index, q2bit, q, k, out_idx, out_dist, 0);
}

extern "C" int gm_min_dist(void *index, const uint64_t *q2bit, int64_t q, uint8_t *out_dist) {
    return knn_host(index, q2bit, q, 1, nullptr, out_dist, 1);
}
