// knn_tc.cu -- K3b: exact Hamming kNN as a one-hot int8 GEMM on the 5th-gen tensor cores (sm_100a).
//
// Hamming(q, t) = L - m(q, t), m = number of matching positions, and m is a genuine dense contraction.  The encoding
// is the rank-minimal one (the 4 x 4 "same base" matrix needs three dimensions plus a constant):
//        target base  ->  g = ([t=C], [t=G], [t=T])                      pure 0/1 bytes (the hot, per-tile expansion)
//        query base   ->  f = ([q=C]-[q=A], [q=G]-[q=A], [q=T]-[q=A])     bytes in {-1, 0, 1}
//        <f(q), g(t)> = [q=t] - [q=A]   =>   sum over positions = m - countA(q),
// and countA(q) is a per-query constant that is folded into the query's bias byte.  THREE bytes per base position
// instead of the four of a one-hot row: a 20-nt guide is 60 bytes + 2 bias bytes = TWO 32-byte MMA K steps, not three.
// A CTA owns 512 queries as two A operands ("sets") of 128 rows, TWO queries per row with weights 1 and 64 (bytes in
// [-65, 65]).  Targets stream through shared memory as the B operand, expanded on the fly from their 8-byte bit planes.
// One `tcgen05.mma.cta_group::1.kind::i8` is M=128 x N=128 x K=32; the int32 accumulator of (row r, target n) in TMEM is
//        (m1 + b1) + 64 * (m2 + b2),      m_i = matching positions of query i,  b_i = 31 - L + tau_i,
// so bit 5 / bit 11 is set iff query 1 / 2 of the row is closer to the target than its current bound tau_i (the
// distance of its k-th best so far).  The accumulator is ONLY a filter: the epilogue warps read it back with
// `tcgen05.ld ... pack::16b`, OR the registers together (one 3-input LOP3 per four accumulators = eight comparisons),
// and hand every 32-target chunk with a set flag bit to the candidate warps, which recompute exact distances from the
// planes and maintain the top-k lists.  A stale (looser) bias only produces extra candidates, so results are
// bit-identical to K3a.
//
// What bounds this kernel is not arithmetic but the latency of the hand-offs between the roles (measured on B200 with
// tools/probes/tc_latency.cu: 3 MMAs + tcgen05.commit -> wake of a waiting warp 363 cycles, 219 of them the MMAs;
// mbarrier.arrive -> try_wait wake in another warp 177; tcgen05.ld x32 + wait 124; try_wait on a completed phase 37;
// the proxy fence after the producers' stores ~150), so every role is spread over enough warps that its per-tile
// chain may take two tile periods, and nothing with a data-dependent duration sits on the MMA <-> epilogue cycle:
//
//   warps  0-15  epilogue: warp w reads TMEM lanes 32(w%4)..+31 (rows) of set (w>>3), accumulator buffer (w>>2)&1 --
//                i.e. the even and the odd target tiles of a set have their own four warps.  A warp waits for its
//                buffer, pulls 128 columns into registers, RELEASES the buffer at once (the next MMA can start) and only
//                then ORs/ballots; flagged chunks go into a small shared-memory queue as (first target, row mask) --
//                one 8-byte store with a generation tag, no fence.
//   warps 16-23  producers, two groups of four taking alternate tiles: planes -> one-hot B tile (one target per
//                thread), canonical no-swizzle K-major layout, then proxy fence + arrive.
//   warps 24-27  MMA issuers, one per (set, accumulator buffer) (elect.sync lane, operands in uniform registers), each
//                taking every other tile of its set: wait for the B stage and for the buffer, three MMAs, ONE
//                tcgen05.commit that publishes the buffer.  A commit stalls the issuing thread until its MMAs have
//                completed (gm_microbench 9-11: three MMAs take 228 cycles of one issuing thread, 467 with one commit
//                behind them, 604 with two), so the shared-memory stage is NOT released by a second commit but by a
//                plain arrive two tiles later, when the issuer has seen the buffer of that tile read out; six stages keep
//                the producers four tiles ahead all the same.  Warp 24 owns the TMEM allocation.  The two issuers of a
//                set pass an issue token, so the set's tiles enter the tensor pipe in ascending order.
//   warps 28-31  candidate warps: warp c serves the queues of TMEM quadrant c (both sets, both buffers) and is the
//                exclusive owner of the lists and bounds of those 128 queries.  Per event lane j loads target j of the
//                chunk (L2), all 32 exact distances of a flagged row's two queries are evaluated at once (2 LOP3 + POPC),
//                hits are inserted by full (distance, index) key into lists kept in shared memory (so the order in
//                which events of different tiles are served does not matter), and a tightened bound is written back
//                into the bias byte of A, where later MMAs pick it up.  Idle candidate warps poll every 2 us: every
//                poll costs shared-memory cycles.
//
// K layout.  A row of K bytes is a string of 32-bit words: word 3g + x holds, for the four positions 4g .. 4g+3, the
// bytes of base x (0 = C, 1 = G, 2 = T); the LAST word of the row carries the bias bytes (A side: b1, b2; B side: 1, 64).
// The index keeps a second copy of the planes with position p stored at bit (p>>2) + 8(p&3): then word (g, x) of a
// target is (mask_x >> g) & 0x01010101 -- two instructions per four bytes.  Hamming distance is invariant under a common
// bit permutation, so the candidate warps work on the permuted planes too (queries are permuted once per CTA).

#include "knn_common.cuh"

namespace gm {

// Timing-only ablations (tools/tc_ablate.py; results are wrong by construction): 1 = epilogue skips the TMEM loads,
// 2 = producers skip the one-hot expansion, 4 = issuers skip the MMAs (commits only), 8 = candidate path disabled,
// 16 = producers skip the proxy fence, 32 = producers do not wait for the stage to be released,
// 128 = candidate warps pop events but do not serve them, 256 = one MMA K step fewer per tile,
// 512 = flags are evaluated and voted on but nothing is queued.
#ifndef GM_TC_ABL
#define GM_TC_ABL 0
#endif
static constexpr int TC_M = 128;           // rows per A operand
static constexpr int TC_N = 128;           // targets per tile
static constexpr int TC_SETS = 2;          // A operands (query tiles) per CTA
#ifndef GM_TC_TOKEN_FENCE
#define GM_TC_TOKEN_FENCE 0
#endif
#ifndef GM_TC_STATS_CTA
#define GM_TC_STATS_CTA 200                 // the CTA whose tile-rate profile -DGM_TC_STATS records
#endif
#ifndef GM_TC_STAGES
#define GM_TC_STAGES 6
#endif
static constexpr int TC_STAGES = GM_TC_STAGES;   // B tiles in shared memory (even: a stage always serves the same producer group)
static constexpr int TC_QT = 256 * TC_SETS;                        // queries per CTA
static constexpr int TC_EPI_WARPS = 8 * TC_SETS;                   // (set, buffer, quadrant)
static constexpr int TC_PROD_WARP0 = TC_EPI_WARPS;
static constexpr int TC_PROD_WARPS = 8;                            // two groups of four, alternate tiles
static constexpr int TC_MMA_WARP = TC_PROD_WARP0 + TC_PROD_WARPS;  // first of the issuer warps
static constexpr int TC_MMA_WARPS = 2 * TC_SETS;                   // (set, buffer): each issues every other tile of its set
static constexpr int TC_CAND_WARP0 = TC_MMA_WARP + TC_MMA_WARPS;
static constexpr int TC_CAND_WARPS = 4;                            // one per TMEM lane quadrant
static constexpr int TC_THREADS = 32 * (TC_CAND_WARP0 + TC_CAND_WARPS);
static constexpr int TC_QN = 64;           // entries per candidate queue (power of two)
#ifndef GM_TC_IDLE_NS
#define GM_TC_IDLE_NS 2000
#endif
static constexpr unsigned TC_IDLE_NS = GM_TC_IDLE_NS;   // idle candidate warps poll their queues this often
static_assert(TC_SETS * 2 * TC_N == 512, "accumulator buffers must tile the 512 TMEM columns");
static_assert(TC_THREADS <= 1024 && TC_EPI_WARPS * 32 == TC_QT, "one epilogue thread per query in the prologue");
static constexpr uint32_t TC_FLAGS = 0x08200820u;   // bit 5 / bit 11 of both 16-bit halves of a packed register

int tc_query_tile() { return TC_QT; }
// 16-byte K chunks per row: 3 words per group of four positions + the bias word, rounded up to whole 32-byte MMA K steps
// (L <= 8: 2, L <= 20: 4, L <= 27: 6)
int tc_k_chunks(int L) {
    const int words = 3 * ((L + 3) / 4) + 1;
    return (((words + 3) / 4) + 1) & ~1;
}

// ---- PTX wrappers ------------------------------------------------------------------------------------
__device__ __forceinline__ void tc_alloc(uint32_t *dst_smem, uint32_t cols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tc_dealloc(uint32_t addr, uint32_t cols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// one lane of a fully active warp (elect.sync): the issuing lane of the MMA warps
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xFFFFFFFF;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void tc_commit_addr(uint32_t bar_smem_addr) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar_smem_addr) : "memory");
}
__device__ __forceinline__ void tc_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_mma_i8(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
        : "memory");
}
// UMMA shared-memory descriptor, SWIZZLE_NONE, K-major: [0,14) start>>4, [16,30) LBO>>4 = stride between the two
// 16-byte K chunks of one MMA, [32,46) SBO>>4 = stride between 8-row groups, [46,48) version 1 (sm_100).
__device__ __forceinline__ uint64_t tc_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((smem_addr & 0x3FFFFu) >> 4) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32) | ((uint64_t)1 << 46);
}

// packed read: 32 registers cover 64 columns (two 16-bit values per register; the accumulators are <= 4030)
#define TC_LD_X32_PACK(r, taddr)                                                                                  \
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.pack::16b.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15," \
                 "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"                       \
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), \
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]),       \
                   "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]),     \
                   "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]),     \
                   "=r"(r[29]), "=r"(r[30]), "=r"(r[31])                                                          \
                 : "r"(taddr)                                                                                     \
                 : "memory")

// position p -> bit (p >> 2) + 8 (p & 3)   (L <= 27 < 32)
__host__ __device__ inline uint32_t tc_permute_bits(uint32_t x) {
    uint32_t r = 0;
    for (int p = 0; p < 32; p++) r |= ((x >> p) & 1u) << ((p >> 2) + 8 * (p & 3));
    return r;
}
__global__ void tc_permute_kernel(const uint2 *__restrict__ in, int64_t n, uint2 *__restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = make_uint2(tc_permute_bits(in[i].x), tc_permute_bits(in[i].y));
}
int tc_permute_planes(const uint2 *planes, int64_t n, uint2 *out, cudaStream_t st) {
    tc_permute_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(planes, n, out);
    count_launch();
    return GM_OK;
}

// word w (< 3 * groups) of a K row from the permuted base masks e[0..2] = C, G, T: base w % 3 of positions 4g .. 4g+3,
// g = w / 3.  w is a compile-time constant at every call site (unrolled loops).
__device__ __forceinline__ uint32_t k_word(const uint32_t (&e)[3], int w) {
    return (e[w % 3] >> (w / 3)) & 0x01010101u;
}

// a spin loop outlived ~10 s: fail the launch instead of hanging the GPU (1 = issue order, 2 = candidate queue, 3 = mbarrier)
__device__ __noinline__ void tc_watchdog(int what) {
    printf("libgm_b200: K3b watchdog %d fired (block %d,%d thread %d)\n", what, blockIdx.x, blockIdx.y, threadIdx.x);
    __trap();
}
// mbarrier wait for the role loops: the first try is inline, the retry loop lives in an outlined function so that it
// costs the callers no registers.  The retry passes a suspend-time hint (the warp stays suspended until the phase completes
// or ~1 ms passes, instead of the short default limit) and the loop body is three instructions: in the round-1 form
// (default limit, clock64() watchdog in every iteration: 15 instructions) the retry loops were 25 % of ALL warp
// instructions the kernel executed (ncu source page, profiles/r02_ncu_full_knn_hamming_tc_c5.csv) -- issue slots the
// producers and the read-out warps compete for.  The watchdog counts retries: 2^14 x ~1 ms ~ 16 s.
__device__ __forceinline__ bool mbar_try_wait_long(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}\n"
        : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity), "r"(1000000u) : "memory");
    return ok != 0;
}
__device__ __noinline__ void tc_wait_slow(uint64_t *bar, uint32_t parity) {
    uint32_t spins = 0;
    while (!mbar_try_wait_long(bar, parity))
        if (++spins == (1u << 14)) tc_watchdog(3);
}
__device__ __forceinline__ void tc_wait(uint64_t *bar, uint32_t parity) {
    if (!mbar_try_wait(bar, parity)) tc_wait_slow(bar, parity);
}
// -DGM_TC_STATS: cycles a role spends inside a wait, accumulated per call site (slot) for the whole grid
#ifdef GM_TC_STATS
#define TC_WAIT_T(slot, bar, parity)                                                                      \
    do {                                                                                                  \
        const long long w0__ = clock64();                                                                 \
        tc_wait(bar, parity);                                                                             \
        if (lane == 0 && a.dbg) wait_clk[slot] += (unsigned long long)(clock64() - w0__);                 \
    } while (0)
#else
#define TC_WAIT_T(slot, bar, parity) tc_wait(bar, parity)
#endif
// OR of 16 registers as a depth-3 tree of 3-input LOP3s (a serial chain would be 8 deep)
__device__ __forceinline__ uint32_t tc_or16(const uint32_t *r) {
    const uint32_t a = r[0] | r[1] | r[2], b = r[3] | r[4] | r[5], c = r[6] | r[7] | r[8], d = r[9] | r[10] | r[11], e = r[12] | r[13] | r[14];
    return (a | b | c) | (d | e | r[15]);
}
__device__ __forceinline__ uint32_t ld_vol(const uint32_t *p) { return *reinterpret_cast<const volatile uint32_t *>(p); }
__device__ __forceinline__ void st_vol(uint32_t *p, uint32_t v) { *reinterpret_cast<volatile uint32_t *>(p) = v; }

// Sorted insert by full key into a list kept in shared memory with stride TC_QT between consecutive ranks.
// Returns the key of the (new) worst entry, KEY_EMPTY while the list is not full.
__device__ __noinline__ uint32_t list_insert_smem(uint32_t *lst, int k, uint32_t key) {
    uint32_t worst = lst[(k - 1) * TC_QT];
    if (key < worst) {
        int pos = k - 1;
        while (pos > 0) {
            const uint32_t v = lst[(pos - 1) * TC_QT];
            if (v <= key) break;
            lst[pos * TC_QT] = v;
            pos--;
        }
        lst[pos * TC_QT] = key;
        worst = lst[(k - 1) * TC_QT];
    }
    return worst;
}

template <int kc /* 16-byte K chunks, even */>
__global__ void __launch_bounds__(TC_THREADS, 1) knn_hamming_tc_kernel(const ScanArgs a) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t b_full[TC_STAGES], b_empty[TC_STAGES], acc_full[TC_SETS][2], acc_empty[TC_SETS][2];
    __shared__ uint32_t s_tmem, s_done, s_issued[TC_SETS];
    __shared__ uint32_t q_head[TC_EPI_WARPS];                             // candidate queues: the consumers' cursors

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int L = a.L;
    constexpr int kwords = 4 * kc;                    // 32-bit words per K row; data words 0 .. 3*ceil(L/4)-1 (< kwords - 1)
    constexpr int kb = kc - 1;                        // the LAST word of the LAST chunk carries the bias bytes: a compile-time position
    const uint32_t lmask = tc_permute_bits((1u << L) - 1u);
    const uint32_t a_bytes = (uint32_t)TC_M * 16u * (uint32_t)kc;
    const uint32_t b_bytes = (uint32_t)TC_N * 16u * (uint32_t)kc;
    uint8_t *sA = smem;                               // TC_SETS query tiles
    uint8_t *sB = smem + TC_SETS * a_bytes;           // TC_STAGES target tiles
    uint2 *sQ = reinterpret_cast<uint2 *>(sB + TC_STAGES * b_bytes);                    // permuted planes of the CTA's queries
    uint32_t *sBound = reinterpret_cast<uint32_t *>(sQ + TC_QT);                        // key bound per query: insert iff key < bound
    uint2 *sQueue = reinterpret_cast<uint2 *>(sBound + TC_QT);                          // [TC_EPI_WARPS][TC_QN] (first target, row mask)
    uint32_t *s_lists = reinterpret_cast<uint32_t *>(sQueue + TC_EPI_WARPS * TC_QN);    // [k][TC_QT] top-k keys

    const int c0 = a.first_chunk + blockIdx.y * a.chunks_per_split;
    const int c1 = min(c0 + a.chunks_per_split, a.n_chunks);
    if (c0 >= c1) return;
    const int tile0 = c0 * (CHUNK / TC_N), n_tiles = (c1 - c0) * (CHUNK / TC_N);
    const int64_t qbase = ((int64_t)blockIdx.x + a.tile_offset) * TC_QT;

    if (tid == 0) {
        for (int s = 0; s < TC_STAGES; s++) { mbar_init(&b_full[s], 4); mbar_init(&b_empty[s], TC_SETS); }   // b_empty: one arrive per set
        for (int q = 0; q < TC_SETS; q++)
            for (int b = 0; b < 2; b++) { mbar_init(&acc_full[q][b], 1); mbar_init(&acc_empty[q][b], 4); }
        mbar_fence_init();
        s_done = 0;
        for (int q = 0; q < TC_SETS; q++) s_issued[q] = 0;
    }
    if (tid < TC_EPI_WARPS) q_head[tid] = 0;
    for (int i = tid; i < TC_EPI_WARPS * TC_QN; i += TC_THREADS) sQueue[i] = make_uint2(0u, 0u);   // generation 0 = nothing yet
    if (warp == TC_MMA_WARP) tc_alloc(&s_tmem, 512);

    // ---- per query: permuted planes, list, key bound ---------------------------------------------------------------
    // Split 0 inherits the warm-start lists (the main scan starts behind the warm sample); the other splits start
    // empty with the bound (w0 << IDX_BITS): a target at the warm k-th distance w0 loses the tie against the k warm
    // entries, which have lower indices.
    if (tid < TC_QT) {
        const int64_t qi = qbase + tid;               // < q_pad by construction
        const uint2 p = a.qplanes[qi];
        const int qslot = ((tid >> 8) * 128 + (tid & 127)) * 2 + ((tid >> 7) & 1);    // the two queries of a row are adjacent
        sQ[qslot] = make_uint2(tc_permute_bits(p.x), tc_permute_bits(p.y));
        uint32_t bound = KEY_EMPTY;
        uint32_t *lst = s_lists + tid;
        if (a.warm) {
            const uint32_t *w = a.warm + (size_t)qi * a.k;
            const uint32_t wk = w[a.k - 1];
            if (a.warm_any_subset) {
                // the warm entries may sit anywhere in the table: start empty and let every guide at a distance <= w0
                // through (at least k of them exist, so the list fills and the usual tightening takes over)
                for (int j = 0; j < a.k; j++) lst[j * TC_QT] = KEY_EMPTY;
                bound = wk == KEY_EMPTY ? KEY_EMPTY : ((wk >> IDX_BITS) + 1u) << IDX_BITS;
            } else if (blockIdx.y == 0) {
                for (int j = 0; j < a.k; j++) lst[j * TC_QT] = w[j];
                bound = wk;
            } else {
                for (int j = 0; j < a.k; j++) lst[j * TC_QT] = KEY_EMPTY;
                bound = wk == KEY_EMPTY ? KEY_EMPTY : (wk >> IDX_BITS) << IDX_BITS;
            }
        } else {
            for (int j = 0; j < a.k; j++) lst[j * TC_QT] = KEY_EMPTY;
        }
        sBound[qslot] = qi < a.q ? bound : 0u;        // padding queries never match
    }
    __syncthreads();
    // ---- A tiles: set s, row r = query (256 s + r) * 1 + query (256 s + 128 + r) * 64 ------------------------------
    if (tid < TC_SETS * TC_M) {
        const int set = tid >> 7, row = tid & 127;
        uint32_t e[2][3], eA[2], tau[2];
#pragma unroll
        for (int s = 0; s < 2; s++) {
            const uint2 p = sQ[(set * 128 + row) * 2 + s];
            tau[s] = sBound[(set * 128 + row) * 2 + s] >> IDX_BITS;
            eA[s] = ~(p.x | p.y) & lmask; e[s][0] = p.x & ~p.y; e[s][1] = p.y & ~p.x; e[s][2] = p.x & p.y;
        }
        uint8_t *myA = sA + (size_t)set * a_bytes;
#pragma unroll
        for (int j = 0; j < kc; j++) {
            uint32_t w[4];
#pragma unroll
            for (int x = 0; x < 4; x++) {
                const int wi = 4 * j + x;
                if (wi == kwords - 1) {
                    // bias bytes: 31 - L + tau + countA(query) (the data words sum to m - countA)
                    w[x] = (uint32_t)(31 - L + (int)tau[0] + __popc(eA[0])) | ((uint32_t)(31 - L + (int)tau[1] + __popc(eA[1])) << 8);
                } else {
                    // per byte: ([q1 = base] - [q1 = A]) + 64 ([q2 = base] - [q2 = A]), two's complement, no carries between bytes
                    const uint32_t a0 = (eA[0] >> (wi / 3)) & 0x01010101u, a1 = (eA[1] >> (wi / 3)) & 0x01010101u;
                    w[x] = __vsub4(k_word(e[0], wi) + (k_word(e[1], wi) << 6), a0 + (a1 << 6));
                }
            }
            *reinterpret_cast<uint4 *>(myA + (size_t)j * (TC_M * 16) + (size_t)row * 16) = make_uint4(w[0], w[1], w[2], w[3]);
        }
        fence_async_smem();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = s_tmem;
#ifdef GM_TC_STATS
    unsigned long long wait_clk[4] = {0, 0, 0, 0};      // per role: [0..1] its waits, [2] loop time
    const long long role_t0 = clock64();
#endif

    if (warp < TC_EPI_WARPS) {
        // ================= epilogue: (set, buffer, quadrant) = (warp >> 3, (warp >> 2) & 1, warp & 3) ===================
        const int set = warp >> 3, par = (warp >> 2) & 1, quad = warp & 3;
        const uint32_t taddr = tmem + ((uint32_t)(quad * 32) << 16) + (uint32_t)set * (2 * TC_N) + (uint32_t)par * TC_N;
        uint2 *queue = sQueue + warp * TC_QN;
        uint32_t tail = 0;
#ifdef GM_TC_STATS                                                  // -DGM_TC_STATS + GM_TC_DEBUG=1: queue stalls, tile-rate profile
        uint32_t stall_clk = 0, n_stalls = 0;                       // lane 0: time (units of 64 cycles) spent behind a full candidate queue
        const long long t_begin = clock64();
#endif
        uint32_t r[32];
        for (int i = par; i < n_tiles; i += 2) {
            TC_WAIT_T(0, &acc_full[set][par], (uint32_t)((i >> 1) & 1));
            tc_fence_after();
#ifdef GM_TC_STATS
            if (a.dbg && (i & 255) == 0 && (i >> 8) < 48 && blockIdx.x + a.tile_offset == GM_TC_STATS_CTA && warp == 0 && lane == 0)     // tile-rate profile of one CTA
                a.dbg[8 + (i >> 8)] = (unsigned long long)(clock64() - t_begin);
#endif
            uint32_t f[4];
#if GM_TC_ABL & 1
            f[0] = f[1] = f[2] = f[3] = 0u;
#else
            // 64 columns per packed load: registers 0..15 = columns 0..31, 16..31 = columns 32..63.  Everything between the
            // buffer's commit and its release is on the MMA -> read-out -> MMA round trip, so the first load is reduced
            // by a depth-3 OR tree (the registers are needed for the second load) and the second load's registers are only
            // reduced AFTER the release.  (A serial OR chain of all 64 registers before the release cost 12 ms of 66.)
            TC_LD_X32_PACK(r, taddr);
            tc_wait_ld();
            f[0] = tc_or16(r);
            f[1] = tc_or16(r + 16);
            TC_LD_X32_PACK(r, taddr + 64);
            tc_wait_ld();
#endif
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&acc_empty[set][par]);       // buffer released before any further work
#if !(GM_TC_ABL & 1)
            f[2] = tc_or16(r);
            f[3] = tc_or16(r + 16);
#endif
            // One vote says whether anything in the tile is flagged (usually not); then one event per flagged chunk.
            // (One event per tile with the union of the rows was measured slower: the candidate warp's extra shared-memory
            // reads cost more than the pushes they save -- shared-memory cycles are what this kernel runs out of.)
#if GM_TC_ABL & 512                                                  // flags evaluated and voted on, nothing queued
            if (__any_sync(0xFFFFFFFFu, ((f[0] | f[1] | f[2] | f[3]) & TC_FLAGS) != 0)) tail++;
#endif
            if (__any_sync(0xFFFFFFFFu, ((f[0] | f[1] | f[2] | f[3]) & TC_FLAGS) != 0) && !(GM_TC_ABL & (8 | 512))) {
                const uint32_t t0 = (uint32_t)(tile0 + i) * TC_N;
#pragma unroll
                for (int h = 0; h < 4; h++) {
                    const uint32_t fl = __ballot_sync(0xFFFFFFFFu, (f[h] & TC_FLAGS) != 0);
                    if (fl == 0) continue;
                    if (lane == 0) {
                        if (tail - ld_vol(&q_head[warp]) >= (uint32_t)TC_QN) {      // queue full: wait for the candidate warp
                            const long long w0 = clock64();
                            while (tail - ld_vol(&q_head[warp]) >= (uint32_t)TC_QN) {
                                __nanosleep(256);
                                if (clock64() - w0 > 20000000000LL) tc_watchdog(2);
                            }
#ifdef GM_TC_STATS
                            stall_clk += (uint32_t)((clock64() - w0) >> 6);
                            n_stalls++;
#endif
                        }
                        // One 8-byte store publishes the event: bit 31 of the first word is a generation tag that flips on
                        // every lap of the ring, so the consumer polls the slot itself and no fence (MEMBAR, ~200 cycles
                        // on this warp's chain) or tail pointer is needed.  (Target indices are < 2^27.)
                        const uint32_t gen = ((tail / TC_QN) & 1u) ^ 1u;
                        *reinterpret_cast<volatile unsigned long long *>(&queue[tail & (TC_QN - 1)]) =
                            (unsigned long long)((t0 + (uint32_t)h * 32u) | (gen << 31)) | ((unsigned long long)fl << 32);
                    }
                    tail++;
                }
            }
        }
        __syncwarp();
#if GM_TC_ABL & 512
        if (tail == 0xFFFFFFFFu) a.lists[0] = tail;                 // keeps the vote alive
#endif
        if (lane == 0) { __threadfence_block(); atomicAdd(&s_done, 1u); }
#ifdef GM_TC_STATS
        if (a.dbg && lane == 0) {
            atomicAdd(&a.dbg[2], (unsigned long long)stall_clk << 6);
            atomicAdd(&a.dbg[3], (unsigned long long)n_stalls);
            atomicAdd(&a.dbg[4], (unsigned long long)(clock64() - t_begin));
        }
#endif
    } else if (warp < TC_MMA_WARP) {
        // ================= producers: planes -> one-hot B tile (one target per thread, group g takes tiles g, g+2, ..) ===
        const int pw = warp - TC_PROD_WARP0, grp = pw >> 2, p = (pw & 3) * 32 + lane;
        const uint2 *tsrc = a.tperm + (size_t)tile0 * TC_N + p;
        uint2 tnext = grp < n_tiles ? tsrc[(size_t)grp * TC_N] : make_uint2(0u, 0u);
        for (int i = grp; i < n_tiles; i += 2) {
            const int s = i % TC_STAGES;
            const uint32_t round = (uint32_t)(i / TC_STAGES);
            const uint2 tp = tnext;
            if (i + 2 < n_tiles) tnext = tsrc[(size_t)(i + 2) * TC_N];       // prefetch this thread's next target
            const uint32_t e[3] = {tp.x & ~tp.y, tp.y & ~tp.x, tp.x & tp.y};     // C, G, T (an A is three zero bytes)
            if (round > 0 && !(GM_TC_ABL & 32)) TC_WAIT_T(0, &b_empty[s], (round - 1) & 1u);
            uint8_t *dstp = sB + (size_t)s * b_bytes + (size_t)p * 16;
#pragma unroll
            for (int j = 0; j < ((GM_TC_ABL & 2) ? 0 : kc); j++) {  // positions beyond L have all-zero masks
                uint4 w = make_uint4(k_word(e, 4 * j), k_word(e, 4 * j + 1), k_word(e, 4 * j + 2), k_word(e, 4 * j + 3));
                if (j == kb) w.w = 1u | (64u << 8);                 // multiplies the bias bytes of A: 1 * b1 + 64 * b2
                *reinterpret_cast<uint4 *>(dstp + j * (TC_N * 16)) = w;
            }
            if (!(GM_TC_ABL & 16)) fence_async_smem();              // generic-proxy writes -> visible to the tensor core
            __syncwarp();
            if (lane == 0) mbar_arrive(&b_full[s]);
        }
    } else if (warp < TC_CAND_WARP0) {
        // ================= MMA issuers: warp TC_MMA_WARP + 2 par + q issues tiles par, par + 2, .. of set q =============
        // The WHOLE warp runs this loop and every operand is made provably warp-uniform (__shfl_sync from lane 0), so the
        // descriptors live in uniform registers and one elect.sync lane issues.  Issuing from `if (lane == 0)` made the
        // compiler wrap every tcgen05.mma in an ELECT + 5 x R2UR.BROADCAST waterfall loop (~65 cycles per instruction).
        const int mw = __shfl_sync(0xFFFFFFFFu, warp, 0) - TC_MMA_WARP;
        const int q = mw & 1, par = mw >> 1;
        const uint32_t tmem_u = __shfl_sync(0xFFFFFFFFu, tmem, 0);
        const uint32_t sA_u = __shfl_sync(0xFFFFFFFFu, smem_u32(sA), 0), sB_u = __shfl_sync(0xFFFFFFFFu, smem_u32(sB), 0);
        const uint32_t bar_f = __shfl_sync(0xFFFFFFFFu, smem_u32(&acc_full[q][par]), 0);
        const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(TC_N >> 3) << 17) | ((uint32_t)(TC_M >> 4) << 24);
        // descriptors differ only in the start-address field (low 14 bits, units of 16 bytes): build once, add offsets
        const uint64_t da = tc_desc(sA_u + (uint32_t)q * a_bytes, TC_M * 16, 128);
        const uint64_t db0 = tc_desc(sB_u, TC_N * 16, 128);
        const uint32_t a_ks = (2u * TC_M * 16u) >> 4, b_stage = b_bytes >> 4, b_ks = (2u * TC_N * 16u) >> 4;
        constexpr int n_ks = kc / 2;
        const bool leader = elect_one();
        const uint32_t d = tmem_u + (uint32_t)q * (2 * TC_N) + (uint32_t)par * TC_N;
        for (int i = par; i < n_tiles; i += 2) {
            const int s = i % TC_STAGES;
            TC_WAIT_T(0, &b_full[s], (uint32_t)((i / TC_STAGES) & 1));
            if (i >= 2) {
                TC_WAIT_T(1, &acc_empty[q][par], (uint32_t)(((i >> 1) - 1) & 1));
                // Tile i - 2 of this set has been multiplied AND read out, so the set is done with its B stage: release it
                // here with a plain arrive.  (A second tcgen05.commit right after the MMAs would release it earlier, but a
                // commit stalls the issuing thread until its MMAs have completed -- measured: 3 MMAs + 1 commit take 467
                // cycles of a single issuing thread, + 2 commits 604, against 228 without -- and the issue path is what
                // this kernel runs out of.  TC_STAGES = 6 keeps the producers four tiles ahead all the same.)
                if (leader) mbar_arrive(&b_empty[(i - 2) % TC_STAGES]);
            }
            // The two issuers of a set take turns: tiles enter the tensor pipe in ascending order, so the bias bytes an
            // MMA reads can only reflect list entries from EARLIER tiles (lower target indices) -- what makes "flag iff
            // strictly closer than the k-th best" exact under the (distance, index) order.
            if (ld_vol(&s_issued[q]) != (uint32_t)i) {
                const long long w0 = clock64();
#ifdef GM_TC_STATS
                const long long w0s = w0;
#endif
                while (ld_vol(&s_issued[q]) != (uint32_t)i) {
                    __nanosleep(20);                                // (a tight loop would eat shared-memory cycles)
                    if (clock64() - w0 > 20000000000LL) tc_watchdog(1);
                }
#ifdef GM_TC_STATS
                if (lane == 0 && a.dbg) wait_clk[2] += (unsigned long long)(clock64() - w0s);
#endif
            }
            tc_fence_after();
            const uint64_t db = db0 + (uint64_t)((uint32_t)s * b_stage);
            if (leader) {
#pragma unroll
                for (int ks = 0; ks < ((GM_TC_ABL & 4) ? 0 : (GM_TC_ABL & 256) ? n_ks - 1 : n_ks); ks++)
                    tc_mma_i8(d, da + (uint64_t)((uint32_t)ks * a_ks), db + (uint64_t)((uint32_t)ks * b_ks), idesc, ks > 0 ? 1u : 0u);
#if GM_TC_TOKEN_FENCE
                __threadfence_block();
#endif
                st_vol(&s_issued[q], (uint32_t)i + 1u);             // the set's other issuer may go ahead
                tc_commit_addr(bar_f);                              // accumulator buffer ready for its epilogue warps (which
                                                                    // also release the smem stage)
            }
            __syncwarp();
        }
    } else {
        // ================= candidate warps: warp c owns the queries of TMEM quadrant c of both sets ====================
        const int c = warp - TC_CAND_WARP0;
#ifdef GM_TC_STATS
        unsigned long long n_events = 0, n_inserts = 0;
#define TC_STAT(x) x
#else
#define TC_STAT(x)
#endif
        // one event = (first target of a 32-target chunk, mask of flagged rows); `set` selects the A operand
        auto serve = [&](const uint2 ev, const uint2 tp, const int set) {
            uint32_t fl = ev.y;
            const uint32_t kpart = ev.x + (uint32_t)lane;
            const bool valid = (int64_t)kpart < a.n_u;
            while (fl) {
                const int row = c * 32 + (__ffs(fl) - 1);
                fl &= fl - 1;
                // both queries of the row at once: one 16-byte and one 8-byte shared-memory read, two ballots
                const uint4 qq = reinterpret_cast<const uint4 *>(sQ)[set * 128 + row];
                const uint2 bb = reinterpret_cast<const uint2 *>(sBound)[set * 128 + row];
                const uint32_t key[2] = {((uint32_t)hamming_planes(qq.x, qq.y, tp.x, tp.y) << IDX_BITS) | kpart,
                                         ((uint32_t)hamming_planes(qq.z, qq.w, tp.x, tp.y) << IDX_BITS) | kpart};
                const uint32_t hit[2] = {__ballot_sync(0xFFFFFFFFu, valid && key[0] < bb.x),
                                         __ballot_sync(0xFFFFFFFFu, valid && key[1] < bb.y)};
#pragma unroll
                for (int e = 0; e < 2; e++) {
                    uint32_t hits = hit[e];
                    if (hits == 0) continue;
                    const uint32_t bound0 = e ? bb.y : bb.x;
                    uint32_t bound = bound0;
                    uint32_t *lst = s_lists + set * 256 + e * 128 + row;
                    while (hits) {                                  // ascending target index
                        const int i = __ffs(hits) - 1;
                        hits &= hits - 1;
                        const uint32_t ki = __shfl_sync(0xFFFFFFFFu, key[e], i);
                        if (ki < bound) {
                            uint32_t w = 0;
                            if (lane == 0) w = list_insert_smem(lst, a.k, ki);
                            w = __shfl_sync(0xFFFFFFFFu, w, 0);
                            bound = min(bound, w);
                            TC_STAT(n_inserts++);
                        }
                    }
                    if (bound != bound0 && lane == 0) {
                        sBound[(set * 128 + row) * 2 + e] = bound;
                        // tighten the bias byte; later MMAs pick it up.  No proxy fence: the byte only has to reach the
                        // tensor core eventually (measured: the fence changes nothing)
                        if ((bound >> IDX_BITS) != (bound0 >> IDX_BITS))
                            sA[(size_t)set * a_bytes + (size_t)kb * (TC_M * 16) + (size_t)row * 16 + 12 + e] =
                                (uint8_t)(31 - L + (int)(bound >> IDX_BITS) + __popc(~((e ? qq.z : qq.x) | (e ? qq.w : qq.y)) & lmask));
                    }
                    __syncwarp();
                }
            }
        };
        uint32_t head[4] = {0u, 0u, 0u, 0u};
        for (;;) {
            const uint32_t done = ld_vol(&s_done);                  // read BEFORE the scan: done + empty queues = finished
            // poll the next slot of each of the four queues (x = (set, buffer)); a slot is valid when its generation tag
            // matches the lap the cursor is on
            uint2 ev[4], tp[4];
            bool have[4];
            bool any = false;
#pragma unroll
            for (int x = 0; x < 4; x++) {
                const int qid = (x >> 1) * 8 + (x & 1) * 4 + c;
                const unsigned long long w = *reinterpret_cast<const volatile unsigned long long *>(&sQueue[qid * TC_QN + (head[x] & (TC_QN - 1))]);
                ev[x] = make_uint2((uint32_t)w, (uint32_t)(w >> 32));
                have[x] = (ev[x].x >> 31) == (((head[x] / TC_QN) & 1u) ^ 1u) && ev[x].y != 0u;
                any |= have[x];
            }
            if (!any) {
                if (done == (uint32_t)TC_EPI_WARPS) break;
                // every poll costs shared-memory cycles the tensor core needs for its operands: sleep generously
                __nanosleep(TC_IDLE_NS);
                continue;
            }
#pragma unroll
            for (int x = 0; x < 4; x++)
                if (have[x]) {
                    ev[x].x &= 0x7FFFFFFFu;
                    tp[x] = a.tperm[(size_t)ev[x].x + lane];          // the loads of all taken events are in flight together
                    head[x]++;
                }
            __syncwarp();
            if (lane == 0) {
#pragma unroll
                for (int x = 0; x < 4; x++)
                    if (have[x]) st_vol(&q_head[(x >> 1) * 8 + (x & 1) * 4 + c], head[x]);   // entries copied: slots may be reused
            }
#pragma unroll
            for (int x = 0; x < 4; x++)
                if (have[x]) { if (!(GM_TC_ABL & 128)) serve(ev[x], tp[x], x >> 1); TC_STAT(n_events++); }
        }
        TC_STAT(if (a.dbg && lane == 0) { atomicAdd(&a.dbg[0], n_events); atomicAdd(&a.dbg[1], n_inserts); })
    }

#ifdef GM_TC_STATS
    if (a.dbg && lane == 0) {
        const int role = warp < TC_EPI_WARPS ? 0 : warp < TC_MMA_WARP ? 1 : warp < TC_CAND_WARP0 ? 2 : 3;
        atomicAdd(&a.dbg[64 + role * 4 + 0], wait_clk[0]);
        atomicAdd(&a.dbg[64 + role * 4 + 1], wait_clk[1]);
        atomicAdd(&a.dbg[64 + role * 4 + 2], wait_clk[2]);
        atomicAdd(&a.dbg[64 + role * 4 + 3], (unsigned long long)(clock64() - role_t0));
    }
#endif
    tc_fence_before();
    __syncthreads();
    if (tid < TC_QT) {                                              // publish the finished lists
        uint32_t *dst = a.lists + ((size_t)blockIdx.y * a.list_stride + (size_t)(qbase + tid - a.list_q0)) * a.k;
        for (int j = 0; j < a.k; j++) dst[j] = s_lists[j * TC_QT + tid];
    }
    if (warp == TC_MMA_WARP) tc_dealloc(tmem, 512);
}

static size_t tc_smem_bytes(int kc, int k) {
    size_t need = (size_t)TC_SETS * TC_M * 16 * kc + (size_t)TC_STAGES * TC_N * 16 * kc + (size_t)TC_QT * (8 + 4) +
                  (size_t)TC_EPI_WARPS * TC_QN * 8 + (size_t)k * TC_QT * 4;
    const size_t one_cta_per_sm = 116 * 1024;      // > half of 227 KB: a second CTA (and its TMEM alloc) can never co-reside
    return need > one_cta_per_sm ? need : one_cta_per_sm;
}

// ---- tensor-pipe roofline denominator: back-to-back kind::i8 MMAs (128 x N x 32) on every SM -------------------
// A_TMEM = false: both operands from shared memory (SS, what the kNN kernel issues); true: A from tensor memory (TS).
__device__ __forceinline__ void tc_mma_i8_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(tmem_d), "r"(tmem_a), "l"(db), "r"(idesc), "r"(accumulate)
        : "memory");
}
template <int N, bool A_TMEM, bool CYCLE = false, int COMMITS = 0>
__global__ void __launch_bounds__(128, 1) mb_mma_i8_kernel(int n_mma, unsigned long long *cycles) {
    extern __shared__ __align__(1024) uint8_t smem[];          // operands: contents irrelevant for throughput
    __shared__ __align__(8) uint64_t bar, side[8];             // side: targets of the COMMITS extra commits per three MMAs
    __shared__ uint32_t s_base;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < (CYCLE ? 64 * 1024 : (128 + 256) * 32) / 4; i += 128) reinterpret_cast<uint32_t *>(smem)[i] = 0x01010101u;
    if (tid == 0) {
        mbar_init(&bar, 1);
        for (int b = 0; b < 8; b++) mbar_init(&side[b], 1);
        mbar_fence_init();
    }
    fence_async_smem();
    if (warp == 0) tc_alloc(&s_base, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    long long t0 = 0;
    if (tid == 0) {
        const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | (((uint32_t)N >> 3) << 17) | ((128u >> 4) << 24);
        const uint64_t da = tc_desc(smem_u32(smem), 128 * 16, 128), db = tc_desc(smem_u32(smem) + 128 * 32, N * 16, 128);
        t0 = clock64();
        // two accumulator tiles alternate so consecutive MMAs do not depend on each other
        // CYCLE: like the kNN kernel, consecutive MMAs read different operand tiles (3 A tiles in the first 12 KB,
        // 12 B tiles behind them), so no operand can be reused from one MMA to the next
        int ia = 0, ib = 0;
        for (int i = 0; i < n_mma; i++) {
            const uint32_t d = s_base + (N == 128 ? (uint32_t)(i & 1) * 128u : 0u);
            const uint64_t oa = CYCLE ? (uint64_t)(ia * (4096 >> 4)) : 0, ob = CYCLE ? (uint64_t)((3 + ib) * (4096 >> 4)) : 0;
            if (A_TMEM) tc_mma_i8_ts(d, s_base + 256 + (CYCLE ? (uint32_t)ia * 8u : 0u), db + ob - (CYCLE ? 128 * 32 / 16 : 0), idesc, i > 1 ? 1u : 0u);
            else tc_mma_i8(d, da + oa, db + ob - (CYCLE ? 128 * 32 / 16 : 0), idesc, i > 1 ? 1u : 0u);
            if (++ia == 3) {
                ia = 0;
                // the kNN kernel's pattern: COMMITS tcgen05.commit after every group of three MMAs (nobody waits on them here)
                for (int c = 0; c < COMMITS; c++) tc_commit(&side[(i + c) & 7]);
            }
            if (++ib == 12) ib = 0;
        }
        tc_commit(&bar);
    }
    mbar_wait(&bar, 0);
    if (tid == 0) cycles[blockIdx.x] = (unsigned long long)(clock64() - t0);
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tc_dealloc(s_base, 512);
}

// int8 tensor ops/s (2 per MAC) of the whole GPU, timed with CUDA events.  variant 0: N = 256 SS (the roofline
// denominator), 1: N = 128 SS, 2: N = 128 TS, 3: N = 256 TS (A operand in tensor memory).
template <int N, bool A_TMEM, bool CYCLE = false, int COMMITS = 0>
static int microbench_mma_i8_t(double *ops_per_s) {
    static bool attr_set = false;
    if (!attr_set) {
        GM_CUDA(cudaFuncSetAttribute(mb_mma_i8_kernel<N, A_TMEM, CYCLE, COMMITS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 120 * 1024));
        attr_set = true;
    }
    unsigned long long *d = nullptr;
    const int grid = device_sm_count(), n_mma = 16384;
    GM_CUDA(dev_alloc((void **)&d, sizeof(unsigned long long) * grid, 0));
    cudaEvent_t e0, e1;
    GM_CUDA(cudaEventCreate(&e0));
    GM_CUDA(cudaEventCreate(&e1));
    float best = 1e30f;
    for (int rep = 0; rep < 3; rep++) {
        GM_CUDA(cudaEventRecord(e0));
        mb_mma_i8_kernel<N, A_TMEM, CYCLE, COMMITS><<<grid, 128, 116 * 1024>>>(n_mma, d);      // > half the SM's shared memory: one CTA per SM
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
    *ops_per_s = 2.0 * 128.0 * (double)N * 32.0 * (double)n_mma * grid / (best * 1e-3);
    return GM_OK;
}
int microbench_mma_i8(int variant, double *ops_per_s) {
    switch (variant) {
    case 1: return microbench_mma_i8_t<128, false>(ops_per_s);
    case 2: return microbench_mma_i8_t<128, true>(ops_per_s);
    case 3: return microbench_mma_i8_t<256, true>(ops_per_s);
    case 4: return microbench_mma_i8_t<128, false, true>(ops_per_s);
    case 5: return microbench_mma_i8_t<128, true, true>(ops_per_s);
    case 6: return microbench_mma_i8_t<128, false, true, 1>(ops_per_s);
    case 7: return microbench_mma_i8_t<128, false, true, 2>(ops_per_s);
    case 8: return microbench_mma_i8_t<64, false, true, 2>(ops_per_s);
    default: return microbench_mma_i8_t<256, false>(ops_per_s);
    }
}

template <int KC>
static int launch_tc_kc(dim3 grid, cudaStream_t st, const ScanArgs &a) {
    static bool attr_set = false;
    if (!attr_set) {
        cudaFuncAttributes fa;
        GM_CUDA(cudaFuncGetAttributes(&fa, knn_hamming_tc_kernel<KC>));
        const int room = 227 * 1024 - (int)((fa.sharedSizeBytes + 1023) / 1024 * 1024);     // static arrays come first, dynamic part 1 KB aligned
        GM_CUDA(cudaFuncSetAttribute(knn_hamming_tc_kernel<KC>, cudaFuncAttributeMaxDynamicSharedMemorySize, room));
        attr_set = true;
    }
    knn_hamming_tc_kernel<KC><<<grid, TC_THREADS, tc_smem_bytes(KC, a.k), st>>>(a);
    count_launch();
    return GM_OK;
}

int launch_hamming_tc(dim3 grid, cudaStream_t st, const ScanArgs &a) {
    switch (tc_k_chunks(a.L)) {
    case 2: return launch_tc_kc<2>(grid, st, a);
    case 4: return launch_tc_kc<4>(grid, st, a);
    default: return launch_tc_kc<6>(grid, st, a);
    }
}

}  // namespace gm
