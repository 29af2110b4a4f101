// knn_tc.cu -- K3b: exact Hamming kNN as a one-hot int8 GEMM on the 5th-gen tensor cores (sm_100a).
//
// Hamming(q, t) = L - <onehot(q), onehot(t)>.  A CTA owns 256 queries as the A operand: 128 rows, TWO queries
// per row with weights 1 and 64 (bytes <= 65), K = 4 bytes per base position (+ one 16-byte chunk that carries
// per-row bias bytes).  Targets stream through shared memory as the B operand, expanded on the fly from
// their 8-byte bit planes to one-hot int8 rows (the multiply trick nibble * 0x00204081 & 0x01010101 turns four
// mask bits into four 0/1 bytes).  One `tcgen05.mma.cta_group::1.kind::i8` tile is M=128 x N=256 x K=32; the
// int32 accumulator of (row r, target n) in TMEM is
//        (m1 + b1) + 64 * (m2 + b2),      m_i = matching positions of query i,  b_i = 31 - L + tau_i,
// so bit 5 / bit 11 is set iff query 1 / 2 of the row beats its current k-th-best distance tau_i.  The epilogue
// warps read the accumulators back with `tcgen05.ld` and only OR them together: ONE 3-input LOP3 per two
// accumulators (= four comparisons); a set flag bit sends the thread to the exact insertion path, which
// recomputes the distance from the planes.  The accumulator is a filter, never the reported value, and a stale
// (looser) bias only produces extra candidates -- results stay bit-exact with K3a.  The candidate path touches only
// shared memory: the raw planes of the last eight target tiles sit in a small ring written by the producers (a
// producer can only reach tile j after the MMA of tile j-4 was issued, which waited for the epilogue of tile j-6),
// and the top-k lists of the CTA's 512 queries live in shared memory until the end of the pass.
//
// Warp roles (448 threads, one CTA per SM).  A CTA owns TWO query tiles (512 queries); each has its own 256-column
// accumulator buffer in TMEM and its own four epilogue warps, so eight warps keep tcgen05.ld requests in flight and
// every target tile is expanded once but multiplied twice:
//   warps 0-3 / 4-7  epilogue of query tile 0 / 1: warp w owns TMEM lanes 32(w%4)..+31 = rows; thread r owns the
//                    two queries of row r (exclusive owner of their lists and thresholds -> no races, in-order inserts)
//   warps 8-11       producers: planes -> one-hot B tile in shared memory (canonical no-swizzle K-major layout)
//   warps 12-13      one MMA issuer per query tile (a single thread each); warp 12 also owns the TMEM allocation;
//                    tcgen05.commit publishes accumulator tiles and releases shared-memory stages
// Bound: TMEM read-out (measured ~170 B/clk/SM => ~86 comparisons/clk/SM at 2 queries per accumulator), with the
// MMA pipe at ~50 % (3 MMAs of 128 clk per 757-clk tile).
#include "knn_common.cuh"

namespace gm {

// Shape of the pipeline.  Measured on B200 (tools/probes/tc_probe.cu, GM_TC_DEBUG=1): issuing a tcgen05.mma costs
// ~65 cycles whatever its N and a tcgen05.commit ~95, and TMEM read-out scales with the number of reading warps.
//   shape 2 (default): TWO query tiles per CTA, N = 128, 2 x 2 x 128 accumulator columns, 8 epilogue warps, 2 MMA
//                      issuer warps                                   -> 140 ms on the 6.3 Mb config (1.33e13 cmp/s)
//   shape 1          : ONE query tile, N = 256, 2 x 256 columns, 4 epilogue warps, 8 producer warps: half the MMA
//                      instructions per comparison, but four warps cannot drain TMEM fast enough next to the running
//                      MMA                                            -> 209 ms
#ifndef GM_TC_SHAPE
#define GM_TC_SHAPE 2
#endif
static constexpr int TC_M = 128;
#if GM_TC_SHAPE == 1
static constexpr int TC_SETS = 1;          // query tiles per CTA
static constexpr int TC_N = 256;           // targets per MMA tile
static constexpr int TC_PROD_WARPS = 8;    // one target per producer thread per tile
static constexpr int TC_STAGES = 3;
#else
static constexpr int TC_SETS = 2;
static constexpr int TC_N = 128;
static constexpr int TC_PROD_WARPS = 4;
static constexpr int TC_STAGES = 4;
#endif
static constexpr int TC_QT = 256 * TC_SETS;   // queries per CTA (2 per row, 128 rows per query tile)
static constexpr int TC_RING = 8;          // raw-plane ring depth >= TC_STAGES + 2 (power of two), see the kernel header
static constexpr int TC_PROD_WARP0 = 4 * TC_SETS;                  // epilogue warps come first (TMEM lane quadrant = warp % 4)
static constexpr int TC_MMA_WARP = TC_PROD_WARP0 + TC_PROD_WARPS;  // first of TC_SETS MMA-issuer warps (one per query tile)
static constexpr int TC_THREADS = 32 * (TC_MMA_WARP + TC_SETS);
static_assert(TC_SETS * 2 * TC_N == 512, "accumulator buffers must tile the 512 TMEM columns");
static_assert(TC_N % 128 == 0, "the epilogue reads 128 columns per round");
static constexpr uint32_t TC_FLAGS = 0x08200820u;   // bit 5 / bit 11 of both 16-bit halves of a packed register

int tc_query_tile() { return TC_QT; }

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

#define TC_LD_X32(r, taddr)                                                                                       \
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15," \
                 "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"                       \
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), \
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]),       \
                   "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]),     \
                   "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]),     \
                   "=r"(r[29]), "=r"(r[30]), "=r"(r[31])                                                          \
                 : "r"(taddr)                                                                                     \
                 : "memory")

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

// four 0/1 bytes from bits [4j, 4j+4) of a position mask
__device__ __forceinline__ uint32_t nibble_bytes(uint32_t mask, int j) {
    return (((mask >> (4 * j)) & 0xFu) * 0x00204081u) & 0x01010101u;
}
// one 16-byte K chunk: positions 4j..4j+3, words = bases A, C, G, T
__device__ __forceinline__ uint4 onehot_chunk(uint32_t eA, uint32_t eC, uint32_t eG, uint32_t eT, int j) {
    return make_uint4(nibble_bytes(eA, j), nibble_bytes(eC, j), nibble_bytes(eG, j), nibble_bytes(eT, j));
}

// Per-role cycle counters and a short event timeline of block (0,0); compiled in only with -DGM_TC_INSTRUMENT
// (then enabled at run time by GM_TC_DEBUG=1).  They are what the numbers in the header and DESIGN.md come from.
#ifdef GM_TC_INSTRUMENT
#define TC_TL_FIRST 2000
#define TC_TL(role, i, ev) do { if (dbg && (i) >= TC_TL_FIRST && (i) < TC_TL_FIRST + 8) a.dbg[64 + (role) * 32 + ((i) - TC_TL_FIRST) * 4 + (ev)] = (unsigned long long)clock64(); } while (0)
#define TC_T0(v) long long v = dbg ? clock64() : 0
#define TC_ADD(acc, v) do { if (dbg) acc += (unsigned long long)(clock64() - v); } while (0)
#define TC_DBG_ON (a.dbg != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && lane == 0)
#else
#define TC_TL(role, i, ev) do { } while (0)
#define TC_T0(v) do { } while (0)
#define TC_ADD(acc, v) do { } while (0)
#define TC_DBG_ON false
#endif

struct TcState {            // per epilogue thread: its two queries
    uint32_t qlo[2], qhi[2], tau[2];
    uint32_t *list[2];
};

// Sorted insert into a list kept in shared memory with stride TC_QT between consecutive ranks (bank-conflict free
// across the lanes of a warp).  Returns the distance of the (new) worst entry, 31 while the list is not full.
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
    return worst >> IDX_BITS;
}

// Candidate path, entered warp-uniformly when any lane's OR over a 32-column chunk has a flag bit.  The accumulators
// are not needed any more.  For every flagged lane (usually one) the WHOLE warp cooperates: lane j holds the raw
// planes of target j of the chunk (from the shared-memory ring), the flagged lane broadcasts its two queries and
// thresholds, all 32 exact distances are evaluated at once (2 LOP3 + 1 POPC each) and two ballots tell the owner
// which targets beat its thresholds; the owner inserts them in ascending index into its shared-memory lists.
// No loop over registers, no divergence, a few dozen instructions per event.
__device__ __noinline__ void tc_candidates(uint32_t flagged, uint32_t t0, const uint2 *ring_chunk, TcState &s, int k, uint32_t n_u,
                                           uint8_t *bias_bytes, int L, int lane) {
    const uint2 tp = ring_chunk[lane];
    const bool valid = t0 + (uint32_t)lane < n_u;
    while (flagged) {
        const int src = __ffs(flagged) - 1;
        flagged &= flagged - 1;
        uint32_t hits[2];       // (evaluating only the query whose flag bit fired was measured: the extra shuffle costs more)
#pragma unroll
        for (int e = 0; e < 2; e++) {
            const uint32_t ql = __shfl_sync(0xFFFFFFFFu, s.qlo[e], src), qh = __shfl_sync(0xFFFFFFFFu, s.qhi[e], src);
            const uint32_t ta = __shfl_sync(0xFFFFFFFFu, s.tau[e], src);
            const uint32_t d = (uint32_t)hamming_planes(ql, qh, tp.x, tp.y);
            hits[e] = __ballot_sync(0xFFFFFFFFu, valid && d < ta);
        }
        if (lane == src && (hits[0] | hits[1])) {
            bool changed = false;
#pragma unroll
            for (int e = 0; e < 2; e++) {
                uint32_t hm = hits[e];
                while (hm) {
                    const int i = __ffs(hm) - 1;
                    hm &= hm - 1;
                    const uint2 ti = ring_chunk[i];
                    const uint32_t d = (uint32_t)hamming_planes(s.qlo[e], s.qhi[e], ti.x, ti.y);
                    if (d < s.tau[e]) {
                        const uint32_t w = list_insert_smem(s.list[e], k, (d << IDX_BITS) | (t0 + (uint32_t)i));
                        if (w < s.tau[e]) { s.tau[e] = w; changed = true; }
                    }
                }
            }
            if (changed) {    // tighten the bias bytes of this row; later MMAs pick them up (a stale value is only looser)
                bias_bytes[0] = (uint8_t)(31 - L + (int)s.tau[0]);
                bias_bytes[1] = (uint8_t)(31 - L + (int)s.tau[1]);
                fence_async_smem();
            }
        }
        __syncwarp();
    }
}

template <int kc /* 16-byte K chunks, even */>
__global__ void __launch_bounds__(TC_THREADS, 1) knn_hamming_tc_kernel(const ScanArgs a) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t b_full[TC_STAGES], b_empty[TC_STAGES], acc_full[TC_SETS][2], acc_empty[TC_SETS][2];
    __shared__ uint32_t s_tmem;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const bool dbg = TC_DBG_ON;
    const long long t_start = dbg ? clock64() : 0;
    const int L = a.L;
    const int nd = (L + 3) >> 2;                      // data chunks; chunk nd carries the bias bytes
    const uint32_t lmask = (1u << L) - 1u;
    const uint32_t a_bytes = (uint32_t)TC_M * 16u * (uint32_t)kc;
    const uint32_t b_bytes = (uint32_t)TC_N * 16u * (uint32_t)kc;
    uint8_t *sA = smem;                               // TC_SETS query tiles
    uint8_t *sB = smem + TC_SETS * a_bytes;           // TC_STAGES target tiles
    uint2 *ring = reinterpret_cast<uint2 *>(sB + TC_STAGES * b_bytes);                  // raw planes of the last TC_RING tiles
    uint32_t *s_lists = reinterpret_cast<uint32_t *>(ring + TC_RING * TC_N);            // [k][TC_QT] top-k keys

    const int c0 = blockIdx.y * a.chunks_per_split;
    const int c1 = min(c0 + a.chunks_per_split, a.n_chunks);
    if (c0 >= c1) return;
    const int tile0 = c0 * (CHUNK / TC_N), n_tiles = (c1 - c0) * (CHUNK / TC_N);
    const int64_t qbase = (int64_t)blockIdx.x * TC_QT;

    if (tid == 0) {
        // arrivals: one elected lane per warp (4 producer warps / 4 epilogue warps per set); b_empty gets one
        // tcgen05.commit from each MMA issuer
        for (int s = 0; s < TC_STAGES; s++) { mbar_init(&b_full[s], TC_PROD_WARPS); mbar_init(&b_empty[s], TC_SETS); }
        for (int q = 0; q < TC_SETS; q++)
            for (int b = 0; b < 2; b++) { mbar_init(&acc_full[q][b], 1); mbar_init(&acc_empty[q][b], 4); }
        mbar_fence_init();
    }
    if (warp == TC_MMA_WARP) tc_alloc(&s_tmem, 512);

    // ---- A tiles: set s, row r = query (qbase + 256 s + r) * 1 + query (qbase + 256 s + 128 + r) * 64 ------------
    TcState st;
    const int set = warp >> 2;                        // valid for the epilogue warps
    const int row = tid & 127;
    if (warp < 4 * TC_SETS) {
        uint32_t e[2][4];
#pragma unroll
        for (int s = 0; s < 2; s++) {
            const int64_t qi = qbase + (int64_t)set * 256 + (int64_t)s * 128 + row;      // < q_pad by construction
            const uint2 p = a.qplanes[qi];
            st.qlo[s] = p.x;
            st.qhi[s] = p.y;
            uint32_t t = 31u;
            if (a.warm) t = min((a.warm[(size_t)qi * a.k + (a.k - 1)] >> IDX_BITS) + 1u, 31u);
            st.tau[s] = qi < a.q ? t : 0u;
            st.list[s] = s_lists + set * 256 + s * 128 + row;
            for (int j = 0; j < a.k; j++) st.list[s][j * TC_QT] = KEY_EMPTY;
            e[s][0] = ~(p.x | p.y) & lmask; e[s][1] = p.x & ~p.y; e[s][2] = p.y & ~p.x; e[s][3] = p.x & p.y;
        }
        uint8_t *myA = sA + (size_t)set * a_bytes;
#pragma unroll
        for (int j = 0; j < kc; j++) {
            uint4 w = make_uint4(0u, 0u, 0u, 0u);
            if (j < nd) {
                const uint4 w0 = onehot_chunk(e[0][0], e[0][1], e[0][2], e[0][3], j);
                const uint4 w1 = onehot_chunk(e[1][0], e[1][1], e[1][2], e[1][3], j);
                w = make_uint4(w0.x + 64u * w1.x, w0.y + 64u * w1.y, w0.z + 64u * w1.z, w0.w + 64u * w1.w);
            } else if (j == nd) {
                w.x = (uint32_t)(31 - L + (int)st.tau[0]) | ((uint32_t)(31 - L + (int)st.tau[1]) << 8);
            }
            *reinterpret_cast<uint4 *>(myA + (size_t)j * (TC_M * 16) + (size_t)row * 16) = w;
        }
        fence_async_smem();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = s_tmem;
    if (dbg && tid == 0) a.dbg[24] = (unsigned long long)(clock64() - t_start);
    const long long t_role = dbg ? clock64() : 0;

    if (warp < 4 * TC_SETS) {
        unsigned long long c_wait = 0, c_ld = 0, c_cand = 0, n_cand = 0, c_arr = 0;
        // ================= epilogue set `set`: two 128-column accumulator buffers, its own 256 queries ===============
        uint8_t *bias_bytes = sA + (size_t)set * a_bytes + (size_t)nd * (TC_M * 16) + (size_t)row * 16;
        const uint32_t lane_addr = tmem + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)set * (2 * TC_N);
        uint32_t ra[32], rb[32];
        for (int i = 0; i < n_tiles; i++) {
            const int buf = i & 1;
            TC_TL(set, i, 3);
            { TC_T0(tw); mbar_wait(&acc_full[set][buf], (uint32_t)((i >> 1) & 1)); TC_ADD(c_wait, tw); }
            TC_TL(set, i, 0);
            tc_fence_after();
            const uint32_t taddr = lane_addr + (uint32_t)buf * TC_N;
            const uint32_t t0 = (uint32_t)(tile0 + i) * TC_N;
            const uint2 *ring_tile = ring + (size_t)(i & (TC_RING - 1)) * TC_N;
#pragma unroll 1
            for (int h = 0; h < TC_N; h += 128) {                   // 128 columns per round: two packed loads in flight
                TC_LD_X32_PACK(ra, taddr + h);                      // columns h .. h+63
                TC_LD_X32_PACK(rb, taddr + h + 64);                 // columns h+64 .. h+127
                { TC_T0(tw); tc_wait_ld(); TC_ADD(c_ld, tw); }
                // pack::16b puts adjacent columns in one register: registers 0..15 = columns 0..31, 16..31 = columns 32..63
#pragma unroll
                for (int half = 0; half < 4; half++) {
                    const uint32_t *r = half < 2 ? ra : rb;
                    const int o = (half & 1) * 16;
                    uint32_t f = 0;
#pragma unroll
                    for (int x = 0; x < 16; x += 2) f |= r[o + x] | r[o + x + 1];
                    const uint32_t fl = __ballot_sync(0xFFFFFFFFu, (f & TC_FLAGS) != 0);
                    if (fl) {
                        TC_T0(tw);
                        tc_candidates(fl, t0 + h + half * 32, ring_tile + h + half * 32, st, a.k, (uint32_t)a.n_u, bias_bytes, L, lane);
                        TC_ADD(c_cand, tw);
                    }
                }
            }
            TC_TL(set, i, 1);
            TC_T0(ta);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&acc_empty[set][buf]);
            TC_ADD(c_arr, ta);
            TC_TL(set, i, 2);
        }
#pragma unroll
        for (int s = 0; s < 2; s++) {                               // publish the finished lists
            const int64_t qi = qbase + (int64_t)set * 256 + (int64_t)s * 128 + row;
            uint32_t *dst = a.lists + ((size_t)blockIdx.y * a.q_pad + qi) * a.k;
            for (int j = 0; j < a.k; j++) dst[j] = st.list[s][j * TC_QT];
        }
        if (dbg && tid == 0) {
            a.dbg[0] = (unsigned long long)(clock64() - t_role); a.dbg[1] = c_wait; a.dbg[2] = c_ld; a.dbg[3] = c_cand; a.dbg[4] = n_cand;
            a.dbg[5] = (unsigned long long)n_tiles; a.dbg[6] = c_arr;
        }
    } else if (warp < TC_MMA_WARP) {
        unsigned long long c_wait = 0, c_exp = 0, c_fence = 0, c_arr = 0;
        // ================= producers: planes -> one-hot B tile (one target per thread) =========================
        static_assert(TC_N == 32 * TC_PROD_WARPS, "one target per producer thread per tile");
        const int p = tid - 32 * TC_PROD_WARP0;
        uint2 tnext = a.tplanes[(size_t)tile0 * TC_N + p];
        for (int i = 0; i < n_tiles; i++) {
            const int s = i % TC_STAGES;
            const uint32_t round = (uint32_t)(i / TC_STAGES);
            if (round > 0) { TC_T0(tw); mbar_wait(&b_empty[s], (round - 1) & 1u); TC_ADD(c_wait, tw); }
            TC_TL(2, i, 0);
            const uint2 tp = tnext;
            if (i + 1 < n_tiles) tnext = a.tplanes[(size_t)(tile0 + i + 1) * TC_N + p];   // prefetch the next tile's planes
            uint8_t *dst = sB + (size_t)s * b_bytes;
            TC_T0(te);
            ring[(size_t)(i & (TC_RING - 1)) * TC_N + p] = tp;      // the slot's previous tile (i - TC_RING) is retired, see header
            const uint32_t eA = ~(tp.x | tp.y) & lmask, eC = tp.x & ~tp.y, eG = tp.y & ~tp.x, eT = tp.x & tp.y;
            uint8_t *dstp = dst + (size_t)p * 16;
#pragma unroll
            for (int j = 0; j < kc; j++) {                          // positions beyond L have all-zero masks
                uint4 w = onehot_chunk(eA, eC, eG, eT, j);
                if (j == nd) w.x = 1u | (64u << 8);                 // multiplies the bias bytes of A: 1 * b1 + 64 * b2
                *reinterpret_cast<uint4 *>(dstp + j * (TC_N * 16)) = w;
            }
            TC_ADD(c_exp, te);
            TC_TL(2, i, 1);
            TC_T0(tf);
            fence_async_smem();                                     // generic-proxy writes -> visible to the tensor core
            TC_ADD(c_fence, tf);
            TC_T0(ta);
            __syncwarp();
            if (lane == 0) mbar_arrive(&b_full[s]);
            TC_ADD(c_arr, ta);
            TC_TL(2, i, 2);
        }
        if (dbg && tid == 32 * TC_PROD_WARP0) { a.dbg[8] = (unsigned long long)(clock64() - t_role); a.dbg[9] = c_wait; a.dbg[10] = c_exp; a.dbg[11] = c_fence; a.dbg[12] = c_arr; }
    } else {
        // ================= MMA issuers: warp TC_MMA_WARP + q feeds query tile q ============================================
        // The WHOLE warp runs this loop and every operand is made provably warp-uniform (__shfl_sync from lane 0), so the
        // descriptors live in uniform registers and one elect.sync lane issues.  Issuing from `if (lane == 0)` made the
        // compiler wrap every tcgen05.mma in an ELECT + 5 x R2UR.BROADCAST waterfall loop (~65 cycles per instruction).
        const int q = __shfl_sync(0xFFFFFFFFu, warp, 0) - TC_MMA_WARP;
        const uint32_t tmem_u = __shfl_sync(0xFFFFFFFFu, tmem, 0);
        const uint32_t sA_u = __shfl_sync(0xFFFFFFFFu, smem_u32(sA), 0), sB_u = __shfl_sync(0xFFFFFFFFu, smem_u32(sB), 0);
        const uint32_t bar_full_u = __shfl_sync(0xFFFFFFFFu, smem_u32(&acc_full[0][0]), 0);
        const uint32_t bar_bempty_u = __shfl_sync(0xFFFFFFFFu, smem_u32(&b_empty[0]), 0);
        const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(TC_N >> 3) << 17) | ((uint32_t)(TC_M >> 4) << 24);
        // descriptors differ only in the start-address field (low 14 bits, units of 16 bytes): build once, add offsets
        const uint64_t da = tc_desc(sA_u + (uint32_t)q * a_bytes, TC_M * 16, 128);
        const uint64_t db0 = tc_desc(sB_u, TC_N * 16, 128);
        const uint32_t a_ks = (2u * TC_M * 16u) >> 4, b_stage = b_bytes >> 4, b_ks = (2u * TC_N * 16u) >> 4;
        constexpr int n_ks = kc / 2;
        const bool leader = elect_one();
        unsigned long long c_wb = 0, c_wa = 0;
        int s = 0;
        uint32_t full_parity = 0;
        for (int i = 0; i < n_tiles; i++) {
            const int buf = i & 1;
            { TC_T0(tw); mbar_wait(&b_full[s], full_parity); TC_ADD(c_wb, tw); }
            TC_TL(3 + q, i, 0);
            if (i >= 2) { TC_T0(tw); mbar_wait(&acc_empty[q][buf], (uint32_t)(((i >> 1) - 1) & 1)); TC_ADD(c_wa, tw); }
            TC_TL(3 + q, i, 1);
            tc_fence_after();
            const uint32_t d = tmem_u + (uint32_t)q * (2 * TC_N) + (uint32_t)buf * TC_N;
            const uint64_t db = db0 + (uint64_t)((uint32_t)s * b_stage);
            const uint32_t bar_f = bar_full_u + (uint32_t)((q * 2 + buf) * 8), bar_e = bar_bempty_u + (uint32_t)(s * 8);
            if (leader) {
#pragma unroll
                for (int ks = 0; ks < n_ks; ks++)
                    tc_mma_i8(d, da + (uint64_t)((uint32_t)ks * a_ks), db + (uint64_t)((uint32_t)ks * b_ks), idesc, ks > 0 ? 1u : 0u);
                tc_commit_addr(bar_f);                              // accumulator tile ready for this set's epilogue warps
                tc_commit_addr(bar_e);                              // this issuer is done with the smem stage
            }
            __syncwarp();
            TC_TL(3 + q, i, 2);
            if (++s == TC_STAGES) { s = 0; full_parity ^= 1u; }
        }
        if (dbg && q == 0) { a.dbg[16] = (unsigned long long)(clock64() - t_role); a.dbg[17] = c_wb; a.dbg[18] = c_wa; }
    }

    const long long t_td = dbg ? clock64() : 0;
    tc_fence_before();
    __syncthreads();
    if (dbg && tid == 0) a.dbg[25] = (unsigned long long)(clock64() - t_td);
    if (warp == TC_MMA_WARP) tc_dealloc(tmem, 512);
}

static size_t tc_smem_bytes(int kc, int k) {
    size_t need = (size_t)TC_SETS * TC_M * 16 * kc + (size_t)TC_STAGES * TC_N * 16 * kc + (size_t)TC_RING * TC_N * 8 +
                  (size_t)k * TC_QT * 4;
    const size_t one_cta_per_sm = 116 * 1024;      // > half of 227 KB: a second CTA (and its TMEM alloc) can never co-reside
    return need > one_cta_per_sm ? need : one_cta_per_sm;
}

// ---- tensor-pipe roofline denominator: back-to-back kind::i8 MMAs (128x256x32) on every SM ----------------------
__global__ void __launch_bounds__(128, 1) mb_mma_i8_kernel(int n_mma, unsigned long long *cycles) {
    extern __shared__ __align__(1024) uint8_t smem[];          // operands: contents irrelevant for throughput
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t s_base;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < (128 + 256) * 32 / 4; i += 128) reinterpret_cast<uint32_t *>(smem)[i] = 0x01010101u;
    if (tid == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
    fence_async_smem();
    if (warp == 0) tc_alloc(&s_base, 256);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    long long t0 = 0;
    if (tid == 0) {
        const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((256u >> 3) << 17) | ((128u >> 4) << 24);
        const uint64_t da = tc_desc(smem_u32(smem), 128 * 16, 128), db = tc_desc(smem_u32(smem) + 128 * 32, 256 * 16, 128);
        t0 = clock64();
        for (int i = 0; i < n_mma; i++) tc_mma_i8(s_base, da, db, idesc, i > 0 ? 1u : 0u);
        tc_commit(&bar);
    }
    mbar_wait(&bar, 0);
    if (tid == 0) cycles[blockIdx.x] = (unsigned long long)(clock64() - t0);
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tc_dealloc(s_base, 256);
}

// int8 tensor ops/s (2 per MAC) of the whole GPU, timed with CUDA events
int microbench_mma_i8(double *ops_per_s) {
    static bool attr_set = false;
    if (!attr_set) {
        GM_CUDA(cudaFuncSetAttribute(mb_mma_i8_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 120 * 1024));
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
        mb_mma_i8_kernel<<<grid, 128, 116 * 1024>>>(n_mma, d);      // > half the SM's shared memory: one CTA per SM
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
    *ops_per_s = 2.0 * 128.0 * 256.0 * 32.0 * (double)n_mma * grid / (best * 1e-3);
    return GM_OK;
}

template <int KC>
static int launch_tc_kc(dim3 grid, cudaStream_t st, const ScanArgs &a) {
    static bool attr_set = false;
    if (!attr_set) {
        GM_CUDA(cudaFuncSetAttribute(knn_hamming_tc_kernel<KC>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
        attr_set = true;
    }
    knn_hamming_tc_kernel<KC><<<grid, TC_THREADS, tc_smem_bytes(KC, a.k), st>>>(a);
    count_launch();
    return GM_OK;
}

int launch_hamming_tc(dim3 grid, cudaStream_t st, const ScanArgs &a) {
    const int nd = (a.L + 3) / 4;
    const int kc = ((nd + 1) + 1) & ~1;         // data chunks + bias chunk, rounded up to whole 32-byte MMA K steps
    switch (kc) {
    case 2: return launch_tc_kc<2>(grid, st, a);
    case 4: return launch_tc_kc<4>(grid, st, a);
    case 6: return launch_tc_kc<6>(grid, st, a);
    default: return launch_tc_kc<8>(grid, st, a);
    }
}

}  // namespace gm
