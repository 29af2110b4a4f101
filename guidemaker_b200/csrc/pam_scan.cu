// pam_scan.cu -- K1: IUPAC-degenerate PAM scan over both strands (sm_100a).
//
// Replaces the four regex.finditer(..., overlapped=True) generators of PamTarget.find_targets
// (core.py:142-246) together with their per-hit slicing, reverse-complementing and check_target.
//
// Pass 0  encode : ASCII genome (read once, 16-byte vector loads) -> three bit planes lo/hi/valid,
//                  32 positions per 32-bit word.  Only upper-case A/C/G/T are valid
//                  (core.py:118-121, :138); everything else (N, lower case, record separators)
//                  clears the valid bit and can neither match a PAM position nor sit in a target.
// Pass 1  count  : one thread per 32-position word evaluates BOTH strands for all 32 positions at
//                  once with bitwise logic: the PAM matches at p iff, for every PAM position j, the
//                  base at p+j is valid and belongs to the IUPAC set of letter j (a 4-entry truth
//                  table on the two planes); the target window is accepted iff its L positions are
//                  valid.  popc of the two hit words -> per-block counts.
// Pass 2  offsets: exclusive scan of the block counts (forward rows first, then reverse rows).
// Pass 3  emit   : recomputes the hit words, block-scans the per-thread counts and writes the
//                  records in ascending position order -- the reference's row order
//                  (core.py:254-284) -- extracting the L-bit windows with funnel shifts;
//                  reverse-strand guides are reverse-complemented with brev + bitwise not.
//
// Algorithmic HBM bytes: n (ASCII in) + 3n/8 (planes out) + 2 * 3n/8 (planes in, twice) + 14 B per hit.
#include "scan.cuh"
#include <new>

namespace gm {

static constexpr int SCAN_THREADS = 256;

struct PamParams {
    int P, L, five_prime;
    int fmask[GM_MAX_PAM];     // IUPAC set of PAM letter j, bit0=A bit1=C bit2=G bit3=T
    int rmask[GM_MAX_PAM];     // sets of revcomp(PAM) as searched on the forward text (core.py:263,279)
    int off_f, off_r;          // target window start relative to the match start, per strand
};

static int iupac_set(char c) {      // core.py:118-121 / :1103-1120
    switch (c) {
    case 'A': return 1; case 'C': return 2; case 'G': return 4; case 'T': return 8;
    case 'M': return 3; case 'R': return 5; case 'W': return 9; case 'S': return 6;
    case 'Y': return 10; case 'K': return 12; case 'V': return 7; case 'H': return 11;
    case 'D': return 13; case 'B': return 14; case 'X': return 15; case 'N': return 15;
    default: return 0;
    }
}
static int complement_set(int m) { return ((m & 1) << 3) | ((m & 2) << 1) | ((m & 4) >> 1) | ((m & 8) >> 3); }

// ---- pass 0 --------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(SCAN_THREADS) encode_kernel(const uint8_t *__restrict__ seq, int64_t n_words,
                                                              uint32_t *__restrict__ lo, uint32_t *__restrict__ hi,
                                                              uint32_t *__restrict__ valid) {
    const int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= n_words) return;
    const uint4 *p = reinterpret_cast<const uint4 *>(seq + w * 32);   // buffer is zero padded to 32 bytes
    const uint4 a = p[0], b = p[1];
    const uint32_t v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
    uint32_t l = 0, h = 0, ok = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) {
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const uint32_t c = (v[i] >> (8 * j)) & 0xFFu;
            const uint32_t good = (c == 'A') | (c == 'C') | (c == 'G') | (c == 'T');
            const uint32_t x = (c >> 1) & 3u;          // A:0 C:1 G:3 T:2
            const uint32_t code = x ^ (x >> 1);        // A:0 C:1 G:2 T:3
            const int bit = 4 * i + j;
            l |= (code & 1u & good) << bit;
            h |= ((code >> 1) & good) << bit;
            ok |= good << bit;
        }
    }
    lo[w + 1] = l;      // word 0 and the words past the end stay zero (= invalid)
    hi[w + 1] = h;
    valid[w + 1] = ok;
}

// ---- window helpers -------------------------------------------------------------------------------------
// x0..x3 hold positions [32(w-1), 32(w+3)).  bits_at(off) = the 32 positions starting at 32w + off,
// off in [-32, 64].
struct Win {
    uint32_t x0, x1, x2, x3;
    __device__ __forceinline__ uint32_t at(int off) const {
        const int a = off + 32, i = a >> 5, sh = a & 31;
        const uint32_t l = i == 0 ? x0 : i == 1 ? x1 : i == 2 ? x2 : x3;
        const uint32_t h = i == 0 ? x1 : i == 1 ? x2 : i == 2 ? x3 : 0u;
        return __funnelshift_r(l, h, sh);
    }
};

__device__ __forceinline__ Win load_win(const uint32_t *__restrict__ plane, int64_t w) {
    Win r;                        // plane[] is stored with one leading zero word
    r.x0 = plane[w];
    r.x1 = plane[w + 1];
    r.x2 = plane[w + 2];
    r.x3 = plane[w + 3];
    return r;
}

__device__ __forceinline__ uint32_t in_set(int set, uint32_t l, uint32_t h) {
    uint32_t r = 0;
    if (set & 1) r |= ~h & ~l;
    if (set & 2) r |= ~h & l;
    if (set & 4) r |= h & ~l;
    if (set & 8) r |= h & l;
    return r;
}

// hit words of the 32 positions of word w: bit b set <=> a row is emitted for match start 32w+b
__device__ __forceinline__ void hit_words(const PamParams &pp, const Win &lo, const Win &hi, const Win &va,
                                          uint32_t &hit_f, uint32_t &hit_r) {
    uint32_t mf = 0xFFFFFFFFu, mr = 0xFFFFFFFFu;
    for (int j = 0; j < pp.P; j++) {
        const uint32_t l = lo.at(j), h = hi.at(j), v = va.at(j);
        mf &= v & in_set(pp.fmask[j], l, h);
        mr &= v & in_set(pp.rmask[j], l, h);
    }
    uint32_t vf = 0xFFFFFFFFu, vr = 0xFFFFFFFFu;
    for (int i = 0; i < pp.L; i++) {                 // all L positions of the target window valid
        vf &= va.at(pp.off_f + i);
        vr &= va.at(pp.off_r + i);
    }
    hit_f = mf & vf;
    hit_r = mr & vr;
}

// ---- pass 1 --------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(SCAN_THREADS) scan_count_kernel(const PamParams pp, int64_t n_words,
                                                                  const uint32_t *__restrict__ lo, const uint32_t *__restrict__ hi,
                                                                  const uint32_t *__restrict__ valid,
                                                                  uint32_t *__restrict__ blk_f, uint32_t *__restrict__ blk_r) {
    __shared__ uint32_t s_f[SCAN_THREADS / 32], s_r[SCAN_THREADS / 32];
    const int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t cf = 0, cr = 0;
    if (w < n_words) {
        uint32_t hf, hr;
        hit_words(pp, load_win(lo, w), load_win(hi, w), load_win(valid, w), hf, hr);
        cf = __popc(hf);
        cr = __popc(hr);
    }
    cf = __reduce_add_sync(0xFFFFFFFFu, cf);
    cr = __reduce_add_sync(0xFFFFFFFFu, cr);
    if ((threadIdx.x & 31) == 0) { s_f[threadIdx.x >> 5] = cf; s_r[threadIdx.x >> 5] = cr; }
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t tf = 0, tr = 0;
        for (int i = 0; i < SCAN_THREADS / 32; i++) { tf += s_f[i]; tr += s_r[i]; }
        blk_f[blockIdx.x] = tf;
        blk_r[blockIdx.x] = tr;
    }
}

// ---- pass 2: exclusive scan of block counts (single CTA; #blocks = n/8192) ---------------------------------
__global__ void __launch_bounds__(1024) block_offsets_kernel(int64_t n_blocks, const uint32_t *__restrict__ blk_f,
                                                             const uint32_t *__restrict__ blk_r, uint64_t *__restrict__ off_f,
                                                             uint64_t *__restrict__ off_r, uint64_t *__restrict__ totals) {
    __shared__ uint64_t s_w[2][32];
    __shared__ uint64_t s_carry[2];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) { s_carry[0] = 0; s_carry[1] = 0; }
    __syncthreads();
    for (int64_t base = 0; base < n_blocks; base += 1024) {
        const int64_t i = base + threadIdx.x;
        uint64_t vf = i < n_blocks ? blk_f[i] : 0, vr = i < n_blocks ? blk_r[i] : 0;
        uint64_t sf = vf, sr = vr;
        for (int d = 1; d < 32; d <<= 1) {
            const uint64_t tf = __shfl_up_sync(0xFFFFFFFFu, sf, d), tr = __shfl_up_sync(0xFFFFFFFFu, sr, d);
            if (lane >= d) { sf += tf; sr += tr; }
        }
        if (lane == 31) { s_w[0][warp] = sf; s_w[1][warp] = sr; }
        __syncthreads();
        if (warp == 0) {
            uint64_t wf = s_w[0][lane], wr = s_w[1][lane];
            for (int d = 1; d < 32; d <<= 1) {
                const uint64_t tf = __shfl_up_sync(0xFFFFFFFFu, wf, d), tr = __shfl_up_sync(0xFFFFFFFFu, wr, d);
                if (lane >= d) { wf += tf; wr += tr; }
            }
            s_w[0][lane] = wf;
            s_w[1][lane] = wr;
        }
        __syncthreads();
        const uint64_t pf = s_carry[0] + (warp ? s_w[0][warp - 1] : 0) + sf - vf;
        const uint64_t pr = s_carry[1] + (warp ? s_w[1][warp - 1] : 0) + sr - vr;
        if (i < n_blocks) { off_f[i] = pf; off_r[i] = pr; }
        __syncthreads();
        if (threadIdx.x == 1023) { s_carry[0] += s_w[0][31]; s_carry[1] += s_w[1][31]; }
        __syncthreads();
    }
    if (threadIdx.x == 0) { totals[0] = s_carry[0]; totals[1] = s_carry[1]; }
}

// ---- pass 3 --------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t bit_reverse_low(uint32_t x, int nbits) { return __brev(x) >> (32 - nbits); }

__global__ void __launch_bounds__(SCAN_THREADS) scan_emit_kernel(const PamParams pp, int64_t n_words,
                                                                 const uint32_t *__restrict__ lo, const uint32_t *__restrict__ hi,
                                                                 const uint32_t *__restrict__ valid,
                                                                 const uint64_t *__restrict__ off_f, const uint64_t *__restrict__ off_r,
                                                                 uint64_t n_fwd_total, uint64_t *__restrict__ guides,
                                                                 uint32_t *__restrict__ start, uint16_t *__restrict__ pamcode) {
    __shared__ uint32_t s_f[SCAN_THREADS / 32], s_r[SCAN_THREADS / 32];
    const int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t hf = 0, hr = 0;
    Win wl, wh;
    wl.x0 = wl.x1 = wl.x2 = wl.x3 = 0;
    wh = wl;
    if (w < n_words) {
        wl = load_win(lo, w);
        wh = load_win(hi, w);
        hit_words(pp, wl, wh, load_win(valid, w), hf, hr);
    }
    const uint32_t cf = __popc(hf), cr = __popc(hr);
    uint32_t sf = cf, sr = cr;                       // inclusive warp scans
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t tf = __shfl_up_sync(0xFFFFFFFFu, sf, d), tr = __shfl_up_sync(0xFFFFFFFFu, sr, d);
        if (lane >= d) { sf += tf; sr += tr; }
    }
    if (lane == 31) { s_f[warp] = sf; s_r[warp] = sr; }
    __syncthreads();
    uint32_t wf = 0, wr = 0;
    for (int i = 0; i < warp; i++) { wf += s_f[i]; wr += s_r[i]; }
    uint64_t out_f = off_f[blockIdx.x] + wf + sf - cf;
    uint64_t out_r = n_fwd_total + off_r[blockIdx.x] + wr + sr - cr;

    const uint32_t lmask = (1u << pp.L) - 1u, pmask = (1u << pp.P) - 1u;
    const int64_t pos0 = w * 32;
    while (hf) {                                     // forward-strand rows, ascending position
        const int b = __ffs(hf) - 1;
        hf &= hf - 1;
        const uint32_t gl = wl.at(b + pp.off_f) & lmask, gh = wh.at(b + pp.off_f) & lmask;
        const uint32_t pl = wl.at(b) & pmask, ph = wh.at(b) & pmask;
        guides[out_f] = from_planes(gl, gh);
        start[out_f] = (uint32_t)(pos0 + b + pp.off_f);
        pamcode[out_f] = (uint16_t)from_planes(pl, ph);
        out_f++;
    }
    while (hr) {                                     // reverse-strand rows: reverse complement
        const int b = __ffs(hr) - 1;
        hr &= hr - 1;
        const uint32_t gl = bit_reverse_low(~wl.at(b + pp.off_r) & lmask, pp.L);
        const uint32_t gh = bit_reverse_low(~wh.at(b + pp.off_r) & lmask, pp.L);
        const uint32_t pl = bit_reverse_low(~wl.at(b) & pmask, pp.P);
        const uint32_t ph = bit_reverse_low(~wh.at(b) & pmask, pp.P);
        guides[out_r] = from_planes(gl, gh);
        start[out_r] = (uint32_t)(pos0 + b + pp.off_r);
        pamcode[out_r] = (uint16_t)from_planes(pl, ph);
        out_r++;
    }
}


// ---- record-ordered emit (sessions) -----------------------------------------------------------------------------------
// The reference lists, PER RECORD, all forward hits and then all reverse hits (core.py:254-284).  With the global
// forward / reverse ranks of a hit (gf, gr: its position among all forward / reverse hits of the joined genome) and the
// ranks F[r], R[r] of the first position of record r, the row of a hit of record r is
//      forward:  F[r] + R[r] + (gf - F[r])                      = R[r] + gf
//      reverse:  F[r] + R[r] + (F[r+1] - F[r]) + (gr - R[r])    = F[r+1] + gr
// so the scan writes every row straight to its final place -- no sort, no regrouping on the host.

// ranks at the record boundaries: one CTA per boundary r (1 <= r < n_rec) counts the hits of its block before rec_start[r]
__global__ void __launch_bounds__(SCAN_THREADS) boundary_rank_kernel(const PamParams pp, int64_t n_words, const uint32_t *__restrict__ lo,
                                                                     const uint32_t *__restrict__ hi, const uint32_t *__restrict__ valid,
                                                                     const uint64_t *__restrict__ off_f, const uint64_t *__restrict__ off_r,
                                                                     const uint64_t *__restrict__ totals, const int64_t *__restrict__ rec_start,
                                                                     int n_rec, uint64_t *__restrict__ F, uint64_t *__restrict__ R) {
    __shared__ uint32_t s_f[SCAN_THREADS / 32], s_r[SCAN_THREADS / 32];
    const int r = blockIdx.x;                              // 0 .. n_rec
    if (r == 0) { if (threadIdx.x == 0) { F[0] = 0; R[0] = 0; } return; }
    const int64_t b = rec_start[r];
    const int64_t wb = b >> 5;
    if (r == n_rec || wb >= n_words) { if (threadIdx.x == 0) { F[r] = totals[0]; R[r] = totals[1]; } return; }
    const int64_t blk = wb / SCAN_THREADS;
    const int64_t w = blk * SCAN_THREADS + threadIdx.x;
    uint32_t cf = 0, cr = 0;
    if (w <= wb) {
        uint32_t hf, hr;
        hit_words(pp, load_win(lo, w), load_win(hi, w), load_win(valid, w), hf, hr);
        if (w == wb) { const uint32_t m = (1u << (b & 31)) - 1u; hf &= m; hr &= m; }
        cf = __popc(hf);
        cr = __popc(hr);
    }
    cf = __reduce_add_sync(0xFFFFFFFFu, cf);
    cr = __reduce_add_sync(0xFFFFFFFFu, cr);
    if ((threadIdx.x & 31) == 0) { s_f[threadIdx.x >> 5] = cf; s_r[threadIdx.x >> 5] = cr; }
    __syncthreads();
    if (threadIdx.x == 0) {
        uint64_t tf = 0, tr = 0;
        for (int i = 0; i < SCAN_THREADS / 32; i++) { tf += s_f[i]; tr += s_r[i]; }
        F[r] = off_f[blk] + tf;
        R[r] = off_r[blk] + tr;
    }
}

__device__ __forceinline__ int record_of(const int64_t *__restrict__ rec_start, int n_rec, int64_t pos) {
    int lo = 0, hi = n_rec;                                // largest r with rec_start[r] <= pos
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (rec_start[mid] <= pos) lo = mid; else hi = mid;
    }
    return lo;
}

__global__ void __launch_bounds__(SCAN_THREADS) scan_emit_records_kernel(const PamParams pp, int64_t n_words,
                                                                         const uint32_t *__restrict__ lo, const uint32_t *__restrict__ hi,
                                                                         const uint32_t *__restrict__ valid,
                                                                         const uint64_t *__restrict__ off_f, const uint64_t *__restrict__ off_r,
                                                                         const int64_t *__restrict__ rec_start, int n_rec,
                                                                         const uint64_t *__restrict__ F, const uint64_t *__restrict__ R,
                                                                         uint64_t *__restrict__ guides, uint32_t *__restrict__ start,
                                                                         uint16_t *__restrict__ pamcode, int32_t *__restrict__ rec,
                                                                         uint8_t *__restrict__ strand) {
    __shared__ uint32_t s_f[SCAN_THREADS / 32], s_r[SCAN_THREADS / 32];
    const int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t hf = 0, hr = 0;
    Win wl, wh;
    wl.x0 = wl.x1 = wl.x2 = wl.x3 = 0;
    wh = wl;
    if (w < n_words) {
        wl = load_win(lo, w);
        wh = load_win(hi, w);
        hit_words(pp, wl, wh, load_win(valid, w), hf, hr);
    }
    const uint32_t cf = __popc(hf), cr = __popc(hr);
    uint32_t sf = cf, sr = cr;                       // inclusive warp scans
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t tf = __shfl_up_sync(0xFFFFFFFFu, sf, d), tr = __shfl_up_sync(0xFFFFFFFFu, sr, d);
        if (lane >= d) { sf += tf; sr += tr; }
    }
    if (lane == 31) { s_f[warp] = sf; s_r[warp] = sr; }
    __syncthreads();
    if ((hf | hr) == 0u) return;
    uint32_t wf = 0, wr = 0;
    for (int i = 0; i < warp; i++) { wf += s_f[i]; wr += s_r[i]; }
    uint64_t gf = off_f[blockIdx.x] + wf + sf - cf;   // global forward / reverse rank of this word's first hit
    uint64_t gr = off_r[blockIdx.x] + wr + sr - cr;

    const uint32_t lmask = (1u << pp.L) - 1u, pmask = (1u << pp.P) - 1u;
    const int64_t pos0 = w * 32;
    const int r0 = record_of(rec_start, n_rec, pos0);
    int r = r0;
    while (hf) {                                     // forward-strand rows, ascending position
        const int b = __ffs(hf) - 1;
        hf &= hf - 1;
        while (pos0 + b >= rec_start[r + 1]) r++;
        const uint64_t o = R[r] + gf;
        const uint32_t gl = wl.at(b + pp.off_f) & lmask, gh = wh.at(b + pp.off_f) & lmask;
        const uint32_t pl = wl.at(b) & pmask, ph = wh.at(b) & pmask;
        guides[o] = from_planes(gl, gh);
        start[o] = (uint32_t)(pos0 + b + pp.off_f - rec_start[r]);
        pamcode[o] = (uint16_t)from_planes(pl, ph);
        rec[o] = r;
        strand[o] = 1;
        gf++;
    }
    r = r0;
    while (hr) {                                     // reverse-strand rows: reverse complement
        const int b = __ffs(hr) - 1;
        hr &= hr - 1;
        while (pos0 + b >= rec_start[r + 1]) r++;
        const uint64_t o = F[r + 1] + gr;
        const uint32_t gl = bit_reverse_low(~wl.at(b + pp.off_r) & lmask, pp.L);
        const uint32_t gh = bit_reverse_low(~wh.at(b + pp.off_r) & lmask, pp.L);
        const uint32_t pl = bit_reverse_low(~wl.at(b) & pmask, pp.P);
        const uint32_t ph = bit_reverse_low(~wh.at(b) & pmask, pp.P);
        guides[o] = from_planes(gl, gh);
        start[o] = (uint32_t)(pos0 + b + pp.off_r - rec_start[r]);
        pamcode[o] = (uint16_t)from_planes(pl, ph);
        rec[o] = r;
        strand[o] = 0;
        gr++;
    }
}

// ---- text columns of the frame, produced from the resident rows and genome ---------------------------------------------
// `target` (core.py:155,183,209,236): the guide as ASCII, L bytes per row.
__global__ void __launch_bounds__(256) decode_rows_kernel(const uint64_t *__restrict__ guides, int64_t n_rows, int L, uint8_t *__restrict__ out) {
    // four output bytes per thread, one 32-bit store (the text matrix is n_rows x L bytes, row-major, 4-byte aligned)
    const int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t total = n_rows * L;
    if (w * 4 >= total) return;
    int64_t i = (w * 4) / L;
    int j = (int)(w * 4 - i * L);
    uint64_t g = guides[i];
    uint32_t word = 0;
#pragma unroll
    for (int b = 0; b < 4; b++) {
        if (w * 4 + b < total) word |= (uint32_t)(uint8_t)("ACGT"[(g >> (2 * j)) & 3u]) << (8 * b);
        if (++j == L) { j = 0; i++; if (i < n_rows) g = guides[i]; }
    }
    if (w * 4 + 4 <= total) *reinterpret_cast<uint32_t *>(out + w * 4) = word;
    else for (int b = 0; w * 4 + b < total; b++) out[w * 4 + b] = (uint8_t)(word >> (8 * b));
}

__constant__ uint8_t c_comp[256];

// `target_seq30` (core.py:156,184,210-211,237): the 30-nt slice of the record around the match -- [ms-3, ms+27) for 5prime
// forward / 3prime reverse hits, [me-27, me+3) for the others -- reverse-complemented for reverse hits, NOT validated.
// Rows whose window leaves the record are flagged in `edge` and filled with '?': the host applies Python's literal slice
// semantics to those few rows.  Four output bytes per thread, one 32-bit store.
__device__ __forceinline__ uint8_t context_byte(const uint8_t *__restrict__ seq, const int64_t *__restrict__ rec_start,
                                                const uint32_t *__restrict__ start, const int32_t *__restrict__ rec,
                                                const uint8_t *__restrict__ strand, int64_t i, int j, int P, int L, int five_prime, int width,
                                                bool *interior_out) {
    const int r = rec[i];
    const int fwd = strand[i];
    const int64_t st = start[i];
    const int64_t ms = five_prime ? (fwd ? st - P : st + L) : (fwd ? st + L : st - P);
    const int64_t a = (fwd == five_prime) ? ms - 3 : ms + P - (width - 3);
    const int64_t len = rec_start[r + 1] - rec_start[r] - 1;
    const bool interior = a >= 0 && a + width <= len;
    *interior_out = interior;
    if (!interior) return (uint8_t)'?';
    const int64_t g0 = rec_start[r] + a;
    return fwd ? seq[g0 + j] : c_comp[seq[g0 + width - 1 - j]];
}

__global__ void __launch_bounds__(256) context_rows_kernel(const uint8_t *__restrict__ seq, const int64_t *__restrict__ rec_start,
                                                           const uint32_t *__restrict__ start, const int32_t *__restrict__ rec,
                                                           const uint8_t *__restrict__ strand, int64_t n_rows, int P, int L,
                                                           int five_prime, int width, uint8_t *__restrict__ out, uint8_t *__restrict__ edge) {
    const int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t total = n_rows * width;
    if (w * 4 >= total) return;
    int64_t i = (w * 4) / width;
    int j = (int)(w * 4 - i * width);
    uint32_t word = 0;
#pragma unroll
    for (int b = 0; b < 4; b++) {
        if (w * 4 + b < total) {
            bool interior;
            word |= (uint32_t)context_byte(seq, rec_start, start, rec, strand, i, j, P, L, five_prime, width, &interior) << (8 * b);
            if (j == 0 && edge) edge[i] = interior ? 0 : 1;
        }
        if (++j == width) { j = 0; i++; }
    }
    if (w * 4 + 4 <= total) *reinterpret_cast<uint32_t *>(out + w * 4) = word;
    else for (int b = 0; w * 4 + b < total; b++) out[w * 4 + b] = (uint8_t)(word >> (8 * b));
}

static int ensure_comp_table() {
    static bool table_set = false;
    if (!table_set) {
        uint8_t comp[256];
        for (int c = 0; c < 256; c++) comp[c] = (uint8_t)c;
        const char *from = "ACGTMRWSYKVHDBXNacgtmrwsykvhdbxn", *to = "TGCAKYWSRMBDHVXNtgcakywsrmbdhvxn";   // Bio.Seq complement
        for (int k = 0; from[k]; k++) comp[(uint8_t)from[k]] = (uint8_t)to[k];
        GM_CUDA(cudaMemcpyToSymbol(c_comp, comp, 256));
        table_set = true;
    }
    return GM_OK;
}

static int fill_params(PamParams &pp, const char *pam, int pam_len, int five_prime, int L) {
    memset(&pp, 0, sizeof pp);
    pp.P = pam_len;
    pp.L = L;
    pp.five_prime = five_prime ? 1 : 0;
    for (int j = 0; j < pam_len; j++) {
        pp.fmask[j] = iupac_set(pam[j]);
        GM_ARG(pp.fmask[j] != 0, "gm_scan_create: '%c' is not an IUPAC letter", pam[j]);
    }
    for (int j = 0; j < pam_len; j++) pp.rmask[j] = complement_set(pp.fmask[pam_len - 1 - j]);
    // target window relative to the match start ms (me = ms + P): core.py:155,183,209,236
    pp.off_f = five_prime ? pam_len : -L;
    pp.off_r = five_prime ? -L : pam_len;
    return GM_OK;
}

// Everything of a scan: upload, encode, count, offsets, emit.  Plain scans (rec_start == nullptr) list all forward rows
// and then all reverse rows with genome-wide coordinates; sessions list rows per record and keep the genome resident.
static int scan_build(Scan *s, const uint8_t *seq_ascii, int64_t n, const PamParams &pp, const int64_t *rec_start, int n_rec) {
    const int64_t n_words = (n + 31) / 32;
    const int64_t n_blocks = (n_words + SCAN_THREADS - 1) / SCAN_THREADS;
    const size_t plane_words = (size_t)n_words + 4;
    const bool session = rec_start != nullptr;
    uint8_t *d_seq = nullptr;
    uint32_t *d_planes = nullptr, *d_blk = nullptr;
    uint64_t *d_off = nullptr, *d_FR = nullptr;
    uint64_t totals[2] = {0, 0};
    const double t0 = now_ms();
    cudaError_t e = dev_alloc((void **)&d_seq, (size_t)n_words * 32, 0);
    if (e == cudaSuccess) e = dev_alloc((void **)&d_planes, 3 * plane_words * 4, 0);
    if (e == cudaSuccess) e = dev_alloc((void **)&d_blk, (size_t)n_blocks * 2 * 4, 0);
    if (e == cudaSuccess) e = dev_alloc((void **)&d_off, ((size_t)n_blocks * 2 + 2) * 8, 0);
    if (e == cudaSuccess) e = cudaMemsetAsync(d_seq + (n_words - 1) * 32, 0, 32, 0);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_seq, seq_ascii, (size_t)n, cudaMemcpyHostToDevice, 0);
    if (e == cudaSuccess) e = cudaMemsetAsync(d_planes, 0, 3 * plane_words * 4, 0);
    if (e == cudaSuccess && session) {
        e = dev_alloc((void **)&s->rec_start, (size_t)(n_rec + 1) * 8, 0);
        if (e == cudaSuccess) e = cudaMemcpyAsync(s->rec_start, rec_start, (size_t)(n_rec + 1) * 8, cudaMemcpyHostToDevice, 0);
        if (e == cudaSuccess) e = dev_alloc((void **)&d_FR, (size_t)(n_rec + 1) * 2 * 8, 0);
    }
    uint32_t *lo = d_planes, *hi = d_planes + plane_words, *va = d_planes + 2 * plane_words;
    uint32_t *blk_f = d_blk, *blk_r = d_blk + n_blocks;
    uint64_t *off_f = d_off, *off_r = d_off + n_blocks, *d_tot = d_off + 2 * n_blocks;
    if (e == cudaSuccess) {
        encode_kernel<<<(unsigned)n_blocks, SCAN_THREADS>>>(d_seq, n_words, lo, hi, va);
        scan_count_kernel<<<(unsigned)n_blocks, SCAN_THREADS>>>(pp, n_words, lo, hi, va, blk_f, blk_r);
        block_offsets_kernel<<<1, 1024>>>(n_blocks, blk_f, blk_r, off_f, off_r, d_tot);
        count_launch(3);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpy(totals, d_tot, sizeof totals, cudaMemcpyDeviceToHost);
    trace("scan: H2D + encode + count", t0);
    const int64_t nf = (int64_t)totals[0], nr = (int64_t)totals[1], nt = nf + nr;
    if (e == cudaSuccess && nt > 0) {
        e = dev_alloc((void **)&s->guides, (size_t)nt * 8, 0);
        if (e == cudaSuccess) e = dev_alloc((void **)&s->start, (size_t)nt * 4, 0);
        if (e == cudaSuccess) e = dev_alloc((void **)&s->pamcode, (size_t)nt * 2, 0);
        if (e == cudaSuccess && session) e = dev_alloc((void **)&s->rec, (size_t)nt * 4, 0);
        if (e == cudaSuccess && session) e = dev_alloc((void **)&s->strand, (size_t)nt, 0);
        if (e == cudaSuccess) {
            if (session) {
                uint64_t *F = d_FR, *R = d_FR + (n_rec + 1);
                boundary_rank_kernel<<<(unsigned)(n_rec + 1), SCAN_THREADS>>>(pp, n_words, lo, hi, va, off_f, off_r, d_tot, s->rec_start, n_rec, F, R);
                scan_emit_records_kernel<<<(unsigned)n_blocks, SCAN_THREADS>>>(pp, n_words, lo, hi, va, off_f, off_r, s->rec_start, n_rec, F, R,
                                                                               s->guides, s->start, s->pamcode, s->rec, s->strand);
                count_launch(2);
            } else {
                scan_emit_kernel<<<(unsigned)n_blocks, SCAN_THREADS>>>(pp, n_words, lo, hi, va, off_f, off_r, (uint64_t)nf,
                                                                       s->guides, s->start, s->pamcode);
                count_launch();
            }
            e = cudaGetLastError();
        }
        if (e == cudaSuccess) e = cudaDeviceSynchronize();
    }
    if (session && e == cudaSuccess) { s->seq = d_seq; s->n_seq = n; d_seq = nullptr; }
    dev_free(d_seq, 0);
    dev_free(d_planes, 0);
    dev_free(d_blk, 0);
    dev_free(d_off, 0);
    dev_free(d_FR, 0);
    if (e != cudaSuccess) return cuda_fail(e, "gm_scan_create", __FILE__, __LINE__);
    s->n_fwd = nf;
    s->n_rev = nr;
    s->session = session;
    s->n_rec = n_rec;
    s->P = pp.P;
    s->L = pp.L;
    s->five_prime = pp.five_prime;
    return GM_OK;
}

}  // namespace gm

using namespace gm;

extern "C" int gm_scan_free(void *scan) {
    Scan *s = (Scan *)scan;
    if (!s) return GM_OK;
    dev_free(s->guides, 0);
    dev_free(s->start, 0);
    dev_free(s->pamcode, 0);
    dev_free(s->rec, 0);
    dev_free(s->strand, 0);
    dev_free(s->seq, 0);
    dev_free(s->rec_start, 0);
    dev_free(s->first32, 0);
    dev_free(s->nb_codes, 0);
    dev_free(s->nb_idx, 0);
    dev_free(s->nb_dist, 0);
    delete s;
    return GM_OK;
}

static int scan_create_common(const uint8_t *seq_ascii, int64_t n, const int64_t *rec_start, int n_rec, const char *pam, int pam_len,
                              int five_prime, int L, void **scan) {
    GM_ARG(n >= 0 && (n == 0 || seq_ascii), "gm_scan_create: bad sequence buffer");
    GM_ARG(pam && pam_len >= 1 && pam_len <= GM_MAX_PAM, "gm_scan_create: PAM length %d outside [1,%d]", pam_len, GM_MAX_PAM);
    GM_ARG(L >= 1 && L <= GM_MAX_L, "gm_scan_create: L=%d outside [1,%d]", L, GM_MAX_L);
    if (n > 0xFFFFFF00LL) { set_error("gm_scan_create: %lld bases exceed the uint32 coordinate range", (long long)n); return GM_ERR_RANGE; }
    PamParams pp;
    int rc = fill_params(pp, pam, pam_len, five_prime, L);
    if (rc) return rc;
    if (rec_start) {
        GM_ARG(n_rec >= 1 && rec_start[0] == 0, "gm_session_create: record table must start at 0");
        for (int r = 0; r < n_rec; r++) GM_ARG(rec_start[r + 1] > rec_start[r], "gm_session_create: record offsets must increase");
        GM_ARG(rec_start[n_rec] == n + 1, "gm_session_create: last record offset must be n + 1 (records joined by one byte)");
    }
    Scan *s = new (std::nothrow) Scan();
    if (!s) { set_error("out of host memory"); return GM_ERR_NOMEM; }
    s->session = rec_start != nullptr;
    s->P = pam_len; s->L = L; s->five_prime = five_prime ? 1 : 0; s->n_rec = n_rec;
    *scan = s;
    if (n == 0) return GM_OK;
    rc = scan_build(s, seq_ascii, n, pp, rec_start, n_rec);
    if (rc) { gm_scan_free(s); *scan = nullptr; }
    return rc;
}

extern "C" int gm_scan_create(const uint8_t *seq_ascii, int64_t n, const char *pam, int pam_len, int five_prime, int L,
                              void **scan, int64_t *n_fwd, int64_t *n_rev) {
    int rc = ensure_init();
    if (rc) return rc;
    GM_ARG(scan && n_fwd && n_rev, "gm_scan_create: NULL output pointer");
    *scan = nullptr;
    *n_fwd = *n_rev = 0;
    rc = scan_create_common(seq_ascii, n, nullptr, 1, pam, pam_len, five_prime, L, scan);
    if (rc) return rc;
    *n_fwd = ((Scan *)*scan)->n_fwd;
    *n_rev = ((Scan *)*scan)->n_rev;
    return GM_OK;
}

extern "C" int gm_session_create(const uint8_t *seq_ascii, int64_t n, const int64_t *rec_start, int n_rec, const char *pam, int pam_len,
                                 int five_prime, int L, void **session, int64_t *n_rows) {
    int rc = ensure_init();
    if (rc) return rc;
    GM_ARG(session && n_rows && rec_start, "gm_session_create: NULL pointer");
    *session = nullptr;
    *n_rows = 0;
    rc = scan_create_common(seq_ascii, n, rec_start, n_rec, pam, pam_len, five_prime, L, session);
    if (rc) return rc;
    *n_rows = ((Scan *)*session)->n_fwd + ((Scan *)*session)->n_rev;
    return GM_OK;
}

extern "C" int gm_session_free(void *session) { return gm_scan_free(session); }

extern "C" int gm_scan_device_ptrs(void *scan, const uint64_t **d_guide2bit, const uint32_t **d_start, const uint16_t **d_pamcode, int64_t *n_rows) {
    Scan *s = (Scan *)scan;
    GM_ARG(s, "gm_scan_device_ptrs: NULL handle");
    if (d_guide2bit) *d_guide2bit = s->guides;
    if (d_start) *d_start = s->start;
    if (d_pamcode) *d_pamcode = s->pamcode;
    if (n_rows) *n_rows = s->n_fwd + s->n_rev;
    return GM_OK;
}

extern "C" int gm_session_fetch_rows(void *session, uint64_t *guide2bit, uint32_t *start, uint16_t *pamcode, int32_t *rec, uint8_t *strand) {
    Scan *s = (Scan *)session;
    GM_ARG(s && s->session, "gm_session_fetch_rows: not a session handle");
    const size_t nt = (size_t)(s->n_fwd + s->n_rev);
    if (nt == 0) return GM_OK;
    const double t0 = now_ms();
    prefault(guide2bit, nt * 8); prefault(start, nt * 4); prefault(pamcode, nt * 2); prefault(rec, nt * 4); prefault(strand, nt);
    if (guide2bit) GM_CUDA(cudaMemcpyAsync(guide2bit, s->guides, nt * 8, cudaMemcpyDeviceToHost, 0));
    if (start) GM_CUDA(cudaMemcpyAsync(start, s->start, nt * 4, cudaMemcpyDeviceToHost, 0));
    if (pamcode) GM_CUDA(cudaMemcpyAsync(pamcode, s->pamcode, nt * 2, cudaMemcpyDeviceToHost, 0));
    if (rec) GM_CUDA(cudaMemcpyAsync(rec, s->rec, nt * 4, cudaMemcpyDeviceToHost, 0));
    if (strand) GM_CUDA(cudaMemcpyAsync(strand, s->strand, nt, cudaMemcpyDeviceToHost, 0));
    GM_CUDA(cudaStreamSynchronize(0));
    trace("session: fetch rows", t0);
    return GM_OK;
}

extern "C" int gm_session_fetch_text(void *session, uint8_t *target_ascii, uint8_t *context, int width, uint8_t *edge) {
    Scan *s = (Scan *)session;
    GM_ARG(s && s->session, "gm_session_fetch_text: not a session handle");
    GM_ARG(!context || (width >= 4 && width <= 1024), "gm_session_fetch_text: context width %d outside [4,1024]", width);
    const int64_t nt = s->n_fwd + s->n_rev;
    if (nt == 0) return GM_OK;
    const double t0 = now_ms();
    int rc = ensure_comp_table();
    if (rc) return rc;
    trace("  text: complement table", t0);
    uint8_t *d_t = nullptr, *d_c = nullptr, *d_e = nullptr;
    cudaError_t e = cudaSuccess;
    if (target_ascii) {
        e = dev_alloc((void **)&d_t, (size_t)nt * s->L, 0);
        if (e == cudaSuccess) {
            decode_rows_kernel<<<(unsigned)(((nt * s->L + 3) / 4 + 255) / 256), 256>>>(s->guides, nt, s->L, d_t);
            count_launch();
            if (trace_on()) { cudaStreamSynchronize(0); trace("  text: decode kernel", t0); }
            prefault(target_ascii, (size_t)nt * s->L);
            trace("  text: prefault target", t0);
            e = cudaMemcpyAsync(target_ascii, d_t, (size_t)nt * s->L, cudaMemcpyDeviceToHost, 0);
        }
    }
    if (e == cudaSuccess && context) {
        e = dev_alloc((void **)&d_c, (size_t)nt * width, 0);
        if (e == cudaSuccess && edge) e = dev_alloc((void **)&d_e, (size_t)nt, 0);
        if (e == cudaSuccess) {
            context_rows_kernel<<<(unsigned)(((nt * width + 3) / 4 + 255) / 256), 256>>>(s->seq, s->rec_start, s->start, s->rec, s->strand, nt, s->P, s->L,
                                                                               s->five_prime, width, d_c, d_e);
            count_launch();
            if (trace_on()) { cudaStreamSynchronize(0); trace("  text: + D2H target, context kernel", t0); }
            prefault(context, (size_t)nt * width);
            prefault(edge, (size_t)nt);
            trace("  text: prefault context", t0);
            e = cudaMemcpyAsync(context, d_c, (size_t)nt * width, cudaMemcpyDeviceToHost, 0);
            if (e == cudaSuccess && edge) e = cudaMemcpyAsync(edge, d_e, (size_t)nt, cudaMemcpyDeviceToHost, 0);
        }
    }
    if (e == cudaSuccess) e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaStreamSynchronize(0);
    dev_free(d_t, 0); dev_free(d_c, 0); dev_free(d_e, 0);
    trace("session: text columns", t0);
    if (e != cudaSuccess) return cuda_fail(e, "gm_session_fetch_text", __FILE__, __LINE__);
    return GM_OK;
}

// ---- sequence context windows (the target_seq30 column) ---------------------------------------------------------
// find_targets also reports a 30-nt slice of the record around every match (core.py:156,184,210-211,237), reverse-
// complemented for reverse-strand hits and NOT validated (it may contain N or lower case).  Row i copies `width` bytes
// starting at win_start[i]; rows flagged in `revcomp` are reversed and complemented with Bio.Seq's IUPAC table
// (other bytes unchanged).  A window that does not lie inside [0, n) is filled with '?' -- the host applies the
// reference's literal Python slice semantics to those few rows near record ends.
__global__ void __launch_bounds__(256) gather_windows_kernel(const uint8_t *__restrict__ seq, int64_t n, const int64_t *__restrict__ win_start,
                                                             const uint8_t *__restrict__ revcomp, int64_t n_rows, int width,
                                                             uint8_t *__restrict__ out) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_rows * width) return;
    const int64_t i = t / width;
    const int j = (int)(t - i * width);
    const int64_t a = win_start[i];
    uint8_t c = (uint8_t)'?';
    if (a >= 0 && a + width <= n) c = revcomp[i] ? c_comp[seq[a + width - 1 - j]] : seq[a + j];
    out[t] = c;
}

extern "C" int gm_gather_windows(const uint8_t *seq_ascii, int64_t n, const int64_t *win_start, const uint8_t *revcomp, int64_t n_rows,
                                 int width, uint8_t *out) {
    int rc = ensure_init();
    if (rc) return rc;
    GM_ARG(n >= 0 && n_rows >= 0 && width >= 1 && width <= 1024, "gm_gather_windows: bad size");
    if (n_rows == 0) return GM_OK;
    GM_ARG(win_start && revcomp && out && (n == 0 || seq_ascii), "gm_gather_windows: NULL buffer");
    rc = ensure_comp_table();
    if (rc) return rc;
    uint8_t *d_seq = nullptr, *d_rc = nullptr, *d_out = nullptr;
    int64_t *d_ws = nullptr;
    cudaError_t e = dev_alloc((void **)&d_seq, (size_t)n + 16, 0);
    if (e == cudaSuccess) e = dev_alloc((void **)&d_ws, (size_t)n_rows * 8, 0);
    if (e == cudaSuccess) e = dev_alloc((void **)&d_rc, (size_t)n_rows, 0);
    if (e == cudaSuccess) e = dev_alloc((void **)&d_out, (size_t)n_rows * width, 0);
    if (e == cudaSuccess && n) e = cudaMemcpyAsync(d_seq, seq_ascii, (size_t)n, cudaMemcpyHostToDevice, 0);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_ws, win_start, (size_t)n_rows * 8, cudaMemcpyHostToDevice, 0);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_rc, revcomp, (size_t)n_rows, cudaMemcpyHostToDevice, 0);
    if (e == cudaSuccess) {
        const int64_t total = n_rows * width;
        gather_windows_kernel<<<(unsigned)((total + 255) / 256), 256>>>(d_seq, n, d_ws, d_rc, n_rows, width, d_out);
        count_launch();
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(out, d_out, (size_t)n_rows * width, cudaMemcpyDeviceToHost, 0);
    if (e == cudaSuccess) e = cudaStreamSynchronize(0);
    dev_free(d_seq, 0); dev_free(d_ws, 0); dev_free(d_rc, 0); dev_free(d_out, 0);
    if (e != cudaSuccess) return cuda_fail(e, "gm_gather_windows", __FILE__, __LINE__);
    return GM_OK;
}

extern "C" int gm_scan_fetch(void *scan, uint64_t *guide2bit, uint32_t *start, uint16_t *pamcode) {
    Scan *s = (Scan *)scan;
    GM_ARG(s, "gm_scan_fetch: NULL handle");
    const size_t nt = (size_t)(s->n_fwd + s->n_rev);
    if (nt == 0) return GM_OK;
    if (guide2bit) GM_CUDA(cudaMemcpy(guide2bit, s->guides, nt * 8, cudaMemcpyDeviceToHost));
    if (start) GM_CUDA(cudaMemcpy(start, s->start, nt * 4, cudaMemcpyDeviceToHost));
    if (pamcode) GM_CUDA(cudaMemcpy(pamcode, s->pamcode, nt * 2, cudaMemcpyDeviceToHost));
    return GM_OK;
}
