/*
 * gm_oracle.c -- CPU ORACLE for the GuideMaker off-target hot path.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load it.  The product
 * (guidemaker_b200/) never imports, links or falls back to anything in oracle/.
 *
 * It restates, in plain C, the algorithm of the reference's hot path
 * (/root/reference/guidemaker/core.py, cited per function below).  The arithmetic
 * of the k-nearest-neighbour step lives in a third-party dependency that is NOT
 * vendored in the reference: nmslib==2.1.1 (requirements.txt:31, setup.py:13),
 * spaces `bit_hamming` (popcount of XOR over the one-hot 4L-bit vector == 2 x
 * mismatching positions) and `leven` (unit-cost Levenshtein).  Their published
 * definitions are restated here as EXACT brute force; the reference's HNSW index
 * is an approximation of exactly this search.
 *
 * Parity pin: tests/test_oracle_golden.py checks this file against (i) the six
 * known-answer vectors of the reference's own tests/test_core.py (:41-65, :95-102,
 * :116-126, :319-347) and (ii) fixtures in tests/golden/ produced by running the
 * reference's real core.py (tests/golden/make_golden.py).
 *
 * Packed guide format used across the whole project ("guide2bit"):
 *   base i (i = 0 is the 5'-most base of the guide) occupies bits [2i, 2i+1] of a
 *   uint64, code A=0 C=1 G=2 T=3; L <= 27 so at most 54 bits are used.
 *
 * Build: see oracle/Makefile (gcc -O3 -march=native -fopenmp -shared -fPIC).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define GMO_MAXK 64
#define GMO_MAXL 27

/* ---- alphabet ------------------------------------------------------------------ */

/* concrete base -> 2-bit code, anything else (N, lower case, ']' ...) -> 4 = invalid.
 * core.py:138 accepts only 'A','T','C','G'; core.py:118-121 regex classes match only
 * those four upper-case letters. */
static inline int base_code(uint8_t c) {
    switch (c) {
    case 'A': return 0;
    case 'C': return 1;
    case 'G': return 2;
    case 'T': return 3;
    default:  return 4;
    }
}

/* IUPAC letter -> 4-bit set over {A=1,C=2,G=4,T=8}; table of core.py:118-121
 * (identical content to extend_ambiguous_dna, core.py:1103-1120). 0 = not a PAM letter. */
static int iupac_mask(char c) {
    switch (c) {
    case 'A': return 1;  case 'C': return 2;  case 'G': return 4;  case 'T': return 8;
    case 'M': return 1|2;   case 'R': return 1|4;   case 'W': return 1|8;
    case 'S': return 2|4;   case 'Y': return 2|8;   case 'K': return 4|8;
    case 'V': return 1|2|4; case 'H': return 1|2|8; case 'D': return 1|4|8;
    case 'B': return 2|4|8; case 'X': return 15;    case 'N': return 15;
    default:  return 0;
    }
}

/* complement of a base set: A<->T, C<->G  (bit0<->bit3, bit1<->bit2) */
static inline int mask_complement(int m) {
    return ((m & 1) << 3) | ((m & 2) << 1) | ((m & 4) >> 1) | ((m & 8) >> 3);
}

/* ---- PAM scan -------------------------------------------------------------------- */

/*
 * gmo_pam_scan: restates PamTarget.find_targets for ONE record (core.py:142-246 hit
 * geometry, :127-140 check_target, :249-287 row order).
 *
 * Rows are written forward-strand hits first (ascending match start), then
 * reverse-strand hits (ascending match start) -- core.py:254-284 appends the forward
 * frame before the reverse frame.  For each hit:
 *   guides[i]  packed guide as it appears in the `target` column (reverse hits are
 *              reverse-complemented, core.py:208,235)
 *   start[i]   0-based forward-strand start of the target window (core.py:160,187,214,240)
 *   pamcode[i] the `exact_pam` string packed 2 bits/base, PAM base j at bits [2j,2j+1]
 *              (reverse hits reverse-complemented, core.py:213,239)
 * Returns 0, or -1 if cap is too small (n_fwd/n_rev still hold the true counts), or -2
 * for a bad PAM / length.  Output pointers may be NULL to count only.
 */
int gmo_pam_scan(const uint8_t *seq, int64_t n, const char *pam, int P, int five_prime, int L,
                 uint64_t *guides, uint32_t *start, uint16_t *pamcode, int64_t cap,
                 int64_t *n_fwd, int64_t *n_rev)
{
    int fmask[8], rmask[8];
    if (P < 1 || P > 8 || L < 1 || L > GMO_MAXL) return -2;
    for (int j = 0; j < P; j++) {
        fmask[j] = iupac_mask(pam[j]);
        if (!fmask[j]) return -2;
    }
    /* reverse strand searches revcomp(PAM) on the forward text (core.py:263,279) */
    for (int j = 0; j < P; j++) rmask[j] = mask_complement(fmask[P - 1 - j]);

    int64_t nf = 0, nr = 0, w = 0;
    int overflow = 0;
    for (int pass = 0; pass < 2; pass++) {          /* pass 0: forward, pass 1: reverse */
        const int *mask = pass == 0 ? fmask : rmask;
        for (int64_t ms = 0; ms + P <= n; ms++) {
            int ok = 1;
            for (int j = 0; j < P && ok; j++) {
                int c = base_code(seq[ms + j]);
                ok = c < 4 && ((mask[j] >> c) & 1);
            }
            if (!ok) continue;
            int64_t me = ms + P, ws;
            /* window of the target on the forward text */
            if (pass == 0) ws = five_prime ? me : ms - L;      /* core.py:155 / :183 */
            else           ws = five_prime ? ms - L : me;      /* core.py:209 / :236 */
            if (ws < 0 || ws + L > n) continue;                /* slice shorter than L -> check_target False */
            uint64_t g = 0;
            for (int i = 0; i < L && ok; i++) {
                int c = base_code(seq[ws + i]);
                if (c >= 4) { ok = 0; break; }
                if (pass == 0) g |= (uint64_t)c << (2 * i);
                else           g |= (uint64_t)(3 - c) << (2 * (L - 1 - i));   /* reverse complement */
            }
            if (!ok) continue;
            uint16_t pc = 0;
            for (int j = 0; j < P; j++) {
                int c = base_code(seq[ms + j]);
                if (pass == 0) pc |= (uint16_t)(c << (2 * j));
                else           pc |= (uint16_t)((3 - c) << (2 * (P - 1 - j)));
            }
            if (pass == 0) nf++; else nr++;
            if (guides) {
                if (w < cap) { guides[w] = g; start[w] = (uint32_t)ws; pamcode[w] = pc; }
                else overflow = 1;
            }
            w++;
        }
    }
    *n_fwd = nf; *n_rev = nr;
    return overflow ? -1 : 0;
}

/* ---- seed region + keep-first duplicate flag ---------------------------------------- */

/* seed key of a packed guide: first lsr bases (5prime) or last lsr bases (3prime), the whole
 * guide when lsr == 0 (core.py:402-412). */
static inline uint64_t seed_key(uint64_t g, int L, int lsr, int five_prime) {
    if (lsr == 0 || lsr >= L) return g;
    if (five_prime) return g & ((1ULL << (2 * lsr)) - 1);
    return g >> (2 * (L - lsr));
}

typedef struct { uint64_t key; int64_t row; } gmo_kr;
static int cmp_kr(const void *a, const void *b) {
    const gmo_kr *x = a, *y = b;
    if (x->key != y->key) return x->key < y->key ? -1 : 1;
    return x->row < y->row ? -1 : (x->row > y->row);
}

/* first_row[i] = smallest row index holding the same key as row i */
static int first_rows(const uint64_t *keys, int64_t n, int64_t *first_row) {
    gmo_kr *a = malloc((size_t)(n > 0 ? n : 1) * sizeof *a);
    if (!a) return -3;
    for (int64_t i = 0; i < n; i++) { a[i].key = keys[i]; a[i].row = i; }
    qsort(a, (size_t)n, sizeof *a, cmp_kr);
    int64_t head = 0;
    for (int64_t i = 0; i < n; i++) {
        if (i == 0 || a[i].key != a[i - 1].key) head = a[i].row;
        first_row[a[i].row] = head;
    }
    free(a);
    return 0;
}

/* gmo_seed_dedup: TargetProcessor.find_unique_near_pam, core.py:414-416 --
 * `seedseq.duplicated()` with pandas' default keep='first': a row is flagged iff an
 * EARLIER row carries the same seed. */
int gmo_seed_dedup(const uint64_t *guides, int64_t n, int L, int lsr, int five_prime, uint8_t *is_dup)
{
    uint64_t *keys = malloc((size_t)(n > 0 ? n : 1) * sizeof *keys);
    int64_t *fr = malloc((size_t)(n > 0 ? n : 1) * sizeof *fr);
    if (!keys || !fr) { free(keys); free(fr); return -3; }
    for (int64_t i = 0; i < n; i++) keys[i] = seed_key(guides[i], L, lsr, five_prime);
    int rc = first_rows(keys, n, fr);
    if (rc == 0) for (int64_t i = 0; i < n; i++) is_dup[i] = fr[i] != i;
    free(keys); free(fr);
    return rc;
}

/* gmo_first_occurrence: the index of create_index holds each distinct guide once
 * (core.py:446, `list(set(...))`, arbitrary order in the reference).  This project fixes the
 * order: distinct guides in order of FIRST OCCURRENCE in the targets frame.
 * first_row[i] = row of the first occurrence of guides[i]. */
int gmo_first_occurrence(const uint64_t *guides, int64_t n, int64_t *first_row)
{
    return first_rows(guides, n, first_row);
}

/* ---- distances ------------------------------------------------------------------------ */

/* number of mismatching positions between two packed guides == nmslib bit_hamming over
 * the one-hot encoding of core.py:379-386, divided by 2 (core.py:512-514). */
static inline int hamming2bit(uint64_t a, uint64_t b) {
    uint64_t x = a ^ b;
    x = (x | (x >> 1)) & 0x5555555555555555ULL;
    return __builtin_popcountll(x);
}

/* unit-cost Levenshtein (nmslib `leven`, core.py:461-466), textbook Wagner-Fischer DP on
 * the decoded bases; both strings have length L. */
static inline int leven2bit(uint64_t a, uint64_t b, int L) {
    int prev[GMO_MAXL + 1], cur[GMO_MAXL + 1];
    for (int j = 0; j <= L; j++) prev[j] = j;
    for (int i = 1; i <= L; i++) {
        int ai = (int)((a >> (2 * (i - 1))) & 3);
        cur[0] = i;
        for (int j = 1; j <= L; j++) {
            int bj = (int)((b >> (2 * (j - 1))) & 3);
            int sub = prev[j - 1] + (ai != bj);
            int del = prev[j] + 1, ins = cur[j - 1] + 1;
            int m = sub < del ? sub : del;
            cur[j] = m < ins ? m : ins;
        }
        memcpy(prev, cur, sizeof(int) * (size_t)(L + 1));
    }
    return prev[L];
}

/*
 * gmo_knn: exact k nearest targets of every query, ascending by (distance, target index)
 * -- the deterministic tie-break of BASELINE.json north_star.  Replaces
 * nmslib knnQueryBatch (core.py:502-503, :603).  metric 0 = hamming (mismatch count, NOT
 * doubled), 1 = leven.  Rows with fewer than k targets are padded with idx=-1, dist=255.
 * threads <= 0 -> all cores.
 */
int gmo_knn(const uint64_t *targets, int64_t n, const uint64_t *queries, int64_t q, int L, int metric,
            int k, int32_t *out_idx, uint8_t *out_dist, int threads)
{
    if (k < 1 || k > GMO_MAXK || L < 1 || L > GMO_MAXL) return -2;
#ifdef _OPENMP
    if (threads > 0) omp_set_num_threads(threads);
#else
    (void)threads;
#endif
#pragma omp parallel for schedule(dynamic, 16)
    for (int64_t qi = 0; qi < q; qi++) {
        int bd[GMO_MAXK]; int32_t bi[GMO_MAXK];
        int cnt = 0;
        uint64_t qq = queries[qi];
        for (int64_t t = 0; t < n; t++) {
            int d = metric == 0 ? hamming2bit(qq, targets[t]) : leven2bit(qq, targets[t], L);
            /* targets stream in ascending index, so on equal distance the earlier one stays */
            if (cnt == k && d >= bd[k - 1]) continue;
            int pos = cnt < k ? cnt : k - 1;
            while (pos > 0 && bd[pos - 1] > d) { bd[pos] = bd[pos - 1]; bi[pos] = bi[pos - 1]; pos--; }
            bd[pos] = d; bi[pos] = (int32_t)t;
            if (cnt < k) cnt++;
        }
        for (int j = 0; j < k; j++) {
            out_idx[qi * k + j] = j < cnt ? bi[j] : -1;
            out_dist[qi * k + j] = j < cnt ? (uint8_t)bd[j] : 255;
        }
    }
    return 0;
}

/* gmo_min_dist: distance of every query to its nearest target (control-sequence query,
 * core.py:603-606 uses only i[1][0]). */
int gmo_min_dist(const uint64_t *targets, int64_t n, const uint64_t *queries, int64_t q, int L, int metric,
                 uint8_t *out_dist, int threads)
{
#ifdef _OPENMP
    if (threads > 0) omp_set_num_threads(threads);
#else
    (void)threads;
#endif
#pragma omp parallel for schedule(dynamic, 64)
    for (int64_t qi = 0; qi < q; qi++) {
        int best = 255;
        uint64_t qq = queries[qi];
        for (int64_t t = 0; t < n; t++) {
            int d = metric == 0 ? hamming2bit(qq, targets[t]) : leven2bit(qq, targets[t], L);
            if (d < best) best = d;
        }
        out_dist[qi] = (uint8_t)best;
    }
    return 0;
}

/* ---- tuned Hamming kNN for the CPU ARM of the benchmark ---------------------------------------------------------------
 * Same result as gmo_knn(metric 0), bit for bit (checked in tests/test_oracle_golden.py); only the schedule differs:
 * guides as two 32-bit planes (mismatch mask = (qlo^tlo)|(qhi^thi), one popcount per pair), 16 targets per AVX-512
 * VPOPCNTD when the CPU has it, queries blocked 64 per task against cache-sized target chunks so the table streams
 * from DRAM once per 64 queries instead of once per query.  gmo_knn stays the checker; this is what bench.py times as
 * `cpu_baseline` / `--impl reference` ("tuned": true). */
#include <immintrin.h>

static inline uint32_t even_bits(uint64_t x) {
    x &= 0x5555555555555555ULL;
    x = (x | (x >> 1)) & 0x3333333333333333ULL;
    x = (x | (x >> 2)) & 0x0F0F0F0F0F0F0F0FULL;
    x = (x | (x >> 4)) & 0x00FF00FF00FF00FFULL;
    x = (x | (x >> 8)) & 0x0000FFFF0000FFFFULL;
    x = (x | (x >> 16)) & 0x00000000FFFFFFFFULL;
    return (uint32_t)x;
}

typedef struct { int cnt; int bd[GMO_MAXK]; int32_t bi[GMO_MAXK]; } gmo_list;

static inline void list_push(gmo_list *l, int k, int d, int32_t t) {
    if (l->cnt == k && d >= l->bd[k - 1]) return;
    int pos = l->cnt < k ? l->cnt : k - 1;
    while (pos > 0 && l->bd[pos - 1] > d) { l->bd[pos] = l->bd[pos - 1]; l->bi[pos] = l->bi[pos - 1]; pos--; }
    l->bd[pos] = d; l->bi[pos] = t;
    if (l->cnt < k) l->cnt++;
}

static void scan_chunk_scalar(const uint32_t *tlo, const uint32_t *thi, int64_t t0, int64_t t1, uint32_t qlo, uint32_t qhi,
                              gmo_list *l, int k) {
    for (int64_t t = t0; t < t1; t++) {
        const int d = __builtin_popcount((tlo[t] ^ qlo) | (thi[t] ^ qhi));
        if (l->cnt == k && d >= l->bd[k - 1]) continue;
        list_push(l, k, d, (int32_t)t);
    }
}

__attribute__((target("avx512f,avx512vpopcntdq")))
static void scan_chunk_avx512(const uint32_t *tlo, const uint32_t *thi, int64_t t0, int64_t t1, uint32_t qlo, uint32_t qhi,
                              gmo_list *l, int k) {
    const __m512i vql = _mm512_set1_epi32((int)qlo), vqh = _mm512_set1_epi32((int)qhi);
    int64_t t = t0;
    int tau = l->cnt == k ? l->bd[k - 1] : 64;                 /* insert iff d < tau */
    __m512i vtau = _mm512_set1_epi32(tau);
    for (; t + 16 <= t1; t += 16) {
        const __m512i a = _mm512_xor_si512(_mm512_loadu_si512((const void *)(tlo + t)), vql);
        const __m512i b = _mm512_xor_si512(_mm512_loadu_si512((const void *)(thi + t)), vqh);
        const __m512i d = _mm512_popcnt_epi32(_mm512_or_si512(a, b));
        __mmask16 m = _mm512_cmplt_epi32_mask(d, vtau);
        if (m) {
            int dd[16];
            _mm512_storeu_si512((void *)dd, d);
            while (m) {                                         /* ascending lane = ascending target index */
                const int j = __builtin_ctz(m);
                m &= (__mmask16)(m - 1);
                list_push(l, k, dd[j], (int32_t)(t + j));
            }
            tau = l->cnt == k ? l->bd[k - 1] : 64;
            vtau = _mm512_set1_epi32(tau);
        }
    }
    scan_chunk_scalar(tlo, thi, t, t1, qlo, qhi, l, k);
}

int gmo_has_avx512_popcnt(void) { return __builtin_cpu_supports("avx512vpopcntdq") && __builtin_cpu_supports("avx512f"); }

int gmo_knn_hamming_fast(const uint64_t *targets, int64_t n, const uint64_t *queries, int64_t q, int L, int k,
                         int32_t *out_idx, uint8_t *out_dist, int threads)
{
    if (k < 1 || k > GMO_MAXK || L < 1 || L > GMO_MAXL) return -2;
#ifdef _OPENMP
    if (threads > 0) omp_set_num_threads(threads);
#else
    (void)threads;
#endif
    uint32_t *tlo = (uint32_t *)malloc(sizeof(uint32_t) * (size_t)(n + 16)), *thi = (uint32_t *)malloc(sizeof(uint32_t) * (size_t)(n + 16));
    if (!tlo || !thi) { free(tlo); free(thi); return -3; }
#pragma omp parallel for schedule(static)
    for (int64_t t = 0; t < n; t++) { tlo[t] = even_bits(targets[t]); thi[t] = even_bits(targets[t] >> 1); }
    const int use512 = gmo_has_avx512_popcnt();
    const int64_t QB = 64, CH = 4096;                           /* 64 queries x 32 KB of targets per pass */
    const int64_t nblk = (q + QB - 1) / QB;
#pragma omp parallel for schedule(dynamic, 1)
    for (int64_t b = 0; b < nblk; b++) {
        const int64_t q0 = b * QB, q1 = q0 + QB < q ? q0 + QB : q;
        gmo_list lists[64];
        uint32_t qlo[64], qhi[64];
        for (int64_t i = q0; i < q1; i++) { lists[i - q0].cnt = 0; qlo[i - q0] = even_bits(queries[i]); qhi[i - q0] = even_bits(queries[i] >> 1); }
        for (int64_t c0 = 0; c0 < n; c0 += CH) {
            const int64_t c1 = c0 + CH < n ? c0 + CH : n;
            for (int64_t i = 0; i < q1 - q0; i++) {
                if (use512) scan_chunk_avx512(tlo, thi, c0, c1, qlo[i], qhi[i], &lists[i], k);
                else scan_chunk_scalar(tlo, thi, c0, c1, qlo[i], qhi[i], &lists[i], k);
            }
        }
        for (int64_t i = q0; i < q1; i++) {
            const gmo_list *l = &lists[i - q0];
            for (int j = 0; j < k; j++) {
                out_idx[i * k + j] = j < l->cnt ? l->bi[j] : -1;
                out_dist[i * k + j] = j < l->cnt ? (uint8_t)l->bd[j] : 255;
            }
        }
    }
    free(tlo); free(thi);
    return 0;
}

int gmo_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* ---- restriction-site flag (core.py:354-377) -------------------------------------------------
 * The reference expands every site and its reverse complement into all concrete strings
 * (extend_ambiguous_dna, core.py:1093-1124) and flags a guide when
 * targets.str.contains('|'.join(expansions)) -- a plain substring search for any expansion.
 * Restated directly: guide i is flagged iff for some motif t and offset o every motif position j
 * accepts base o + j of the guide.  iupac[t * 32 + j] holds the IUPAC LETTER of position j; the
 * accepted bases come from the table of core.py:1103-1120 (X and N accept all four).  A motif of
 * length 0 is the empty pattern and matches every guide, as the regex does. */
static int gmo_letter_set(char c)
{
    switch (c) {
    case 'A': return 1; case 'C': return 2; case 'G': return 4; case 'T': return 8;
    case 'M': return 1 | 2; case 'R': return 1 | 4; case 'W': return 1 | 8; case 'S': return 2 | 4;
    case 'Y': return 2 | 8; case 'K': return 4 | 8; case 'V': return 1 | 2 | 4; case 'H': return 1 | 2 | 8;
    case 'D': return 1 | 4 | 8; case 'B': return 2 | 4 | 8; case 'X': case 'N': return 15;
    default: return 0;
    }
}

int gmo_restriction(const uint64_t *guides, int64_t n, int L, const char *iupac, const int32_t *motif_len, int n_motifs,
                    uint8_t *has_site)
{
    for (int t = 0; t < n_motifs; t++)
        for (int j = 0; j < motif_len[t] && j < 32; j++)
            if (!gmo_letter_set(iupac[t * 32 + j])) return -2;
    for (int64_t i = 0; i < n; i++) {
        int hit = 0;
        for (int t = 0; t < n_motifs && !hit; t++) {
            const int len = motif_len[t];
            for (int o = 0; o + len <= L && !hit; o++) {
                int ok = 1;
                for (int j = 0; j < len && ok; j++) {
                    const int base = (int)((guides[i] >> (2 * (o + j))) & 3u);
                    ok = (gmo_letter_set(iupac[t * 32 + j]) >> base) & 1;
                }
                hit = ok;
            }
        }
        has_site[i] = (uint8_t)hit;
    }
    return 0;
}
