/*
 * gm_b200.h -- C ABI of libgm_b200.so: the B200 (sm_100a) engine behind GuideMaker's off-target
 * hot path.  Plain pointers and sizes only; no C++/torch types cross this boundary.
 *
 * The reference (USDA-ARS-GBRU/GuideMaker, pure Python) has no FFI of its own for this path: its
 * four hot methods call three third-party native engines.  Each entry point below replaces one
 * of those call sites (file:line in /root/reference/guidemaker/core.py):
 *
 *   gm_scan_*            regex.finditer(pam_regex, seq, overlapped=True) + slicing/revcomp/
 *                        check_target                                   core.py:142-246 (call sites :154,:182,:207,:234)
 *   gm_gather_windows    the seq[...] slices (+ reverse_complement) behind target_seq30  core.py:156,184,210-211,237
 *   gm_seed_dedup        Series.duplicated() over the seed strings      core.py:402-416
 *   gm_first_occurrence  list(set(targets))  (made deterministic: first-occurrence order) core.py:446
 *   gm_restriction_scan  targets.str.contains('|'.join(expanded sites))  core.py:354-377 (call site :375)
 *   gm_index_create      nmslib.init + addDataPointBatch + createIndex  core.py:451-457, :461-467
 *   gm_knn               index.knnQueryBatch(queries, k=knum)           core.py:502-503
 *   gm_min_dist          index.knnQueryBatch(binseq, k=2) -> i[1][0]    core.py:603-606
 *
 * Conventions
 *   - "guide2bit": one guide per uint64, base i (0 = 5'-most) in bits [2i,2i+1], A=0 C=1 G=2 T=3,
 *     1 <= L <= 27.
 *   - Every function returns GM_OK (0) or a negative GM_ERR_* code; gm_last_error() gives the
 *     message (thread-local).  Nothing throws across the ABI.
 *   - The caller owns every buffer it passes; the library owns only what is behind its opaque
 *     handles (freed by the matching *_free).  Host-pointer calls are synchronous on return.
 *   - *_dev variants take DEVICE pointers and enqueue on the given cudaStream_t (passed as
 *     void*, NULL = default stream) without synchronising; they exist for HBM-resident
 *     pipelines (bench `value`, multi-GPU sharding over torch.distributed/NCCL buffers).
 *   - Distances are true mismatch / edit counts (NOT doubled as nmslib's one-hot bit_hamming).
 *     Neighbours are ordered ascending by (distance, target index); rows with fewer than k
 *     targets are padded with idx = -1, dist = 255.
 *   - There is no CPU fallback: without a CUDA device of compute capability 10.x gm_init fails.
 *   - Threads and streams: a handle (scan, session, index) must be used by ONE thread and ONE stream at a time -- an
 *     index owns a scratch buffer that every query call on it reuses, so two concurrent gm_knn* calls on the same index
 *     (from two threads, or enqueued on two streams without an event between them) race.  Different handles are
 *     independent.  The profiling counters (gm_prof_*) and the tuning defaults (gm_knn_tune / gm_knn_engine) are
 *     process-wide and unsynchronised: set them before starting worker threads.
 *     gm_session_create / gm_scan_create / the host-pointer calls use the default stream.
 */
#ifndef GM_B200_H
#define GM_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GM_OK          0
#define GM_ERR_CUDA   -1   /* a CUDA runtime call or kernel failed            */
#define GM_ERR_ARG    -2   /* invalid argument (bad PAM letter, L, k, NULL)   */
#define GM_ERR_NOMEM  -3   /* host or device allocation failed               */
#define GM_ERR_NODEV  -4   /* no usable sm_100 device                        */
#define GM_ERR_RANGE  -5   /* input exceeds a documented limit               */

#define GM_METRIC_HAMMING 0
#define GM_METRIC_LEVEN   1

#define GM_MAX_L    27      /* cli.py:43  guidelength 10..27 (one 32-bit plane per guide) */
#define GM_MAX_K    32      /* cli.py:55  knum 2..20                                      */
#define GM_MAX_PAM  8       /* cli.py:84  PAM length 2..8                                 */

/* ---- library ------------------------------------------------------------------------------- */
int         gm_init(int device);              /* select device, check sm_100, warm the context   */
const char *gm_last_error(void);
int         gm_version(void);
/* sm_count, compute capability and total HBM of the device chosen by gm_init */
int         gm_device_info(int *sm_count, int *cc_major, int *cc_minor, int64_t *mem_bytes);
/* The library keeps released device scratch blocks for reuse (no driver allocation in the steady state); gm_trim returns
 * the cached blocks to the driver (handles stay valid). */
int         gm_trim(void);

/* ---- K1: IUPAC PAM scan over both strands ------------------------------------------------------
 * seq_ascii: n bytes of one record (or several records joined by any non-ACGT byte).  Only
 * upper-case A/C/G/T match a PAM position or are allowed inside a target (core.py:118-121,:138).
 * Rows come out forward-strand hits first (ascending position), then reverse-strand hits
 * (ascending position) -- the reference's row order (core.py:254-284).
 * gm_scan_fetch fills, for n_fwd + n_rev rows: the guide as it appears in the `target` column
 * (reverse hits reverse-complemented), the 0-based forward-strand start of the target window,
 * and exact_pam packed 2 bits/base (PAM base j in bits [2j,2j+1], reverse hits
 * reverse-complemented).  Any output pointer may be NULL. */
int gm_scan_create(const uint8_t *seq_ascii, int64_t n, const char *pam, int pam_len,
                   int five_prime, int L, void **scan, int64_t *n_fwd, int64_t *n_rev);
int gm_scan_fetch(void *scan, uint64_t *guide2bit, uint32_t *start, uint16_t *pamcode);
int gm_scan_free(void *scan);
/* The 30-nt context column of find_targets (core.py:156,184,210-211,237): out[i*width + j] = byte j of the window
 * starting at win_start[i] in seq_ascii; rows with revcomp[i] != 0 are reversed and complemented (Bio.Seq IUPAC
 * table, other bytes unchanged).  Nothing is validated.  Windows not inside [0, n) are filled with '?'. */
int gm_gather_windows(const uint8_t *seq_ascii, int64_t n, const int64_t *win_start, const uint8_t *revcomp,
                      int64_t n_rows, int width, uint8_t *out);

/* ---- K2: keep-first duplicate flags -----------------------------------------------------------
 * is_dup[i] = 1 iff an earlier row has the same seed (first lsr bases if five_prime, last lsr
 * bases otherwise, the whole guide if lsr == 0) -- pandas duplicated(keep='first'). */
int gm_seed_dedup(const uint64_t *guide2bit, int64_t n, int L, int lsr, int five_prime,
                  uint8_t *is_dup);
/* first_row[i] = smallest row whose key equals keys[i] (distinct guides in first-occurrence
 * order are the rows with first_row[i] == i). n < 2^31. */
int gm_first_occurrence(const uint64_t *keys, int64_t n, int64_t *first_row);

/* device-resident variants: pointers into HBM, enqueued on `stream`, no synchronisation.  d_first_row is int32 (rows < 2^31). */
int gm_seed_dedup_dev(const uint64_t *d_guide2bit, int64_t n, int L, int lsr, int five_prime, uint8_t *d_is_dup, void *stream);
int gm_first_occurrence_dev(const uint64_t *d_keys, int64_t n, int32_t *d_first_row, void *stream);
/* the rows of a finished scan, still in HBM (valid until gm_scan_free): feed them to the *_dev entry points */
int gm_scan_device_ptrs(void *scan, const uint64_t **d_guide2bit, const uint32_t **d_start, const uint16_t **d_pamcode, int64_t *n_rows);

/* ---- session: the four hot methods chained on the device ------------------------------------------------------------
 * A session is a scan that (a) takes a multi-record genome -- records joined by ONE invalid byte, rec_start[r] = offset
 * of record r in seq_ascii, rec_start[n_rec] = n + 1 -- and emits the rows in the reference's order (per record: forward
 * hits ascending, then reverse hits ascending; core.py:254-284) with record-relative coordinates, and (b) keeps genome and
 * rows resident in HBM so that the later stages run off the handle without re-uploading anything:
 *   gm_session_fetch_rows   the numeric columns of find_targets' frame (any pointer may be NULL)
 *   gm_session_pam_histogram / gm_session_pam_categories   `exact_pam` as a categorical without n strings: the counts of the
 *                           65536 possible packed PAM codes, then -- with a caller-made code -> category table -- one int8
 *                           category code per row (core.py:167,195,222,248 build the column from Python strings)
 *   gm_session_fetch_text   `target` as ASCII (n_rows x L) and the 30-nt context `target_seq30` (n_rows x width) gathered
 *                           from the resident genome (core.py:156,184,210-211,237); edge[i] = 1 where the window leaves
 *                           the record (filled with '?': the caller applies Python's slice semantics to those rows)
 *   gm_session_seed_dedup   find_unique_near_pam's keep-first flags (core.py:402-416)
 *   gm_session_restriction  check_restriction_enzymes' flags (core.py:354-377), motifs as for gm_restriction_scan
 *   gm_session_index        distinct guides in first-occurrence order -> kNN index (core.py:446-467); uniq2bit must have
 *                           room for n_rows entries (n_u are written), row2uniq[i] = position of row i's guide in it
 *   gm_session_knn          get_neighbors (core.py:495-503): kNN of the rows with qmask[i] != 0 (host bytes), in row order;
 *                           n_q must equal the number of selected rows.  _dev: outputs are device pointers, no sync.
 * All calls on one session / index handle must come from one thread at a time. */
int gm_session_create(const uint8_t *seq_ascii, int64_t n, const int64_t *rec_start, int n_rec, const char *pam, int pam_len,
                      int five_prime, int L, void **session, int64_t *n_rows);
int gm_session_info(void *session, int64_t *n_rows, int *n_rec, int *L, int *P, int *five_prime);
int gm_session_fetch_rows(void *session, uint64_t *guide2bit, uint32_t *start, uint16_t *pamcode, int32_t *rec, uint8_t *strand);
int gm_session_fetch_text(void *session, uint8_t *target_ascii, uint8_t *context, int width, uint8_t *edge);
int gm_session_pam_histogram(void *session, uint32_t *hist65536);
int gm_session_pam_categories(void *session, const int8_t *lut65536, int8_t *codes);
int gm_session_seed_dedup(void *session, int lsr, uint8_t *is_dup);
int gm_session_restriction(void *session, const uint8_t *motif_sets, const int32_t *motif_len, int n_motifs, uint8_t *has_site);
int gm_session_index(void *session, int metric, void **index, uint64_t *uniq2bit, int32_t *row2uniq, int64_t *n_u);
int gm_session_knn(void *session, void *index, const uint8_t *qmask, int64_t n_q, int k, int32_t *out_idx, uint8_t *out_dist);
int gm_session_knn_dev(void *session, void *index, const uint8_t *qmask, int64_t n_q, int k, int32_t *d_out_idx,
                       uint8_t *d_out_dist, void *stream);
/* get_neighbors in one call (core.py:495-523): kNN of the masked rows, then -- on the device -- the distance filter
 * (keep a query iff its nearest OTHER guide is >= editdist away, core.py:512,518) and the one-entry-per-guide rule of the
 * reference's dict (the first query row of every guide); *n_kept rows stay on the device, *n_short = rows with fewer than
 * two hits (the reference raises IndexError then).  gm_session_fetch_neighbors copies the kept rows (guide2bit codes,
 * idx n_kept x k, dist n_kept x k, in row order) and releases them. */
int gm_session_neighbors(void *session, void *index, const uint8_t *qmask, int64_t n_q, int k, int editdist, int64_t *n_kept,
                         int64_t *n_short);
/* the same selection for kNN rows already on the device (n_q x k, query-row order) -- multi-GPU: after the all-gather */
int gm_session_filter_dev(void *session, const uint8_t *qmask, int64_t n_q, int k, int editdist, const int32_t *d_idx,
                          const uint8_t *d_dist, void *stream, int64_t *n_kept, int64_t *n_short);
int gm_session_fetch_neighbors(void *session, uint64_t *codes, int32_t *idx, uint8_t *dist);
int gm_session_free(void *session);

/* ---- K6: restriction-site flag ------------------------------------------------------------------
 * has_site[i] = 1 iff some motif occurs in guide i at any offset.  A motif is a string of letter sets:
 * motif_sets[t * 32 + j] = accepted bases of position j of motif t (bit 0 = A, 1 = C, 2 = G, 3 = T; 1..15),
 * motif_len[t] its length (0 = the empty pattern, which matches every guide as the reference's regex
 * does; a motif longer than L never matches).  The caller passes each enzyme site AND its reverse
 * complement, as core.py:367-370 does.  n_motifs = 0 clears the flags. */
int gm_restriction_scan(const uint64_t *guide2bit, int64_t n, int L, const uint8_t *motif_sets,
                        const int32_t *motif_len, int n_motifs, uint8_t *has_site);
int gm_restriction_scan_dev(const uint64_t *d_guide2bit, int64_t n, int L, const uint8_t *motif_sets,
                            const int32_t *motif_len, int n_motifs, uint8_t *d_has_site, void *stream);

/* ---- CFD off-target score (core.py:1129-1148, cfd_score_calculator.py:62-85) ------------------------------------------
 * out[i*k + j] = product, over the last 20 positions p where guide i and off-target (i, j) differ, of
 * mm_table[rna base of the guide][dna base = complement of the off-target base][pos - 1]; mm_table is 4 x 4 x 20 doubles
 * (A,C,G,U x A,C,G,T x position 1..20).  Double precision, positions in ascending order: equal to the reference's
 * Python float.  wt2bit: n guides, off2bit: n*k off-targets (row-major), both guide2bit of length L. */
int gm_cfd_scores(const uint64_t *wt2bit, const uint64_t *off2bit, int64_t n, int k, int L, const double *mm_table, double *out);

/* ---- K3/K4/K5: exact brute-force kNN index -----------------------------------------------------
 * The index is the table of distinct guides resident in HBM (bit-plane layout).
 * n_u < 2^27. */
int gm_index_create(const uint64_t *uniq2bit, int64_t n_u, int L, int metric, void **index);
int gm_index_create_dev(const uint64_t *d_uniq2bit, int64_t n_u, int L, int metric, void **index,
                        void *stream);
int gm_index_info(void *index, int64_t *n_u, int *L, int *metric);
int gm_index_free(void *index);

/* out_idx: q*k int32, out_dist: q*k uint8, row-major */
int gm_knn(void *index, const uint64_t *q2bit, int64_t q, int k, int32_t *out_idx,
           uint8_t *out_dist);
int gm_knn_dev(void *index, const uint64_t *d_q2bit, int64_t q, int k, int32_t *d_out_idx,
               uint8_t *d_out_dist, void *stream);
/* distance to the nearest indexed guide */
int gm_min_dist(void *index, const uint64_t *q2bit, int64_t q, uint8_t *out_dist);
int gm_min_dist_dev(void *index, const uint64_t *d_q2bit, int64_t q, uint8_t *d_out_dist,
                    void *stream);

/* ---- multi-GPU: one process per GPU, query rows sharded, guide table replicated -----------------------
 * The reference has no multi-device path.  Every rank builds the same index (gm_index_create / gm_session_index on its
 * own GPU), rank 0 makes an id with gm_comm_unique_id and the HOST carries those 128 bytes to the other ranks (MPI, a
 * file, a socket, torch.distributed ...); each rank then calls gm_comm_create (ncclCommInitRank on the gm_init device).
 * gm_knn_sharded is gm_knn for all ranks at once: every rank passes the SAME q query rows, uploads and searches only its
 * contiguous share [rank*q/world ...), the fixed-size (idx, dist) rows are all-gathered device-to-device with
 * ncclAllGather over NVLink, and every rank receives all q result rows in its host buffers.  NCCL is loaded lazily
 * (dlopen libnccl.so.2) by the first gm_comm_* call. */
#define GM_COMM_ID_BYTES 128
int gm_comm_unique_id(uint8_t *id128);
int gm_comm_create(const uint8_t *id128, int rank, int world, void **comm);
int gm_comm_free(void *comm);
int gm_knn_sharded(void *index, void *comm, const uint64_t *q2bit, int64_t q, int k, int32_t *out_idx, uint8_t *out_dist);

/* ---- measurement hooks ---------------------------------------------------------------------------
 * With profiling on, the library brackets its dominant kernel (the pair-scan of gm_knn*) with
 * CUDA events on the launching stream.  gm_prof_read synchronises those events and returns the
 * accumulated kernel time, the number of those launches, the (query, target) pairs they
 * evaluated, and the total number of kernels of this library launched since gm_prof_reset. */
int gm_prof_enable(int on);
int gm_prof_reset(void);
int gm_prof_read(double *scan_kernel_ms, int64_t *scan_kernel_launches, double *pairs,
                 int64_t *all_kernel_launches);

/* tuning knob for experiments and tests: queries per thread (4 or 8), target splits (0 = auto),
 * warm-start sample size (0 = off, -1 = default: for K3b on tables of >= 16384 guides the bound comes from the guides
 * around every query's rank in three sorted copies of the table (warm.cu), no sample; > 0 = the first `warm_sample`
 * guides of the table, scanned by a K3a / K4 launch) */
int gm_knn_tune(int queries_per_thread, int splits, int warm_sample);
/* Pair-scan engine.  Hamming: 1 = K3b tcgen05 kind::i8 GEMM over the 3-byte base code with the threshold test on the
 * TMEM read-out (default, ~9x faster), 0 = K3a XOR/POPC on the INT pipes.  Levenshtein: 1 = K4p, Myers' recurrence over
 * the prefix-sorted copy of the table with the DP states shared between consecutive guides (default, 1.55x), 0 = K4, the
 * plain scan in index order.  Every engine is exact and returns identical bits.  DESIGN.md section 4. */
int gm_knn_engine(int engine);
/* The same knobs for ONE index handle, overriding the process-wide defaults above: engine -1 / queries_per_thread -1 /
 * splits -1 / warm_sample -2 = follow the default.  Two indices with different engines can live in one process. */
int gm_index_tune(void *index, int engine, int queries_per_thread, int splits, int warm_sample);

/* microbenchmarks used as roofline denominators, whole GPU:
 * what = 0: POPC, 1: LOP3, 2: IMAD -> register-resident lane-operations per second;
 * what = 3: back-to-back tcgen05.mma kind::i8 (128x256x32, both operands in shared memory) -> int8 tensor
 *           operations per second (2 per MAC); 4: 128x128x32; 5 / 6: 128x128x32 / 128x256x32 with A in tensor memory;
 *           7 / 8: 128x128x32 SS / TS with consecutive MMAs reading different operand tiles (as the kNN kernel does);
 *           9 / 10: as 7 with one / two tcgen05.commit after every three MMAs; 11: as 10 with 128x64x32 MMAs. */
int gm_microbench(int what, double *ops_per_s);

#ifdef __cplusplus
}
#endif
#endif /* GM_B200_H */
