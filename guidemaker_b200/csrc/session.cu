// session.cu -- the device-resident chain behind the four hot methods of guidemaker.core.
//
// A session (gm_session_create, pam_scan.cu) keeps the genome and the rows of the PAM scan in HBM.  The stages that the
// reference runs one after the other on host strings
//      find_unique_near_pam (core.py:388-416)  ->  create_index (core.py:418-467)  ->  get_neighbors (core.py:471-523)
// take their input from that handle on one stream: nothing but the per-row results the host frame needs (1-byte flags,
// the distinct-guide table, the (idx, dist) rows) crosses PCIe, and no stage re-uploads the guides.
#include <cub/device/device_scan.cuh>
#include <cub/device/device_select.cuh>

#include "knn_common.cuh"
#include "scan.cuh"

namespace gm {

__global__ void first_flag_kernel(const int32_t *__restrict__ first32, int64_t n, int32_t *__restrict__ flag) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) flag[i] = first32[i] == (int32_t)i ? 1 : 0;
}

// uniq[rank[i]] = guides[i] for first occurrences; row2uniq[i] = rank[first32[i]]
__global__ void uniq_scatter_kernel(const uint64_t *__restrict__ guides, const int32_t *__restrict__ first32, const int32_t *__restrict__ rank,
                                    int64_t n, uint64_t *__restrict__ uniq, int32_t *__restrict__ row2uniq) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int32_t f = first32[i];
    const int32_t r = rank[f];
    row2uniq[i] = r;
    if (f == (int32_t)i) uniq[r] = guides[i];
}

// get_neighbors' selection (core.py:505-522) on the device: a query row is kept iff its nearest OTHER guide is at least
// `editdist` away (dist[1] >= editdist) and it is the first query row carrying its guide (the reference's dict keeps one
// entry per guide string).  short_rows counts rows with fewer than two hits (the reference raises IndexError there).
__global__ void neighbor_flag_kernel(const int32_t *__restrict__ idx, const uint8_t *__restrict__ dist, const int32_t *__restrict__ first_q,
                                     int64_t n_q, int k, int editdist, uint8_t *__restrict__ flag, unsigned long long *__restrict__ short_rows) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_q) return;
    const bool has2 = k >= 2 && idx[i * k + 1] >= 0;
    if (!has2) atomicAdd(short_rows, 1ULL);
    flag[i] = (has2 && (int)dist[i * k + 1] >= editdist && first_q[i] == (int32_t)i) ? 1 : 0;
}

__global__ void neighbor_gather_kernel(const int32_t *__restrict__ rows, int64_t n_keep, int k, const uint64_t *__restrict__ q,
                                       const int32_t *__restrict__ idx, const uint8_t *__restrict__ dist, uint64_t *__restrict__ o_codes,
                                       int32_t *__restrict__ o_idx, uint8_t *__restrict__ o_dist) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_keep * k) return;
    const int64_t r = t / k;
    const int j = (int)(t - r * k);
    const int64_t src = rows[r];
    o_idx[t] = idx[src * k + j];
    o_dist[t] = dist[src * k + j];
    if (j == 0) o_codes[r] = q[src];
}

__global__ void iota_kernel(int32_t *__restrict__ out, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = (int32_t)i;
}

static Scan *as_session(void *h) {
    Scan *s = (Scan *)h;
    return (s && s->session) ? s : nullptr;
}

}  // namespace gm

using namespace gm;

// ---- exact_pam as a categorical on the device (core.py:167,195,222,248 build it from n Python strings) -----------------
// histogram of the packed PAM codes: lanes holding the same code elect one to add their count (a genome has a handful of
// distinct PAMs, so plain atomics would all hit the same few words)
__global__ void pam_hist_kernel(const uint16_t *__restrict__ pamcode, int64_t n, unsigned int *__restrict__ hist) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool in = i < n;
    const unsigned active = __ballot_sync(0xFFFFFFFFu, in);
    if (!in) return;
    const unsigned code = pamcode[i];
    const unsigned peers = __match_any_sync(active, code);
    if ((threadIdx.x & 31) == __ffs(peers) - 1) atomicAdd(&hist[code], (unsigned)__popc(peers));
}
__global__ void pam_lut_kernel(const uint16_t *__restrict__ pamcode, int64_t n, const int8_t *__restrict__ lut, int8_t *__restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = lut[pamcode[i]];
}

extern "C" int gm_session_info(void *session, int64_t *n_rows, int *n_rec, int *L, int *P, int *five_prime) {
    Scan *s = as_session(session);
    GM_ARG(s, "gm_session_info: not a session handle");
    if (n_rows) *n_rows = s->n_fwd + s->n_rev;
    if (n_rec) *n_rec = s->n_rec;
    if (L) *L = s->L;
    if (P) *P = s->P;
    if (five_prime) *five_prime = s->five_prime;
    return GM_OK;
}

extern "C" int gm_session_seed_dedup(void *session, int lsr, uint8_t *is_dup) {
    Scan *s = as_session(session);
    GM_ARG(s && is_dup, "gm_session_seed_dedup: bad argument");
    const int64_t n = s->n_fwd + s->n_rev;
    if (n == 0) return GM_OK;
    const double t0 = now_ms();
    uint8_t *d = nullptr;
    GM_CUDA(dev_alloc((void **)&d, (size_t)n, 0));
    int rc = dedup_dev(s->guides, n, s->L, lsr, s->five_prime, d, nullptr, nullptr, 0);
    cudaError_t e = cudaSuccess;
    if (rc == GM_OK) e = cudaMemcpyAsync(is_dup, d, (size_t)n, cudaMemcpyDeviceToHost, 0);
    if (e == cudaSuccess) e = cudaStreamSynchronize(0);
    dev_free(d, 0);
    trace("session: seed flags", t0);
    if (rc) return rc;
    if (e != cudaSuccess) return cuda_fail(e, "gm_session_seed_dedup", __FILE__, __LINE__);
    return GM_OK;
}

extern "C" int gm_session_restriction(void *session, const uint8_t *motif_sets, const int32_t *motif_len, int n_motifs, uint8_t *has_site) {
    Scan *s = as_session(session);
    GM_ARG(s && has_site, "gm_session_restriction: bad argument");
    const int64_t n = s->n_fwd + s->n_rev;
    if (n == 0) return GM_OK;
    uint8_t *d = nullptr;
    GM_CUDA(dev_alloc((void **)&d, (size_t)n, 0));
    int rc = restriction_dev(s->guides, n, s->L, motif_sets, motif_len, n_motifs, d, 0);
    cudaError_t e = cudaSuccess;
    if (rc == GM_OK) e = cudaMemcpyAsync(has_site, d, (size_t)n, cudaMemcpyDeviceToHost, 0);
    if (e == cudaSuccess) e = cudaStreamSynchronize(0);
    dev_free(d, 0);
    if (rc) return rc;
    if (e != cudaSuccess) return cuda_fail(e, "gm_session_restriction", __FILE__, __LINE__);
    return GM_OK;
}

// Distinct guides in first-occurrence order (the deterministic stand-in for list(set(targets)), core.py:446) -> index.
extern "C" int gm_session_index(void *session, int metric, void **index, uint64_t *uniq2bit, int32_t *row2uniq, int64_t *n_u) {
    Scan *s = as_session(session);
    GM_ARG(s && index && n_u, "gm_session_index: bad argument");
    *index = nullptr;
    *n_u = 0;
    const int64_t n = s->n_fwd + s->n_rev;
    GM_ARG(n >= 1, "gm_session_index: empty guide table");
    const double t0 = now_ms();
    cudaStream_t st = 0;
    int32_t *d_flag = nullptr, *d_rank = nullptr, *d_r2u = nullptr;
    uint64_t *d_uniq = nullptr;
    void *d_tmp = nullptr;
    size_t tmp_bytes = 0;
    int rc = GM_OK;
    cudaError_t e = cudaSuccess;
    if (!s->first32) e = dev_alloc((void **)&s->first32, (size_t)n * 4, st);
    if (e == cudaSuccess) rc = dedup_dev(s->guides, n, GM_MAX_L, 0, 0, nullptr, nullptr, s->first32, st);
    if (trace_on()) { cudaStreamSynchronize(st); trace("  index: first occurrence", t0); }
    if (rc == GM_OK && e == cudaSuccess) e = dev_alloc((void **)&d_flag, (size_t)(n + 1) * 4, st);
    if (rc == GM_OK && e == cudaSuccess) e = dev_alloc((void **)&d_rank, (size_t)(n + 1) * 4, st);
    if (rc == GM_OK && e == cudaSuccess) e = dev_alloc((void **)&d_r2u, (size_t)n * 4, st);
    if (rc == GM_OK && e == cudaSuccess) e = dev_alloc((void **)&d_uniq, (size_t)n * 8, st);
    int32_t last[2] = {0, 0};
    if (rc == GM_OK && e == cudaSuccess) {
        const unsigned grid = (unsigned)((n + 255) / 256);
        first_flag_kernel<<<grid, 256, 0, st>>>(s->first32, n, d_flag);
        e = cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, d_flag, d_rank, (int)n, st);
        if (e == cudaSuccess) e = dev_alloc(&d_tmp, tmp_bytes, st);
        if (e == cudaSuccess) e = cub::DeviceScan::ExclusiveSum(d_tmp, tmp_bytes, d_flag, d_rank, (int)n, st);
        if (e == cudaSuccess) {
            uniq_scatter_kernel<<<grid, 256, 0, st>>>(s->guides, s->first32, d_rank, n, d_uniq, d_r2u);
            count_launch(4);
            e = cudaMemcpyAsync(&last[0], d_rank + (n - 1), 4, cudaMemcpyDeviceToHost, st);
        }
        if (e == cudaSuccess) e = cudaMemcpyAsync(&last[1], d_flag + (n - 1), 4, cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    }
    trace("  index: + compaction", t0);
    const int64_t nu = (int64_t)last[0] + last[1];
    if (rc == GM_OK && e == cudaSuccess) {
        rc = gm_index_create_dev(d_uniq, nu, s->L, metric, index, st);
        if (trace_on()) { cudaStreamSynchronize(st); trace("  index: + planes", t0); }
        prefault(uniq2bit, (size_t)nu * 8);
        prefault(row2uniq, (size_t)n * 4);
        if (rc == GM_OK && uniq2bit) e = cudaMemcpyAsync(uniq2bit, d_uniq, (size_t)nu * 8, cudaMemcpyDeviceToHost, st);
        if (rc == GM_OK && e == cudaSuccess && row2uniq) e = cudaMemcpyAsync(row2uniq, d_r2u, (size_t)n * 4, cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    }
    dev_free(d_flag, st); dev_free(d_rank, st); dev_free(d_r2u, st); dev_free(d_uniq, st); dev_free(d_tmp, st);
    trace("session: distinct guides + index", t0);
    if (rc == GM_OK && e != cudaSuccess) {
        if (*index) { gm_index_free(*index); *index = nullptr; }
        return cuda_fail(e, "gm_session_index", __FILE__, __LINE__);
    }
    if (rc) return rc;
    *n_u = nu;
    return GM_OK;
}

// kNN of the rows selected by `qmask` (one byte per row, host; the query mask of core.py:495), in row order.
// out_* receive n_q rows (n_q = number of non-zero mask bytes, checked).  With `device_out` the outputs are DEVICE
// pointers and the call returns without synchronising `stream` (multi-GPU: the caller all-gathers them first).
static int session_knn(void *session, void *index, const uint8_t *qmask, int64_t n_q, int k, int32_t *out_idx, uint8_t *out_dist,
                       bool device_out, cudaStream_t st) {
    Scan *s = as_session(session);
    GM_ARG(s && index && qmask, "gm_session_knn: bad argument");
    GM_ARG(k >= 1 && k <= GM_MAX_K, "gm_session_knn: k=%d outside [1,%d]", k, GM_MAX_K);
    const int64_t n = s->n_fwd + s->n_rev;
    if (n == 0 || n_q == 0) return GM_OK;
    GM_ARG(out_idx && out_dist && n_q > 0 && n_q <= n, "gm_session_knn: bad output buffers / n_q");
    const double t0 = now_ms();
    uint8_t *d_mask = nullptr;
    uint64_t *d_q = nullptr;
    int64_t *d_cnt = nullptr;
    int32_t *d_idx = device_out ? out_idx : nullptr;
    uint8_t *d_dist = device_out ? out_dist : nullptr;
    void *d_tmp = nullptr;
    size_t tmp_bytes = 0;
    int64_t cnt = -1;
    int rc = GM_OK;
    cudaError_t e = dev_alloc((void **)&d_mask, (size_t)n, st);
    if (e == cudaSuccess) e = dev_alloc((void **)&d_q, (size_t)n_q * 8 + 8, st);
    if (e == cudaSuccess) e = dev_alloc((void **)&d_cnt, 8, st);
    if (e == cudaSuccess && !device_out) e = dev_alloc((void **)&d_idx, (size_t)n_q * k * 4, st);
    if (e == cudaSuccess && !device_out) e = dev_alloc((void **)&d_dist, (size_t)n_q * k, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_mask, qmask, (size_t)n, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) {
        if (n_q == n) {                                            // every row is a query: no compaction (mask checked below)
            e = cudaMemcpyAsync(d_q, s->guides, (size_t)n * 8, cudaMemcpyDeviceToDevice, st);
            cnt = n;
        } else {
            // the mask must select exactly n_q rows or the compaction would overrun d_q: count first
            e = cub::DeviceSelect::Flagged(nullptr, tmp_bytes, s->guides, d_mask, d_q, d_cnt, (int)n, st);
            if (e == cudaSuccess) e = dev_alloc(&d_tmp, tmp_bytes, st);
            int64_t host_cnt = 0;
            for (int64_t i = 0; i < n; i++) host_cnt += qmask[i] != 0;
            if (host_cnt != n_q) { rc = GM_ERR_ARG; set_error("gm_session_knn: qmask selects %lld rows, n_q = %lld", (long long)host_cnt, (long long)n_q); }
            if (rc == GM_OK && e == cudaSuccess) e = cub::DeviceSelect::Flagged(d_tmp, tmp_bytes, s->guides, d_mask, d_q, d_cnt, (int)n, st);
            count_launch(2);
            cnt = n_q;
        }
    }
    if (rc == GM_OK && e == cudaSuccess) rc = gm_knn_dev(index, d_q, cnt, k, d_idx, d_dist, st);
    if (rc == GM_OK && e == cudaSuccess && !device_out) {
        prefault(out_idx, (size_t)n_q * k * 4);                    // overlaps the kernels enqueued above
        prefault(out_dist, (size_t)n_q * k);
        e = cudaMemcpyAsync(out_idx, d_idx, (size_t)n_q * k * 4, cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaMemcpyAsync(out_dist, d_dist, (size_t)n_q * k, cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    }
    dev_free(d_mask, st); dev_free(d_q, st); dev_free(d_cnt, st); dev_free(d_tmp, st);
    if (!device_out) { dev_free(d_idx, st); dev_free(d_dist, st); }
    trace("session: kNN", t0);
    if (rc) return rc;
    if (e != cudaSuccess) return cuda_fail(e, "gm_session_knn", __FILE__, __LINE__);
    return GM_OK;
}

extern "C" int gm_session_knn(void *session, void *index, const uint8_t *qmask, int64_t n_q, int k, int32_t *out_idx, uint8_t *out_dist) {
    int rc = ensure_init();
    if (rc) return rc;
    return session_knn(session, index, qmask, n_q, k, out_idx, out_dist, false, 0);
}

extern "C" int gm_session_knn_dev(void *session, void *index, const uint8_t *qmask, int64_t n_q, int k, int32_t *d_out_idx,
                                  uint8_t *d_out_dist, void *stream) {
    int rc = ensure_init();
    if (rc) return rc;
    return session_knn(session, index, qmask, n_q, k, d_out_idx, d_out_dist, true, (cudaStream_t)stream);
}

// The selection of get_neighbors (core.py:505-522) for kNN rows that are already on the device: d_q = the n_q query codes
// in row order, d_idx / d_dist = their (idx, dist) rows.  Leaves the kept rows compact in the session (nb_*).
static int neighbors_filter(Scan *s, const uint64_t *d_q, int64_t n_q, int k, int editdist, const int32_t *d_idx, const uint8_t *d_dist,
                            cudaStream_t st, int64_t *n_kept, int64_t *n_short) {
    dev_free(s->nb_codes, st); dev_free(s->nb_idx, st); dev_free(s->nb_dist, st);
    s->nb_codes = nullptr; s->nb_idx = nullptr; s->nb_dist = nullptr; s->nb_rows = 0; s->nb_k = k;
    uint8_t *d_flag = nullptr;
    int32_t *d_first = nullptr, *d_iota = nullptr, *d_rows = nullptr;
    unsigned long long *d_cnt = nullptr;          // [0] rows short of two hits, [1] kept rows
    void *d_tmp = nullptr;
    size_t tmp_bytes = 0;
    long long kept = 0;
    unsigned long long short_rows = 0;
    int rc = GM_OK;
    cudaError_t e = dev_alloc((void **)&d_flag, (size_t)n_q, st);
    if (e == cudaSuccess) e = dev_alloc((void **)&d_first, (size_t)n_q * 4, st);
    if (e == cudaSuccess) e = dev_alloc((void **)&d_iota, (size_t)n_q * 4, st);
    if (e == cudaSuccess) e = dev_alloc((void **)&d_rows, (size_t)n_q * 4, st);
    if (e == cudaSuccess) e = dev_alloc((void **)&d_cnt, 16, st);
    if (e == cudaSuccess) e = cudaMemsetAsync(d_cnt, 0, 16, st);
    if (e == cudaSuccess) e = cub::DeviceSelect::Flagged(nullptr, tmp_bytes, d_iota, d_flag, d_rows, (long long *)(d_cnt + 1), (int)n_q, st);
    if (e == cudaSuccess) e = dev_alloc(&d_tmp, tmp_bytes, st);
    if (e == cudaSuccess) rc = dedup_dev(d_q, n_q, GM_MAX_L, 0, 0, nullptr, nullptr, d_first, st);   // first query row per guide
    if (rc == GM_OK && e == cudaSuccess) {
        const unsigned grid = (unsigned)((n_q + 255) / 256);
        neighbor_flag_kernel<<<grid, 256, 0, st>>>(d_idx, d_dist, d_first, n_q, k, editdist, d_flag, d_cnt);
        iota_kernel<<<grid, 256, 0, st>>>(d_iota, n_q);
        e = cub::DeviceSelect::Flagged(d_tmp, tmp_bytes, d_iota, d_flag, d_rows, (long long *)(d_cnt + 1), (int)n_q, st);
        count_launch(4);
        if (e == cudaSuccess) e = cudaMemcpyAsync(&short_rows, d_cnt, 8, cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaMemcpyAsync(&kept, d_cnt + 1, 8, cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    }
    if (rc == GM_OK && e == cudaSuccess && kept > 0) {
        e = dev_alloc((void **)&s->nb_codes, (size_t)kept * 8, st);
        if (e == cudaSuccess) e = dev_alloc((void **)&s->nb_idx, (size_t)kept * k * 4, st);
        if (e == cudaSuccess) e = dev_alloc((void **)&s->nb_dist, (size_t)kept * k, st);
        if (e == cudaSuccess) {
            neighbor_gather_kernel<<<(unsigned)((kept * k + 255) / 256), 256, 0, st>>>(d_rows, kept, k, d_q, d_idx, d_dist, s->nb_codes, s->nb_idx, s->nb_dist);
            count_launch();
            e = cudaGetLastError();
        }
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    }
    dev_free(d_flag, st); dev_free(d_first, st); dev_free(d_iota, st); dev_free(d_rows, st); dev_free(d_cnt, st); dev_free(d_tmp, st);
    if (rc) return rc;
    if (e != cudaSuccess) return cuda_fail(e, "neighbour filter", __FILE__, __LINE__);
    s->nb_rows = kept;
    *n_kept = kept;
    *n_short = (int64_t)short_rows;
    return GM_OK;
}

// compaction of the masked rows' guides: d_q[0 .. n_q) = guides[qmask != 0]
static int masked_queries(Scan *s, const uint8_t *qmask, int64_t n_q, uint64_t *d_q, cudaStream_t st) {
    const int64_t n = s->n_fwd + s->n_rev;
    int64_t host_cnt = 0;
    for (int64_t i = 0; i < n; i++) host_cnt += qmask[i] != 0;
    if (host_cnt != n_q) { set_error("qmask selects %lld rows, n_q = %lld", (long long)host_cnt, (long long)n_q); return GM_ERR_ARG; }
    if (n_q == n) { GM_CUDA(cudaMemcpyAsync(d_q, s->guides, (size_t)n * 8, cudaMemcpyDeviceToDevice, st)); return GM_OK; }
    uint8_t *d_mask = nullptr;
    long long *d_cnt = nullptr;
    void *d_tmp = nullptr;
    size_t tmp_bytes = 0;
    cudaError_t e = dev_alloc((void **)&d_mask, (size_t)n, st);
    if (e == cudaSuccess) e = dev_alloc((void **)&d_cnt, 8, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_mask, qmask, (size_t)n, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) e = cub::DeviceSelect::Flagged(nullptr, tmp_bytes, s->guides, d_mask, d_q, d_cnt, (int)n, st);
    if (e == cudaSuccess) e = dev_alloc(&d_tmp, tmp_bytes, st);
    if (e == cudaSuccess) e = cub::DeviceSelect::Flagged(d_tmp, tmp_bytes, s->guides, d_mask, d_q, d_cnt, (int)n, st);
    count_launch(2);
    dev_free(d_mask, st); dev_free(d_cnt, st); dev_free(d_tmp, st);
    if (e != cudaSuccess) return cuda_fail(e, "query compaction", __FILE__, __LINE__);
    return GM_OK;
}

// get_neighbors in one call (core.py:495-523): kNN of the masked rows, the distance filter and the one-entry-per-guide
// rule applied on the device; only the kept rows are copied to the host (gm_session_fetch_neighbors), already compact.
extern "C" int gm_session_neighbors(void *session, void *index, const uint8_t *qmask, int64_t n_q, int k, int editdist, int64_t *n_kept,
                                    int64_t *n_short) {
    int rc = ensure_init();
    if (rc) return rc;
    Scan *s = as_session(session);
    GM_ARG(s && index && qmask && n_kept && n_short, "gm_session_neighbors: bad argument");
    GM_ARG(k >= 1 && k <= GM_MAX_K, "gm_session_neighbors: k=%d outside [1,%d]", k, GM_MAX_K);
    *n_kept = *n_short = 0;
    const int64_t n = s->n_fwd + s->n_rev;
    if (n == 0 || n_q == 0) return GM_OK;
    GM_ARG(n_q > 0 && n_q <= n, "gm_session_neighbors: bad n_q");
    const double t0 = now_ms();
    cudaStream_t st = 0;
    uint64_t *d_q = nullptr;
    int32_t *d_idx = nullptr;
    uint8_t *d_dist = nullptr;
    cudaError_t e = dev_alloc((void **)&d_q, (size_t)n_q * 8 + 8, st);
    if (e == cudaSuccess) e = dev_alloc((void **)&d_idx, (size_t)n_q * k * 4, st);
    if (e == cudaSuccess) e = dev_alloc((void **)&d_dist, (size_t)n_q * k, st);
    if (e == cudaSuccess) rc = masked_queries(s, qmask, n_q, d_q, st);
    if (rc == GM_OK && e == cudaSuccess) rc = gm_knn_dev(index, d_q, n_q, k, d_idx, d_dist, st);
    if (rc == GM_OK && e == cudaSuccess) rc = neighbors_filter(s, d_q, n_q, k, editdist, d_idx, d_dist, st, n_kept, n_short);
    dev_free(d_q, st); dev_free(d_idx, st); dev_free(d_dist, st);
    trace("session: kNN + neighbour filter", t0);
    if (rc) return rc;
    if (e != cudaSuccess) return cuda_fail(e, "gm_session_neighbors", __FILE__, __LINE__);
    return GM_OK;
}

// The same selection for kNN rows the caller already holds on the device -- the multi-GPU path: every rank searches its
// shard (gm_session_knn_dev), the rows are all-gathered device to device, and each rank filters the gathered table here.
// d_idx / d_dist: n_q x k rows in query-row order (device); enqueued on `stream`, synchronised before returning.
extern "C" int gm_session_filter_dev(void *session, const uint8_t *qmask, int64_t n_q, int k, int editdist, const int32_t *d_idx,
                                     const uint8_t *d_dist, void *stream, int64_t *n_kept, int64_t *n_short) {
    int rc = ensure_init();
    if (rc) return rc;
    Scan *s = as_session(session);
    GM_ARG(s && qmask && d_idx && d_dist && n_kept && n_short, "gm_session_filter_dev: bad argument");
    GM_ARG(k >= 1 && k <= GM_MAX_K, "gm_session_filter_dev: k=%d outside [1,%d]", k, GM_MAX_K);
    *n_kept = *n_short = 0;
    const int64_t n = s->n_fwd + s->n_rev;
    if (n == 0 || n_q == 0) return GM_OK;
    GM_ARG(n_q > 0 && n_q <= n, "gm_session_filter_dev: bad n_q");
    const double t0 = now_ms();
    cudaStream_t st = (cudaStream_t)stream;
    uint64_t *d_q = nullptr;
    GM_CUDA(dev_alloc((void **)&d_q, (size_t)n_q * 8 + 8, st));
    rc = masked_queries(s, qmask, n_q, d_q, st);
    if (rc == GM_OK) rc = neighbors_filter(s, d_q, n_q, k, editdist, d_idx, d_dist, st, n_kept, n_short);
    dev_free(d_q, st);
    trace("session: neighbour filter (gathered rows)", t0);
    return rc;
}

extern "C" int gm_session_fetch_neighbors(void *session, uint64_t *codes, int32_t *idx, uint8_t *dist) {
    Scan *s = as_session(session);
    GM_ARG(s, "gm_session_fetch_neighbors: not a session handle");
    if (s->nb_rows == 0) return GM_OK;
    GM_ARG(codes && idx && dist, "gm_session_fetch_neighbors: NULL buffer");
    const double t0 = now_ms();
    const size_t r = (size_t)s->nb_rows, k = (size_t)s->nb_k;
    prefault(codes, r * 8); prefault(idx, r * k * 4); prefault(dist, r * k);
    GM_CUDA(cudaMemcpyAsync(codes, s->nb_codes, r * 8, cudaMemcpyDeviceToHost, 0));
    GM_CUDA(cudaMemcpyAsync(idx, s->nb_idx, r * k * 4, cudaMemcpyDeviceToHost, 0));
    GM_CUDA(cudaMemcpyAsync(dist, s->nb_dist, r * k, cudaMemcpyDeviceToHost, 0));
    GM_CUDA(cudaStreamSynchronize(0));
    dev_free(s->nb_codes, 0); dev_free(s->nb_idx, 0); dev_free(s->nb_dist, 0);
    s->nb_codes = nullptr; s->nb_idx = nullptr; s->nb_dist = nullptr; s->nb_rows = 0;
    trace("session: fetch neighbours", t0);
    return GM_OK;
}

extern "C" int gm_session_pam_histogram(void *session, uint32_t *hist65536) {
    Scan *s = (Scan *)session;
    GM_ARG(s && s->session && hist65536, "gm_session_pam_histogram: bad argument");
    const int64_t nt = s->n_fwd + s->n_rev;
    unsigned int *d = nullptr;
    GM_CUDA(dev_alloc((void **)&d, 65536 * sizeof(unsigned int), 0));
    cudaError_t e = cudaMemsetAsync(d, 0, 65536 * sizeof(unsigned int), 0);
    if (e == cudaSuccess && nt > 0) {
        pam_hist_kernel<<<(unsigned)((nt + 255) / 256), 256>>>(s->pamcode, nt, d);
        count_launch();
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(hist65536, d, 65536 * sizeof(unsigned int), cudaMemcpyDeviceToHost, 0);
    if (e == cudaSuccess) e = cudaStreamSynchronize(0);
    dev_free(d, 0);
    if (e != cudaSuccess) return cuda_fail(e, "gm_session_pam_histogram", __FILE__, __LINE__);
    return GM_OK;
}

extern "C" int gm_session_pam_categories(void *session, const int8_t *lut65536, int8_t *codes) {
    Scan *s = (Scan *)session;
    GM_ARG(s && s->session && lut65536, "gm_session_pam_categories: bad argument");
    const int64_t nt = s->n_fwd + s->n_rev;
    if (nt == 0) return GM_OK;
    GM_ARG(codes, "gm_session_pam_categories: NULL output");
    int8_t *d_lut = nullptr, *d_out = nullptr;
    cudaError_t e = dev_alloc((void **)&d_lut, 65536, 0);
    if (e == cudaSuccess) e = dev_alloc((void **)&d_out, (size_t)nt, 0);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_lut, lut65536, 65536, cudaMemcpyHostToDevice, 0);
    if (e == cudaSuccess) {
        pam_lut_kernel<<<(unsigned)((nt + 255) / 256), 256>>>(s->pamcode, nt, d_lut, d_out);
        count_launch();
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) {
        prefault(codes, (size_t)nt);
        e = cudaMemcpyAsync(codes, d_out, (size_t)nt, cudaMemcpyDeviceToHost, 0);
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(0);
    dev_free(d_lut, 0); dev_free(d_out, 0);
    if (e != cudaSuccess) return cuda_fail(e, "gm_session_pam_categories", __FILE__, __LINE__);
    return GM_OK;
}
