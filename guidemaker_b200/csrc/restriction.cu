// restriction.cu -- K6: restriction-site flag on packed guides.
//
// Replaces TargetProcessor.check_restriction_enzymes' `targets.str.contains('|'.join(expansions))`
// (core.py:354-377): a guide is flagged iff some IUPAC motif (each enzyme site and its reverse complement, expanded by
// the host wrapper) occurs in it at any offset.  The reference expands every ambiguous motif into all concrete
// strings and regex-searches their alternation; that is the same predicate as "at some offset o every motif position
// j accepts the guide base at o + j", which is evaluated here for all offsets at once on the guide's bit planes:
// for motif position j, pos_j = OR of the one-hot base masks of the accepted letters; match = AND_j (pos_j >> j),
// restricted to offsets 0 .. L - len.  One thread per guide, 8 bytes in, 1 byte out: HBM bound.
#include "scan.cuh"

namespace gm {

static constexpr int RS_MAX_MOTIFS = 64;      // per launch; longer lists run in batches that OR into the flags
static constexpr int RS_MAX_LEN = 32;

struct RsMotifs {
    int n;
    uint8_t len[RS_MAX_MOTIFS];
    uint8_t set[RS_MAX_MOTIFS][RS_MAX_LEN];   // accepted letters of motif position j: bit 0 = A, 1 = C, 2 = G, 3 = T
};

__global__ void __launch_bounds__(256) restriction_kernel(const uint64_t *__restrict__ guides, int64_t n, int L, const RsMotifs m,
                                                          int accumulate, uint8_t *__restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint2 p = to_planes(guides[i]);
    const uint32_t lmask = L >= 32 ? 0xFFFFFFFFu : (1u << L) - 1u;
    const uint32_t e[4] = {~(p.x | p.y) & lmask, p.x & ~p.y & lmask, p.y & ~p.x & lmask, p.x & p.y & lmask};
    bool hit = false;
    for (int t = 0; t < m.n; t++) {
        const int len = m.len[t];
        if (len > L) continue;                                        // a motif longer than the guide cannot occur
        uint32_t ok = (L - len + 1) >= 32 ? 0xFFFFFFFFu : (1u << (L - len + 1)) - 1u;   // offsets 0 .. L - len (len = 0: the empty
        for (int j = 0; j < len; j++) {                                                 // regex matches every guide)
            const uint32_t s = m.set[t][j];
            const uint32_t pos = ((s & 1u) ? e[0] : 0u) | ((s & 2u) ? e[1] : 0u) | ((s & 4u) ? e[2] : 0u) | ((s & 8u) ? e[3] : 0u);
            ok &= pos >> j;
        }
        hit |= ok != 0u;
    }
    out[i] = (uint8_t)((accumulate ? out[i] : 0) | (hit ? 1 : 0));
}

static int restriction_run(const uint64_t *d_guides, int64_t n, int L, const uint8_t *motif_sets, const int32_t *motif_len,
                           int n_motifs, uint8_t *d_out, cudaStream_t st) {
    GM_ARG(n >= 0 && n_motifs >= 0, "gm_restriction_scan: negative count");
    if (n == 0) return GM_OK;
    GM_ARG(d_guides && d_out, "gm_restriction_scan: NULL buffer");
    GM_ARG(L >= 1 && L <= GM_MAX_L, "gm_restriction_scan: L=%d outside [1,%d]", L, GM_MAX_L);
    GM_ARG(n_motifs == 0 || (motif_sets && motif_len), "gm_restriction_scan: NULL motif table");
    if (n_motifs == 0) { GM_CUDA(cudaMemsetAsync(d_out, 0, (size_t)n, st)); return GM_OK; }
    for (int base = 0; base < n_motifs; base += RS_MAX_MOTIFS) {
        RsMotifs m;
        memset(&m, 0, sizeof m);
        m.n = n_motifs - base < RS_MAX_MOTIFS ? n_motifs - base : RS_MAX_MOTIFS;
        for (int t = 0; t < m.n; t++) {
            const int len = motif_len[base + t];
            GM_ARG(len >= 0, "gm_restriction_scan: negative motif length");
            m.len[t] = (uint8_t)(len > RS_MAX_LEN ? RS_MAX_LEN + 1 : len);      // > GM_MAX_L: skipped by the kernel
            for (int j = 0; j < len && j < RS_MAX_LEN; j++) {
                const uint8_t s = motif_sets[(size_t)(base + t) * RS_MAX_LEN + j];
                GM_ARG(s >= 1 && s <= 15, "gm_restriction_scan: motif %d position %d has letter set %d outside [1,15]", base + t, j, (int)s);
                m.set[t][j] = s;
            }
        }
        restriction_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(d_guides, n, L, m, base > 0, d_out);
        count_launch();
    }
    GM_CUDA(cudaGetLastError());
    return GM_OK;
}

int restriction_dev(const uint64_t *d_guides, int64_t n, int L, const uint8_t *motif_sets, const int32_t *motif_len, int n_motifs,
                    uint8_t *d_has_site, cudaStream_t st) {
    return restriction_run(d_guides, n, L, motif_sets, motif_len, n_motifs, d_has_site, st);
}

}  // namespace gm

using namespace gm;

extern "C" int gm_restriction_scan_dev(const uint64_t *d_guide2bit, int64_t n, int L, const uint8_t *motif_sets, const int32_t *motif_len,
                                       int n_motifs, uint8_t *d_has_site, void *stream) {
    int rc = ensure_init();
    if (rc) return rc;
    return restriction_run(d_guide2bit, n, L, motif_sets, motif_len, n_motifs, d_has_site, (cudaStream_t)stream);
}

extern "C" int gm_restriction_scan(const uint64_t *guide2bit, int64_t n, int L, const uint8_t *motif_sets, const int32_t *motif_len,
                                   int n_motifs, uint8_t *has_site) {
    int rc = ensure_init();
    if (rc) return rc;
    GM_ARG(n >= 0, "gm_restriction_scan: negative row count");
    if (n == 0) return GM_OK;
    GM_ARG(guide2bit && has_site, "gm_restriction_scan: NULL buffer");
    uint64_t *d_g = nullptr;
    uint8_t *d_o = nullptr;
    cudaError_t e = dev_alloc((void **)&d_g, (size_t)n * 8, 0);
    if (e == cudaSuccess) e = dev_alloc((void **)&d_o, (size_t)n, 0);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_g, guide2bit, (size_t)n * 8, cudaMemcpyHostToDevice, 0);
    if (e == cudaSuccess) {
        rc = restriction_run(d_g, n, L, motif_sets, motif_len, n_motifs, d_o, 0);
        if (rc == GM_OK) e = cudaMemcpyAsync(has_site, d_o, (size_t)n, cudaMemcpyDeviceToHost, 0);
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(0);
    dev_free(d_g, 0);
    dev_free(d_o, 0);
    if (rc) return rc;
    if (e != cudaSuccess) return cuda_fail(e, "gm_restriction_scan", __FILE__, __LINE__);
    return GM_OK;
}
