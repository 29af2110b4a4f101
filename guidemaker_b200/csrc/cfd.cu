// cfd.cu -- CFD (cutting frequency determination) score of guide / off-target pairs on packed guides.
//
// Replaces the per-row Python loop of guidemaker.core.cfd_score (core.py:1129-1148) over cfd_score_calculator.calc_cfd
// (cfd_score_calculator.py:62-85): score = product over the last 20 positions i of mm[r wt_i : d comp(off_i), pos] for
// every position where the two sequences differ; guides longer than 20 nt ignore their 5' end, shorter ones score the
// positions they have (pos = 20 + i + 1 - L).  The product runs in double precision in the reference's order
// (i ascending), so the result equals the Python float bit for bit.
#include "common.cuh"

namespace gm {

__constant__ double c_mm[4 * 4 * 20];      // [rna base A,C,G,U][dna base A,C,G,T][pos - 1]

__global__ void __launch_bounds__(256) cfd_kernel(const uint64_t *__restrict__ wt, const uint64_t *__restrict__ off, int64_t n, int k, int L,
                                                  double *__restrict__ out) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n * k) return;
    const uint64_t w = wt[t / k], o = off[t];
    double score = 1.0;
    for (int i = 0; i < L; i++) {
        if (L - 20 - i > 0) continue;                              // 5' positions beyond 20 nt are ignored
        const int wb = (int)((w >> (2 * i)) & 3u), ob = (int)((o >> (2 * i)) & 3u);
        if (wb != ob) score *= c_mm[(wb * 4 + (3 - ob)) * 20 + (20 + i - L)];      // dna base = complement: A<->T, C<->G = 3 - code
    }
    out[t] = score;
}

}  // namespace gm

using namespace gm;

extern "C" int gm_cfd_scores(const uint64_t *wt2bit, const uint64_t *off2bit, int64_t n, int k, int L, const double *mm_table, double *out) {
    int rc = ensure_init();
    if (rc) return rc;
    GM_ARG(n >= 0 && k >= 1 && L >= 1 && L <= GM_MAX_L, "gm_cfd_scores: bad size");
    if (n == 0) return GM_OK;
    GM_ARG(wt2bit && off2bit && mm_table && out, "gm_cfd_scores: NULL buffer");
    GM_CUDA(cudaMemcpyToSymbol(c_mm, mm_table, sizeof(double) * 320));
    uint64_t *d_w = nullptr, *d_o = nullptr;
    double *d_s = nullptr;
    cudaError_t e = dev_alloc((void **)&d_w, (size_t)n * 8, 0);
    if (e == cudaSuccess) e = dev_alloc((void **)&d_o, (size_t)n * k * 8, 0);
    if (e == cudaSuccess) e = dev_alloc((void **)&d_s, (size_t)n * k * 8, 0);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_w, wt2bit, (size_t)n * 8, cudaMemcpyHostToDevice, 0);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_o, off2bit, (size_t)n * k * 8, cudaMemcpyHostToDevice, 0);
    if (e == cudaSuccess) {
        cfd_kernel<<<(unsigned)((n * k + 255) / 256), 256>>>(d_w, d_o, n, k, L, d_s);
        count_launch();
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(out, d_s, (size_t)n * k * 8, cudaMemcpyDeviceToHost, 0);
    if (e == cudaSuccess) e = cudaStreamSynchronize(0);
    dev_free(d_w, 0); dev_free(d_o, 0); dev_free(d_s, 0);
    if (e != cudaSuccess) return cuda_fail(e, "gm_cfd_scores", __FILE__, __LINE__);
    return GM_OK;
}
