// alloc_probe.cu -- how long do first-touch device allocations and host<->device copies take on this box?
// (decides how libgm_b200 gets its scratch memory: stream-ordered pool vs cudaMalloc, pageable vs registered host buffers)
#include <cuda_runtime.h>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
static double now() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
int main() {
    cudaFree(0);
    const size_t MB = 1 << 20, sz = 512 * MB;
    double t0;
    void *p = nullptr, *q = nullptr;
    t0 = now(); cudaMalloc(&p, sz); cudaDeviceSynchronize(); printf("cudaMalloc 512 MB (first)            %8.2f ms\n", now() - t0);
    t0 = now(); cudaMemset(p, 1, sz); cudaDeviceSynchronize(); printf("  memset                              %8.2f ms\n", now() - t0);
    t0 = now(); cudaFree(p); printf("  cudaFree                            %8.2f ms\n", now() - t0);
    t0 = now(); cudaMalloc(&p, sz); cudaDeviceSynchronize(); printf("cudaMalloc 512 MB (second)           %8.2f ms\n", now() - t0);
    cudaFree(p);
    cudaMemPool_t pool; cudaDeviceGetDefaultMemPool(&pool, 0);
    unsigned long long keep = ~0ULL; cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    t0 = now(); cudaMallocAsync(&q, sz, 0); cudaStreamSynchronize(0); printf("cudaMallocAsync 512 MB (pool grows)  %8.2f ms\n", now() - t0);
    t0 = now(); cudaMemsetAsync(q, 1, sz, 0); cudaStreamSynchronize(0); printf("  memset                              %8.2f ms\n", now() - t0);
    cudaFreeAsync(q, 0); cudaStreamSynchronize(0);
    t0 = now(); cudaMallocAsync(&q, sz, 0); cudaStreamSynchronize(0); printf("cudaMallocAsync 512 MB (from pool)   %8.2f ms\n", now() - t0);
    t0 = now(); for (int i = 0; i < 8; i++) { void *r; cudaMallocAsync(&r, 64 * MB, 0); } cudaStreamSynchronize(0);
    printf("8 x cudaMallocAsync 64 MB (grows)    %8.2f ms\n", now() - t0);
    // host <-> device
    char *h = (char *)malloc(sz); memset(h, 1, sz);
    t0 = now(); cudaMemcpy(q, h, sz, cudaMemcpyHostToDevice); printf("H2D 512 MB pageable                  %8.2f ms\n", now() - t0);
    t0 = now(); cudaMemcpy(h, q, sz, cudaMemcpyDeviceToHost); printf("D2H 512 MB pageable (touched)        %8.2f ms\n", now() - t0);
    char *h2 = (char *)malloc(sz);
    t0 = now(); cudaMemcpy(h2, q, sz, cudaMemcpyDeviceToHost); printf("D2H 512 MB pageable (untouched dst)  %8.2f ms\n", now() - t0);
    t0 = now(); cudaHostRegister(h, sz, cudaHostRegisterDefault); printf("cudaHostRegister 512 MB              %8.2f ms\n", now() - t0);
    t0 = now(); cudaMemcpy(h, q, sz, cudaMemcpyDeviceToHost); printf("D2H 512 MB registered                %8.2f ms\n", now() - t0);
    t0 = now(); cudaMemcpy(q, h, sz, cudaMemcpyHostToDevice); printf("H2D 512 MB registered                %8.2f ms\n", now() - t0);
    t0 = now(); cudaHostUnregister(h); printf("cudaHostUnregister                   %8.2f ms\n", now() - t0);
    char *hp = nullptr;
    t0 = now(); cudaMallocHost(&hp, sz); printf("cudaMallocHost 512 MB                %8.2f ms\n", now() - t0);
    t0 = now(); cudaMemcpy(hp, q, sz, cudaMemcpyDeviceToHost); printf("D2H 512 MB pinned                    %8.2f ms\n", now() - t0);
    t0 = now(); memcpy(h2, hp, sz); printf("host memcpy 512 MB pinned->pageable  %8.2f ms\n", now() - t0);
    return 0;
}
