// Experiment: DRAM bytes per random 8-byte gather on B200 as a function of
// cudaLimitMaxL2FetchGranularity and of the load flavour.  Times a pure gather
// kernel over a 400 MB vector (x of BASELINE config 4) and prints ns and the
// implied gather rate.  Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o l2fg l2_fetch_granularity.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint64_t splitmix64(uint64_t x) {
    uint64_t z = x + 0x9e3779b97f4a7c15ULL;
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
    return z ^ (z >> 31);
}

template <int MODE>
__global__ void gather(const double *__restrict__ x, uint64_t n, double *out, int per_thread) {
    uint64_t t = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    double acc = 0;
    for (int i = 0; i < per_thread; i += 8) {
        double v[8];
#pragma unroll
        for (int u = 0; u < 8; u++) {
            uint64_t idx = __umul64hi(splitmix64(t * per_thread + i + u), n);
            const double *p = x + idx;
            if (MODE == 0) v[u] = __ldg(p);
            else if (MODE == 1) asm("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(v[u]) : "l"(p));
            else if (MODE == 2) asm("ld.global.cg.f64 %0, [%1];" : "=d"(v[u]) : "l"(p));
            else asm("ld.global.cv.f64 %0, [%1];" : "=d"(v[u]) : "l"(p));
        }
#pragma unroll
        for (int u = 0; u < 8; u++) acc += v[u];
    }
    if (acc == 12345.678) out[t] = acc;
}

int main(int argc, char **argv) {
    int gran = argc > 1 ? atoi(argv[1]) : -1;
    size_t cur = 0;
    cudaDeviceGetLimit(&cur, cudaLimitMaxL2FetchGranularity);
    printf("default cudaLimitMaxL2FetchGranularity = %zu\n", cur);
    if (gran > 0) {
        cudaError_t e = cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, gran);
        cudaDeviceGetLimit(&cur, cudaLimitMaxL2FetchGranularity);
        printf("set %d -> %s, now %zu\n", gran, cudaGetErrorString(e), cur);
    }
    const uint64_t n = 50000000;
    double *x, *out;
    cudaMalloc(&x, n * 8);
    cudaMemset(x, 0, n * 8);
    cudaMalloc(&out, 8);
    const int per_thread = 64, threads = 256;
    const uint64_t total = 1600000000ULL;
    const unsigned blocks = (unsigned)(total / per_thread / threads);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int mode = 0; mode < 4; mode++) {
        float best = 1e9;
        for (int rep = 0; rep < 3; rep++) {
            cudaEventRecord(e0);
            if (mode == 0) gather<0><<<blocks, threads>>>(x, n, out, per_thread);
            if (mode == 1) gather<1><<<blocks, threads>>>(x, n, out, per_thread);
            if (mode == 2) gather<2><<<blocks, threads>>>(x, n, out, per_thread);
            if (mode == 3) gather<3><<<blocks, threads>>>(x, n, out, per_thread);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            if (ms < best) best = ms;
        }
        printf("mode %d (%s): %.3f ms for %.1f G gathers -> %.1f G gathers/s\n", mode,
               mode == 0 ? "ld.global.nc" : mode == 1 ? "nc.L1::no_allocate" : mode == 2 ? "ld.global.cg" : "ld.global.cv",
               best, blocks * (double)threads * per_thread * 1e-9, blocks * (double)threads * per_thread / best * 1e-6);
    }
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
