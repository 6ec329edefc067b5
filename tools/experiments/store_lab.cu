// store_lab.cu -- phase 1 of the staged gather spends as much on storing the gathered values as on
// 0.6 gathers per entry (profiles/r2_phase1_lab.md: 277 G/s without the value stream, 176 G/s with
// it).  The shipped kernel gives each thread 4 consecutive entries and stores them as two 16-byte
// halves of a 32-byte sector: every store instruction half-fills 32 sectors.  Which store shape is
// cheap?  Short-lived threads as in sg_gather_kernel, 200 M entries, 48 MB table (32 MB hot).
//   V0  as shipped: thread owns 4 consecutive entries, two st.cs.v2.f64 (half sectors)
//   V1  thread owns 4 consecutive entries, one 256-bit st.v4.f64 (one full sector per lane)
//   V2  warp owns 128 consecutive entries, lane i the pairs (2i, 2i+1) and (64+2i, 64+2i+1):
//       every st.v2.f64 instruction writes 512 contiguous bytes
//   V3  warp owns 128 consecutive entries, lane i entries i, 32+i, 64+i, 96+i: 8-byte stores,
//       256 contiguous bytes per instruction
//   each with st.cs (evict first) and default st (suffix b)
//
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o store_lab store_lab.cu
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>

#define CK(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) { \
    fprintf(stderr, "%s:%d: %s -> %s\n", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); exit(1); } } while (0)

__device__ __forceinline__ uint32_t mix32(uint32_t v)
{
    v ^= v >> 16; v *= 0x7feb352dU; v ^= v >> 15; v *= 0x846ca68bU; v ^= v >> 16;
    return v;
}
__global__ void fill_index_kernel(int *__restrict__ idx, int64_t n, uint32_t mask)
{
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        idx[i] = (int)(mix32((uint32_t)i * 2654435761U + (uint32_t)(i >> 32)) & mask);
}
__global__ void spin_kernel(double *out, int iters)
{
    double a = threadIdx.x;
    for (int i = 0; i < iters; i++) a = a * 1.0000001 + 1e-9;
    if (a == 1.2345e300) out[0] = a;
}

template <bool CS> __device__ __forceinline__ void st2(double *p, double a, double b)
{
    if (CS) __stcs(reinterpret_cast<double2 *>(p), make_double2(a, b));
    else *reinterpret_cast<double2 *>(p) = make_double2(a, b);
}
template <bool CS> __device__ __forceinline__ void st1(double *p, double a)
{
    if (CS) __stcs(p, a); else *p = a;
}
template <bool CS> __device__ __forceinline__ void st4(double *p, double a, double b, double c, double d)
{
    if (CS) asm volatile("st.global.cs.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(p), "d"(a), "d"(b), "d"(c), "d"(d) : "memory");
    else asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(p), "d"(a), "d"(b), "d"(c), "d"(d) : "memory");
}

template <int V, bool CS>
__global__ void __launch_bounds__(256) gather_kernel(const int *__restrict__ idx, const double *__restrict__ x,
                                                     double *__restrict__ xg, int64_t n)
{
    const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (V == 0 || V == 1) {
        const int64_t e = t * 4;
        if (e >= n) return;
        const int4 c = __ldcs(reinterpret_cast<const int4 *>(idx + e));
        const double v0 = __ldg(x + c.x), v1 = __ldg(x + c.y), v2 = __ldg(x + c.z), v3 = __ldg(x + c.w);
        if (V == 0) { st2<CS>(xg + e, v0, v1); st2<CS>(xg + e + 2, v2, v3); }
        else st4<CS>(xg + e, v0, v1, v2, v3);
    } else if (V == 2) {
        const int lane = threadIdx.x & 31;
        const int64_t w0 = (t >> 5) * 128;                 // the warp's 128 entries
        if (w0 >= n) return;
        const int2 ca = __ldcs(reinterpret_cast<const int2 *>(idx + w0) + lane);
        const int2 cb = __ldcs(reinterpret_cast<const int2 *>(idx + w0 + 64) + lane);
        const double v0 = __ldg(x + ca.x), v1 = __ldg(x + ca.y), v2 = __ldg(x + cb.x), v3 = __ldg(x + cb.y);
        st2<CS>(xg + w0 + 2 * lane, v0, v1);
        st2<CS>(xg + w0 + 64 + 2 * lane, v2, v3);
    } else {
        const int lane = threadIdx.x & 31;
        const int64_t w0 = (t >> 5) * 128;
        if (w0 >= n) return;
        int c[4]; double v[4];
#pragma unroll
        for (int j = 0; j < 4; j++) c[j] = __ldcs(idx + w0 + 32 * j + lane);
#pragma unroll
        for (int j = 0; j < 4; j++) v[j] = __ldg(x + c[j]);
#pragma unroll
        for (int j = 0; j < 4; j++) st1<CS>(xg + w0 + 32 * j + lane, v[j]);
    }
}

static float time_ms(cudaEvent_t a, cudaEvent_t b) { float ms = 0; CK(cudaEventElapsedTime(&ms, a, b)); return ms; }

template <int V, bool CS>
static void run(const char *name, const int *idx, const double *x, double *xg, int64_t n, cudaEvent_t e0, cudaEvent_t e1)
{
    const unsigned grid = (unsigned)((n / 4 + 255) / 256);
    float best = 1e30f;
    for (int rep = 0; rep < 6; rep++) {
        CK(cudaEventRecord(e0));
        gather_kernel<V, CS><<<grid, 256>>>(idx, x, xg, n);
        CK(cudaEventRecord(e1));
        CK(cudaDeviceSynchronize());
        if (rep) best = time_ms(e0, e1) < best ? time_ms(e0, e1) : best;
    }
    printf("{\"lab\": \"store\", \"variant\": \"%s\", \"cs\": %d, \"entries\": %.3g, \"ms\": %.3f, \"Ggathers_per_s\": %.1f}\n",
           name, (int)CS, (double)n, best, (double)n / best * 1e-6);
    fflush(stdout);
}

int main()
{
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    const int sms = prop.multiProcessorCount;
    printf("# %s, %d SMs\n", prop.name, sms);
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    const int64_t n = 200LL * 1000 * 1000 / 1024 * 1024;
    const uint32_t words = 6u << 20, mask = (4u << 20) - 1;
    int *idx; double *x, *xg, *sink;
    CK(cudaMalloc(&idx, (size_t)n * 4)); CK(cudaMalloc(&x, (size_t)words * 8)); CK(cudaMalloc(&xg, (size_t)n * 8));
    CK(cudaMalloc(&sink, 1024));
    CK(cudaMemset(x, 0, (size_t)words * 8));
    fill_index_kernel<<<sms * 8, 256>>>(idx, n, mask);
    for (int i = 0; i < 50; i++) spin_kernel<<<sms * 8, 256>>>(sink, 400000);
    CK(cudaDeviceSynchronize());
    run<0, true>("V0 4 consecutive entries per thread, two 16-byte stores (as shipped)", idx, x, xg, n, e0, e1);
    run<0, false>("V0 4 consecutive entries per thread, two 16-byte stores (as shipped)", idx, x, xg, n, e0, e1);
    run<1, true>("V1 4 consecutive entries per thread, one 32-byte store", idx, x, xg, n, e0, e1);
    run<1, false>("V1 4 consecutive entries per thread, one 32-byte store", idx, x, xg, n, e0, e1);
    run<2, true>("V2 warp-contiguous 16-byte stores", idx, x, xg, n, e0, e1);
    run<2, false>("V2 warp-contiguous 16-byte stores", idx, x, xg, n, e0, e1);
    run<3, true>("V3 warp-contiguous 8-byte stores", idx, x, xg, n, e0, e1);
    run<3, false>("V3 warp-contiguous 8-byte stores", idx, x, xg, n, e0, e1);
    return 0;
}
