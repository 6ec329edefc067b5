// gather_paths.cu -- how many scattered 8-byte reads per second does a B200 serve, by path?
//
// Background (profiles/r1_staged_gather.md): phase 1 of the staged gather and the column-blocked
// kernel both top out at ~175-195 G gathers/s because every scattered LDG is its own 128-byte-line
// request on the SM -> L2 port.  This micro-benchmark measures the alternatives that DESIGN.md 8
// lists before anyone builds a kernel on them:
//   ldg     scattered ld.global.nc from an L2-resident table (the baseline; expect ~175 G/s)
//   dsmem   the table spread over the shared memory of an 8-CTA cluster, scattered
//           ld.shared::cluster reads (local and remote CTAs alike)
//   smem    the same reads from the CTA's own shared memory only (upper bound of a table that fits one SM)
//   bulk    scattered 16-byte cp.async.bulk global -> shared copies (the bulk-copy engine's own
//           request path), completion counted on an mbarrier
// Indices are generated in registers (a hash of a counter), so nothing but the gathers touches memory.
//
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o gather_paths gather_paths.cu
//   timeout 60 ./gather_paths
#include <cooperative_groups.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>

namespace cg = cooperative_groups;

#define CK(call)                                                                                  \
    do {                                                                                          \
        cudaError_t e__ = (call);                                                                 \
        if (e__ != cudaSuccess) {                                                                 \
            fprintf(stderr, "%s:%d: %s -> %s\n", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
            exit(1);                                                                              \
        }                                                                                         \
    } while (0)

__device__ __forceinline__ uint32_t mix32(uint32_t v)
{
    v ^= v >> 16; v *= 0x7feb352dU; v ^= v >> 15; v *= 0x846ca68bU; v ^= v >> 16;
    return v;
}

constexpr int kUnroll = 8;

// ---- ldg: scattered reads from a global table (L2-resident when it is <= ~48 MB) -------------
__global__ void __launch_bounds__(256) ldg_kernel(const double *__restrict__ table, uint32_t mask, int iters,
                                                  double *__restrict__ out)
{
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
    double acc = 0.0;
    uint32_t ctr = tid * 2654435761U;
    for (int it = 0; it < iters; it++) {
        double v[kUnroll];
#pragma unroll
        for (int u = 0; u < kUnroll; u++) v[u] = __ldg(table + (mix32(ctr + u) & mask));
#pragma unroll
        for (int u = 0; u < kUnroll; u++) acc += v[u];
        ctr += kUnroll;
    }
    if (acc == 1.2345e300) out[tid] = acc;   // never true: keeps the loads alive
}

// ---- dsmem / smem: the table lives in the cluster's (or the CTA's) shared memory ----------------
template <bool CLUSTER>
__global__ void __launch_bounds__(1024) smem_kernel(int words_per_cta, int iters, double *__restrict__ out)
{
    extern __shared__ double tab[];
    cg::cluster_group cluster = cg::this_cluster();
    for (int i = threadIdx.x; i < words_per_cta; i += blockDim.x) tab[i] = (double)(i + blockIdx.x);
    if (CLUSTER) cluster.sync(); else __syncthreads();
    const unsigned nranks = CLUSTER ? cluster.num_blocks() : 1;
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
    double acc = 0.0;
    uint32_t ctr = tid * 2654435761U;
    const uint32_t wmask = (uint32_t)words_per_cta - 1;        // power of two
    for (int it = 0; it < iters; it++) {
        double v[kUnroll];
#pragma unroll
        for (int u = 0; u < kUnroll; u++) {
            const uint32_t h = mix32(ctr + u);
            const double *p = tab + (h & wmask);
            if (CLUSTER) p = cluster.map_shared_rank(const_cast<double *>(p), (h >> 20) % nranks);
            v[u] = *p;
        }
#pragma unroll
        for (int u = 0; u < kUnroll; u++) acc += v[u];
        ctr += kUnroll;
    }
    if (CLUSTER) cluster.sync();                                // nobody exits while peers still read its memory
    if (acc == 1.2345e300) out[tid] = acc;
}

// ---- bulk: scattered 16-byte bulk copies global -> shared ------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void __launch_bounds__(128) bulk_kernel(const double *__restrict__ table, uint32_t mask, int iters,
                                                   double *__restrict__ out)
{
    __shared__ __align__(16) double land[128 * kUnroll * 2];     // 16 bytes per copy
    __shared__ __align__(8) uint64_t bar;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t ctr = tid * 2654435761U;
    double acc = 0.0;
    for (int it = 0; it < iters; it++) {
        if (threadIdx.x == 0)
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;"
                         ::"r"(smem_u32(&bar)), "r"(128u * kUnroll * 16u) : "memory");
        __syncthreads();
#pragma unroll
        for (int u = 0; u < kUnroll; u++) {
            const double *src = table + ((mix32(ctr + u) & mask) & ~1u);          // 16-byte aligned
            double *dst = land + (threadIdx.x * kUnroll + u) * 2;
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], 16, [%2];"
                         ::"r"(smem_u32(dst)), "l"(src), "r"(smem_u32(&bar)) : "memory");
        }
        uint32_t ok = 0;
        for (int spins = 0; !ok && spins < (1 << 22); spins++)
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                         "selp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(&bar)), "r"((uint32_t)(it & 1)) : "memory");
        acc += land[threadIdx.x * kUnroll * 2];
        ctr += kUnroll;
        __syncthreads();                       // everyone has read before the buffer is refilled
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (acc == 1.2345e300) out[tid] = acc;
}

// ---- phase1: the staged gather's phase 1 in isolation -- indices streamed from HBM, gathered
// values streamed back -- with short-lived threads (one group of 4 per thread, as shipped in
// round 1) and with persistent threads (grid-stride over G groups, all index loads of a batch
// first, then all gathers, then the stores)
__global__ void fill_index_kernel(int *__restrict__ idx, int64_t n, uint32_t mask)
{
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        idx[i] = (int)(mix32((uint32_t)i * 2654435761U + (uint32_t)(i >> 32)) & mask);
}

__global__ void __launch_bounds__(256) phase1_short_kernel(const int *__restrict__ idx, const double *__restrict__ table,
                                                           double *__restrict__ xg, int64_t groups)
{
    const int64_t g = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (g >= groups) return;
    const int4 c = __ldcs(reinterpret_cast<const int4 *>(idx) + g);
    const double v0 = __ldg(table + c.x), v1 = __ldg(table + c.y), v2 = __ldg(table + c.z), v3 = __ldg(table + c.w);
    __stcs(reinterpret_cast<double2 *>(xg) + 2 * g, make_double2(v0, v1));
    __stcs(reinterpret_cast<double2 *>(xg) + 2 * g + 1, make_double2(v2, v3));
}

template <int G>
__global__ void __launch_bounds__(256) phase1_persistent_kernel(const int *__restrict__ idx, const double *__restrict__ table,
                                                                double *__restrict__ xg, int64_t groups)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t g0 = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; g0 < groups; g0 += stride * G) {
        int4 c[G];
        double v[G][4];
#pragma unroll
        for (int j = 0; j < G; j++) {
            const int64_t g = g0 + j * stride;
            c[j] = g < groups ? __ldcs(reinterpret_cast<const int4 *>(idx) + g) : make_int4(0, 0, 0, 0);
        }
#pragma unroll
        for (int j = 0; j < G; j++) {
            v[j][0] = __ldg(table + c[j].x); v[j][1] = __ldg(table + c[j].y);
            v[j][2] = __ldg(table + c[j].z); v[j][3] = __ldg(table + c[j].w);
        }
#pragma unroll
        for (int j = 0; j < G; j++) {
            const int64_t g = g0 + j * stride;
            if (g < groups) {
                __stcs(reinterpret_cast<double2 *>(xg) + 2 * g, make_double2(v[j][0], v[j][1]));
                __stcs(reinterpret_cast<double2 *>(xg) + 2 * g + 1, make_double2(v[j][2], v[j][3]));
            }
        }
    }
}

static float time_ms(cudaEvent_t a, cudaEvent_t b) { float ms = 0; CK(cudaEventElapsedTime(&ms, a, b)); return ms; }

int main()
{
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    const int sms = prop.multiProcessorCount;
    printf("# %s, %d SMs\n", prop.name, sms);
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    double *out;
    CK(cudaMalloc(&out, (size_t)sms * 64 * 1024 * 8));

    // ---- ldg, tables of 1.6 MB (the size a cluster's shared memory would hold), 32 MB and 1 GB
    for (uint32_t words : {1u << 18, 1u << 22, 1u << 27}) {
        double *table;
        CK(cudaMalloc(&table, (size_t)words * 8));
        CK(cudaMemset(table, 0, (size_t)words * 8));
        const int iters = 64, grid = sms * 64, threads = 256;
        ldg_kernel<<<grid, threads>>>(table, words - 1, 4, out);          // warm-up: pulls the table into L2
        CK(cudaEventRecord(e0));
        ldg_kernel<<<grid, threads>>>(table, words - 1, iters, out);
        CK(cudaEventRecord(e1));
        CK(cudaDeviceSynchronize());
        const double n = (double)grid * threads * iters * kUnroll;
        printf("{\"path\": \"ldg\", \"table_MB\": %.1f, \"gathers\": %.3g, \"ms\": %.3f, \"Ggathers_per_s\": %.1f}\n",
               words * 8 / 1e6, n, time_ms(e0, e1), n / time_ms(e0, e1) * 1e-6);
        CK(cudaFree(table));
    }

    // ---- smem (own CTA) and dsmem (8-CTA cluster): 128 KB of table per CTA
    {
        const int words = 16384, threads = 1024, iters = 256;
        const size_t smem = (size_t)words * 8;
        CK(cudaFuncSetAttribute(smem_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        CK(cudaFuncSetAttribute(smem_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        // 16 clusters of 8: a GPC holds two of them, so all run in one wave; one CTA per SM (shared memory)
        const int grid = 128;
        smem_kernel<false><<<grid, threads, smem>>>(words, 4, out);
        CK(cudaEventRecord(e0));
        smem_kernel<false><<<grid, threads, smem>>>(words, iters, out);
        CK(cudaEventRecord(e1));
        CK(cudaDeviceSynchronize());
        double n = (double)grid * threads * iters * kUnroll;
        printf("{\"path\": \"smem\", \"table_MB\": %.2f, \"sms_used\": %d, \"gathers\": %.3g, \"ms\": %.3f, \"Ggathers_per_s\": %.1f, \"scaled_to_all_sms\": %.1f}\n",
               words * 8 / 1e6, grid, n, time_ms(e0, e1), n / time_ms(e0, e1) * 1e-6, n / time_ms(e0, e1) * 1e-6 * sms / grid);

        cudaLaunchConfig_t lc = {};
        lc.gridDim = dim3(grid); lc.blockDim = dim3(threads); lc.dynamicSmemBytes = smem;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = 8; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        lc.attrs = attr; lc.numAttrs = 1;
        CK(cudaLaunchKernelEx(&lc, smem_kernel<true>, words, 4, out));
        CK(cudaEventRecord(e0));
        CK(cudaLaunchKernelEx(&lc, smem_kernel<true>, words, iters, out));
        CK(cudaEventRecord(e1));
        CK(cudaDeviceSynchronize());
        printf("{\"path\": \"dsmem\", \"cluster\": 8, \"table_MB\": %.2f, \"sms_used\": %d, \"gathers\": %.3g, \"ms\": %.3f, \"Ggathers_per_s\": %.1f, \"scaled_to_all_sms\": %.1f}\n",
               8 * words * 8 / 1e6, grid, n, time_ms(e0, e1), n / time_ms(e0, e1) * 1e-6, n / time_ms(e0, e1) * 1e-6 * sms / grid);
    }

    // ---- phase 1 in isolation: 200 M entries (one column block of BASELINE config 4), 48 MB table
    {
        const int64_t n = 200LL * 1000 * 1000, groups = n / 4;
        const uint32_t words = 6u << 20;                     // 48 MB of doubles; mask below keeps 32 MB of it hot
        int *idx; double *table, *xg;
        CK(cudaMalloc(&idx, (size_t)n * 4)); CK(cudaMalloc(&table, (size_t)words * 8)); CK(cudaMalloc(&xg, (size_t)n * 8));
        CK(cudaMemset(table, 0, (size_t)words * 8));
        fill_index_kernel<<<sms * 8, 256>>>(idx, n, (4u << 20) - 1);
        auto report = [&](const char *name, int per_thread) {
            CK(cudaDeviceSynchronize());
            printf("{\"path\": \"phase1\", \"variant\": \"%s\", \"groups_per_thread_batch\": %d, \"entries\": %.3g, \"ms\": %.3f, \"Ggathers_per_s\": %.1f}\n",
                   name, per_thread, (double)n, time_ms(e0, e1), (double)n / time_ms(e0, e1) * 1e-6);
        };
        const unsigned short_grid = (unsigned)((groups + 255) / 256);
        phase1_short_kernel<<<short_grid, 256>>>(idx, table, xg, groups);
        CK(cudaEventRecord(e0));
        phase1_short_kernel<<<short_grid, 256>>>(idx, table, xg, groups);
        CK(cudaEventRecord(e1));
        report("short-lived threads (round 1)", 1);
        for (int ctas_per_sm : {4, 8}) {
            const unsigned grid = (unsigned)(sms * ctas_per_sm);
            phase1_persistent_kernel<2><<<grid, 256>>>(idx, table, xg, groups);
            CK(cudaEventRecord(e0));
            phase1_persistent_kernel<2><<<grid, 256>>>(idx, table, xg, groups);
            CK(cudaEventRecord(e1));
            report(ctas_per_sm == 4 ? "persistent, 4 CTAs/SM" : "persistent, 8 CTAs/SM", 2);
            phase1_persistent_kernel<4><<<grid, 256>>>(idx, table, xg, groups);
            CK(cudaEventRecord(e0));
            phase1_persistent_kernel<4><<<grid, 256>>>(idx, table, xg, groups);
            CK(cudaEventRecord(e1));
            report(ctas_per_sm == 4 ? "persistent, 4 CTAs/SM" : "persistent, 8 CTAs/SM", 4);
        }
        CK(cudaFree(idx)); CK(cudaFree(table)); CK(cudaFree(xg));
    }

    // ---- bulk: 16-byte copies from a 32 MB table
    {
        const uint32_t words = 1u << 22;
        double *table;
        CK(cudaMalloc(&table, (size_t)words * 8));
        CK(cudaMemset(table, 0, (size_t)words * 8));
        const int iters = 64, grid = sms * 16, threads = 128;
        bulk_kernel<<<grid, threads>>>(table, words - 1, 2, out);
        CK(cudaEventRecord(e0));
        bulk_kernel<<<grid, threads>>>(table, words - 1, iters, out);
        CK(cudaEventRecord(e1));
        CK(cudaDeviceSynchronize());
        const double n = (double)grid * threads * iters * kUnroll;
        printf("{\"path\": \"bulk16\", \"table_MB\": %.1f, \"copies\": %.3g, \"ms\": %.3f, \"Gcopies_per_s\": %.1f}\n",
               words * 8 / 1e6, n, time_ms(e0, e1), n / time_ms(e0, e1) * 1e-6);
        CK(cudaFree(table));
    }
    return 0;
}
