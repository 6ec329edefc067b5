// barrier.cu -- step barrier between the GPUs of a row-sharded repeated SpMV.
//
// After the fused SpMV+push kernel of step k, every rank must know that all
// peers' pushes into its next-x vector have landed before step k+1 gathers
// from it (and that peers are done reading the vector it is about to
// overwrite).  Instead of a host-driven collective, one warp does it on the
// device: lane p stores the step number into slot [rank] of peer p's flag
// array (peer-mapped HBM over NVLink, system-scope release), then spins on
// its own slot [p] until it shows the step number (system-scope acquire).
// Stream order puts this kernel after the SpMV kernel, whose peer stores are
// complete when it retires.  Each rank runs on its own GPU, so the spinning
// kernels are always co-resident.
#include "common.cuh"

namespace ellspmv {

struct BarrierArgs {
    int rank, nranks;
    long long epoch;
    long long *local_flags;
    long long *peer_flags[kMaxRanks];
    int *error_flag;
};

__global__ void peer_barrier_kernel(const BarrierArgs a)
{
    const int p = threadIdx.x;
    if (p >= a.nranks) return;
    __threadfence_system();
    long long *dst = a.peer_flags[p] + a.rank;
    asm volatile("st.release.sys.global.s64 [%0], %1;" ::"l"(dst), "l"(a.epoch) : "memory");
    const long long *src = a.local_flags + p;
    const long long t0 = clock64();
    long long seen;
    for (;;) {
        asm volatile("ld.acquire.sys.global.s64 %0, [%1];" : "=l"(seen) : "l"(src) : "memory");
        if (seen >= a.epoch) break;
        if (clock64() - t0 > 40000000000LL) {   // ~20 s at 2 GHz: a peer died; do not hang the GPU
            if (a.error_flag) *a.error_flag = 1 + p;
            break;
        }
        __nanosleep(64);
    }
}

cudaError_t launch_peer_barrier(int rank, int nranks, long long epoch, long long *local_flags,
                                long long *const *peer_flags, int *error_flag, cudaStream_t stream)
{
    if (nranks < 1 || nranks > kMaxRanks || rank < 0 || rank >= nranks) return cudaErrorInvalidValue;
    BarrierArgs a = {};
    a.rank = rank;
    a.nranks = nranks;
    a.epoch = epoch;
    a.local_flags = local_flags;
    a.error_flag = error_flag;
    for (int p = 0; p < nranks; p++) a.peer_flags[p] = peer_flags[p];
    peer_barrier_kernel<<<1, 32, 0, stream>>>(a);
    return cudaGetLastError();
}

__global__ void peer_sync_kernel(const StepSync a)
{
    const int p = threadIdx.x;
    if (p >= a.num_peers) return;
    __threadfence_system();
    asm volatile("st.release.sys.global.s64 [%0], %1;" ::"l"(a.peer_flags[p] + a.rank), "l"(a.epoch) : "memory");
    const long long *src = a.local_flags + a.peer_rank[p];
    const long long t0 = clock64();
    long long seen;
    for (;;) {
        asm volatile("ld.acquire.sys.global.s64 %0, [%1];" : "=l"(seen) : "l"(src) : "memory");
        if (seen >= a.epoch) break;
        if (clock64() - t0 > 40000000000LL) {
            if (a.error) *a.error = 1 + a.peer_rank[p];
            break;
        }
        __nanosleep(64);
    }
}

cudaError_t launch_peer_sync(const StepSync &sync, cudaStream_t stream)
{
    if (sync.num_peers <= 0) return cudaSuccess;
    peer_sync_kernel<<<1, 32, 0, stream>>>(sync);
    return cudaGetLastError();
}

// CUDA loads a kernel's code on its first launch, and that load can wait for kernels already
// running in the context.  A hand-shake kernel that is first launched while a peer on the SAME
// device is already spinning for its signal would never get there (two shards of one process on
// one GPU: tests/test_gpu_ell.py::test_exchange_on_one_gpu).  Called at upload time, before
// anything can spin.
cudaError_t preload_sync_kernels()
{
    cudaFuncAttributes attr;
    cudaError_t e = cudaFuncGetAttributes(&attr, peer_barrier_kernel);
    if (e == cudaSuccess) e = cudaFuncGetAttributes(&attr, peer_sync_kernel);
    return e;
}

}  // namespace ellspmv
