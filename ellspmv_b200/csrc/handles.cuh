// handles.cuh -- the opaque handle types behind include/ellspmv_cuda.h, shared by
// api.cu (one GPU) and group.cu (several GPUs driven by one host thread).
#pragma once

#include <vector>

#include "common.cuh"

namespace ellspmv {

struct DeviceGuard {
    int prev = -1;
    bool ok = false;
    explicit DeviceGuard(int dev) {
        if (cudaGetDevice(&prev) != cudaSuccess) { prev = -1; }
        ok = cudaSetDevice(dev) == cudaSuccess;
    }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};


}  // namespace ellspmv

struct ellspmv_cuda_matrix {
    int device = 0;
    ellspmv::EllLayout lay = {};
    int host_idx_bits = 32, dev_idx_bits = 32;
    int64_t num_columns = 0, row_begin = 0, global_rows = 0;
    unsigned flags = 0;
    ellspmv::EllLaunchCfg cfg = {};
    double *vals = nullptr;
    void *cols = nullptr;
    long long *d_minmax = nullptr;
    double *d_ad = nullptr;                  // separately stored diagonal (shard rows), optional
    int sd_order = 0;
    ellspmv::PatternSet pat;                 // offset patterns of the index stream (pattern.cu), optional
    ellspmv::SgMatrix *sg = nullptr;         // staged-gather copy (ELLSPMV_CUDA_STAGED_GATHER or KERNEL_AUTO's choice), optional
    int *d_rowlen = nullptr;                 // per-row lengths (CSR view built by csrspmv_cuda_*), optional
    int staged_mode = 0;                     // 0 direct gather, 1 staged on request, 2 staged chosen by the timed trial
    double tune_ms[2] = {0.0, 0.0};          // KERNEL_AUTO's trial at upload: direct, staged
    bool kernel_auto = false;                // the caller left the kernel choice to the library
    bool rpt_auto = false;                   // ... and the rows per thread
    ellspmv::CbMatrix *cb = nullptr;         // column-blocked copy (ELLSPMV_CUDA_COLUMN_BLOCKED), optional
    ellspmv::SellMatrix *sell = nullptr;     // SELL-128-sigma copy without the trailing padding (ELLSPMV_CUDA_SKIP_PADDING), optional
    int64_t min_col = 0, max_col = -1;
    unsigned char *d_remote = nullptr;       // per slice: reads columns outside the shard's rows (fused step sync)
    unsigned *d_done = nullptr;              // completion counter of the fused step sync
    std::vector<unsigned char> h_remote;     // host copy of d_remote
    ellspmv::PushTargets sync_plan = {};     // the push ranges the boundary description below was made for
    unsigned sync_total = 0;                 // warps of the boundary slices (push or halo) under that plan
    int sync_num_ranges = 0;                 // boundary slices as <= 4 index ranges, or -1: per-slice table d_boundary
    long long sync_range_lo[4] = {0, 0, 0, 0}, sync_range_hi[4] = {0, 0, 0, 0};
    unsigned char *d_boundary = nullptr;
    int64_t sync_boundary_slices = 0;
    bool sync_plan_valid = false;
    cudaStream_t side = nullptr;             // the step hand-shake runs here, next to the interior slices
    cudaEvent_t ev_start = nullptr, ev_boundary = nullptr, ev_handshake = nullptr;
    bool handshake_pending = false;
    cudaStream_t stream = nullptr;
    cudaStream_t stream_out = nullptr;       // D2H stream of the pipelined host call
    cudaStream_t stream_in = nullptr;        // H2D stream of the pipelined host call
    double *d_x = nullptr, *d_y = nullptr;   // vectors of the host-facing spmv
    std::vector<long long> chunk_max;        // pipelined host call: largest column each row chunk references
    int64_t vec_len = 0;
    std::vector<cudaEvent_t> events;
    int64_t device_bytes = 0;
    int64_t launches = 0;

    // ---- group handle (num_gpus > 1): one shard handle per GPU ----------------
    std::vector<ellspmv_cuda_matrix *> shards;
    std::vector<double *> xb[2];            // per device: two full-length vectors (current / next x)
    std::vector<long long *> bflags;        // per device: 32 int64 barrier flags
    std::vector<std::vector<cudaEvent_t>> gevents;
    long long epoch = 0;
    int64_t vec_elems = 0;
};

struct csrspmv_cuda_matrix {
    int device = 0;
    int idx_bits = 32;
    int64_t num_rows = 0, num_columns = 0, csrsize = 0;
    int64_t row_begin = 0;                     // global index of row 0 (shard of a group handle)
    unsigned flags = 0;
    int kernel = ELLSPMV_CUDA_KERNEL_THREAD;   // 1 stream, 2 vector, 3 scalar (see csr_kernels.cu)
    bool auto_kernel = true;                   // pick scalar vs stream from the row lengths
    int64_t max_row_len = 0;
    bool fma = false;
    int64_t *rowptr = nullptr;
    void *cols = nullptr;
    double *vals = nullptr;
    cudaStream_t stream = nullptr;
    double *d_x = nullptr, *d_y = nullptr;
    double *d_ad = nullptr;                  // separately stored diagonal, optional
    unsigned long long *d_scratch = nullptr; // 32 bytes for the upload-time inspection (csr_inspect)
    ellspmv_cuda_matrix *ell = nullptr;      // sliced-ELL view of the same entries (KERNEL_AUTO, balanced rows): the
                                             //   launches go through the ELL kernels with per-row lengths
    int64_t min_row_len = 0;
    ellspmv::SellMatrix *sell = nullptr;     // SELL-128-sigma copy (KERNEL_AUTO, unbalanced rows; or kernel selector 5)
    int64_t min_col = 0, max_col = -1;       // range of the stored column indices
    std::vector<cudaEvent_t> events;
    int64_t device_bytes = 0;

    // ---- group handle (num_gpus > 1): nnz-balanced contiguous row blocks, one per GPU ----
    std::vector<csrspmv_cuda_matrix *> shards;
    std::vector<int64_t> row_lo;             // first row of every shard (+ num_rows at the end)
};


namespace ellspmv {
// group.cu
int group_upload(ellspmv_cuda_matrix **out, int idx_width_bits, int64_t num_rows, int64_t num_columns,
                 int64_t rowsize, const void *colidx, const double *a, int num_gpus, unsigned flags);
int group_generate(ellspmv_cuda_matrix **out, int kind, const int64_t dims[3], const double vals[2],
                   uint64_t seed, int idx_width_bits, int num_gpus, unsigned flags);
int group_spmv(ellspmv_cuda_matrix *G, double *y, const double *x, int repeat, int mode, double *seconds);
int group_download(const ellspmv_cuda_matrix *G, void *colidx, double *a);
int group_info(const ellspmv_cuda_matrix *G, ellspmv_cuda_info *info);
void group_free(ellspmv_cuda_matrix *G);
int csr_group_upload(csrspmv_cuda_matrix **out, int idx_width_bits, int64_t num_rows, int64_t num_columns,
                     const int64_t *rowptr, const void *colidx, const double *a, int num_gpus, unsigned flags);
int csr_group_spmv(csrspmv_cuda_matrix *G, double *y, const double *x, int repeat, int mode, double *seconds);
void csr_group_free(csrspmv_cuda_matrix *G);
int csr_upload_on(csrspmv_cuda_matrix **out, int idx_width_bits, int64_t num_rows, int64_t num_columns,
                  const int64_t *rowptr, const void *colidx, const double *a, int device, unsigned flags,
                  int64_t row_begin = 0);
// api.cu internals the group needs
int launch_shard(ellspmv_cuda_matrix *A, double *y_dev, const double *x_dev, int beta,
                 const PushTargets *push, cudaStream_t stream);
// fused SpMV + push + step signalling (falls back to push + peer_sync kernel where the kernel in
// use does not carry the fused form); sync.done / sync.remote are filled in from the handle
int launch_shard_exchange(ellspmv_cuda_matrix *A, double *y_dev, const double *x_dev, int beta,
                          const PushTargets *push, StepSync sync, cudaStream_t stream);
int ensure_event_count(std::vector<cudaEvent_t> &ev, size_t n);
void shard_x_range(const ellspmv_cuda_matrix *A, int64_t *lo, int64_t *hi);   // the part of x a shard's kernels read
int csr_launch(csrspmv_cuda_matrix *A, double *y_dev, const double *x_dev, int beta, cudaStream_t stream);
void csr_x_range(const csrspmv_cuda_matrix *A, int64_t *lo, int64_t *hi);
}  // namespace ellspmv
