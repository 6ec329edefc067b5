// common.cuh -- shared declarations of the sm_100a SpMV library.
//
// Internal header (C++/CUDA).  The public boundary is the C ABI in
// include/ellspmv_cuda.h; nothing here is exported.
#pragma once

#include <cuda_runtime.h>
#include <errno.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/ellspmv_cuda.h"

namespace ellspmv {

// ---- error plumbing -----------------------------------------------------
// The C ABI returns errno-style ints like the reference's kernels do
// (ellspmv.c:1197 returns EINVAL); the text of the underlying CUDA error is
// kept per host thread for ellspmv_cuda_last_error().
void set_last_error(const char *fmt, ...);
int  cuda_to_errno(cudaError_t e);

#define ELL_CK(call)                                                        \
    do {                                                                    \
        cudaError_t e__ = (call);                                           \
        if (e__ != cudaSuccess) {                                           \
            ::ellspmv::set_last_error("%s:%d: %s -> %s", __FILE__, __LINE__, \
                                      #call, cudaGetErrorString(e__));      \
            return ::ellspmv::cuda_to_errno(e__);                           \
        }                                                                   \
    } while (0)

#define ELL_FAIL(err, ...)                       \
    do {                                         \
        ::ellspmv::set_last_error(__VA_ARGS__);  \
        return (err);                            \
    } while (0)

constexpr int kBlockThreads = 128;  // threads per CTA of the thread-per-row kernels
constexpr int kMaxPeers = 8;
constexpr int kMaxPatterns = 16;    // offset-pattern dictionary size, one id per group (pattern.cu)
constexpr int kMaxLanePatterns = 32; // ... one id per thread (the dictionary holds boundary rows' vectors too)

// ---- sliced-ELL device layout -------------------------------------------
// Rows are grouped into slices of S = kBlockThreads * R rows (R = rows per
// thread).  Inside a slice the K slots are stored slot-major: element
// (row r, slot l) of slice s lives at  s*S*K + l*S + (r mod S).  A CTA owns
// one slice; for a fixed slot its threads read one contiguous run of S
// values (S*8 bytes) and S indices, so every load is a fully coalesced
// 128/256-bit vector load, and a slice is one contiguous S*K*(8+idx) byte
// region of HBM.  Rows past num_rows in the last slice hold (col 0, 0.0).
struct EllLayout {
    int64_t num_rows;     // rows in this shard
    int64_t num_slices;   // ceil(num_rows / slice_rows)
    int     rowsize;      // K
    int     slice_rows;   // S
    __host__ __device__ int64_t padded_rows() const { return num_slices * (int64_t)slice_rows; }
    __host__ __device__ int64_t entries() const { return padded_rows() * rowsize; }
    __host__ __device__ int64_t offset(int64_t row, int slot) const {
        int64_t s = row / slice_rows;
        int64_t r = row - s * slice_rows;
        return (s * rowsize + slot) * slice_rows + r;
    }
};

// ---- fused exchange targets ---------------------------------------------
struct PushTargets {
    int      num_peers;
    double  *x[kMaxPeers];       // peer vectors, indexed by GLOBAL row
    int64_t  row_lo[kMaxPeers];  // global row range the peer needs
    int64_t  row_hi[kMaxPeers];
};

// ---- fused step synchronisation of the row-sharded y -> x loop (ell_kernels.cu) ----------
// Instead of a barrier kernel between two steps, the SpMV kernel itself signals and waits:
// its last warp to finish stores the step number into slot [rank] of every listed rank's flag
// array (peer-mapped HBM, system-scope release), and only the warps of CTAs that read halo
// columns or push to a peer wait -- at their start -- until the listed ranks have signalled the
// previous step.  Interior CTAs never wait, so the flag round trip over NVLink hides behind
// them.  Everything is per warp (no CTA barrier): threads past the last row can simply exit.
struct StepSync {
    long long *local_flags;            // this rank's flag array, slot [q] = last step rank q finished; NULL = off
    long long *peer_flags[kMaxPeers];  // flag arrays of the ranks below (peer-mapped)
    int        peer_rank[kMaxPeers];   // the ranks this one exchanges with (pushes to or is pushed by)
    int        num_peers;
    int        rank;
    long long  epoch;                  // this step's number (>= 1): wait for epoch-1, signal epoch
    unsigned  *done;                   // boundary warps finished so far (device memory, zero between launches)
    unsigned   total_warps;            // warps (with at least one row) of the boundary slices
    // boundary slices = those that push to a peer or read columns outside the (16-aligned) local row
    // range.  Given as up to 4 ranges of slice indices (a stencil shard: its first and last few
    // slices), so that an interior CTA decides with a few compares on kernel parameters -- a
    // per-slice table lookup put a dependent global load at the head of every warp and cost 10 %
    // of a step.  num_ranges < 0: more than 4 runs, look slice s up in `table`.
    int        num_ranges;
    long long  range_lo[4], range_hi[4];
    const unsigned char *table;
    int       *error;                  // set when a peer never showed up
    unsigned   poll_ns;                // back-off between two looks at a flag
    int        debug_nowait;           // experiments only (ELLSPMV_CUDA_SYNC_NOWAIT): do not wait -- results undefined
};

struct EllSpmvArgs {
    const double *vals;     // sliced layout
    const void   *cols;     // sliced layout, int32 or int64
    const double *x;        // num_columns
    double       *y;        // shard rows (in/out)
    int64_t       num_rows; // shard rows
    int64_t       row_begin;// global index of shard row 0 (for push)
    int64_t       slice_begin; // first slice of this launch (chunked launches of the pipelined host call)
    int           rowsize;
    int           beta;     // 1: y += A x, 0: y = A x
    const double *ad;       // separately stored diagonal of the shard rows, or NULL
    int           sd_order; // 0: y += ad*x + yi (ellgemvsd); 1: sum starts at ad*x (ellgemv16sd)
    PushTargets   push;
    int           prefetch;     // slices ahead whose value stream is requested into L2 (0 = none)
    const unsigned char *patid; // offset patterns (pattern.cu): one id per warp (32*R rows), 0xff = explicit indices; or NULL
    const unsigned long long *patinfo; // with lane masks (opt-in) instead: low byte = pattern id, high half = lanes whose rows deviate
    const unsigned char *patlane; // or one id per THREAD (R rows): all 0xff or none in a warp (pattern.cu, lane patterns)
    const long long     *pat;   // dictionary [kMaxPatterns][K] of column offsets relative to the GLOBAL row
    const double        *vpat;  // value patterns: the dictionary entry's K coefficients too (then the value stream of a patterned thread is not read), or NULL
    StepSync      sync;
    const int    *rowlen;   // per row: how many leading slots count (CSR view: the rest is never touched
                            //   arithmetically, so no 0*inf from a padded slot); NULL = all K
};

struct EllLaunchCfg {
    int  idx_bits;         // 32 / 64 (device storage)
    int  rows_per_thread;  // 1, 2, 4
    int  kernel;           // ELLSPMV_CUDA_KERNEL_THREAD / _WARP
    int  variant;          // 0 direct loads, 1 bulk-async staged
    bool fma;
    bool persist_x;        // attach an L2 access-policy window over x
    int64_t x_bytes;
    int  num_sms;
};

cudaError_t launch_ell_spmv(const EllLaunchCfg &cfg, const EllSpmvArgs &args,
                            int64_t num_slices, cudaStream_t stream);
// few, long rows: CTA per row group on the row-major layout (slice height 1), ell_longrow.cu
cudaError_t launch_ell_longrow(const EllLaunchCfg &cfg, const EllSpmvArgs &args, cudaStream_t stream);
// the thread-per-row kernel with per-row lengths (the CSR view), ell_kernels_len.cu
cudaError_t launch_ell_thread_len(const EllLaunchCfg &cfg, const EllSpmvArgs &args, bool yvec, cudaLaunchConfig_t &lc);
constexpr int kKernelLongRow = 4;   // ELLSPMV_CUDA_KERNEL_LONGROW
// persistent bulk-async (TMA) staged variant, ell_bulk.cu; *handled = false: not applicable
cudaError_t launch_ell_bulk(const EllLaunchCfg &cfg, const EllSpmvArgs &args, int64_t num_slices,
                            cudaStream_t stream, bool *handled);

// ---- CSR ---------------------------------------------------------------
struct CsrSpmvArgs {
    const int64_t *rowptr;
    const void    *cols;
    const double  *vals;
    const double  *x;
    double        *y;
    int64_t        num_rows;
    int            beta;
    const double  *ad;      // separately stored diagonal of these rows, or NULL (csrgemvsd)
    int64_t        row_begin; // global index of row 0 (a row block of a larger matrix): csrgemvsd reads x[global row]
};
cudaError_t launch_csr_spmv(int idx_bits, bool fma, int kernel, const CsrSpmvArgs &args,
                            cudaStream_t stream);
struct CsrInspection { int64_t max_row_len, min_row_len, bad_rows, min_col, max_col; };
cudaError_t csr_inspect(int idx_bits, const int64_t *rowptr, const void *cols, int64_t num_rows, int64_t csrsize,
                        unsigned long long *scratch /* device, 40 bytes */, CsrInspection *res, cudaStream_t stream);

// ---- layout / generators (layout.cu) -------------------------------------
// row-major chunk (rows [row0, row0+rows) of the shard) -> sliced layout
cudaError_t relayout_chunk(int src_idx_bits, int dst_idx_bits, const void *src_cols,
                           const double *src_vals, void *dst_cols, double *dst_vals,
                           const EllLayout &lay, int64_t row0, int64_t rows,
                           long long *minmax /* device [2] */, cudaStream_t stream);
// sliced layout -> row-major chunk
cudaError_t unlayout_chunk(int dev_idx_bits, int host_idx_bits, const void *src_cols,
                           const double *src_vals, void *dst_cols, double *dst_vals,
                           const EllLayout &lay, int64_t row0, int64_t rows,
                           cudaStream_t stream);
cudaError_t generate_sliced(int kind, const int64_t dims[3], const double vals[2], uint64_t seed,
                            int dst_idx_bits, void *dst_cols, double *dst_vals,
                            const EllLayout &lay, int64_t row_begin,
                            long long *minmax, cudaStream_t stream);
cudaError_t generate_csr_random(const int64_t dims[3], uint64_t seed, int idx_bits,
                                int64_t *rowptr, void *cols, double *vals, cudaStream_t stream);
// CSR form of the laplace2d / stencil27 generators (no padding, entries in the ELL order)
int64_t csr_stencil_nnz(int kind, const int64_t dims[3]);
cudaError_t generate_csr_stencil(int kind, const int64_t dims[3], const double vals[2], int idx_bits,
                                 int64_t *rowptr, void *cols, double *out_vals, cudaStream_t stream);
cudaError_t init_minmax(long long *minmax, cudaStream_t stream);
// CSR rows -> sliced-ELL layout of width lay.rowsize (>= the longest row): entries keep their order,
// the unused slots of a row get (its last column or 0, 0.0), rowlen[r] = the row's length
cudaError_t csr_to_sliced(int src_idx_bits, int dst_idx_bits, const int64_t *rowptr, const void *src_cols,
                          const double *src_vals, void *dst_cols, double *dst_vals, int *rowlen,
                          const EllLayout &lay, cudaStream_t stream);
cudaError_t chunk_max_cols(int idx_bits, const void *cols, const EllLayout &lay, int64_t chunk_slices, int nchunks,
                           long long *d_out /* device, nchunks */, cudaStream_t stream);
cudaError_t mark_remote_slices(int idx_bits, const void *cols, const EllLayout &lay, int64_t lo, int64_t hi,
                               unsigned char *remote, cudaStream_t stream);

// ---- offset patterns: groups of 32 rows whose column indices are row + d[l] (pattern.cu) ----
struct PatternSet {
    unsigned char *patid = nullptr;   // device: padded_rows / 32 ids
    unsigned long long *patinfo = nullptr; // device: per group, id (low byte) | lanes that keep explicit indices << 32
    int64_t explicit_lanes = 0;       // lanes flagged in the masks of the patterned groups
    int max_explicit = 0;             // 0: whole groups only (patinfo unused by the kernel)
    unsigned char *patlane = nullptr; // device: padded_rows / R ids, one per thread of the thread-per-row kernel (lane patterns)
    long long *pat = nullptr;         // device: kMaxLanePatterns * K offsets
    double *vpat = nullptr;           // device: kMaxLanePatterns * K coefficients (value patterns), or NULL
    int num_patterns = 0;
    int group_rows = 32;              // 32 * rows per thread
    int64_t groups = 0, covered = 0;  // groups: all / patterned
    int64_t bytes = 0;
    bool any() const { return patid || patlane; }
};
// max_explicit: lanes of a patterned group that may deviate and keep explicit indices (0 = whole groups only)
// lanes: also try one pattern id per thread and keep it when it saves more index bytes than the ids cost
// vals: also look for value patterns (rows that share offsets AND coefficients), or NULL
cudaError_t pattern_build(PatternSet *ps, int idx_bits, const void *cols, const double *vals, const EllLayout &lay,
                          int rows_per_thread, int64_t row_begin, int max_explicit, bool lanes, cudaStream_t stream);
void pattern_free(PatternSet *ps);

// ---- column-blocked ELL (ell_blocked.cu) ----------------------------------------
struct CbMatrix;
cudaError_t cb_build(CbMatrix **out, int idx_bits, const double *vals, const void *cols, const EllLayout &lay,
                     int64_t num_columns, int64_t target_x_bytes, cudaStream_t stream);
cudaError_t cb_spmv(const CbMatrix *cb, bool fma, const double *x, double *y, int64_t num_rows, int64_t num_columns,
                    int beta, cudaStream_t stream);
void cb_free(CbMatrix *cb);
int64_t cb_bytes(const CbMatrix *cb);
int cb_blocks(const CbMatrix *cb);
int64_t cb_entries(const CbMatrix *cb);

// ---- column-blocked, staged gather: bit-exact (ell_staged.cu) ---------------------
struct SgMatrix;
cudaError_t sg_build(SgMatrix **out, int idx_bits, const void *cols, const double *vals, const EllLayout &lay,
                     int64_t num_columns, int64_t target_x_bytes, cudaStream_t stream);
cudaError_t sg_spmv(const SgMatrix *sg, const double *x, double *y, const double *ad,
                    int sd_order, int64_t num_rows, int64_t row_begin, int beta, const PushTargets *push,
                    cudaStream_t stream, const int *rowlen = nullptr);
cudaError_t sg_scatter_estimate(int idx_bits, const void *cols, const EllLayout &lay, double *lines_per_gather,
                                cudaStream_t stream);
int64_t sg_bytes_estimate(int idx_bits, const EllLayout &lay);
void sg_free(SgMatrix *sg);
int64_t sg_bytes(const SgMatrix *sg);
int sg_launches(const SgMatrix *sg);

// ---- SELL-128-sigma: per-slice widths, rows sorted by length in windows (sell.cu) --------------
struct SellMatrix;
cudaError_t sell_build_csr(SellMatrix **out, int src_idx_bits, int dst_idx_bits, int64_t num_rows, const int64_t *rowptr,
                           const void *cols, const double *vals, cudaStream_t stream);
cudaError_t sell_build_ell(SellMatrix **out, int idx_bits, const void *cols, const double *vals, const EllLayout &lay,
                           int64_t row_begin, int64_t num_columns, cudaStream_t stream);
cudaError_t sell_spmv(const SellMatrix *m, bool fma, const int64_t *csr_rowptr, const void *csr_cols,
                      const double *csr_vals, int csr_idx_bits, const double *x, double *y,
                      const double *ad, int64_t row_begin, int beta, cudaStream_t stream);
void sell_free(SellMatrix *m);
int64_t sell_bytes(const SellMatrix *m);
int64_t sell_entries(const SellMatrix *m);        // stored slots
int64_t sell_real_entries(const SellMatrix *m);   // slots that count
int64_t sell_long_rows(const SellMatrix *m);
int sell_launches(const SellMatrix *m);
int sell_long_row(const SellMatrix *m);       // rows longer than this run one CTA each (0: none)

// ---- COO -> ELL / CSR on the device (convert.cu) ------------------------------
struct CooEllJob {
    int idx_bits = 32;
    int64_t nnz = 0, num_rows = 0, num_columns = 0;
    const void *d_rowidx = nullptr, *d_colidx = nullptr;   // 1-based, device copies of the file's arrays
    const double *d_a = nullptr;
    void *state = nullptr;                                 // sort workspace, owned by convert.cu
    int64_t rowsize = 0;                                   // K, known after phase 1
};
cudaError_t coo_to_ell_phase1(CooEllJob &job, cudaStream_t stream, int *bad);
cudaError_t coo_to_ell_phase2(CooEllJob &job, int dst_idx_bits, void *dst_cols, double *dst_vals,
                              const EllLayout &lay, long long *minmax, cudaStream_t stream, int *bad);
void coo_to_ell_release(CooEllJob &job);
cudaError_t coo_to_csr(int idx_bits, const void *d_rowidx, const void *d_colidx, const double *d_a, int64_t nnz,
                       int64_t num_rows, int64_t num_columns, int64_t *rowptr, void *csrcolidx, double *csra,
                       cudaStream_t stream, int *bad);

// ---- cross-GPU step barrier (barrier.cu) ----------------------------------
// rank writes `epoch` into slot [rank] of every peer's flag array, then waits
// until its own array shows `epoch` from every rank.  Flags are int64 in
// peer-mapped device memory (one array of kMaxRanks slots per rank).
constexpr int kMaxRanks = 16;
cudaError_t launch_peer_barrier(int rank, int nranks, long long epoch, long long *local_flags,
                                long long *const *peer_flags, int *error_flag, cudaStream_t stream);
// the same between a rank and the ranks it exchanges with only (the protocol of StepSync, for the
// kernels that do not carry the fused form): signal `epoch` to them, wait for `epoch` from them
cudaError_t launch_peer_sync(const StepSync &sync, cudaStream_t stream);
cudaError_t preload_sync_kernels();      // load the hand-shake kernels' code now (see barrier.cu)

}  // namespace ellspmv
