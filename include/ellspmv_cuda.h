/*
 * ellspmv_cuda.h -- C ABI of the B200 (sm_100a) SpMV library.
 *
 * This is the drop-in boundary for the hot path of jamtrott/ellspmv: the
 * fp64 ELLPACK kernel `ellgemv` (reference ellspmv.c:1129-1153) and, as the
 * comparison path, the CSR kernel `csrgemv` (reference csrspmv.c:1565-1595).
 * The reference has no FFI of its own; the entry points below are what a
 * maintainer binds at the reference's three kernel call sites
 * (ellspmv.c:1766-1767 warm-up, 1841-1842 timed; csrspmv.c:2766-2767,
 * 2857-2858).  See INTEGRATION.md for the exact patch.
 *
 * Conventions (mirroring the reference):
 *   - every function returns 0 on success or a positive errno value
 *     (EINVAL, ENOMEM, ENODEV, EIO ...), like `ellgemv` does
 *     (ellspmv.c:1197); ellspmv_cuda_last_error() has the detail;
 *   - index width is a run-time argument here (32 or 64 bits) where the
 *     reference fixes idx_t at compile time (ellspmv.c:112-130);
 *   - host arrays stay owned by the caller; `upload` copies, the handle
 *     owns all device memory, `free` releases it;
 *   - calls on one handle must come from one host thread at a time (the
 *     reference calls its kernel from every OpenMP thread; call this from
 *     the master thread only).
 *
 * There is no CPU fallback: without a CUDA device every compute entry
 * point fails with ENODEV.
 */
#ifndef ELLSPMV_CUDA_H
#define ELLSPMV_CUDA_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ELLSPMV_CUDA_VERSION 100 /* 1.00 */

typedef struct ellspmv_cuda_matrix ellspmv_cuda_matrix; /* opaque, ELL  */
typedef struct csrspmv_cuda_matrix csrspmv_cuda_matrix; /* opaque, CSR  */

/* ---- flags for upload/generate (OR together) ------------------------- */
enum {
    /* kernel selection (low 4 bits) */
    ELLSPMV_CUDA_KERNEL_AUTO   = 0, /* ELL: by nnz-per-row and row count --      */
                                    /*   thread-per-row on the sliced layout     */
                                    /*   (fully coalesced at any K), the long-row*/
                                    /*   kernel for few long rows; both bit-exact*/
                                    /*   CSR: see CSRSPMV_CUDA_KERNEL_SCALAR     */
    ELLSPMV_CUDA_KERNEL_THREAD = 1, /* thread-per-row, sequential slot order:   */
                                    /*   bit-exact with the reference loop      */
    ELLSPMV_CUDA_KERNEL_WARP   = 2, /* sub-warp-per-row + shuffle reduction:    */
                                    /*   tolerance mode (summation order differs)*/
    CSRSPMV_CUDA_KERNEL_SCALAR = 3, /* CSR only: thread-per-row, bit-exact.  CSR AUTO:     */
                                    /*   balanced rows (padding <= 25 %) run through a     */
                                    /*   sliced-ELL view with per-row lengths (the ELL     */
                                    /*   kernels, same bits as csrgemv); otherwise scalar  */
                                    /*   or the smem-staged stream kernel (KERNEL_THREAD)  */
                                    /*   by row raggedness.  With ELLSPMV_CUDA_FMA the     */
                                    /*   stream kernel cannot contract (it parks rounded   */
                                    /*   products): AUTO then never picks it, so the bits  */
                                    /*   do not depend on the row lengths                  */
    ELLSPMV_CUDA_KERNEL_LONGROW = 4,/* ELL: few, long rows -- a CTA per small group of rows  */
                                    /*   on the reference's own row-major layout: eight loader */
                                    /*   warps stream and gather (cp.async ring), the rounded  */
                                    /*   products are parked in shared memory, one lane per    */
                                    /*   row of a summing warp adds them in slot order:        */
                                    /*   bit-exact.  AUTO takes it for rowsize >= 64 on        */
                                    /*   matrices of at most 32768 rows       */
    CSRSPMV_CUDA_KERNEL_SELL   = 5, /* CSR only: SELL-128-sigma (sell.cu); AUTO takes it for    */
                                    /*   unbalanced rows                                        */
    ELLSPMV_CUDA_KERNEL_MASK   = 0xf,
    /* arithmetic: default is mul-then-add (__dmul_rn/__dadd_rn), the bits the
     * reference's compiled loop produces; FMA allows contraction (tolerance) */
    ELLSPMV_CUDA_FMA           = 1 << 4,
    /* ask for an L2 persisting access-policy window over x when it fits */
    ELLSPMV_CUDA_L2_PERSIST_X  = 1 << 5,
    /* 64-bit column indices are kept as 32-bit on the device whenever
     * num_columns < 2^31: a device-layout choice, results and
     * ellspmv_cuda_download() stay 64-bit and bit-exact, the index stream
     * halves.  This is the default; NARROW_INDEX is accepted for
     * compatibility, WIDE_INDEX keeps the caller's width on the device. */
    ELLSPMV_CUDA_NARROW_INDEX  = 1 << 6,
    ELLSPMV_CUDA_WIDE_INDEX    = 1 << 16,
    /* bin the entries by column block so that each block's slice of x
     * (<= 48 MB) stays in L2, and run y += A_b*x block after block: for
     * matrices with scattered columns and an x larger than L2 (BASELINE
     * config 4).  Tolerance mode: inside a block the reference's order is
     * kept, but the per-block partial sums are added block by block; stored
     * zeros are dropped.  Keeps the regular layout as well (download, push
     * and separate-diagonal calls use it). */
    ELLSPMV_CUDA_COLUMN_BLOCKED = 1 << 7,
    /* the bit-exact way to block by columns: store the entries a second time
     * sorted by (column block, slice); per SpMV, walk them block after block
     * (each block's slice of x stays in L2), gather x, multiply, and park the
     * ROUNDED products as a flat stream in HBM; then add each row's products in
     * slot order 0..K-1 (bulk-async copies into shared memory) -- mul, then
     * left-to-right adds: the reference's bits, ~30 B instead of ~110 B of DRAM
     * traffic per stored entry when x is much larger than L2; costs idx + 18
     * bytes of device memory per entry.  KERNEL_AUTO takes this path by itself
     * after a timed trial (see NO_STAGED_GATHER).  No-op when x fits one block or
     * a slice's K*128 staged products do not fit shared memory.  With
     * ELLSPMV_CUDA_FMA the products are still parked rounded (no contraction). */
    ELLSPMV_CUDA_STAGED_GATHER  = 1 << 17,
    /* offset patterns (on by default, thread-per-row kernel with one row
     * per thread): at upload, groups of 32 consecutive rows whose column
     * indices are all row + d[l] for one offset vector d (structured grids:
     * nearly every row) are found and verified entry by entry; the kernel
     * then computes those columns from a 16-entry dictionary of offset
     * vectors instead of streaming them from HBM.  A device-layout choice
     * like the index narrowing: same columns, same bits, download()
     * unchanged; only the index bytes of patterned rows are no longer read.
     * NO_PATTERN keeps every group on the explicit index stream. */
    ELLSPMV_CUDA_NO_PATTERN     = 1 << 18,
    /* KERNEL_AUTO builds the staged gather by itself for matrices whose x is
     * larger than L2, whose rows follow no offset pattern and whose gathers are
     * scattered (a warp's 32 columns fall into more than 16 different 128-byte
     * lines on average), times both paths once at upload and keeps the faster
     * one -- the results are the same bits either way.  NO_STAGED_GATHER keeps
     * the direct gather without a trial. */
    ELLSPMV_CUDA_NO_STAGED_GATHER = 1 << 19,
    /* SELL-128-sigma for ELL matrices with ragged rows (opt-in): the reference pads every
     * row to the longest one (ellspmv.c:944-955, 1111-1117) and streams the padding; with
     * this flag the handle also keeps the rows sorted by their length without trailing
     * padding (windows of 4096 rows), sliced with a width per slice, and runs
     * y += A*x from that copy -- only the slots that count are streamed, in the
     * reference's order.  Bit-exact for finite x; where x is inf/NaN on a row's PADDING
     * column the reference produces 0*inf = NaN and this path does not: hence opt-in.
     * (The CSR path uses the same layout by itself for unbalanced rows: csrgemv has no
     * padding, so there it is exact for every x.) */
    ELLSPMV_CUDA_SKIP_PADDING     = 1 << 20,
    /* offset patterns with lane masks: a group of 32*R rows stays on its pattern when at
     * most 4 of its lanes hold deviating rows (a grid boundary); those rows are recomputed
     * from the explicit indices by the whole warp.  Raises the coverage (27-point 384^3:
     * 83 % -> 99.5 % of the rows) but measured slower than whole groups on the BASELINE
     * shapes (profiles/r2_offset_patterns.md): opt-in.  Same bits either way. */
    ELLSPMV_CUDA_PATTERN_MASKS    = 1 << 21,
    /* ellspmv_cuda_spmv_exchange: do the step hand-shake INSIDE the SpMV kernel (boundary
     * warps wait at their start and signal at their end) instead of in a one-warp kernel
     * after it.  Saves a launch per step, but the SpMV instantiation that carries it
     * measured 10 % slower than kernel + hand-shake kernel on the 8192^2 shard
     * (profiles/r2_scaling.md): opt-in.  Same protocol, same bits. */
    ELLSPMV_CUDA_FUSED_SYNC       = 1 << 22,
    /* lane patterns: where whole groups leave index bytes on the table (a grid boundary
     * puts one deviating row into every 12th group of the 27-point 384^3 stencil and voids
     * it), the upload also tries ONE PATTERN ID PER THREAD instead of one per group: the
     * boundary row's own offset vector goes into the (32-entry) dictionary like any other
     * and the row stays in the warp of its interior neighbours.  Kept when the index bytes
     * saved exceed twice the byte per thread the ids cost (27-point 384^3: 83 % -> 99.9 %
     * of the rows; 2D 5-point 8192^2: not worth it, stays on group ids).  Same kernel
     * otherwise, same bits.  NO_PATTERN_LANES keeps group ids. */
    ELLSPMV_CUDA_NO_PATTERN_LANES = 1 << 23,
    /* value patterns (OPT-IN): in a constant-coefficient stencil the rows share not only their
     * column offsets but their COEFFICIENTS (2D 5-point Laplacian: 4, -1, -1, -1, -1 in every
     * interior row).  With this flag the upload looks for that too -- the pattern signature then
     * includes the bit patterns of the values, every row is verified entry by entry, bit for bit
     * -- and a dictionary entry carries offsets and coefficients: a patterned thread streams
     * neither indices nor values from HBM; what is left is the x gather and the y store (the
     * matrix-free limit of the same arithmetic).  Taken when it covers at least 9/10 of the rows
     * the index-only search covers; a matrix with variable coefficients keeps index patterns and
     * its value stream.  Same arithmetic on the same numbers in the same order: same bits.  Not
     * searched with ELLSPMV_CUDA_FMA.  Off by default, and never set by bench.py's headline:
     * constant coefficients are a property of synthetic matrices, and a benchmark that does not
     * read A no longer measures an SpMV. */
    ELLSPMV_CUDA_VALUE_PATTERN    = 1 << 24,
    /* rows handled per thread in the thread-per-row kernel (1, 2 or 4):
     * 0 = auto = 2 for rows of at most 12 entries, else 1 */
    ELLSPMV_CUDA_ROWS_PER_THREAD_SHIFT = 8,
    ELLSPMV_CUDA_ROWS_PER_THREAD_MASK  = 0x7 << 8,
    /* kernel variant for experiments (0 = default direct loads,
     * 1 = bulk-async (TMA) staged through shared memory) */
    ELLSPMV_CUDA_VARIANT_SHIFT = 12,
    ELLSPMV_CUDA_VARIANT_MASK  = 0xf << 12
};

/* ---- modes for spmv -------------------------------------------------- */
enum {
    ELLSPMV_CUDA_ACCUMULATE = 0, /* y <- y + A*x, `repeat` times (reference semantics,   */
                                 /*   x constant: ellspmv.c:1150, 1824-1843)             */
    ELLSPMV_CUDA_OVERWRITE  = 1, /* y <- A*x                                             */
    ELLSPMV_CUDA_ITERATE    = 2  /* x_{k+1} <- A*x_k, `repeat` times, result in y        */
                                 /*   (BASELINE config 5; square A only)                 */
};

/* ---- synthetic matrix kinds for generate ------------ */
enum {
    ELLSPMV_CUDA_GEN_LAPLACE2D = 1, /* dims = {nx, ny, -}, K = 5,  vals = {centre, off} */
    ELLSPMV_CUDA_GEN_STENCIL27 = 2, /* dims = {nx, ny, nz}, K = 27, vals = {centre, off} */
    ELLSPMV_CUDA_GEN_RANDOM    = 3  /* dims = {rows, cols, K}, seeded, vals unused       */
};

typedef struct ellspmv_cuda_info {
    int64_t num_rows;        /* rows held by this handle (the shard)        */
    int64_t num_columns;
    int64_t rowsize;         /* K                                           */
    int64_t row_begin;       /* first global row of the shard               */
    int64_t global_rows;     /* rows of the whole matrix                    */
    int     idx_width_bits;  /* as given by the caller (host view)          */
    int     dev_idx_bits;    /* as stored on the device                     */
    int     slice_rows;      /* sliced-ELL slice height                     */
    int     rows_per_thread;
    int     kernel;          /* ELLSPMV_CUDA_KERNEL_THREAD / _WARP in use   */
    int     fma;
    int     device;          /* CUDA device ordinal                         */
    int64_t device_bytes;    /* bytes of HBM held by the handle             */
    int64_t min_col, max_col;/* column range referenced by this shard       */
    int64_t launches;        /* SpMV kernel launches issued so far          */
    int     num_gpus;        /* GPUs behind this handle (row shards)        */
    int64_t pattern_rows;    /* rows whose column indices come from an      */
                             /*   offset pattern instead of the index stream */
    int     staged;          /* 1: the staged gather (column blocks) is in use; */
                             /*   2: chosen by KERNEL_AUTO from a timed trial  */
    int     launches_per_spmv; /* kernel launches one spmv_device call issues  */
    double  tune_ms[2];      /* AUTO's timed trial at upload: direct gather,  */
                             /*   staged gather (0 = not tried)               */
    int64_t exception_entries; /* entries patched after the pattern lookup    */
    int64_t long_rows;       /* 0, or the row length from which the long-row  */
                             /*   kernel is used (KERNEL_AUTO)                */
    int64_t sell_slots;      /* ELLSPMV_CUDA_SKIP_PADDING: slots stored in the */
                             /*   SELL-128-sigma copy (vs num_rows * rowsize)  */
    int64_t value_pattern_rows; /* rows whose coefficients come from the       */
                             /*   dictionary too (value patterns): = pattern_rows */
                             /*   or 0                                         */
    int64_t pattern_id_bytes;/* bytes of pattern ids one SpMV reads: one per  */
                             /*   group of 32*R rows, or one per thread (lane */
                             /*   patterns); 0 without offset patterns        */
} ellspmv_cuda_info;

/* ---- ELL ------------------------------------------------------------- */

/*
 * Copy a row-major ELL matrix (the arrays ell_from_coo produces,
 * ellspmv.c:1081-1127: colidx[i*K+l], a[i*K+l], 0-based, padded) to the
 * device and re-lay it as sliced ELL.  Replaces nothing in the reference;
 * call it once after ell_from_coo (~ellspmv.c:1745).
 *   idx_width_bits  32 or 64 (sizeof(idx_t)*8)
 *   num_gpus        1, or N > 1: rows are split in N contiguous blocks
 *                   (rows/N + (p < rows%N), the reference's static split,
 *                   csrspmv.c:2238) over CUDA devices 0..N-1 of this process;
 *                   ellspmv_cuda_spmv then drives all of them from the calling
 *                   thread (peer access over NVLink; ITERATE uses the fused
 *                   SpMV+push kernel and the device-side barrier, no NCCL).
 *                   spmv_device/spmv_push need a single-GPU handle.
 */
int ellspmv_cuda_upload(
    ellspmv_cuda_matrix **out, int idx_width_bits,
    int64_t num_rows, int64_t num_columns, int64_t rowsize,
    const void *colidx, const double *a, int num_gpus, unsigned flags);

/*
 * Same, for rows [row_begin, row_end) of a larger matrix on CUDA device
 * `device`: colidx/a point at the shard's first row, column indices stay
 * global.  Used for row-sharded repeated SpMV.
 */
int ellspmv_cuda_upload_shard(
    ellspmv_cuda_matrix **out, int idx_width_bits,
    int64_t global_rows, int64_t num_columns, int64_t rowsize,
    int64_t row_begin, int64_t row_end,
    const void *colidx, const double *a, int device, unsigned flags);

/*
 * COO -> ELL on the device: takes the arrays as read from the Matrix Market
 * file (1-based rowidx/colidx, file order) and builds the same device matrix
 * as ell_from_coo (ellspmv.c:931-958, 1081-1127) followed by
 * ellspmv_cuda_upload would: K = widest row, entries in file order inside a
 * row, (min(i, ncols-1), 0.0) padding.  Replaces the reference's serial host
 * scatter; the result is bit-identical (ellspmv_cuda_download shows it).
 */
int ellspmv_cuda_upload_coo(
    ellspmv_cuda_matrix **out, int idx_width_bits,
    int64_t num_rows, int64_t num_columns, int64_t num_nonzeros,
    const void *rowidx, const void *colidx, const double *a, int device, unsigned flags);

/*
 * Build rows [row_begin, row_end) of a synthetic matrix directly on the
 * device, in device layout (for shapes too large to build on the host).
 * Bit-identical to uploading the arrays the reference's ell_from_coo
 * produces from the same entries (tested).  device < 0: current device.
 */
int ellspmv_cuda_generate(
    ellspmv_cuda_matrix **out, int kind, const int64_t dims[3],
    const double vals[2], uint64_t seed, int idx_width_bits,
    int64_t row_begin, int64_t row_end, int device, unsigned flags);

/* the same over num_gpus row shards (see ellspmv_cuda_upload) */
int ellspmv_cuda_generate_sharded(
    ellspmv_cuda_matrix **out, int kind, const int64_t dims[3],
    const double vals[2], uint64_t seed, int idx_width_bits, int num_gpus, unsigned flags);

/*
 * y <- y + A*x (mode ACCUMULATE; replaces the ellgemv call,
 * ellspmv.c:1841-1842), with HOST vectors: x has num_columns entries, y has
 * num_rows entries.  Copies x and y in, runs `repeat` kernel launches,
 * copies y out.  seconds (may be NULL) receives `repeat` per-launch device
 * times measured with CUDA events, the counterpart of the reference's
 * per-iteration CLOCK_MONOTONIC pair (ellspmv.c:1825-1847).
 */
int ellspmv_cuda_spmv(
    ellspmv_cuda_matrix *A, double *y, const double *x,
    int repeat, int mode, double *seconds);

/*
 * One launch on DEVICE vectors, asynchronous on `stream` (a cudaStream_t;
 * NULL = the legacy default stream).  x_dev: num_columns doubles; y_dev:
 * the shard's num_rows doubles.  mode: ACCUMULATE or OVERWRITE.
 */
int ellspmv_cuda_spmv_device(
    ellspmv_cuda_matrix *A, double *y_dev, const double *x_dev,
    int mode, void *stream);

/*
 * As above, and additionally store each computed y[i] (global row
 * row_begin+i) into up to 8 peer vectors: for peer p, rows in
 * [peer_row_lo[p], peer_row_hi[p]) are written to peer_x[p][global row].
 * peer_x are device pointers valid on this device (peer-mapped or IPC
 * opened): the fused SpMV + exchange used for row-sharded repeated SpMV.
 */
int ellspmv_cuda_spmv_push(
    ellspmv_cuda_matrix *A, double *y_dev, const double *x_dev, int mode,
    int num_peers, double *const *peer_x,
    const int64_t *peer_row_lo, const int64_t *peer_row_hi, void *stream);

/*
 * One step of the row-sharded y -> x loop (BASELINE config 5): spmv_push plus the hand-shake
 * between this rank and the ranks it exchanges rows with -- no all-ranks barrier, no NCCL call
 * between two steps.  After its SpMV + push kernel a one-warp kernel stores `epoch` into slot
 * [rank] of each listed rank's flag array (system-scope release over NVLink) and waits until
 * those ranks have stored `epoch` here: their pushes of this step have landed, and they are
 * done reading the vector the next step overwrites.  With ELLSPMV_CUDA_FUSED_SYNC (upload
 * flag) the hand-shake runs inside the SpMV kernel instead: warps of the slices that read halo
 * columns or push wait -- at their start -- for epoch-1, interior warps never wait, the last
 * boundary warp to finish signals `epoch`.
 *   sync_ranks/sync_flags  the ranks this one pushes to or is pushed by, and their flag arrays
 *                          (peer-mapped; see ellspmv_cuda_peer_barrier for the array's shape)
 *   local_flags            this rank's own flag array; slot [16] turns non-zero if a peer never
 *                          arrived (~20 s), slot [rank] is unused
 *   epoch                  1, 2, 3, ... one per step, the same on every rank
 */
int ellspmv_cuda_spmv_exchange(
    ellspmv_cuda_matrix *A, double *y_dev, const double *x_dev, int mode,
    int num_peers, double *const *peer_x,
    const int64_t *peer_row_lo, const int64_t *peer_row_hi,
    int rank, int num_sync, const int *sync_ranks, int64_t *const *sync_flags,
    int64_t *local_flags, int64_t epoch, void *stream);

/*
 * Attach a separately stored diagonal: afterwards every spmv computes
 * y <- y + (ad .* x + A*x), the reference's ellgemvsd (ellspmv.c:1155-1180;
 * order 0: yi summed from 0, then ad*x + yi) or ellgemv16sd
 * (ellspmv.c:1182-1221; order 1: the sum starts at ad*x).  ad has the shard's
 * num_rows entries (host or device memory) and is copied; NULL detaches it.
 * Needs rows <= columns, like the reference's kernels (they read x[i]).
 */
int ellspmv_cuda_set_diagonal(ellspmv_cuda_matrix *A, const double *ad, int order);

/* Copy the device matrix back as row-major host arrays of the caller's
 * index width (inverse of upload; used to test bit-exactness). */
int ellspmv_cuda_download(
    const ellspmv_cuda_matrix *A, void *colidx, double *a);

int  ellspmv_cuda_get_info(const ellspmv_cuda_matrix *A, ellspmv_cuda_info *info);
void ellspmv_cuda_free(ellspmv_cuda_matrix *A);

/* ---- CSR (comparison path) ------------------------------------------- */

/* rowptr: num_rows+1 int64 (always 64-bit in the reference, csrspmv.c:1573);
 * colidx: csrsize idx_t, 0-based; a: csrsize doubles.
 * num_gpus > 1: contiguous row blocks balanced by entries (the reference's
 * --partition-nonzeros rule, csrspmv.c:1700-1708, rounded to row boundaries)
 * over devices 0..N-1; csrspmv_cuda_spmv then drives all of them. */
int csrspmv_cuda_upload(
    csrspmv_cuda_matrix **out, int idx_width_bits,
    int64_t num_rows, int64_t num_columns,
    const int64_t *rowptr, const void *colidx, const double *a,
    int num_gpus, unsigned flags);

/* COO -> CSR on the device (general matrices): the stable sort by row of
 * csr_from_coo's default branch (csrspmv.c:1436-1465), bit-identical arrays */
int csrspmv_cuda_upload_coo(
    csrspmv_cuda_matrix **out, int idx_width_bits,
    int64_t num_rows, int64_t num_columns, int64_t num_nonzeros,
    const void *rowidx, const void *colidx, const double *a, unsigned flags);

/* copy the CSR arrays back to the host (rowptr: num_rows+1 int64) */
int csrspmv_cuda_download(const csrspmv_cuda_matrix *A, int64_t *rowptr, void *colidx, double *a);

/* the generators of ellspmv_cuda_generate in CSR form, as csr_from_coo (csrspmv.c:1390-1475)
 * stores the same canonical COO stream: ELLSPMV_CUDA_GEN_RANDOM (every row has exactly K
 * entries: rowptr[i] = i*K), ELLSPMV_CUDA_GEN_LAPLACE2D / _STENCIL27 (no padding: rows at the
 * grid boundary are shorter; vals = {centre, off}) */
int csrspmv_cuda_generate(
    csrspmv_cuda_matrix **out, int kind, const int64_t dims[3],
    const double vals[2], uint64_t seed, int idx_width_bits,
    int device, unsigned flags);

/* attach the separately stored diagonal of csrgemvsd (csrspmv.c:1598-1629):
 * y <- y + (ad .* x + A*x); ad has num_rows entries, NULL detaches it */
int csrspmv_cuda_set_diagonal(csrspmv_cuda_matrix *A, const double *ad);

/* replaces the csrgemv call, csrspmv.c:2857-2858 */
int csrspmv_cuda_spmv(
    csrspmv_cuda_matrix *A, double *y, const double *x,
    int repeat, int mode, double *seconds);

int csrspmv_cuda_spmv_device(
    csrspmv_cuda_matrix *A, double *y_dev, const double *x_dev,
    int mode, void *stream);

typedef struct csrspmv_cuda_info {
    int64_t num_rows, num_columns, csrsize;
    int64_t min_row_len, max_row_len;
    int64_t min_col, max_col;
    int64_t device_bytes;
    int     kernel;            /* kernel in use: 1 stream, 2 vector, 3 scalar, 5 SELL-128-sigma  */
    int64_t sell_slots;        /* kernel 5: slots stored in the slices (padding included) ...    */
    int64_t sell_real;         /*   ... of which count; */
    int64_t sell_long_rows;    /*   rows longer than sell_long_len entries, run one CTA per row  */
    int64_t sell_long_len;     /*   that length (256; environment SELL_LONG_ROW overrides)       */
    int     ell_view;          /* 0: native CSR kernels; 1: sliced-ELL view with per-row        */
                               /*   lengths; 2: view of rows of one length (no length array)    */
    int     ell_staged;        /* the view runs the staged gather (see ellspmv_cuda_info)       */
    int     launches_per_spmv;
    int64_t ell_pattern_rows;  /* rows of the view on an offset pattern                         */
    int     num_gpus;
    int     fma;
    int64_t ell_pattern_id_bytes; /* pattern-id bytes one SpMV of the view reads             */
    int     ell_dev_idx_bits;  /* index width of the view on the device (0 without a view)  */
    int     ell_rows_per_thread;
} csrspmv_cuda_info;
int csrspmv_cuda_get_info(const csrspmv_cuda_matrix *A, csrspmv_cuda_info *info);

int64_t csrspmv_cuda_device_bytes(const csrspmv_cuda_matrix *A);
void csrspmv_cuda_free(csrspmv_cuda_matrix *A);

/* ---- utilities -------------------------------------------------------- */

/* pinned host memory for x / y so the copies inside spmv run at full PCIe
 * rate (the host program allocates its vectors with these) */
int  ellspmv_cuda_malloc_host(void **ptr, int64_t bytes);
void ellspmv_cuda_free_host(void *ptr);

/* plain device allocations (cudaMalloc) for the exchange vectors of the
 * row-sharded path; unlike a sub-allocated framework tensor these can be
 * exported over CUDA IPC as they are */
int  ellspmv_cuda_malloc_device(void **ptr, int64_t bytes);
void ellspmv_cuda_free_device(void *ptr);

/* CUDA IPC plumbing for one-process-per-GPU sharding: export a device
 * allocation as a 64-byte handle / open a peer's handle */
int ellspmv_cuda_ipc_export(const void *dev_ptr, unsigned char handle[64]);
int ellspmv_cuda_ipc_open(const unsigned char handle[64], void **dev_ptr);
int ellspmv_cuda_ipc_close(void *dev_ptr);

/*
 * Device-side step barrier between the ranks of a row-sharded repeated SpMV,
 * asynchronous on `stream`: writes `epoch` into slot [rank] of every rank's
 * flag array and waits until the local array shows `epoch` from all ranks.
 * Flag arrays are 32 zero-initialised int64 in ellspmv_cuda_malloc_device
 * memory, peers' arrays opened over CUDA IPC; epochs must increase.  Slot
 * [16] of the local array turns non-zero if a peer never arrived (~20 s).
 */
int ellspmv_cuda_peer_barrier(int rank, int nranks, int64_t epoch, int64_t *local_flags,
                              int64_t *const *peer_flags, void *stream);

int ellspmv_cuda_device_count(int *count);
const char *ellspmv_cuda_strerror(int err);
const char *ellspmv_cuda_last_error(void);
int ellspmv_cuda_version(void);

#ifdef __cplusplus
}
#endif
#endif /* ELLSPMV_CUDA_H */
