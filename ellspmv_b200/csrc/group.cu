// group.cu -- one handle over several GPUs of this process (num_gpus > 1).
//
// Rows are split into contiguous blocks by the reference's static OpenMP rule,
// rows/N + (p < rows % N) (csrspmv.c:2238); shard p lives on CUDA device p and
// keeps global column indices.  Every device holds two full-length vectors.
// One host thread drives all devices (the reference calls its kernel from every
// OpenMP thread of one process; here the "threads" are GPUs):
//
//   ACCUMULATE / OVERWRITE   x is constant (the reference's semantics, Q6), so
//       the shards are independent: x goes to device 0 over PCIe and on to the
//       others over NVLink, every device runs its launches, y slices come back.
//   ITERATE (x <- A*x)       per step every device runs the fused SpMV + push
//       kernel (each fresh y[i] is stored into its own next-x slice and into
//       the next-x vectors of the peers that reference row i, peer-mapped HBM
//       over NVLink) followed by the device-side flag barrier (barrier.cu).
//       No NCCL, no host synchronisation inside the loop.
//
// Row sharding keeps each row's summation order, so results are bit-identical
// to the single-GPU handle (tests/test_gpu_group.py).
#include <string.h>

#include "handles.cuh"

namespace ellspmv {

namespace {

void split_rows(int64_t rows, int n, std::vector<int64_t> &lo, std::vector<int64_t> &hi)
{
    lo.resize(n); hi.resize(n);
    const int64_t base = rows / n, rem = rows % n;
    int64_t at = 0;
    for (int p = 0; p < n; p++) {
        lo[p] = at;
        at += base + (p < rem ? 1 : 0);
        hi[p] = at;
    }
}

// the global rows of `owner` that `reader`'s stored entries reference, if any
bool overlap(const ellspmv_cuda_matrix *reader, const ellspmv_cuda_matrix *owner, int64_t *lo, int64_t *hi)
{
    if (reader->max_col < reader->min_col) return false;
    *lo = reader->min_col > owner->row_begin ? reader->min_col : owner->row_begin;
    const int64_t end = owner->row_begin + owner->lay.num_rows;
    *hi = reader->max_col + 1 < end ? reader->max_col + 1 : end;
    return *lo < *hi;
}

int enable_peers(int n)
{
    int count = 0;
    ELL_CK(cudaGetDeviceCount(&count));
    if (n > count) ELL_FAIL(ENODEV, "num_gpus=%d but only %d CUDA device(s) are visible", n, count);
    if (n > kMaxPeers) ELL_FAIL(EINVAL, "num_gpus=%d: at most %d", n, kMaxPeers);
    for (int p = 0; p < n; p++) {
        ELL_CK(cudaSetDevice(p));
        for (int q = 0; q < n; q++) {
            if (q == p) continue;
            int ok = 0;
            ELL_CK(cudaDeviceCanAccessPeer(&ok, p, q));
            if (!ok) ELL_FAIL(ENOTSUP, "device %d cannot access device %d's memory", p, q);
            cudaError_t e = cudaDeviceEnablePeerAccess(q, 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) ELL_CK(e);
            cudaGetLastError();
        }
    }
    return 0;
}

// allocate the per-device vectors and barrier flags once the shards exist
int finish_group(ellspmv_cuda_matrix *G)
{
    const int n = (int)G->shards.size();
    G->vec_elems = G->num_columns > G->global_rows ? G->num_columns : G->global_rows;
    if (G->vec_elems < 1) G->vec_elems = 1;
    for (int b = 0; b < 2; b++) G->xb[b].assign(n, nullptr);
    G->bflags.assign(n, nullptr);
    G->gevents.resize(n);
    G->min_col = 0x7fffffffffffffffLL;
    G->max_col = -1;
    G->device_bytes = 0;
    for (int p = 0; p < n; p++) {
        ELL_CK(cudaSetDevice(p));
        for (int b = 0; b < 2; b++) ELL_CK(cudaMalloc(&G->xb[b][p], (size_t)G->vec_elems * 8));
        ELL_CK(cudaMalloc(&G->bflags[p], 32 * sizeof(long long)));
        ELL_CK(cudaMemset(G->bflags[p], 0, 32 * sizeof(long long)));
        ellspmv_cuda_matrix *S = G->shards[p];
        if (S->max_col >= S->min_col) {
            if (S->min_col < G->min_col) G->min_col = S->min_col;
            if (S->max_col > G->max_col) G->max_col = S->max_col;
        }
        G->device_bytes += S->device_bytes + 2 * G->vec_elems * 8;
    }
    if (G->max_col < 0) G->min_col = 0;
    G->cfg = G->shards[0]->cfg;
    G->dev_idx_bits = G->shards[0]->dev_idx_bits;
    G->lay.slice_rows = G->shards[0]->lay.slice_rows;
    G->lay.rowsize = G->shards[0]->lay.rowsize;
    return 0;
}

ellspmv_cuda_matrix *new_group(int idx_bits, int64_t rows, int64_t cols, unsigned flags)
{
    ellspmv_cuda_matrix *G = new (std::nothrow) ellspmv_cuda_matrix();
    if (!G) return nullptr;
    G->device = 0;
    G->host_idx_bits = idx_bits;
    G->num_columns = cols;
    G->global_rows = rows;
    G->row_begin = 0;
    G->lay.num_rows = rows;
    G->flags = flags;
    return G;
}

}  // namespace

void group_free(ellspmv_cuda_matrix *G)
{
    if (!G) return;
    int prev = -1;
    cudaGetDevice(&prev);
    const int n = (int)G->shards.size();
    for (int p = 0; p < n; p++) {
        cudaSetDevice(p);
        cudaDeviceSynchronize();
        for (int b = 0; b < 2; b++) if ((int)G->xb[b].size() > p && G->xb[b][p]) cudaFree(G->xb[b][p]);
        if ((int)G->bflags.size() > p && G->bflags[p]) cudaFree(G->bflags[p]);
        if ((int)G->gevents.size() > p) for (cudaEvent_t e : G->gevents[p]) cudaEventDestroy(e);
    }
    std::vector<ellspmv_cuda_matrix *> shards;
    shards.swap(G->shards);                    // so that ellspmv_cuda_free sees plain handles
    for (ellspmv_cuda_matrix *S : shards) ellspmv_cuda_free(S);
    if (prev >= 0) cudaSetDevice(prev);
    delete G;
}

int group_upload(ellspmv_cuda_matrix **out, int idx_width_bits, int64_t num_rows, int64_t num_columns,
                 int64_t rowsize, const void *colidx, const double *a, int num_gpus, unsigned flags)
{
    if (!out) ELL_FAIL(EINVAL, "out is NULL");
    *out = nullptr;
    if (idx_width_bits != 32 && idx_width_bits != 64) ELL_FAIL(EINVAL, "idx_width_bits must be 32 or 64");
    if (num_rows < 0 || num_columns < 0 || rowsize < 0) ELL_FAIL(EINVAL, "negative dimension");
    int prev = -1;
    cudaGetDevice(&prev);
    int err = enable_peers(num_gpus);
    if (err) { if (prev >= 0) cudaSetDevice(prev); return err; }
    ellspmv_cuda_matrix *G = new_group(idx_width_bits, num_rows, num_columns, flags);
    if (!G) ELL_FAIL(ENOMEM, "out of host memory");
    std::vector<int64_t> lo, hi;
    split_rows(num_rows, num_gpus, lo, hi);
    const int64_t ib = idx_width_bits / 8;
    for (int p = 0; p < num_gpus && !err; p++) {
        ellspmv_cuda_matrix *S = nullptr;
        err = ellspmv_cuda_upload_shard(&S, idx_width_bits, num_rows, num_columns, rowsize, lo[p], hi[p],
                                        colidx ? (const char *)colidx + lo[p] * rowsize * ib : nullptr,
                                        a ? a + lo[p] * rowsize : nullptr, p, flags);
        if (!err) G->shards.push_back(S);
    }
    if (!err) err = finish_group(G);
    if (prev >= 0) cudaSetDevice(prev);
    if (err) { group_free(G); return err; }
    *out = G;
    return 0;
}

int group_generate(ellspmv_cuda_matrix **out, int kind, const int64_t dims[3], const double vals[2],
                   uint64_t seed, int idx_width_bits, int num_gpus, unsigned flags)
{
    if (!out || !dims) ELL_FAIL(EINVAL, "NULL argument");
    *out = nullptr;
    int64_t rows, cols;
    switch (kind) {
    case ELLSPMV_CUDA_GEN_LAPLACE2D: rows = cols = dims[0] * dims[1]; break;
    case ELLSPMV_CUDA_GEN_STENCIL27: rows = cols = dims[0] * dims[1] * dims[2]; break;
    case ELLSPMV_CUDA_GEN_RANDOM: rows = dims[0]; cols = dims[1]; break;
    default: ELL_FAIL(EINVAL, "unknown generator kind %d", kind);
    }
    int prev = -1;
    cudaGetDevice(&prev);
    int err = enable_peers(num_gpus);
    if (err) { if (prev >= 0) cudaSetDevice(prev); return err; }
    ellspmv_cuda_matrix *G = new_group(idx_width_bits, rows, cols, flags);
    if (!G) ELL_FAIL(ENOMEM, "out of host memory");
    std::vector<int64_t> lo, hi;
    split_rows(rows, num_gpus, lo, hi);
    for (int p = 0; p < num_gpus && !err; p++) {
        ellspmv_cuda_matrix *S = nullptr;
        err = ellspmv_cuda_generate(&S, kind, dims, vals, seed, idx_width_bits, lo[p], hi[p], p, flags);
        if (!err) G->shards.push_back(S);
    }
    if (!err) err = finish_group(G);
    if (prev >= 0) cudaSetDevice(prev);
    if (err) { group_free(G); return err; }
    *out = G;
    return 0;
}

int group_info(const ellspmv_cuda_matrix *G, ellspmv_cuda_info *info)
{
    memset(info, 0, sizeof(*info));
    info->num_rows = G->global_rows;
    info->num_columns = G->num_columns;
    info->rowsize = G->lay.rowsize;
    info->row_begin = 0;
    info->global_rows = G->global_rows;
    info->idx_width_bits = G->host_idx_bits;
    info->dev_idx_bits = G->dev_idx_bits;
    info->slice_rows = G->lay.slice_rows;
    info->rows_per_thread = G->cfg.rows_per_thread;
    info->kernel = G->cfg.kernel;
    info->fma = G->cfg.fma ? 1 : 0;
    info->device = 0;
    info->device_bytes = G->device_bytes;
    info->min_col = G->min_col;
    info->max_col = G->max_col;
    int64_t launches = 0;
    for (const ellspmv_cuda_matrix *S : G->shards) {
        launches += S->launches;
        info->pattern_rows += S->pat.covered * S->pat.group_rows - S->pat.explicit_lanes * S->cfg.rows_per_thread;
        info->exception_entries += S->pat.explicit_lanes * S->cfg.rows_per_thread * S->lay.rowsize;
        if (S->pat.vpat) info->value_pattern_rows += S->pat.covered * S->pat.group_rows;
        info->pattern_id_bytes += S->pat.patlane ? S->pat.groups * 32 : (S->pat.max_explicit ? S->pat.groups * 8 : (S->pat.patid ? S->pat.groups : 0));
    }
    info->launches = launches;
    info->num_gpus = (int)G->shards.size();
    return 0;
}

int group_download(const ellspmv_cuda_matrix *G, void *colidx, double *a)
{
    const int64_t K = G->lay.rowsize, ib = G->host_idx_bits / 8;
    for (const ellspmv_cuda_matrix *S : G->shards) {
        int err = ellspmv_cuda_download(S, colidx ? (char *)colidx + S->row_begin * K * ib : nullptr,
                                        a ? a + S->row_begin * K : nullptr);
        if (err) return err;
    }
    return 0;
}

int group_spmv(ellspmv_cuda_matrix *G, double *y, const double *x, int repeat, int mode, double *seconds)
{
    const int n = (int)G->shards.size();
    const int64_t rows = G->global_rows, ncols = G->num_columns;
    if ((rows > 0 && !y) || (ncols > 0 && !x)) ELL_FAIL(EINVAL, "NULL host vector");
    if (mode == ELLSPMV_CUDA_ITERATE && rows != ncols) ELL_FAIL(EINVAL, "ITERATE needs a square matrix");
    if (repeat == 0) return 0;
    int prev = -1;
    cudaGetDevice(&prev);
    struct Restore { int d; ~Restore() { if (d >= 0) cudaSetDevice(d); } } restore{prev};

    // events: one per device per iteration boundary
    for (int p = 0; p < n; p++) {
        ELL_CK(cudaSetDevice(p));
        int err = ensure_event_count(G->gevents[p], (size_t)repeat + 1);
        if (err) return err;
    }
    // x: every device takes the range its shard references (its own slice plus the halo for a
    // stencil, everything for a scattered matrix) straight from the host vector, all copies in
    // flight at once -- not N copies of the whole vector (api.cu: shard_x_range)
    for (int p = 0; p < n; p++) {
        ellspmv_cuda_matrix *S = G->shards[p];
        int64_t lo, hi;
        shard_x_range(S, &lo, &hi);
        ELL_CK(cudaSetDevice(p));
        if (hi > lo) ELL_CK(cudaMemcpyAsync(G->xb[0][p] + lo, x + lo, (size_t)(hi - lo) * 8, cudaMemcpyDefault, S->stream));
    }
    const bool iterate = mode == ELLSPMV_CUDA_ITERATE;
    const int beta = mode == ELLSPMV_CUDA_ACCUMULATE ? 1 : 0;
    if (!iterate && beta) {
        for (int p = 0; p < n; p++) {      // y slices: host -> their device (kept at their global offset in xb[1])
            ellspmv_cuda_matrix *S = G->shards[p];
            ELL_CK(cudaSetDevice(p));
            if (S->lay.num_rows > 0)
                ELL_CK(cudaMemcpyAsync(G->xb[1][p] + S->row_begin, y + S->row_begin, (size_t)S->lay.num_rows * 8,
                                       cudaMemcpyDefault, S->stream));
        }
    }
    if (iterate) {
        // all devices must hold x before anybody pushes into a next-x vector
        for (int p = 0; p < n; p++) { ELL_CK(cudaSetDevice(p)); ELL_CK(cudaStreamSynchronize(G->shards[p]->stream)); }
    }
    for (int p = 0; p < n; p++) {
        ELL_CK(cudaSetDevice(p));
        ELL_CK(cudaEventRecord(G->gevents[p][0], G->shards[p]->stream));
    }
    int cur = 0;
    for (int r = 0; r < repeat; r++) {
        for (int p = 0; p < n; p++) {
            ellspmv_cuda_matrix *S = G->shards[p];
            ELL_CK(cudaSetDevice(p));
            int err;
            if (iterate) {
                const int nxt = 1 - cur;
                // one kernel per device and step: SpMV, push of the rows the peers reference, and the
                // step hand-shake with exactly those peers (fused where the kernel carries it)
                PushTargets pt = {};
                StepSync sy = {};
                sy.local_flags = G->bflags[p];
                sy.rank = p;
                sy.epoch = G->epoch + 1;
                sy.error = reinterpret_cast<int *>(G->bflags[p] + kMaxRanks);
                for (int q = 0; q < n; q++) {
                    if (q == p) continue;
                    int64_t lo, hi;
                    const bool to_q = overlap(G->shards[q], S, &lo, &hi);      // rows of p that q reads
                    if (to_q) {
                        pt.x[pt.num_peers] = G->xb[nxt][q];
                        pt.row_lo[pt.num_peers] = lo;
                        pt.row_hi[pt.num_peers] = hi;
                        pt.num_peers++;
                    }
                    int64_t lo2, hi2;
                    if (to_q || overlap(S, G->shards[q], &lo2, &hi2)) {        // ... or rows of q that p reads
                        sy.peer_flags[sy.num_peers] = G->bflags[q];
                        sy.peer_rank[sy.num_peers] = q;
                        sy.num_peers++;
                    }
                }
                err = launch_shard_exchange(S, G->xb[nxt][p] + S->row_begin, G->xb[cur][p], 0, &pt, sy, S->stream);
            } else {
                err = launch_shard(S, G->xb[1][p] + S->row_begin, G->xb[0][p], beta, nullptr, S->stream);
            }
            if (err) return err;
            ELL_CK(cudaEventRecord(G->gevents[p][(size_t)r + 1], S->stream));
        }
        if (iterate) { cur = 1 - cur; G->epoch++; }
    }
    // results: every device returns its own rows
    for (int p = 0; p < n; p++) {
        ellspmv_cuda_matrix *S = G->shards[p];
        ELL_CK(cudaSetDevice(p));
        const double *src = (iterate ? G->xb[cur][p] : G->xb[1][p]) + S->row_begin;
        if (S->lay.num_rows > 0)
            ELL_CK(cudaMemcpyAsync(y + S->row_begin, src, (size_t)S->lay.num_rows * 8, cudaMemcpyDefault, S->stream));
    }
    for (int p = 0; p < n; p++) {
        ELL_CK(cudaSetDevice(p));
        ELL_CK(cudaStreamSynchronize(G->shards[p]->stream));
        if (G->shards[p]->side) ELL_CK(cudaStreamSynchronize(G->shards[p]->side));   // the last step's hand-shake
    }
    if (seconds) {
        for (int r = 0; r < repeat; r++) {
            double worst = 0.0;
            for (int p = 0; p < n; p++) {
                float ms = 0.f;
                ELL_CK(cudaSetDevice(p));
                ELL_CK(cudaEventElapsedTime(&ms, G->gevents[p][(size_t)r], G->gevents[p][(size_t)r + 1]));
                if (ms * 1e-3 > worst) worst = ms * 1e-3;
            }
            seconds[r] = worst;       // the slowest GPU sets the iteration time
        }
    }
    for (int p = 0; p < n; p++) {      // a peer that never arrived leaves a mark in slot 16
        long long mark = 0;
        ELL_CK(cudaSetDevice(p));
        ELL_CK(cudaMemcpy(&mark, G->bflags[p] + kMaxRanks, sizeof(mark), cudaMemcpyDeviceToHost));
        if (mark) {
            // report once: clear the mark so that the handle stays usable after the error
            cudaMemset(G->bflags[p] + kMaxRanks, 0, sizeof(mark));
            ELL_FAIL(EIO, "device %d gave up waiting for device %lld in the step synchronisation", p,
                     (long long)(mark & 0xffffffffLL) - 1);
        }
    }
    return 0;
}

// ---- CSR over several GPUs (comparison path) ---------------------------------------
// Contiguous row blocks balanced by entries, the reference's --partition-nonzeros
// rule (csrspmv.c:1700-1708: start = p * ceil(nnz / T)), rounded to row boundaries.
// x is constant in the reference's loop, so the shards are independent.

void csr_group_free(csrspmv_cuda_matrix *G)
{
    if (!G) return;
    std::vector<csrspmv_cuda_matrix *> shards;
    shards.swap(G->shards);
    for (csrspmv_cuda_matrix *S : shards) csrspmv_cuda_free(S);
    delete G;
}

int csr_group_upload(csrspmv_cuda_matrix **out, int idx_width_bits, int64_t num_rows, int64_t num_columns,
                     const int64_t *rowptr, const void *colidx, const double *a, int num_gpus, unsigned flags)
{
    if (!out) ELL_FAIL(EINVAL, "out is NULL");
    *out = nullptr;
    if (!rowptr) ELL_FAIL(EINVAL, "rowptr is NULL");
    if (num_rows < 0 || num_columns < 0) ELL_FAIL(EINVAL, "negative dimension");
    int count = 0;
    ELL_CK(cudaGetDeviceCount(&count));
    if (num_gpus > count) ELL_FAIL(ENODEV, "num_gpus=%d but only %d CUDA device(s) are visible", num_gpus, count);
    // the partition needs rowptr on the host
    std::vector<int64_t> rp((size_t)num_rows + 1);
    ELL_CK(cudaMemcpy(rp.data(), rowptr, (size_t)(num_rows + 1) * 8, cudaMemcpyDefault));
    const int64_t nnz = rp[num_rows] - rp[0];
    const int64_t per = (nnz + num_gpus - 1) / num_gpus;
    csrspmv_cuda_matrix *G = new (std::nothrow) csrspmv_cuda_matrix();
    if (!G) ELL_FAIL(ENOMEM, "out of host memory");
    G->idx_bits = idx_width_bits;
    G->num_rows = num_rows;
    G->num_columns = num_columns;
    G->csrsize = nnz;
    G->flags = flags;
    G->row_lo.assign((size_t)num_gpus + 1, num_rows);
    int64_t r = 0;
    for (int p = 0; p < num_gpus; p++) {
        const int64_t startnz = rp[0] + (int64_t)p * per;
        while (r < num_rows && rp[r] < startnz) r++;          // first row that starts at or after the cut
        G->row_lo[p] = p == 0 ? 0 : r;
    }
    int prev = -1;
    cudaGetDevice(&prev);
    int err = 0;
    const int64_t ib = idx_width_bits / 8;
    (void)ib;
    for (int p = 0; p < num_gpus && !err; p++) {
        const int64_t lo = G->row_lo[p], hi = G->row_lo[p + 1];
        csrspmv_cuda_matrix *S = nullptr;
        // the shard keeps the parent's entry offsets; csr_upload_on rebases rowptr and slices colidx / a
        err = csr_upload_on(&S, idx_width_bits, hi - lo, num_columns, rowptr + lo, colidx, a, p, flags, lo);
        if (!err) { G->shards.push_back(S); G->device_bytes += S->device_bytes; }
    }
    if (prev >= 0) cudaSetDevice(prev);
    if (err) { csr_group_free(G); return err; }
    *out = G;
    return 0;
}

int csr_group_spmv(csrspmv_cuda_matrix *G, double *y, const double *x, int repeat, int mode, double *seconds)
{
    const int n = (int)G->shards.size();
    int prev = -1;
    cudaGetDevice(&prev);
    struct Restore { int d; ~Restore() { if (d >= 0) cudaSetDevice(d); } } restore{prev};
    std::vector<std::vector<double>> secs((size_t)n, std::vector<double>((size_t)repeat, 0.0));
    // every shard owns its x copy and y slice: issue everything asynchronously, then wait
    for (int p = 0; p < n; p++) {
        csrspmv_cuda_matrix *S = G->shards[p];
        ELL_CK(cudaSetDevice(p));
        int err = ensure_event_count(S->events, (size_t)repeat + 1);
        if (err) return err;
        cudaStream_t s = S->stream;
        int64_t xlo, xhi;
        csr_x_range(S, &xlo, &xhi);
        if (xhi > xlo) ELL_CK(cudaMemcpyAsync(S->d_x + xlo, x + xlo, (size_t)(xhi - xlo) * 8, cudaMemcpyDefault, s));
        if (mode == ELLSPMV_CUDA_ACCUMULATE && S->num_rows > 0)
            ELL_CK(cudaMemcpyAsync(S->d_y, y + G->row_lo[p], (size_t)S->num_rows * 8, cudaMemcpyDefault, s));
        ELL_CK(cudaEventRecord(S->events[0], s));
        for (int r = 0; r < repeat; r++) {
            if ((err = csr_launch(S, S->d_y, S->d_x, mode == ELLSPMV_CUDA_ACCUMULATE ? 1 : 0, s))) return err;
            ELL_CK(cudaEventRecord(S->events[(size_t)r + 1], s));
        }
        if (S->num_rows > 0)
            ELL_CK(cudaMemcpyAsync(y + G->row_lo[p], S->d_y, (size_t)S->num_rows * 8, cudaMemcpyDefault, s));
    }
    for (int p = 0; p < n; p++) {
        csrspmv_cuda_matrix *S = G->shards[p];
        ELL_CK(cudaSetDevice(p));
        ELL_CK(cudaStreamSynchronize(S->stream));
        for (int r = 0; r < repeat; r++) {
            float ms = 0.f;
            ELL_CK(cudaEventElapsedTime(&ms, S->events[(size_t)r], S->events[(size_t)r + 1]));
            secs[(size_t)p][(size_t)r] = ms * 1e-3;
        }
    }
    if (seconds)
        for (int r = 0; r < repeat; r++) {
            double worst = 0.0;
            for (int p = 0; p < n; p++) if (secs[(size_t)p][(size_t)r] > worst) worst = secs[(size_t)p][(size_t)r];
            seconds[r] = worst;
        }
    return 0;
}

}  // namespace ellspmv
