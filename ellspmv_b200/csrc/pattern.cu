// pattern.cu -- offset patterns: column indices that need not be streamed.
//
// The SpMV is HBM-bound and 4 of every 12 streamed bytes (8 of 16 with wide
// indices) are column indices.  In a matrix that comes from a structured grid,
// almost every row has the SAME column offsets relative to its own index:
// colidx[i][l] = i + d[l] (2D 5-point Laplacian: d = -n, -1, 0, +1, +n).  At
// upload the library looks for that: a group of 32*R consecutive rows (one warp of
// the thread-per-row kernel, R rows per thread) is "patterned" when all of its rows share one offset
// vector d[0..K-1]; the up to 16 most common vectors form a dictionary, every
// group gets a one-byte pattern id (0xff = none), and for a patterned group the
// kernel computes col = row + d[l] from the dictionary (a warp-uniform, L1-resident
// load) instead of loading the 128/256-byte line of indices from HBM.
//
// Lane masks (ELLSPMV_CUDA_PATTERN_MASKS, opt-in): a grid boundary puts ONE deviating row into
// an otherwise regular group (its entries shift left when a neighbour is missing,
// ellspmv.c:1102-1117), and by default that voids the group.  With the flag a group stays on
// the pattern when at most 4 of its 32 lanes deviate: they are flagged in a 32-bit mask (read
// together with the id as one 64-bit word per warp) and their rows are recomputed from the
// explicit indices by the whole warp after the main loop.  27-point 384^3: 83 % -> 99.5 % of
// the rows stop streaming indices -- and the kernel gets SLOWER (2.11 -> 2.25 ms): without
// the index stream it is latency-bound, and the extra round of loads per masked group costs
// more than the 6 % of bytes it saves.  Three variants measured, all slower than whole groups
// (profiles/r2_offset_patterns.md); hence opt-in.
//
// Lane patterns (default where they pay, ELLSPMV_CUDA_NO_PATTERN_LANES turns them off): the same
// boundary row costs nothing when every THREAD has its own pattern id.  The dictionary then
// holds the up to 32 most common per-thread offset vectors (interior rows and the few kinds of
// boundary rows alike), a group is patterned when each of its 32 threads matches some entry,
// and the kernel is the whole-group one with a per-thread id (no tail, no extra round trip:
// id -> offsets from L1 -> gather).  One byte per thread instead of one per group, so it is
// kept only when the index bytes it saves exceed twice that (27-point 384^3: 18 B/row saved).
//
// This is a device-layout choice like the 64->32-bit index narrowing: the column
// used for every entry is the stored one (every group is verified against the
// dictionary entry by entry, not by hash), so results stay bit-exact; the explicit
// index array is kept (download, the other kernels and the un-patterned groups use
// it).  Only the index stream is touched: values are always read, so matrices with
// variable coefficients on a regular grid profit just the same.
#include <algorithm>
#include <unordered_map>
#include <vector>

#include "common.cuh"

namespace ellspmv {

namespace {

struct PatHashes { unsigned long long h[kMaxLanePatterns]; };

__device__ __forceinline__ unsigned long long pat_mix(unsigned long long h, long long d)
{
    h = (h ^ (unsigned long long)d) * 0x9E3779B97F4A7C15ull;
    return h ^ (h >> 29);
}


// offsets of the lane's R rows at slot l, relative to each row: equal for all R rows, or not
template <typename IdxT>
__device__ __forceinline__ bool lane_offset(const IdxT *__restrict__ cols, const EllLayout &lay, int R,
                                            int64_t row_begin, int64_t row, int l, long long *d)
{
    *d = (long long)cols[lay.offset(row, l)] - (row_begin + row);
    bool same = true;
    for (int r = 1; r < R; r++) same = same && (long long)cols[lay.offset(row + r, l)] - (row_begin + row + r) == *d;
    return same;
}

// same for the values (only when the handle looks for value patterns): the bit pattern of the
// lane's R rows at slot l, equal for all R rows, or not
__device__ __forceinline__ bool lane_value(const double *__restrict__ vals, const EllLayout &lay, int R,
                                           int64_t row, int l, long long *bits)
{
    *bits = __double_as_longlong(vals[lay.offset(row, l)]);
    bool same = true;
    for (int r = 1; r < R; r++) same = same && __double_as_longlong(vals[lay.offset(row + r, l)]) == *bits;
    return same;
}

// one warp per group of 32*R rows (the rows one warp of the thread-per-row kernel owns; lane j
// holds rows j*R .. j*R+R-1 of it).  Every lane hashes its own offset vector; the group's
// signature is the hash that at least 32 - kPatMaxExplicit lanes share, 0 if there is none.
template <typename IdxT>
__global__ void __launch_bounds__(256)
pat_signature_kernel(const IdxT *__restrict__ cols, const double *__restrict__ vals, EllLayout lay, int R, int64_t row_begin,
                     int64_t num_groups, int max_explicit, unsigned long long *__restrict__ sig)
{
    const int lane = threadIdx.x & 31;
    const int64_t g = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    if (g >= num_groups) return;
    const int64_t row = (g * 32 + lane) * R;
    const bool whole = __all_sync(0xffffffffu, row + R <= lay.num_rows);  // the ragged last group stays explicit
    unsigned long long h = 0x243F6A8885A308D3ull;
    bool lane_ok = whole;
    for (int l = 0; l < lay.rowsize && lane_ok; l++) {
        long long d;
        lane_ok = lane_offset(cols, lay, R, row_begin, row, l, &d);
        h = pat_mix(h, d);
        if (vals && lane_ok) { lane_ok = lane_value(vals, lay, R, row, l, &d); h = pat_mix(h, d); }
    }
    h |= 1ull;
    if (!lane_ok) h = 2ull * (unsigned long long)(lane + 1);               // even: never equal to a real hash or each other
    const unsigned peers = __match_any_sync(0xffffffffu, h);
    const unsigned key = ((unsigned)__popc(peers) << 8) | (unsigned)(31 - lane);
    const unsigned best = __reduce_max_sync(0xffffffffu, key);
    const unsigned long long hmaj = __shfl_sync(0xffffffffu, h, 31 - (int)(best & 0xffu));
    if (lane == 0) sig[g] = (whole && (hmaj & 1ull) && (int)(best >> 8) >= 32 - max_explicit) ? hmaj : 0ull;
}

__global__ void pat_sample_kernel(const unsigned long long *__restrict__ sig, int64_t stride, int64_t n,
                                  unsigned long long *__restrict__ out)
{
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) out[i] = sig[i * stride];
}

// dictionary entry p = the offset vector most lanes of its representative group share (the lane
// is found again here: the first lane whose own hash is the group's signature)
template <typename IdxT>
__global__ void pat_extract_kernel(const IdxT *__restrict__ cols, const double *__restrict__ vals, EllLayout lay, int R, int64_t row_begin,
                                   const long long *__restrict__ reps, int npat,
                                   long long *__restrict__ pat, double *__restrict__ vpat)
{
    const int p = blockIdx.x;
    if (p >= npat) return;
    __shared__ int s_lane;
    if (threadIdx.x < 32) {
        const int lane = threadIdx.x;
        const int64_t row = (reps[p] * 32 + lane) * R;
        unsigned long long h = 0x243F6A8885A308D3ull;
        bool lane_ok = true;
        for (int l = 0; l < lay.rowsize && lane_ok; l++) {
            long long d;
            lane_ok = lane_offset(cols, lay, R, row_begin, row, l, &d);
            h = pat_mix(h, d);
            if (vals && lane_ok) { lane_ok = lane_value(vals, lay, R, row, l, &d); h = pat_mix(h, d); }
        }
        h |= 1ull;
        if (!lane_ok) h = 2ull * (unsigned long long)(lane + 1);
        const unsigned peers = __match_any_sync(0xffffffffu, h);
        const unsigned key = ((unsigned)__popc(peers) << 8) | (unsigned)(31 - lane);
        const unsigned best = __reduce_max_sync(0xffffffffu, key);
        if (lane == 0) s_lane = 31 - (int)(best & 0xffu);
    }
    __syncthreads();
    const int64_t row = (reps[p] * 32 + s_lane) * R;
    for (int l = threadIdx.x; l < lay.rowsize; l += blockDim.x) {
        pat[(int64_t)p * lay.rowsize + l] = (long long)cols[lay.offset(row, l)] - (row_begin + row);
        if (vals) vpat[(int64_t)p * lay.rowsize + l] = vals[lay.offset(row, l)];
    }
}

// final word: every lane is checked entry by entry against dictionary pattern p; the lanes that
// differ go into the group's mask (they keep their explicit indices), and the group takes the
// pattern if at most kPatMaxExplicit lanes do
template <typename IdxT>
__global__ void __launch_bounds__(256)
pat_classify_kernel(const IdxT *__restrict__ cols, const double *__restrict__ vals, EllLayout lay, int R, int64_t row_begin,
                    int64_t num_groups, int max_explicit, const unsigned long long *__restrict__ sig, PatHashes hashes, int npat,
                    const long long *__restrict__ pat, const double *__restrict__ vpat, unsigned char *__restrict__ patid,
                    unsigned long long *__restrict__ patinfo, unsigned long long *__restrict__ covered)
{
    const int lane = threadIdx.x & 31;
    const int64_t g = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    if (g >= num_groups) return;
    const unsigned long long s = sig[g];
    int p = -1;
    if (s != 0)
        for (int q = 0; q < npat; q++)
            if (hashes.h[q] == s) { p = q; break; }
    bool ok = p >= 0;
    unsigned mask = 0;
    if (ok) {
        const int64_t row = (g * 32 + lane) * R;
        const long long *d = pat + (int64_t)p * lay.rowsize;
        bool mine = true;
        for (int l = 0; l < lay.rowsize && mine; l++)
            for (int r = 0; r < R; r++) {
                mine = mine && (long long)cols[lay.offset(row + r, l)] - (row_begin + row + r) == d[l];
                if (vals) mine = mine && __double_as_longlong(vals[lay.offset(row + r, l)]) ==
                                             __double_as_longlong(vpat[(int64_t)p * lay.rowsize + l]);
            }
        mask = ~__ballot_sync(0xffffffffu, mine);
        ok = __popc(mask) <= max_explicit;
    }
    if (lane == 0) {
        patid[g] = ok ? (unsigned char)p : (unsigned char)0xff;
        // what the kernel reads, ONE load per warp: pattern id in the low byte, lane mask on top
        patinfo[g] = ok ? (((unsigned long long)mask << 32) | (unsigned long long)p) : 0xffull;
        if (ok) { atomicAdd(covered, 1ull); atomicAdd(covered + 1, (unsigned long long)__popc(mask)); }
    }
}

template <typename IdxT>
cudaError_t pattern_build_typed(PatternSet *ps, const IdxT *cols, const double *vals, const EllLayout &lay, int R,
                                int64_t row_begin, int max_explicit, cudaStream_t stream)
{
    const int64_t groups = lay.padded_rows() / (32 * R);
    const int K = lay.rowsize;
    cudaError_t e;
    unsigned long long *sig = nullptr, *sample = nullptr, *covered = nullptr;
    long long *reps = nullptr;
    auto cleanup = [&]() { cudaFree(sig); cudaFree(sample); cudaFree(covered); cudaFree(reps); };
    if ((e = cudaMalloc(&sig, (size_t)groups * 8)) != cudaSuccess) return e;
    const unsigned grid = (unsigned)((groups * 32 + 255) / 256);
    pat_signature_kernel<IdxT><<<grid, 256, 0, stream>>>(cols, vals, lay, R, row_begin, groups, max_explicit, sig);
    if ((e = cudaGetLastError()) != cudaSuccess) { cleanup(); return e; }

    // dictionary candidates: the most common signatures of a strided sample (the
    // classification below verifies every group, so sampling cannot cost correctness)
    const int64_t want = 1 << 20;
    const int64_t stride = (groups + want - 1) / want;
    const int64_t n = (groups + stride - 1) / stride;
    if ((e = cudaMalloc(&sample, (size_t)n * 8)) != cudaSuccess) { cleanup(); return e; }
    pat_sample_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(sig, stride, n, sample);
    std::vector<unsigned long long> hs((size_t)n);
    if ((e = cudaMemcpyAsync(hs.data(), sample, (size_t)n * 8, cudaMemcpyDeviceToHost, stream)) != cudaSuccess ||
        (e = cudaStreamSynchronize(stream)) != cudaSuccess) { cleanup(); return e; }
    struct Cand { unsigned long long h; int64_t count, first; };
    std::unordered_map<unsigned long long, Cand> seen;
    for (int64_t i = 0; i < n; i++) {
        if (hs[(size_t)i] == 0) continue;
        auto it = seen.find(hs[(size_t)i]);
        if (it == seen.end()) seen.emplace(hs[(size_t)i], Cand{hs[(size_t)i], 1, i * stride});
        else it->second.count++;
    }
    std::vector<Cand> cands;
    cands.reserve(seen.size());
    for (auto &kv : seen) cands.push_back(kv.second);
    std::sort(cands.begin(), cands.end(), [](const Cand &a, const Cand &b) {
        return a.count != b.count ? a.count > b.count : a.first < b.first;
    });
    const int npat = (int)std::min<size_t>(cands.size(), (size_t)kMaxPatterns);
    if (npat == 0) { cleanup(); return cudaSuccess; }

    PatHashes hashes = {};
    long long hreps[kMaxPatterns] = {};
    for (int p = 0; p < npat; p++) { hashes.h[p] = cands[(size_t)p].h; hreps[p] = cands[(size_t)p].first; }
    if ((e = cudaMalloc(&reps, sizeof(hreps))) != cudaSuccess) { cleanup(); return e; }
    if ((e = cudaMalloc(&covered, 16)) != cudaSuccess) { cleanup(); return e; }
    if ((e = cudaMalloc(&ps->pat, (size_t)kMaxLanePatterns * K * 8)) != cudaSuccess) { cleanup(); return e; }
    if (vals && (e = cudaMalloc(&ps->vpat, (size_t)kMaxLanePatterns * K * 8)) != cudaSuccess) { cleanup(); return e; }
    if ((e = cudaMalloc(&ps->patid, (size_t)groups)) != cudaSuccess) { cleanup(); return e; }
    if ((e = cudaMalloc(&ps->patinfo, (size_t)groups * sizeof(unsigned long long))) != cudaSuccess) { cleanup(); return e; }
    if ((e = cudaMemcpyAsync(reps, hreps, sizeof(hreps), cudaMemcpyHostToDevice, stream)) != cudaSuccess ||
        (e = cudaMemsetAsync(covered, 0, 16, stream)) != cudaSuccess ||
        (e = cudaMemsetAsync(ps->pat, 0, (size_t)kMaxLanePatterns * K * 8, stream)) != cudaSuccess) { cleanup(); return e; }
    pat_extract_kernel<IdxT><<<npat, 128, 0, stream>>>(cols, vals, lay, R, row_begin, reps, npat, ps->pat, ps->vpat);
    if ((e = cudaGetLastError()) != cudaSuccess) { cleanup(); return e; }
    pat_classify_kernel<IdxT><<<grid, 256, 0, stream>>>(cols, vals, lay, R, row_begin, groups, max_explicit, sig, hashes, npat, ps->pat,
                                                        ps->vpat, ps->patid, ps->patinfo, covered);
    unsigned long long hc[2] = {0, 0};
    if ((e = cudaGetLastError()) != cudaSuccess ||
        (e = cudaMemcpyAsync(hc, covered, 16, cudaMemcpyDeviceToHost, stream)) != cudaSuccess ||
        (e = cudaStreamSynchronize(stream)) != cudaSuccess) { cleanup(); return e; }
    cleanup();
    ps->num_patterns = npat;
    ps->groups = groups;
    ps->group_rows = 32 * R;
    ps->covered = (int64_t)hc[0];
    ps->explicit_lanes = (int64_t)hc[1];
    ps->max_explicit = max_explicit;
    ps->bytes = groups * 9 + (int64_t)kMaxLanePatterns * K * 8;
    return cudaSuccess;
}

// ---- lane patterns: one id per thread ---------------------------------------------------------
// where sample i of n looks among `threads` threads: every one when they all fit, else a
// pseudo-random one (a stride would alias with the grid: every 54th row of a 384-row line never
// meets the line's last row)
__host__ __device__ inline int64_t lane_sample_pos(int64_t i, int64_t n, int64_t threads)
{
    if (threads <= n) return i;
    unsigned long long z = (unsigned long long)i + 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z ^= z >> 31;
    return (int64_t)(z % (unsigned long long)threads);
}

// hash of thread t's offset vector (odd), or 0 when its R rows differ or pass the last row
template <typename IdxT>
__device__ __forceinline__ unsigned long long lane_hash(const IdxT *__restrict__ cols, const double *__restrict__ vals,
                                                        const EllLayout &lay, int R, int64_t row_begin, int64_t t)
{
    const int64_t row = t * R;
    if (row + R > lay.num_rows) return 0ull;
    unsigned long long h = 0x243F6A8885A308D3ull;
    for (int l = 0; l < lay.rowsize; l++) {
        long long d;
        if (!lane_offset(cols, lay, R, row_begin, row, l, &d)) return 0ull;
        h = pat_mix(h, d);
        if (vals) {
            if (!lane_value(vals, lay, R, row, l, &d)) return 0ull;
            h = pat_mix(h, d);
        }
    }
    return h | 1ull;
}

template <typename IdxT>
__global__ void pat_lane_sample_kernel(const IdxT *__restrict__ cols, const double *__restrict__ vals, EllLayout lay, int R,
                                       int64_t row_begin, int64_t n, int64_t threads, unsigned long long *__restrict__ out)
{
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) out[i] = lane_hash(cols, vals, lay, R, row_begin, lane_sample_pos(i, n, threads));
}

// dictionary entry p = the offset vector of its representative thread
template <typename IdxT>
__global__ void pat_lane_extract_kernel(const IdxT *__restrict__ cols, const double *__restrict__ vals, EllLayout lay, int R,
                                        int64_t row_begin, const long long *__restrict__ reps, int npat,
                                        long long *__restrict__ pat, double *__restrict__ vpat)
{
    const int p = blockIdx.x;
    if (p >= npat) return;
    const int64_t row = reps[p] * R;
    for (int l = threadIdx.x; l < lay.rowsize; l += blockDim.x) {
        pat[(int64_t)p * lay.rowsize + l] = (long long)cols[lay.offset(row, l)] - (row_begin + row);
        if (vals) vpat[(int64_t)p * lay.rowsize + l] = vals[lay.offset(row, l)];
    }
}

// one warp per group: every thread looks its vector up by hash, verifies it entry by entry, and
// the group is patterned when all 32 succeed -- else all 32 ids are 0xff (the kernel's branch on
// the id is warp-uniform by construction)
template <typename IdxT>
__global__ void __launch_bounds__(256)
pat_lane_classify_kernel(const IdxT *__restrict__ cols, const double *__restrict__ vals, EllLayout lay, int R, int64_t row_begin,
                         int64_t num_groups, PatHashes hashes, int npat, const long long *__restrict__ pat,
                         const double *__restrict__ vpat, unsigned char *__restrict__ patlane,
                         unsigned long long *__restrict__ covered)
{
    const int lane = threadIdx.x & 31;
    const int64_t g = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    if (g >= num_groups) return;
    const int64_t t = g * 32 + lane;
    const unsigned long long h = lane_hash(cols, vals, lay, R, row_begin, t);
    int p = -1;
    if (h != 0ull)
        for (int q = 0; q < npat; q++)
            if (hashes.h[q] == h) { p = q; break; }
    bool mine = p >= 0;
    if (mine) {
        const int64_t row = t * R;
        const long long *d = pat + (int64_t)p * lay.rowsize;
        for (int l = 0; l < lay.rowsize && mine; l++)
            for (int r = 0; r < R; r++) {
                mine = mine && (long long)cols[lay.offset(row + r, l)] - (row_begin + row + r) == d[l];
                if (vals) mine = mine && __double_as_longlong(vals[lay.offset(row + r, l)]) ==
                                             __double_as_longlong(vpat[(int64_t)p * lay.rowsize + l]);
            }
    }
    const bool ok = __all_sync(0xffffffffu, mine);
    patlane[t] = ok ? (unsigned char)p : (unsigned char)0xff;
    if (ok && lane == 0) atomicAdd(covered, 1ull);
}

template <typename IdxT>
cudaError_t pattern_build_lanes_typed(PatternSet *ps, const IdxT *cols, const double *vals, const EllLayout &lay, int R,
                                      int64_t row_begin, cudaStream_t stream)
{
    const int64_t groups = lay.padded_rows() / (32 * R);
    const int64_t threads = groups * 32;
    const int K = lay.rowsize;
    cudaError_t e;
    unsigned long long *sample = nullptr, *covered = nullptr;
    long long *reps = nullptr;
    auto cleanup = [&]() { cudaFree(sample); cudaFree(covered); cudaFree(reps); };
    const int64_t n = std::min<int64_t>(threads, 1 << 20);
    if ((e = cudaMalloc(&sample, (size_t)n * 8)) != cudaSuccess) return e;
    pat_lane_sample_kernel<IdxT><<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(cols, vals, lay, R, row_begin, n, threads, sample);
    std::vector<unsigned long long> hs((size_t)n);
    if ((e = cudaGetLastError()) != cudaSuccess ||
        (e = cudaMemcpyAsync(hs.data(), sample, (size_t)n * 8, cudaMemcpyDeviceToHost, stream)) != cudaSuccess ||
        (e = cudaStreamSynchronize(stream)) != cudaSuccess) { cleanup(); return e; }
    // the most common vectors of the sample, ties by the lowest thread that shows them
    struct Cand { unsigned long long h; int64_t count, first; };
    std::unordered_map<unsigned long long, Cand> seen;
    for (int64_t i = 0; i < n; i++) {
        if (hs[(size_t)i] == 0) continue;
        const int64_t t = lane_sample_pos(i, n, threads);
        auto it = seen.find(hs[(size_t)i]);
        if (it == seen.end()) seen.emplace(hs[(size_t)i], Cand{hs[(size_t)i], 1, t});
        else { it->second.count++; if (t < it->second.first) it->second.first = t; }
    }
    std::vector<Cand> cands;
    cands.reserve(seen.size());
    for (auto &kv : seen) cands.push_back(kv.second);
    std::sort(cands.begin(), cands.end(), [](const Cand &a, const Cand &b) {
        return a.count != b.count ? a.count > b.count : a.first < b.first;
    });
    const int npat = (int)std::min<size_t>(cands.size(), (size_t)kMaxLanePatterns);
    if (npat == 0) { cleanup(); return cudaSuccess; }
    PatHashes hashes = {};
    long long hreps[kMaxLanePatterns] = {};
    for (int p = 0; p < npat; p++) { hashes.h[p] = cands[(size_t)p].h; hreps[p] = cands[(size_t)p].first; }
    if ((e = cudaMalloc(&reps, sizeof(hreps))) != cudaSuccess) { cleanup(); return e; }
    if ((e = cudaMalloc(&covered, 8)) != cudaSuccess) { cleanup(); return e; }
    if ((e = cudaMalloc(&ps->pat, (size_t)kMaxLanePatterns * K * 8)) != cudaSuccess) { cleanup(); return e; }
    if (vals && (e = cudaMalloc(&ps->vpat, (size_t)kMaxLanePatterns * K * 8)) != cudaSuccess) { cleanup(); return e; }
    if ((e = cudaMalloc(&ps->patlane, (size_t)threads)) != cudaSuccess) { cleanup(); return e; }
    if ((e = cudaMemcpyAsync(reps, hreps, sizeof(hreps), cudaMemcpyHostToDevice, stream)) != cudaSuccess ||
        (e = cudaMemsetAsync(covered, 0, 8, stream)) != cudaSuccess ||
        (e = cudaMemsetAsync(ps->pat, 0, (size_t)kMaxLanePatterns * K * 8, stream)) != cudaSuccess) { cleanup(); return e; }
    pat_lane_extract_kernel<IdxT><<<npat, 128, 0, stream>>>(cols, vals, lay, R, row_begin, reps, npat, ps->pat, ps->vpat);
    if ((e = cudaGetLastError()) != cudaSuccess) { cleanup(); return e; }
    pat_lane_classify_kernel<IdxT><<<(unsigned)((threads + 255) / 256), 256, 0, stream>>>(cols, vals, lay, R, row_begin, groups, hashes,
                                                                                          npat, ps->pat, ps->vpat, ps->patlane, covered);
    unsigned long long hc = 0;
    if ((e = cudaGetLastError()) != cudaSuccess ||
        (e = cudaMemcpyAsync(&hc, covered, 8, cudaMemcpyDeviceToHost, stream)) != cudaSuccess ||
        (e = cudaStreamSynchronize(stream)) != cudaSuccess) { cleanup(); return e; }
    cleanup();
    ps->num_patterns = npat;
    ps->groups = groups;
    ps->group_rows = 32 * R;
    ps->covered = (int64_t)hc;
    ps->bytes = threads + (int64_t)kMaxLanePatterns * K * 8;
    return cudaSuccess;
}

}  // namespace

void pattern_free(PatternSet *ps)
{
    cudaFree(ps->patid);
    cudaFree(ps->patinfo);
    cudaFree(ps->patlane);
    cudaFree(ps->vpat);
    cudaFree(ps->pat);
    *ps = PatternSet{};
}

// one search, with the values in the signatures or without: group ids first, then (lanes) one id
// per thread when that wins back more bytes than twice the byte per thread the ids cost
static cudaError_t pattern_search(PatternSet *ps, int idx_bits, const void *cols, const double *vals, const EllLayout &lay,
                                  int R, int64_t row_begin, int max_explicit, bool lanes, cudaStream_t stream)
{
    *ps = PatternSet{};
    cudaError_t e = idx_bits == 64 ? pattern_build_typed<int64_t>(ps, (const int64_t *)cols, vals, lay, R, row_begin, max_explicit, stream)
                                   : pattern_build_typed<int32_t>(ps, (const int32_t *)cols, vals, lay, R, row_begin, max_explicit, stream);
    if (e != cudaSuccess) { pattern_free(ps); return e; }
    const int64_t groups = lay.padded_rows() / (32 * R);
    if (lanes && max_explicit == 0 && ps->covered < groups) {
        PatternSet pl = PatternSet{};
        e = idx_bits == 64 ? pattern_build_lanes_typed<int64_t>(&pl, (const int64_t *)cols, vals, lay, R, row_begin, stream)
                           : pattern_build_lanes_typed<int32_t>(&pl, (const int32_t *)cols, vals, lay, R, row_begin, stream);
        if (e != cudaSuccess) { pattern_free(&pl); pattern_free(ps); return e; }
        const int64_t per_entry = idx_bits / 8 + (vals ? 8 : 0);
        const int64_t won = (pl.covered - ps->covered) * 32 * R * lay.rowsize * per_entry;
        if (pl.patlane && won > 2 * groups * 32) { pattern_free(ps); *ps = pl; }
        else pattern_free(&pl);
    }
    return cudaSuccess;
}

// Leaves *ps empty (and returns success) when fewer than 1 group in 10 is patterned:
// the table would cost a byte per group and buy nothing.
// vals != NULL: also look for VALUE patterns -- rows that share offsets AND coefficients (a
// constant-coefficient stencil); the dictionary entry then carries both and a patterned thread
// streams neither indices nor values.  Taken when it covers at least 9/10 of the rows the
// index-only search covers (a matrix with variable coefficients has none and keeps index patterns).
cudaError_t pattern_build(PatternSet *ps, int idx_bits, const void *cols, const double *vals, const EllLayout &lay,
                          int rows_per_thread, int64_t row_begin, int max_explicit, bool lanes, cudaStream_t stream)
{
    *ps = PatternSet{};
    const int R = rows_per_thread;
    if (lay.num_rows < 32 * R || lay.rowsize <= 0 || lay.rowsize > 4096 || lay.slice_rows != kBlockThreads * R)
        return cudaSuccess;
    if (max_explicit < 0) max_explicit = 0;
    if (max_explicit > 8) max_explicit = 8;
    cudaError_t e = pattern_search(ps, idx_bits, cols, nullptr, lay, R, row_begin, max_explicit, lanes, stream);
    if (e != cudaSuccess) return e;
    if (ps->covered * 10 < ps->groups) { pattern_free(ps); return cudaSuccess; }
    if (vals && max_explicit == 0) {
        PatternSet pv = PatternSet{};
        e = pattern_search(&pv, idx_bits, cols, vals, lay, R, row_begin, 0, lanes, stream);
        if (e != cudaSuccess) { pattern_free(ps); return e; }
        if (pv.vpat && pv.covered * 10 >= ps->covered * 9) { pattern_free(ps); *ps = pv; }
        else pattern_free(&pv);
    }
    return cudaSuccess;
}

}  // namespace ellspmv
