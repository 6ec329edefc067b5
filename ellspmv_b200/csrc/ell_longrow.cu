// ell_longrow.cu -- fp64 ELL y <- beta*y + A*x for FEW, LONG rows, bit-exact.
//
// The thread-per-row kernel (ell_kernels.cu) needs ~10^5 rows to keep enough loads in flight
// for HBM; a matrix with a few thousand rows of a few thousand entries each leaves most of the
// GPU idle there (profiles/r2_k_sweep.md).  The reference's loop (ellspmv.c:1146-1151) fixes the
// ORDER of the additions inside a row but says nothing about who loads the operands, so here a
// whole CTA works on a small group of rows:
//
//   layout   row-major, exactly the reference's arrays (slice height 1): entries of a row are
//            contiguous, so 32 lanes reading 32 consecutive slots of one row is a fully
//            coalesced request at any K, and a matrix of 3 rows is not padded to 128;
//   loads    all 128 threads stream the CTA's rows tile by tile (1024 entries), gather x and
//            park the ROUNDED products a*x in shared memory;
//   sums     one lane per row then adds its row's products in slot order with __dadd_rn:
//            mul, then left-to-right adds -- the reference's rounding sequence, bit for bit;
//   overlap  two tiles are in flight in registers (values + indices of tile t+2, values + gathered
//            x of tile t+1) while tile t is being summed, one CTA barrier per tile.
//
// A row is one dependent chain of K additions whoever runs it (the CPU has the same chain), so a
// single very long row is latency-bound by construction; from a few hundred rows on the chains
// of different rows overlap and the kernel is HBM-bound like the others.  ELLSPMV_CUDA_FMA cannot
// contract here (the products are parked rounded): tolerance mode gets the same bits as the
// exact mode.
#include <stdlib.h>

#include "common.cuh"

namespace ellspmv {

constexpr int kLrThreads = 128;
constexpr int kLrU = 8;                          // entries per thread per tile
constexpr int kLrTile = kLrThreads * kLrU;       // 1024 entries per tile, shared by the CTA's rows

template <typename IdxT>
__device__ __forceinline__ void lr_load(const double *vp, const IdxT *cp, double &v, int64_t &c)
{
    asm("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(v) : "l"(vp));
    if (sizeof(IdxT) == 4) { int t; asm("ld.global.nc.L1::no_allocate.s32 %0, [%1];" : "=r"(t) : "l"(cp)); c = t; }
    else { long long t; asm("ld.global.nc.L1::no_allocate.s64 %0, [%1];" : "=l"(t) : "l"(cp)); c = t; }
}

// rpc = rows per CTA = 1 << rshift (1..32); every row gets T_row = 1024 >> rshift slots of a tile
template <typename IdxT>
__global__ void __launch_bounds__(kLrThreads)
ell_longrow_kernel(const EllSpmvArgs a, int rshift)
{
    __shared__ double prod[2][kLrTile + 32];     // row r of a tile starts at r * (T_row + 1)
    const int K = a.rowsize;
    const int rpc = 1 << rshift;
    const int tshift = 10 - rshift;              // log2(T_row)
    const int T_row = 1 << tshift;
    const int tid = threadIdx.x;
    const int64_t row0 = (int64_t)blockIdx.x * rpc;
    const int nrows = (int)((a.num_rows - row0 < rpc) ? a.num_rows - row0 : rpc);
    const int ntiles = (K + T_row - 1) >> tshift;
    const double *__restrict__ x = a.x;
    const IdxT *cols = reinterpret_cast<const IdxT *>(a.cols);

    // this thread's entries of a tile: flat f = u*128 + tid -> row f >> tshift, slot f & (T_row-1)
    // (T_row >= 32 is a multiple of 32, so a warp's 32 lanes share one row: coalesced)
    auto issue_vc = [&](int t, double (&v)[kLrU], int64_t (&c)[kLrU]) {
#pragma unroll
        for (int u = 0; u < kLrU; u++) {
            const int f = u * kLrThreads + tid;
            const int r = f >> tshift, l = (t << tshift) + (f & (T_row - 1));
            v[u] = 0.0; c[u] = -1;
            if (t < ntiles && r < nrows && l < K) {
                const int64_t e = (row0 + r) * (int64_t)K + l;
                lr_load<IdxT>(a.vals + e, cols + e, v[u], c[u]);
            }
        }
    };
    auto issue_x = [&](const int64_t (&c)[kLrU], double (&xv)[kLrU]) {
#pragma unroll
        for (int u = 0; u < kLrU; u++) xv[u] = c[u] >= 0 ? __ldg(x + c[u]) : 0.0;
    };

    double v1[kLrU], x1[kLrU];                   // tile t+1 (at loop entry: tile 0): values and gathered x
    double v2[kLrU]; int64_t c2[kLrU];           // tile t+2 (at loop entry: tile 1): values and indices
    {
        int64_t c1[kLrU];
        issue_vc(0, v1, c1);
        issue_vc(1, v2, c2);
        issue_x(c1, x1);
    }

    // the lane that sums row r: separately stored diagonal first (ellgemvsd / ellgemv16sd orders)
    const bool summer = tid < nrows;
    const int64_t row = row0 + tid;
    // CSR view (api.cu): only the first rowlen[row] slots of a row enter the arithmetic
    const int len = (summer && a.rowlen) ? a.rowlen[row] : K;
    double acc = 0.0, dx = 0.0;
    if (summer && a.ad) {
        dx = __dmul_rn(a.ad[row], __ldg(x + a.row_begin + row));
        if (a.sd_order) acc = dx;
    }

    for (int t = 0; t < ntiles; t++) {
        double *p = prod[t & 1];
#pragma unroll
        for (int u = 0; u < kLrU; u++) {
            const int f = u * kLrThreads + tid;
            p[(f >> tshift) * (T_row + 1) + (f & (T_row - 1))] = __dmul_rn(v1[u], x1[u]);
        }
        __syncthreads();                         // tile t is parked (and tile t-1's sum is over: its buffer is free)
        // next tiles: gathers of t+1 (indices arrived during the previous sum), loads of t+2
#pragma unroll
        for (int u = 0; u < kLrU; u++) v1[u] = v2[u];
        issue_x(c2, x1);
        issue_vc(t + 2, v2, c2);
        if (summer) {
            const int n = (len - (t << tshift) < T_row) ? len - (t << tshift) : T_row;   // may be <= 0: nothing left
            // the next 8 operands leave shared memory while the current 8 are being added: the chain
            // runs at the DADD latency, not DADD + LDS
            const double *q = p + tid * (T_row + 1);
            int l = 0;
            if (n >= 8) {
                double w[8];
#pragma unroll
                for (int j = 0; j < 8; j++) w[j] = q[j];
                for (l = 8; l + 8 <= n; l += 8) {
                    double wn[8];
#pragma unroll
                    for (int j = 0; j < 8; j++) wn[j] = q[l + j];
#pragma unroll
                    for (int j = 0; j < 8; j++) acc = __dadd_rn(acc, w[j]);
#pragma unroll
                    for (int j = 0; j < 8; j++) w[j] = wn[j];
                }
#pragma unroll
                for (int j = 0; j < 8; j++) acc = __dadd_rn(acc, w[j]);
            }
            for (; l < n; l++) acc = __dadd_rn(acc, q[l]);
        }
    }
    if (!summer) return;
    if (a.ad && !a.sd_order) acc = __dadd_rn(dx, acc);
    const double yold = a.beta ? a.y[row] : 0.0;
    const double out = __dadd_rn(yold, acc);
    a.y[row] = out;
    const int64_t g = a.row_begin + row;
    for (int q = 0; q < a.push.num_peers; q++)
        if (g >= a.push.row_lo[q] && g < a.push.row_hi[q]) a.push.x[q][g] = out;
}

// Rows per CTA (1 << rshift).  A tile is 1024 entries; with r rows per CTA each row contributes
// T_row = 1024 / r consecutive slots per tile.  Two opposite costs (profiles/r2_k_sweep.md):
//   * the summing lane of a row runs T_row dependent additions per tile while the loads of the
//     next tiles are in flight: a long T_row (few rows per CTA) makes the kernel chain-bound;
//   * a row's share of a tile is one contiguous piece of HBM: a short T_row (many rows per CTA)
//     means 256-byte pieces scattered over rows that lie K*8 bytes apart, which DRAM serves badly.
// T_row = 128 (8 rows per CTA: 1 KB value pieces, 128-step chains) balances the two; short rows
// (K < 128) take more rows per CTA so that a tile is not mostly empty, and matrices with few
// rows take fewer so that there are ~4 CTAs per SM.
int longrow_rshift(int64_t num_rows, int rowsize, int num_sms)
{
    int rshift = 3;
    while (rshift < 5 && (1024 >> rshift) >= 2 * rowsize) rshift++;       // K <= 64: 16 rows, K <= 32: 32 rows
    while (rshift > 0 && (num_rows >> rshift) < (int64_t)num_sms * 4) rshift--;
    return rshift;
}

cudaError_t launch_ell_longrow(const EllLaunchCfg &cfg, const EllSpmvArgs &args, cudaStream_t stream)
{
    if (args.num_rows <= 0) return cudaSuccess;
    int rshift = longrow_rshift(args.num_rows, args.rowsize, cfg.num_sms > 0 ? cfg.num_sms : 148);
    static const int rshift_env = getenv("ELLSPMV_CUDA_LONGROW_RSHIFT") ? atoi(getenv("ELLSPMV_CUDA_LONGROW_RSHIFT")) : -1;
    if (rshift_env >= 0 && rshift_env <= 5) rshift = rshift_env;     // experiments
    const int64_t grid = (args.num_rows + (1 << rshift) - 1) >> rshift;
    if (grid > 0x7fffffffLL) return cudaErrorInvalidValue;
    if (cfg.idx_bits == 64)
        ell_longrow_kernel<int64_t><<<(unsigned)grid, kLrThreads, 0, stream>>>(args, rshift);
    else
        ell_longrow_kernel<int32_t><<<(unsigned)grid, kLrThreads, 0, stream>>>(args, rshift);
    return cudaGetLastError();
}

}  // namespace ellspmv
