// ell_kernels_len.cu -- the thread-per-row kernel with per-row lengths (LEN): the sliced-ELL
// view of a CSR matrix whose rows differ in length (api.cu::csr_build_ell_view).  A slot past a
// row's end is loaded but never enters the arithmetic, so the result is csrgemv's
// (csrspmv.c:1588-1593) bit for bit.  One row per thread; explicit indices, one pattern id per
// group or one per thread (a CSR stencil: its boundary rows are shorter, their unused slots
// repeat the last column, so they are just more kinds of rows for the dictionary); K = 5 and
// K = 27 -- the two stencils of BASELINE.json in CSR form -- are unrolled like their ELL
// counterparts (the run-time-K form holds 8 slots of loads in 56 registers and ran the 5-point
// CSR matrix at 1.0 ms; see profiles/r2_csr_structured.md).
#include "ell_thread.cuh"

namespace ellspmv {

template <typename IdxT, int KU, bool FMA>
static cudaError_t launch_len_pat(const EllSpmvArgs &args, bool yvec, cudaLaunchConfig_t &lc)
{
    if (args.patinfo || args.vpat) return cudaErrorInvalidValue;
    if (args.patlane) {
        if (yvec) return cudaLaunchKernelEx(&lc, ell_thread_kernel<IdxT, 1, KU, FMA, true, 0, 3, true>, args);
        return cudaLaunchKernelEx(&lc, ell_thread_kernel<IdxT, 1, KU, FMA, false, 0, 3, true>, args);
    }
    if (args.patid) {
        if (yvec) return cudaLaunchKernelEx(&lc, ell_thread_kernel<IdxT, 1, KU, FMA, true, 0, 1, true>, args);
        return cudaLaunchKernelEx(&lc, ell_thread_kernel<IdxT, 1, KU, FMA, false, 0, 1, true>, args);
    }
    if (yvec) return cudaLaunchKernelEx(&lc, ell_thread_kernel<IdxT, 1, KU, FMA, true, 0, 0, true>, args);
    return cudaLaunchKernelEx(&lc, ell_thread_kernel<IdxT, 1, KU, FMA, false, 0, 0, true>, args);
}

template <typename IdxT, bool FMA>
static cudaError_t launch_len_k(const EllSpmvArgs &args, bool yvec, cudaLaunchConfig_t &lc)
{
    switch (args.rowsize) {
    case 5:  return launch_len_pat<IdxT, 5, FMA>(args, yvec, lc);
    case 27: return launch_len_pat<IdxT, 27, FMA>(args, yvec, lc);
    default: return launch_len_pat<IdxT, 0, FMA>(args, yvec, lc);
    }
}

cudaError_t launch_ell_thread_len(const EllLaunchCfg &cfg, const EllSpmvArgs &args, bool yvec, cudaLaunchConfig_t &lc)
{
    if (cfg.rows_per_thread != 1 || !args.rowlen || args.sync.local_flags) return cudaErrorInvalidValue;
    if (cfg.idx_bits == 64) return cfg.fma ? launch_len_k<int64_t, true>(args, yvec, lc) : launch_len_k<int64_t, false>(args, yvec, lc);
    return cfg.fma ? launch_len_k<int32_t, true>(args, yvec, lc) : launch_len_k<int32_t, false>(args, yvec, lc);
}

}  // namespace ellspmv
