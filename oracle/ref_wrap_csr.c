/*
 * ref_wrap_csr.c -- function-level access to the UNMODIFIED reference
 * csrspmv.c (see ref_wrap_ell.c for the rules).  Built once per
 * IDXTYPEWIDTH into libref_csr32.so / libref_csr64.so (separate shared
 * objects: the two reference files define clashing global symbols).
 */
#define main csrspmv_reference_main
#include "csrspmv.c"
#undef main

int ref_idx_bytes(void) { return (int)sizeof(idx_t); }

int ref_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

int ref_csr_from_coo_size(
    int symmetric, int64_t num_rows, int64_t num_columns, int64_t num_nonzeros,
    const void *rowidx, const void *colidx, const double *a,
    int64_t *rowptr, int64_t *csrsize, int64_t *rowsizemin, int64_t *rowsizemax)
{
    idx_t lo = 0, hi = 0, ds = 0;
    int err = csr_from_coo_size(
        symmetric ? mtxsymmetric : mtxgeneral,
        (idx_t)num_rows, (idx_t)num_columns, num_nonzeros,
        (const idx_t *)rowidx, (const idx_t *)colidx, a,
        rowptr, csrsize, &lo, &hi, &ds, false, partition_rows);
    *rowsizemin = lo; *rowsizemax = hi;
    return err;
}

/* arrays zero-filled first (csrspmv.c:2122-2204, Q19) */
int ref_csr_from_coo(
    int symmetric, int64_t num_rows, int64_t num_columns, int64_t num_nonzeros,
    const void *rowidx, const void *colidx, const double *a,
    int64_t *rowptr, int64_t csrsize, int64_t rowsizemin, int64_t rowsizemax,
    void *csrcolidx, double *csra)
{
    idx_t *cc = (idx_t *)csrcolidx;
    for (int64_t k = 0; k < csrsize; k++) { cc[k] = 0; csra[k] = 0; }
    return csr_from_coo(
        symmetric ? mtxsymmetric : mtxgeneral,
        (idx_t)num_rows, (idx_t)num_columns, num_nonzeros,
        (const idx_t *)rowidx, (const idx_t *)colidx, a,
        rowptr, csrsize, (idx_t)rowsizemin, (idx_t)rowsizemax,
        cc, csra, NULL, false, false, partition_rows);
}

/* called from every thread of a parallel region, barriers on both sides
 * of the timed call like csrspmv.c:2838-2872 */
int ref_csrgemv(
    int64_t num_rows, double *y, int64_t num_columns, const double *x,
    int64_t csrsize, int64_t rowsizemin, int64_t rowsizemax,
    const int64_t *rowptr, const void *colidx, const double *a,
    int repeat, double *seconds)
{
    int err = 0;
    struct timespec t0, t1;
#ifdef _OPENMP
    #pragma omp parallel
#endif
    for (int r = 0; r < repeat; r++) {
#ifdef _OPENMP
        #pragma omp barrier
        #pragma omp master
#endif
        clock_gettime(CLOCK_MONOTONIC, &t0);
#ifdef _OPENMP
        #pragma omp barrier
#endif
        int priverr = csrgemv(
            (idx_t)num_rows, y, (idx_t)num_columns, x, csrsize,
            (idx_t)rowsizemin, (idx_t)rowsizemax, rowptr, (const idx_t *)colidx, a);
#ifdef _OPENMP
        #pragma omp barrier
        #pragma omp master
#endif
        {
            clock_gettime(CLOCK_MONOTONIC, &t1);
            if (seconds) seconds[r] = timespec_duration(t0, t1);
        }
        if (priverr) {
#ifdef _OPENMP
            #pragma omp critical
#endif
            err = priverr;
        }
    }
    return err;
}

/* ---- separate diagonal (square matrices; csrspmv.c:1249-1252, 1428-1435) ---- */
int ref_csr_from_coo_sd_size(
    int symmetric, int64_t num_rows, int64_t num_columns, int64_t num_nonzeros,
    const void *rowidx, const void *colidx, const double *a,
    int64_t *rowptr, int64_t *csrsize, int64_t *rowsizemin, int64_t *rowsizemax, int64_t *diagsize)
{
    idx_t lo = 0, hi = 0, ds = 0;
    int err = csr_from_coo_size(
        symmetric ? mtxsymmetric : mtxgeneral,
        (idx_t)num_rows, (idx_t)num_columns, num_nonzeros,
        (const idx_t *)rowidx, (const idx_t *)colidx, a,
        rowptr, csrsize, &lo, &hi, &ds, true, partition_rows);
    *rowsizemin = lo; *rowsizemax = hi; *diagsize = ds;
    return err;
}

int ref_csr_from_coo_sd(
    int symmetric, int64_t num_rows, int64_t num_columns, int64_t num_nonzeros,
    const void *rowidx, const void *colidx, const double *a,
    int64_t *rowptr, int64_t csrsize, int64_t rowsizemin, int64_t rowsizemax,
    void *csrcolidx, double *csra, double *csrad)
{
    idx_t *cc = (idx_t *)csrcolidx;
    for (int64_t k = 0; k < csrsize; k++) { cc[k] = 0; csra[k] = 0; }
    for (int64_t k = 0; k < num_rows; k++) csrad[k] = 0;
    return csr_from_coo(
        symmetric ? mtxsymmetric : mtxgeneral,
        (idx_t)num_rows, (idx_t)num_columns, num_nonzeros,
        (const idx_t *)rowidx, (const idx_t *)colidx, a,
        rowptr, csrsize, (idx_t)rowsizemin, (idx_t)rowsizemax,
        cc, csra, csrad, true, false, partition_rows);
}

int ref_csrgemvsd(
    int64_t num_rows, double *y, int64_t num_columns, const double *x,
    int64_t csrsize, int64_t rowsizemin, int64_t rowsizemax,
    const int64_t *rowptr, const void *colidx, const double *a, const double *ad)
{
    int err = 0;
#ifdef _OPENMP
    #pragma omp parallel
#endif
    {
        int priverr = csrgemvsd(
            (idx_t)num_rows, y, (idx_t)num_columns, x, csrsize,
            (idx_t)rowsizemin, (idx_t)rowsizemax, rowptr, (const idx_t *)colidx, a, ad);
        if (priverr) {
#ifdef _OPENMP
            #pragma omp critical
#endif
            err = priverr;
        }
    }
    return err;
}

/* ---- --sort-rows: the reference's rowsort on a CSR matrix (csrspmv.c:1269-1388) ---- */
int ref_rowsort(int64_t num_rows, int64_t num_columns, int64_t *rowptr, int64_t rowsizemax,
                void *colidx, double *a)
{
    return rowsort((idx_t)num_rows, (idx_t)num_columns, rowptr, (idx_t)rowsizemax, (idx_t *)colidx, a);
}
