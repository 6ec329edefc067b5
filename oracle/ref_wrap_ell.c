/*
 * ref_wrap_ell.c -- function-level access to the UNMODIFIED reference
 * ellspmv.c, compiled where it lies (-I/root/reference); nothing of the
 * reference is copied into this repository.  Output goes to oracle/_ref/
 * only (git-ignored).  TEST INFRASTRUCTURE / CPU BASELINE ONLY.
 *
 * The reference's functions are `static`, so the translation unit is
 * included and thin exported wrappers are added below it.  Built once per
 * IDXTYPEWIDTH (32, 64) into libref_ell32.so / libref_ell64.so.
 */
#define main ellspmv_reference_main
#include "ellspmv.c"
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

int ref_ell_from_coo_size(
    int64_t num_rows, int64_t num_columns, int64_t num_nonzeros,
    const void *rowidx, const void *colidx, const double *a,
    int64_t *rowptr, int64_t *ellsize, int64_t *rowsize, int64_t *diagsize)
{
    idx_t rs = 0, ds = 0;
    int err = ell_from_coo_size(
        (idx_t)num_rows, (idx_t)num_columns, num_nonzeros,
        (const idx_t *)rowidx, (const idx_t *)colidx, a,
        rowptr, ellsize, &rs, &ds, false);
    *rowsize = rs; *diagsize = ds;
    return err;
}

/* arrays are zero-filled first, exactly like main() does (ellspmv.c:1425-1467);
 * the two bools are passed in DECLARED order, both false (Q1). */
int ref_ell_from_coo(
    int64_t num_rows, int64_t num_columns, int64_t num_nonzeros,
    const void *rowidx, const void *colidx, const double *a,
    int64_t *rowptr, int64_t ellsize, int64_t rowsize,
    void *ellcolidx, double *ella, double *ellad)
{
    idx_t *ec = (idx_t *)ellcolidx;
    for (int64_t k = 0; k < ellsize; k++) { ec[k] = 0; ella[k] = 0; }
    int64_t ds = num_rows < num_columns ? num_rows : num_columns;
    for (int64_t k = 0; k < ds; k++) ellad[k] = 0;
    return ell_from_coo(
        (idx_t)num_rows, (idx_t)num_columns, num_nonzeros,
        (const idx_t *)rowidx, (const idx_t *)colidx, a,
        rowptr, ellsize, (idx_t)rowsize, ec, ella, ellad, false, false);
}

/*
 * The reference calls ellgemv from every thread of an enclosing parallel
 * region (ellspmv.c:1821-1843); its orphaned `omp for` splits the rows.
 * seconds[r] is measured like the reference does: t0 on the master before
 * the call, t1 after the barrier (ellspmv.c:1825-1847), plus a barrier in
 * front so that every repeat starts with all threads present.
 */
int ref_ellgemv(
    int64_t num_rows, double *y, int64_t num_columns, const double *x,
    int64_t ellsize, int64_t rowsize, const void *colidx, const double *a,
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
        int priverr = ellgemv(
            (idx_t)num_rows, y, (idx_t)num_columns, x, ellsize,
            (idx_t)rowsize, (const idx_t *)colidx, a);
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

/* first-touch initialisation the way the reference's main() does it
 * (ellspmv.c:1425-1431, 1460-1467, 1502-1505, 1610-1613), so that the CPU
 * baseline sees the same NUMA placement the reference binary would. */
void ref_first_touch(int64_t num_rows, int64_t rowsize, void *colidx, double *a,
                     int64_t num_columns, double *x, double *y)
{
    idx_t *ec = (idx_t *)colidx;
#ifdef _OPENMP
    #pragma omp parallel for
#endif
    for (int64_t i = 0; i < num_rows; i++)
        for (int64_t l = 0; l < rowsize; l++) { ec[i*rowsize+l] = 0; a[i*rowsize+l] = 0; }
#ifdef _OPENMP
    #pragma omp parallel for
#endif
    for (int64_t j = 0; j < num_columns; j++) x[j] = 1.0;
#ifdef _OPENMP
    #pragma omp parallel for
#endif
    for (int64_t i = 0; i < num_rows; i++) y[i] = 0.0;
}

/* ---- separate-diagonal functions, flags in DECLARED order (Q1) ---------- */
int ref_ell_from_coo_sd_size(
    int64_t num_rows, int64_t num_columns, int64_t num_nonzeros,
    const void *rowidx, const void *colidx, const double *a,
    int64_t *rowptr, int64_t *ellsize, int64_t *rowsize, int64_t *diagsize)
{
    idx_t rs = 0, ds = 0;
    int err = ell_from_coo_size(
        (idx_t)num_rows, (idx_t)num_columns, num_nonzeros,
        (const idx_t *)rowidx, (const idx_t *)colidx, a,
        rowptr, ellsize, &rs, &ds, true);
    *rowsize = rs; *diagsize = ds;
    return err;
}

int ref_ell_from_coo_sd(
    int64_t num_rows, int64_t num_columns, int64_t num_nonzeros,
    const void *rowidx, const void *colidx, const double *a,
    int64_t *rowptr, int64_t ellsize, int64_t rowsize,
    void *ellcolidx, double *ella, double *ellad)
{
    idx_t *ec = (idx_t *)ellcolidx;
    for (int64_t k = 0; k < ellsize; k++) { ec[k] = 0; ella[k] = 0; }
    int64_t ds = num_rows < num_columns ? num_rows : num_columns;
    for (int64_t k = 0; k < ds; k++) ellad[k] = 0;
    return ell_from_coo(
        (idx_t)num_rows, (idx_t)num_columns, num_nonzeros,
        (const idx_t *)rowidx, (const idx_t *)colidx, a,
        rowptr, ellsize, (idx_t)rowsize, ec, ella, ellad, true, false);
}

/* which = 0: ellgemvsd, 1: ellgemv16sd (returns EINVAL unless rowsize == 16) */
int ref_ellgemvsd(
    int which, int64_t num_rows, double *y, int64_t num_columns, const double *x,
    int64_t ellsize, int64_t rowsize, const void *colidx, const double *a, const double *ad)
{
    int err = 0;
#ifdef _OPENMP
    #pragma omp parallel
#endif
    {
        int priverr = which
            ? ellgemv16sd((idx_t)num_rows, y, (idx_t)num_columns, x, ellsize, (idx_t)rowsize, (const idx_t *)colidx, a, ad)
            : ellgemvsd((idx_t)num_rows, y, (idx_t)num_columns, x, ellsize, (idx_t)rowsize, (const idx_t *)colidx, a, ad);
        if (priverr) {
#ifdef _OPENMP
            #pragma omp critical
#endif
            err = priverr;
        }
    }
    return err;
}
