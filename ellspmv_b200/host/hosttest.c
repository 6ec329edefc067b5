/*
 * hosttest.c -- exports the host programs' reader + converters from a
 * shared library (bin/libhost{32,64}.so) so that the CPU test-suite can
 * check them against the golden vectors without a GPU.  Not linked into the
 * programs.
 */
#include <errno.h>
#include <stdlib.h>

#include "convert.h"
#include "hostutil.h"
#include "mtxfile.h"

int host_idx_bits(void) { return IDX_BITS; }
void host_free(void *p) { free(p); }

static int read_coo(const char *path, int gzip, struct mtx_header *h, idx_t **ri, idx_t **ci, double **a,
                    int64_t *lines)
{
    struct mtx_stream *s = mtx_open(path, gzip);
    if (!s) return errno ? errno : EIO;
    int64_t bytes = 0;
    *lines = 0;
    int err = mtx_read_header(s, h, lines, &bytes);
    if (!err && !(h->object == MTX_MATRIX && h->format == MTX_COORDINATE)) err = EINVAL;
    if (err) { mtx_close(s); return err; }
    size_t nz = h->num_nonzeros > 0 ? (size_t)h->num_nonzeros : 1;
    *ri = malloc(nz * sizeof(idx_t)); *ci = malloc(nz * sizeof(idx_t)); *a = malloc(nz * sizeof(double));
    if (!*ri || !*ci || !*a) { mtx_close(s); return ENOMEM; }
    err = (getenv("ELLSPMV_SERIAL_READER") ? mtx_read_coordinate : mtx_read_coordinate_parallel)(s, h, *ri, *ci, *a, lines, &bytes);
    mtx_close(s);
    if (err) { free(*ri); free(*ci); free(*a); }
    return err;
}

/* dims = {rows, cols, nnz, rowsize, ellsize, diagsize, lines_read}; ad may be NULL */
int host_ell_from_file_sd(const char *path, int gzip, int separate_diagonal, int64_t dims[7],
                          void **colidx, double **a, double **ad);

int host_ell_from_file(const char *path, int gzip, int64_t dims[7], void **colidx, double **a)
{
    double *ad = NULL;
    int err = host_ell_from_file_sd(path, gzip, 0, dims, colidx, a, &ad);
    free(ad);
    return err;
}

int host_ell_from_file_sd(const char *path, int gzip, int separate_diagonal, int64_t dims[7],
                          void **colidx, double **a, double **ad)
{
    struct mtx_header h;
    idx_t *ri, *ci; double *v;
    int err = read_coo(path, gzip, &h, &ri, &ci, &v, &dims[6]);
    if (err) return err;
    struct ell_matrix ell;
    /* bit 1 of separate_diagonal asks for --sort-rows as well */
    err = ell_from_coo(&ell, h.num_rows, h.num_columns, h.num_nonzeros, ri, ci, v, separate_diagonal & 1);
    if (!err && (separate_diagonal & 2)) err = ell_sort_rows(&ell);
    free(ri); free(ci); free(v);
    if (err) return err;
    dims[0] = ell.num_rows; dims[1] = ell.num_columns; dims[2] = h.num_nonzeros;
    dims[3] = ell.rowsize; dims[4] = ell.ellsize; dims[5] = ell.diagsize;
    *colidx = ell.colidx; *a = ell.a; *ad = ell.ad;
    free(ell.rowcount);
    return 0;
}

/* dims = {rows, cols, nnz, csrsize, rowsizemin, rowsizemax, lines_read} */
int host_csr_from_file_sd(const char *path, int gzip, int separate_diagonal, int64_t dims[7],
                          int64_t **rowptr, void **colidx, double **a, double **ad);

int host_csr_from_file(const char *path, int gzip, int64_t dims[7], int64_t **rowptr, void **colidx, double **a)
{
    double *ad = NULL;
    int err = host_csr_from_file_sd(path, gzip, 0, dims, rowptr, colidx, a, &ad);
    free(ad);
    return err;
}

int host_csr_from_file_sd(const char *path, int gzip, int separate_diagonal, int64_t dims[7],
                          int64_t **rowptr, void **colidx, double **a, double **ad)
{
    struct mtx_header h;
    idx_t *ri, *ci; double *v;
    int err = read_coo(path, gzip, &h, &ri, &ci, &v, &dims[6]);
    if (err) return err;
    struct csr_matrix csr;
    /* bit 1 of separate_diagonal asks for --sort-rows as well */
    err = csr_from_coo(&csr, h.symmetry == MTX_SYMMETRIC, h.num_rows, h.num_columns, h.num_nonzeros, ri, ci, v,
                       separate_diagonal & 1);
    if (!err && (separate_diagonal & 2)) err = csr_sort_rows(&csr);
    free(ri); free(ci); free(v);
    if (err) return err;
    dims[0] = csr.num_rows; dims[1] = csr.num_columns; dims[2] = h.num_nonzeros;
    dims[3] = csr.csrsize; dims[4] = csr.rowsizemin; dims[5] = csr.rowsizemax;
    *rowptr = csr.rowptr; *colidx = csr.colidx; *a = csr.a; *ad = csr.ad;
    return 0;
}

int host_vector_from_file(const char *path, int gzip, int64_t n, double *v)
{
    return read_vector_file(path, gzip, (idx_t)n, v, 0);
}
