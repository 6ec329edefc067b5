/*
 * convert.h -- COO -> ELL / CSR conversion of the host programs.
 *
 * Produces, bit for bit, the arrays the reference's converters produce on
 * their default paths (ell_from_coo_size / ell_from_coo, ellspmv.c:931-958,
 * 1081-1127; csr_from_coo_size / csr_from_coo, csrspmv.c:1219-1267,
 * 1390-1475): 0-based, row-major ELL with file-order slots and
 * (min(i, ncols-1), 0.0) padding; CSR sorted by row, file order inside a
 * row, with symmetric expansion for square symmetric input.
 */
#ifndef ELLSPMV_HOST_CONVERT_H
#define ELLSPMV_HOST_CONVERT_H

#include <stdint.h>

#include "idx.h"

struct ell_matrix {
    idx_t num_rows, num_columns;
    idx_t rowsize;      /* K */
    idx_t diagsize;     /* min(rows, cols): only used by the printed model */
    int64_t ellsize;    /* rows * K */
    idx_t *colidx;      /* ellsize */
    double *a;          /* ellsize */
    double *ad;         /* diagsize entries when the diagonal is stored separately, else NULL */
    int64_t *rowcount;  /* entries per row before the padding (kept for --sort-rows) */
};

struct csr_matrix {
    idx_t num_rows, num_columns;
    idx_t rowsizemin, rowsizemax;
    int64_t csrsize;
    int64_t *rowptr;    /* num_rows + 1 */
    idx_t *colidx;      /* csrsize */
    double *a;          /* csrsize */
    idx_t diagsize;     /* num_rows when the diagonal is stored separately, else 0 */
    double *ad;         /* diagsize entries, or NULL */
};

/* rowidx/colidx are 1-based as read from the file.  Return 0 or errno.
 *
 * separate_diagonal: entries with row == column are summed (file order) into
 * `ad` and left out of the ELL/CSR arrays.  For ELL this is what the
 * reference's converters do when their two flags arrive in DECLARED order
 * (ellspmv.c:946-949, 1098-1101); its own main() swaps them (Q1).  For CSR it
 * is csrspmv.c:1249-1252 / 1409-1435 (square matrices; rowsizemin/max then
 * count the diagonal). */
int ell_from_coo(struct ell_matrix *ell, idx_t num_rows, idx_t num_columns, int64_t num_nonzeros,
                 const idx_t *rowidx, const idx_t *colidx, const double *a, int separate_diagonal);
int csr_from_coo(struct csr_matrix *csr, int symmetric, idx_t num_rows, idx_t num_columns,
                 int64_t num_nonzeros, const idx_t *rowidx, const idx_t *colidx, const double *a,
                 int separate_diagonal);
/* --sort-rows for CSR: sort every row by column exactly like the reference's
 * rowsort (csrspmv.c:1269-1388), including its tie order for duplicate columns */
int csr_sort_rows(struct csr_matrix *csr);
/* --sort-rows for ELL: the entries of every row (not the padding) sorted by
 * column with the same order the reference's rowsort gives the CSR rows, so
 * ELL row i == sorted CSR row i followed by padding.  The reference's own ELL
 * --sort-rows hands rowsort per-row counts where it expects offsets and
 * scrambles the matrix (ellspmv.c:1121-1123); this is the intended
 * behaviour. */
int ell_sort_rows(struct ell_matrix *ell);
void ell_free(struct ell_matrix *ell);
void csr_free(struct csr_matrix *csr);

#endif
