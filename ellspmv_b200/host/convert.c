/* convert.c -- see convert.h. */
#include "convert.h"

#include <errno.h>
#include <stdlib.h>
#include <string.h>

void ell_free(struct ell_matrix *ell)
{
    free(ell->colidx);
    free(ell->a);
    free(ell->ad);
    free(ell->rowcount);
    memset(ell, 0, sizeof(*ell));
}

void csr_free(struct csr_matrix *csr)
{
    free(csr->rowptr);
    free(csr->colidx);
    free(csr->a);
    free(csr->ad);
    memset(csr, 0, sizeof(*csr));
}

int ell_from_coo(struct ell_matrix *ell, idx_t num_rows, idx_t num_columns, int64_t num_nonzeros,
                 const idx_t *rowidx, const idx_t *colidx, const double *a, int separate_diagonal)
{
    memset(ell, 0, sizeof(*ell));
    ell->num_rows = num_rows;
    ell->num_columns = num_columns;
    ell->diagsize = num_rows < num_columns ? num_rows : num_columns;

    /* pass 1: entries per row; K is the widest row (ellspmv.c:944-955) */
    int64_t *fill = calloc((size_t)num_rows + 1, sizeof(*fill));
    if (!fill) return ENOMEM;
    int64_t widest = 0;
    for (int64_t k = 0; k < num_nonzeros; k++) {
        if (separate_diagonal && rowidx[k] == colidx[k]) continue;
        int64_t n = ++fill[rowidx[k] - 1];
        if (n > widest) widest = n;
    }
    if (widest > IDX_T_MAX || (widest > 0 && (int64_t)num_rows > INT64_MAX / widest)) { free(fill); return EOVERFLOW; }
    const int64_t K = widest;
    ell->rowsize = (idx_t)K;
    ell->ellsize = (int64_t)num_rows * K;
#if IDX_T_MAX == INT_MAX || (defined(IDXTYPEWIDTH) && IDXTYPEWIDTH == 32)
    /* the reference computes ellsize and slot offsets in idx_t (Q9); a 32-bit
     * build cannot address more than 2^31-1 slots, so say so instead of wrapping */
    if (ell->ellsize > IDX_T_MAX) { free(fill); return EOVERFLOW; }
#endif
    size_t n = ell->ellsize > 0 ? (size_t)ell->ellsize : 1;
    ell->colidx = malloc(n * sizeof(idx_t));
    ell->a = malloc(n * sizeof(double));
    if (!ell->colidx || !ell->a) { free(fill); ell_free(ell); return ENOMEM; }
    if (separate_diagonal) {
        ell->ad = calloc(ell->diagsize > 0 ? (size_t)ell->diagsize : 1, sizeof(double));
        if (!ell->ad) { free(fill); ell_free(ell); return ENOMEM; }
    }

    /* pass 2: scatter in file order (ellspmv.c:1098-1107) */
    memset(fill, 0, ((size_t)num_rows + 1) * sizeof(*fill));
    for (int64_t k = 0; k < num_nonzeros; k++) {
        const int64_t r = (int64_t)rowidx[k] - 1;
        if (separate_diagonal && rowidx[k] == colidx[k]) { ell->ad[r] += a[k]; continue; }
        const int64_t slot = r * K + fill[r]++;
        ell->colidx[slot] = colidx[k] - 1;
        ell->a[slot] = a[k];
    }
    /* pass 3: padding (ellspmv.c:1111-1117) */
#ifdef _OPENMP
#pragma omp parallel for
#endif
    for (int64_t r = 0; r < (int64_t)num_rows; r++) {
        const idx_t padcol = r < (int64_t)num_columns ? (idx_t)r : (idx_t)(num_columns - 1);
        for (int64_t l = fill[r]; l < K; l++) {
            ell->colidx[r * K + l] = padcol;
            ell->a[r * K + l] = 0.0;
        }
    }
    ell->rowcount = fill;      /* entries per row, for ell_sort_rows */
    return 0;
}

int csr_from_coo(struct csr_matrix *csr, int symmetric, idx_t num_rows, idx_t num_columns,
                 int64_t num_nonzeros, const idx_t *rowidx, const idx_t *colidx, const double *a,
                 int separate_diagonal)
{
    memset(csr, 0, sizeof(*csr));
    csr->num_rows = num_rows;
    csr->num_columns = num_columns;
    /* symmetric expansion only for square matrices (csrspmv.c:1244) */
    const int expand = symmetric && num_rows == num_columns;
    /* like the reference, the diagonal is only split off for square matrices */
    const int split = separate_diagonal && num_rows == num_columns;
    int64_t *rowptr = calloc((size_t)num_rows + 2, sizeof(*rowptr));
    if (!rowptr) return ENOMEM;
    /* counts at rowptr[row] with 1-based rows -> after the prefix sum
     * rowptr[i] is the start of 0-based row i */
    for (int64_t k = 0; k < num_nonzeros; k++) {
        if (split && rowidx[k] == colidx[k]) continue;
        rowptr[rowidx[k]]++;
        if (expand && rowidx[k] != colidx[k]) rowptr[colidx[k]]++;
    }
    int64_t lo = num_rows > 0 ? rowptr[1] : 0, hi = 0;
    for (int64_t i = 1; i <= (int64_t)num_rows; i++) {
        if (rowptr[i] < lo) lo = rowptr[i];
        if (rowptr[i] > hi) hi = rowptr[i];
        rowptr[i] += rowptr[i - 1];
    }
    csr->rowsizemin = (idx_t)(lo + (split ? 1 : 0));      /* csrspmv.c:1261 */
    csr->rowsizemax = (idx_t)(hi + (split ? 1 : 0));
    csr->csrsize = rowptr[num_rows];
    csr->diagsize = split ? num_rows : 0;
    size_t n = csr->csrsize > 0 ? (size_t)csr->csrsize : 1;
    csr->colidx = malloc(n * sizeof(idx_t));
    csr->a = malloc(n * sizeof(double));
    if (!csr->colidx || !csr->a) { free(rowptr); csr_free(csr); return ENOMEM; }
    if (split) {
        csr->ad = calloc(num_rows > 0 ? (size_t)num_rows : 1, sizeof(double));
        if (!csr->ad) { free(rowptr); csr_free(csr); return ENOMEM; }
    }
    /* stable placement: entry k goes to the next free slot of its row
     * (and, when expanding, its mirror image right after it:
     * csrspmv.c:1421-1425) */
    for (int64_t k = 0; k < num_nonzeros; k++) {
        const int64_t i = (int64_t)rowidx[k] - 1, j = (int64_t)colidx[k] - 1;
        if (split && i == j) { csr->ad[i] += a[k]; continue; }
        int64_t d = rowptr[i]++;
        csr->colidx[d] = (idx_t)j;
        csr->a[d] = a[k];
        if (expand && i != j) {
            d = rowptr[j]++;
            csr->colidx[d] = (idx_t)i;
            csr->a[d] = a[k];
        }
    }
    for (int64_t i = num_rows; i > 0; i--) rowptr[i] = rowptr[i - 1];
    rowptr[0] = 0;
    csr->rowptr = rowptr;
    return 0;
}

/* stable insertion sort of entries [lo, hi) of one row by column */
static void insertion_sort(idx_t *c, double *v, int64_t lo, int64_t hi)
{
    for (int64_t k = lo + 1; k < hi; k++) {
        const idx_t ck = c[k];
        const double vk = v[k];
        int64_t l = k;
        while (l > lo && c[l - 1] > ck) { c[l] = c[l - 1]; v[l] = v[l - 1]; l--; }
        c[l] = ck;
        v[l] = vk;
    }
}

/* sort n entries (c, v) by column: runs of 16 by stable insertion sort, then
 * bottom-up merging where, on equal columns, the right-hand run wins -- the
 * order the reference's rowsort produces (csrspmv.c:1279-1376).  tc/tv: scratch of n. */
static void sort_one_row(idx_t *c, double *v, int64_t n, idx_t *tc, double *tv)
{
    enum { RUN = 16 };
    if (n <= RUN) { insertion_sort(c, v, 0, n); return; }
    for (int64_t q = 0; q < n - 1; q += RUN) insertion_sort(c, v, q, q + RUN < n ? q + RUN : n);
    for (int64_t w = RUN; w < n; w *= 2) {
        memcpy(tc, c, (size_t)n * sizeof(idx_t));
        memcpy(tv, v, (size_t)n * sizeof(double));
        for (int64_t q = 0; q < n - 1; q += 2 * w) {
            const int64_t mid = q + w < n ? q + w : n, end = q + 2 * w < n ? q + 2 * w : n;
            int64_t o = q, l = q, r = mid;
            while (l < mid && r < end) {
                if (tc[l] < tc[r]) { c[o] = tc[l]; v[o++] = tv[l++]; }
                else { c[o] = tc[r]; v[o++] = tv[r++]; }
            }
            while (l < mid) { c[o] = tc[l]; v[o++] = tv[l++]; }
            while (r < end) { c[o] = tc[r]; v[o++] = tv[r++]; }
        }
    }
}

/* rows [0, num_rows): row i holds n_i = len(i) entries starting at start(i) */
static int sort_rows(idx_t *colidx, double *a, int64_t num_rows, const int64_t *rowptr /* or NULL */,
                     int64_t stride, const int64_t *rowcount /* with stride */)
{
    int err = 0;
#ifdef _OPENMP
#pragma omp parallel
#endif
    {
        idx_t *tc = NULL;
        double *tv = NULL;
        int64_t cap = 0;
#ifdef _OPENMP
#pragma omp for schedule(dynamic, 256)
#endif
        for (int64_t i = 0; i < num_rows; i++) {
            const int64_t start = rowptr ? rowptr[i] : i * stride;
            const int64_t n = rowptr ? rowptr[i + 1] - rowptr[i] : rowcount[i];
            if (n > 16 && n > cap) {
                free(tc); free(tv);
                tc = malloc((size_t)n * sizeof(idx_t));
                tv = malloc((size_t)n * sizeof(double));
                cap = n;
                if (!tc || !tv) { err = ENOMEM; cap = 0; continue; }
            }
            sort_one_row(colidx + start, a + start, n, tc, tv);
        }
        free(tc);
        free(tv);
    }
    return err;
}

int csr_sort_rows(struct csr_matrix *csr)
{
    return sort_rows(csr->colidx, csr->a, csr->num_rows, csr->rowptr, 0, NULL);
}

int ell_sort_rows(struct ell_matrix *ell)
{
    if (!ell->rowcount) return EINVAL;
    return sort_rows(ell->colidx, ell->a, ell->num_rows, NULL, ell->rowsize, ell->rowcount);
}
