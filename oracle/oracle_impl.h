/*
 * oracle_impl.h -- index-width-generic body of the CPU oracle.
 *
 * TEST INFRASTRUCTURE ONLY (see oracle/README.md).  Included twice by
 * ellspmv_oracle.c, once with OIDX=int32_t/SUF=32 and once with
 * OIDX=int64_t/SUF=64, mirroring the reference's compile-time idx_t
 * (ellspmv.c:112-130, csrspmv.c:153-171).
 *
 * Every function is a plain-C restatement of the reference algorithm it
 * cites; none of the reference's source text is reproduced.  Arithmetic
 * in the SpMV loops is mul-then-add in left-to-right slot order, the
 * order the reference's compiled loop uses; this file
 * is built with -ffp-contract=off so no FMA can appear.
 */

#define OCAT2(a, b) a##b
#define OCAT(a, b) OCAT2(a, b)
#define ONAME(base) OCAT(base, SUF)

/*
 * ELL sizing: per-row entry counts, K = max count, ellsize = N*K.
 * Follows ell_from_coo_size, ellspmv.c:931-958 (default path,
 * separate_diagonal = false).  rowidx is 1-based as read from the
 * Matrix Market file.  rowcount has num_rows+1 slots; on return
 * rowcount[i] holds the inclusive prefix sum like the reference's
 * rowptr (callers only rely on K/ellsize/diagsize).
 */
int ONAME(oracle_ell_from_coo_size)(
    OIDX num_rows, OIDX num_columns, int64_t num_nonzeros,
    const OIDX *rowidx, int64_t *rowcount,
    int64_t *ellsize, OIDX *rowsize, OIDX *diagsize)
{
    for (int64_t i = 0; i <= (int64_t)num_rows; i++) rowcount[i] = 0;
    for (int64_t k = 0; k < num_nonzeros; k++) rowcount[rowidx[k]]++;
    int64_t widest = 0;
    for (int64_t i = 1; i <= (int64_t)num_rows; i++) {
        if (rowcount[i] > widest) widest = rowcount[i];
        rowcount[i] += rowcount[i - 1];
    }
    *rowsize = (OIDX)widest;
    /* the reference multiplies in idx_t (Q9); callers keep N*K in range */
    *ellsize = (int64_t)num_rows * widest;
    *diagsize = num_rows < num_columns ? num_rows : num_columns;
    return 0;
}

/*
 * COO -> ELL scatter in file order, then padding.
 * Follows ell_from_coo, ellspmv.c:1081-1127 (default path).  Entry k goes
 * to slot (row, fill[row]) and fill[row] advances (ellspmv.c:1102-1105);
 * unused slots get column min(row, ncols-1) and value 0.0
 * (ellspmv.c:1111-1116).  Duplicates keep separate slots (Q17).
 */
int ONAME(oracle_ell_from_coo)(
    OIDX num_rows, OIDX num_columns, int64_t num_nonzeros,
    const OIDX *rowidx, const OIDX *colidx, const double *a,
    int64_t *fill, OIDX rowsize, OIDX *ellcolidx, double *ella)
{
    const int64_t K = rowsize;
    for (int64_t i = 0; i <= (int64_t)num_rows; i++) fill[i] = 0;
    for (int64_t k = 0; k < num_nonzeros; k++) {
        int64_t r = (int64_t)rowidx[k] - 1;
        int64_t slot = r * K + fill[r];
        ellcolidx[slot] = colidx[k] - 1;
        ella[slot] = a[k];
        fill[r]++;
    }
    for (int64_t r = 0; r < (int64_t)num_rows; r++) {
        OIDX padcol = r < (int64_t)num_columns ? (OIDX)r : (OIDX)(num_columns - 1);
        for (int64_t l = fill[r]; l < K; l++) {
            ellcolidx[r * K + l] = padcol;
            ella[r * K + l] = 0.0;
        }
    }
    return 0;
}

/*
 * y <- y + A*x on row-major ELL.  Follows ellgemv, ellspmv.c:1146-1151:
 * per row, yi starts at 0, accumulates a*x left to right, then y += yi.
 */
int ONAME(oracle_ellgemv)(
    OIDX num_rows, double *y, const double *x,
    OIDX rowsize, const OIDX *colidx, const double *a)
{
    const int64_t K = rowsize;
#ifdef _OPENMP
#pragma omp parallel for schedule(static)
#endif
    for (int64_t r = 0; r < (int64_t)num_rows; r++) {
        double acc = 0.0;
        const OIDX *c = colidx + r * K;
        const double *v = a + r * K;
        for (int64_t l = 0; l < K; l++) {
            double prod = v[l] * x[c[l]];
            acc = acc + prod;
        }
        y[r] = y[r] + acc;
    }
    return 0;
}

/*
 * ITERATE mode oracle (BASELINE config 5; not a reference feature, see
 * SURVEY.md 8(e)): repeat { y = 0; y += A*x; swap(x, y) }.  Square A only.
 * On return xa holds the final vector; xb is scratch.
 */
int ONAME(oracle_ell_iterate)(
    OIDX num_rows, double *xa, double *xb, int iterations,
    OIDX rowsize, const OIDX *colidx, const double *a)
{
    double *cur = xa, *nxt = xb;
    for (int it = 0; it < iterations; it++) {
        for (int64_t r = 0; r < (int64_t)num_rows; r++) nxt[r] = 0.0;
        ONAME(oracle_ellgemv)(num_rows, nxt, cur, rowsize, colidx, a);
        double *t = cur; cur = nxt; nxt = t;
    }
    if (cur != xa) memcpy(xa, cur, (size_t)num_rows * sizeof(double));
    return 0;
}

/*
 * CSR sizing, general (non-symmetric, no diagonal split) branch.
 * Follows csr_from_coo_size, csrspmv.c:1219-1267 (the final else at 1253,
 * min/max/prefix loop 1254-1260).
 */
int ONAME(oracle_csr_from_coo_size)(
    OIDX num_rows, OIDX num_columns, int64_t num_nonzeros,
    const OIDX *rowidx, int64_t *rowptr,
    int64_t *csrsize, OIDX *rowsizemin, OIDX *rowsizemax)
{
    (void)num_columns;
    for (int64_t i = 0; i <= (int64_t)num_rows; i++) rowptr[i] = 0;
    for (int64_t k = 0; k < num_nonzeros; k++) rowptr[rowidx[k]]++;
    int64_t lo = num_rows > 0 ? rowptr[1] : 0, hi = 0;
    for (int64_t i = 1; i <= (int64_t)num_rows; i++) {
        if (rowptr[i] < lo) lo = rowptr[i];
        if (rowptr[i] > hi) hi = rowptr[i];
        rowptr[i] += rowptr[i - 1];
    }
    *rowsizemin = (OIDX)lo;
    *rowsizemax = (OIDX)hi;
    *csrsize = rowptr[num_rows];
    return 0;
}

/*
 * COO -> CSR, general branch: stable counting sort by row (file order
 * kept inside a row), 1-based -> 0-based columns.  Follows csr_from_coo,
 * csrspmv.c:1436-1465.  rowptr comes in as the exclusive prefix produced
 * by the size pass and leaves as the usual CSR row pointer.
 */
int ONAME(oracle_csr_from_coo)(
    OIDX num_rows, int64_t num_nonzeros,
    const OIDX *rowidx, const OIDX *colidx, const double *a,
    int64_t *rowptr, OIDX *csrcolidx, double *csra)
{
    for (int64_t k = 0; k < num_nonzeros; k++) {
        int64_t r = (int64_t)rowidx[k] - 1;
        int64_t dst = rowptr[r]++;
        csrcolidx[dst] = colidx[k] - 1;
        csra[dst] = a[k];
    }
    for (int64_t i = num_rows; i > 0; i--) rowptr[i] = rowptr[i - 1];
    rowptr[0] = 0;
    return 0;
}

/*
 * y <- y + A*x on CSR.  Follows csrgemv (scalar), csrspmv.c:1588-1593.
 */
int ONAME(oracle_csrgemv)(
    OIDX num_rows, double *y, const double *x,
    const int64_t *rowptr, const OIDX *colidx, const double *a)
{
#ifdef _OPENMP
#pragma omp parallel for schedule(static)
#endif
    for (int64_t r = 0; r < (int64_t)num_rows; r++) {
        double acc = 0.0;
        for (int64_t k = rowptr[r]; k < rowptr[r + 1]; k++) {
            double prod = a[k] * x[colidx[k]];
            acc = acc + prod;
        }
        y[r] = y[r] + acc;
    }
    return 0;
}

/* ------------------------------------------------------------------ */
/* Separate-diagonal variants.                 */
/* ------------------------------------------------------------------ */

/*
 * ELL conversion with the diagonal split off, as ell_from_coo_size /
 * ell_from_coo behave when their flags are passed in DECLARED order
 * (separate_diagonal = true, sort_rows = false; ellspmv.c:946-949,
 * 1098-1107).  The reference's own main() passes the two flags swapped
 * (Q1), so this is the function-level contract, not the program's.
 * K counts off-diagonal entries only; ellad[i] accumulates every (i,i)
 * entry in file order; diagsize = min(rows, cols).
 */
int ONAME(oracle_ell_from_coo_sd_size)(
    OIDX num_rows, OIDX num_columns, int64_t num_nonzeros,
    const OIDX *rowidx, const OIDX *colidx, int64_t *rowcount,
    int64_t *ellsize, OIDX *rowsize, OIDX *diagsize)
{
    for (int64_t i = 0; i <= (int64_t)num_rows; i++) rowcount[i] = 0;
    for (int64_t k = 0; k < num_nonzeros; k++)
        if (rowidx[k] != colidx[k]) rowcount[rowidx[k]]++;
    int64_t widest = 0;
    for (int64_t i = 1; i <= (int64_t)num_rows; i++) {
        if (rowcount[i] > widest) widest = rowcount[i];
        rowcount[i] += rowcount[i - 1];
    }
    *rowsize = (OIDX)widest;
    *ellsize = (int64_t)num_rows * widest;
    *diagsize = num_rows < num_columns ? num_rows : num_columns;
    return 0;
}

int ONAME(oracle_ell_from_coo_sd)(
    OIDX num_rows, OIDX num_columns, int64_t num_nonzeros,
    const OIDX *rowidx, const OIDX *colidx, const double *a,
    int64_t *fill, OIDX rowsize, OIDX *ellcolidx, double *ella, double *ellad)
{
    const int64_t K = rowsize;
    const int64_t diagsize = num_rows < num_columns ? num_rows : num_columns;
    for (int64_t i = 0; i <= (int64_t)num_rows; i++) fill[i] = 0;
    for (int64_t i = 0; i < diagsize; i++) ellad[i] = 0.0;
    for (int64_t k = 0; k < num_nonzeros; k++) {
        int64_t r = (int64_t)rowidx[k] - 1;
        if (rowidx[k] == colidx[k]) { ellad[r] += a[k]; continue; }
        int64_t slot = r * K + fill[r];
        ellcolidx[slot] = colidx[k] - 1;
        ella[slot] = a[k];
        fill[r]++;
    }
    for (int64_t r = 0; r < (int64_t)num_rows; r++) {
        OIDX padcol = r < (int64_t)num_columns ? (OIDX)r : (OIDX)(num_columns - 1);
        for (int64_t l = fill[r]; l < K; l++) { ellcolidx[r * K + l] = padcol; ella[r * K + l] = 0.0; }
    }
    return 0;
}

/*
 * y <- y + (ad.*x + A*x).  order 0 follows ellgemvsd, ellspmv.c:1173-1178:
 * yi = sum of the slots from 0, then y += ad*x + yi.  order 1 follows the
 * hand-unrolled ellgemv16sd, ellspmv.c:1201-1219, generalised to any K:
 * the sum starts from ad*x and the slots are added to it left to right.
 */
int ONAME(oracle_ellgemvsd)(
    OIDX num_rows, double *y, const double *x,
    OIDX rowsize, const OIDX *colidx, const double *a, const double *ad, int order)
{
    const int64_t K = rowsize;
#ifdef _OPENMP
#pragma omp parallel for schedule(static)
#endif
    for (int64_t r = 0; r < (int64_t)num_rows; r++) {
        const OIDX *c = colidx + r * K;
        const double *v = a + r * K;
        double dx = ad[r] * x[r];
        double acc = order ? dx : 0.0;
        for (int64_t l = 0; l < K; l++) {
            double prod = v[l] * x[c[l]];
            acc = acc + prod;
        }
        if (!order) acc = dx + acc;
        y[r] = y[r] + acc;
    }
    return 0;
}

/*
 * CSR with the diagonal split off, square general matrices: follows
 * csr_from_coo_size (csrspmv.c:1249-1252, 1261, 1265) and csr_from_coo
 * (csrspmv.c:1428-1435).  rowsizemin/max count the diagonal (+1).
 */
int ONAME(oracle_csr_from_coo_sd)(
    OIDX num_rows, int64_t num_nonzeros,
    const OIDX *rowidx, const OIDX *colidx, const double *a,
    int64_t *rowptr, int64_t *csrsize, OIDX *rowsizemin, OIDX *rowsizemax,
    OIDX *csrcolidx, double *csra, double *csrad)
{
    for (int64_t i = 0; i <= (int64_t)num_rows; i++) rowptr[i] = 0;
    for (int64_t k = 0; k < num_nonzeros; k++)
        if (rowidx[k] != colidx[k]) rowptr[rowidx[k]]++;
    int64_t lo = num_rows > 0 ? rowptr[1] : 0, hi = 0;
    for (int64_t i = 1; i <= (int64_t)num_rows; i++) {
        if (rowptr[i] < lo) lo = rowptr[i];
        if (rowptr[i] > hi) hi = rowptr[i];
        rowptr[i] += rowptr[i - 1];
    }
    *rowsizemin = (OIDX)(lo + 1);
    *rowsizemax = (OIDX)(hi + 1);
    *csrsize = rowptr[num_rows];
    if (!csrcolidx) return 0;                       /* size pass only */
    for (int64_t i = 0; i < (int64_t)num_rows; i++) csrad[i] = 0.0;
    for (int64_t k = 0; k < num_nonzeros; k++) {
        int64_t i = (int64_t)rowidx[k] - 1, j = (int64_t)colidx[k] - 1;
        if (i == j) { csrad[i] += a[k]; continue; }
        int64_t dst = rowptr[i]++;
        csrcolidx[dst] = (OIDX)j;
        csra[dst] = a[k];
    }
    for (int64_t i = num_rows; i > 0; i--) rowptr[i] = rowptr[i - 1];
    rowptr[0] = 0;
    return 0;
}

/* y <- y + (ad.*x + A*x) on CSR; follows csrgemvsd, csrspmv.c:1622-1627 */
int ONAME(oracle_csrgemvsd)(
    OIDX num_rows, double *y, const double *x,
    const int64_t *rowptr, const OIDX *colidx, const double *a, const double *ad)
{
#ifdef _OPENMP
#pragma omp parallel for schedule(static)
#endif
    for (int64_t r = 0; r < (int64_t)num_rows; r++) {
        double acc = 0.0;
        for (int64_t k = rowptr[r]; k < rowptr[r + 1]; k++) {
            double prod = a[k] * x[colidx[k]];
            acc = acc + prod;
        }
        double dx = ad[r] * x[r];
        acc = dx + acc;
        y[r] = y[r] + acc;
    }
    return 0;
}

/*
 * Sort the entries of every CSR row by column (--sort-rows).  Follows
 * rowsort, csrspmv.c:1269-1388: rows of at most 16 entries get a stable
 * insertion sort; longer rows are insertion-sorted in blocks of 16 and the
 * blocks are merged bottom-up, doubling the width, where on EQUAL columns
 * the entry of the right-hand run goes first (csrspmv.c:1356-1360) -- so
 * duplicates of a column keep file order only inside a 16-block.
 */
int ONAME(oracle_rowsort)(
    OIDX num_rows, const int64_t *rowptr, OIDX *colidx, double *a)
{
    const int64_t block = 16;
    int64_t longest = 0;
    for (int64_t i = 0; i < (int64_t)num_rows; i++)
        if (rowptr[i + 1] - rowptr[i] > longest) longest = rowptr[i + 1] - rowptr[i];
    OIDX *tc = (OIDX *)malloc((size_t)(longest > 0 ? longest : 1) * sizeof(OIDX));
    double *ta = (double *)malloc((size_t)(longest > 0 ? longest : 1) * sizeof(double));
    if (!tc || !ta) { free(tc); free(ta); return 12; }
    for (int64_t i = 0; i < (int64_t)num_rows; i++) {
        OIDX *c = colidx + rowptr[i];
        double *v = a + rowptr[i];
        const int64_t n = rowptr[i + 1] - rowptr[i];
        const int64_t width = n <= block ? (n > 0 ? n : 1) : block;
        for (int64_t q = 0; q < n - 1; q += width) {
            const int64_t e = q + width < n ? q + width : n;
            for (int64_t k = q + 1; k < e; k++) {
                OIDX cj = c[k]; double vj = v[k];
                int64_t l = k - 1;
                while (l >= q && c[l] > cj) { c[l + 1] = c[l]; v[l + 1] = v[l]; l--; }
                c[l + 1] = cj; v[l + 1] = vj;
            }
        }
        if (n <= block) continue;
        for (int64_t p = block; p < n; p *= 2) {
            memcpy(tc, c, (size_t)n * sizeof(OIDX));
            memcpy(ta, v, (size_t)n * sizeof(double));
            for (int64_t q = 0; q < n - 1; q += 2 * p) {
                const int64_t mid = q + p < n ? q + p : n, end = q + 2 * p < n ? q + 2 * p : n;
                int64_t out = q, l = q, r = mid;
                while (l < mid && r < end) {
                    if (tc[l] < tc[r]) { c[out] = tc[l]; v[out] = ta[l]; l++; }
                    else { c[out] = tc[r]; v[out] = ta[r]; r++; }
                    out++;
                }
                while (l < mid) { c[out] = tc[l]; v[out] = ta[l]; l++; out++; }
                while (r < end) { c[out] = tc[r]; v[out] = ta[r]; r++; out++; }
            }
        }
    }
    free(tc); free(ta);
    return 0;
}

/* ------------------------------------------------------------------ */
/* Synthetic matrices of BASELINE.json's shapes.      */
/* Each generator can emit the ELL arrays for a row range directly     */
/* (row-major, reference padding rule) and the 1-based COO stream in   */
/* canonical order, so tests can push the COO stream through           */
/* ell_from_coo and check both routes agree.                           */
/* ------------------------------------------------------------------ */

/*
 * 2D 5-point Laplacian on an nx-by-ny grid, row r = i*ny + j.
 * Slot order: (i-1,j), (i,j-1), (i,j), (i,j+1), (i+1,j), only those in
 * range; centre = cval, neighbours = oval.  K = 5.
 */
int64_t ONAME(oracle_gen_laplace2d_ell)(
    int64_t nx, int64_t ny, double cval, double oval,
    int64_t row_begin, int64_t row_end, OIDX *ellcolidx, double *ella)
{
    const int64_t K = 5, ncols = nx * ny;
    int64_t real = 0;
    for (int64_t r = row_begin; r < row_end; r++) {
        int64_t i = r / ny, j = r % ny, n = 0;
        OIDX *c = ellcolidx + (r - row_begin) * K;
        double *v = ella + (r - row_begin) * K;
        if (i > 0)      { c[n] = (OIDX)(r - ny); v[n] = oval; n++; }
        if (j > 0)      { c[n] = (OIDX)(r - 1);  v[n] = oval; n++; }
        c[n] = (OIDX)r; v[n] = cval; n++;
        if (j + 1 < ny) { c[n] = (OIDX)(r + 1);  v[n] = oval; n++; }
        if (i + 1 < nx) { c[n] = (OIDX)(r + ny); v[n] = oval; n++; }
        real += n;
        OIDX padcol = r < ncols ? (OIDX)r : (OIDX)(ncols - 1);
        for (; n < K; n++) { c[n] = padcol; v[n] = 0.0; }
    }
    return real;
}

int64_t ONAME(oracle_gen_laplace2d_coo)(
    int64_t nx, int64_t ny, double cval, double oval,
    OIDX *rowidx, OIDX *colidx, double *a)
{
    int64_t k = 0;
    for (int64_t r = 0; r < nx * ny; r++) {
        int64_t i = r / ny, j = r % ny;
#define OEMIT(cc, vv) do { if (rowidx) { rowidx[k] = (OIDX)(r + 1); colidx[k] = (OIDX)((cc) + 1); a[k] = (vv); } k++; } while (0)
        if (i > 0)      OEMIT(r - ny, oval);
        if (j > 0)      OEMIT(r - 1, oval);
        OEMIT(r, cval);
        if (j + 1 < ny) OEMIT(r + 1, oval);
        if (i + 1 < nx) OEMIT(r + ny, oval);
    }
    return k;
}

/*
 * 3D 27-point stencil on an nx-by-ny-by-nz grid, row r = (i*ny + j)*nz + k.
 * Slot order: (di,dj,dk) in {-1,0,1}^3 lexicographic, in-range only;
 * centre = cval, others = oval.  K = 27.
 */
int64_t ONAME(oracle_gen_stencil27_ell)(
    int64_t nx, int64_t ny, int64_t nz, double cval, double oval,
    int64_t row_begin, int64_t row_end, OIDX *ellcolidx, double *ella)
{
    const int64_t K = 27, ncols = nx * ny * nz;
    int64_t real = 0;
    for (int64_t r = row_begin; r < row_end; r++) {
        int64_t k = r % nz, j = (r / nz) % ny, i = r / (nz * ny), n = 0;
        OIDX *c = ellcolidx + (r - row_begin) * K;
        double *v = ella + (r - row_begin) * K;
        for (int di = -1; di <= 1; di++)
            for (int dj = -1; dj <= 1; dj++)
                for (int dk = -1; dk <= 1; dk++) {
                    int64_t ii = i + di, jj = j + dj, kk = k + dk;
                    if (ii < 0 || ii >= nx || jj < 0 || jj >= ny || kk < 0 || kk >= nz) continue;
                    c[n] = (OIDX)((ii * ny + jj) * nz + kk);
                    v[n] = (di == 0 && dj == 0 && dk == 0) ? cval : oval;
                    n++;
                }
        real += n;
        OIDX padcol = r < ncols ? (OIDX)r : (OIDX)(ncols - 1);
        for (; n < K; n++) { c[n] = padcol; v[n] = 0.0; }
    }
    return real;
}

int64_t ONAME(oracle_gen_stencil27_coo)(
    int64_t nx, int64_t ny, int64_t nz, double cval, double oval,
    OIDX *rowidx, OIDX *colidx, double *a)
{
    int64_t cnt = 0;
    for (int64_t r = 0; r < nx * ny * nz; r++) {
        int64_t k = r % nz, j = (r / nz) % ny, i = r / (nz * ny);
        for (int di = -1; di <= 1; di++)
            for (int dj = -1; dj <= 1; dj++)
                for (int dk = -1; dk <= 1; dk++) {
                    int64_t ii = i + di, jj = j + dj, kk = k + dk;
                    if (ii < 0 || ii >= nx || jj < 0 || jj >= ny || kk < 0 || kk >= nz) continue;
                    if (rowidx) {
                        rowidx[cnt] = (OIDX)(r + 1);
                        colidx[cnt] = (OIDX)((ii * ny + jj) * nz + kk + 1);
                        a[cnt] = (di == 0 && dj == 0 && dk == 0) ? cval : oval;
                    }
                    cnt++;
                }
    }
    return cnt;
}

/*
 * Random N-by-ncols matrix with exactly K entries per row, counter-based:
 * u = splitmix64(seed ^ (r*K + l)); col = mulhi64(u, ncols);
 * val = 2*((splitmix64(u) >> 11) * 2^-53) - 1.  Duplicates allowed (Q17).
 */
int64_t ONAME(oracle_gen_random_ell)(
    int64_t num_rows, int64_t num_columns, int64_t K, uint64_t seed,
    int64_t row_begin, int64_t row_end, OIDX *ellcolidx, double *ella)
{
    (void)num_rows;
    for (int64_t r = row_begin; r < row_end; r++) {
        for (int64_t l = 0; l < K; l++) {
            uint64_t u = oracle_splitmix64(seed ^ (uint64_t)(r * K + l));
            uint64_t col = (uint64_t)(((unsigned __int128)u * (unsigned __int128)(uint64_t)num_columns) >> 64);
            double val = 2.0 * ((double)(oracle_splitmix64(u) >> 11) * 0x1.0p-53) - 1.0;
            ellcolidx[(r - row_begin) * K + l] = (OIDX)col;
            ella[(r - row_begin) * K + l] = val;
        }
    }
    return (row_end - row_begin) * K;
}

int64_t ONAME(oracle_gen_random_coo)(
    int64_t num_rows, int64_t num_columns, int64_t K, uint64_t seed,
    OIDX *rowidx, OIDX *colidx, double *a)
{
    int64_t cnt = 0;
    for (int64_t r = 0; r < num_rows; r++) {
        for (int64_t l = 0; l < K; l++) {
            uint64_t u = oracle_splitmix64(seed ^ (uint64_t)(r * K + l));
            uint64_t col = (uint64_t)(((unsigned __int128)u * (unsigned __int128)(uint64_t)num_columns) >> 64);
            double val = 2.0 * ((double)(oracle_splitmix64(u) >> 11) * 0x1.0p-53) - 1.0;
            if (rowidx) { rowidx[cnt] = (OIDX)(r + 1); colidx[cnt] = (OIDX)(col + 1); a[cnt] = val; }
            cnt++;
        }
    }
    return cnt;
}

#undef OEMIT
#undef ONAME
#undef OCAT
#undef OCAT2
