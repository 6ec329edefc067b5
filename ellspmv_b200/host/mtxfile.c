/* mtxfile.c -- see mtxfile.h. */
#include "mtxfile.h"

#include <errno.h>
#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

#include <sys/mman.h>
#include <sys/stat.h>

#ifdef _OPENMP
#include <omp.h>
#endif

#ifdef HAVE_LIBZ
#include <zlib.h>
#endif

struct mtx_stream {
    FILE *f;
#ifdef HAVE_LIBZ
    gzFile gz;
#endif
    char *line;      /* line buffer, line_max + 1 bytes */
    int line_max;
};

struct mtx_stream *mtx_open(const char *path, int gzip)
{
    struct mtx_stream *s = calloc(1, sizeof(*s));
    if (!s) return NULL;
    long lm = sysconf(_SC_LINE_MAX);
    s->line_max = lm > 0 ? (int)lm : 2048;
    s->line = malloc((size_t)s->line_max + 1);
    if (!s->line) { free(s); return NULL; }
    if (gzip) {
#ifdef HAVE_LIBZ
        s->gz = gzopen(path, "r");
        if (!s->gz) { int e = errno ? errno : EIO; free(s->line); free(s); errno = e; return NULL; }
#else
        free(s->line); free(s); errno = ENOTSUP; return NULL;
#endif
    } else {
        s->f = fopen(path, "r");
        if (!s->f) { int e = errno; free(s->line); free(s); errno = e; return NULL; }
    }
    return s;
}

void mtx_close(struct mtx_stream *s)
{
    if (!s) return;
    if (s->f) fclose(s->f);
#ifdef HAVE_LIBZ
    if (s->gz) gzclose(s->gz);
#endif
    free(s->line);
    free(s);
}

/* one line into s->line; -1 at end of file, EOVERFLOW if it does not fit */
static int next_line(struct mtx_stream *s)
{
    char *got;
    int at_eof;
    if (s->f) {
        got = fgets(s->line, s->line_max + 1, s->f);
        at_eof = !got && feof(s->f);
    } else {
#ifdef HAVE_LIBZ
        got = gzgets(s->gz, s->line, s->line_max + 1);
        at_eof = !got && gzeof(s->gz);
#else
        return EINVAL;
#endif
    }
    if (!got) return at_eof ? -1 : (errno ? errno : EIO);
    size_t n = strlen(got);
    if (n > 0 && n == (size_t)s->line_max && got[n - 1] != '\n') return EOVERFLOW;
    return 0;
}

/* integer field via strtoll with range check against [lo, hi] */
static int scan_int(const char *p, char **end, long long lo, long long hi, long long *out)
{
    errno = 0;
    long long v = strtoll(p, end, 10);
    if (errno == ERANGE) return ERANGE;
    if (errno != 0 && v == 0) return errno;
    if (*end == p) return EINVAL;
    if (v < lo || v > hi) return ERANGE;
    *out = v;
    return 0;
}

static int scan_double(const char *p, char **end, double *out)
{
    errno = 0;
    double v = strtod(p, end);
    if (errno == ERANGE && (v == HUGE_VAL || v == -HUGE_VAL)) return ERANGE;
    if (*end == p) return EINVAL;
    *out = v;
    return 0;
}

/* consume `word` at *p or fail */
static int expect(const char **p, const char *word)
{
    size_t n = strlen(word);
    if (strncmp(*p, word, n) != 0) return 0;
    *p += n;
    return 1;
}

int mtx_read_header(struct mtx_stream *s, struct mtx_header *h, int64_t *lines, int64_t *bytes)
{
    int err = next_line(s);
    if (err) return err;
    const char *p = s->line;
    if (!expect(&p, "%%MatrixMarket ")) return EINVAL;
    if (expect(&p, "matrix ")) h->object = MTX_MATRIX;
    else if (expect(&p, "vector ")) h->object = MTX_VECTOR;
    else return EINVAL;
    if (expect(&p, "array ")) h->format = MTX_ARRAY;
    else if (expect(&p, "coordinate ")) h->format = MTX_COORDINATE;
    else return EINVAL;
    if (expect(&p, "real ")) h->field = MTX_REAL;
    else if (expect(&p, "integer ")) h->field = MTX_INTEGER;
    else if (expect(&p, "pattern ")) h->field = MTX_PATTERN;
    else return EINVAL;
    if (expect(&p, "general")) h->symmetry = MTX_GENERAL;
    else if (expect(&p, "symmetric")) h->symmetry = MTX_SYMMETRIC;
    else return EINVAL;
    *bytes += p - s->line;

    /* comment lines */
    do {
        (*lines)++;
        err = next_line(s);
        if (err) return err;
    } while (s->line[0] == '%');

    /* size line */
    char *q;
    long long v;
    p = s->line;
    h->num_columns = 0;
    h->num_nonzeros = 0;
    if (h->object == MTX_MATRIX && h->format == MTX_COORDINATE) {
        if ((err = scan_int(p, &q, IDX_T_MIN, IDX_T_MAX, &v))) return err;
        if (*q != ' ') return EINVAL;
        h->num_rows = (idx_t)v;
        *bytes += q - p + 1;
        p = q + 1;
        if ((err = scan_int(p, &q, IDX_T_MIN, IDX_T_MAX, &v))) return err;
        if (*q != ' ') return EINVAL;
        h->num_columns = (idx_t)v;
        *bytes += q - p + 1;
        p = q + 1;
        if ((err = scan_int(p, &q, INT64_MIN, INT64_MAX, &v))) return err;
        h->num_nonzeros = v;
        *bytes += q - p;
        if (h->num_rows < 0 || h->num_columns < 0 || h->num_nonzeros < 0) return EINVAL;
    } else if (h->object == MTX_VECTOR && h->format == MTX_ARRAY) {
        if ((err = scan_int(p, &q, IDX_T_MIN, IDX_T_MAX, &v))) return err;
        h->num_rows = (idx_t)v;
        *bytes += q - p;
        if (h->num_rows < 0) return EINVAL;
    } else {
        return EINVAL;
    }
    (*lines)++;
    return 0;
}

int mtx_read_coordinate(struct mtx_stream *s, const struct mtx_header *h,
                        idx_t *rowidx, idx_t *colidx, double *a, int64_t *lines, int64_t *bytes)
{
    const int with_value = h->field != MTX_PATTERN;
    for (int64_t k = 0; k < h->num_nonzeros; k++) {
        int err = next_line(s);
        if (err) return err;
        const char *p = s->line;
        char *q;
        long long v;
        if ((err = scan_int(p, &q, IDX_T_MIN, IDX_T_MAX, &v))) return err;
        if (*q != ' ') return EINVAL;
        /* the reference trusts the file; an index outside the matrix would
         * write outside the ELL/CSR arrays, so it is rejected here */
        if (v < 1 || v > h->num_rows) return EINVAL;
        rowidx[k] = (idx_t)v;
        *bytes += q - p + 1;
        p = q + 1;
        if ((err = scan_int(p, &q, IDX_T_MIN, IDX_T_MAX, &v))) return err;
        if (v < 1 || v > h->num_columns) return EINVAL;
        colidx[k] = (idx_t)v;
        *bytes += q - p;
        if (with_value) {
            /* real and integer values both go through strtod (the reference's
             * integer branch is unreachable, ellspmv.c:824 vs 845) */
            if (*q != ' ') return EINVAL;
            (*bytes)++;
            p = q + 1;
            if ((err = scan_double(p, &q, &a[k]))) return err;
            *bytes += q - p;
        } else {
            a[k] = 1.0;
        }
        (*lines)++;
    }
    return 0;
}

/* ---- parallel entry parser ---------------------------------------------------- */

/* unsigned decimal field starting at p (must start with a digit); returns the
 * end, or NULL when the field is not plain digits or does not fit int64 */
static const char *scan_digits(const char *p, const char *end, int64_t *out)
{
    if (p >= end || *p < '0' || *p > '9') return NULL;
    int64_t v = 0;
    int nd = 0;
    while (p < end && *p >= '0' && *p <= '9') {
        if (++nd > 18) return NULL;
        v = v * 10 + (*p - '0');
        p++;
    }
    *out = v;
    return p;
}

/* parse lines [first, first+count) that start at text; returns 0 or 1 (= use the serial reader) */
static int parse_block(const char *text, const char *end, int64_t first, int64_t count,
                       const struct mtx_header *h, int line_max,
                       idx_t *rowidx, idx_t *colidx, double *a, int64_t *bytes_out)
{
    const int with_value = h->field != MTX_PATTERN;
    const char *p = text;
    int64_t bytes = 0;
    char tmp[64];
    for (int64_t k = first; k < first + count; k++) {
        const char *nl = memchr(p, '\n', (size_t)(end - p));
        const char *eol = nl ? nl : end;
        if (eol - p + (nl ? 1 : 0) > line_max) return 1;
        int64_t v;
        const char *q = scan_digits(p, eol, &v);
        if (!q || q >= eol || *q != ' ' || v < 1 || v > h->num_rows) return 1;
        rowidx[k] = (idx_t)v;
        bytes += q - p + 1;
        const char *c0 = q + 1;
        q = scan_digits(c0, eol, &v);
        if (!q || v < 1 || v > h->num_columns) return 1;
        colidx[k] = (idx_t)v;
        bytes += q - c0;
        if (with_value) {
            if (q >= eol || *q != ' ') return 1;
            const char *v0 = q + 1;
            if (v0 >= eol || *v0 == ' ' || *v0 == '\t') return 1;   /* strtod would skip blanks: leave that to the serial reader */
            char *stop;
            double d;
            errno = 0;
            if (nl) {
                d = strtod(v0, &stop);                              /* the newline ends the number */
            } else {                                               /* last line without '\n': no terminator in the map */
                size_t n = (size_t)(eol - v0);
                if (n >= sizeof(tmp)) return 1;
                memcpy(tmp, v0, n);
                tmp[n] = '\0';
                d = strtod(tmp, &stop);
                stop = (char *)v0 + (stop - tmp);
            }
            if (stop == v0 || stop > eol || errno == ERANGE) return 1;
            a[k] = d;
            bytes += 1 + (stop - v0);
        } else {
            a[k] = 1.0;
        }
        p = nl ? nl + 1 : end;
        if (!nl && k + 1 < first + count) return 1;
    }
    *bytes_out = bytes;
    return 0;
}

int mtx_read_coordinate_parallel(struct mtx_stream *s, const struct mtx_header *h,
                                 idx_t *rowidx, idx_t *colidx, double *a, int64_t *lines, int64_t *bytes)
{
    if (!s->f || h->num_nonzeros < (1 << 16))
        return mtx_read_coordinate(s, h, rowidx, colidx, a, lines, bytes);
    const long pos = ftell(s->f);
    struct stat st;
    if (pos < 0 || fstat(fileno(s->f), &st) != 0 || !S_ISREG(st.st_mode) || st.st_size <= pos)
        return mtx_read_coordinate(s, h, rowidx, colidx, a, lines, bytes);
    char *map = mmap(NULL, (size_t)st.st_size, PROT_READ, MAP_PRIVATE, fileno(s->f), 0);
    if (map == MAP_FAILED) return mtx_read_coordinate(s, h, rowidx, colidx, a, lines, bytes);
    const char *text = map + pos, *end = map + st.st_size;
    int nthreads = 1;
#ifdef _OPENMP
    nthreads = omp_get_max_threads();
#endif
    if (nthreads > 256) nthreads = 256;
    /* cut the text into nthreads pieces at line starts, count lines per piece */
    const char *cut[257];
    int64_t nlines[256], first[257], pbytes[256];
    int bad = 0;
    cut[0] = text;
    for (int t = 1; t < nthreads; t++) {
        const char *c = text + (int64_t)(end - text) / nthreads * t;
        if (c < cut[t - 1]) c = cut[t - 1];
        const char *nl = c < end ? memchr(c, '\n', (size_t)(end - c)) : NULL;
        cut[t] = nl ? nl + 1 : end;
    }
    cut[nthreads] = end;
#ifdef _OPENMP
#pragma omp parallel for schedule(static, 1) num_threads(nthreads)
#endif
    for (int t = 0; t < nthreads; t++) {
        int64_t n = 0;
        const char *p = cut[t];
        while (p < cut[t + 1]) {
            const char *nl = memchr(p, '\n', (size_t)(cut[t + 1] - p));
            n++;
            if (!nl) break;
            p = nl + 1;
        }
        nlines[t] = n;
    }
    first[0] = 0;
    for (int t = 0; t < nthreads; t++) first[t + 1] = first[t] + nlines[t];
    if (first[nthreads] < h->num_nonzeros) bad = 1;          /* premature end of file */
    if (!bad) {
#ifdef _OPENMP
#pragma omp parallel for schedule(static, 1) num_threads(nthreads) reduction(|:bad)
#endif
        for (int t = 0; t < nthreads; t++) {
            pbytes[t] = 0;
            int64_t f0 = first[t], cnt = nlines[t];
            if (f0 >= h->num_nonzeros) continue;            /* lines past the declared count are ignored */
            if (f0 + cnt > h->num_nonzeros) cnt = h->num_nonzeros - f0;
            bad |= parse_block(cut[t], cut[t + 1], f0, cnt, h, s->line_max, rowidx, colidx, a, &pbytes[t]);
        }
    }
    munmap(map, (size_t)st.st_size);
    if (bad) {
        /* let the serial reader produce the reference's exact answer (or error) */
        if (fseek(s->f, pos, SEEK_SET) != 0) return EIO;
        return mtx_read_coordinate(s, h, rowidx, colidx, a, lines, bytes);
    }
    for (int t = 0; t < nthreads; t++) *bytes += pbytes[t];
    *lines += h->num_nonzeros;
    return 0;
}

int mtx_read_vector(struct mtx_stream *s, enum mtx_field field, int64_t n, double *x,
                    int64_t *lines, int64_t *bytes)
{
    if (field != MTX_REAL && field != MTX_INTEGER) return EINVAL;
    for (int64_t i = 0; i < n; i++) {
        int err = next_line(s);
        if (err) return err;
        char *q;
        if (field == MTX_REAL) {
            if ((err = scan_double(s->line, &q, &x[i]))) return err;
        } else {
            long long v;
            if ((err = scan_int(s->line, &q, INT_MIN, INT_MAX, &v))) return err;
            x[i] = (double)v;
        }
        *bytes += q - s->line;
        (*lines)++;
    }
    return 0;
}

void mtx_write_vector(FILE *f, idx_t n, const double *y)
{
    fprintf(f, "%%%%MatrixMarket vector array real general\n");
    fprintf(f, "%" PRIdx "\n", n);
    for (idx_t i = 0; i < n; i++) fprintf(f, "%.*g\n", DBL_DIG, y[i]);
}
