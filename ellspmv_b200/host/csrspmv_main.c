/*
 * csrspmv_main.c -- host program `csrspmv`: y := A*x + y with A in CSR
 * format on a B200; the comparison path of the ELL program.
 *
 * Drop-in for the reference's `csrspmv` (csrspmv.c:1766-2959) on its default
 * path: same arguments, reader, COO->CSR conversion (stable by row, file
 * order inside a row, symmetric expansion for square symmetric input:
 * csrspmv.c:1219-1267, 1390-1475), timing lines and output.  The kernel call
 * sites (csrspmv.c:2766-2767, 2857-2858) become
 *     csrspmv_cuda_upload() / csrspmv_cuda_spmv() / csrspmv_cuda_free().
 *
 * The reference's OpenMP thread-partitioning options (--partition-rows,
 * --partition-nonzeros, --precompute-partition, --rows-per-thread,
 * --columns-per-thread) describe how CPU threads split the rows; they are
 * accepted and ignored (a note is printed with -v).  --separate-diagonal is
 * supported for square matrices (csrgemvsd, csrspmv.c:1598-1629; on a
 * non-square matrix the reference skips the split but still reads the absent
 * diagonal array, so that case is refused).  --sort-rows sorts every row by
 * column on the host with the reference's rowsort order (csrspmv.c:1269-1388).
 */
#include <errno.h>
#include <locale.h>
#include <stdbool.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "../../include/ellspmv_cuda.h"
#include "convert.h"
#include "hostutil.h"
#include "idx.h"
#include "mtxfile.h"

static const char *version = "1.10-b200";

struct options {
    const char *Apath, *xpath, *ypath;
    int gzip;
    bool separate_diagonal, sort_rows, ignored_partition, device_convert;
    int repeat, warmup, verbose, quiet, gpus;
    unsigned flags;
    const char *synthetic;
};

static void usage(FILE *f) { fprintf(f, "Usage: %s [OPTION..] A [x] [y]\n", prog); }

static void help(FILE *f)
{
    usage(f);
    fprintf(f, "\n");
    fprintf(f, " Multiply a matrix by a vector on a CUDA device.\n");
    fprintf(f, "\n");
    fprintf(f, " The operation performed is ‘y := A*x + y’, where\n");
    fprintf(f, " ‘A’ is a matrix, and ‘x’ and ‘y’ are vectors.\n");
    fprintf(f, "\n");
    fprintf(f, " Positional arguments are:\n");
    fprintf(f, "  A        path to Matrix Market file for the matrix A\n");
    fprintf(f, "  x        optional path to Matrix Market file for the vector x\n");
    fprintf(f, "  y        optional path for to Matrix Market file for the vector y\n");
    fprintf(f, "\n");
    fprintf(f, " Other options are:\n");
#ifdef HAVE_LIBZ
    fprintf(f, "  -z, --gzip, --gunzip, --ungzip    filter files through gzip\n");
#endif
    fprintf(f, "  --separate-diagonal       store diagonal nonzeros separately\n");
    fprintf(f, "  --sort-rows               sort nonzeros by column within each row\n");
    fprintf(f, "  --partition-rows, --partition-nonzeros, --precompute-partition,\n");
    fprintf(f, "  --rows-per-thread=N.., --columns-per-thread=N..   accepted, ignored (CPU thread partitioning)\n");
    fprintf(f, "  --repeat=N                repeat matrix-vector multiplication N times\n");
    fprintf(f, "  --warmup=N                perform N additional warmup iterations\n");
    fprintf(f, "  -q, --quiet               do not print Matrix Market output\n");
    fprintf(f, "  -v, --verbose             be more verbose\n");
    fprintf(f, "\n");
    fprintf(f, " Options for the CUDA path are:\n");
    fprintf(f, "  --kernel=auto|stream|scalar|sell|warp\n");
    fprintf(f, "                            auto (default): a sliced-ELL view for balanced rows, SELL-128-sigma\n");
    fprintf(f, "                            for skewed ones; stream/scalar: the native CSR kernels (all bit-exact);\n");
    fprintf(f, "                            warp: sub-warp-per-row (tolerance mode)\n");
    fprintf(f, "  --fma                     allow fused multiply-add (tolerance mode)\n");
    fprintf(f, "  --device-convert          convert COO to CSR on the device (general matrices)\n");
    fprintf(f, "  --synthetic=SPEC          build A on the device instead of reading a file:\n");
    fprintf(f, "                            laplace2d:NX,NY | stencil27:NX,NY,NZ | random:ROWS,COLS,K[,SEED]\n");
    fprintf(f, "  --gpus=N                  split the rows in N nonzero-balanced blocks over devices 0..N-1 [1]\n");
    fprintf(f, "\n");
    fprintf(f, "  -h, --help                display this help and exit\n");
    fprintf(f, "  --version                 display version information and exit\n");
}

static int parse_options(int argc, char **argv, struct options *o, int *bad)
{
    memset(o, 0, sizeof(*o));
    o->repeat = 1;
    o->gpus = 1;
    int npos = 0;
    bool only_positional = false;
    for (int i = 1; i < argc; i++) {
        *bad = i;
        const char *a = argv[i], *v;
        if (!only_positional) {
            if (!strcmp(a, "--separate-diagonal")) { o->separate_diagonal = true; continue; }
            if (!strcmp(a, "--sort-rows")) { o->sort_rows = true; continue; }
            if (!strcmp(a, "--partition-rows") || !strcmp(a, "--partition-nonzeros") ||
                !strcmp(a, "--precompute-partition")) { o->ignored_partition = true; continue; }
            if (!strncmp(a, "--rows-per-thread", 17) && (a[17] == '=' || a[17] == '\0')) {
                if (!optval(argc, argv, &i, "--rows-per-thread")) return EINVAL;
                o->ignored_partition = true; continue;
            }
            if (!strncmp(a, "--columns-per-thread", 20) && (a[20] == '=' || a[20] == '\0')) {
                if (!optval(argc, argv, &i, "--columns-per-thread")) return EINVAL;
                o->ignored_partition = true; continue;
            }
            if (!strncmp(a, "--repeat", 8) && (a[8] == '=' || a[8] == '\0')) {
                if (!(v = optval(argc, argv, &i, "--repeat"))) return EINVAL;
                int err = to_int(v, &o->repeat);
                if (err) return err;
                continue;
            }
            if (!strncmp(a, "--warmup", 8) && (a[8] == '=' || a[8] == '\0')) {
                if (!(v = optval(argc, argv, &i, "--warmup"))) return EINVAL;
                if (to_int(v, &o->warmup)) return EINVAL;
                continue;
            }
#ifdef HAVE_LIBZ
            if (!strcmp(a, "-z") || !strcmp(a, "--gzip") || !strcmp(a, "--gunzip") || !strcmp(a, "--ungzip")) {
                o->gzip = 1; continue;
            }
#endif
            if (!strcmp(a, "-q") || !strcmp(a, "--quiet")) { o->quiet = 1; continue; }
            if (!strcmp(a, "-v") || !strcmp(a, "--verbose")) { o->verbose++; continue; }
            if (!strncmp(a, "--kernel", 8) && (a[8] == '=' || a[8] == '\0')) {
                if (!(v = optval(argc, argv, &i, "--kernel"))) return EINVAL;
                o->flags &= ~(unsigned)ELLSPMV_CUDA_KERNEL_MASK;
                if (!strcmp(v, "thread") || !strcmp(v, "stream")) o->flags |= ELLSPMV_CUDA_KERNEL_THREAD;
                else if (!strcmp(v, "warp")) o->flags |= ELLSPMV_CUDA_KERNEL_WARP;
                else if (!strcmp(v, "scalar")) o->flags |= CSRSPMV_CUDA_KERNEL_SCALAR;
                else if (!strcmp(v, "sell")) o->flags |= CSRSPMV_CUDA_KERNEL_SELL;
                else if (strcmp(v, "auto")) return EINVAL;
                continue;
            }
            if (!strcmp(a, "--fma")) { o->flags |= ELLSPMV_CUDA_FMA; continue; }
            if (!strcmp(a, "--device-convert")) { o->device_convert = true; continue; }
            if (!strncmp(a, "--synthetic", 11) && (a[11] == '=' || a[11] == '\0')) {
                if (!(o->synthetic = optval(argc, argv, &i, "--synthetic"))) return EINVAL;
                continue;
            }
            if (!strncmp(a, "--gpus", 6) && (a[6] == '=' || a[6] == '\0')) {
                if (!(v = optval(argc, argv, &i, "--gpus")) || to_int(v, &o->gpus) || o->gpus < 1) return EINVAL;
                continue;
            }
            if (!strcmp(a, "-h") || !strcmp(a, "--help")) { help(stdout); exit(EXIT_SUCCESS); }
            if (!strcmp(a, "--version")) {
                printf("%s %s\nrow/column offsets: %d-bit\n", prog, version, IDX_BITS);
                exit(EXIT_SUCCESS);
            }
            if (!strcmp(a, "--")) { only_positional = true; continue; }
        }
        if (npos == 0) o->Apath = a;
        else if (npos == 1) o->xpath = a;
        else if (npos == 2) o->ypath = a;
        else return EINVAL;
        npos++;
    }
    if (o->synthetic && npos > 0) {
        /* with --synthetic the positionals are [x] [y] */
        o->ypath = o->xpath; o->xpath = o->Apath; o->Apath = NULL;
        if (npos > 2) return EINVAL;
    } else if (npos < 1 && !o->synthetic) {
        usage(stdout);
        exit(EXIT_FAILURE);
    }
    return 0;
}

static int parse_synthetic(const char *spec, int *kind, int64_t dims[3], double vals[2], uint64_t *seed)
{
    char name[32];
    long long d[4] = {0, 0, 0, 42};
    const char *colon = strchr(spec, ':');
    if (!colon || (size_t)(colon - spec) >= sizeof(name)) return EINVAL;
    memcpy(name, spec, (size_t)(colon - spec));
    name[colon - spec] = '\0';
    int n = sscanf(colon + 1, "%lld,%lld,%lld,%lld", &d[0], &d[1], &d[2], &d[3]);
    *seed = 42;
    if (!strcmp(name, "laplace2d") && n == 2) { *kind = ELLSPMV_CUDA_GEN_LAPLACE2D; vals[0] = 4.0; vals[1] = -1.0; }
    else if (!strcmp(name, "stencil27") && n == 3) { *kind = ELLSPMV_CUDA_GEN_STENCIL27; vals[0] = 26.0; vals[1] = -1.0; }
    else if (!strcmp(name, "random") && n >= 3) { *kind = ELLSPMV_CUDA_GEN_RANDOM; if (n == 4) *seed = (uint64_t)d[3]; }
    else return EINVAL;
    dims[0] = d[0]; dims[1] = d[1]; dims[2] = d[2];
    return 0;
}

int main(int argc, char *argv[])
{
    struct timespec t0, t1;
    setlocale(LC_ALL, "");
    const char *slash = strrchr(argv[0], '/');
    prog = slash ? slash + 1 : argv[0];

    struct options o;
    int bad = 0;
    int err = parse_options(argc, argv, &o, &bad);
    if (err) {
        fprintf(stderr, "%s: %s %s\n", prog, strerror(err), bad < argc ? argv[bad] : "");
        return EXIT_FAILURE;
    }
    if (o.ignored_partition && o.verbose > 0)
        fprintf(stderr, "%s: note: CPU thread-partitioning options are ignored on the CUDA path\n", prog);

    idx_t num_rows = 0, num_columns = 0;
    int64_t num_nonzeros = 0;
    csrspmv_cuda_matrix *A = NULL;
    int64_t csrsize = 0, diagsize = 0;

    if (o.synthetic) {
        /* build the matrix on the device, in the form csr_from_coo gives it (shapes too large for a text file) */
        int kind = 0;
        int64_t dims[3];
        double vals[2] = {0, 0};
        uint64_t seed;
        if (o.separate_diagonal || o.sort_rows || o.gpus > 1) {
            fprintf(stderr, "%s: --synthetic takes neither --separate-diagonal, --sort-rows nor --gpus\n", prog);
            return EXIT_FAILURE;
        }
        if (parse_synthetic(o.synthetic, &kind, dims, vals, &seed)) {
            fprintf(stderr, "%s: %s --synthetic=%s\n", prog, strerror(EINVAL), o.synthetic);
            return EXIT_FAILURE;
        }
        if (o.verbose > 0) { fprintf(stderr, "cuda_generate: "); clock_gettime(CLOCK_MONOTONIC, &t0); }
        err = csrspmv_cuda_generate(&A, kind, dims, vals, seed, IDX_BITS, -1, o.flags);
        if (err) {
            if (o.verbose > 0) fprintf(stderr, "\n");
            fprintf(stderr, "%s: %s (%s)\n", prog, strerror(err), ellspmv_cuda_last_error());
            return EXIT_FAILURE;
        }
        csrspmv_cuda_info info;
        csrspmv_cuda_get_info(A, &info);
        num_rows = (idx_t)info.num_rows;
        num_columns = (idx_t)info.num_columns;
        num_nonzeros = csrsize = info.csrsize;
        if (o.verbose > 0) {
            clock_gettime(CLOCK_MONOTONIC, &t1);
            fprintf(stderr, "%'.6f seconds, %'" PRIdx " rows, %'" PRIdx " columns, %'" PRId64 " nonzeros"
                            ", %'" PRId64 " to %'" PRId64 " nonzeros per row\n",
                    seconds_between(t0, t1), num_rows, num_columns, csrsize, info.min_row_len, info.max_row_len);
        }
    } else {
    /* 2. read the matrix (csrspmv.c:1843-1909) */
    if (o.verbose > 0) { fprintf(stderr, "mtxfile_read: "); clock_gettime(CLOCK_MONOTONIC, &t0); }
    struct mtx_stream *s = mtx_open(o.Apath, o.gzip);
    if (!s) { fprintf(stderr, "%s: %s: %s\n", prog, o.Apath, strerror(errno)); return EXIT_FAILURE; }
    struct mtx_header h;
    int64_t lines = 0, bytes = 0;
    err = mtx_read_header(s, &h, &lines, &bytes);
    if (!err && !(h.object == MTX_MATRIX && h.format == MTX_COORDINATE)) err = EINVAL;
    if (err) {
        if (o.verbose > 0) fprintf(stderr, "\n");
        fprintf(stderr, "%s: %s:%" PRId64 ": %s\n", prog, o.Apath, lines + 1, strerror(err));
        mtx_close(s);
        return EXIT_FAILURE;
    }
    num_rows = h.num_rows; num_columns = h.num_columns;
    num_nonzeros = h.num_nonzeros;
    size_t nz = num_nonzeros > 0 ? (size_t)num_nonzeros : 1;
    idx_t *rowidx = malloc(nz * sizeof(idx_t));
    idx_t *colidx = malloc(nz * sizeof(idx_t));
    double *a = malloc(nz * sizeof(double));
    if (!rowidx || !colidx || !a) { fprintf(stderr, "%s: %s\n", prog, strerror(ENOMEM)); return EXIT_FAILURE; }
    err = (getenv("ELLSPMV_SERIAL_READER") ? mtx_read_coordinate : mtx_read_coordinate_parallel)(s, &h, rowidx, colidx, a, &lines, &bytes);
    if (err) {
        if (o.verbose > 0) fprintf(stderr, "\n");
        fprintf(stderr, "%s: %s:%" PRId64 ": %s\n", prog, o.Apath, lines + 1, strerror(err));
        mtx_close(s);
        return EXIT_FAILURE;
    }
    if (o.verbose > 0) {
        clock_gettime(CLOCK_MONOTONIC, &t1);
        fprintf(stderr, "%'.6f seconds (%'.1f MB/s)\n", seconds_between(t0, t1),
                1.0e-6 * (double)bytes / seconds_between(t0, t1));
    }
    mtx_close(s);

    /* 3. convert to CSR (csrspmv.c:1911-2287) */
    if (o.verbose > 0) { fprintf(stderr, "csr_from_coo: "); clock_gettime(CLOCK_MONOTONIC, &t0); }
    if (o.separate_diagonal && num_rows != num_columns) {
        if (o.verbose > 0) fprintf(stderr, "\n");
        fprintf(stderr, "%s: --separate-diagonal needs a square matrix\n", prog);
        return EXIT_FAILURE;
    }
    if (o.device_convert && !o.separate_diagonal && !o.sort_rows && h.symmetry == MTX_GENERAL && o.gpus == 1) {
        /* stable sort by row on the device; rowsizemin/max are only printed, count them here */
        int64_t *cnt = calloc((size_t)num_rows + 1, sizeof(*cnt));
        if (!cnt) { fprintf(stderr, "%s: %s\n", prog, strerror(ENOMEM)); return EXIT_FAILURE; }
        for (int64_t k = 0; k < num_nonzeros; k++) cnt[rowidx[k] - 1]++;
        int64_t lo = num_rows > 0 ? cnt[0] : 0, hi = 0;
        for (idx_t i = 0; i < num_rows; i++) { if (cnt[i] < lo) lo = cnt[i]; if (cnt[i] > hi) hi = cnt[i]; }
        free(cnt);
        err = csrspmv_cuda_upload_coo(&A, IDX_BITS, num_rows, num_columns, num_nonzeros, rowidx, colidx, a, o.flags);
        free(a); free(colidx); free(rowidx);
        if (err) {
            if (o.verbose > 0) fprintf(stderr, "\n");
            fprintf(stderr, "%s: %s (%s)\n", prog, strerror(err), ellspmv_cuda_last_error());
            return EXIT_FAILURE;
        }
        csrsize = num_nonzeros;
        if (o.verbose > 0) {
            clock_gettime(CLOCK_MONOTONIC, &t1);
            fprintf(stderr, "%'.6f seconds, %'" PRIdx " rows, %'" PRIdx " columns, %'" PRId64 " nonzeros"
                            ", %'" PRId64 " to %'" PRId64 " nonzeros per row\n",
                    seconds_between(t0, t1), num_rows, num_columns, csrsize, lo, hi);
        }
    } else {
        struct csr_matrix csr;
        err = csr_from_coo(&csr, h.symmetry == MTX_SYMMETRIC, num_rows, num_columns, num_nonzeros, rowidx, colidx, a,
                           o.separate_diagonal);
        free(a); free(colidx); free(rowidx);
        if (!err && o.sort_rows) err = csr_sort_rows(&csr);
        if (err) {
            if (o.verbose > 0) fprintf(stderr, "\n");
            fprintf(stderr, "%s: %s\n", prog, strerror(err));
            return EXIT_FAILURE;
        }
        csrsize = csr.csrsize;
        diagsize = csr.diagsize;
        if (o.verbose > 0) {
            clock_gettime(CLOCK_MONOTONIC, &t1);
            fprintf(stderr, "%'.6f seconds, %'" PRIdx " rows, %'" PRIdx " columns, %'" PRId64 " nonzeros"
                            ", %'" PRIdx " to %'" PRIdx " nonzeros per row\n",
                    seconds_between(t0, t1), num_rows, num_columns, csrsize + diagsize, csr.rowsizemin, csr.rowsizemax);
        }
        if (o.verbose > 0) { fprintf(stderr, "cuda_upload: "); clock_gettime(CLOCK_MONOTONIC, &t0); }
        err = csrspmv_cuda_upload(&A, IDX_BITS, num_rows, num_columns, csr.rowptr, csr.colidx, csr.a, o.gpus, o.flags);
        if (!err && csr.ad) err = csrspmv_cuda_set_diagonal(A, csr.ad);
        csr_free(&csr);
        if (err) {
            if (o.verbose > 0) fprintf(stderr, "\n");
            fprintf(stderr, "%s: %s (%s)\n", prog, strerror(err), ellspmv_cuda_last_error());
            return EXIT_FAILURE;
        }
        if (o.verbose > 0) {
            clock_gettime(CLOCK_MONOTONIC, &t1);
            fprintf(stderr, "%'.6f seconds, %'" PRId64 " bytes on the device\n", seconds_between(t0, t1),
                    csrspmv_cuda_device_bytes(A));
        }
    }

    }   /* !synthetic */

    /* 4. vectors (csrspmv.c:2340-2631) */
    double *x = NULL, *y = NULL;
    if (ellspmv_cuda_malloc_host((void **)&x, (int64_t)(num_columns > 0 ? num_columns : 1) * 8) ||
        ellspmv_cuda_malloc_host((void **)&y, (int64_t)(num_rows > 0 ? num_rows : 1) * 8)) {
        fprintf(stderr, "%s: %s (%s)\n", prog, strerror(ENOMEM), ellspmv_cuda_last_error());
        csrspmv_cuda_free(A);
        return EXIT_FAILURE;
    }
    for (idx_t j = 0; j < num_columns; j++) x[j] = 1.0;
    for (idx_t i = 0; i < num_rows; i++) y[i] = 0.0;
    if (o.xpath && read_vector_file(o.xpath, o.gzip, num_columns, x, o.verbose)) goto fail;
    if (o.ypath && read_vector_file(o.ypath, o.gzip, num_rows, y, o.verbose)) goto fail;

    /* 5. warm-up + timed multiplications (csrspmv.c:2742-2901) */
    {
        const int total = (o.warmup > 0 ? o.warmup : 0) + (o.repeat > 0 ? o.repeat : 0);
        double *secs = calloc((size_t)(total > 0 ? total : 1), sizeof(double));
        if (!secs) { fprintf(stderr, "%s: %s\n", prog, strerror(ENOMEM)); goto fail; }
        err = csrspmv_cuda_spmv(A, y, x, total, ELLSPMV_CUDA_ACCUMULATE, secs);
        if (err) {
            fprintf(stderr, "%s: %s (%s)\n", prog, strerror(err), ellspmv_cuda_last_error());
            free(secs);
            goto fail;
        }
        if (o.verbose > 0) {
            /* the reference's model (csrspmv.c:2882-2887) */
            const int64_t num_flops = 2 * (csrsize + diagsize);
            const int64_t min_bytes = (int64_t)num_rows * 8 + (int64_t)num_columns * 8 +
                                      ((int64_t)num_rows + 1) * 8 + csrsize * (int64_t)sizeof(idx_t) + csrsize * 8 +
                                      diagsize * 8;
            const int64_t max_bytes = (int64_t)num_rows * 8 + csrsize * 8 + (int64_t)num_rows * 8 +
                                      csrsize * (int64_t)sizeof(idx_t) + csrsize * 8 + diagsize * 8 + diagsize * 8;
            for (int r = 0; r < total; r++) {
                const double t = secs[r];
                const char *label = o.separate_diagonal ? "gemvsd" : "gemv";
                fprintf(stderr, r < o.warmup ? "%s (warmup): " : "%s: ", label);
                fprintf(stderr, "%'.6f seconds (%'.3f Gnz/s, %'.3f Gflop/s, %'.1f to %'.1f GB/s)\n", t,
                        (double)num_nonzeros * 1e-9 / t, (double)num_flops * 1e-9 / t,
                        (double)min_bytes * 1e-9 / t, (double)max_bytes * 1e-9 / t);
            }
        }
        free(secs);
    }

    /* 6. result vector (csrspmv.c:2940-2954) */
    if (!o.quiet) {
        if (o.verbose > 0) { fprintf(stderr, "mtxfile_write:\n"); clock_gettime(CLOCK_MONOTONIC, &t0); }
        mtx_write_vector(stdout, num_rows, y);
        if (o.verbose > 0) {
            clock_gettime(CLOCK_MONOTONIC, &t1);
            fprintf(stderr, "mtxfile_write done in %'.6f seconds\n", seconds_between(t0, t1));
        }
    }
    ellspmv_cuda_free_host(x);
    ellspmv_cuda_free_host(y);
    csrspmv_cuda_free(A);
    return EXIT_SUCCESS;

fail:
    ellspmv_cuda_free_host(x);
    ellspmv_cuda_free_host(y);
    csrspmv_cuda_free(A);
    return EXIT_FAILURE;
}
