/*
 * ellspmv_main.c -- host program `ellspmv`: y := A*x + y with A in ELLPACK
 * format, the multiplication running on a B200 through the C ABI in
 * include/ellspmv_cuda.h.
 *
 * Drop-in for the reference's `ellspmv` program (ellspmv.c:1226-1917): same
 * positional arguments and options, same Matrix Market reader, same COO->ELL
 * conversion (bit-exact arrays), same --repeat/--warmup/--verbose timing
 * lines on stderr and the same result vector on stdout.  The only part that
 * differs is the call site of the kernel (ellspmv.c:1766-1767, 1841-1842):
 * instead of every OpenMP thread calling `ellgemv`, the master thread calls
 *     ellspmv_cuda_upload()   once, after ell_from_coo
 *     ellspmv_cuda_spmv()     for the warm-up + repeat launches
 *     ellspmv_cuda_free()
 *
 * Known differences, on purpose:
 *   - --separate-diagonal does what the reference's functions do when their
 *     flags arrive in DECLARED order (diagonal summed into `ad`, K counted
 *     without it, kernel ellgemvsd / ellgemv16sd): the reference's main()
 *     passes the two flags swapped into ell_from_coo and overruns its arrays
 *     (ellspmv.c:1094-1095 vs 1468-1471).  --sort-rows sorts every row's
 *     entries by column like the reference's CSR program does; the
 *     reference's ELL rowsort sorts the wrong ranges (ellspmv.c:1121-1123);
 *   - x from a file is read with num_columns entries (the reference reads
 *     num_rows, ellspmv.c:1574-1575, which is only right for square A);
 *   - a matrix with more rows than columns works (the reference overruns its
 *     `ellad` array, ellspmv.c:1447-1467, and aborts);
 *   - new options for the GPU path (see --help).
 */
#include <errno.h>
#include <float.h>
#include <locale.h>
#include <stdbool.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <unistd.h>

#include "../../include/ellspmv_cuda.h"
#include "convert.h"
#include "idx.h"
#include "hostutil.h"
#include "mtxfile.h"


static const char *version = "1.10-b200";

struct options {
    const char *Apath, *xpath, *ypath;
    int gzip;
    bool separate_diagonal, sort_rows;
    int repeat, warmup, verbose, quiet;
    /* GPU options */
    unsigned flags;
    bool iterate, device_convert;
    const char *synthetic;
    int device, gpus;
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
    fprintf(f, "  A    path to Matrix Market file for the matrix A\n");
    fprintf(f, "  x    optional path to Matrix Market file for the vector x\n");
    fprintf(f, "  y    optional path for to Matrix Market file for the vector y\n");
    fprintf(f, "\n");
    fprintf(f, " Other options are:\n");
#ifdef HAVE_LIBZ
    fprintf(f, "  -z, --gzip, --gunzip, --ungzip    filter files through gzip\n");
#endif
    fprintf(f, "  --separate-diagonal  store diagonal nonzeros separately\n");
    fprintf(f, "  --sort-rows          sort nonzeros by column within each row\n");
    fprintf(f, "  --repeat=N           repeat matrix-vector multiplication N times\n");
    fprintf(f, "  --warmup=N                perform N additional warmup iterations\n");
    fprintf(f, "  -q, --quiet          do not print Matrix Market output\n");
    fprintf(f, "  -v, --verbose        be more verbose\n");
    fprintf(f, "\n");
    fprintf(f, " Options for the CUDA path are:\n");
    fprintf(f, "  --kernel=auto|thread|longrow|warp\n");
    fprintf(f, "                       auto (default): thread-per-row, or the long-row kernel for few long\n");
    fprintf(f, "                       rows -- both bit-exact; warp: sub-warp-per-row (tolerance mode)\n");
    fprintf(f, "  --skip-padding       stream only the slots that count (rows sorted by length, width per\n");
    fprintf(f, "                       slice); exact for finite x\n");
    fprintf(f, "  --no-staged-gather   never take the staged gather for scattered matrices (auto tries it)\n");
    fprintf(f, "  --fma                allow fused multiply-add (tolerance mode)\n");
    fprintf(f, "  --rows-per-thread=N  1, 2 or 4 rows per thread [by row length]\n");
    fprintf(f, "  --l2-persist-x       L2 persisting access window over x\n");
    fprintf(f, "  --wide-index         keep 64-bit column indices 64-bit on the device (default: stored as\n");
    fprintf(f, "                       32-bit when the matrix has fewer than 2^31 columns)\n");
    fprintf(f, "  --column-blocked     bin entries by column block so x stays in L2 (tolerance mode;\n");
    fprintf(f, "                       for scattered matrices whose x is larger than the L2 cache)\n");
    fprintf(f, "  --staged-gather      column blocks with the gather staged through device memory: the\n");
    fprintf(f, "                       same bits as the default kernel, for the same kind of matrix\n");
    fprintf(f, "  --no-pattern         always stream the column indices (default: groups of rows whose\n");
    fprintf(f, "                       indices are row + fixed offsets take them from a small table)\n");
    fprintf(f, "  --iterate            compute x := A*x repeatedly (y := A^N x); square A only\n");
    fprintf(f, "  --synthetic=SPEC     build A on the device instead of reading a file:\n");
    fprintf(f, "                       laplace2d:NX,NY | stencil27:NX,NY,NZ | random:ROWS,COLS,K[,SEED]\n");
    fprintf(f, "  --device-convert     convert COO to ELL on the device (same arrays as the host conversion)\n");
    fprintf(f, "  --gpus=N             shard the rows over CUDA devices 0..N-1 of this machine [1]\n");
    fprintf(f, "  --device=N           CUDA device ordinal [current]\n");
    fprintf(f, "\n");
    fprintf(f, "  -h, --help           display this help and exit\n");
    fprintf(f, "  --version            display version information and exit\n");
}

static void print_version(FILE *f)
{
    fprintf(f, "%s %s\n", prog, version);
    fprintf(f, "row/column offsets: %d-bit\n", IDX_BITS);
#ifdef HAVE_LIBZ
    fprintf(f, "zlib: yes\n");
#else
    fprintf(f, "zlib: no\n");
#endif
    int n = 0;
    ellspmv_cuda_device_count(&n);
    fprintf(f, "CUDA: libellspmv_cuda %d.%02d, %d device(s)\n", ellspmv_cuda_version() / 100,
            ellspmv_cuda_version() % 100, n);
}

/* returns 0, or errno with *bad set to the offending argument index */
static int parse_options(int argc, char **argv, struct options *o, int *bad)
{
    memset(o, 0, sizeof(*o));
    o->repeat = 1;
    o->device = -1;
    o->gpus = 1;
    int npos = 0;
    bool only_positional = false;
    for (int i = 1; i < argc; i++) {
        *bad = i;
        const char *a = argv[i], *v;
        if (!only_positional) {
            if (!strcmp(a, "--separate-diagonal")) { o->separate_diagonal = true; continue; }
            if (!strcmp(a, "--sort-rows")) { o->sort_rows = true; continue; }
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
                if (!strcmp(v, "thread")) o->flags |= ELLSPMV_CUDA_KERNEL_THREAD;
                else if (!strcmp(v, "warp")) o->flags |= ELLSPMV_CUDA_KERNEL_WARP;
                else if (!strcmp(v, "longrow")) o->flags |= ELLSPMV_CUDA_KERNEL_LONGROW;
                else if (strcmp(v, "auto")) return EINVAL;
                continue;
            }
            if (!strcmp(a, "--fma")) { o->flags |= ELLSPMV_CUDA_FMA; continue; }
            if (!strcmp(a, "--skip-padding")) { o->flags |= ELLSPMV_CUDA_SKIP_PADDING; continue; }
            if (!strcmp(a, "--no-staged-gather")) { o->flags |= ELLSPMV_CUDA_NO_STAGED_GATHER; continue; }
            if (!strcmp(a, "--l2-persist-x")) { o->flags |= ELLSPMV_CUDA_L2_PERSIST_X; continue; }
            if (!strcmp(a, "--narrow-index")) { o->flags |= ELLSPMV_CUDA_NARROW_INDEX; continue; }
            if (!strcmp(a, "--wide-index")) { o->flags |= ELLSPMV_CUDA_WIDE_INDEX; continue; }
            if (!strcmp(a, "--column-blocked")) { o->flags |= ELLSPMV_CUDA_COLUMN_BLOCKED; continue; }
            if (!strcmp(a, "--staged-gather")) { o->flags |= ELLSPMV_CUDA_STAGED_GATHER; continue; }
            if (!strcmp(a, "--no-pattern")) { o->flags |= ELLSPMV_CUDA_NO_PATTERN; continue; }
            if (!strncmp(a, "--rows-per-thread", 17) && (a[17] == '=' || a[17] == '\0')) {
                int r;
                if (!(v = optval(argc, argv, &i, "--rows-per-thread")) || to_int(v, &r)) return EINVAL;
                if (r != 1 && r != 2 && r != 4) return EINVAL;
                o->flags = (o->flags & ~(unsigned)ELLSPMV_CUDA_ROWS_PER_THREAD_MASK) |
                           ((unsigned)r << ELLSPMV_CUDA_ROWS_PER_THREAD_SHIFT);
                continue;
            }
            if (!strcmp(a, "--iterate")) { o->iterate = true; continue; }
            if (!strcmp(a, "--device-convert")) { o->device_convert = true; continue; }
            if (!strncmp(a, "--synthetic", 11) && (a[11] == '=' || a[11] == '\0')) {
                if (!(o->synthetic = optval(argc, argv, &i, "--synthetic"))) return EINVAL;
                continue;
            }
            if (!strncmp(a, "--gpus", 6) && (a[6] == '=' || a[6] == '\0')) {
                if (!(v = optval(argc, argv, &i, "--gpus")) || to_int(v, &o->gpus) || o->gpus < 1) return EINVAL;
                continue;
            }
            if (!strncmp(a, "--device", 8) && (a[8] == '=' || a[8] == '\0')) {
                if (!(v = optval(argc, argv, &i, "--device")) || to_int(v, &o->device)) return EINVAL;
                continue;
            }
            if (!strcmp(a, "-h") || !strcmp(a, "--help")) { help(stdout); exit(EXIT_SUCCESS); }
            if (!strcmp(a, "--version")) { print_version(stdout); exit(EXIT_SUCCESS); }
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
    else if (!strcmp(name, "stencil27s") && n == 3) { *kind = ELLSPMV_CUDA_GEN_STENCIL27; vals[0] = 0.5; vals[1] = -1.0 / 52.0; }
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
    if (o.sort_rows && o.synthetic) {
        fprintf(stderr, "%s: --sort-rows is not supported with --synthetic\n", prog);
        return EXIT_FAILURE;
    }
    if (o.separate_diagonal && o.synthetic) {
        fprintf(stderr, "%s: --separate-diagonal is not supported with --synthetic\n", prog);
        return EXIT_FAILURE;
    }

    ellspmv_cuda_matrix *A = NULL;
    idx_t num_rows = 0, num_columns = 0, rowsize = 0, diagsize = 0;
    int64_t num_nonzeros = 0, ellsize = 0;

    if (o.synthetic) {
        /* build the matrix on the device (shapes too large for a text file) */
        int kind = 0;
        int64_t dims[3];
        double vals[2] = {0, 0};
        uint64_t seed;
        if (parse_synthetic(o.synthetic, &kind, dims, vals, &seed)) {
            fprintf(stderr, "%s: %s --synthetic=%s\n", prog, strerror(EINVAL), o.synthetic);
            return EXIT_FAILURE;
        }
        if (o.verbose > 0) { fprintf(stderr, "cuda_generate: "); clock_gettime(CLOCK_MONOTONIC, &t0); }
        if (o.gpus > 1) err = ellspmv_cuda_generate_sharded(&A, kind, dims, vals, seed, IDX_BITS, o.gpus, o.flags);
        else err = ellspmv_cuda_generate(&A, kind, dims, vals, seed, IDX_BITS, 0, -1, o.device, o.flags);
        if (err) {
            if (o.verbose > 0) fprintf(stderr, "\n");
            fprintf(stderr, "%s: %s (%s)\n", prog, strerror(err), ellspmv_cuda_last_error());
            return EXIT_FAILURE;
        }
        ellspmv_cuda_info info;
        ellspmv_cuda_get_info(A, &info);
        num_rows = (idx_t)info.num_rows;
        num_columns = (idx_t)info.num_columns;
        rowsize = (idx_t)info.rowsize;
        ellsize = info.num_rows * info.rowsize;
        diagsize = num_rows < num_columns ? num_rows : num_columns;
        num_nonzeros = ellsize;
        if (o.verbose > 0) {
            clock_gettime(CLOCK_MONOTONIC, &t1);
            fprintf(stderr, "%'.6f seconds, %'" PRIdx " rows, %'" PRId64 " nonzeros, %'" PRIdx " nonzeros per row\n",
                    seconds_between(t0, t1), num_rows, ellsize + num_rows, rowsize);
        }
    } else {
        /* 2. read the matrix (ellspmv.c:1264-1377) */
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
        num_rows = h.num_rows; num_columns = h.num_columns; num_nonzeros = h.num_nonzeros;
        size_t nz = num_nonzeros > 0 ? (size_t)num_nonzeros : 1;
        idx_t *rowidx = malloc(nz * sizeof(idx_t));
        idx_t *colidx = malloc(nz * sizeof(idx_t));
        double *a = malloc(nz * sizeof(double));
        if (!rowidx || !colidx || !a) {
            if (o.verbose > 0) fprintf(stderr, "\n");
            fprintf(stderr, "%s: %s\n", prog, strerror(ENOMEM));
            return EXIT_FAILURE;
        }
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

        /* 3. convert to ELLPACK (ellspmv.c:1379-1486) */
        if (o.verbose > 0) { fprintf(stderr, "ell_from_coo: "); clock_gettime(CLOCK_MONOTONIC, &t0); }
        if (o.device_convert && !o.separate_diagonal && !o.sort_rows && o.gpus == 1) {
            /* stable sort by row on the device instead of the serial host scatter */
            err = ellspmv_cuda_upload_coo(&A, IDX_BITS, num_rows, num_columns, num_nonzeros, rowidx, colidx, a,
                                          o.device, o.flags);
            free(a); free(colidx); free(rowidx);
            if (err) {
                if (o.verbose > 0) fprintf(stderr, "\n");
                fprintf(stderr, "%s: %s (%s)\n", prog, strerror(err), ellspmv_cuda_last_error());
                return EXIT_FAILURE;
            }
            ellspmv_cuda_info info;
            ellspmv_cuda_get_info(A, &info);
            rowsize = (idx_t)info.rowsize;
            ellsize = info.num_rows * info.rowsize;
            diagsize = num_rows < num_columns ? num_rows : num_columns;
            if (o.verbose > 0) {
                clock_gettime(CLOCK_MONOTONIC, &t1);
                fprintf(stderr, "%'.6f seconds, %'" PRIdx " rows, %'" PRId64 " nonzeros, %'" PRIdx " nonzeros per row\n",
                        seconds_between(t0, t1), num_rows, ellsize + num_rows, rowsize);
            }
        } else {
        struct ell_matrix ell;
        if (o.separate_diagonal && num_rows > num_columns) {
            if (o.verbose > 0) fprintf(stderr, "\n");
            fprintf(stderr, "%s: --separate-diagonal needs rows <= columns (the kernel reads x[i] for every row)\n", prog);
            return EXIT_FAILURE;
        }
        err = ell_from_coo(&ell, num_rows, num_columns, num_nonzeros, rowidx, colidx, a, o.separate_diagonal);
        free(a); free(colidx); free(rowidx);
        /* intended --sort-rows: every row's entries by column, padding last (the
         * reference's ELL version is broken, see convert.h) */
        if (!err && o.sort_rows) err = ell_sort_rows(&ell);
        if (err) {
            if (o.verbose > 0) fprintf(stderr, "\n");
            fprintf(stderr, "%s: %s\n", prog, strerror(err));
            return EXIT_FAILURE;
        }
        rowsize = ell.rowsize; ellsize = ell.ellsize; diagsize = ell.diagsize;
        if (o.verbose > 0) {
            clock_gettime(CLOCK_MONOTONIC, &t1);
            fprintf(stderr, "%'.6f seconds, %'" PRIdx " rows, %'" PRId64 " nonzeros, %'" PRIdx " nonzeros per row\n",
                    seconds_between(t0, t1), num_rows, ellsize + num_rows, rowsize);
        }

        /* device copy + re-layout: the one step the reference does not have */
        if (o.verbose > 0) { fprintf(stderr, "cuda_upload: "); clock_gettime(CLOCK_MONOTONIC, &t0); }
        if (o.device >= 0 && o.gpus == 1)
            err = ellspmv_cuda_upload_shard(&A, IDX_BITS, num_rows, num_columns, rowsize, 0, num_rows,
                                            ell.colidx, ell.a, o.device, o.flags);
        else
            err = ellspmv_cuda_upload(&A, IDX_BITS, num_rows, num_columns, rowsize, ell.colidx, ell.a, o.gpus, o.flags);
        /* same dispatch as the reference: K == 16 takes the unrolled kernel's
         * summation order (ellspmv.c:1759-1768) */
        if (!err && o.separate_diagonal) err = ellspmv_cuda_set_diagonal(A, ell.ad, rowsize == 16 ? 1 : 0);
        ell_free(&ell);
        if (err) {
            if (o.verbose > 0) fprintf(stderr, "\n");
            fprintf(stderr, "%s: %s (%s)\n", prog, strerror(err), ellspmv_cuda_last_error());
            return EXIT_FAILURE;
        }
        if (o.verbose > 0) {
            ellspmv_cuda_info info;
            ellspmv_cuda_get_info(A, &info);
            clock_gettime(CLOCK_MONOTONIC, &t1);
            fprintf(stderr, "%'.6f seconds, %d GPU(s), %'" PRId64 " bytes, sliced ELL %d rows/slice, "
                            "%d rows/thread, %d-bit indices, %'" PRId64 " rows on offset patterns\n",
                    seconds_between(t0, t1), info.num_gpus, info.device_bytes, info.slice_rows,
                    info.rows_per_thread, info.dev_idx_bits, info.pattern_rows);
        }
        }
    }

    /* 4. vectors (ellspmv.c:1488-1702), in pinned memory so the copies run at PCIe rate */
    double *x = NULL, *y = NULL;
    if (ellspmv_cuda_malloc_host((void **)&x, (int64_t)(num_columns > 0 ? num_columns : 1) * 8) ||
        ellspmv_cuda_malloc_host((void **)&y, (int64_t)(num_rows > 0 ? num_rows : 1) * 8)) {
        fprintf(stderr, "%s: %s (%s)\n", prog, strerror(ENOMEM), ellspmv_cuda_last_error());
        ellspmv_cuda_free(A);
        return EXIT_FAILURE;
    }
    for (idx_t j = 0; j < num_columns; j++) x[j] = 1.0;
    for (idx_t i = 0; i < num_rows; i++) y[i] = 0.0;
    if (o.xpath && read_vector_file(o.xpath, o.gzip, num_columns, x, o.verbose)) goto fail;
    if (o.ypath && read_vector_file(o.ypath, o.gzip, num_rows, y, o.verbose)) goto fail;

    /* 5. warm-up and timed multiplications (ellspmv.c:1745-1876): y keeps
     * accumulating across all of them, x stays fixed */
    {
        const int total = (o.warmup > 0 ? o.warmup : 0) + (o.repeat > 0 ? o.repeat : 0);
        double *secs = calloc((size_t)(total > 0 ? total : 1), sizeof(double));
        if (!secs) { fprintf(stderr, "%s: %s\n", prog, strerror(ENOMEM)); goto fail; }
        err = ellspmv_cuda_spmv(A, y, x, total, o.iterate ? ELLSPMV_CUDA_ITERATE : ELLSPMV_CUDA_ACCUMULATE, secs);
        if (err) {
            fprintf(stderr, "%s: %s (%s)\n", prog, strerror(err), ellspmv_cuda_last_error());
            free(secs);
            goto fail;
        }
        if (o.verbose > 0) {
            /* the reference's throughput model, padding and diagsize included (ellspmv.c:1857-1862) */
            const int64_t num_flops = 2 * (ellsize + diagsize);
            const int64_t min_bytes = (int64_t)num_rows * 8 + (int64_t)num_columns * 8 +
                                      ellsize * (int64_t)sizeof(idx_t) + ellsize * 8 + (int64_t)diagsize * 8;
            const int64_t max_bytes = (int64_t)num_rows * 8 + ellsize * 8 + ellsize * (int64_t)sizeof(idx_t) +
                                      ellsize * 8 + (int64_t)diagsize * 8 + (int64_t)diagsize * 8;
            double best = 0.0;
            for (int r = 0; r < total; r++) {
                const double t = secs[r];
                const char *label = !o.separate_diagonal ? "gemv" : (rowsize == 16 ? "gemv16sd" : "gemvsd");
                fprintf(stderr, r < o.warmup ? "%s (warmup): " : "%s: ", label);
                fprintf(stderr, "%'.6f seconds (%'.3f Gnz/s, %'.3f Gflop/s, %'.1f to %'.1f GB/s)\n", t,
                        (double)num_nonzeros * 1e-9 / t, (double)num_flops * 1e-9 / t,
                        (double)min_bytes * 1e-9 / t, (double)max_bytes * 1e-9 / t);
                if (r >= o.warmup && (best == 0.0 || t < best)) best = t;
            }
            if (best > 0.0) {
                /* effective-bytes roofline (values + indices + x + y read and written) */
                const double eff = (double)ellsize * (8.0 + sizeof(idx_t)) + 8.0 * num_columns + 16.0 * num_rows;
                fprintf(stderr, "cuda: best %'.6f seconds, %'.1f Gflop/s, %'.1f GB/s effective "
                                "(values+indices+x+y), device-event time per launch\n",
                        best, 2.0 * (double)ellsize * 1e-9 / best, eff * 1e-9 / best);
            }
        }
        free(secs);
    }

    /* 6. result vector (ellspmv.c:1898-1912) */
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
    ellspmv_cuda_free(A);
    return EXIT_SUCCESS;

fail:
    ellspmv_cuda_free_host(x);
    ellspmv_cuda_free_host(y);
    ellspmv_cuda_free(A);
    return EXIT_FAILURE;
}
