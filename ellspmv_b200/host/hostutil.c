/* hostutil.c -- helpers shared by the ellspmv and csrspmv host programs. */
#include "hostutil.h"

#include <errno.h>
#include <limits.h>
#include <stdlib.h>
#include <string.h>

const char *prog = "ellspmv";

/* value of "--name=V" or "--name V"; advances *i when the next argv is used */
const char *optval(int argc, char **argv, int *i, const char *name)
{
    size_t n = strlen(name);
    if (strncmp(argv[*i], name, n) != 0) return NULL;
    if (argv[*i][n] == '=') return argv[*i] + n + 1;
    if (argv[*i][n] == '\0' && *i + 1 < argc) return argv[++*i];
    return NULL;
}

int to_int(const char *s, int *out)
{
    char *end;
    errno = 0;
    long long v = strtoll(s, &end, 10);
    if (errno || end == s || *end != '\0') return EINVAL;
    if (v < INT_MIN || v > INT_MAX) return ERANGE;
    *out = (int)v;
    return 0;
}

double seconds_between(struct timespec t0, struct timespec t1)
{
    return (double)(t1.tv_sec - t0.tv_sec) + (double)(t1.tv_nsec - t0.tv_nsec) * 1e-9;
}

/* read a dense vector file into v[0..n); the "expected vector" message is the reference's */
int read_vector_file(const char *path, int gzip, idx_t n, double *v, int verbose)
{
    struct timespec t0, t1;
    if (verbose > 0) { fprintf(stderr, "mtxfile_read: "); clock_gettime(CLOCK_MONOTONIC, &t0); }
    struct mtx_stream *s = mtx_open(path, gzip);
    if (!s) { fprintf(stderr, "%s: %s: %s\n", prog, path, strerror(errno)); return -2; }
    struct mtx_header h;
    int64_t lines = 0, bytes = 0;
    int err = mtx_read_header(s, &h, &lines, &bytes);
    if (err) {
        if (verbose > 0) fprintf(stderr, "\n");
        fprintf(stderr, "%s: %s:%" PRId64 ": %s\n", prog, path, lines + 1, strerror(err));
        mtx_close(s);
        return -2;
    }
    if (h.object != MTX_VECTOR || h.format != MTX_ARRAY || h.num_rows != n) {
        if (verbose > 0) fprintf(stderr, "\n");
        fprintf(stderr, "%s: %s:%" PRId64 ": expected vector in array format of size %" PRIdx "\n",
                prog, path, lines + 1, n);
        mtx_close(s);
        return -2;
    }
    err = mtx_read_vector(s, h.field, n, v, &lines, &bytes);
    if (err) {
        if (verbose > 0) fprintf(stderr, "\n");
        fprintf(stderr, "%s: %s:%" PRId64 ": %s\n", prog, path, lines + 1, strerror(err));
        mtx_close(s);
        return -2;
    }
    if (verbose > 0) {
        clock_gettime(CLOCK_MONOTONIC, &t1);
        fprintf(stderr, "%'.6f seconds (%'.1f MB/s)\n", seconds_between(t0, t1),
                1.0e-6 * (double)bytes / seconds_between(t0, t1));
    }
    mtx_close(s);
    return 0;
}

