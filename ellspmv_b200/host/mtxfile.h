/*
 * mtxfile.h -- strict Matrix Market reader/writer of the host programs.
 *
 * Accepts exactly what the reference's reader accepts (ellspmv.c:657-929):
 * a "%%MatrixMarket <object> <format> <field> <symmetry>" header with single
 * spaces, '%' comment lines, a size line, then one entry per line with
 * single-space separators; lines are limited to _SC_LINE_MAX bytes.
 * Optional gzip input when built with -DHAVE_LIBZ.
 */
#ifndef ELLSPMV_HOST_MTXFILE_H
#define ELLSPMV_HOST_MTXFILE_H

#include <stdint.h>
#include <stdio.h>

#include "idx.h"

enum mtx_object { MTX_MATRIX, MTX_VECTOR };
enum mtx_format { MTX_ARRAY, MTX_COORDINATE };
enum mtx_field { MTX_REAL, MTX_INTEGER, MTX_PATTERN };
enum mtx_symmetry { MTX_GENERAL, MTX_SYMMETRIC };

struct mtx_stream;   /* stdio or zlib */

struct mtx_header {
    enum mtx_object object;
    enum mtx_format format;
    enum mtx_field field;
    enum mtx_symmetry symmetry;
    idx_t num_rows, num_columns;
    int64_t num_nonzeros;
};

/* returns NULL and sets errno on failure */
struct mtx_stream *mtx_open(const char *path, int gzip);
void mtx_close(struct mtx_stream *s);

/* Each returns 0, a positive errno value, or -1 for a premature end of file
 * (the reference's convention, ellspmv.c:665); *lines and *bytes advance as
 * input is consumed so the caller can report "path:line: error". */
int mtx_read_header(struct mtx_stream *s, struct mtx_header *h, int64_t *lines, int64_t *bytes);
int mtx_read_coordinate(struct mtx_stream *s, const struct mtx_header *h,
                        idx_t *rowidx, idx_t *colidx, double *a, int64_t *lines, int64_t *bytes);
/*
 * Same result as mtx_read_coordinate (arrays, *lines, *bytes, error code and
 * error line), but the entry lines are parsed by all OpenMP threads from a
 * memory map of the file: the reference's serial fgets/strtoll/strtod loop
 * (ellspmv.c:808-888) runs at ~100 MB/s and dominates the wall clock for real
 * matrices (README: 32 s of reading for a 0.01 s SpMV).  Lines that are not
 * in the plain "digits SP digits [SP number] NL" form, gzip streams, and any
 * error make it fall back to the serial reader, which then reports exactly
 * what the reference would.
 */
int mtx_read_coordinate_parallel(struct mtx_stream *s, const struct mtx_header *h,
                                 idx_t *rowidx, idx_t *colidx, double *a, int64_t *lines, int64_t *bytes);

int mtx_read_vector(struct mtx_stream *s, enum mtx_field field, int64_t n, double *x,
                    int64_t *lines, int64_t *bytes);

/* "%%MatrixMarket vector array real general", n, then %.15g per line
 * (ellspmv.c:1905-1907) */
void mtx_write_vector(FILE *f, idx_t n, const double *y);

#endif
