/* hostutil.h -- helpers shared by the ellspmv and csrspmv host programs. */
#ifndef ELLSPMV_HOST_UTIL_H
#define ELLSPMV_HOST_UTIL_H

#include <inttypes.h>
#include <stdio.h>
#include <time.h>

#include "idx.h"
#include "mtxfile.h"

extern const char *prog;   /* program_invocation_short_name of the reference */

/* value of "--name=V" or "--name V"; advances *i when the next argv is used */
const char *optval(int argc, char **argv, int *i, const char *name);
int to_int(const char *s, int *out);
double seconds_between(struct timespec t0, struct timespec t1);
/* dense vector file -> v[0..n); prints the reference's messages; 0 or nonzero */
int read_vector_file(const char *path, int gzip, idx_t n, double *v, int verbose);

#endif
