/*
 * ellspmv_oracle.c -- CPU oracle for the ELL/CSR SpMV hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  This is a plain-C restatement of the
 * reference algorithms (jamtrott/ellspmv: ellspmv.c, csrspmv.c) used as
 * the checker for the CUDA path.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load it; the
 * product (ellspmv_b200/, include/) never does.
 *
 * Parity pinning: the reference has no tests or golden vectors of its own
 * (SURVEY.md 4).  The oracle is pinned two ways: (1) against the
 * reference's only fixture, test.mtx, through known answers recorded in
 * tests/golden/ from the unmodified reference run in the build container;
 * (2) function by function against the unmodified reference compiled
 * from /root/reference into oracle/_ref/ (ref_wrap_*.c), in
 * tests/test_oracle_vs_reference.py on seeded random inputs.
 *
 * Build: see oracle/Makefile (-O2 -fopenmp -ffp-contract=off, never
 * -ffast-math: the reference's compiled loop is mul-then-add).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

static inline uint64_t oracle_splitmix64(uint64_t x)
{
    uint64_t z = x + 0x9e3779b97f4a7c15ULL;
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
    return z ^ (z >> 31);
}

uint64_t oracle_splitmix64_export(uint64_t x) { return oracle_splitmix64(x); }

#define OIDX int32_t
#define SUF 32
#include "oracle_impl.h"
#undef OIDX
#undef SUF

#define OIDX int64_t
#define SUF 64
#include "oracle_impl.h"
#undef OIDX
#undef SUF

/* number of OpenMP threads the SpMV loops will use (1 without OpenMP) */
#ifdef _OPENMP
#include <omp.h>
int oracle_num_threads(void) { return omp_get_max_threads(); }
#else
int oracle_num_threads(void) { return 1; }
#endif
