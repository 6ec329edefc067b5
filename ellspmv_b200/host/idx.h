/*
 * idx.h -- compile-time index type of the host programs, selected with
 * -DIDXTYPEWIDTH=32|64 exactly like the reference (ellspmv.c:112-130,
 * csrspmv.c:153-171).  The CUDA library itself takes the width at run time.
 */
#ifndef ELLSPMV_HOST_IDX_H
#define ELLSPMV_HOST_IDX_H

#include <inttypes.h>
#include <limits.h>
#include <stdint.h>

#ifndef IDXTYPEWIDTH
typedef int idx_t;
#define PRIdx "d"
#define IDX_T_MAX INT_MAX
#define IDX_T_MIN INT_MIN
#elif IDXTYPEWIDTH == 32
typedef int32_t idx_t;
#define PRIdx PRId32
#define IDX_T_MAX INT32_MAX
#define IDX_T_MIN INT32_MIN
#elif IDXTYPEWIDTH == 64
typedef int64_t idx_t;
#define PRIdx PRId64
#define IDX_T_MAX INT64_MAX
#define IDX_T_MIN INT64_MIN
#else
#error "IDXTYPEWIDTH must be 32 or 64"
#endif

#define IDX_BITS ((int)(sizeof(idx_t) * CHAR_BIT))

#endif
