/* fmd_internal.h -- shared between the C and CUDA translation units of libfmdemod. */
#ifndef FMD_INTERNAL_H
#define FMD_INTERNAL_H

#include "fmdemod.h"

#ifdef __cplusplus
extern "C" {
#endif

/* records a printf-style message for fmd_last_error() and returns `code` */
int fmd_set_error(int code, const char *fmt, ...)
#if defined(__GNUC__)
    __attribute__((format(printf, 2, 3)))
#endif
    ;

static inline unsigned fmd_sections(unsigned degree) { return (degree >> 1) + (degree & 1u); }

#ifdef __cplusplus
}
#endif
#endif
