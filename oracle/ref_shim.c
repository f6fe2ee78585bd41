/*
 * oracle/ref_shim.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Compiles the UNMODIFIED reference sources where they lie under
 * $(REF)=/root/reference (src/filter.c, src/matrix.c are pulled in with
 * #include; nothing is copied into this repository) and exposes their
 * file-static hot-path functions through thin extern wrappers so that
 * Python/ctypes and the C tests can call the real thing.
 *
 * Built only by oracle/Makefile into oracle/_ref/libdemod_ref.so.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load it.
 *
 * Reference entry points wrapped (file:line in /root/reference):
 *   processFilterOption   src/matrix.c:25-80
 *   shiftOrigin           src/matrix.c:82-98
 *   fmDemod               src/matrix.c:159-176
 *   applyFilter           src/filter.c:212-231
 *   applyComplexFilter    src/filter.c:233-259
 *   processMatrix loop    src/matrix.c:236-273 (driven here one chunk at a
 *                         time, synchronously, with the same scratch layout)
 */
#include <stdlib.h>

/* highpassDc (-q2, src/matrix.c:142-157) reads a few floats past its own calloc'd scratch.  In the
 * reference process that scratch (bufSize floats >= 128 KiB, allocated before anything is freed)
 * is a fresh mmap'd block whose slack is zero -- the reference CLI's output, which the fixtures
 * are generated from, shows exactly that.  Inside another process (Python) the same calloc may be
 * served from a recycled heap chunk, so the allocator behaviour is pinned here: every calloc made
 * by the included sources gets 64 zeroed bytes of slack.  The sources themselves are untouched. */
static void *shim_calloc(size_t n, size_t size) {
    return calloc(n * size + 64, 1);
}
#define calloc(n, size) shim_calloc((n), (size))
#include "src/filter.c"
#include "src/matrix.c"
#undef calloc

static size_t sectionsFor(size_t degree) {
    return (degree >> 1) + (degree & 1);
}

void ref_design(unsigned mode, size_t degree, double fc, double fs, double eps,
                float *sosOut /* [ceil(degree/2)][6] */) {
    REAL(*sos)[6] = (REAL(*)[6]) sosOut;
    processFilterOption((uint8_t) mode, degree, sos, (LREAL) fc, (LREAL) fs, (LREAL) eps);
}

void ref_shift_origin(const uint8_t *in, size_t len, float *out) {
    shiftOrigin((void *) in, len, out);
}

void ref_normalize_input(const uint8_t *in, size_t len, float *out) {
    normalizeInput((void *) in, len, out);
}

void ref_fm_demod(const float *in, size_t len, float *out) {
    fmDemod(in, len, out);
}

void ref_apply_filter(float *x, float *y, size_t len, size_t sosLen, const float *sos) {
    applyFilter(x, y, len, sosLen, (const REAL(*)[6]) sos);
}

void ref_apply_complex_filter(float *x, float *y, size_t len, size_t sosLen, const float *sos) {
    applyComplexFilter(x, y, len, sosLen, (const REAL(*)[6]) sos);
}

/*
 * Run `nchunks` whole chunks of `bufSize` bytes through the reference's
 * default-mode (-q0, demod on) pipeline.  The scratch arena is laid out the
 * way the worker thread lays it out (one calloc'd float arena of 2*B or 4*B
 * floats whose thirds/halves are used as x / y regions, cleared after every
 * chunk; a B-float demod buffer cleared once), so that the overrun of
 * applyComplexFilter into the region applyFilter accumulates into is produced
 * by the real code and not by a model of it.
 * Returns the number of floats written (nchunks * bufSize/4), or 0 on error.
 */
size_t ref_run_chunks(const uint8_t *in, size_t nchunks, size_t bufSize, unsigned mode,
                      double fs, double fcIn, double fcOut, double eps,
                      size_t degIn, size_t degOut, float *out) {
    if (fcIn != 0. && !degIn) degIn = degOut;            /* src/matrix.c:190-192 */
    if (fcIn == 0. && degIn) return 0;  /* -D without -L: reference reads an undesigned filter */
    if (!degOut || (bufSize & 3)) return 0;
    if (((mode >> 2) & 3) == 1) return 0; /* -q1: state in a function-local static, cannot be driven chunk by chunk from here */
    if (((mode >> 2) & 3) == 2) {
        /* -q2: highpassDc allocates its scratch once, on its first call, for that call's length:
         * one chunk size per process (see shim_calloc above for the slack it reads) */
        static size_t q2_len = 0;
        if (q2_len && q2_len != bufSize) return 0;
        q2_len = bufSize;
        samplingRate = (LREAL) fs; /* src/matrix.c:206 */
    }
    const size_t nOut = bufSize >> 2;
    const size_t secOut = sectionsFor(degOut), secIn = sectionsFor(degIn);
    REAL sosOut[secOut ? secOut : 1][6];
    REAL sosIn[secIn ? secIn : 1][6];
    const size_t arenaFloats = (bufSize << 1) << (fcIn != 0. ? 1 : 0);
    REAL *arena = calloc(arenaFloats, sizeof(REAL));
    REAL *demod = calloc(bufSize, sizeof(REAL));
    size_t c;
    if (!arena || !demod) return 0;

    processFilterOption(mode & 1, degOut, sosOut, (LREAL) fcOut, (LREAL) fs, (LREAL) eps);
    if (fcIn != 0.)
        processFilterOption((mode >> 1) & 1, degIn, sosIn, (LREAL) fcIn, (LREAL) fs, (LREAL) eps);

    for (c = 0; c < nchunks; ++c) {
        REAL *res;
        if (((mode >> 2) & 3) == 3) normalizeInput((void *) (in + c * bufSize), bufSize, arena);  /* -q3 */
        else if (((mode >> 2) & 3) == 2) highpassDc((void *) (in + c * bufSize), bufSize, arena); /* -q2 */
        else shiftOrigin((void *) (in + c * bufSize), bufSize, arena);
        if (!degIn) {
            res = arena + bufSize;
            fmDemod(arena, bufSize, demod);
            applyFilter(demod, res, nOut, secOut, sosOut);
        } else {
            res = arena + (bufSize << 1);
            applyComplexFilter(arena, arena + bufSize, bufSize, secIn, sosIn);
            fmDemod(arena + bufSize, bufSize, demod);
            applyFilter(demod, res, nOut, secOut, sosOut);
        }
        memcpy(out + c * nOut, res, nOut * sizeof(REAL));
        memset(arena, 0, arenaFloats * sizeof(REAL));
    }
    free(arena);
    free(demod);
    return nchunks * nOut;
}
