/*
 * fmd_compat.c -- the reference's backend seam, re-implemented over the C ABI.
 *
 * peads/demodulator's main links against exactly two backend symbols
 * (/root/reference/src/main.c:44-47, include/matrix.h:59-62):
 *
 *     void *processMatrix(void *ctx);               pthread start routine, ctx = consumerArgs*
 *     void  allocateBuffer(void **buf, size_t len); backend allocates the reader's 1-slot buffer
 *
 * and chooses the implementation at link time (src/matrix.c:178,282 or cuda/matrix.cu:63,106).
 * This file is a third implementation: build the reference's own, unmodified src/main.c with
 * -DIS_NVIDIA and link it against libfmd_compat.so + libfmdemod.so (recipe in INTEGRATION.md;
 * tests/test_cli_gpu.py::test_reference_main_over_this_backend runs the prebuilt binary on the GPU box).
 *
 * The struct below restates the reference's consumerArgs field for field
 * (include/matrix.h:43-57; x86-64 offsets in SURVEY.md section 8b) -- it is the ABI of the seam.
 *
 * Hand-off protocol (src/matrix.c:236-242): wait `full`; lock; take bufSize bytes out of buf;
 * unlock; post `empty`; process; fwrite to outFile; loop while !exitFlag.  Deliberate
 * deviation, documented in DESIGN.md "tail policy": bufSize is sampled once, under the lock
 * (the reference re-reads it after the reader may have overwritten it); a hand-off of 0 bytes
 * emits nothing; a short final hand-off emits r/4 floats.
 */
#include <errno.h>
#include <pthread.h>
#include <semaphore.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "fmdemod.h"

typedef struct {
    sem_t *full, *empty;
    pthread_mutex_t mutex;
    FILE *outFile;
    void *buf;
    int exitFlag;
    uint8_t mode;
    long double sampleRate;
    long double lowpassIn;
    long double lowpassOut;
    long double epsilon;
    size_t inFilterDegree;
    size_t outFilterDegree;
    size_t bufSize;
} consumerArgs;

#if defined(__x86_64__)
_Static_assert(sizeof(consumerArgs) == 176, "consumerArgs layout differs from the reference's");
_Static_assert(__builtin_offsetof(consumerArgs, bufSize) == 160, "consumerArgs layout differs");
#endif

void allocateBuffer(void **buf, const size_t len) {
    if (fmd_host_alloc(buf, len)) {
        fprintf(stderr, "allocateBuffer: %s\n", fmd_last_error());
        *buf = NULL;
    }
}

/* wait on `full`; once exitFlag is up, give the reader a grace period to post its last hand-off */
static int take_full(consumerArgs *a) {
    if (!a->exitFlag) {
        while (sem_wait(a->full) && errno == EINTR) {}
        return 1;
    }
    struct timespec ts;
    clock_gettime(CLOCK_REALTIME, &ts);
    ts.tv_nsec += 50 * 1000 * 1000;
    if (ts.tv_nsec >= 1000000000L) { ts.tv_sec += 1; ts.tv_nsec -= 1000000000L; }
    for (;;) {
        if (!sem_timedwait(a->full, &ts)) return 1;
        if (errno != EINTR) return 0;
    }
}

void *processMatrix(void *ctx) {
    consumerArgs *args = ctx;
    const size_t chunk = args->bufSize;
    fmd_config cfg;
    fmd_ctx *fm = NULL;
    uint8_t *in[2] = {NULL, NULL};
    float *out[2] = {NULL, NULL};
    size_t pending[2] = {0, 0};
    unsigned n = 0;
    const char *dev = getenv("FMD_DEVICE");

    fmd_config_default(&cfg);
    cfg.sample_rate = (double) args->sampleRate;
    cfg.lowpass_in = (double) args->lowpassIn;
    cfg.lowpass_out = (double) args->lowpassOut;
    cfg.epsilon = (double) args->epsilon;
    cfg.in_degree = (uint32_t) args->inFilterDegree;
    cfg.out_degree = (uint32_t) args->outFilterDegree;
    cfg.mode = args->mode;
    cfg.chunk_bytes = chunk;

    if (fmd_create(&cfg, dev ? atoi(dev) : 0, &fm) || fmd_ring_create(fm, 2, chunk) ||
        fmd_host_alloc((void **) &in[0], chunk) || fmd_host_alloc((void **) &in[1], chunk) ||
        fmd_host_alloc((void **) &out[0], chunk) || fmd_host_alloc((void **) &out[1], chunk)) {
        fprintf(stderr, "processMatrix: %s\n", fmd_last_error());
        args->exitFlag = -5;
        sem_post(args->empty); /* do not leave the reader blocked */
        return NULL;
    }
    /* the reference resolves the inherited input degree in place (src/matrix.c:190-192) */
    fmd_get_config(fm, &cfg);
    args->inFilterDegree = cfg.in_degree;

    int failed = 0;
    for (;;) {
        if (!take_full(args)) break;
        const unsigned s = n & 1u;
        pthread_mutex_lock(&args->mutex);
        size_t len = args->bufSize;
        if (len > chunk) len = chunk;
        memcpy(in[s], args->buf, len);
        pthread_mutex_unlock(&args->mutex);
        sem_post(args->empty);

        /* After a failure the hand-offs are still taken (and dropped): the reference's reader
         * (src/main.c:72-91) blocks in sem_wait(empty) until the worker takes its buffer and
         * overwrites exitFlag with ferror() on every turn, so the only way not to hang it is to
         * keep the protocol going until it sees the end of its input. */
        if (failed) continue;
        len &= ~(size_t) 3;
        if (len) {
            if (fmd_submit(fm, s, in[s], len, out[s])) {
                failed = 1;
                fprintf(stderr, "processMatrix: %s (dropping the rest of the stream)\n", fmd_last_error());
                continue;
            }
            pending[s] = len >> 2;
        }
        /* while this chunk is in flight, write the previous one */
        if (pending[s ^ 1u]) {
            if (fmd_wait(fm, s ^ 1u)) {
                failed = 1;
                fprintf(stderr, "processMatrix: %s (dropping the rest of the stream)\n", fmd_last_error());
                continue;
            }
            fwrite(out[s ^ 1u], sizeof(float), pending[s ^ 1u], args->outFile);
            pending[s ^ 1u] = 0;
        }
        if (len) ++n;
    }
    for (unsigned k = 0; k < 2 && !failed; ++k) {
        const unsigned s = (n + k) & 1u; /* oldest first */
        if (pending[s]) {
            if (fmd_wait(fm, s)) {
                fprintf(stderr, "processMatrix: %s\n", fmd_last_error());
                break;
            }
            fwrite(out[s], sizeof(float), pending[s], args->outFile);
            pending[s] = 0;
        }
    }
    fflush(args->outFile);
    fmd_destroy(fm);
    for (unsigned k = 0; k < 2; ++k) {
        fmd_host_free(in[k]);
        fmd_host_free(out[k]);
    }
    fmd_host_free(args->buf);
    args->buf = NULL;
    return NULL;
}
