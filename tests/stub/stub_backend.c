/*
 * stub_backend.c -- TEST INFRASTRUCTURE, not product: a stand-in for libfmdemod.so that lets the
 * C host (demodulator_b200/csrc/demodulator.c) run on a machine without a GPU, so that its slot
 * ring, its reader / writer threads and its positioned I/O can be tested for ORDER.  It does not
 * demodulate anything: output float k of a submit is a checksum of input bytes 4k..4k+3, so every
 * output value identifies the input position it came from.  fmd_wait sleeps a pseudo-random time
 * to shake the thread interleavings.  Only tests/test_cli_host_logic.py links this.
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "fmdemod.h"

struct fmd_ctx {
    fmd_config cfg;
    unsigned nslots;
    size_t slot_bytes;
    unsigned seed;
};

static char g_err[128] = "";
const char *fmd_last_error(void) { return g_err; }
int fmd_device_count(void) { return getenv("STUB_DEVICES") ? atoi(getenv("STUB_DEVICES")) : 1; }

void fmd_config_default(fmd_config *cfg) {
    memset(cfg, 0, sizeof *cfg);
    cfg->sample_rate = 125000.0;
    cfg->lowpass_out = 12500.0;
    cfg->epsilon = 0.3;
    cfg->out_degree = 3;
    cfg->mode = 0x10;
    cfg->chunk_bytes = FMD_DEFAULT_CHUNK_BYTES;
}

int fmd_create(const fmd_config *cfg, int device, fmd_ctx **out) {
    fmd_ctx *c = calloc(1, sizeof *c);
    c->cfg = *cfg;
    c->seed = 12345u + (unsigned) device;
    *out = c;
    return 0;
}
void fmd_destroy(fmd_ctx *c) { free(c); }
int fmd_ring_create(fmd_ctx *c, unsigned slots, size_t slot_bytes) {
    c->nslots = slots;
    c->slot_bytes = slot_bytes;
    return 0;
}
int fmd_host_alloc(void **p, size_t n) { *p = malloc(n ? n : 1); return *p ? 0 : -1; }
int fmd_host_free(void *p) { free(p); return 0; }
size_t fmd_out_count(const fmd_ctx *c, size_t nbytes) { (void) c; return nbytes >> 2; }
size_t fmd_max_out_count(const fmd_ctx *c, size_t nbytes) { (void) c; return nbytes >> 2; }

int fmd_submit(fmd_ctx *c, unsigned slot, const uint8_t *in, size_t nbytes, float *out) {
    if (slot >= c->nslots || nbytes > c->slot_bytes) {
        snprintf(g_err, sizeof g_err, "stub: bad submit");
        return -1;
    }
    for (size_t k = 0; k < nbytes / 4; ++k) {
        uint32_t w;
        memcpy(&w, in + 4 * k, 4);
        out[k] = (float) (w & 0xffffffu); /* exact in float32 */
    }
    return 0;
}

/* segmented quality mode: the stub checks what the host must get right -- the warm-up prefix is
 * exactly the stream bytes that precede the payload (input word k holds k) -- and then behaves
 * like fmd_submit on the payload */
uint64_t fmd_warmup_samples(const fmd_ctx *c) { (void) c; return getenv("STUB_WARM") ? (uint64_t) atoll(getenv("STUB_WARM")) : 0; }
size_t fmd_segment_out_count(const fmd_ctx *c, uint64_t first_sample, size_t nsamples) { (void) c; (void) first_sample; return nsamples >> 1; }
int fmd_submit_segment(fmd_ctx *c, unsigned slot, const uint8_t *in, size_t warm_bytes, size_t nbytes, uint64_t first_sample,
                       float *out) {
    if ((warm_bytes & 15u) || warm_bytes / 2 > first_sample || ((uintptr_t) in & 15u)) {
        snprintf(g_err, sizeof g_err, "stub: bad segment geometry (warm %zu first %llu)", warm_bytes, (unsigned long long) first_sample);
        return -1;
    }
    const uint64_t w0 = (first_sample * 2 - warm_bytes) / 4; /* index of the first word of the prefix */
    for (size_t k = 0; k < (warm_bytes + nbytes) / 4; ++k) {
        uint32_t w;
        memcpy(&w, in + 4 * k, 4);
        if (w != (uint32_t) (w0 + k)) {
            snprintf(g_err, sizeof g_err, "stub: segment word %zu holds %u, expected %u", k, w, (uint32_t) (w0 + k));
            return -1;
        }
    }
    const size_t keep = c->slot_bytes; /* the ring was sized for warm-up + payload */
    c->slot_bytes = (size_t) -1;
    int rc = fmd_submit(c, slot, in + warm_bytes, nbytes, out);
    c->slot_bytes = keep;
    return rc;
}

int fmd_wait(fmd_ctx *c, unsigned slot) {
    (void) slot;
    if (getenv("STUB_JITTER")) {
        c->seed = c->seed * 1103515245u + 12345u; /* racy on purpose: only the amount of sleep depends on it */
        struct timespec ts = {0, (long) ((c->seed >> 16) % 3000u) * 1000L};
        nanosleep(&ts, NULL);
    }
    return 0;
}
