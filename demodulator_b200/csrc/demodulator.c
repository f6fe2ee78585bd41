/*
 * demodulator.c -- the C host of the B200 FM demodulation path: same command line and
 * streaming contract as peads/demodulator's main (/root/reference/src/main.c:100-198),
 * over the C ABI of include/fmdemod.h (no CUDA call appears in this file).
 *
 *   demodulator -i <file|-> -o <file|-> [-L fcIn] [-l fcOut] [-S fs] [-d degOut] [-D degIn]
 *               [-m bits] [-e eps] [-b shift] [-q n] [-c n] [-r x] [-n x]
 *
 * in : raw unsigned 8-bit interleaved I,Q      out: raw native-endian float32,
 *      exactly bufSize/4 floats per bufSize input bytes, in order.
 *
 * Where the reference hands ONE 256 KiB buffer from its reader to its worker under a mutex
 * (src/main.c:72-91, src/matrix.c:238-242), this host keeps a ring of pinned slots, each
 * holding many chunks: a reader thread fills slots, the main thread submits them
 * (H2D / kernels / D2H on three streams inside the library), a writer thread drains them in
 * order.  Slots are cut short when the producer stalls, so a live source sees one-chunk
 * latency like the reference while a file or fast pipe gets multi-megabyte transfers.
 *
 * Environment (extensions; the option letters stay the reference's):
 *   FMD_DEVICES   comma list of GPU ordinals or "all" (default "0"): slots round-robin
 *   FMD_SLOT_MB   slot size in MiB (default 32), rounded to whole chunks
 *   FMD_SLOTS     ring depth per GPU (default 3)
 *   FMD_EXACT     1 = op-for-op audit kernel for every output
 *   FMD_STATS     1 = print throughput to stderr at exit
 */
#define _GNU_SOURCE
#include <errno.h>
#include <fcntl.h>
#include <poll.h>
#include <pthread.h>
#include <semaphore.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <unistd.h>

#include "fmdemod.h"

#define MAX_DEV 16

typedef struct {
    uint8_t *in;   /* pinned */
    float *out;    /* pinned */
    size_t nbytes; /* filled by the reader */
    int last;      /* reader hit EOF in (or right after) this slot */
    int dev;       /* which context owns it */
    unsigned ring; /* slot index inside that context's ring */
} slot_t;

typedef struct {
    int in_fd, out_fd;
    size_t chunk, slot_bytes;
    unsigned nslots;
    slot_t *slots;
    fmd_ctx *ctx[MAX_DEV];
    int ndev;
    sem_t free_, filled, inflight;
    volatile int error;
    unsigned long long bytes_in, floats_out;
} host_t;

static void die(const char *what) {
    fprintf(stderr, "demodulator: %s: %s\n", what, fmd_last_error());
    exit(2);
}

/* read until `want` bytes, EOF, or -- once at least one whole chunk is in -- the producer stalls */
static size_t fill_slot(host_t *h, uint8_t *dst, size_t want, int *eof) {
    size_t got = 0;
    *eof = 0;
    while (got < want) {
        if (got >= h->chunk && got % h->chunk == 0) {
            struct pollfd p = {h->in_fd, POLLIN, 0};
            if (poll(&p, 1, 0) == 0) break; /* nothing pending: ship what we have */
        }
        ssize_t n = read(h->in_fd, dst + got, want - got);
        if (n < 0) {
            if (errno == EINTR) continue;
            perror("demodulator: read");
            h->error = -2; /* the reference's exitFlag for ferror (src/main.c:78-81) */
            *eof = 1;
            break;
        }
        if (n == 0) {
            *eof = 1;
            break;
        }
        got += (size_t) n;
    }
    return got;
}

static void *reader_main(void *arg) {
    host_t *h = arg;
    unsigned i = 0;
    for (;;) {
        sem_wait(&h->free_);
        slot_t *s = &h->slots[i % h->nslots];
        int eof;
        s->nbytes = fill_slot(h, s->in, h->slot_bytes, &eof);
        s->last = eof;
        h->bytes_in += s->nbytes;
        sem_post(&h->filled);
        if (eof) break;
        ++i;
    }
    return NULL;
}

static int write_all(int fd, const void *buf, size_t n) {
    const char *p = buf;
    while (n) {
        ssize_t w = write(fd, p, n);
        if (w < 0) {
            if (errno == EINTR) continue;
            return -1;
        }
        p += w;
        n -= (size_t) w;
    }
    return 0;
}

static void *writer_main(void *arg) {
    host_t *h = arg;
    unsigned i = 0;
    for (;;) {
        sem_wait(&h->inflight);
        slot_t *s = &h->slots[i % h->nslots];
        if (s->nbytes) {
            if (fmd_wait(h->ctx[s->dev], s->ring)) die("fmd_wait");
            const size_t nf = s->nbytes >> 2;
            if (!h->error && write_all(h->out_fd, s->out, nf * sizeof(float))) {
                perror("demodulator: write");
                h->error = -4;
            }
            h->floats_out += nf;
        }
        const int last = s->last;
        sem_post(&h->free_);
        if (last) break;
        ++i;
    }
    return NULL;
}

static int parse_devices(const char *spec, int *devs) {
    int n = 0;
    if (!spec || !*spec) {
        devs[0] = 0;
        return 1;
    }
    if (!strcmp(spec, "all")) {
        int c = fmd_device_count();
        if (c <= 0) return 0;
        if (c > MAX_DEV) c = MAX_DEV;
        for (n = 0; n < c; ++n) devs[n] = n;
        return n;
    }
    char *dup = strdup(spec), *save = NULL;
    for (char *t = strtok_r(dup, ",", &save); t && n < MAX_DEV; t = strtok_r(NULL, ",", &save)) devs[n++] = atoi(t);
    free(dup);
    return n;
}

int main(int argc, char **argv) {
    fmd_config cfg;
    host_t h;
    int opt, ret = 0, devs[MAX_DEV];
    long shift;
    const char *e;

    fmd_config_default(&cfg);
    memset(&h, 0, sizeof h);
    h.in_fd = h.out_fd = -1;

    if (argc < 3) return -1; /* src/main.c:122-123 */

    while ((opt = getopt(argc, argv, "i:o:r:L:l:S:D:d:e:m:b:c:q:n:")) != -1) {
        switch (opt) {
            case 'i': /* any argument containing '-' means stdin (src/main.c:128) */
                h.in_fd = strstr(optarg, "-") ? STDIN_FILENO : open(optarg, O_RDONLY);
                if (h.in_fd < 0) { perror(NULL); ret += 1; }
                break;
            case 'o':
                h.out_fd = strstr(optarg, "-") ? STDOUT_FILENO : open(optarg, O_WRONLY | O_CREAT | O_TRUNC, 0644);
                if (h.out_fd < 0) { perror(NULL); ret += 1; }
                break;
            case 'L': cfg.lowpass_in = (double) strtold(optarg, NULL); break;
            case 'l': cfg.lowpass_out = (double) strtold(optarg, NULL); break;
            case 'S': cfg.sample_rate = (double) strtold(optarg, NULL); break;
            case 'D': cfg.in_degree = (uint32_t) strtoul(optarg, NULL, 10); break;
            case 'd': cfg.out_degree = (uint32_t) strtoul(optarg, NULL, 10); break;
            case 'e': cfg.epsilon = (double) (strtold(optarg, NULL) / 10.L); break;
            case 'm': cfg.mode |= (uint32_t) strtoul(optarg, NULL, 10); break;
            case 'b': /* src/main.c:164-173 */
                shift = strtol(optarg, NULL, 10);
                if (labs(shift) < 17)
                    cfg.chunk_bytes = shift < 1 ? (uint64_t) FMD_DEFAULT_CHUNK_BYTES >> -shift
                                                : (uint64_t) FMD_DEFAULT_CHUNK_BYTES << shift;
                break;
            case 'c': cfg.mode |= (uint32_t) strtoul(optarg, NULL, 10) << 4; break;
            case 'q': cfg.mode |= (uint32_t) strtoul(optarg, NULL, 10) << 2; break;
            default: break; /* -r, -n: accepted and ignored, like the reference */
        }
    }
    cfg.mode &= 0xffu; /* consumerArgs.mode is a uint8_t */
    if (ret) return ret;
    if (h.in_fd < 0 || h.out_fd < 0) {
        fprintf(stderr, "demodulator: both -i and -o are required\n");
        return 1;
    }
    if ((e = getenv("FMD_EXACT")) && atoi(e)) cfg.flags |= FMD_FLAG_EXACT;

    h.ndev = parse_devices(getenv("FMD_DEVICES"), devs);
    if (h.ndev <= 0) die("no CUDA device");
    for (int d = 0; d < h.ndev; ++d)
        if (fmd_create(&cfg, devs[d], &h.ctx[d])) die("fmd_create");

    h.chunk = (size_t) cfg.chunk_bytes;
    {
        size_t mb = (e = getenv("FMD_SLOT_MB")) ? (size_t) atol(e) : 32;
        unsigned depth = (e = getenv("FMD_SLOTS")) ? (unsigned) atoi(e) : 3;
        if (mb < 1) mb = 1;
        if (depth < 2) depth = 2;
        if (depth > 16) depth = 16;
        h.slot_bytes = (mb << 20) / h.chunk * h.chunk;
        if (!h.slot_bytes) h.slot_bytes = h.chunk;
        h.nslots = depth * (unsigned) h.ndev;
        h.slots = calloc(h.nslots, sizeof *h.slots);
        for (int d = 0; d < h.ndev; ++d)
            if (fmd_ring_create(h.ctx[d], depth, h.slot_bytes)) die("fmd_ring_create");
        for (unsigned i = 0; i < h.nslots; ++i) {
            slot_t *s = &h.slots[i];
            s->dev = (int) (i % (unsigned) h.ndev);
            s->ring = i / (unsigned) h.ndev;
            /* the backend chooses how the reader's buffer is allocated, as in the reference's
             * allocateBuffer seam (src/main.c:65): pinned, so copies are true DMA */
            if (fmd_host_alloc((void **) &s->in, h.slot_bytes) || fmd_host_alloc((void **) &s->out, h.slot_bytes))
                die("fmd_host_alloc");
        }
    }
    sem_init(&h.free_, 0, h.nslots);
    sem_init(&h.filled, 0, 0);
    sem_init(&h.inflight, 0, 0);

    struct timespec t0, t1;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    pthread_t rd, wr;
    pthread_create(&rd, NULL, reader_main, &h);
    pthread_create(&wr, NULL, writer_main, &h);
    for (unsigned i = 0;; ++i) {
        sem_wait(&h.filled);
        slot_t *s = &h.slots[i % h.nslots];
        /* tail policy: a final partial chunk of r bytes is a chunk of r&~3 bytes */
        s->nbytes &= ~(size_t) 3;
        if (s->nbytes && fmd_submit(h.ctx[s->dev], s->ring, s->in, s->nbytes, s->out)) die("fmd_submit");
        const int last = s->last;
        sem_post(&h.inflight);
        if (last) break;
    }
    pthread_join(rd, NULL);
    pthread_join(wr, NULL);
    clock_gettime(CLOCK_MONOTONIC, &t1);

    if ((e = getenv("FMD_STATS")) && atoi(e)) {
        const double sec = (double) (t1.tv_sec - t0.tv_sec) + 1e-9 * (double) (t1.tv_nsec - t0.tv_nsec);
        fprintf(stderr, "demodulator: %llu bytes in, %llu floats out, %.3f s, %.1f MS/s, %d GPU(s)\n", h.bytes_in,
                h.floats_out, sec, (double) h.bytes_in / 2e6 / sec, h.ndev);
    }
    for (int d = 0; d < h.ndev; ++d) fmd_destroy(h.ctx[d]);
    for (unsigned i = 0; i < h.nslots; ++i) {
        fmd_host_free(h.slots[i].in);
        fmd_host_free(h.slots[i].out);
    }
    close(h.out_fd);
    close(h.in_fd);
    return h.error ? 1 : ret;
}
