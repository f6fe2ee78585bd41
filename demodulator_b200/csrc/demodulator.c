/*
 * demodulator.c -- the C host of the B200 FM demodulation path: same command line and
 * streaming contract as peads/demodulator's main (/root/reference/src/main.c:100-198),
 * over the C ABI of include/fmdemod.h (no CUDA call appears in this file).
 *
 *   demodulator -i <file|-> -o <file|-> [-L fcIn] [-l fcOut] [-S fs] [-d degOut] [-D degIn]
 *               [-m bits] [-e eps] [-b shift] [-q n] [-c n] [-r x] [-n x]
 *               [-Q -r inputRate [-F s16]]  (extension: the opt-in quality pipeline, FMD_FLAG_QUALITY --
 *                                       recursive filters with carried state, adjacent-sample
 *                                       discriminator, output decimated to -S; over several GPUs the
 *                                       stream is cut into segments with a warm-up prefix)
 *
 * in : raw unsigned 8-bit interleaved I,Q      out: raw native-endian float32,
 *      exactly bufSize/4 floats per bufSize input bytes, in order.
 *
 * Where the reference hands ONE 256 KiB buffer from its reader to its worker under a mutex
 * (src/main.c:72-91, src/matrix.c:238-242), this host keeps a ring of pinned slots, each
 * holding many chunks: reader threads fill slots, the main thread submits them (H2D / kernels /
 * D2H on three streams inside the library), writer threads drain them; a slot passes from hand to
 * hand by an explicit per-slot turn counter (slot_acquire / slot_release).  A stream (pipe, tty,
 * socket) has one reader and one writer, and its slots are cut short when the producer stalls,
 * so a live source sees one-chunk latency like the reference; regular files are read and written
 * with positioned I/O by several threads, because one thread copies only ~5 GB/s out of the page
 * cache while the GPU path takes 44 GB/s.
 *
 * Environment (extensions; the option letters stay the reference's):
 *   FMD_DEVICES   comma list of GPU ordinals or "all" (default "0"): slots round-robin
 *   FMD_SLOT_MB   slot size in MiB (default 32), rounded to whole chunks
 *   FMD_SLOTS     ring depth per GPU (default 3)
 *   FMD_IO_THREADS reader / writer threads when the input / output is a regular file (default: a quarter
 *                 of the cores, 4..16)
 *   FMD_NO_MMAP   1 = write a regular output file with pwrite instead of through a shared mapping
 *   FMD_EXACT     1 = op-for-op audit kernel for every output
 *   FMD_STATS     1 = print throughput to stderr at exit
 */
#define _GNU_SOURCE
#include <errno.h>
#include <fcntl.h>
#include <poll.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <time.h>
#include <unistd.h>

#include "fmdemod.h"

#define MAX_DEV 16
#define MAX_IO_THREADS 16

typedef struct {
    uint8_t *in;   /* pinned */
    float *out;    /* pinned */
    size_t nbytes; /* filled by the reader */
    size_t nout;   /* floats this slot yields (set at submit time) */
    int last;      /* nothing follows this slot */
    size_t warm;   /* segmented -Q: bytes of warm-up prefix that sit right before the payload */
    unsigned long long first_sample; /* segmented -Q: global index of the payload's first sample */
    int dev;       /* which context owns it */
    unsigned ring; /* slot index inside that context's ring */
    /* Whose turn it is: slot number n (n % nslots == this slot, generation g = n / nslots) may be
     * filled when stage == 3g, submitted when stage == 3g + 1, drained when stage == 3g + 2; the
     * writer then sets 3(g + 1).  A thread waits for the stage of ITS slot number, so slot numbers
     * n and n + nslots can never be confused however the threads are scheduled (a binary
     * semaphore per phase would let the reader of n + nslots overtake the reader of n). */
    unsigned long long stage;
} slot_t;

typedef struct {
    int in_fd, out_fd;
    int in_regular, out_regular; /* regular files: positioned I/O, several threads */
    unsigned long long in_size;  /* bytes, when in_regular */
    size_t chunk, slot_bytes;
    size_t out_elem; /* bytes per output sample: 4, or 2 with -F s16 */
    unsigned nslots;
    slot_t *slots;
    fmd_ctx *ctx[MAX_DEV];
    int ndev;
    unsigned n_readers, n_writers;
    /* -Q over several GPUs: every slot is an independent segment that starts warm_cap bytes early
     * (include/fmdemod.h, fmd_submit_segment); the payload of a slot sits at in + warm_cap */
    int segmented;
    size_t warm_cap;
    uint8_t *hist;               /* stream input: the last warm_cap bytes read so far */
    unsigned long long stream_bytes;
    off_t in_base, out_base;     /* where the descriptors stood when we got them (inherited stdin / stdout) */
    /* file output through a shared mapping: writers memcpy into it instead of pwrite -- buffered
     * writes to ONE inode serialise on its lock (measured: 4 pwrite threads into a tmpfs file move
     * 3.1 GB/s together), page faults on a shared mapping do not */
    uint8_t *out_map;
    size_t out_map_len;          /* bytes mapped, from the page that holds out_base */
    size_t out_map_skew;         /* out_base minus the start of that page */
    volatile int error;
    unsigned long long bytes_in, floats_out;
    pthread_mutex_t stat_lock;
    /* one lock + one condition for all slot hand-offs: an event per 32 MiB slot, no contention */
    pthread_mutex_t turn_lock;
    pthread_cond_t turn_cond;
    int finished;                 /* the stream has ended ... */
    unsigned long long last_n;    /* ... with this slot number */
} host_t;

typedef struct { host_t *h; unsigned id; } io_arg_t;

static void die(const char *what) {
    fprintf(stderr, "demodulator: %s: %s\n", what, fmd_last_error());
    exit(2);
}

enum { PH_FREE = 0, PH_FILLED = 1, PH_INFLIGHT = 2 };

/* Block until slot number n reaches `phase`.  Returns 0 when the stream ended before slot number n
 * (the caller has nothing left to do), 1 when the slot is the caller's. */
static int slot_acquire(host_t *h, unsigned long long n, int phase) {
    slot_t *s = &h->slots[n % h->nslots];
    const unsigned long long want = 3ull * (n / h->nslots) + (unsigned) phase;
    int mine = 1;
    pthread_mutex_lock(&h->turn_lock);
    while (s->stage != want) {
        if (h->finished && n > h->last_n) {
            mine = 0;
            break;
        }
        pthread_cond_wait(&h->turn_cond, &h->turn_lock);
    }
    pthread_mutex_unlock(&h->turn_lock);
    return mine;
}

/* hand slot number n on to the next phase (PH_FREE = to the reader of slot number n + nslots) */
static void slot_release(host_t *h, unsigned long long n, int next_phase) {
    slot_t *s = &h->slots[n % h->nslots];
    pthread_mutex_lock(&h->turn_lock);
    s->stage = next_phase == PH_FREE ? 3ull * (n / h->nslots + 1ull) : 3ull * (n / h->nslots) + (unsigned) next_phase;
    pthread_cond_broadcast(&h->turn_cond);
    pthread_mutex_unlock(&h->turn_lock);
}

static void stream_finished(host_t *h, unsigned long long last_n) {
    pthread_mutex_lock(&h->turn_lock);
    h->finished = 1;
    h->last_n = last_n;
    pthread_cond_broadcast(&h->turn_cond);
    pthread_mutex_unlock(&h->turn_lock);
}

/* stream input: read until `want` bytes, EOF, or -- once at least one whole chunk is in -- the
 * producer stalls (a live source then sees one-chunk latency, like the reference) */
static size_t fill_slot_stream(host_t *h, uint8_t *dst, size_t want, int *eof) {
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

static int pread_all(host_t *h, uint8_t *dst, size_t want, unsigned long long off) {
    size_t got = 0;
    while (got < want) {
        ssize_t r = pread(h->in_fd, dst + got, want - got, h->in_base + (off_t) (off + got));
        if (r < 0 && errno == EINTR) continue;
        if (r <= 0) return -1;
        got += (size_t) r;
    }
    return 0;
}

/* file input: slot number n is bytes [n*slot_bytes, (n+1)*slot_bytes) of the file */
static size_t fill_slot_file(host_t *h, uint8_t *dst, unsigned long long n, int *eof) {
    const unsigned long long off = n * (unsigned long long) h->slot_bytes;
    size_t want = h->slot_bytes, got = 0;
    if (off >= h->in_size) want = 0;
    else if (h->in_size - off < want) want = (size_t) (h->in_size - off);
    *eof = off + want >= h->in_size;
    while (got < want) {
        ssize_t r = pread(h->in_fd, dst + got, want - got, h->in_base + (off_t) (off + got));
        if (r < 0) {
            if (errno == EINTR) continue;
            perror("demodulator: pread");
            h->error = -2;
            *eof = 1;
            break;
        }
        if (r == 0) { /* file shrank */
            *eof = 1;
            break;
        }
        got += (size_t) r;
    }
    return got;
}

/* reader `id` of n_readers fills slot numbers id, id + n_readers, ... (one reader for a stream) */
static void *reader_main(void *arg) {
    host_t *h = ((io_arg_t *) arg)->h;
    for (unsigned long long n = ((io_arg_t *) arg)->id;; n += h->n_readers) {
        /* file input: slot numbers past the end of the file are never produced (slot 0 always is) */
        if (h->in_regular && n > 0 && n * (unsigned long long) h->slot_bytes >= h->in_size) break;
        slot_t *s = &h->slots[n % h->nslots];
        if (!slot_acquire(h, n, PH_FREE)) break;
        int eof;
        uint8_t *payload = s->in + h->warm_cap;
        if (h->segmented) {
            /* the bytes that precede this slot in the stream, as far back as the warm-up reaches */
            const unsigned long long at = h->in_regular ? n * (unsigned long long) h->slot_bytes : h->stream_bytes;
            s->warm = at < h->warm_cap ? (size_t) at : h->warm_cap;
            s->first_sample = at / 2;
            if (s->warm && h->in_regular && pread_all(h, payload - s->warm, s->warm, at - s->warm)) {
                perror("demodulator: pread");
                h->error = -2;
            }
            if (s->warm && !h->in_regular) memcpy(payload - s->warm, h->hist + h->warm_cap - s->warm, s->warm);
        }
        s->nbytes = h->in_regular ? fill_slot_file(h, payload, n, &eof) : fill_slot_stream(h, payload, h->slot_bytes, &eof);
        if (h->segmented && !h->in_regular) { /* one reader: remember the end of the stream so far */
            if (s->nbytes >= h->warm_cap) {
                memcpy(h->hist, payload + s->nbytes - h->warm_cap, h->warm_cap);
            } else if (s->nbytes) {
                memmove(h->hist, h->hist + s->nbytes, h->warm_cap - s->nbytes);
                memcpy(h->hist + h->warm_cap - s->nbytes, payload, s->nbytes);
            }
            h->stream_bytes += s->nbytes;
        }
        s->last = eof;
        pthread_mutex_lock(&h->stat_lock);
        h->bytes_in += s->nbytes;
        pthread_mutex_unlock(&h->stat_lock);
        slot_release(h, n, PH_FILLED);
        if (eof) break;
    }
    return NULL;
}

static int write_all(int fd, const void *buf, size_t n, int positioned, off_t off) {
    const char *p = buf;
    while (n) {
        ssize_t w = positioned ? pwrite(fd, p, n, off) : write(fd, p, n);
        if (w < 0) {
            if (errno == EINTR) continue;
            return -1;
        }
        p += w;
        off += w;
        n -= (size_t) w;
    }
    return 0;
}

/* writer `id` of n_writers drains slot numbers id, id + n_writers, ... (one writer for a stream:
 * output order is the order of the writes) */
static void *writer_main(void *arg) {
    host_t *h = ((io_arg_t *) arg)->h;
    for (unsigned long long n = ((io_arg_t *) arg)->id;; n += h->n_writers) {
        slot_t *s = &h->slots[n % h->nslots];
        if (!slot_acquire(h, n, PH_INFLIGHT)) break; /* the stream ended before this slot number */
        if (s->nbytes) {
            if (fmd_wait(h->ctx[s->dev], s->ring)) die("fmd_wait");
            const size_t nf = s->nout;
            /* reference-compatible mode: a slot yields as many output bytes as it had input bytes,
             * so slot n lands at n*slot_bytes (positioned writes are off in quality mode) */
            const unsigned long long at = n * (unsigned long long) h->slot_bytes;
            if (h->out_map && at + nf * h->out_elem <= h->out_map_len - h->out_map_skew) {
                if (!h->error) memcpy(h->out_map + h->out_map_skew + at, s->out, nf * h->out_elem);
            } else if (!h->error && write_all(h->out_fd, s->out, nf * h->out_elem, h->out_regular,
                                              h->out_base + (off_t) at)) {
                perror("demodulator: write");
                h->error = -4;
            }
            pthread_mutex_lock(&h->stat_lock);
            h->floats_out += nf;
            pthread_mutex_unlock(&h->stat_lock);
        }
        const int last = s->last;
        slot_release(h, n, PH_FREE);
        if (last) break;
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

    while ((opt = getopt(argc, argv, "i:o:r:L:l:S:D:d:e:m:b:c:q:n:QF:")) != -1) {
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
            /* -Q (not a reference flag): the opt-in quality pipeline -- recursive filters with state
             * carried through the stream, adjacent-sample discriminator, decimation from the input
             * rate -r to the output rate -S.  Without -Q, -r is parsed and ignored like the
             * reference does (src/main.c:180). */
            case 'Q': cfg.flags |= FMD_FLAG_QUALITY; break;
            /* -F s16 (with -Q): signed 16-bit output, what the sox stage of the reference's pipelines
             * would make of the floats -- `... -Q -r2400000 -S48000 -F s16 | dsd -i - ...` */
            case 'F':
                if (!strcmp(optarg, "s16")) cfg.flags |= FMD_FLAG_S16;
                else if (strcmp(optarg, "f32")) { fprintf(stderr, "demodulator: -F takes f32 or s16\n"); ret += 1; }
                break;
            case 'r': cfg.input_rate = (double) strtold(optarg, NULL); break;
            default: break; /* -n: accepted and ignored, like the reference */
        }
    }
    cfg.mode &= 0xffu; /* consumerArgs.mode is a uint8_t */
    if (ret) return ret;
    if (h.in_fd < 0 || h.out_fd < 0) {
        fprintf(stderr, "demodulator: both -i and -o are required\n");
        return 1;
    }
    if ((e = getenv("FMD_EXACT")) && atoi(e)) cfg.flags |= FMD_FLAG_EXACT;
#ifdef F_SETPIPE_SZ
    /* a pipe moves 64 KiB per wake-up by default; ask for the largest buffer the kernel allows
     * (fails harmlessly on regular files and when the limit is lower) */
    for (long sz = 16L << 20; sz >= (1L << 20); sz >>= 1) {
        int ok_in = fcntl(h.in_fd, F_SETPIPE_SZ, sz) >= 0, ok_out = fcntl(h.out_fd, F_SETPIPE_SZ, sz) >= 0;
        if (ok_in || ok_out) break;
    }
#endif

    {
        struct stat sb;
        /* Regular files get positioned I/O from several threads.  A descriptor we inherited
         * (`-i -` / `-o -` with a redirection) may already stand past byte 0, and may be shared with
         * other writers: offsets are taken relative to where it stands now, and the descriptor is
         * moved past what was consumed / produced at exit, so `{ echo hdr; demodulator -o -; } > f`
         * behaves like the reference's sequential fread / fwrite.  With O_APPEND pwrite ignores
         * its offset (Linux), so such an output is written sequentially by one thread. */
        if (!fstat(h.in_fd, &sb) && S_ISREG(sb.st_mode)) {
            const off_t base = lseek(h.in_fd, 0, SEEK_CUR);
            if (base >= 0) {
                h.in_regular = 1;
                h.in_base = base;
                h.in_size = sb.st_size > base ? (unsigned long long) (sb.st_size - base) : 0ull;
            }
        }
        if (!fstat(h.out_fd, &sb) && S_ISREG(sb.st_mode)) {
            const off_t base = lseek(h.out_fd, 0, SEEK_CUR);
            const int fl = fcntl(h.out_fd, F_GETFL);
            if (base >= 0 && fl >= 0 && !(fl & O_APPEND)) {
                h.out_regular = 1;
                h.out_base = base;
            }
        }
        /* one thread copies ~4 GB/s between the page cache and a pinned slot: a quarter of the
         * cores (4..16) by default */
        long ncpu = sysconf(_SC_NPROCESSORS_ONLN);
        unsigned dflt = ncpu >= 64 ? 16u : (ncpu >= 16 ? (unsigned) ncpu / 4u : 4u);
        unsigned want = (e = getenv("FMD_IO_THREADS")) ? (unsigned) atoi(e) : dflt;
        if (want < 1) want = 1;
        if (want > MAX_IO_THREADS) want = MAX_IO_THREADS;
        h.n_readers = h.in_regular ? want : 1;
        /* positioned writes need every earlier slot to be full-sized: true when the input is a file
         * (only its last slot is short); a stream may cut slots short, so its output stays sequential */
        h.n_writers = (h.out_regular && h.in_regular) ? want : 1;
        if (!(h.out_regular && h.in_regular)) h.out_regular = 0;
        if ((cfg.flags & FMD_FLAG_QUALITY) || ((cfg.mode >> 2) & 3u) == 1u) {
            /* one ordered stream: -Q output sizes depend on the position; -q1 carries its offsets on */
            h.n_writers = 1;
            h.out_regular = 0;
        }
        pthread_mutex_init(&h.stat_lock, NULL);
        /* reference-compatible mode writes exactly as many bytes as it reads (4 B per 4 B): with a
         * regular input the size of the output is known, so it can be sized and mapped up front */
        if (h.out_regular && h.in_regular && h.n_writers > 1 && !((e = getenv("FMD_NO_MMAP")) && atoi(e))) {
            const unsigned long long total = h.in_size & ~3ull;
            const long pg = sysconf(_SC_PAGESIZE);
            const off_t map_from = h.out_base / pg * pg;
            if (total && !ftruncate(h.out_fd, h.out_base + (off_t) total)) {
                const size_t len = (size_t) (h.out_base - map_from) + (size_t) total;
                void *m = mmap(NULL, len, PROT_READ | PROT_WRITE, MAP_SHARED, h.out_fd, map_from);
                if (m != MAP_FAILED) {
                    h.out_map = m;
                    h.out_map_len = len;
                    h.out_map_skew = (size_t) (h.out_base - map_from);
                }
            }
        }
    }

    h.ndev = parse_devices(getenv("FMD_DEVICES"), devs);
    if (h.ndev <= 0) die("no CUDA device");
    if (((cfg.mode >> 2) & 3u) == 1u && h.ndev > 1) {
        fprintf(stderr, "demodulator: -q1 carries its offsets through the stream: using device %d only\n", devs[0]);
        h.ndev = 1;
    }
    for (int d = 0; d < h.ndev; ++d)
        if (fmd_create(&cfg, devs[d], &h.ctx[d])) die("fmd_create");

    h.chunk = (size_t) cfg.chunk_bytes;
    h.out_elem = (cfg.flags & FMD_FLAG_S16) ? sizeof(int16_t) : sizeof(float);
    if ((cfg.flags & FMD_FLAG_QUALITY) && h.ndev > 1) {
        /* recursive filters carry state: a slot that goes to another GPU becomes a segment with a
         * warm-up prefix (one GPU keeps the true carried state and needs none) */
        h.segmented = 1;
        h.warm_cap = (size_t) fmd_warmup_samples(h.ctx[0]) * 2u;
        if (!h.in_regular) h.hist = calloc(1, h.warm_cap ? h.warm_cap : 1);
    }
    {
        size_t mb = (e = getenv("FMD_SLOT_MB")) ? (size_t) atol(e) : 32;
        unsigned depth = (e = getenv("FMD_SLOTS")) ? (unsigned) atoi(e) : 3;
        if (mb < 1) mb = 1;
        if (depth < 2) depth = 2;
        if (depth > 16) depth = 16;
        h.slot_bytes = (mb << 20) / h.chunk * h.chunk;
        if (!h.slot_bytes) h.slot_bytes = h.chunk;
        while (h.segmented && h.slot_bytes < 4 * h.warm_cap) h.slot_bytes += h.chunk; /* keep the overlap under 25 % */
        /* enough slots for every I/O thread to own one while the GPUs own theirs */
        while (depth * (unsigned) h.ndev < h.n_readers + h.n_writers + 1u && depth < 16) ++depth;
        h.nslots = depth * (unsigned) h.ndev;
        h.slots = calloc(h.nslots, sizeof *h.slots);
        for (int d = 0; d < h.ndev; ++d)
            if (fmd_ring_create(h.ctx[d], depth, h.slot_bytes + (h.warm_cap + h.chunk - 1) / h.chunk * h.chunk)) die("fmd_ring_create");
        for (unsigned i = 0; i < h.nslots; ++i) {
            slot_t *s = &h.slots[i];
            s->dev = (int) (i % (unsigned) h.ndev);
            s->ring = i / (unsigned) h.ndev;
            /* the backend chooses how the reader's buffer is allocated, as in the reference's
             * allocateBuffer seam (src/main.c:65): pinned, so copies are true DMA */
            size_t out_bytes = fmd_max_out_count(h.ctx[s->dev], h.slot_bytes) * sizeof(float);
            if (out_bytes < h.slot_bytes) out_bytes = h.slot_bytes;
            if (fmd_host_alloc((void **) &s->in, h.warm_cap + h.slot_bytes) || fmd_host_alloc((void **) &s->out, out_bytes))
                die("fmd_host_alloc");
        }
    }
    pthread_mutex_init(&h.turn_lock, NULL);
    pthread_cond_init(&h.turn_cond, NULL);

    struct timespec t0, t1;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    pthread_t rd[MAX_IO_THREADS], wr[MAX_IO_THREADS];
    io_arg_t ra[MAX_IO_THREADS], wa[MAX_IO_THREADS];
    for (unsigned i = 0; i < h.n_readers; ++i) {
        ra[i].h = &h; ra[i].id = i;
        pthread_create(&rd[i], NULL, reader_main, &ra[i]);
    }
    for (unsigned i = 0; i < h.n_writers; ++i) {
        wa[i].h = &h; wa[i].id = i;
        pthread_create(&wr[i], NULL, writer_main, &wa[i]);
    }
    unsigned long long n = 0;
    for (;; ++n) {
        slot_t *s = &h.slots[n % h.nslots];
        slot_acquire(&h, n, PH_FILLED);
        /* tail policy: a final partial chunk of r bytes is a chunk of r&~3 bytes */
        s->nbytes &= ~(size_t) 3;
        if (((cfg.mode >> 2) & 3u) == 1u) s->nbytes -= s->nbytes % h.chunk; /* -q1 works on whole chunks only */
        if (h.segmented) {
            s->nout = fmd_segment_out_count(h.ctx[s->dev], s->first_sample, s->nbytes / 2);
            if (s->nbytes && fmd_submit_segment(h.ctx[s->dev], s->ring, s->in + h.warm_cap - s->warm, s->warm, s->nbytes,
                                                s->first_sample, s->out))
                die("fmd_submit_segment");
        } else {
            s->nout = fmd_out_count(h.ctx[s->dev], s->nbytes); /* before the submit moves a -Q stream on */
            if (s->nbytes && fmd_submit(h.ctx[s->dev], s->ring, s->in, s->nbytes, s->out)) die("fmd_submit");
        }
        const int last = s->last;
        slot_release(&h, n, PH_INFLIGHT);
        if (last) break;
    }
    /* readers and writers whose next slot number lies past the end of the stream stop waiting */
    stream_finished(&h, n);
    for (unsigned i = 0; i < h.n_readers; ++i) pthread_join(rd[i], NULL);
    for (unsigned i = 0; i < h.n_writers; ++i) pthread_join(wr[i], NULL);
    clock_gettime(CLOCK_MONOTONIC, &t1);

    if ((e = getenv("FMD_STATS")) && atoi(e)) {
        const double sec = (double) (t1.tv_sec - t0.tv_sec) + 1e-9 * (double) (t1.tv_nsec - t0.tv_nsec);
        fprintf(stderr, "demodulator: %llu bytes in, %llu floats out, %.3f s, %.1f MS/s, %d GPU(s)\n", h.bytes_in,
                h.floats_out, sec, (double) h.bytes_in / 2e6 / sec, h.ndev);
    }
    if (h.out_map) munmap(h.out_map, h.out_map_len);
    /* leave inherited descriptors where sequential I/O would have left them */
    if (h.in_regular) lseek(h.in_fd, h.in_base + (off_t) h.bytes_in, SEEK_SET);
    if (h.out_regular) lseek(h.out_fd, h.out_base + (off_t) (h.floats_out * h.out_elem), SEEK_SET);
    for (int d = 0; d < h.ndev; ++d) fmd_destroy(h.ctx[d]);
    for (unsigned i = 0; i < h.nslots; ++i) {
        fmd_host_free(h.slots[i].in);
        fmd_host_free(h.slots[i].out);
    }
    free(h.hist);
    free(h.slots);
    close(h.out_fd);
    close(h.in_fd);
    return h.error ? 1 : ret;
}
