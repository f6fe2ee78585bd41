/*
 * fmdemod.h -- C ABI of the B200-native FM demodulation backend.
 *
 * Drop-in boundary for the hot path of peads/demodulator: everything the
 * reference's worker thread does per chunk between "bytes arrived" and
 * "floats written" (src/matrix.c:236-273), plus the set-up it does once
 * (src/matrix.c:181-234).  Plain C: pointers, sizes, ints.  No CUDA, C++ or
 * torch types appear in any signature; `stream` arguments are an opaque
 * cudaStream_t passed as void* (NULL = CUDA's default stream).
 *
 * Every entry point returns 0 on success or a negative FMD_E* code and leaves
 * a human-readable message for fmd_last_error() (thread-local).  There is no
 * CPU fallback: without a CUDA device fmd_create fails with FMD_ENODEV.
 *
 * Units: a "sample" is one complex IQ pair = 2 input bytes (u8 I, u8 Q).
 * A chunk of B bytes yields exactly B/4 float32 outputs (out[k] is computed
 * from input samples 2k and 2k+1 -- src/matrix.c:166-175); chunks are
 * processed independently of one another (src/matrix.c:272 clears all filter
 * memory between chunks).
 */
#ifndef FMDEMOD_H
#define FMDEMOD_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FMD_DEFAULT_CHUNK_BYTES 262144u /* DEFAULT_BUF_SIZE, include/matrix.h:37-39 */
#define FMD_MAX_SECTIONS 8              /* ceil(degree/2) <= 8, i.e. -d/-D up to 16 */
#define FMD_MAX_TAPS 6                  /* folded stencil width for FMD_MAX_SECTIONS */

enum {
    FMD_OK = 0,
    FMD_EINVAL = -1, /* bad argument / configuration the reference leaves undefined */
    FMD_ENODEV = -2, /* no usable CUDA device */
    FMD_ECUDA = -3,  /* CUDA runtime error (message has the cudaError string) */
    FMD_ENOMEM = -4,
    FMD_EBUSY = -5   /* slot still in flight */
};

/* filter family, numbering of processFilterOption's `mode` (src/matrix.c:48-73) */
enum { FMD_BUTTER_LP = 0, FMD_CHEBY1_LP = 1, FMD_BUTTER_HP = 2, FMD_CHEBY1_HP = 3 };

/* fmd_config.flags */
enum {
    FMD_FLAG_EXACT = 1u,  /* run every output through the op-for-op kernel (slow; for audits) */
    /*
     * Opt-in "quality" pipeline -- NOT what the reference computes, never the
     * default.  The same designed sections (src/matrix.c:25-80, designed at `input_rate`) run as
     * true recursive filters whose state is carried through the whole stream, the discriminator
     * works on ADJACENT samples (x[n] = arg z[n] conj z[n-1]), and the output is decimated by
     * M = input_rate / sample_rate inside the output filter's kernel: float m of the stream is
     * the filtered discriminator at input sample m*M.  Samples are centred by -127.5.  The
     * context is a STREAM: successive fmd_demod_* / fmd_submit calls continue it (issue them in
     * order on one stream), fmd_reset starts a new one.  Needs input_rate / sample_rate to be an
     * integer >= 1, degrees <= 6, -q0, and cut-offs of at least input_rate / 1000 (float32 state).
     */
    FMD_FLAG_QUALITY = 2u,
    /*
     * Egress adapter of the quality pipeline (with FMD_FLAG_QUALITY only): the decimated outputs
     * leave as signed 16-bit samples -- full scale 1.0 = 32768, round to nearest, clip -- i.e. what
     * the `sox -e signed-integer -b16` stage of the reference's pipelines (README.md:22,29) does to
     * the float stream.  Every `float *out` argument then points at int16_t storage (cast it);
     * fmd_out_count still counts samples.
     */
    FMD_FLAG_S16 = 4u
};

/*
 * Mirrors the filter/geometry fields of the reference's consumerArgs
 * (include/matrix.h:43-57) with the same meaning and defaults
 * (src/main.c:102-113).  Doubles instead of long doubles: every value the
 * CLI can produce (strtold of a decimal option) that matters is exact.
 */
typedef struct fmd_config {
    double sample_rate;   /* -S  consumerArgs.sampleRate; filter DESIGN rate only (default 125000) */
    double lowpass_in;    /* -L  consumerArgs.lowpassIn;  0 = no input filter                     */
    double lowpass_out;   /* -l  consumerArgs.lowpassOut (default 12500)                          */
    double epsilon;       /* -e/10 consumerArgs.epsilon (default 0.3), Chebyshev ripple           */
    uint32_t in_degree;   /* -D  consumerArgs.inFilterDegree; 0 with -L => out_degree             */
    uint32_t out_degree;  /* -d  consumerArgs.outFilterDegree (default 3)                         */
    uint32_t mode;        /* consumerArgs.mode: bit0 out Chebyshev, bit1 in Chebyshev,
                             bits2-3 input conditioner (-q: 0 shiftOrigin, 1 correctIq, 2 highpassDc,
                             3 normalizeInput).  -q1 makes the context a STREAM like the reference's
                             static offsets do: calls take whole chunks and continue it, fmd_reset
                             starts over.  bits4-5 demod (-c) */
    uint32_t flags;       /* FMD_FLAG_*                                                           */
    uint64_t chunk_bytes; /* consumerArgs.bufSize (default 262144); multiple of 4                 */
    double input_rate;    /* -r  rate of the IQ input; read only under FMD_FLAG_QUALITY (the
                             reference parses and ignores -r, src/main.c:180)                     */
} fmd_config;

typedef struct fmd_ctx fmd_ctx;

/* fill *cfg with the reference CLI's defaults (src/main.c:102-113) */
void fmd_config_default(fmd_config *cfg);

const char *fmd_last_error(void);

/* number of CUDA devices visible, or a negative FMD_E* */
int fmd_device_count(void);

/* ---- set-up: replaces processFilterOption + transformBilinear + zp2Sos
 *      (src/matrix.c:25-80, src/filter.c:22-210).  Host only, long double inside. ---- */

/* writes ceil(degree/2) rows {b0,b1,b2,a0,a1,a2} as float32 */
int fmd_design_sos(unsigned kind, unsigned degree, double fc, double fs, double eps, float *sos);

/*
 * Fold `sections` rows into the steady-state stencil the section sweep of
 * applyFilter/applyComplexFilter (src/filter.c:221-230, 242-258) is equivalent
 * to:  y[k] = *dc + sum_{t<*ntaps} taps[t] * x[k+t].   `bias` = 1 for lanes that
 * carry the literal "+1" (real filter, I lane), 0 for the Q lane.
 */
int fmd_fold_taps(const float *sos, unsigned sections, int bias,
                  float taps[FMD_MAX_TAPS], float *dc, unsigned *ntaps);

/* ---- context: replaces the worker's set-up block (src/matrix.c:181-234) ---- */

int fmd_create(const fmd_config *cfg, int device, fmd_ctx **out);
void fmd_destroy(fmd_ctx *ctx);

/* effective configuration after the reference's defaulting rules (in_degree inheritance) */
int fmd_get_config(const fmd_ctx *ctx, fmd_config *cfg);
/* designed coefficient rows; `which` 0 = output filter, 1 = input filter. returns rows written */
int fmd_get_sos(const fmd_ctx *ctx, int which, float *sos /* [FMD_MAX_SECTIONS][6] */);

/* ---- pinned host memory: replaces allocateBuffer (src/matrix.c:282-285,
 *      cuda/matrix.cu:106-109) ---- */
int fmd_host_alloc(void **ptr, size_t bytes);
int fmd_host_free(void *ptr);
/* pin / unpin caller-owned memory in place (for buffers the caller cannot reallocate) */
int fmd_host_register(void *ptr, size_t bytes);
int fmd_host_unregister(void *ptr);

/* floats produced for `nbytes` of input under the tail policy below.  Under FMD_FLAG_QUALITY:
 * for the NEXT `nbytes` of the stream (depends on the stream position modulo M). */
size_t fmd_out_count(const fmd_ctx *ctx, size_t nbytes);
/* upper bound of fmd_out_count over every stream position (sizes output buffers) */
size_t fmd_max_out_count(const fmd_ctx *ctx, size_t nbytes);
/* FMD_FLAG_QUALITY / -q1: forget the carried state and stream position (waits for the device).
 * No effect otherwise: the reference keeps no state between chunks (src/matrix.c:272). */
int fmd_reset(fmd_ctx *ctx);

/*
 * ---- the hot path: replaces one or many iterations of the worker loop
 *      (src/matrix.c:236-273: shiftOrigin -> [applyComplexFilter] -> fmDemod ->
 *      applyFilter), for `nbytes` of input starting ON A CHUNK BOUNDARY. ----
 *
 * Whole chunks behave exactly as the reference.  A final partial chunk of r
 * bytes is processed as a chunk of length r&~3 with the same edge rules and
 * yields r/4 floats (the reference's own behaviour there is a data race, see
 * DESIGN.md "tail policy").
 */

/* device-resident buffers; asynchronous on `stream` */
int fmd_demod_device(fmd_ctx *ctx, const uint8_t *d_in, size_t nbytes, float *d_out, void *stream);
/* Synchronise `stream` and report what only the device can know: in the quality pipeline a tile
 * whose look-back gave up waiting raises a flag (its outputs are then invalid).  fmd_wait does this
 * for ring slots; callers of fmd_demod_device use this.  The flag is cleared once reported. */
int fmd_check(fmd_ctx *ctx, void *stream);

/* host buffers; copies in, computes, copies out, returns when `out` is complete.
 * Pageable memory works; pinned memory (fmd_host_alloc) gets full copy/compute overlap. */
int fmd_demod_host(fmd_ctx *ctx, const uint8_t *in, size_t nbytes, float *out, size_t *nout);

/*
 * Asynchronous ring for streaming hosts (the CLI): `slots` staging slots of
 * `slot_bytes` input bytes each are allocated on the device at ring creation.
 * submit: H2D on the copy-in stream -> kernels on the compute stream -> D2H on
 * the copy-out stream, chained by events; returns at once.  `in`/`out` must be
 * pinned and stay valid until fmd_wait(slot) returns.
 */
int fmd_ring_create(fmd_ctx *ctx, unsigned slots, size_t slot_bytes);
int fmd_submit(fmd_ctx *ctx, unsigned slot, const uint8_t *in, size_t nbytes, float *out);
int fmd_wait(fmd_ctx *ctx, unsigned slot);

/*
 * ---- one quality-pipeline stream over several GPUs (BASELINE north star: "a single long stream is
 *      split into overlapped chunks with warm-up state handoff").  The reference has nothing to
 *      replace here: its sweep keeps no state (src/matrix.c:272), which is why the
 *      reference-compatible path needs no overlap at all.  With FMD_FLAG_QUALITY the recursive
 *      filters do carry state, so a segment handed to another GPU starts `fmd_warmup_samples`
 *      samples early from a zero state and drops the outputs of that prefix: the state has then
 *      converged to the running stream's to ~1e-9 (the length follows from the slowest pole). ----
 */
uint64_t fmd_warmup_samples(const fmd_ctx *ctx);
/* the same for one cascade given its section rows {b0,b1,b2,a0,a1,a2} (no context, no GPU needed):
 * fmd_warmup_samples = the output cascade's + the input cascade's, rounded up to 8 samples */
uint64_t fmd_warmup_for_sections(const float *sos6, unsigned rows);
/* outputs of `nsamples` samples whose first has global index `first_sample` (decimation keeps the
 * samples whose GLOBAL index is a multiple of the factor) */
size_t fmd_segment_out_count(const fmd_ctx *ctx, uint64_t first_sample, size_t nsamples);
/* like fmd_submit, for a segment: `in` holds warm_bytes of the stream that precede the payload (a
 * multiple of 16; 0 at the stream start) and then nbytes of payload whose first sample has global
 * index first_sample; `out` receives fmd_segment_out_count(first_sample, nbytes / 2) outputs */
int fmd_submit_segment(fmd_ctx *ctx, unsigned slot, const uint8_t *in, size_t warm_bytes, size_t nbytes,
                       uint64_t first_sample, float *out);

/* ---- introspection for benches/tests ---- */

/* kernels launched by this context since creation (fast + exact) */
uint64_t fmd_launch_count(const fmd_ctx *ctx);
/* per-launch timing of the dominant (fast) kernel: while enabled, every fast-kernel launch is
 * bracketed by CUDA events on its own stream.  fmd_kernel_time synchronises, returns the
 * accumulated milliseconds and launch count since the last call, and resets both. */
int fmd_profile_kernels(fmd_ctx *ctx, int enable);
int fmd_kernel_time(fmd_ctx *ctx, double *ms, uint64_t *launches);
/* The copies of fmd_demod_host without the kernels between them (same slices, same streams, same
 * event chain; `out` receives whatever the device buffers hold): what the host side of the box
 * can move for this call pattern, i.e. the ceiling of the end-to-end number. */
int fmd_copy_host(fmd_ctx *ctx, const uint8_t *in, size_t nbytes, float *out);
/* time the device leg only: `reps` back-to-back fmd_demod_device calls bracketed by CUDA
 * events on the launching stream; returns total milliseconds in *ms */
int fmd_time_device(fmd_ctx *ctx, const uint8_t *d_in, size_t nbytes, float *d_out,
                    void *stream, int reps, float *ms);

#ifdef __cplusplus
}
#endif
#endif /* FMDEMOD_H */
