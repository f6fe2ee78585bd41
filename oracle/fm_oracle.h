/*
 * oracle/fm_oracle.h -- CPU restatement of peads/demodulator's FM path (every -q conditioner).
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product (demodulator_b200/, the
 * C-ABI library, the CLI) may include, link or call this.  Callers allowed:
 * tests/, __graft_entry__.smoke(), bench.py's cpu_baseline / --impl reference.
 *
 * Parity status: PINNED.  The reference ships no golden vectors (SURVEY.md
 * section 4), so this port is pinned against the reference itself: the
 * unmodified sources compiled in place (oracle/_ref/libdemod_ref.so, see
 * ref_shim.c) -- bit-exact on every stage and on the whole chunk pipeline
 * (tests/test_oracle_vs_ref.py), and against the fixtures under tests/golden/
 * that were generated from that build and from the reference CLI binary
 * (tests/golden/make_golden.py).
 */
#ifndef FM_ORACLE_H
#define FM_ORACLE_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* filter family selector, same numbering as the `mode` argument of the
 * reference's processFilterOption (src/matrix.c:48-73) */
enum { FMO_BUTTER_LP = 0, FMO_CHEBY1_LP = 1, FMO_BUTTER_HP = 2, FMO_CHEBY1_HP = 3 };

typedef struct {
    double fs;        /* -S  (filter design rate only) */
    double fc_in;     /* -L  (0 = no input filter)     */
    double fc_out;    /* -l                            */
    double eps;       /* -e / 10                       */
    unsigned deg_in;  /* -D  (0 = inherit deg_out when fc_in != 0) */
    unsigned deg_out; /* -d                            */
    unsigned mode;    /* bit0: output filter Chebyshev, bit1: input filter Chebyshev,
                         bits 2-3: input conditioner (-q): 0 shiftOrigin, 1 correctIq (state
                         crosses chunks: one fmo_process_chunks call = one process run),
                         2 highpassDc, 3 normalizeInput */
} fmo_config;

size_t fmo_sections(unsigned degree);

/* design `degree`-th order filter; writes ceil(degree/2) rows {b0,b1,b2,a0,a1,a2} */
void fmo_design_sos(unsigned kind, unsigned degree, double fc, double fs, double eps, float *sos);

/* u8 -> float, value-127 wrapped to int8 */
void fmo_shift(const uint8_t *in, size_t n, float *out);
/* u8 -> float in [-1, 1] (-q3) */
void fmo_normalize(const uint8_t *in, size_t n, float *out);

/* u8 -> float minus a running offset per lane (-q1); `off` carries on from chunk to chunk */
void fmo_correct_iq(const uint8_t *in, size_t n, float esr, float off[2], float *out);

/* n floats (n/2 complex, n/4 pairs) -> n/4 phase differences */
void fmo_discriminate(const float *iq, size_t n, float *out);

/*
 * Section sweep over one real lane.
 *   x    : n + 2*L readable floats (caller zero-extends)
 *   y    : n + 2*L floats; on entry the initial contents of the accumulator
 *          region, on exit the result (indices L .. n+2L-2 may be rewritten)
 *   bias : 1 for the lane that carries the literal "+1", 0 for the lane that
 *          does not (the Q lane of the complex filter)
 */
void fmo_sweep_lane(const float *x, float *y, size_t n, size_t L, const float *sos, int bias);

/* whole chunks of `chunk_bytes` bytes -> chunk_bytes/4 floats each. returns floats written */
size_t fmo_process_chunks(const uint8_t *in, size_t nchunks, size_t chunk_bytes,
                          const fmo_config *cfg, float *out);

/*
 * ---- opt-in "quality" pipeline (fm_quality_oracle.c): recursive filters with state carried
 * through the stream, adjacent-sample discriminator, decimation by M = input_rate / cfg->fs.
 * Not a restatement of the reference (which computes none of this): pinned against
 * scipy.signal.sosfilt instead, see the header of fm_quality_oracle.c. ----
 */
size_t fmo_quality_state_size(void);
/* returns M, or 0 if input_rate / cfg->fs is not an integer >= 1 or a degree is out of range */
unsigned fmo_quality_init(void *state, const fmo_config *cfg, double input_rate);
/* the next nsamples complex samples of the stream -> floats written (global index % M == 0) */
size_t fmo_quality_run(void *state, const uint8_t *in, size_t nsamples, float *out);
/* undecimated intermediate signals: z[2n], z[2n+1], x[n], v[n] */
void fmo_quality_stages(void *state, const uint8_t *in, size_t nsamples, double *z, double *x, double *v);

#ifdef __cplusplus
}
#endif
#endif
