/*
 * oracle/fm_quality_oracle.c -- CPU statement of the opt-in "quality" pipeline
 * (what the north star describes and the reference
 * does NOT compute: recursive filters with state carried through the stream,
 * adjacent-sample discriminator, decimation to the -S rate).
 *
 * TEST INFRASTRUCTURE ONLY -- same rules as fm_oracle.h.
 *
 * Parity status: NOT checkable against the reference (its applyFilter /
 * applyComplexFilter are not recursive in time and it never decimates beyond
 * the pair discriminator -- SURVEY.md F2-F4).  What this file is pinned
 * against instead: scipy.signal.sosfilt (double precision, direct form II
 * transposed) on the coefficient rows the reference's own design code
 * produces -- fixtures under tests/golden/quality_*.npz, generator
 * tests/golden/make_quality_golden.py, test tests/test_quality_oracle.py.
 *
 * The arithmetic here is double precision, serial, sample by sample:
 *
 *   s[n]  = (I[n] - 127.5, Q[n] - 127.5)                     u8 -> centred
 *   z[n]  = cascade_in(s[n])      per lane, only with fc_in   causal IIR (rows of
 *                                                             processFilterOption,
 *                                                             src/matrix.c:25-80, designed
 *                                                             at the INPUT rate)
 *   x[n]  = atan2(Im, Re)(z[n] * conj(z[n-1])),  z[-1] = 0    adjacent samples
 *   v[n]  = cascade_out(x[n])                                 causal IIR
 *   y[m]  = v[m*M],  M = input_rate / fs                      keep global n % M == 0
 *
 * A cascade section {b0,b1,b2,1,a1,a2} is run in direct form II transposed:
 *   w = b0*u + s1;  s1 = b1*u - a1*w + s2;  s2 = b2*u - a2*w.
 */
#include "fm_oracle.h"

#include <math.h>
#include <string.h>

#define FMO_Q_MAXSEC 8

struct fmo_qstate {
    double in_s[2][FMO_Q_MAXSEC][2]; /* lane, section, (s1, s2) */
    double out_s[FMO_Q_MAXSEC][2];
    double zprev[2];
    unsigned long long pos; /* global index of the next input sample */
    float sos_in[FMO_Q_MAXSEC][6], sos_out[FMO_Q_MAXSEC][6];
    unsigned lin, lout, decim;
};

size_t fmo_quality_state_size(void) { return sizeof(struct fmo_qstate); }

/* returns the decimation factor, 0 if the configuration is not usable */
unsigned fmo_quality_init(void *state, const fmo_config *cfg, double input_rate) {
    struct fmo_qstate *st = state;
    memset(st, 0, sizeof *st);
    if (!(input_rate > 0.) || !(cfg->fs > 0.)) return 0;
    const double ratio = input_rate / cfg->fs;
    const double m = floor(ratio + 0.5);
    if (m < 1. || m > 65536. || fabs(ratio - m) > 1e-9 * ratio) return 0;
    unsigned deg_in = cfg->deg_in;
    if (cfg->fc_in != 0. && deg_in == 0) deg_in = cfg->deg_out; /* same inheritance as matrix.c:190-192 */
    if (cfg->deg_out == 0 || fmo_sections(cfg->deg_out) > FMO_Q_MAXSEC || fmo_sections(deg_in) > FMO_Q_MAXSEC) return 0;
    st->lout = (unsigned) fmo_sections(cfg->deg_out);
    fmo_design_sos(cfg->mode & 1u, cfg->deg_out, cfg->fc_out, input_rate, cfg->eps, &st->sos_out[0][0]);
    if (cfg->fc_in != 0.) {
        st->lin = (unsigned) fmo_sections(deg_in);
        fmo_design_sos((cfg->mode >> 1) & 1u, deg_in, cfg->fc_in, input_rate, cfg->eps, &st->sos_in[0][0]);
    }
    st->decim = (unsigned) m;
    return st->decim;
}

static double cascade_step(const float (*sos)[6], unsigned L, double (*s)[2], double u) {
    for (unsigned i = 0; i < L; ++i) {
        const double b0 = sos[i][0], b1 = sos[i][1], b2 = sos[i][2], a1 = sos[i][4], a2 = sos[i][5];
        const double w = b0 * u + s[i][0];
        s[i][0] = b1 * u - a1 * w + s[i][1];
        s[i][1] = b2 * u - a2 * w;
        u = w;
    }
    return u;
}

/* the next `nsamples` complex samples of the stream; returns floats written */
size_t fmo_quality_run(void *state, const uint8_t *in, size_t nsamples, float *out) {
    struct fmo_qstate *st = state;
    size_t nout = 0;
    for (size_t n = 0; n < nsamples; ++n) {
        double zr = (double) in[2 * n] - 127.5, zi = (double) in[2 * n + 1] - 127.5;
        if (st->lin) {
            zr = cascade_step(st->sos_in, st->lin, st->in_s[0], zr);
            zi = cascade_step(st->sos_in, st->lin, st->in_s[1], zi);
        }
        const double re = zr * st->zprev[0] + zi * st->zprev[1];
        const double im = zi * st->zprev[0] - zr * st->zprev[1];
        st->zprev[0] = zr;
        st->zprev[1] = zi;
        const double v = cascade_step(st->sos_out, st->lout, st->out_s, atan2(im, re));
        if (st->pos % st->decim == 0) out[nout++] = (float) v;
        st->pos++;
    }
    return nout;
}

/* intermediate signals for stage-level fixtures: z (2 doubles per sample), x, v (1 each), no decimation */
void fmo_quality_stages(void *state, const uint8_t *in, size_t nsamples, double *z, double *x, double *v) {
    struct fmo_qstate *st = state;
    for (size_t n = 0; n < nsamples; ++n) {
        double zr = (double) in[2 * n] - 127.5, zi = (double) in[2 * n + 1] - 127.5;
        if (st->lin) {
            zr = cascade_step(st->sos_in, st->lin, st->in_s[0], zr);
            zi = cascade_step(st->sos_in, st->lin, st->in_s[1], zi);
        }
        const double re = zr * st->zprev[0] + zi * st->zprev[1];
        const double im = zi * st->zprev[0] - zr * st->zprev[1];
        st->zprev[0] = zr;
        st->zprev[1] = zi;
        z[2 * n] = zr;
        z[2 * n + 1] = zi;
        x[n] = atan2(im, re);
        v[n] = cascade_step(st->sos_out, st->lout, st->out_s, x[n]);
        st->pos++;
    }
}
