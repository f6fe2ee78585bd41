/*
 * oracle/fm_oracle.c -- plain-C restatement of the reference's default FM
 * demodulation path, written stage-major (whole-array passes) instead of the
 * reference's sample-major in-place loops.  TEST INFRASTRUCTURE ONLY; see the
 * header for who may call this and for the parity-pinning statement.
 *
 * Every function names the reference lines it restates (paths relative to
 * /root/reference).  Build with -ffp-contract=off: the reference's generic
 * x86-64 build rounds every float multiply and add separately.
 */
#include "fm_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

typedef long double ld;

size_t fmo_sections(unsigned degree) { return (degree >> 1) + (degree & 1u); }

/* ------------------------------------------------------------------ design */

typedef struct { ld re, im, ret; } fmo_root; /* pole (re,im) and the generator's return value */

/* src/filter.c:22-40 (warpButterGeneric), :42-58 (Hp / Lp wrappers) */
static fmo_root butter_root(ld alpha, ld beta, unsigned k, unsigned n, int highpass) {
    fmo_root r;
    const ld ang = M_PI_2 * (1. / (ld) n * (-1. + (ld) (k << 1)) + 1.);
    const ld c = cosl(ang);
    const ld inv = 1. / (c - alpha);
    const ld zr = (-beta + c) * inv;
    r.im = sinl(ang) * inv;
    r.re = -zr + 1.;
    r.ret = highpass ? 2. - zr : zr;
    return r;
}

/* src/filter.c:60-83 (warpCheby1Generic), :85-101 (wrappers; Hp inverts the tangent) */
static fmo_root cheby1_root(ld tng, ld ep, unsigned k, unsigned n, int highpass) {
    fmo_root r;
    const ld t0 = highpass ? 1. / tng : tng;
    const ld inv_n = 1. / (ld) n;
    const ld v = logl((1. + powl(10., 0.5 * ep)) / sqrtl(-1. + powl(10., ep))) * inv_n;
    const ld t = M_PI_2 * (inv_n * (-1. + (ld) (k << 1)));
    const ld a = cosl(t) * coshl(v) * t0;
    const ld b = sinl(t) * sinhl(v) * t0;
    const ld c = a * a + b * b;
    const ld d = 1. / (1. + c + 2. * b);
    const ld zj = 2. * a * d;
    const ld zr = 2. * (b + c) * d;
    r.re = -zr + 1.;
    r.im = zj;
    r.ret = zr;
    return r;
}

/*
 * src/matrix.c:25-80 (processFilterOption) + src/filter.c:142-210
 * (transformBilinear) + src/filter.c:104-140 (zp2Sos), folded into one routine:
 * roots -> running gain -> rows {1,-2Re z,|z|^2, 1,-2Re p,|p|^2}, odd order adds a
 * first-order row, the gain scales the b's of the LAST conjugate-pair row
 * (index npc-1 for even order, npc for odd order).
 */
void fmo_design_sos(unsigned kind, unsigned degree, double fc_d, double fs_d, double eps_d, float *sosf) {
    const unsigned n = degree;
    const size_t rows = fmo_sections(n);
    const unsigned npc = n >> 1;
    const int odd = (int) (n & 1u);
    const int highpass = (kind == FMO_BUTTER_HP || kind == FMO_CHEBY1_HP);
    const int cheby = (kind == FMO_CHEBY1_LP || kind == FMO_CHEBY1_HP);
    const ld fc = (ld) fc_d, fs = (ld) fs_d, eps = (ld) eps_d;
    const ld w = M_PI * fc / fs;                                     /* matrix.c:34 */
    ld alpha, beta;
    ld g_re, g_im = 0.;
    ld zero = highpass ? 1. : -1.;                                   /* filter.c:159 */
    ld (*sos)[6];
    fmo_root *roots;
    size_t i, j;
    unsigned k;

    if (!rows) return;
    if (cheby) {                                                     /* matrix.c:37,53,65 */
        const ld wh = coshl(1. / (ld) n * acoshl(1. / sqrtl(powl(10., eps) - 1.)));
        alpha = tanl(w * wh);
        beta = eps;
    } else {                                                         /* matrix.c:59,71 */
        alpha = 1. / sinl(2. * w);
        beta = tanl(w);
    }
    /* filter.c:150,153 -- only the LOWPASS Chebyshev generator gets the even-order 1/sqrt2 */
    g_re = (!odd && kind == FMO_CHEBY1_LP) ? M_SQRT1_2 : 1.;

    sos = calloc(rows, sizeof *sos);
    roots = calloc(rows, sizeof *roots);
    for (k = 1; k <= rows; ++k) {                                    /* filter.c:165-182 */
        fmo_root r = cheby ? cheby1_root(alpha, beta, k, n, highpass)
                           : butter_root(alpha, beta, k, n, highpass);
        const ld zr = r.ret, zj = r.im;
        roots[k - 1] = r;
        if (k <= npc) {
            const ld m2 = zr * zr + zj * zj;
            g_re *= m2;
            g_im *= m2;
        } else {
            const ld t = zr * g_re - zj * g_im;
            g_im = zr * g_im + zj * g_re;
            g_re = t;
        }
    }
    g_re /= (ld) (1 << n);                                           /* filter.c:184 */

    for (j = 0; j < npc; ++j) {                                      /* filter.c:117-123 */
        sos[j][0] = sos[j][3] = 1.;
        sos[j][1] = -2. * zero;
        sos[j][2] = zero * zero + 0. * 0.;
        sos[j][4] = -2. * roots[j].re;
        sos[j][5] = roots[j].re * roots[j].re + roots[j].im * roots[j].im;
    }
    if (odd) {                                                       /* filter.c:125-130 */
        sos[npc][0] = 1.;
        sos[npc][1] = -zero;
        sos[npc][2] = sos[npc][5] = 0.;
        sos[npc][3] = 1.;
        sos[npc][4] = -roots[npc].re;
    } else {                                                         /* filter.c:131-135 */
        sos[0][0] = 1.;
        sos[0][2] = 1.;
        sos[0][1] = -2. * zero;
    }
    {
        const size_t last = odd ? npc : npc - 1;                     /* filter.c:113,137-139 */
        sos[last][0] *= g_re;
        sos[last][1] *= g_re;
        sos[last][2] *= g_re;
    }
    for (i = 0; i < rows; ++i)
        for (j = 0; j < 6; ++j) sosf[i * 6 + j] = (float) sos[i][j]; /* matrix.c:75-79 */
    free(sos);
    free(roots);
}

/* ------------------------------------------------------------- sample ops */

/* src/matrix.c:82-98: out = (float)(int8_t)(u8 - 127); 255 -> -128 */
void fmo_shift(const uint8_t *in, size_t n, float *out) {
    size_t i;
    for (i = 0; i < n; ++i) {
        int v = (int) in[i] - 127;
        if (v > 127) v -= 256;
        out[i] = (float) v;
    }
}

/* src/matrix.c:100-118 (-q3): out = (float)((double)((float)u * (float)(2/255)) - 1.0) */
void fmo_normalize(const uint8_t *in, size_t n, float *out) {
    const float denom = 2. / 255.;
    size_t i;
    for (i = 0; i < n; ++i) {
        const float p = (float) in[i] * denom;
        out[i] = (float) ((double) p - 1.);
    }
}

/* src/matrix.c:159-176: s0*conj(s1) over NON-overlapping pairs, atan2f, NaN->0 */
/*
 * src/matrix.c:120-140 (-q1, correctIq): walks the chunk from both ends; sample pair k (complex k
 * from the front, complex nc-1-k from the back) is reduced by the running offsets off[0] (I) and
 * off[1] (Q), which then move towards the mean of what they were subtracted from:
 *     off += ((u_front - off) + (u_back - off)) * esr,        esr = (float)(50 / -S)
 * `off` is a function-local static in the reference: it lives on from chunk to chunk.
 */
void fmo_correct_iq(const uint8_t *in, size_t n, float esr, float off[2], float *out) {
    const size_t half = n >> 1;
    size_t i;
    for (i = 0; i < half; i += 2) {
        out[i] = (float) in[i] - off[0];
        out[n - i - 2] = (float) in[n - i - 2] - off[0];
        out[i + 1] = (float) in[i + 1] - off[1];
        out[n - i - 1] = (float) in[n - i - 1] - off[1];
        {
            const float s0 = out[i] + out[n - i - 2], s1 = out[i + 1] + out[n - i - 1];
            const float d0 = s0 * esr, d1 = s1 * esr;
            off[0] = off[0] + d0;
            off[1] = off[1] + d1;
        }
    }
}

void fmo_discriminate(const float *iq, size_t n, float *out) {
    size_t p;
    for (p = 0; p + 3 < n; p += 4) {
        const float a = iq[p], b = iq[p + 1], c = iq[p + 2], d = iq[p + 3];
        const float ac = a * c, bd = b * d;
        const float nad = -a * d, bc = b * c;
        const float re = ac + bd;
        const float im = nad + bc;
        const float ph = atan2f(im, re);
        out[p >> 2] = isnan(ph) ? 0.f : ph;
    }
}

/* ----------------------------------------------------------- section sweep */

/*
 * src/filter.c:221-230 (real) and :242-258 (one lane of the complex filter).
 * The reference updates y in place, sample-major: at outer step i, section m
 * rewrites y[i+L+m] from y[i+L+m], y[i+L+m+1], x[i+L+m], x[i+L+m+1].  Index k
 * is therefore rewritten by section s at step i=k-L-s (if 0<=i<n), sections in
 * DEcreasing s, and what it reads as y[k+1] has by then seen sections >= s+2
 * only.  Stage-major that is  Y_s[k] = f(Y_{s+1}[k], Y_{s+2}[k+1], x[k], x[k+1]).
 *
 * `g0` is the global index of element 0 of x/y (lets a caller evaluate a
 * sub-range: results at index k depend only on indices >= k).
 */
static void sweep_range(const float *x, float *y, size_t g0, size_t count,
                        size_t n, size_t L, const float *sos, int bias) {
    float *prev1, *prev2, *cur, *tmp;
    size_t s, t;
    if (!L || !count) return;
    prev1 = malloc(3 * count * sizeof(float));
    prev2 = prev1 + count;
    cur = prev2 + count;
    memcpy(prev1, y, count * sizeof(float));
    memcpy(prev2, y, count * sizeof(float));
    for (s = L; s-- > 0;) {
        const float b0 = sos[s * 6 + 0], b1 = sos[s * 6 + 1];
        const float a0 = sos[s * 6 + 3], a1 = sos[s * 6 + 4], a2 = sos[s * 6 + 5];
        const size_t lo = L + s, hi = n - 1 + L + s; /* global indices section s rewrites */
        for (t = 0; t < count; ++t) {
            const size_t k = g0 + t;
            if (k >= lo && k <= hi && t + 1 < count) {
                const float p0 = b0 * prev1[t];
                const float p1 = b1 * prev2[t + 1];
                float acc = p0 + p1;
                float rhs;
                if (bias) acc = acc + 1;                 /* filter.c:227,252 vs :255 */
                rhs = a0 + a1 * x[t];
                rhs = rhs + a2 * x[t + 1];
                cur[t] = acc - rhs;
            } else {
                cur[t] = prev1[t];
            }
        }
        tmp = prev2; prev2 = prev1; prev1 = cur; cur = tmp;
    }
    memcpy(y, prev1, count * sizeof(float));
    /* prev1/prev2/cur are rotations of one allocation; free its base */
    {
        float *base = prev1;
        if (prev2 < base) base = prev2;
        if (cur < base) base = cur;
        free(base);
    }
}

void fmo_sweep_lane(const float *x, float *y, size_t n, size_t L, const float *sos, int bias) {
    sweep_range(x, y, 0, n + 2 * L, n, L, sos, bias);
}

/* ---------------------------------------------------------- chunk pipeline */

/*
 * One pass of applyComplexFilter (src/filter.c:233-259) as the reference's buffers make it behave.
 *   xs    : B interleaved floats (nc = B/2 complex samples), the x region
 *   init  : initial contents of the first `ninit` floats of the y region (what an earlier pass
 *           spilled there), zero elsewhere
 *   alias : 1 -- the y region follows the x region in the same arena (src/matrix.c:260-267), so what
 *               the sweep READS as x[B+t] is the already-final head of y;
 *           0 -- x is a buffer of its own whose slack is zero (highpassDc, src/matrix.c:142-157: a
 *               calloc'd block of B floats; the reference reads a few floats past it)
 *   ys    : B floats, the valid outputs;  spill : the 4L-2 floats the sweep WRITES past y[B-1]
 * Modelled as: (1) sweep with the extension every valid output can see -- the never-rewritten head
 * of y (= init) when aliased, zero otherwise; (2) when aliased, re-sweep just the indices >= nc with
 * the extension set to the final head.
 */
static void complex_pass(const float *xs, size_t nc, size_t L, const float *sos, const float *init, size_t ninit,
                         int alias, float *ys, float *spill) {
    const size_t clen = nc + 2 * L;
    float *lane_x = calloc(clen, sizeof(float)), *lane_y = calloc(clen, sizeof(float));
    float *tail_x = calloc(2 * L + 1, sizeof(float)), *tail_y = calloc(2 * L + 1, sizeof(float));
    size_t i;
    int q;
    for (q = 0; q < 2; ++q) {
        memset(lane_x, 0, clen * sizeof(float));
        memset(lane_y, 0, clen * sizeof(float));
        for (i = 0; i < nc; ++i) lane_x[i] = xs[2 * i + (size_t) q];
        for (i = 0; 2 * i + (size_t) q < ninit && i < nc; ++i) lane_y[i] = init[2 * i + (size_t) q];
        if (alias)
            for (i = 0; i < L; ++i) lane_x[nc + i] = lane_y[i]; /* heads are never rewritten */
        sweep_range(lane_x, lane_y, 0, clen, nc, L, sos, q == 0);
        for (i = 0; i < nc; ++i) ys[2 * i + (size_t) q] = lane_y[i];
        if (alias) {
            for (i = 0; i < 2 * L + 1; ++i) {
                tail_x[i] = i < 2 * L ? lane_y[i] : 0.f;
                tail_y[i] = 0.f;
            }
            sweep_range(tail_x, tail_y, nc, 2 * L + 1, nc, L, sos, q == 0);
        }
        /* float f past the end of y is lane f&1 of complex index nc + f/2 */
        for (i = 0; i + 1 < 2 * L; ++i) spill[2 * i + (size_t) q] = alias ? tail_y[i] : lane_y[nc + i];
    }
    free(lane_x);
    free(lane_y);
    free(tail_x);
    free(tail_y);
}

/*
 * src/matrix.c:236-273 (demodulation on), one chunk:
 *   conditioner (-q0 shiftOrigin, -q3 normalizeInput, -q2 highpassDc = shiftOrigin + a degree-3
 *   Butterworth HIGHPASS at fc = 1 designed at -S, applied like any complex filter)
 *   -> [input filter, -L] -> discriminator -> output filter.
 * Every complex pass writes 4L-2 floats past its output region; those land at the start of the
 * region the NEXT pass accumulates into (y_c of the input filter, or y_o of the output filter),
 * whose first L entries are never rewritten -- so they are observable in the first outputs of
 * every chunk.
 */
static int one_chunk(const uint8_t *in, size_t B, const fmo_config *cfg,
                     size_t Lin, size_t Lout, const float *sos_in, const float *sos_out,
                     float off[2], float *out) {
    const size_t n_out = B >> 2, nc = B >> 1;
    const size_t ylen = n_out + 2 * Lout;
    const unsigned conditioner = (cfg->mode >> 2) & 3u;
    float *xs = malloc(B * sizeof(float)), *ys = malloc(B * sizeof(float));
    float *dem = calloc(ylen, sizeof(float));
    float *y = calloc(ylen, sizeof(float));
    float spill[4 * 8], sos_hp[2 * 6];
    size_t nspill = 0, i;
    if (!xs || !ys || !dem || !y) return -1;
    if (conditioner == 3u) fmo_normalize(in, B, xs);   /* matrix.c:208-222 */
    else if (conditioner == 1u) fmo_correct_iq(in, B, (float) (50. / cfg->fs), off, xs); /* matrix.c:209-212 */
    else fmo_shift(in, B, xs);
    if (conditioner == 2u) {                           /* matrix.c:142-157 */
        fmo_design_sos(FMO_BUTTER_HP, 3, 1., cfg->fs, 0., sos_hp);
        complex_pass(xs, nc, 2, sos_hp, NULL, 0, 0, ys, spill);
        nspill = 4 * 2 - 2;
        memcpy(xs, ys, B * sizeof(float));
    }
    if (Lin) {
        float next[4 * 8];
        complex_pass(xs, nc, Lin, sos_in, spill, nspill, 1, ys, next);
        nspill = 4 * Lin - 2;
        memcpy(spill, next, nspill * sizeof(float));
        memcpy(xs, ys, B * sizeof(float));
    }
    for (i = 0; i < nspill && i < ylen; ++i) y[i] = spill[i];
    fmo_discriminate(xs, B, dem);
    sweep_range(dem, y, 0, ylen, n_out, Lout, sos_out, 1);
    memcpy(out, y, n_out * sizeof(float));
    free(xs);
    free(ys);
    free(dem);
    free(y);
    return 0;
}

size_t fmo_process_chunks(const uint8_t *in, size_t nchunks, size_t B,
                          const fmo_config *cfg, float *out) {
    unsigned deg_in = cfg->deg_in;
    size_t Lin, Lout, c;
    float *sos_in = NULL, *sos_out;
    if (cfg->fc_in != 0. && !deg_in) deg_in = cfg->deg_out;          /* matrix.c:190-192 */
    if (cfg->fc_in == 0. && deg_in) return 0;                        /* reference: undefined */
    if (!cfg->deg_out || (B & 3)) return 0;
    Lin = fmo_sections(deg_in);
    Lout = fmo_sections(cfg->deg_out);
    /* the aliasing model above assumes the chunk is longer than the head it re-reads */
    if ((B >> 1) < 2 * Lin + 2) return 0;
    sos_out = malloc(Lout * 6 * sizeof(float));
    fmo_design_sos(cfg->mode & 1u, cfg->deg_out, cfg->fc_out, cfg->fs, cfg->eps, sos_out);
    if (Lin) {
        sos_in = malloc(Lin * 6 * sizeof(float));
        fmo_design_sos((cfg->mode >> 1) & 1u, deg_in, cfg->fc_in, cfg->fs, cfg->eps, sos_in);
    }
    {
        float off[2] = {0.f, 0.f}; /* -q1: the reference's static offsets, zero at process start */
        for (c = 0; c < nchunks; ++c)
            if (one_chunk(in + c * B, B, cfg, Lin, Lout, sos_in, sos_out, off, out + c * (B >> 2))) return 0;
    }
    free(sos_in);
    free(sos_out);
    return nchunks * (B >> 2);
}
