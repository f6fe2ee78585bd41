/*
 * fmd_design.c -- host-side set-up for the FM path: IIR prototype design in
 * long double, and folding of the designed sections into the finite stencil the
 * reference's section sweep actually computes.
 *
 * Replaces (same results, float-for-float -- tests/test_oracle_golden.py::test_design_table_port_and_product
 * checks the 288-design table generated from the reference):
 *   processFilterOption                 /root/reference/src/matrix.c:25-80
 *   warpButter.. / warpCheby1..         /root/reference/src/filter.c:22-101
 *   zp2Sos                              /root/reference/src/filter.c:104-140
 *   transformBilinear                   /root/reference/src/filter.c:142-210
 *
 * Plain C (gcc): nvcc's host pass is not asked to handle long double.
 */
#include "fmd_internal.h"

#include <math.h>
#include <string.h>

typedef long double ld;

typedef struct {
    ld pole_re, pole_im; /* digital pole of the k-th conjugate pair (or the real pole) */
    ld gain_re, gain_im; /* the factor this root contributes to the running gain      */
} root_t;

/* k-th Butterworth root after the reference's pre-warped bilinear map.
 * alpha = 1/sin(2w), beta = tan(w), w = pi*fc/fs. */
static root_t butterworth_root(ld alpha, ld beta, unsigned k, unsigned n, int highpass) {
    root_t r;
    const ld theta = M_PI_2 * (1. / (ld) n * (-1. + (ld) (k << 1)) + 1.);
    const ld ct = cosl(theta);
    const ld inv = 1. / (ct - alpha);
    const ld zr = (-beta + ct) * inv;
    r.pole_re = -zr + 1.;
    r.pole_im = sinl(theta) * inv;
    r.gain_re = highpass ? 2. - zr : zr;
    r.gain_im = r.pole_im;
    return r;
}

/* k-th Chebyshev type-1 root; tng = tan(w*wh) (lowpass) or its reciprocal (highpass) */
static root_t chebyshev1_root(ld tng, ld ripple, unsigned k, unsigned n) {
    root_t r;
    const ld inv_n = 1. / (ld) n;
    const ld v = logl((1. + powl(10., 0.5 * ripple)) / sqrtl(-1. + powl(10., ripple))) * inv_n;
    const ld theta = M_PI_2 * (inv_n * (-1. + (ld) (k << 1)));
    const ld a = cosl(theta) * coshl(v) * tng;
    const ld b = sinl(theta) * sinhl(v) * tng;
    const ld c = a * a + b * b;
    const ld d = 1. / (1. + c + 2. * b);
    const ld zi = 2. * a * d;
    const ld zr = 2. * (b + c) * d;
    r.pole_re = -zr + 1.;
    r.pole_im = zi;
    r.gain_re = zr;
    r.gain_im = zi;
    return r;
}

int fmd_design_sos(unsigned kind, unsigned degree, double fc_in, double fs_in, double eps_in, float *out) {
    const unsigned n = degree;
    const unsigned pairs = n >> 1;
    const unsigned rows = pairs + (n & 1u);
    const int odd = (int) (n & 1u);
    const int highpass = (kind & 2u) != 0;
    const int chebyshev = (kind & 1u) != 0;
    const ld fc = (ld) fc_in, fs = (ld) fs_in, ripple = (ld) eps_in;
    const ld w = M_PI * fc / fs;
    const ld zero_at = highpass ? 1. : -1.; /* all n zeros sit at z = -1 (LP) or +1 (HP) */
    ld table[FMD_MAX_SECTIONS][6];
    root_t roots[FMD_MAX_SECTIONS];
    ld alpha = 0., beta = 0., tng = 0.;
    ld gr, gi = 0.;
    unsigned k, r, c;

    if (!out || kind > 3u) return fmd_set_error(FMD_EINVAL, "fmd_design_sos: bad kind %u", kind);
    if (!n || rows > FMD_MAX_SECTIONS)
        return fmd_set_error(FMD_EINVAL, "fmd_design_sos: degree %u outside 1..%d", degree, 2 * FMD_MAX_SECTIONS);

    if (chebyshev) {
        const ld wh = coshl(1. / (ld) n * acoshl(1. / sqrtl(powl(10., ripple) - 1.)));
        tng = tanl(w * wh);
        if (highpass) tng = 1. / tng;
    } else {
        alpha = 1. / sinl(2. * w);
        beta = tanl(w);
    }
    /* even-order lowpass Chebyshev starts the gain at 1/sqrt(2) (passband edge convention) */
    gr = (!odd && chebyshev && !highpass) ? M_SQRT1_2 : 1.;

    for (k = 1; k <= rows; ++k) {
        const root_t rt = chebyshev ? chebyshev1_root(tng, ripple, k, n)
                                    : butterworth_root(alpha, beta, k, n, highpass);
        roots[k - 1] = rt;
        if (k <= pairs) { /* conjugate pair: multiply by |g|^2 */
            const ld m2 = rt.gain_re * rt.gain_re + rt.gain_im * rt.gain_im;
            gr *= m2;
            gi *= m2;
        } else { /* the unpaired root of an odd order: complex multiply */
            const ld t = rt.gain_re * gr - rt.gain_im * gi;
            gi = rt.gain_re * gi + rt.gain_im * gr;
            gr = t;
        }
    }
    gr /= (ld) (1 << n);

    for (r = 0; r < pairs; ++r) {
        const ld pr = roots[r].pole_re, pi = roots[r].pole_im;
        table[r][0] = 1.;
        table[r][1] = -2. * zero_at;
        table[r][2] = zero_at * zero_at + 0. * 0.;
        table[r][3] = 1.;
        table[r][4] = -2. * pr;
        table[r][5] = pr * pr + pi * pi;
    }
    if (odd) { /* first-order row for the real pole */
        table[pairs][0] = 1.;
        table[pairs][1] = -zero_at;
        table[pairs][2] = 0.;
        table[pairs][3] = 1.;
        table[pairs][4] = -roots[pairs].pole_re;
        table[pairs][5] = 0.;
    }
    { /* overall gain rides on the numerator of the last conjugate-pair row (odd: the extra row) */
        const unsigned g_row = odd ? pairs : pairs - 1;
        table[g_row][0] *= gr;
        table[g_row][1] *= gr;
        table[g_row][2] *= gr;
    }
    for (r = 0; r < rows; ++r)
        for (c = 0; c < 6; ++c) out[r * 6 + c] = (float) table[r][c];
    return FMD_OK;
}

/*
 * The reference applies the sections as an in-place sweep, not as a recursion
 * in time.  Away from the first 2L-1 samples of a chunk
 * the result is affine in a short window of the input:
 *
 *   Y_s[k] = b0_s*Y_{s+1}[k] + b1_s*Y_{s+2}[k+1] + bias - a0_s - a1_s*x[k] - a2_s*x[k+1],
 *   Y_L = Y_{L+1} = 0,   y[k] = Y_0[k].
 *
 * Propagate (dc, taps[]) through that recurrence in double.
 */
int fmd_fold_taps(const float *sos, unsigned sections, int bias,
                  float taps[FMD_MAX_TAPS], float *dc, unsigned *ntaps) {
    double c[FMD_MAX_SECTIONS + 2][FMD_MAX_TAPS + 1];
    double d[FMD_MAX_SECTIONS + 2];
    unsigned t, nt;
    int s;
    if (!sos || !taps || !dc || !ntaps || !sections || sections > FMD_MAX_SECTIONS)
        return fmd_set_error(FMD_EINVAL, "fmd_fold_taps: bad arguments");
    memset(c, 0, sizeof c);
    memset(d, 0, sizeof d);
    for (s = (int) sections - 1; s >= 0; --s) {
        const double b0 = sos[s * 6 + 0], b1 = sos[s * 6 + 1];
        const double a0 = sos[s * 6 + 3], a1 = sos[s * 6 + 4], a2 = sos[s * 6 + 5];
        d[s] = b0 * d[s + 1] + b1 * d[s + 2] + ((bias ? 1.0 : 0.0) - a0);
        for (t = 0; t <= FMD_MAX_TAPS; ++t)
            c[s][t] = b0 * c[s + 1][t] + (t ? b1 * c[s + 2][t - 1] : 0.0);
        c[s][0] -= a1;
        c[s][1] -= a2;
    }
    nt = (sections - 1) / 2 + 2; /* structural width; non-finite coefficients stay "non-zero" */
    while (nt > 1 && c[0][nt - 1] == 0.0) --nt;
    if (nt > FMD_MAX_TAPS) return fmd_set_error(FMD_EINVAL, "fmd_fold_taps: stencil wider than %d", FMD_MAX_TAPS);
    for (t = 0; t < FMD_MAX_TAPS; ++t) taps[t] = t < nt ? (float) c[0][t] : 0.f;
    *dc = (float) d[0];
    *ntaps = nt;
    return FMD_OK;
}
