/*
 * fmd_quality.c -- host-side tables of the opt-in quality pipeline (see fmd_quality.h).
 *
 * Turns the designed section rows into the state-space matrices of the cascade and into every
 * power of the transition matrix the device-side scans need.  Long double throughout; the
 * device sees float32 copies.  Plain C (gcc), like fmd_design.c.
 */
#include "fmd_quality.h"

#include <string.h>

#include "fmd_internal.h"

typedef long double ld;

/* one sample through the cascade in direct form II transposed: state q (2 per section) */
static ld cascade_step(const fmq_design *d, ld *q, ld u) {
    for (int i = 0; i < d->L; ++i) {
        const ld b0 = d->sos[i][0], b1 = d->sos[i][1], b2 = d->sos[i][2], a1 = d->sos[i][3], a2 = d->sos[i][4];
        const ld w = b0 * u + q[2 * i];
        q[2 * i] = b1 * u - a1 * w + q[2 * i + 1];
        q[2 * i + 1] = b2 * u - a2 * w;
        u = w;
    }
    return u;
}

int fmq_design_from_sos(const float *sos6, unsigned sections, fmq_design *d) {
    if (!sos6 || !d || sections == 0 || sections > FMQ_MAXL)
        return fmd_set_error(FMD_EINVAL, "quality mode takes 1..%d sections per filter (degree <= %d)", FMQ_MAXL,
                             2 * FMQ_MAXL);
    memset(d, 0, sizeof *d);
    d->L = (int) sections;
    d->D = 2 * (int) sections;
    for (unsigned i = 0; i < sections; ++i) {
        const float *r = sos6 + 6 * i;
        if (!(r[3] == 1.0f)) return fmd_set_error(FMD_EINVAL, "section %u has a0 != 1", i);
        for (int k = 0; k < 6; ++k)
            if (!(r[k] == r[k]) || r[k] > 3.0e38f || r[k] < -3.0e38f)
                return fmd_set_error(FMD_EINVAL, "filter design is not finite");
        d->sos[i][0] = r[0];
        d->sos[i][1] = r[1];
        d->sos[i][2] = r[2];
        d->sos[i][3] = r[4];
        d->sos[i][4] = r[5];
    }
    /* the system is linear: probe it with unit vectors */
    for (int j = 0; j < d->D; ++j) {
        ld q[FMQ_MAXD] = {0};
        q[j] = 1.0L;
        d->C[j] = cascade_step(d, q, 0.0L);
        for (int i = 0; i < d->D; ++i) d->A[i][j] = q[i];
    }
    {
        ld q[FMQ_MAXD] = {0};
        d->Dd = cascade_step(d, q, 1.0L);
        for (int i = 0; i < d->D; ++i) d->B[i] = q[i];
    }
    return FMD_OK;
}

static void mat_mul(int D, const ld *x, const ld *y, ld *out) { /* out = x * y, may not alias */
    for (int i = 0; i < D; ++i)
        for (int j = 0; j < D; ++j) {
            ld acc = 0.0L;
            for (int k = 0; k < D; ++k) acc += x[i * D + k] * y[k * D + j];
            out[i * D + j] = acc;
        }
}

static void mat_identity(int D, ld *m) {
    for (int i = 0; i < D * D; ++i) m[i] = 0.0L;
    for (int i = 0; i < D; ++i) m[i * D + i] = 1.0L;
}

static void mat_pack(const fmq_design *d, ld *m) {
    for (int i = 0; i < d->D; ++i)
        for (int j = 0; j < d->D; ++j) m[i * d->D + j] = d->A[i][j];
}

void fmq_matpow(const fmq_design *d, unsigned long long k, ld *out) {
    const int D = d->D;
    ld base[FMQ_MAXD * FMQ_MAXD], tmp[FMQ_MAXD * FMQ_MAXD];
    mat_pack(d, base);
    mat_identity(D, out);
    while (k) {
        if (k & 1ull) {
            mat_mul(D, base, out, tmp);
            memcpy(out, tmp, sizeof(ld) * (size_t) (D * D));
        }
        mat_mul(D, base, base, tmp);
        memcpy(base, tmp, sizeof(ld) * (size_t) (D * D));
        k >>= 1;
    }
}

static void to_float(const ld *m, int n, float *out) {
    for (int i = 0; i < n; ++i) out[i] = (float) m[i];
}

void fmq_matpow_f(const fmq_design *d, unsigned long long k, float *out) {
    ld m[FMQ_MAXD * FMQ_MAXD];
    fmq_matpow(d, k, m);
    to_float(m, d->D * d->D, out);
}

void fmq_build_consts(const fmq_design *d, unsigned unit, unsigned tile, fmq_consts *c) {
    memset(c, 0, sizeof *c);
    c->L = d->L;
    c->D = d->D;
    memcpy(c->sos, d->sos, sizeof c->sos);
    for (int j = 0; j < 5; ++j) fmq_matpow_f(d, (unsigned long long) unit << j, c->lvl[j]);
    fmq_matpow_f(d, 32ull * unit, c->warp);
    fmq_matpow_f(d, (unsigned long long) tile, c->tile);
}

void fmq_table_layout(const fmq_design *d, unsigned T, fmq_table_sizes *sz) {
    const unsigned D = (unsigned) d->D;
    sz->lane = 32u * D * D;
    sz->apow = (T + 1u) * D * D;
    sz->ck = T * D;
    sz->ckn = T * FMQ_THREADS * D;
    sz->lane_tile = 32u * D * D;
}

unsigned fmq_build_tables(const fmq_design *d, unsigned T, float *lane, float *apow, float *ck, float *ckn,
                          float *lane_tile) {
    const int D = d->D;
    const unsigned TILE = T * FMQ_THREADS;
    ld a[FMQ_MAXD * FMQ_MAXD], pt[FMQ_MAXD * FMQ_MAXD], cur[FMQ_MAXD * FMQ_MAXD], tmp[FMQ_MAXD * FMQ_MAXD];
    unsigned zir_len = 1;
    mat_pack(d, a);
    fmq_matpow(d, T, pt);
    mat_identity(D, cur);
    for (int l = 0; l < 32; ++l) { /* (A^T)^l */
        to_float(cur, D * D, lane + l * D * D);
        mat_mul(D, pt, cur, tmp);
        memcpy(cur, tmp, sizeof tmp);
    }
    fmq_matpow(d, TILE, pt);
    mat_identity(D, cur);
    for (int l = 0; l < 32; ++l) { /* (A^TILE)^l */
        to_float(cur, D * D, lane_tile + l * D * D);
        mat_mul(D, pt, cur, tmp);
        memcpy(cur, tmp, sizeof tmp);
    }
    mat_identity(D, cur);
    for (unsigned k = 0; k <= T; ++k) { /* A^k */
        to_float(cur, D * D, apow + k * D * D);
        mat_mul(D, a, cur, tmp);
        memcpy(cur, tmp, sizeof tmp);
    }
    /* row vector C A^k, k = 0 .. TILE-1 (the first T of them are the thread-level table) */
    ld row[FMQ_MAXD], nxt[FMQ_MAXD];
    for (int j = 0; j < D; ++j) row[j] = d->C[j];
    for (unsigned k = 0; k < TILE; ++k) {
        for (int j = 0; j < D; ++j) {
            ckn[k * D + j] = (float) row[j];
            if (k < T) ck[k * D + j] = (float) row[j];
            if (row[j] >= FMQ_ZIR_EPS || row[j] <= -FMQ_ZIR_EPS || row[j] != row[j]) zir_len = (unsigned) k + 1u;
        }
        for (int j = 0; j < D; ++j) {
            ld acc = 0.0L;
            for (int i = 0; i < D; ++i) acc += row[i] * a[i * D + j];
            nxt[j] = acc;
        }
        memcpy(row, nxt, sizeof row);
    }
    return zir_len;
}

/* ---- flat exports for the CPU-side test of the scan algebra (tests/test_quality_tables.py);
 *      not part of include/fmdemod.h ---- */

int fmq_test_sizes(unsigned sections, unsigned T, unsigned *consts_bytes, unsigned sizes[5]) {
    fmq_design d;
    memset(&d, 0, sizeof d);
    d.L = (int) sections;
    d.D = 2 * (int) sections;
    fmq_table_sizes sz;
    fmq_table_layout(&d, T, &sz);
    *consts_bytes = (unsigned) sizeof(fmq_consts);
    sizes[0] = sz.lane;
    sizes[1] = sz.apow;
    sizes[2] = sz.ck;
    sizes[3] = sz.ckn;
    sizes[4] = sz.lane_tile;
    return FMD_OK;
}

int fmq_test_build(const float *sos6, unsigned sections, unsigned T, fmq_consts *c, fmq_consts *c_tile, float *lane, float *apow,
                   float *ck, float *ckn, float *lane_tile, double *abcd /* D*D + D + D + 1 */, unsigned *zir_len) {
    fmq_design d;
    int rc = fmq_design_from_sos(sos6, sections, &d);
    if (rc) return rc;
    fmq_build_consts(&d, T, T * FMQ_THREADS, c);
    fmq_build_consts(&d, T * FMQ_THREADS, T * FMQ_THREADS, c_tile);
    *zir_len = fmq_build_tables(&d, T, lane, apow, ck, ckn, lane_tile);
    int p = 0;
    for (int i = 0; i < d.D; ++i)
        for (int j = 0; j < d.D; ++j) abcd[p++] = (double) d.A[i][j];
    for (int i = 0; i < d.D; ++i) abcd[p++] = (double) d.B[i];
    for (int i = 0; i < d.D; ++i) abcd[p++] = (double) d.C[i];
    abcd[p++] = (double) d.Dd;
    return FMD_OK;
}
