/*
 * fmd_quality_host.cu -- host side of the opt-in quality pipeline (FMD_FLAG_QUALITY): table
 * upload, per-call launch sequence, stream state.  Kernels: fmd_quality.cuh; tables: fmd_quality.c;
 * what it computes and why it is not a reference mode: fmd_quality.h / DESIGN.md section 9.
 */
#include <cuda_runtime.h>

#include <cmath>
#include <cstddef>
#include <cstdlib>
#include <cstring>
#include <new>

#include "fmd_ctx.h"
#include "fmd_correctiq.cuh"
#include "fmd_q1_fused.cuh"
#include "fmd_quality.cuh"

/* one cascade of the quality pipeline: host design + constants + device tables */
struct fmq_filter {
    fmq_design des;
    fmq_consts c;      /* scan over the threads of a CTA (unit T)   */
    fmq_consts c_tile; /* scan over the tiles of a launch (unit TILE) */
    fmq_table_sizes sz;
    unsigned zir_len = 0; /* tile offsets whose zero-input response is not negligible */
    float *d_tab = nullptr; /* lane | apow | ck | ckn | lane_tile */
    const float *lane() const { return d_tab; }
    const float *apow() const { return d_tab + sz.lane; }
    const float *ck() const { return d_tab + sz.lane + sz.apow; }
    const float *ckn() const { return d_tab + sz.lane + sz.apow + sz.ck; }
    const float *lane_tile() const { return d_tab + sz.lane + sz.apow + sz.ck + sz.ckn; }
};

/* FMD_FLAG_QUALITY state of a context (fmd_quality.cuh) */
struct fmq_ctx {
    unsigned decim = 1;
    unsigned T = FMQ_T, tile = FMQ_TILE; /* samples per thread / per CTA: FMQ_T_BIG for large decimation factors */
    unsigned lin = 0, lout = 0; /* sections */
    fmq_filter fin, fout;
    FmqState *d_state = nullptr;
    float *d_ein = nullptr, *d_qin = nullptr, *d_eout = nullptr, *d_qout = nullptr; /* per-tile scratch */
    unsigned *d_lb = nullptr; /* look-back ticket + per-tile flags */
    float *d_y = nullptr;     /* FMD_FLAG_S16: the float outputs of a launch, before the conversion */
    size_t y_cap = 0;
    unsigned tile_cap = 0;
    bool smem_opted_in = false;
    unsigned long long pos = 0; /* stream position: global index of the next input sample */
};

/* outputs of the next `nsamp` samples when the stream stands at `pos` */
static size_t quality_out_count(unsigned long long pos, size_t nsamp, unsigned M) {
    const unsigned long long r0 = (M - pos % M) % M;
    return nsamp > r0 ? (size_t) ((nsamp - r0 + M - 1) / M) : 0;
}

static int quality_upload_filter(fmq_filter &f, const float *sos6, unsigned sections, unsigned T) {
    int rc = fmq_design_from_sos(sos6, sections, &f.des);
    if (rc) return rc;
    fmq_build_consts(&f.des, T, T * FMQ_THREADS, &f.c);
    fmq_build_consts(&f.des, T * FMQ_THREADS, T * FMQ_THREADS, &f.c_tile);
    fmq_table_layout(&f.des, T, &f.sz);
    const size_t n = (size_t) f.sz.lane + f.sz.apow + f.sz.ck + f.sz.ckn + f.sz.lane_tile;
    float *h = (float *) malloc(n * sizeof(float));
    if (!h) return fmd_set_error(FMD_ENOMEM, "out of host memory");
    f.zir_len = fmq_build_tables(&f.des, T, h, h + f.sz.lane, h + f.sz.lane + f.sz.apow, h + f.sz.lane + f.sz.apow + f.sz.ck,
                                 h + f.sz.lane + f.sz.apow + f.sz.ck + f.sz.ckn);
    cudaError_t e = cudaMalloc(&f.d_tab, n * sizeof(float));
    if (e == cudaSuccess) e = cudaMemcpy(f.d_tab, h, n * sizeof(float), cudaMemcpyHostToDevice);
    free(h);
    if (e != cudaSuccess) return fmd_set_error(FMD_ECUDA, "quality table upload failed: %s", cudaGetErrorString(e));
    return FMD_OK;
}

int fmq_setup(fmd_ctx *c, unsigned decim) {
    fmq_ctx *q = new (std::nothrow) fmq_ctx();
    if (!q) return fmd_set_error(FMD_ENOMEM, "out of host memory");
    c->q = q;
    q->decim = decim;
    q->lout = c->lout;
    q->lin = c->lin;
    /* measured on B200: 128 samples per thread is +11 % at M = 50 and -25 % at M = 1 (shared memory
     * for the kept samples of a tile); 64 is the better span up to M = 4 */
    q->T = decim >= 8 ? FMQ_T_BIG : FMQ_T;
    if (const char *e = getenv("FMD_QUALITY_T")) { /* tuning experiments only */
        const int t = atoi(e);
        if (t == FMQ_T || t == FMQ_T_BIG) q->T = (unsigned) t;
    }
    q->tile = q->T * FMQ_THREADS;
    int rc = quality_upload_filter(q->fout, &c->sec.out[0][0], q->lout, q->T);
    if (!rc && q->lin) rc = quality_upload_filter(q->fin, &c->sec.in[0][0], q->lin, q->T);
    if (rc) return rc;
    /* the kernels for odd degrees skip the second state of the last (first-order) section */
    if (((c->cfg.out_degree & 1u) && (c->sec.out[q->lout - 1][2] != 0.0f || c->sec.out[q->lout - 1][5] != 0.0f)) ||
        (q->lin && (c->cfg.in_degree & 1u) && (c->sec.in[q->lin - 1][2] != 0.0f || c->sec.in[q->lin - 1][5] != 0.0f)))
        return fmd_set_error(FMD_EINVAL, "odd-degree design whose last section is not first order");
    FMD_CUDA(cudaMalloc(&q->d_state, sizeof(FmqState)));
    FMD_CUDA(cudaMemset(q->d_state, 0, sizeof(FmqState)));
    return FMD_OK;
}

void fmq_free(fmd_ctx *c) {
    fmq_ctx *q = c->q;
    if (!q) return;
    if (q->fin.d_tab) cudaFree(q->fin.d_tab);
    if (q->fout.d_tab) cudaFree(q->fout.d_tab);
    if (q->d_state) cudaFree(q->d_state);
    if (q->d_ein) cudaFree(q->d_ein);
    if (q->d_qin) cudaFree(q->d_qin);
    if (q->d_eout) cudaFree(q->d_eout);
    if (q->d_qout) cudaFree(q->d_qout);
    if (q->d_lb) cudaFree(q->d_lb);
    if (q->d_y) cudaFree(q->d_y);
    delete q;
    c->q = nullptr;
}

static int quality_reserve_tiles(fmq_ctx *q, unsigned ntiles) {
    if (ntiles <= q->tile_cap) return FMD_OK;
    /* cudaFree waits for the kernels that may still read the old scratch */
    if (q->d_ein) cudaFree(q->d_ein);
    if (q->d_qin) cudaFree(q->d_qin);
    if (q->d_eout) cudaFree(q->d_eout);
    if (q->d_qout) cudaFree(q->d_qout);
    if (q->d_lb) cudaFree(q->d_lb);
    q->d_lb = nullptr;
    q->d_ein = q->d_qin = q->d_eout = q->d_qout = nullptr;
    q->tile_cap = 0;
    const unsigned cap = ntiles < 4096u ? 4096u : ntiles;
    const size_t per_in = (size_t) cap * 2 * FMQ_MAXD * sizeof(float), per_out = (size_t) cap * FMQ_MAXD * sizeof(float);
    FMD_CUDA(cudaMalloc(&q->d_ein, per_in));
    FMD_CUDA(cudaMalloc(&q->d_qin, per_in));
    FMD_CUDA(cudaMalloc(&q->d_eout, per_out));
    FMD_CUDA(cudaMalloc(&q->d_qout, per_out));
    FMD_CUDA(cudaMalloc(&q->d_lb, ((size_t) cap + 1u) * sizeof(unsigned)));
    q->tile_cap = cap;
    return FMD_OK;
}

typedef void (*fmq_main_fn)(const uint8_t *, float *, const FmqGeom, const fmq_consts, const fmq_consts, const float *,
                            const float *, const float *, const float *, const float *, const float *,
                            const FmqLookback, const FmqMat, const FmqMat, float *, FmqState *);

template <int NIN, int T>
static fmq_main_fn quality_pick_out(unsigned nout) {
    switch (nout) {
        case 1: return fmq_main_kernel<NIN, 1, T>;
        case 2: return fmq_main_kernel<NIN, 2, T>;
        case 3: return fmq_main_kernel<NIN, 3, T>;
        case 4: return fmq_main_kernel<NIN, 4, T>;
        case 5: return fmq_main_kernel<NIN, 5, T>;
        case 6: return fmq_main_kernel<NIN, 6, T>;
        default: return nullptr;
    }
}

/* by filter DEGREE: an odd degree ends in a first-order section, which the kernels exploit */
template <int T>
static fmq_main_fn quality_pick_in(unsigned nin, unsigned nout) {
    switch (nin) {
        case 0: return quality_pick_out<0, T>(nout);
        case 1: return quality_pick_out<1, T>(nout);
        case 2: return quality_pick_out<2, T>(nout);
        case 3: return quality_pick_out<3, T>(nout);
        case 4: return quality_pick_out<4, T>(nout);
        case 5: return quality_pick_out<5, T>(nout);
        case 6: return quality_pick_out<6, T>(nout);
        default: return nullptr;
    }
}

static fmq_main_fn quality_pick_main(unsigned nin, unsigned nout, unsigned T) {
    return T == FMQ_T_BIG ? quality_pick_in<FMQ_T_BIG>(nin, nout) : quality_pick_in<FMQ_T>(nin, nout);
}

/* one launch group: `n` samples (< 2^31) that continue the stream; *nout = floats written */
static int quality_launch(fmd_ctx *c, const uint8_t *d_in, unsigned n, float *d_out, cudaStream_t s, size_t *nout) {
    fmq_ctx *q = c->q;
    FmqGeom g;
    g.n = n;
    g.ntiles = (n + q->tile - 1) / q->tile;
    g.decim = q->decim;
    g.r0 = (unsigned) ((q->decim - q->pos % q->decim) % q->decim);
    g.n_out = (unsigned) quality_out_count(q->pos, n, q->decim);
    *nout = g.n_out;
    int rc = quality_reserve_tiles(q, g.ntiles);
    if (rc) return rc;
    const unsigned len_last = n - (g.ntiles - 1u) * q->tile;
    /* the scan over tiles: 32 warps, each a run of `runlen` tiles (a multiple of 32) */
    const unsigned nwarps = FMQ_SCAN_THREADS / 32;
    const unsigned runlen = ((g.ntiles + nwarps - 1u) / nwarps + 31u) & ~31u;
    FmqMat prun, pl;

    FmqLookback lb = {nullptr, nullptr, nullptr, nullptr};
    FmqMat pl_in, wtile_in;
    memset(&pl_in, 0, sizeof pl_in);
    memset(&wtile_in, 0, sizeof wtile_in);
    if (q->lin) {
        /* the input cascade's carry travels between the tiles INSIDE the main kernel (look-back):
         * ticket + one flag per tile, cleared per launch; E and S records per tile */
        FMD_CUDA(cudaMemsetAsync(q->d_lb, 0, ((size_t) g.ntiles + 1u) * sizeof(unsigned), s));
        lb.ticket = q->d_lb;
        lb.flag = q->d_lb + 1;
        lb.E = q->d_ein;
        lb.S = q->d_qin;
        fmq_matpow_f(&q->fin.des, len_last, pl_in.m);
        memcpy(wtile_in.m, q->fin.c_tile.warp, sizeof wtile_in.m); /* A^(32 TILE): one look-back round */
    }

    /* FMD_FLAG_S16: the kernels write floats into a scratch of the context; the caller's buffer
     * receives the converted samples at the end */
    const bool s16 = (c->cfg.flags & FMD_FLAG_S16) != 0;
    float *const user_out = d_out;
    if (s16) {
        if (g.n_out > q->y_cap) {
            if (q->d_y) cudaFree(q->d_y);
            q->d_y = nullptr;
            q->y_cap = 0;
            const size_t cap = g.n_out < (1u << 20) ? (1u << 20) : g.n_out;
            FMD_CUDA(cudaMalloc(&q->d_y, cap * sizeof(float)));
            q->y_cap = cap;
        }
        d_out = q->d_y;
    }

    fmq_main_fn fn = quality_pick_main(q->lin ? c->cfg.in_degree : 0u, c->cfg.out_degree, q->T);
    if (!fn) return fmd_set_error(FMD_EINVAL, "no quality kernel for degrees %u/%u", c->cfg.in_degree, c->cfg.out_degree);
    /* shared memory for the kept samples of one tile */
    const unsigned keep_cap = q->tile / q->decim + 2u;
    const unsigned keep = keep_cap < q->tile ? keep_cap : q->tile;
    const size_t smem = (size_t) (keep + FMQ_THREADS) * sizeof(float); /* one pad word per owning thread */
    cudaEvent_t *pe = nullptr;
    rc = fmd_prof_events(c, &pe);
    if (rc) return rc;
    if (pe) FMD_CUDA(cudaEventRecord(pe[0], s));
    if (!q->smem_opted_in) { /* static + dynamic shared memory may pass 48 KiB for small M (M = 1 keeps a whole tile) */
        FMD_CUDA(cudaFuncSetAttribute((const void *) fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
        q->smem_opted_in = true;
    }
    fn<<<g.ntiles, FMQ_THREADS, smem, s>>>(d_in, d_out, g, q->lin ? q->fin.c : q->fout.c, q->fout.c,
                                           q->lin ? q->fin.lane() : nullptr, q->fout.lane(), q->fout.apow(),
                                           q->fout.ck(), q->lin ? q->fin.apow() : nullptr,
                                           q->lin ? q->fin.lane_tile() : nullptr, lb, pl_in, wtile_in, q->d_eout, q->d_state);
    FMD_CUDA(cudaGetLastError());
    if (pe) FMD_CUDA(cudaEventRecord(pe[1], s));

    memset(&prun, 0, sizeof prun);
    memset(&pl, 0, sizeof pl);
    fmq_matpow_f(&q->fout.des, (unsigned long long) q->tile * runlen, prun.m);
    fmq_matpow_f(&q->fout.des, len_last, pl.m);
    const uint8_t *last = d_in + 2ull * (n - 1u);
    /* zero-input response of the tile carry: only the leading zir_len offsets of every tile */
    const unsigned zl = q->fout.zir_len < q->tile ? q->fout.zir_len : q->tile;
    const dim3 zgrid(g.ntiles, (zl / q->decim + 1u + 255u) / 256u);
    switch (q->lout) {
        case 1:
            fmq_scan_tiles_kernel<2, 1><<<1, FMQ_SCAN_THREADS, 0, s>>>(q->d_eout, q->d_qout, q->d_state->out_q, q->fout.c_tile, q->fout.lane_tile(), prun, pl, g.ntiles, runlen, last, q->d_state->s_prev);
            if (g.n_out) fmq_zir_kernel<2><<<zgrid, 256, 0, s>>>(d_out, g, zl, q->tile, q->fout.ckn(), q->d_qout);
            break;
        case 2:
            fmq_scan_tiles_kernel<4, 1><<<1, FMQ_SCAN_THREADS, 0, s>>>(q->d_eout, q->d_qout, q->d_state->out_q, q->fout.c_tile, q->fout.lane_tile(), prun, pl, g.ntiles, runlen, last, q->d_state->s_prev);
            if (g.n_out) fmq_zir_kernel<4><<<zgrid, 256, 0, s>>>(d_out, g, zl, q->tile, q->fout.ckn(), q->d_qout);
            break;
        default:
            fmq_scan_tiles_kernel<6, 1><<<1, FMQ_SCAN_THREADS, 0, s>>>(q->d_eout, q->d_qout, q->d_state->out_q, q->fout.c_tile, q->fout.lane_tile(), prun, pl, g.ntiles, runlen, last, q->d_state->s_prev);
            if (g.n_out) fmq_zir_kernel<6><<<zgrid, 256, 0, s>>>(d_out, g, zl, q->tile, q->fout.ckn(), q->d_qout);
            break;
    }
    if (s16 && g.n_out) {
        fmq_to_s16_kernel<<<(g.n_out + 255u) / 256u, 256, 0, s>>>(d_out, reinterpret_cast<short *>(user_out), g.n_out);
        c->launches++;
    }
    FMD_CUDA(cudaGetLastError());
    c->launches += g.n_out ? 3 : 2;
    q->pos += n;
    return FMD_OK;
}

int fmq_demod_device(fmd_ctx *c, const uint8_t *d_in, size_t nbytes, float *d_out, cudaStream_t s) {
    if (nbytes & 1u) return fmd_set_error(FMD_EINVAL, "quality mode takes whole IQ samples (an even number of bytes)");
    if (((uintptr_t) d_in % 16) || ((uintptr_t) d_out % ((c->cfg.flags & FMD_FLAG_S16) ? 2 : 4)))
        return fmd_set_error(FMD_EINVAL, "quality mode needs a 16-byte aligned input and an aligned output");
    size_t left = nbytes / 2;
    const size_t step = (size_t) 1 << 30; /* a multiple of the tile: later groups stay 16-byte aligned */
    while (left) {
        const unsigned n = (unsigned) (left < step ? left : step);
        size_t nout = 0;
        int rc = quality_launch(c, d_in, n, d_out, s, &nout);
        if (rc) return rc;
        d_in += 2ull * n;
        if (c->cfg.flags & FMD_FLAG_S16) d_out = reinterpret_cast<float *>(reinterpret_cast<short *>(d_out) + nout);
        else d_out += nout;
        left -= n;
    }
    return FMD_OK;
}

int fmq_reset(fmd_ctx *c) {
    FmdDeviceScope on_device(c->device);
    FMD_CUDA(cudaDeviceSynchronize());
    FMD_CUDA(cudaMemset(c->q->d_state, 0, sizeof(FmqState)));
    c->q->pos = 0;
    return FMD_OK;
}

/* ---- one stream over several GPUs: independent segments with a warm-up prefix ---- */

/* Start a fresh segment at global sample index `pos`: zero filter state (asynchronously, on the
 * stream the segment's kernels will run on), the decimation phase follows the global index. */
int fmq_seek(fmd_ctx *c, unsigned long long pos, cudaStream_t s) {
    FMD_CUDA(cudaMemsetAsync(c->q->d_state, 0, offsetof(FmqState, error), s));
    c->q->pos = pos;
    return FMD_OK;
}

size_t fmq_out_count_at(const fmd_ctx *c, unsigned long long pos, size_t nsamp) {
    return quality_out_count(pos, nsamp, c->q->decim);
}

/* Samples after which a cascade started from a zero state agrees with the running stream to far
 * below the 1e-5 the pipeline is held to: the slowest pole radius rho of each cascade gives a decay
 * rho^n (times a polynomial in n for the cascade's nearly coinciding poles, hence the margin); the
 * output cascade only starts converging once the input cascade has, so the two lengths add. */
extern "C" uint64_t fmd_warmup_for_sections(const float *sos6, unsigned rows) {
    double rho = 0.0;
    for (unsigned r = 0; sos6 && r < rows; ++r) {
        const double a1 = sos6[6 * r + 4], a2 = sos6[6 * r + 5];
        const double rr = a2 != 0.0 ? sqrt(fabs(a2)) : fabs(a1);
        if (rr > rho) rho = rr;
    }
    if (!(rho > 0.0)) return 64u * rows;
    if (rho >= 0.99999) rho = 0.99999;
    return (uint64_t) ceil(1.5 * (log(1e-9) / log(rho)) + 64.0 * rows);
}

unsigned long long fmq_warmup_samples(const fmd_ctx *c) {
    unsigned long long n = fmd_warmup_for_sections(&c->sec.out[0][0], c->lout);
    if (c->lin) n += fmd_warmup_for_sections(&c->sec.in[0][0], c->lin);
    return (n + 7ull) & ~7ull; /* 16 bytes of IQ: the payload after it stays aligned */
}

size_t fmq_out_count(const fmd_ctx *c, size_t nsamp) { return quality_out_count(c->q->pos, nsamp, c->q->decim); }

size_t fmq_max_out_count(const fmd_ctx *c, size_t nsamp) { return quality_out_count(0, nsamp, c->q->decim); }

int fmq_check_error(fmd_ctx *c) {
    if (!c->q->lin) return FMD_OK; /* only the input cascade's look-back can raise it */
    /* the calling thread's current device may be another one (multi-GPU hosts) */
    int prev = -1;
    FMD_CUDA(cudaGetDevice(&prev));
    if (prev != c->device) FMD_CUDA(cudaSetDevice(c->device));
    unsigned err = 0;
    cudaError_t e = cudaMemcpy(&err, &c->q->d_state->error, sizeof err, cudaMemcpyDeviceToHost);
    /* report once: the flag is cleared so that the calls after an fmd_reset (or a retry) start clean */
    if (e == cudaSuccess && err) e = cudaMemset(&c->q->d_state->error, 0, sizeof err);
    if (prev != c->device && prev >= 0) cudaSetDevice(prev);
    if (e != cudaSuccess) return fmd_set_error(FMD_ECUDA, "quality pipeline: reading the error flag failed: %s", cudaGetErrorString(e));
    if (err)
        return fmd_set_error(FMD_ECUDA, "quality pipeline: a tile gave up waiting for its predecessor's carry "
                                        "(the outputs of that call are invalid; fmd_reset and resubmit)");
    return FMD_OK;
}

/* ---------------------------------------------------------------- -q1 (correctIq) as a scan */

struct fmq1_ctx {
    fmq_design des;       /* D = 1: A = [1 - 2 esr] */
    fmq_consts c, c_tile;
    fmq_table_sizes sz;
    float *d_tab = nullptr; /* lane | apow | ck | ckn | lane_tile (ck, ckn unused) */
    float *d_off = nullptr; /* the two running offsets: the carry between calls */
    float *d_e = nullptr, *d_q = nullptr; /* per-tile end values / start values */
    unsigned tile_cap = 0;
    fmq_consts c_run;             /* fused form: the scan over the CTAs' runs (unit FMQ1_SUBS tiles) */
    float *d_lane_run = nullptr;  /* (a^run)^l, l = 0..31 */
    float *d_sbuf = nullptr; /* conditioned samples of the current call (three-pass form) */
    size_t sbuf_cap = 0;
    float2 *d_heads = nullptr; /* fused form: the conditioned first FMQ1_HEADS complex samples of every chunk */
    size_t heads_cap = 0;      /* chunks */

    float esr = 0.f, a = 0.f;
};

int fmq1_setup(fmd_ctx *c) {
    fmq1_ctx *q = new (std::nothrow) fmq1_ctx();
    if (!q) return fmd_set_error(FMD_ENOMEM, "out of host memory");
    c->q1 = q;
    q->esr = (float) (50. / c->cfg.sample_rate); /* matrix.c:210 */
    const double a = 1.0 - 2.0 * (double) q->esr;
    q->a = (float) a;
    memset(&q->des, 0, sizeof q->des);
    q->des.L = 1;
    q->des.D = 1;
    q->des.A[0][0] = (long double) a;
    q->des.C[0] = 1.0L;
    fmq_build_consts(&q->des, FMQ1_T, FMQ1_TILE, &q->c);
    fmq_build_consts(&q->des, FMQ1_TILE, FMQ1_TILE, &q->c_tile);
    fmq_table_layout(&q->des, FMQ1_T, &q->sz);
    const size_t n = (size_t) q->sz.lane + q->sz.apow + q->sz.ck + q->sz.ckn + q->sz.lane_tile;
    float *h = (float *) malloc(n * sizeof(float));
    if (!h) return fmd_set_error(FMD_ENOMEM, "out of host memory");
    fmq_build_tables(&q->des, FMQ1_T, h, h + q->sz.lane, h + q->sz.lane + q->sz.apow, h + q->sz.lane + q->sz.apow + q->sz.ck,
                     h + q->sz.lane + q->sz.apow + q->sz.ck + q->sz.ckn);
    cudaError_t e = cudaMalloc(&q->d_tab, n * sizeof(float));
    if (e == cudaSuccess) e = cudaMemcpy(q->d_tab, h, n * sizeof(float), cudaMemcpyHostToDevice);
    free(h);
    if (e != cudaSuccess) return fmd_set_error(FMD_ECUDA, "-q1 table upload failed: %s", cudaGetErrorString(e));
    FMD_CUDA(cudaMalloc(&q->d_off, 2 * sizeof(float)));
    FMD_CUDA(cudaMemset(q->d_off, 0, 2 * sizeof(float))); /* static REAL off[2] = {} */
    {
        const unsigned long long run = (unsigned long long) FMQ1_SUBS * FMQ1_TILE;
        fmq_build_consts(&q->des, (unsigned) run, (unsigned) run, &q->c_run);
        float lane_run[32];
        for (unsigned l = 0; l < 32; ++l) lane_run[l] = (float) powl((long double) a, (long double) (run * l));
        FMD_CUDA(cudaMalloc(&q->d_lane_run, sizeof lane_run));
        FMD_CUDA(cudaMemcpy(q->d_lane_run, lane_run, sizeof lane_run, cudaMemcpyHostToDevice));
    }
    return FMD_OK;
}

void fmq1_free(fmd_ctx *c) {
    fmq1_ctx *q = c->q1;
    if (!q) return;
    if (q->d_tab) cudaFree(q->d_tab);
    if (q->d_off) cudaFree(q->d_off);
    if (q->d_e) cudaFree(q->d_e);
    if (q->d_q) cudaFree(q->d_q);
    if (q->d_sbuf) cudaFree(q->d_sbuf);
    if (q->d_heads) cudaFree(q->d_heads);
    if (q->d_lane_run) cudaFree(q->d_lane_run);
    delete q;
    c->q1 = nullptr;
}

int fmq1_reset(fmd_ctx *c) {
    FmdDeviceScope on_device(c->device);
    FMD_CUDA(cudaDeviceSynchronize());
    FMD_CUDA(cudaMemset(c->q1->d_off, 0, 2 * sizeof(float)));
    return FMD_OK;
}

/* the offsets at the start of every tile of this call (q->d_q): per-tile end values from zero, then
 * the scan over the tiles with the carry of the previous call in and the next one's out */
static int fmq1_scan(fmd_ctx *c, const uint8_t *d_in, unsigned long long nchunks, cudaStream_t s, FmdQ1Geom *gout,
                     unsigned *ntiles_out) {
    fmq1_ctx *q = c->q1;
    const unsigned long long B = c->cfg.chunk_bytes;
    FmdQ1Geom g;
    g.chunk_bytes = B;
    g.steps_per_chunk = (unsigned) (B >> 2);
    g.nsteps = nchunks * g.steps_per_chunk;
    g.esr = q->esr;
    g.a = q->a;
    const unsigned long long ntiles64 = (g.nsteps + FMQ1_TILE - 1) / FMQ1_TILE;
    if (ntiles64 > 0x7fffffffull) return fmd_set_error(FMD_EINVAL, "too many chunks in one -q1 call");
    const unsigned ntiles = (unsigned) ntiles64;
    if (ntiles > q->tile_cap) {
        if (q->d_e) cudaFree(q->d_e);
        if (q->d_q) cudaFree(q->d_q);
        q->d_e = q->d_q = nullptr;
        q->tile_cap = 0;
        const unsigned cap = ntiles < 4096u ? 4096u : ntiles;
        FMD_CUDA(cudaMalloc(&q->d_e, (size_t) cap * 2 * sizeof(float)));
        FMD_CUDA(cudaMalloc(&q->d_q, (size_t) cap * 2 * sizeof(float)));
        q->tile_cap = cap;
    }
    const float *lane = q->d_tab, *apow = q->d_tab + q->sz.lane;
    const float *lane_tile = q->d_tab + q->sz.lane + q->sz.apow + q->sz.ck + q->sz.ckn;
    const unsigned nwarps = FMQ_SCAN_THREADS / 32;
    const unsigned runlen = ((ntiles + nwarps - 1u) / nwarps + 31u) & ~31u;
    const unsigned long long len_last = g.nsteps - (unsigned long long) (ntiles - 1u) * FMQ1_TILE;
    FmqMat prun, pl;
    memset(&prun, 0, sizeof prun);
    memset(&pl, 0, sizeof pl);
    fmq_matpow_f(&q->des, (unsigned long long) FMQ1_TILE * runlen, prun.m);
    fmq_matpow_f(&q->des, len_last, pl.m);
    fmd_correctiq_kernel<false><<<ntiles, FMQ_THREADS, 0, s>>>(d_in, nullptr, g, q->c, lane, apow, nullptr, q->d_e);
    fmq_scan_tiles_kernel<1, 2><<<1, FMQ_SCAN_THREADS, 0, s>>>(q->d_e, q->d_q, q->d_off, q->c_tile, lane_tile, prun, pl, ntiles,
                                                                runlen, nullptr, nullptr);
    FMD_CUDA(cudaGetLastError());
    c->launches += 2;
    *gout = g;
    *ntiles_out = ntiles;
    return FMD_OK;
}

/* three-pass form: conditioned floats (one per input byte) for the float-input streaming kernel */
int fmq1_condition(fmd_ctx *c, const uint8_t *d_in, unsigned long long nchunks, cudaStream_t s, const float **sbuf) {
    fmq1_ctx *q = c->q1;
    const size_t need = (size_t) (nchunks * c->cfg.chunk_bytes); /* floats: one per input byte */
    if (need > q->sbuf_cap) {
        if (q->d_sbuf) cudaFree(q->d_sbuf);
        q->d_sbuf = nullptr;
        q->sbuf_cap = 0;
        FMD_CUDA(cudaMalloc(&q->d_sbuf, need * sizeof(float)));
        q->sbuf_cap = need;
    }
    FmdQ1Geom g;
    unsigned ntiles = 0;
    int rc = fmq1_scan(c, d_in, nchunks, s, &g, &ntiles);
    if (rc) return rc;
    const float *lane = q->d_tab, *apow = q->d_tab + q->sz.lane;
    fmd_correctiq_kernel<true><<<ntiles, FMQ_THREADS, 0, s>>>(d_in, q->d_sbuf, g, q->c, lane, apow, q->d_q, nullptr);
    FMD_CUDA(cudaGetLastError());
    c->launches += 1;
    *sbuf = q->d_sbuf;
    return FMD_OK;
}

/* ---- fused form (fmd_q1_fused.cuh): conditioning and demodulation in one kernel ---- */
typedef void (*fmq1_fused_fn)(const uint8_t *, float *, const FmdQ1Geom, const fmq_consts, const float *, const float *,
                              const unsigned, const FmdStencil, const unsigned, float2 *);

template <int NTI>
static fmq1_fused_fn fmq1_pick_out(unsigned nto) {
    switch (nto) {
        case 1: return fmd_q1_fused_kernel<NTI, 1>;
        case 2: return fmd_q1_fused_kernel<NTI, 2>;
        case 3: return fmd_q1_fused_kernel<NTI, 3>;
        case 4: return fmd_q1_fused_kernel<NTI, 4>;
        case 5: return fmd_q1_fused_kernel<NTI, 5>;
        default: return nullptr;
    }
}

static fmq1_fused_fn fmq1_pick(unsigned nti, unsigned nto) {
    switch (nti) {
        case 0: return fmq1_pick_out<0>(nto);
        case 1: return fmq1_pick_out<1>(nto);
        case 2: return fmq1_pick_out<2>(nto);
        case 3: return fmq1_pick_out<3>(nto);
        default: return nullptr;
    }
}

/* *done = false: this design / chunk size has no fused form (the caller takes the three-pass one).
 * On success d_out holds every output except the chunk heads [0, c->head), and *heads the conditioned
 * first FMQ1_HEADS complex samples of every chunk for the op-for-op path that produces those. */
int fmq1_fused(fmd_ctx *c, const uint8_t *d_in, unsigned long long nchunks, float *d_out, cudaStream_t s,
               const float **heads, unsigned *heads_stride_bytes, bool *done) {
    fmq1_ctx *q = c->q1;
    *done = false;
    const unsigned long long B = c->cfg.chunk_bytes;
    fmq1_fused_fn fn = fmq1_pick(c->nti, c->nto);
    if (const char *e = getenv("FMD_Q1_FUSED")) {
        if (!atoi(e)) fn = nullptr; /* audits: the three-pass form */
    }
    if (!fn || !c->fast_in_ok || (c->cfg.flags & FMD_FLAG_EXACT) || (B >> 2) % FMQ1_TILE != 0 || c->head > FMQ1_HEADS / 4u ||
        nchunks > 0x7fffffffull)
        return FMD_OK;
    if (nchunks > q->heads_cap) {
        if (q->d_heads) cudaFree(q->d_heads);
        q->d_heads = nullptr;
        q->heads_cap = 0;
        const size_t cap = nchunks < 1024 ? 1024 : (size_t) nchunks;
        FMD_CUDA(cudaMalloc(&q->d_heads, cap * FMQ1_HEADS * sizeof(float2)));
        q->heads_cap = cap;
    }
    FmdQ1Geom g;
    g.chunk_bytes = B;
    g.steps_per_chunk = (unsigned) (B >> 2);
    g.nsteps = nchunks * g.steps_per_chunk;
    g.esr = q->esr;
    g.a = q->a;
    const unsigned long long ntiles64 = g.nsteps / FMQ1_TILE;
    if (ntiles64 > 0x7fffffffull) return fmd_set_error(FMD_EINVAL, "too many chunks in one -q1 call");
    const unsigned ntiles = (unsigned) ntiles64;
    const unsigned nruns = (ntiles + FMQ1_SUBS - 1u) / FMQ1_SUBS; /* one CTA walks a run of FMQ1_SUBS tiles */
    if (nruns > q->tile_cap) {
        if (q->d_e) cudaFree(q->d_e);
        if (q->d_q) cudaFree(q->d_q);
        q->d_e = q->d_q = nullptr;
        q->tile_cap = 0;
        const unsigned cap = nruns < 4096u ? 4096u : nruns;
        FMD_CUDA(cudaMalloc(&q->d_e, (size_t) cap * 2 * sizeof(float)));
        FMD_CUDA(cudaMalloc(&q->d_q, (size_t) cap * 2 * sizeof(float)));
        q->tile_cap = cap;
    }
    /* end values of the offsets over every run from zero, chained by the scan over the runs (carry of the
     * previous call in, the next one's out) */
    const unsigned long long run = (unsigned long long) FMQ1_SUBS * FMQ1_TILE;
    const unsigned nwarps = FMQ_SCAN_THREADS / 32;
    const unsigned runlen = ((nruns + nwarps - 1u) / nwarps + 31u) & ~31u;
    const unsigned long long len_last = g.nsteps - (unsigned long long) (nruns - 1u) * run;
    FmqMat prun, pl;
    memset(&prun, 0, sizeof prun);
    memset(&pl, 0, sizeof pl);
    fmq_matpow_f(&q->des, run * runlen, prun.m);
    fmq_matpow_f(&q->des, len_last, pl.m);
    fmd_q1_ends_kernel<<<nruns, FMQ_THREADS, 0, s>>>(d_in, g, q->c, q->d_tab, ntiles, q->d_e);
    fmq_scan_tiles_kernel<1, 2><<<1, FMQ_SCAN_THREADS, 0, s>>>(q->d_e, q->d_q, q->d_off, q->c_run, q->d_lane_run, prun, pl, nruns,
                                                                runlen, nullptr, nullptr);
    FMD_CUDA(cudaGetLastError());
    c->launches += 2;
    int rc;
    cudaEvent_t *pe = nullptr;
    rc = fmd_prof_events(c, &pe);
    if (rc) return rc;
    if (pe) FMD_CUDA(cudaEventRecord(pe[0], s));
    fn<<<nruns, FMQ_THREADS, 0, s>>>(d_in, d_out, g, q->c, q->d_tab, q->d_q, ntiles, c->st, c->head, q->d_heads);
    FMD_CUDA(cudaGetLastError());
    if (pe) FMD_CUDA(cudaEventRecord(pe[1], s));
    c->launches += 1;
    *heads = reinterpret_cast<const float *>(q->d_heads);
    *heads_stride_bytes = FMQ1_HEADS * (unsigned) sizeof(float2);
    *done = true;
    return FMD_OK;
}
