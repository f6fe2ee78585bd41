/*
 * fmdemod.cu -- the C ABI of include/fmdemod.h over the sm_100a kernels.
 *
 * What the reference does on its worker thread per 256 KiB chunk
 * (/root/reference/src/matrix.c:236-273) this library does for an arbitrary
 * number of chunks per launch; what the reference does with a 1-slot buffer, a
 * mutex and two semaphores (src/main.c:72-91) is here a ring of pinned slots
 * chained through three CUDA streams (copy-in, compute, copy-out) by events.
 *
 * There is no CPU implementation of the data path in this library.
 */
#include <cuda_runtime.h>

#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>

#define FMD_WITH_EXACT_KERNELS
#include "fmd_ctx.h"
#include "fmd_slices.h"
#include "fmd_kernels.cuh"

/* ------------------------------------------------------------------ errors */

static thread_local char g_err[512] = "";

extern "C" int fmd_set_error(int code, const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
    return code;
}

extern "C" const char *fmd_last_error(void) { return g_err; }

/* ----------------------------------------------------------------- context */

static int prof_flush(fmd_ctx *c) {
    for (unsigned i = 0; i < c->pev_used; ++i) {
        float ms = 0.f;
        FMD_CUDA(cudaEventSynchronize(c->pev[2 * i + 1]));
        FMD_CUDA(cudaEventElapsedTime(&ms, c->pev[2 * i], c->pev[2 * i + 1]));
        c->prof_ms += ms;
        c->prof_n++;
    }
    c->pev_used = 0;
    return FMD_OK;
}

int fmd_prof_events(fmd_ctx *c, cudaEvent_t **pe) {
    *pe = nullptr;
    if (!c->profile) return FMD_OK;
    if (c->pev_used == c->pev_cap) {
        int rc = prof_flush(c);
        if (rc) return rc;
    }
    *pe = &c->pev[2 * c->pev_used++];
    return FMD_OK;
}

extern "C" int fmd_profile_kernels(fmd_ctx *c, int enable) {
    if (!c) return fmd_set_error(FMD_EINVAL, "null argument");
    FmdDeviceScope on_device(c->device); /* events belong to the device they are created on */
    if (enable && !c->pev) {
        c->pev_cap = 256;
        c->pev = new (std::nothrow) cudaEvent_t[2 * c->pev_cap];
        if (!c->pev) return fmd_set_error(FMD_ENOMEM, "out of host memory");
        for (unsigned i = 0; i < 2 * c->pev_cap; ++i) FMD_CUDA(cudaEventCreate(&c->pev[i]));
    }
    if (!enable && c->profile) prof_flush(c);
    c->profile = enable != 0;
    return FMD_OK;
}

extern "C" int fmd_kernel_time(fmd_ctx *c, double *ms, uint64_t *launches) {
    if (!c || !ms || !launches) return fmd_set_error(FMD_EINVAL, "null argument");
    int rc = prof_flush(c);
    if (rc) return rc;
    *ms = c->prof_ms;
    *launches = c->prof_n;
    c->prof_ms = 0.0;
    c->prof_n = 0;
    return FMD_OK;
}

extern "C" void fmd_config_default(fmd_config *cfg) {
    if (!cfg) return;
    memset(cfg, 0, sizeof *cfg);
    cfg->sample_rate = 125000.0;
    cfg->lowpass_in = 0.0;
    cfg->lowpass_out = 12500.0;
    cfg->epsilon = 0.3;
    cfg->in_degree = 0;
    cfg->out_degree = 3;
    cfg->mode = 0x10;
    cfg->flags = 0;
    cfg->chunk_bytes = FMD_DEFAULT_CHUNK_BYTES;
}

/* FMD_VERBOSE=1: what the reference prints to stderr when built with -DIS_VERBOSE (filter
 * parameters and section rows, src/matrix.c:39-46, src/filter.c:193-204), plus the folded stencils
 * the streaming kernel runs */
static void fmd_dump_design(const fmd_ctx *c, unsigned conditioner) {
    const fmd_config &g = c->cfg;
    const bool quality = (g.flags & FMD_FLAG_QUALITY) != 0;
    const double fs = quality ? g.input_rate : g.sample_rate;
    auto rows = [](const char *what, const float (*sos)[6], unsigned n) {
        for (unsigned i = 0; i < n; ++i)
            fprintf(stderr, "  %s sos[%u] = {%.9g, %.9g, %.9g, %.9g, %.9g, %.9g}\n", what, i, sos[i][0], sos[i][1],
                    sos[i][2], sos[i][3], sos[i][4], sos[i][5]);
    };
    auto taps = [](const char *what, const float *t, unsigned n, float dc) {
        fprintf(stderr, "  %s stencil:", what);
        for (unsigned i = 0; i < n; ++i) fprintf(stderr, " %.9g", t[i]);
        fprintf(stderr, "  dc %.9g\n", dc);
    };
    fprintf(stderr, "fmdemod: design rate %.1f Hz, chunk %llu bytes, conditioner -q%u%s\n", fs,
            (unsigned long long) g.chunk_bytes, conditioner, quality ? ", quality pipeline" : "");
    fprintf(stderr, " output lowpass: %s, cutoff %.1f Hz, w = %.6f, degree %u (%u sections)\n",
            (g.mode & 1u) ? "Chebyshev type 1" : "Butterworth", g.lowpass_out, M_PI * g.lowpass_out / fs, g.out_degree,
            c->lout);
    if (g.mode & 1u) fprintf(stderr, "  ripple parameter %.3f\n", g.epsilon);
    rows("out", c->sec.out, c->lout);
    if (!quality) taps("out", c->st.tout, c->nto, c->st.dco);
    if (c->lin) {
        fprintf(stderr, " input lowpass: %s, cutoff %.1f Hz, w = %.6f, degree %u (%u sections)\n",
                (g.mode & 2u) ? "Chebyshev type 1" : "Butterworth", g.lowpass_in, M_PI * g.lowpass_in / fs, g.in_degree,
                c->lin);
        rows("in ", c->sec.in, c->lin);
        if (!quality)
            taps(c->fast_in_ok ? "in  (divided by its first tap; dc is the Q lane's)" : "in  (op-for-op path only)",
                 c->st.tin, c->nti, c->st.dcq);
    }
    if (conditioner == 2u) rows("-q2 highpass", c->sec.hp, 2);
}

extern "C" int fmd_device_count(void) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) return fmd_set_error(FMD_ENODEV, "cudaGetDeviceCount: %s", cudaGetErrorString(e));
    return n;
}

extern "C" int fmd_create(const fmd_config *cfg_in, int device, fmd_ctx **out) {
    if (!cfg_in || !out) return fmd_set_error(FMD_EINVAL, "fmd_create: null argument");
    *out = nullptr;
    fmd_config cfg = *cfg_in;
    /* the reference's defaulting rule: -L without -D inherits the output degree (matrix.c:190-192) */
    if (cfg.lowpass_in != 0.0 && cfg.in_degree == 0) cfg.in_degree = cfg.out_degree;
    if (cfg.lowpass_in == 0.0 && cfg.in_degree != 0)
        return fmd_set_error(FMD_EINVAL, "-D without -L: the reference applies an undesigned filter (undefined)");
    if (cfg.out_degree == 0 || fmd_sections(cfg.out_degree) > FMD_MAX_SECTIONS ||
        fmd_sections(cfg.in_degree) > FMD_MAX_SECTIONS)
        return fmd_set_error(FMD_EINVAL, "filter degree outside 1..%d", 2 * FMD_MAX_SECTIONS);
    if (cfg.chunk_bytes == 0 || (cfg.chunk_bytes & 3u) || cfg.chunk_bytes > (1ull << 32))
        return fmd_set_error(FMD_EINVAL, "chunk_bytes must be a positive multiple of 4, at most 4 GiB (got %llu)",
                             (unsigned long long) cfg.chunk_bytes);
    if (cfg.in_degree && cfg.chunk_bytes / 2 < 2ull * fmd_sections(cfg.in_degree) + 2ull)
        return fmd_set_error(FMD_EINVAL,
                             "chunk of %llu bytes is shorter than the head of a degree-%u input filter: in the "
                             "reference the result then depends on which overrun it reads first",
                             (unsigned long long) cfg.chunk_bytes, cfg.in_degree);
    const unsigned conditioner = (cfg.mode >> 2) & 3u;
    const bool quality = (cfg.flags & FMD_FLAG_QUALITY) != 0;
    if ((cfg.flags & FMD_FLAG_S16) && !quality)
        return fmd_set_error(FMD_EINVAL, "FMD_FLAG_S16 is an egress of the quality pipeline (FMD_FLAG_QUALITY) only");
    unsigned q_decim = 1;
    if (quality) {
        if (cfg.flags & FMD_FLAG_EXACT)
            return fmd_set_error(FMD_EINVAL, "FMD_FLAG_EXACT audits the reference-compatible path; not with FMD_FLAG_QUALITY");
        if (conditioner != 0u) return fmd_set_error(FMD_EINVAL, "quality mode centres the samples itself: -q0 only");
        if (!(cfg.input_rate > 0.0) || !(cfg.sample_rate > 0.0))
            return fmd_set_error(FMD_EINVAL, "quality mode needs the input rate (-r) and the output rate (-S)");
        const double ratio = cfg.input_rate / cfg.sample_rate, m = floor(ratio + 0.5);
        if (m < 1.0 || m > 65536.0 || fabs(ratio - m) > 1e-9 * ratio)
            return fmd_set_error(FMD_EINVAL, "input rate / output rate = %.9g is not an integer in 1..65536", ratio);
        q_decim = (unsigned) m;
        if (fmd_sections(cfg.out_degree) > FMQ_MAXL || fmd_sections(cfg.in_degree) > FMQ_MAXL)
            return fmd_set_error(FMD_EINVAL, "quality mode takes filter degrees up to %d", 2 * FMQ_MAXL);
        const double lo = cfg.input_rate / 1000.0, hi = cfg.input_rate / 2.0;
        if (!(cfg.lowpass_out >= lo && cfg.lowpass_out < hi) ||
            (cfg.lowpass_in != 0.0 && !(cfg.lowpass_in >= lo && cfg.lowpass_in < hi)))
            return fmd_set_error(FMD_EINVAL,
                                 "quality mode keeps the filter state in float32: cut-offs must lie in "
                                 "[input_rate/1000, input_rate/2)");
    }
    if (conditioner == 1u && (quality || (cfg.chunk_bytes % 32u) != 0)) /* 8 steps (4 bytes each) per thread, fmd_correctiq.cuh */
        return fmd_set_error(FMD_EINVAL, "-q1 (correctIq) needs chunks that are a multiple of 32 bytes (and no -Q)");
    if (conditioner == 2u && cfg.chunk_bytes * sizeof(float) < (128u << 10))
        return fmd_set_error(FMD_EINVAL,
                             "-q2 with chunks under 32 KiB: the reference's highpassDc reads past a scratch buffer "
                             "that is then a recycled heap chunk (undefined); with larger chunks its slack is zero");
    if (((cfg.mode >> 4) & 3u) == 0)
        return fmd_set_error(FMD_EINVAL, "demodulation off (-c) is unreachable in the reference CLI and unsupported");

    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev <= 0)
        return fmd_set_error(FMD_ENODEV, "no CUDA device: %s (this library has no CPU path)",
                             e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
    if (device < 0 || device >= ndev) return fmd_set_error(FMD_EINVAL, "device %d not in 0..%d", device, ndev - 1);
    FmdDeviceScope on_device(device);
    cudaDeviceProp prop;
    FMD_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10)
        return fmd_set_error(FMD_ENODEV, "device %d is sm_%d%d; this library ships sm_100a code only", device,
                             prop.major, prop.minor);

    fmd_ctx *c = new (std::nothrow) fmd_ctx();
    if (!c) return fmd_set_error(FMD_ENOMEM, "out of host memory");
    c->cfg = cfg;
    c->device = device;
    c->sm_count = prop.multiProcessorCount;
    c->lout = fmd_sections(cfg.out_degree);
    c->lin = fmd_sections(cfg.in_degree);
    memset(&c->sec, 0, sizeof c->sec);
    memset(&c->st, 0, sizeof c->st);
    c->sec.lin = (int) c->lin;
    c->sec.lout = (int) c->lout;
    c->sec.conditioner = (int) conditioner;
    /* -q3 (normalizeInput) streams with or without the input filter (the kernel clears the halo
     * samples where they are the zero extension past a chunk: no byte normalises to 0.0).
     * -q2 (highpassDc) is one more complex sweep: without -L it IS the input stencil of the streaming
     * kernel; under -L the two stencils are composed into one (below).  The highpass sweep writes
     * three complex values past its output region, which the input filter then reads: the first two
     * are fully swept -- exactly what the composed stencil gives on the zero extension -- and only
     * the third is partial, so the composition is used while the input stencil has at most 3 taps
     * (degree <= 8: it then never reaches the third); wider ones stay on the op-for-op path. */
    /* -q1 (correctIq): the samples are conditioned by a scan of their own (fmd_correctiq.cuh); the
     * conditioned floats then stream through the float-input kernel (up to 3 input taps) */

    /* quality mode runs both filters at the input rate (the output one BEFORE the decimation) */
    const double design_rate = quality ? cfg.input_rate : cfg.sample_rate;
    int rc = fmd_design_sos(cfg.mode & 1u, cfg.out_degree, cfg.lowpass_out, design_rate, cfg.epsilon,
                            &c->sec.out[0][0]);
    if (!rc) rc = fmd_fold_taps(&c->sec.out[0][0], c->lout, 1, c->st.tout, &c->st.dco, &c->nto);
    if (!rc && c->lin) {
        float dci = 0.f;
        unsigned nti_q = 0;
        float taps_q[FMD_MAX_TAPS];
        rc = fmd_design_sos((cfg.mode >> 1) & 1u, cfg.in_degree, cfg.lowpass_in, design_rate, cfg.epsilon,
                            &c->sec.in[0][0]);
        if (!rc) rc = fmd_fold_taps(&c->sec.in[0][0], c->lin, 1, c->st.tin, &dci, &c->nti);
        /* the Q lane lacks the "+1" (filter.c:255): same taps, different constant */
        if (!rc) rc = fmd_fold_taps(&c->sec.in[0][0], c->lin, 0, taps_q, &c->st.dcq, &nti_q);
        c->st.dci = dci;
        /* the streaming kernel applies the input stencil divided by its first tap (the factor
         * cancels in atan2 of s0*conj(s1)); designs where that is impossible use the exact path */
        const float c0 = c->st.tin[0];
        if (!rc && c->nti >= 1 && std::isfinite(c0) && c0 != 0.0f && dci == 0.0f && std::isfinite(c->st.dcq)) {
            for (unsigned t = 1; t < c->nti; ++t) c->st.tin[t] = (float) ((double) c->st.tin[t] / (double) c0);
            c->st.dcq = (float) ((double) c->st.dcq / (double) c0);
            c->st.tin[0] = 1.0f;
        } else if (!rc) {
            c->fast_in_ok = false;
        }
    }
    if (!rc && conditioner == 2u) {
        /* matrix.c:151: processFilterOption(2, 3, sos, 1., samplingRate, 0.) */
        rc = fmd_design_sos(FMD_BUTTER_HP, 3, 1.0, cfg.sample_rate, 0.0, &c->sec.hp[0][0]);
        if (!rc && c->lin && c->fast_in_ok) {
            /* highpass, then lowpass: z[c] = dL + sum_t L[t] x[c+t] with x[c] = dH + sum_u H[u] s[c+u]
             * is one stencil with taps L*H and constant dL + dH sum L (per lane; both I-lane
             * constants are 0).  c->st.tin / dcq hold the lowpass stencil divided by its first tap
             * at this point; the highpass one is folded here the same way. */
            float hp[FMD_MAX_TAPS], hq[FMD_MAX_TAPS], dhi = 0.f, dhq = 0.f;
            unsigned nh = 0, nhq = 0;
            rc = fmd_fold_taps(&c->sec.hp[0][0], 2, 1, hp, &dhi, &nh);
            if (!rc) rc = fmd_fold_taps(&c->sec.hp[0][0], 2, 0, hq, &dhq, &nhq);
            const unsigned nl = c->nti, ncomp = nl + nh - 1u;
            if (!rc && nl <= 3u && ncomp <= 5u && std::isfinite(hp[0]) && hp[0] != 0.0f && dhi == 0.0f && std::isfinite(dhq)) {
                double comp[2 * FMD_MAX_TAPS] = {0.0}, suml = 0.0;
                for (unsigned t = 0; t < nl; ++t) {
                    suml += (double) c->st.tin[t];
                    for (unsigned u = 0; u < nh; ++u) comp[t + u] += (double) c->st.tin[t] * (double) hp[u];
                }
                const double dq = (double) c->st.dcq + (double) dhq * suml, c0 = comp[0];
                for (unsigned t = 0; t < ncomp; ++t) c->st.tin[t] = (float) (comp[t] / c0);
                for (unsigned t = ncomp; t < FMD_MAX_TAPS; ++t) c->st.tin[t] = 0.f;
                c->st.dcq = (float) (dq / c0);
                c->st.tin[0] = 1.0f;
                c->nti = ncomp;
                if (!(c0 > 0.0) && !(c0 < 0.0)) c->fast_in_ok = false;
            } else if (!rc) {
                c->fast_in_ok = false;
            }
        } else if (!rc && c->lin) {
            c->fast_in_ok = false;
        }
        if (!rc && !c->lin) {
            float dci = 0.f, taps_q[FMD_MAX_TAPS];
            unsigned nti_q = 0;
            rc = fmd_fold_taps(&c->sec.hp[0][0], 2, 1, c->st.tin, &dci, &c->nti);
            if (!rc) rc = fmd_fold_taps(&c->sec.hp[0][0], 2, 0, taps_q, &c->st.dcq, &nti_q);
            c->st.dci = dci;
            const float c0 = c->st.tin[0];
            if (!rc && c->nti >= 1 && std::isfinite(c0) && c0 != 0.0f && dci == 0.0f && std::isfinite(c->st.dcq)) {
                for (unsigned t = 1; t < c->nti; ++t) c->st.tin[t] = (float) ((double) c->st.tin[t] / (double) c0);
                c->st.dcq = (float) ((double) c->st.dcq / (double) c0);
                c->st.tin[0] = 1.0f;
            } else if (!rc) {
                c->fast_in_ok = false;
            }
        }
    }
    if (rc) {
        delete c;
        return rc;
    }
    if ((conditioner == 3u || conditioner == 1u) && c->nti > 3u) c->fast_in_ok = false; /* -q3 / -q1 under -L stream up to 3 input taps (degree <= 8) */
    {
        /* the discriminator's error reaches the output multiplied by sum|tout|: keep it near 2e-5 */
        double gain = 0.0;
        for (unsigned t = 0; t < c->nto; ++t) gain += fabs((double) c->st.tout[t]);
        c->atan_terms = (gain * 1.75e-6 <= 2.0e-5) ? 6u : 8u;
        if (const char *e = getenv("FMD_ATAN_TERMS")) c->atan_terms = (unsigned) atoi(e); /* audits */
    }
    if (const char *e = getenv("FMD_VERBOSE")) {
        if (atoi(e)) fmd_dump_design(c, conditioner);
    }
    c->head = 2 * c->lout - 1;
    if (c->lin && 4 * c->lin - 2 > c->head) c->head = 4 * c->lin - 2;
    if (conditioner == 2u && 4 * 2 - 2 > c->head) c->head = 4 * 2 - 2; /* the highpass sweep's spill */
    c->head = (c->head + 7u) & ~7u;

    if (cudaStreamCreateWithFlags(&c->s_in, cudaStreamNonBlocking) != cudaSuccess ||
        cudaStreamCreateWithFlags(&c->s_k, cudaStreamNonBlocking) != cudaSuccess ||
        cudaStreamCreateWithFlags(&c->s_out, cudaStreamNonBlocking) != cudaSuccess) {
        fmd_destroy(c);
        return fmd_set_error(FMD_ECUDA, "cudaStreamCreate failed: %s", cudaGetErrorString(cudaGetLastError()));
    }
    if (cudaMalloc(&c->d_sec, sizeof(FmdSections)) != cudaSuccess ||
        cudaMemcpy(c->d_sec, &c->sec, sizeof(FmdSections), cudaMemcpyHostToDevice) != cudaSuccess) {
        fmd_destroy(c);
        return fmd_set_error(FMD_ECUDA, "coefficient upload failed: %s", cudaGetErrorString(cudaGetLastError()));
    }
    if (quality) {
        /* The recursion state is float32: a pole at radius rho amplifies its rounding noise by
         * roughly 1/(1-rho)^2.  Measured on B200 against the double-precision serial oracle:
         * rho = 0.9806 (Butterworth 3, fc/rate 0.006) 3e-6, rho = 0.9987 1e-4 (the bar),
         * rho = 0.9994 (Chebyshev 5, fc/rate 0.0017) 1.8e-2 -- profiles/r02_quality_narrow_diag.txt.
         * Designs past 0.9985 are refused instead of answered wrongly. */
        double rho = 0.0;
        for (unsigned r = 0; r < c->lout + c->lin; ++r) {
            const float *row = r < c->lout ? c->sec.out[r] : c->sec.in[r - c->lout];
            const double rr = row[5] != 0.0f ? sqrt(fabs((double) row[5])) : fabs((double) row[4]);
            if (rr > rho) rho = rr;
        }
        if (!(rho <= 0.9985)) {
            fmd_destroy(c);
            return fmd_set_error(FMD_EINVAL,
                                 "quality mode keeps the filter state in float32: the slowest pole of this design has "
                                 "radius %.5f (limit 0.9985) -- raise the cut-off or lower the degree", rho);
        }
        rc = fmq_setup(c, q_decim);
        if (rc) {
            fmd_destroy(c);
            return rc;
        }
    }
    if (conditioner == 1u) {
        rc = fmq1_setup(c);
        if (rc) {
            fmd_destroy(c);
            return rc;
        }
    }
    *out = c;
    return FMD_OK;
}

static void fmd_ring_free(fmd_ctx *c) {
    if (!c->slots) return;
    for (unsigned i = 0; i < c->nslots; ++i) {
        fmd_slot &s = c->slots[i];
        if (s.ev_out && s.busy) cudaEventSynchronize(s.ev_out);
        if (s.d_in) cudaFree(s.d_in);
        if (s.d_out) cudaFree(s.d_out);
        if (s.ev_in) cudaEventDestroy(s.ev_in);
        if (s.ev_k) cudaEventDestroy(s.ev_k);
        if (s.ev_out) cudaEventDestroy(s.ev_out);
    }
    delete[] c->slots;
    c->slots = nullptr;
    c->nslots = 0;
    c->slot_bytes = 0;
}

extern "C" void fmd_destroy(fmd_ctx *c) {
    if (!c) return;
    FmdDeviceScope on_device(c->device);
    fmd_ring_free(c);
    if (c->pev) {
        for (unsigned i = 0; i < 2 * c->pev_cap; ++i) cudaEventDestroy(c->pev[i]);
        delete[] c->pev;
    }
    if (c->d_sec) cudaFree(c->d_sec);
    fmq_free(c);
    fmq1_free(c);
    if (c->s_in) cudaStreamDestroy(c->s_in);
    if (c->s_k) cudaStreamDestroy(c->s_k);
    if (c->s_out) cudaStreamDestroy(c->s_out);
    delete c;
}

extern "C" int fmd_get_config(const fmd_ctx *c, fmd_config *cfg) {
    if (!c || !cfg) return fmd_set_error(FMD_EINVAL, "null argument");
    *cfg = c->cfg;
    return FMD_OK;
}

extern "C" int fmd_get_sos(const fmd_ctx *c, int which, float *sos) {
    if (!c || !sos) return fmd_set_error(FMD_EINVAL, "null argument");
    const unsigned rows = which ? c->lin : c->lout;
    memcpy(sos, which ? &c->sec.in[0][0] : &c->sec.out[0][0], rows * 6 * sizeof(float));
    return (int) rows;
}

extern "C" uint64_t fmd_launch_count(const fmd_ctx *c) { return c ? c->launches : 0; }

/* ------------------------------------------------------------ pinned memory */

extern "C" int fmd_host_alloc(void **ptr, size_t bytes) {
    if (!ptr) return fmd_set_error(FMD_EINVAL, "null argument");
    *ptr = nullptr;
    FMD_CUDA(cudaHostAlloc(ptr, bytes ? bytes : 1, cudaHostAllocPortable));
    return FMD_OK;
}

extern "C" int fmd_host_free(void *ptr) {
    if (ptr) FMD_CUDA(cudaFreeHost(ptr));
    return FMD_OK;
}

extern "C" int fmd_host_register(void *ptr, size_t bytes) {
    FMD_CUDA(cudaHostRegister(ptr, bytes, cudaHostRegisterPortable));
    return FMD_OK;
}

extern "C" int fmd_host_unregister(void *ptr) {
    FMD_CUDA(cudaHostUnregister(ptr));
    return FMD_OK;
}

/* ---------------------------------------------------------------- launching */

extern "C" size_t fmd_out_count(const fmd_ctx *c, size_t nbytes) {
    if (c && c->q) return fmq_out_count(c, nbytes / 2);
    return nbytes >> 2; /* whole chunks give B/4 each; the ragged tail gives r/4 */
}

extern "C" size_t fmd_max_out_count(const fmd_ctx *c, size_t nbytes) {
    if (c && c->q) return fmq_max_out_count(c, nbytes / 2);
    return nbytes >> 2;
}

/* the streaming kernel's instantiations are compiled in groups (fmd_fast_inst.cu, one object per
 * group, so that the build runs in parallel); `at` = terms of the atan polynomial: 6 (1.75e-6)
 * unless the output stencil amplifies too much */
static fmd_fast_fn pick_fast(unsigned nti, unsigned nto, unsigned at, unsigned conditioner) {
    const unsigned a8 = at <= 6 ? 0u : 1u;
    if (conditioner == 3u) return fmd_pick_q3(nti, nto, a8);
    if (nti <= 1) return fmd_pick_n01(nti, nto, a8);
    if (nti <= 3) return fmd_pick_n23(nti, nto, a8);
    return fmd_pick_n45(nti, nto, a8);
}

/* every output of every chunk through the op-for-op path (staged through shared memory) */
static int launch_exact(fmd_ctx *c, const uint8_t *d_in, float *d_out, FmdGeom g, cudaStream_t s) {
    if (!g.chunk_out || !g.nchunks) return FMD_OK;
    g.head = g.chunk_out;
    g.fast_out = g.chunk_out;
    if (g.chunk_out < FMD_XT) {
        /* chunks shorter than a tile (-b down to 4 bytes): one thread per output, no staging */
        const unsigned long long total = (unsigned long long) g.nchunks * g.chunk_out;
        const unsigned long long nb = (total + FMD_BLOCK - 1) / FMD_BLOCK;
        if (nb > 0x7fffffffull) return fmd_set_error(FMD_EINVAL, "exact kernel grid too large");
        g.exact_blocks = (unsigned) nb;
        fmd_exact_kernel<<<(unsigned) nb, FMD_BLOCK, 0, s>>>(d_in, d_out, g, c->d_sec);
        FMD_CUDA(cudaGetLastError());
        c->launches++;
        return FMD_OK;
    }
    const unsigned long long blocks = (unsigned long long) g.nchunks * ((g.chunk_out + FMD_XT - 1u) / FMD_XT);
    if (blocks > 0x7fffffffull) return fmd_set_error(FMD_EINVAL, "exact kernel grid too large");
    g.exact_blocks = (unsigned) blocks;
    fmd_exact_tiled_kernel<<<(unsigned) blocks, FMD_XT, 0, s>>>(d_in, d_out, g, c->d_sec);
    FMD_CUDA(cudaGetLastError());
    c->launches++;
    return FMD_OK;
}

/* `nchunks` chunks of `chunk_out` outputs each, chunk starts `stride` bytes apart */
static int launch_chunks(fmd_ctx *c, const uint8_t *d_in, float *d_out, unsigned nchunks, unsigned chunk_out,
                         unsigned long long stride, cudaStream_t s, bool float_in = false) {
    if (!nchunks || !chunk_out) return FMD_OK;
    FmdGeom g;
    g.chunk_stride = stride;
    g.chunk_out = chunk_out;
    g.nchunks = nchunks;
    g.fast_out = 0;
    g.iters_per_seg = 1;
    g.segs_per_chunk = 1;
    g.head = 0;
    g.exact_blocks = 0;

    /* float_in: `d_in` holds conditioned floats, one per input byte (-q1); same geometry, the
     * float-input instantiation of the streaming kernel */
    fmd_fast_fn fn = float_in ? fmd_pick_float(c->nti, c->nto, c->atan_terms <= 6 ? 0u : 1u)
                              : pick_fast(c->nti, c->nto, c->atan_terms, (unsigned) c->sec.conditioner);
    const bool aligned = ((uintptr_t) d_in % 16 == 0) && ((uintptr_t) d_out % 16 == 0) && (stride % 16 == 0) &&
                         (nchunks == 1 || chunk_out % 4 == 0);
    /* the fast kernel covers whole 128-output passes; if the chunk is not a multiple of 128
     * outputs it stops one pass early so that its drain pass still reads real samples */
    const unsigned fast_out = (chunk_out % FMD_PASS_OUT == 0)
                                  ? chunk_out
                                  : (chunk_out >= 2 * FMD_PASS_OUT ? (chunk_out / FMD_PASS_OUT - 1) * FMD_PASS_OUT : 0);
    const bool fast_ok = fn && aligned && c->fast_in_ok && !(c->cfg.flags & FMD_FLAG_EXACT) && fast_out >= FMD_PASS_OUT &&
                         fast_out > c->head;
    if (!fast_ok) return launch_exact(c, d_in, d_out, g, s);

    g.fast_out = fast_out;
    const unsigned iters_chunk = g.fast_out >> 7;
    /* passes per warp segment: long enough to amortise the one extra (drain) pass, short
     * enough that the grid still holds a few waves of warps for small inputs */
    const unsigned long long total_iters = (unsigned long long) iters_chunk * nchunks;
    const unsigned long long want_warps = (unsigned long long) c->sm_count * 64ull * 2ull;
    /* measured on B200 (profiles/): without the input filter the kernel is HBM-bound and short
     * segments keep the chip inside a compact window of memory (16 passes: 6.6 TB/s, 64: 6.4);
     * with it the kernel is pipe-bound and longer segments amortise the drain pass (64: 5.6 TB/s,
     * 16: 5.4) */
    unsigned ips = (c->nti && !float_in) ? 64u : 16u; /* the float-input kernel is HBM-bound like the plain one */
    if (const char *e = getenv("FMD_IPS")) ips = (unsigned) atoi(e); /* tuning experiments only */
    if (ips < 1) ips = 1;
    while (ips > 4 && total_iters / ips < want_warps / 4) ips >>= 1;
    if (ips > iters_chunk) ips = iters_chunk;
    g.iters_per_seg = ips;
    g.segs_per_chunk = (iters_chunk + ips - 1) / ips;
    const unsigned long long nseg = (unsigned long long) g.segs_per_chunk * nchunks;
    if (nseg > 0xffffffffull) return fmd_set_error(FMD_EINVAL, "too many segments in one launch");
    /* chunk heads and the ragged remainder ride along as the first CTAs of the same launch */
    g.head = c->head;
    const unsigned long long exact_threads = (unsigned long long) nchunks * (g.head + (chunk_out - fast_out));
    g.exact_blocks = (unsigned) ((exact_threads + FMD_BLOCK - 1) / FMD_BLOCK);
    const unsigned long long blocks = (nseg + FMD_WARPS - 1) / FMD_WARPS + g.exact_blocks;
    if (blocks > 0x7fffffffull) return fmd_set_error(FMD_EINVAL, "fast kernel grid too large");
    cudaEvent_t *pe = nullptr;
    if (c->profile) {
        if (c->pev_used == c->pev_cap) {
            int rc = prof_flush(c);
            if (rc) return rc;
        }
        pe = &c->pev[2 * c->pev_used++];
        FMD_CUDA(cudaEventRecord(pe[0], s));
    }
    fn<<<(unsigned) blocks, FMD_BLOCK, 0, s>>>(d_in, d_out, g, c->st, c->d_sec);
    FMD_CUDA(cudaGetLastError());
    if (pe) FMD_CUDA(cudaEventRecord(pe[1], s));
    c->launches++;
    return FMD_OK;
}


extern "C" int fmd_reset(fmd_ctx *c) {
    if (!c) return fmd_set_error(FMD_EINVAL, "null argument");
    if (c->q1) return fmq1_reset(c);
    return c->q ? fmq_reset(c) : FMD_OK;
}

extern "C" int fmd_demod_device(fmd_ctx *c, const uint8_t *d_in, size_t nbytes, float *d_out, void *stream) {
    if (!c || (!d_in && nbytes) || (!d_out && nbytes)) return fmd_set_error(FMD_EINVAL, "null argument");
    cudaStream_t s = (cudaStream_t) stream; /* NULL is CUDA's default stream, as everywhere in CUDA */
    const unsigned long long B = c->cfg.chunk_bytes;
    const unsigned long long nfull = nbytes / B;
    const unsigned long long rem = (nbytes - nfull * B) & ~3ull;
    if (nfull > 0xffffffffull) return fmd_set_error(FMD_EINVAL, "too many chunks in one call");
    /* launch on the context's device whatever the calling thread's current device is */
    int prev = -1;
    FMD_CUDA(cudaGetDevice(&prev));
    if (prev != c->device) FMD_CUDA(cudaSetDevice(c->device));
    int rc;
    if (c->q) {
        rc = fmq_demod_device(c, d_in, nbytes, d_out, s);
    } else if (c->q1) {
        /* -q1: the offsets live on from chunk to chunk, so a call is a run of WHOLE chunks that
         * continues the stream; condition them, then the op-for-op path on the conditioned floats */
        if (rem || nbytes != nfull * B)
            rc = fmd_set_error(FMD_EINVAL, "-q1 takes whole chunks (%llu bytes each)", B);
        else if (((uintptr_t) d_in % 16) || ((uintptr_t) d_out % 16))
            rc = fmd_set_error(FMD_EINVAL, "-q1 needs 16-byte aligned device buffers");
        else if (!nfull)
            rc = FMD_OK;
        else {
            /* one kernel conditions and demodulates (fmd_q1_fused.cuh); the chunk heads then come from
             * the op-for-op path over the conditioned head samples it left behind */
            const float *heads = nullptr;
            unsigned heads_stride = 0;
            bool fused = false;
            rc = fmq1_fused(c, d_in, nfull, d_out, s, &heads, &heads_stride, &fused);
            if (!rc && fused && c->head) {
                FmdGeom g;
                memset(&g, 0, sizeof g);
                g.chunk_stride = heads_stride;
                g.chunk_out = (unsigned) (B >> 2);
                g.nchunks = (unsigned) nfull;
                g.fast_out = g.chunk_out;
                g.head = c->head;
                const unsigned long long nb = ((unsigned long long) g.nchunks * g.head + FMD_BLOCK - 1) / FMD_BLOCK;
                g.exact_blocks = (unsigned) nb;
                fmd_exact_kernel<<<(unsigned) nb, FMD_BLOCK, 0, s>>>(reinterpret_cast<const uint8_t *>(heads), d_out, g, c->d_sec);
                if (cudaGetLastError() != cudaSuccess) rc = fmd_set_error(FMD_ECUDA, "head kernel launch failed");
                c->launches++;
            }
            const float *sbuf = nullptr;
            if (!rc && !fused) rc = fmq1_condition(c, d_in, nfull, s, &sbuf);
            /* the conditioned floats stream through the float-input kernel (chunk heads on its exact
             * CTAs); designs it has no instantiation for fall back to the op-for-op path inside */
            if (!rc && !fused)
                rc = launch_chunks(c, reinterpret_cast<const uint8_t *>(sbuf), d_out, (unsigned) nfull, (unsigned) (B >> 2),
                                   B * sizeof(float), s, true);
        }
    } else {
        rc = launch_chunks(c, d_in, d_out, (unsigned) nfull, (unsigned) (B >> 2), B, s);
        if (!rc && rem)
            rc = launch_chunks(c, d_in + nfull * B, d_out + nfull * (B >> 2), 1u, (unsigned) (rem >> 2), B, s);
    }
    if (prev != c->device && prev >= 0) cudaSetDevice(prev);
    return rc;
}

extern "C" int fmd_time_device(fmd_ctx *c, const uint8_t *d_in, size_t nbytes, float *d_out, void *stream,
                               int reps, float *ms) {
    if (!c || !ms || reps <= 0) return fmd_set_error(FMD_EINVAL, "bad argument");
    cudaStream_t s = (cudaStream_t) stream;
    /* events belong to a device: record them on the context's, whatever the caller's current one is */
    int prev = -1;
    FMD_CUDA(cudaGetDevice(&prev));
    if (prev != c->device) FMD_CUDA(cudaSetDevice(c->device));
    cudaEvent_t a = nullptr, b = nullptr;
    cudaError_t e = cudaEventCreate(&a);
    if (e == cudaSuccess) e = cudaEventCreate(&b);
    if (e == cudaSuccess) e = cudaEventRecord(a, s);
    int rc = FMD_OK;
    for (int i = 0; i < reps && !rc && e == cudaSuccess; ++i) rc = fmd_demod_device(c, d_in, nbytes, d_out, s);
    if (!rc && e == cudaSuccess) e = cudaEventRecord(b, s);
    if (!rc && e == cudaSuccess) e = cudaEventSynchronize(b);
    if (!rc && e == cudaSuccess) e = cudaEventElapsedTime(ms, a, b);
    if (a) cudaEventDestroy(a);
    if (b) cudaEventDestroy(b);
    if (prev != c->device && prev >= 0) cudaSetDevice(prev);
    if (!rc && e != cudaSuccess) rc = fmd_set_error(FMD_ECUDA, "fmd_time_device: %s", cudaGetErrorString(e));
    return rc;
}

/* --------------------------------------------------------------------- ring */

extern "C" int fmd_ring_create(fmd_ctx *c, unsigned slots, size_t slot_bytes) {
    if (!c || slots == 0 || slots > 64 || slot_bytes == 0) return fmd_set_error(FMD_EINVAL, "bad ring geometry");
    if (slot_bytes % c->cfg.chunk_bytes)
        return fmd_set_error(FMD_EINVAL, "slot_bytes must be a multiple of chunk_bytes");
    FmdDeviceScope on_device(c->device);
    fmd_ring_free(c);
    c->slots = new (std::nothrow) fmd_slot[slots];
    if (!c->slots) return fmd_set_error(FMD_ENOMEM, "out of host memory");
    c->nslots = slots;
    c->slot_bytes = slot_bytes;
    const size_t out_bytes = fmd_max_out_count(c, slot_bytes) * sizeof(float);
    for (unsigned i = 0; i < slots; ++i) {
        fmd_slot &s = c->slots[i];
        if (cudaMalloc(&s.d_in, slot_bytes) != cudaSuccess ||
            cudaMalloc(&s.d_out, out_bytes ? out_bytes : sizeof(float)) != cudaSuccess ||
            cudaEventCreateWithFlags(&s.ev_in, cudaEventDisableTiming) != cudaSuccess ||
            cudaEventCreateWithFlags(&s.ev_k, cudaEventDisableTiming) != cudaSuccess ||
            cudaEventCreateWithFlags(&s.ev_out, cudaEventDisableTiming) != cudaSuccess) {
            const char *m = cudaGetErrorString(cudaGetLastError());
            fmd_ring_free(c);
            return fmd_set_error(FMD_ECUDA, "ring allocation failed: %s", m);
        }
    }
    return FMD_OK;
}

/* `run_kernel` = false: the same copies in the same order without the compute in between (the
 * copy ceiling of the host side, fmd_copy_host) */
static int submit_slot(fmd_ctx *c, unsigned slot, const uint8_t *in, size_t nbytes, float *out, bool run_kernel,
                       bool device_ordered = false) {
    if (!c || slot >= c->nslots || !in || !out) return fmd_set_error(FMD_EINVAL, "bad slot or null buffer");
    if (nbytes > c->slot_bytes) return fmd_set_error(FMD_EINVAL, "nbytes exceeds slot_bytes");
    fmd_slot &s = c->slots[slot];
    if (s.busy && !device_ordered) return fmd_set_error(FMD_EBUSY, "slot %u still in flight: call fmd_wait first", slot);
    FmdDeviceScope on_device(c->device);
    const size_t nout = fmd_out_count(c, nbytes);
    if (s.busy) {
        /* device_ordered (fmd_demod_host): the slot's previous occupant is still in flight and the
         * HOST does not wait for it -- the streams do: the copy-in waits until the previous kernel has
         * read d_in, the kernel until the previous copy-out has read d_out.  The host only enqueues,
         * so its wake-up latency never sits between two transfers. */
        FMD_CUDA(cudaStreamWaitEvent(c->s_in, s.ev_k, 0));
        FMD_CUDA(cudaStreamWaitEvent(c->s_k, s.ev_out, 0));
    }
    /* (otherwise the slot's device buffers are free: the previous occupant's D2H completed in fmd_wait) */
    FMD_CUDA(cudaMemcpyAsync(s.d_in, in, nbytes, cudaMemcpyHostToDevice, c->s_in));
    FMD_CUDA(cudaEventRecord(s.ev_in, c->s_in));
    FMD_CUDA(cudaStreamWaitEvent(c->s_k, s.ev_in, 0));
    int rc = run_kernel ? fmd_demod_device(c, s.d_in, nbytes, s.d_out, c->s_k) : FMD_OK;
    if (rc) return rc;
    FMD_CUDA(cudaEventRecord(s.ev_k, c->s_k));
    FMD_CUDA(cudaStreamWaitEvent(c->s_out, s.ev_k, 0));
    const size_t out_elem = (c->cfg.flags & FMD_FLAG_S16) ? sizeof(short) : sizeof(float);
    FMD_CUDA(cudaMemcpyAsync(out, s.d_out, nout * out_elem, cudaMemcpyDeviceToHost, c->s_out));
    FMD_CUDA(cudaEventRecord(s.ev_out, c->s_out));
    s.busy = true;
    return FMD_OK;
}

extern "C" int fmd_submit(fmd_ctx *c, unsigned slot, const uint8_t *in, size_t nbytes, float *out) {
    return submit_slot(c, slot, in, nbytes, out, true);
}

/* ---- the quality pipeline's stream over several GPUs: segments with a warm-up prefix ---- */

extern "C" uint64_t fmd_warmup_samples(const fmd_ctx *c) { return (c && c->q) ? fmq_warmup_samples(c) : 0; }

extern "C" size_t fmd_segment_out_count(const fmd_ctx *c, uint64_t first_sample, size_t nsamples) {
    if (c && c->q) return fmq_out_count_at(c, first_sample, nsamples);
    return nsamples >> 1;
}

extern "C" int fmd_submit_segment(fmd_ctx *c, unsigned slot, const uint8_t *in, size_t warm_bytes, size_t nbytes,
                                  uint64_t first_sample, float *out) {
    if (!c || slot >= c->nslots || !in || !out) return fmd_set_error(FMD_EINVAL, "bad slot or null buffer");
    if (!c->q)
        return fmd_set_error(FMD_EINVAL, "segments belong to the quality pipeline; in reference-compatible mode chunks "
                                         "are independent and fmd_submit needs no warm-up");
    if ((warm_bytes & 15u) || (nbytes & 1u) || warm_bytes / 2 > first_sample)
        return fmd_set_error(FMD_EINVAL, "warm-up must be a multiple of 16 bytes that does not reach before the stream start");
    if (warm_bytes + nbytes > c->slot_bytes) return fmd_set_error(FMD_EINVAL, "warm-up + payload exceed slot_bytes");
    fmd_slot &s = c->slots[slot];
    if (s.busy) return fmd_set_error(FMD_EBUSY, "slot %u still in flight: call fmd_wait first", slot);
    FmdDeviceScope on_device(c->device);
    const size_t out_elem = (c->cfg.flags & FMD_FLAG_S16) ? sizeof(short) : sizeof(float);
    const uint64_t seg_start = first_sample - warm_bytes / 2;
    const size_t n_warm = fmq_out_count_at(c, seg_start, warm_bytes / 2);
    const size_t n_pay = fmq_out_count_at(c, first_sample, nbytes / 2);
    FMD_CUDA(cudaMemcpyAsync(s.d_in, in, warm_bytes + nbytes, cudaMemcpyHostToDevice, c->s_in));
    FMD_CUDA(cudaEventRecord(s.ev_in, c->s_in));
    FMD_CUDA(cudaStreamWaitEvent(c->s_k, s.ev_in, 0));
    int rc = fmq_seek(c, seg_start, c->s_k);
    if (!rc) rc = fmd_demod_device(c, s.d_in, warm_bytes + nbytes, s.d_out, c->s_k);
    if (rc) return rc;
    FMD_CUDA(cudaEventRecord(s.ev_k, c->s_k));
    FMD_CUDA(cudaStreamWaitEvent(c->s_out, s.ev_k, 0));
    if (n_pay)
        FMD_CUDA(cudaMemcpyAsync(out, reinterpret_cast<const char *>(s.d_out) + n_warm * out_elem, n_pay * out_elem,
                                 cudaMemcpyDeviceToHost, c->s_out));
    FMD_CUDA(cudaEventRecord(s.ev_out, c->s_out));
    s.busy = true;
    return FMD_OK;
}

extern "C" int fmd_wait(fmd_ctx *c, unsigned slot) {
    if (!c || slot >= c->nslots) return fmd_set_error(FMD_EINVAL, "bad slot");
    fmd_slot &s = c->slots[slot];
    if (!s.busy) return FMD_OK;
    FMD_CUDA(cudaEventSynchronize(s.ev_out));
    s.busy = false;
    if (c->q) return fmq_check_error(c); /* a look-back that gave up waiting raises a device-side flag */
    return FMD_OK;
}

extern "C" int fmd_check(fmd_ctx *c, void *stream) {
    if (!c) return fmd_set_error(FMD_EINVAL, "null argument");
    int prev = -1;
    FMD_CUDA(cudaGetDevice(&prev));
    if (prev != c->device) FMD_CUDA(cudaSetDevice(c->device));
    cudaError_t e = cudaStreamSynchronize((cudaStream_t) stream);
    if (prev != c->device && prev >= 0) cudaSetDevice(prev);
    if (e != cudaSuccess) return fmd_set_error(FMD_ECUDA, "fmd_check: %s", cudaGetErrorString(e));
    return c->q ? fmq_check_error(c) : FMD_OK;
}

/* ------------------------------------------------------- host-buffer wrapper */

static int host_pipeline(fmd_ctx *c, const uint8_t *in, size_t nbytes, float *out, size_t *nout, bool run_kernel);

extern "C" int fmd_demod_host(fmd_ctx *c, const uint8_t *in, size_t nbytes, float *out, size_t *nout) {
    return host_pipeline(c, in, nbytes, out, nout, true);
}

extern "C" int fmd_copy_host(fmd_ctx *c, const uint8_t *in, size_t nbytes, float *out) {
    if (c && c->q) return fmd_set_error(FMD_EINVAL, "fmd_copy_host measures the reference-compatible pipeline's copies");
    return host_pipeline(c, in, nbytes, out, nullptr, false);
}

static int host_pipeline(fmd_ctx *c, const uint8_t *in, size_t nbytes, float *out, size_t *nout, bool run_kernel) {
    if (!c || (!in && nbytes) || (!out && nbytes)) return fmd_set_error(FMD_EINVAL, "null argument");
    if (nout) *nout = 0;
    const size_t B = c->cfg.chunk_bytes;
    /* Slices of 64 MiB (whole chunks), 3 in flight, graded at both ends of the call.  Measured on the
     * B200 box (profiles/r02_e2e_probe.txt): every slice that has a kernel between its two copies costs
     * the link about 40 us whatever its size -- 43.2 / 45.7 / 46.9 GB/s per direction with 16 / 32 / 64 MiB
     * against 48.6-49.0 for the same copies alone -- while the first copy-in and the last copy-out of a
     * call run with one direction idle, which is what the 4, 8, 16 ... MiB start and the halving end
     * keep short (47.1 with both; 128 MiB and more gain nothing). */
    size_t slice = (size_t) 64 << 20;
    bool ramp = true;
    if (const char *e = getenv("FMD_HOST_SLICE_MB")) { /* tuning experiments */
        const long v = atol(e);
        if (v >= 1 && v <= 1024) slice = (size_t) v << 20;
    }
    if (const char *e = getenv("FMD_HOST_RAMP")) ramp = atoi(e) != 0;
    slice = slice / B * B;
    if (slice == 0) slice = B;
    if (slice > nbytes) slice = (nbytes + B - 1) / B * B;
    if (slice == 0) slice = B;
    unsigned want_slots = 3;
    if (const char *e = getenv("FMD_HOST_SLOTS")) { /* tuning experiments */
        const int v = atoi(e);
        if (v >= 2 && v <= 16) want_slots = (unsigned) v;
    }
    if (c->nslots < want_slots || c->slot_bytes < slice) {
        int rc = fmd_ring_create(c, want_slots, slice);
        if (rc) return rc;
    }
    const unsigned ns = c->nslots;
    size_t off = 0, out_off = 0;
    unsigned n = 0;
    bool host_ordered = false;
    if (const char *e = getenv("FMD_HOST_ORDERED")) host_ordered = atoi(e) != 0;
    /* slices grow from 4 MiB at the start of the call and shrink towards its end (fmd_slices.h;
     * FMD_HOST_RAMP=0: equal slices) */
    while (off < nbytes) {
        const unsigned slot = n % ns;
        const size_t len = fmd_slice_len(n, nbytes - off, B, c->slot_bytes, ramp ? 1 : 0);
        const size_t produced = fmd_out_count(c, len); /* before the submit moves the stream on */
        float *dst = (c->cfg.flags & FMD_FLAG_S16) ? reinterpret_cast<float *>(reinterpret_cast<short *>(out) + out_off)
                                                   : out + out_off;
        /* slot reuse is ordered on the device (stream waits on the slot's events), not by the host */
        int rc = submit_slot(c, slot, in + off, len, dst, run_kernel, !host_ordered);
        if (rc == FMD_EBUSY) { /* FMD_HOST_ORDERED=1 (the round-1 form, kept for A/B): wait, then submit */
            rc = fmd_wait(c, slot);
            if (!rc) rc = submit_slot(c, slot, in + off, len, dst, run_kernel);
        }
        if (rc) return rc;
        off += len;
        out_off += produced;
        ++n;
    }
    for (unsigned i = 0; i < ns; ++i) {
        int rc = fmd_wait(c, i);
        if (rc) return rc;
    }
    if (nout) *nout = out_off;
    return FMD_OK;
}
