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

#include "fmd_internal.h"
#include "fmd_kernels.cuh"
#include "fmd_quality.cuh"

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

#define FMD_CUDA(call)                                                                              \
    do {                                                                                            \
        cudaError_t e__ = (call);                                                                   \
        if (e__ != cudaSuccess)                                                                     \
            return fmd_set_error(FMD_ECUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), \
                                 __FILE__, __LINE__);                                               \
    } while (0)

/* ----------------------------------------------------------------- context */

struct fmd_ctx;
static int quality_setup(fmd_ctx *c, unsigned decim);
static void quality_free(fmd_ctx *c);
static size_t quality_out_count(unsigned long long pos, size_t nsamp, unsigned M);

struct fmd_slot {
    uint8_t *d_in = nullptr;
    float *d_out = nullptr;
    cudaEvent_t ev_in = nullptr;   /* H2D landed              */
    cudaEvent_t ev_k = nullptr;    /* kernels done            */
    cudaEvent_t ev_out = nullptr;  /* D2H landed              */
    bool busy = false;
};

/* one cascade of the quality pipeline: host design + constants + device tables */
struct fmq_filter {
    fmq_design des;
    fmq_consts c;      /* scan over the threads of a CTA (unit T)   */
    fmq_consts c_tile; /* scan over the tiles of a launch (unit TILE) */
    fmq_table_sizes sz;
    unsigned zir_len = FMQ_TILE; /* tile offsets whose zero-input response is not negligible */
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
    unsigned lin = 0, lout = 0; /* sections */
    fmq_filter fin, fout;
    FmqState *d_state = nullptr;
    float *d_ein = nullptr, *d_qin = nullptr, *d_eout = nullptr, *d_qout = nullptr; /* per-tile scratch */
    unsigned *d_lb = nullptr; /* look-back ticket + per-tile flags */
    unsigned tile_cap = 0;
    bool smem_opted_in = false;
    unsigned long long pos = 0; /* stream position: global index of the next input sample */
};

struct fmd_ctx {
    fmd_config cfg;
    fmq_ctx *q = nullptr;
    int device = 0;
    unsigned lin = 0, lout = 0;
    FmdSections sec;
    FmdSections *d_sec = nullptr; /* device copy, read by the exact path */
    FmdStencil st;
    unsigned nti = 0, nto = 0;
    unsigned atan_terms = 6;
    bool fast_in_ok = true; /* input stencil could be normalised by its first tap */
    unsigned head = 0; /* outputs at the start of every chunk that are not in steady state */
    cudaStream_t s_in = nullptr, s_k = nullptr, s_out = nullptr;
    fmd_slot *slots = nullptr;
    unsigned nslots = 0;
    size_t slot_bytes = 0;
    uint64_t launches = 0;
    int sm_count = 148;
    /* optional per-launch event timing of the fast kernel */
    bool profile = false;
    cudaEvent_t *pev = nullptr; /* pairs: [2*i] before, [2*i+1] after */
    unsigned pev_used = 0, pev_cap = 0;
    double prof_ms = 0.0;
    uint64_t prof_n = 0;
};

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

extern "C" int fmd_profile_kernels(fmd_ctx *c, int enable) {
    if (!c) return fmd_set_error(FMD_EINVAL, "null argument");
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
    if (conditioner == 1u || conditioner == 2u)
        return fmd_set_error(FMD_EINVAL,
                             "input conditioner -q%u is outside the accelerated path (-q0 and -q3 only: -q1 is a "
                             "serial recurrence across chunks, -q2 reads past its buffer in the reference)",
                             conditioner);
    if (((cfg.mode >> 4) & 3u) == 0)
        return fmd_set_error(FMD_EINVAL, "demodulation off (-c) is unreachable in the reference CLI and unsupported");

    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev <= 0)
        return fmd_set_error(FMD_ENODEV, "no CUDA device: %s (this library has no CPU path)",
                             e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
    if (device < 0 || device >= ndev) return fmd_set_error(FMD_EINVAL, "device %d not in 0..%d", device, ndev - 1);
    FMD_CUDA(cudaSetDevice(device));
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
    /* -q3 (normalizeInput) streams without the input filter; with it the zero extension past a
     * chunk is not a byte value (no u8 normalises to 0.0), so that combination stays op-for-op */
    if (conditioner == 3u && c->lin) c->fast_in_ok = false;

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
    if (rc) {
        delete c;
        return rc;
    }
    {
        /* the discriminator's error reaches the output multiplied by sum|tout|: keep it near 2e-5 */
        double gain = 0.0;
        for (unsigned t = 0; t < c->nto; ++t) gain += fabs((double) c->st.tout[t]);
        c->atan_terms = (gain * 1.75e-6 <= 2.0e-5) ? 6u : 8u;
        if (const char *e = getenv("FMD_ATAN_TERMS")) c->atan_terms = (unsigned) atoi(e); /* audits */
    }
    c->head = 2 * c->lout - 1;
    if (c->lin && 4 * c->lin - 2 > c->head) c->head = 4 * c->lin - 2;
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
        rc = quality_setup(c, q_decim);
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
    cudaSetDevice(c->device);
    fmd_ring_free(c);
    if (c->pev) {
        for (unsigned i = 0; i < 2 * c->pev_cap; ++i) cudaEventDestroy(c->pev[i]);
        delete[] c->pev;
    }
    if (c->d_sec) cudaFree(c->d_sec);
    quality_free(c);
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
    if (c && c->q) return quality_out_count(c->q->pos, nbytes / 2, c->q->decim);
    return nbytes >> 2; /* whole chunks give B/4 each; the ragged tail gives r/4 */
}

extern "C" size_t fmd_max_out_count(const fmd_ctx *c, size_t nbytes) {
    if (c && c->q) return quality_out_count(0, nbytes / 2, c->q->decim);
    return nbytes >> 2;
}

typedef void (*fast_fn)(const uint8_t *, float *, const FmdGeom, const FmdStencil, const FmdSections *);

template <int NTI, int AT>
static fast_fn pick_out(unsigned nto) {
    switch (nto) {
        case 1: return fmd_fast_kernel<NTI, 1, AT>;
        case 2: return fmd_fast_kernel<NTI, 2, AT>;
        case 3: return fmd_fast_kernel<NTI, 3, AT>;
        case 4: return fmd_fast_kernel<NTI, 4, AT>;
        case 5: return fmd_fast_kernel<NTI, 5, AT>;
        default: return nullptr;
    }
}

template <int AT>
static fast_fn pick_in(unsigned nti, unsigned nto) {
    switch (nti) {
        case 0: return pick_out<0, AT>(nto);
        case 1: return pick_out<1, AT>(nto);
        case 2: return pick_out<2, AT>(nto);
        case 3: return pick_out<3, AT>(nto);
        case 4: return pick_out<4, AT>(nto);
        case 5: return pick_out<5, AT>(nto);
        default: return nullptr;
    }
}

/* -q3 (normalizeInput) streams when there is no input filter */
template <int AT>
static fast_fn pick_q3(unsigned nto) {
    switch (nto) {
        case 1: return fmd_fast_kernel<0, 1, AT, 3>;
        case 2: return fmd_fast_kernel<0, 2, AT, 3>;
        case 3: return fmd_fast_kernel<0, 3, AT, 3>;
        case 4: return fmd_fast_kernel<0, 4, AT, 3>;
        case 5: return fmd_fast_kernel<0, 5, AT, 3>;
        default: return nullptr;
    }
}

/* `at` = terms of the atan polynomial: 6 (1.75e-6) unless the output stencil amplifies too much */
static fast_fn pick_fast(unsigned nti, unsigned nto, unsigned at, unsigned conditioner) {
    if (conditioner == 3u) return nti ? nullptr : (at <= 6 ? pick_q3<6>(nto) : pick_q3<8>(nto));
    return at <= 6 ? pick_in<6>(nti, nto) : pick_in<8>(nti, nto);
}

/* every output of every chunk through the op-for-op path */
static int launch_exact(fmd_ctx *c, const uint8_t *d_in, float *d_out, FmdGeom g, cudaStream_t s) {
    if (!g.chunk_out || !g.nchunks) return FMD_OK;
    g.head = g.chunk_out;
    g.fast_out = g.chunk_out;
    const unsigned long long total = (unsigned long long) g.nchunks * g.chunk_out;
    const unsigned long long blocks = (total + FMD_BLOCK - 1) / FMD_BLOCK;
    if (blocks > 0x7fffffffull) return fmd_set_error(FMD_EINVAL, "exact kernel grid too large");
    g.exact_blocks = (unsigned) blocks;
    fmd_exact_kernel<<<(unsigned) blocks, FMD_BLOCK, 0, s>>>(d_in, d_out, g, c->d_sec);
    FMD_CUDA(cudaGetLastError());
    c->launches++;
    return FMD_OK;
}

/* `nchunks` chunks of `chunk_out` outputs each, chunk starts `stride` bytes apart */
static int launch_chunks(fmd_ctx *c, const uint8_t *d_in, float *d_out, unsigned nchunks, unsigned chunk_out,
                         unsigned long long stride, cudaStream_t s) {
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

    fast_fn fn = pick_fast(c->nti, c->nto, c->atan_terms, (unsigned) c->sec.conditioner);
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
    unsigned ips = c->nti ? 64u : 16u;
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


/* ------------------------------------------------------------ quality pipeline (opt-in) */

static int quality_upload_filter(fmq_filter &f, const float *sos6, unsigned sections) {
    int rc = fmq_design_from_sos(sos6, sections, &f.des);
    if (rc) return rc;
    fmq_build_consts(&f.des, FMQ_T, &f.c);
    fmq_build_consts(&f.des, FMQ_TILE, &f.c_tile);
    fmq_table_layout(&f.des, &f.sz);
    const size_t n = (size_t) f.sz.lane + f.sz.apow + f.sz.ck + f.sz.ckn + f.sz.lane_tile;
    float *h = (float *) malloc(n * sizeof(float));
    if (!h) return fmd_set_error(FMD_ENOMEM, "out of host memory");
    f.zir_len = fmq_build_tables(&f.des, h, h + f.sz.lane, h + f.sz.lane + f.sz.apow, h + f.sz.lane + f.sz.apow + f.sz.ck,
                                 h + f.sz.lane + f.sz.apow + f.sz.ck + f.sz.ckn);
    cudaError_t e = cudaMalloc(&f.d_tab, n * sizeof(float));
    if (e == cudaSuccess) e = cudaMemcpy(f.d_tab, h, n * sizeof(float), cudaMemcpyHostToDevice);
    free(h);
    if (e != cudaSuccess) return fmd_set_error(FMD_ECUDA, "quality table upload failed: %s", cudaGetErrorString(e));
    return FMD_OK;
}

static int quality_setup(fmd_ctx *c, unsigned decim) {
    fmq_ctx *q = new (std::nothrow) fmq_ctx();
    if (!q) return fmd_set_error(FMD_ENOMEM, "out of host memory");
    c->q = q;
    q->decim = decim;
    q->lout = c->lout;
    q->lin = c->lin;
    int rc = quality_upload_filter(q->fout, &c->sec.out[0][0], q->lout);
    if (!rc && q->lin) rc = quality_upload_filter(q->fin, &c->sec.in[0][0], q->lin);
    if (rc) return rc;
    /* the kernels for odd degrees skip the second state of the last (first-order) section */
    if (((c->cfg.out_degree & 1u) && (c->sec.out[q->lout - 1][2] != 0.0f || c->sec.out[q->lout - 1][5] != 0.0f)) ||
        (q->lin && (c->cfg.in_degree & 1u) && (c->sec.in[q->lin - 1][2] != 0.0f || c->sec.in[q->lin - 1][5] != 0.0f)))
        return fmd_set_error(FMD_EINVAL, "odd-degree design whose last section is not first order");
    FMD_CUDA(cudaMalloc(&q->d_state, sizeof(FmqState)));
    FMD_CUDA(cudaMemset(q->d_state, 0, sizeof(FmqState)));
    return FMD_OK;
}

static void quality_free(fmd_ctx *c) {
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

/* outputs of the next `nsamp` samples when the stream stands at `pos` */
static size_t quality_out_count(unsigned long long pos, size_t nsamp, unsigned M) {
    const unsigned long long r0 = (M - pos % M) % M;
    return nsamp > r0 ? (size_t) ((nsamp - r0 + M - 1) / M) : 0;
}

typedef void (*fmq_main_fn)(const uint8_t *, float *, const FmqGeom, const fmq_consts, const fmq_consts, const float *,
                            const float *, const float *, const float *, const float *, const float *,
                            const FmqLookback, const FmqMat, const FmqMat, float *, FmqState *);

template <int NIN>
static fmq_main_fn quality_pick_out(unsigned nout) {
    switch (nout) {
        case 1: return fmq_main_kernel<NIN, 1>;
        case 2: return fmq_main_kernel<NIN, 2>;
        case 3: return fmq_main_kernel<NIN, 3>;
        case 4: return fmq_main_kernel<NIN, 4>;
        case 5: return fmq_main_kernel<NIN, 5>;
        case 6: return fmq_main_kernel<NIN, 6>;
        default: return nullptr;
    }
}

/* by filter DEGREE: an odd degree ends in a first-order section, which the kernels exploit */
static fmq_main_fn quality_pick_main(unsigned nin, unsigned nout) {
    switch (nin) {
        case 0: return quality_pick_out<0>(nout);
        case 1: return quality_pick_out<1>(nout);
        case 2: return quality_pick_out<2>(nout);
        case 3: return quality_pick_out<3>(nout);
        case 4: return quality_pick_out<4>(nout);
        case 5: return quality_pick_out<5>(nout);
        case 6: return quality_pick_out<6>(nout);
        default: return nullptr;
    }
}

/* one launch group: `n` samples (< 2^31) that continue the stream; *nout = floats written */
static int quality_launch(fmd_ctx *c, const uint8_t *d_in, unsigned n, float *d_out, cudaStream_t s, size_t *nout) {
    fmq_ctx *q = c->q;
    FmqGeom g;
    g.n = n;
    g.ntiles = (n + FMQ_TILE - 1) / FMQ_TILE;
    g.decim = q->decim;
    g.r0 = (unsigned) ((q->decim - q->pos % q->decim) % q->decim);
    g.n_out = (unsigned) quality_out_count(q->pos, n, q->decim);
    *nout = g.n_out;
    int rc = quality_reserve_tiles(q, g.ntiles);
    if (rc) return rc;
    const unsigned len_last = n - (g.ntiles - 1u) * FMQ_TILE;
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

    fmq_main_fn fn = quality_pick_main(q->lin ? c->cfg.in_degree : 0u, c->cfg.out_degree);
    if (!fn) return fmd_set_error(FMD_EINVAL, "no quality kernel for degrees %u/%u", c->cfg.in_degree, c->cfg.out_degree);
    /* shared memory for the kept samples of one tile */
    const unsigned keep_cap = FMQ_TILE / q->decim + 2u;
    const unsigned keep = keep_cap < FMQ_TILE ? keep_cap : FMQ_TILE;
    const size_t smem = (size_t) (keep + FMQ_THREADS) * sizeof(float); /* one pad word per owning thread */
    cudaEvent_t *pe = nullptr;
    if (c->profile) {
        if (c->pev_used == c->pev_cap) {
            rc = prof_flush(c);
            if (rc) return rc;
        }
        pe = &c->pev[2 * c->pev_used++];
        FMD_CUDA(cudaEventRecord(pe[0], s));
    }
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
    fmq_matpow_f(&q->fout.des, (unsigned long long) FMQ_TILE * runlen, prun.m);
    fmq_matpow_f(&q->fout.des, len_last, pl.m);
    const uint8_t *last = d_in + 2ull * (n - 1u);
    /* zero-input response of the tile carry: only the leading zir_len offsets of every tile */
    const unsigned zl = q->fout.zir_len < FMQ_TILE ? q->fout.zir_len : FMQ_TILE;
    const dim3 zgrid(g.ntiles, (zl / q->decim + 1u + 255u) / 256u);
    switch (q->lout) {
        case 1:
            fmq_scan_tiles_kernel<2, 1><<<1, FMQ_SCAN_THREADS, 0, s>>>(q->d_eout, q->d_qout, q->d_state->out_q, q->fout.c_tile, q->fout.lane_tile(), prun, pl, g.ntiles, runlen, last, q->d_state->s_prev);
            if (g.n_out) fmq_zir_kernel<2><<<zgrid, 256, 0, s>>>(d_out, g, zl, q->fout.ckn(), q->d_qout);
            break;
        case 2:
            fmq_scan_tiles_kernel<4, 1><<<1, FMQ_SCAN_THREADS, 0, s>>>(q->d_eout, q->d_qout, q->d_state->out_q, q->fout.c_tile, q->fout.lane_tile(), prun, pl, g.ntiles, runlen, last, q->d_state->s_prev);
            if (g.n_out) fmq_zir_kernel<4><<<zgrid, 256, 0, s>>>(d_out, g, zl, q->fout.ckn(), q->d_qout);
            break;
        default:
            fmq_scan_tiles_kernel<6, 1><<<1, FMQ_SCAN_THREADS, 0, s>>>(q->d_eout, q->d_qout, q->d_state->out_q, q->fout.c_tile, q->fout.lane_tile(), prun, pl, g.ntiles, runlen, last, q->d_state->s_prev);
            if (g.n_out) fmq_zir_kernel<6><<<zgrid, 256, 0, s>>>(d_out, g, zl, q->fout.ckn(), q->d_qout);
            break;
    }
    FMD_CUDA(cudaGetLastError());
    c->launches += g.n_out ? 3 : 2;
    q->pos += n;
    return FMD_OK;
}

static int quality_demod_device(fmd_ctx *c, const uint8_t *d_in, size_t nbytes, float *d_out, cudaStream_t s) {
    if (nbytes & 1u) return fmd_set_error(FMD_EINVAL, "quality mode takes whole IQ samples (an even number of bytes)");
    if (((uintptr_t) d_in % 16) || ((uintptr_t) d_out % 4))
        return fmd_set_error(FMD_EINVAL, "quality mode needs a 16-byte aligned input and a 4-byte aligned output");
    size_t left = nbytes / 2;
    const size_t step = (size_t) 1 << 30; /* a multiple of the tile: later groups stay 16-byte aligned */
    while (left) {
        const unsigned n = (unsigned) (left < step ? left : step);
        size_t nout = 0;
        int rc = quality_launch(c, d_in, n, d_out, s, &nout);
        if (rc) return rc;
        d_in += 2ull * n;
        d_out += nout;
        left -= n;
    }
    return FMD_OK;
}

extern "C" int fmd_reset(fmd_ctx *c) {
    if (!c) return fmd_set_error(FMD_EINVAL, "null argument");
    if (!c->q) return FMD_OK;
    FMD_CUDA(cudaSetDevice(c->device));
    FMD_CUDA(cudaDeviceSynchronize());
    FMD_CUDA(cudaMemset(c->q->d_state, 0, sizeof(FmqState)));
    c->q->pos = 0;
    return FMD_OK;
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
        rc = quality_demod_device(c, d_in, nbytes, d_out, s);
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
    cudaEvent_t a, b;
    FMD_CUDA(cudaEventCreate(&a));
    FMD_CUDA(cudaEventCreate(&b));
    FMD_CUDA(cudaEventRecord(a, s));
    int rc = FMD_OK;
    for (int i = 0; i < reps && !rc; ++i) rc = fmd_demod_device(c, d_in, nbytes, d_out, s);
    if (!rc) {
        FMD_CUDA(cudaEventRecord(b, s));
        FMD_CUDA(cudaEventSynchronize(b));
        FMD_CUDA(cudaEventElapsedTime(ms, a, b));
    }
    cudaEventDestroy(a);
    cudaEventDestroy(b);
    return rc;
}

/* --------------------------------------------------------------------- ring */

extern "C" int fmd_ring_create(fmd_ctx *c, unsigned slots, size_t slot_bytes) {
    if (!c || slots == 0 || slots > 64 || slot_bytes == 0) return fmd_set_error(FMD_EINVAL, "bad ring geometry");
    if (slot_bytes % c->cfg.chunk_bytes)
        return fmd_set_error(FMD_EINVAL, "slot_bytes must be a multiple of chunk_bytes");
    FMD_CUDA(cudaSetDevice(c->device));
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

extern "C" int fmd_submit(fmd_ctx *c, unsigned slot, const uint8_t *in, size_t nbytes, float *out) {
    if (!c || slot >= c->nslots || !in || !out) return fmd_set_error(FMD_EINVAL, "bad slot or null buffer");
    if (nbytes > c->slot_bytes) return fmd_set_error(FMD_EINVAL, "nbytes exceeds slot_bytes");
    fmd_slot &s = c->slots[slot];
    if (s.busy) return fmd_set_error(FMD_EBUSY, "slot %u still in flight: call fmd_wait first", slot);
    FMD_CUDA(cudaSetDevice(c->device));
    const size_t nout = fmd_out_count(c, nbytes);
    /* the slot's device buffers are free: the previous occupant's D2H completed in fmd_wait */
    FMD_CUDA(cudaMemcpyAsync(s.d_in, in, nbytes, cudaMemcpyHostToDevice, c->s_in));
    FMD_CUDA(cudaEventRecord(s.ev_in, c->s_in));
    FMD_CUDA(cudaStreamWaitEvent(c->s_k, s.ev_in, 0));
    int rc = fmd_demod_device(c, s.d_in, nbytes, s.d_out, c->s_k);
    if (rc) return rc;
    FMD_CUDA(cudaEventRecord(s.ev_k, c->s_k));
    FMD_CUDA(cudaStreamWaitEvent(c->s_out, s.ev_k, 0));
    FMD_CUDA(cudaMemcpyAsync(out, s.d_out, nout * sizeof(float), cudaMemcpyDeviceToHost, c->s_out));
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
    if (c->q && c->q->lin) { /* a look-back that gave up waiting raises a device-side flag */
        unsigned err = 0;
        FMD_CUDA(cudaMemcpy(&err, &c->q->d_state->error, sizeof err, cudaMemcpyDeviceToHost));
        if (err) return fmd_set_error(FMD_ECUDA, "quality pipeline: a tile gave up waiting for its predecessor's carry");
    }
    return FMD_OK;
}

/* ------------------------------------------------------- host-buffer wrapper */

extern "C" int fmd_demod_host(fmd_ctx *c, const uint8_t *in, size_t nbytes, float *out, size_t *nout) {
    if (!c || (!in && nbytes) || (!out && nbytes)) return fmd_set_error(FMD_EINVAL, "null argument");
    if (nout) *nout = 0;
    const size_t B = c->cfg.chunk_bytes;
    /* slice size: ~32 MiB rounded to whole chunks, 3 slices in flight (measured on the B200 box:
     * 22.2 GS/s against 20.8 GS/s with 16 MiB x 4; PCIe Gen5 x16 moves 47-49 GB/s each way when
     * both directions are busy, profiles/r01_pcie_ceiling.txt) */
    size_t slice = (size_t) 32 << 20;
    slice = slice / B * B;
    if (slice == 0) slice = B;
    if (slice > nbytes) slice = (nbytes + B - 1) / B * B;
    if (slice == 0) slice = B;
    if (c->nslots < 3 || c->slot_bytes < slice) {
        int rc = fmd_ring_create(c, 3, slice);
        if (rc) return rc;
    }
    const unsigned ns = c->nslots;
    size_t off = 0, out_off = 0;
    unsigned n = 0;
    while (off < nbytes) {
        const unsigned slot = n % ns;
        const size_t len = nbytes - off < c->slot_bytes ? nbytes - off : c->slot_bytes;
        int rc = fmd_wait(c, slot);
        const size_t produced = fmd_out_count(c, len); /* before the submit moves the stream on */
        if (!rc) rc = fmd_submit(c, slot, in + off, len, out + out_off);
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
