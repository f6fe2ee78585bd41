/*
 * fmd_ctx.h -- the context behind include/fmdemod.h, shared by the translation units of
 * libfmdemod: fmdemod.cu (C ABI, reference-compatible path, streams, ring) and
 * fmd_quality_host.cu (host side of the opt-in quality pipeline).  C++ / CUDA only.
 */
#ifndef FMD_CTX_H
#define FMD_CTX_H

#include <cuda_runtime.h>
#include <stdint.h>

#include "fmd_internal.h"
#include "fmd_quality.h"
#include "fmd_types.h"

#define FMD_CUDA(call)                                                                              \
    do {                                                                                            \
        cudaError_t e__ = (call);                                                                   \
        if (e__ != cudaSuccess)                                                                     \
            return fmd_set_error(FMD_ECUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), \
                                 __FILE__, __LINE__);                                               \
    } while (0)

/* Every entry point works on its context's device and leaves the calling thread's current device
 * as it found it (a host that drives several GPUs from one thread must not be surprised). */
struct FmdDeviceScope {
    int prev = -1;
    bool switched = false;
    explicit FmdDeviceScope(int device) {
        if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
        if (prev != device) switched = cudaSetDevice(device) == cudaSuccess;
    }
    ~FmdDeviceScope() {
        if (switched && prev >= 0) cudaSetDevice(prev);
    }
    FmdDeviceScope(const FmdDeviceScope &) = delete;
    FmdDeviceScope &operator=(const FmdDeviceScope &) = delete;
};

struct fmd_slot {
    uint8_t *d_in = nullptr;
    float *d_out = nullptr;
    cudaEvent_t ev_in = nullptr;   /* H2D landed              */
    cudaEvent_t ev_k = nullptr;    /* kernels done            */
    cudaEvent_t ev_out = nullptr;  /* D2H landed              */
    bool busy = false;
};

struct fmq_ctx;  /* FMD_FLAG_QUALITY state, fmd_quality_host.cu */
struct fmq1_ctx; /* -q1 (correctIq) state, fmd_quality_host.cu */

struct fmd_ctx {
    fmd_config cfg;
    fmq_ctx *q = nullptr;
    fmq1_ctx *q1 = nullptr;
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

/* per-launch event timing (fmd_profile_kernels): *pe = pair of events to record around the
 * dominant kernel of a launch, or nullptr when profiling is off */
int fmd_prof_events(fmd_ctx *c, cudaEvent_t **pe);

/* the streaming kernel's instantiations, compiled in groups (fmd_fast_inst.cu); a8 = 1: 8-term atan */
typedef void (*fmd_fast_fn)(const uint8_t *, float *, const FmdGeom, const FmdStencil, const FmdSections *);
fmd_fast_fn fmd_pick_n01(unsigned nti, unsigned nto, unsigned a8);
fmd_fast_fn fmd_pick_n23(unsigned nti, unsigned nto, unsigned a8);
fmd_fast_fn fmd_pick_n45(unsigned nti, unsigned nto, unsigned a8);
fmd_fast_fn fmd_pick_q3(unsigned nti, unsigned nto, unsigned a8);
fmd_fast_fn fmd_pick_float(unsigned nti, unsigned nto, unsigned a8); /* float-input kernel (-q1); takes const float* as its first argument */

/* ---- quality pipeline, fmd_quality_host.cu ---- */
int fmq_setup(fmd_ctx *c, unsigned decim);
void fmq_free(fmd_ctx *c);
size_t fmq_out_count(const fmd_ctx *c, size_t nsamp);
size_t fmq_max_out_count(const fmd_ctx *c, size_t nsamp);
int fmq_demod_device(fmd_ctx *c, const uint8_t *d_in, size_t nbytes, float *d_out, cudaStream_t s);
int fmq_reset(fmd_ctx *c);
/* 0, or FMD_ECUDA if a look-back gave up waiting (reads a device-side flag: synchronises) */
int fmq_check_error(fmd_ctx *c);
int fmq_seek(fmd_ctx *c, unsigned long long pos, cudaStream_t s);
size_t fmq_out_count_at(const fmd_ctx *c, unsigned long long pos, size_t nsamp);
unsigned long long fmq_warmup_samples(const fmd_ctx *c);

/* ---- -q1 (correctIq): offsets that live on from chunk to chunk, as a scan; fmd_quality_host.cu ---- */
int fmq1_setup(fmd_ctx *c);
void fmq1_free(fmd_ctx *c);
int fmq1_reset(fmd_ctx *c);
/* conditions `nchunks` whole chunks (continuing the stream) into floats, one per input byte; *sbuf is
 * owned by the context and valid until the next call on it */
int fmq1_condition(fmd_ctx *c, const uint8_t *d_in, unsigned long long nchunks, cudaStream_t s, const float **sbuf);
int fmq1_fused(fmd_ctx *c, const uint8_t *d_in, unsigned long long nchunks, float *d_out, cudaStream_t s,
               const float **heads, unsigned *heads_stride_bytes, bool *done);

#endif /* FMD_CTX_H */
