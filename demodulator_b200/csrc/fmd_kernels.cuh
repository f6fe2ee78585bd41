/*
 * fmd_kernels.cuh -- sm_100a device code of the FM demodulation hot path.
 *
 * One launch per call: fmd_fast_kernel<NTI, NTO, AT> (input taps, output taps, atan terms).
 *
 *  streaming CTAs -- the bandwidth part.  One 16-byte load per lane per pass (8 complex u8
 *      samples = 4 discriminator pairs), one 16-byte store per lane per pass (4 floats); every
 *      intermediate lives in registers.  A warp walks a contiguous segment of a chunk; the
 *      stencil neighbours that fall outside a lane come from the next lane by shuffle, and the
 *      ones that fall outside the warp come from the SAME warp's next pass (the output stage
 *      runs one pass behind the discriminator), so no value is computed twice, nothing is
 *      exchanged through memory and there is no block-level barrier.
 *      Computes, per chunk-local output k:
 *          s[c]  = (int8)(u8 - 127)                        shiftOrigin        matrix.c:82-98
 *                  (-q3: u * (2/255) - 1, normalizeInput, matrix.c:100-118)
 *          z[c]  = dc + sum_t tin[t]  * s[c+t]   (-L only)  applyComplexFilter filter.c:233-259
 *                  (without -L, -q2's highpass sweep, matrix.c:142-157, takes this place)
 *                  (evaluated as z/tin[0]: the common factor cancels in the atan2 below)
 *          x[k]  = atan2(Im, Re)(z[2k] * conj(z[2k+1]))    fmDemod            matrix.c:159-176
 *          y[k]  = dco + sum_t tout[t] * x[k+t]            applyFilter        filter.c:212-231
 *      with zero extension past the chunk end, i.e. the steady-state form of the reference's
 *      section sweep (fmd_fold_taps in fmd_design.c).
 *
 *  exact CTAs -- the first CTAs of the same launch (or fmd_exact_kernel stand-alone): one thread
 *      per output evaluates the reference's section recurrence op for op (separately rounded
 *      mul/add, literal "+1", head rules, and the x|y_c|y_o arena aliasing that seeds the first
 *      outputs of every chunk when -L is on).  They produce the first few outputs of every
 *      chunk (where the sweep is not in steady state), ragged remainders, chunk sizes the
 *      streaming part does not take, and everything under FMD_FLAG_EXACT (audits).
 *
 *  fmd_exact_tiled_kernel -- the same cones for WHOLE chunks, staged through shared memory (audits
 *      under FMD_FLAG_EXACT, and the designs the streaming kernels have no instantiation for: more
 *      than 5 folded taps, a conditioner under an input filter of more than 3 taps).
 *
 *  fmd_fastf_kernel -- the streaming pass over FLOAT samples (the three-pass form of -q1,
 *      fmd_correctiq.cuh; the fused form is fmd_q1_fused.cuh).
 */
#ifndef FMD_KERNELS_CUH
#define FMD_KERNELS_CUH

#include <cuda_runtime.h>
#include <stdint.h>

#include "fmdemod.h"

#include "fmd_math.cuh"
#include "fmd_types.h"

/* -------------------------------------------------------------- exact kernel */

#define FMD_CONE_W (FMD_MAX_SECTIONS / 2 + 3)

/*
 * Value of index k after the reference's section sweep over a lane of n samples
 * (filter.c:221-230 / 242-258), from the dependency cone of k alone:
 *   Y_s[k] = ((b0*Y_{s+1}[k] + b1*Y_{s+2}[k+1]) [+ 1]) - ((a0 + a1*x[k]) + a2*x[k+1])
 * if section s rewrites k (L+s <= k <= n-1+L+s), else Y_{s+1}[k]; every product
 * and sum rounded separately.  getx(i) = lane input, init(i) = what the y region
 * held before the sweep.
 */
template <class GetX, class Init>
__device__ float fmd_sweep_cone(long long k, long long n, int L, const float (*sos)[6], bool bias,
                                const GetX &getx, const Init &init) {
    /* The cone of index k is W = L/2 + 3 wide.  The arrays are walked over their full compile-time
     * width with the extra columns masked, so that they are indexed statically and live in registers
     * (indexed up to the run-time W they were a 384-byte stack frame per thread and 435 MB of
     * local-memory write traffic per cfg2 launch, round-1 review). */
    float p1[FMD_CONE_W], p2[FMD_CONE_W], xv[FMD_CONE_W + 1];
    const int W = L / 2 + 3;
#pragma unroll
    for (int j = 0; j <= FMD_CONE_W; ++j) xv[j] = j <= W ? getx(k + j) : 0.0f;
#pragma unroll
    for (int j = 0; j < FMD_CONE_W; ++j) p1[j] = p2[j] = j < W ? init(k + j) : 0.0f;
    for (int s = L - 1; s >= 0; --s) {
        const float b0 = sos[s][0], b1 = sos[s][1], a0 = sos[s][3], a1 = sos[s][4], a2 = sos[s][5];
        const long long lo = L + s, hi = n - 1 + L + s;
        float nw[FMD_CONE_W];
#pragma unroll
        for (int j = 0; j < FMD_CONE_W; ++j) {
            const long long kk = k + j;
            float val = p1[j];
            if (j + 1 < FMD_CONE_W && j + 1 < W && kk >= lo && kk <= hi) {
                float acc = __fadd_rn(__fmul_rn(b0, p1[j]), __fmul_rn(b1, p2[j + 1 < FMD_CONE_W ? j + 1 : j]));
                if (bias) acc = __fadd_rn(acc, 1.0f);
                const float rhs = __fadd_rn(__fadd_rn(a0, __fmul_rn(a1, xv[j])), __fmul_rn(a2, xv[j + 1]));
                val = __fsub_rn(acc, rhs);
            }
            nw[j] = val;
        }
#pragma unroll
        for (int j = 0; j < FMD_CONE_W; ++j) {
            p2[j] = p1[j];
            p1[j] = nw[j];
        }
    }
    return p1[0];
}

struct FmdZeroInit {
    __device__ float operator()(long long) const { return 0.0f; }
};

/* shifted input lane q (0 = I, 1 = Q) of one chunk, zero past its end */
struct FmdShiftedLane {
    const uint8_t *cin;
    long long nc;
    int q;
    int conditioner;
    __device__ float operator()(long long c) const {
        if (c >= nc) return 0.0f;
        /* -q1 (correctIq): the samples were conditioned by fmd_correctiq_kernel; `cin` then points at
         * floats, one per input byte, same layout */
        if (conditioner == 1) return reinterpret_cast<const float *>(cin)[2 * c + q];
        const uint8_t u = cin[2 * c + q];
        /* -q3, matrix.c:100-118: (float)((double)((float)u * (float)(2/255)) - 1.0); the double
         * subtraction of a float and 1 is exact, so one rounded float subtraction is the same */
        if (conditioner == 3) return __fsub_rn(__fmul_rn((float) u, (float) (2. / 255.)), 1.0f);
        return (float) (int8_t) (u - 127); /* matrix.c:82-98 */
    }
};

/*
 * What the conditioner leaves in the x region, lane q, at complex index c.
 *   -q0 / -q3: the shifted / normalised sample, zero past the end of the chunk.
 *   -q2 (highpassDc, matrix.c:142-157): shiftOrigin, then a degree-3 Butterworth highpass at 1 Hz
 *       swept like any complex filter (filter.c:242-258) from a scratch buffer whose slack is zero.
 *       Indices nc .. nc+2 are what that sweep WRITES past its output region -- the start of the
 *       region the next stage accumulates into -- and 0 beyond (no section reaches them).
 */
struct FmdConditionedLane {
    const uint8_t *cin;
    long long nc;
    const FmdSections *sec;
    int q;
    __device__ float operator()(long long c) const {
        FmdShiftedLane sl{cin, nc, q, sec->conditioner};
        if (sec->conditioner != 2) return sl(c);
        return fmd_sweep_cone(c, nc, 2, sec->hp, q == 0, sl, FmdZeroInit());
    }
};

/* initial contents of the region a stage accumulates into, lane q, complex index c: what the
 * stage before it spilled there (only -q2 spills ahead of the input filter) */
struct FmdConditionerSpill {
    FmdConditionedLane x;
    __device__ float operator()(long long c) const {
        if (x.sec->conditioner != 2 || c >= 3) return 0.0f;
        return x(x.nc + c);
    }
};

/* final y_c lane value at complex index c < nc (input lowpass on), or the conditioned sample.
 * Past the end of the chunk the input sweep reads the head of its own output region, which it
 * never rewrites: zero, or what -q2 spilled there -- FmdConditionedLane at c >= nc is exactly that. */
struct FmdComplexLane {
    const uint8_t *cin;
    long long nc;
    const FmdSections *sec;
    int q;
    __device__ float operator()(long long c) const {
        FmdConditionedLane xl{cin, nc, sec, q};
        if (sec->lin == 0) return xl(c);
        return fmd_sweep_cone(c, nc, sec->lin, sec->in, q == 0, xl, FmdConditionerSpill{xl});
    }
};

/* what the complex sweep reads as x[nc + t] while it produces its own spill: the already-final
 * head of y_c */
struct FmdAliasedLane {
    FmdComplexLane z;
    __device__ float operator()(long long c) const { return z(c - z.nc); }
};

/* discriminator output j of the chunk, zero past its end (demodRet slack, matrix.c:184) */
struct FmdDemodX {
    const uint8_t *cin;
    long long nc, n_out;
    const FmdSections *sec;
    __device__ float operator()(long long j) const {
        if (j >= n_out) return 0.0f;
        FmdComplexLane zi{cin, nc, sec, 0}, zq{cin, nc, sec, 1};
        const float a = zi(2 * j), b = zq(2 * j), c = zi(2 * j + 1), d = zq(2 * j + 1);
        const float re = __fadd_rn(__fmul_rn(a, c), __fmul_rn(b, d));
        const float im = __fadd_rn(__fmul_rn(-a, d), __fmul_rn(b, c));
        return fmd_atan2(im, re);
    }
};

/* initial contents of the output accumulator region: what the LAST complex sweep spilled past
 * the end of its output region (matrix.c:260-267 + filter.c:242-258; with -q2 and no -L that is
 * the conditioner's own sweep), zero elsewhere */
struct FmdSpillInit {
    const uint8_t *cin;
    long long nc;
    const FmdSections *sec;
    __device__ float operator()(long long f) const {
        const int L = sec->lin;
        const int q = (int) (f & 1);
        if (L == 0) return FmdConditionerSpill{FmdConditionedLane{cin, nc, sec, q}}(f >> 1);
        if (f >= 4 * L - 2) return 0.0f;
        FmdAliasedLane al{FmdComplexLane{cin, nc, sec, q}};
        return fmd_sweep_cone(nc + (f >> 1), nc, L, sec->in, q == 0, al, FmdZeroInit());
    }
};

/* output k of one chunk, op for op */
static __device__ __noinline__ float fmd_exact_output(const uint8_t *cin, unsigned chunk_out, const FmdSections *sec,
                                               unsigned k) {
    const long long n_out = chunk_out, nc = 2ll * chunk_out;
    FmdDemodX dx{cin, nc, n_out, sec};
    FmdSpillInit si{cin, nc, sec};
    return fmd_sweep_cone((long long) k, n_out, sec->lout, sec->out, true, dx, si);
}

/*
 * Thread `tid` of the exact path: outputs [0, head) and [fast_out, chunk_out) of every chunk
 * (the chunk heads, where the sweep is not in steady state, and the ragged remainder).
 */
__device__ __forceinline__ void fmd_exact_thread(const uint8_t *__restrict__ in, float *__restrict__ out,
                                                 const FmdGeom &g, const FmdSections *sec,
                                                 unsigned long long tid) {
    const unsigned per_chunk = g.head + (g.chunk_out - g.fast_out);
    if (tid >= (unsigned long long) g.nchunks * per_chunk) return;
    const unsigned chunk = (unsigned) (tid / per_chunk);
    const unsigned r = (unsigned) (tid - (unsigned long long) chunk * per_chunk);
    const unsigned k = r < g.head ? r : g.fast_out + (r - g.head);
    if (k >= g.chunk_out) return;
    out[(unsigned long long) chunk * g.chunk_out + k] =
        fmd_exact_output(in + (unsigned long long) chunk * g.chunk_stride, g.chunk_out, sec, k);
}

#ifdef FMD_WITH_EXACT_KERNELS /* defined by the one translation unit that launches them (fmdemod.cu) */
/* stand-alone exact kernel: every output of every chunk (geometry: head = chunk_out, fast_out = chunk_out) */
__global__ void __launch_bounds__(FMD_BLOCK)
fmd_exact_kernel(const uint8_t *__restrict__ in, float *__restrict__ out, const FmdGeom g,
                 const FmdSections *__restrict__ sec) {
    fmd_exact_thread(in, out, g, sec, (unsigned long long) blockIdx.x * blockDim.x + threadIdx.x);
}


/*
 * The same op-for-op evaluation for WHOLE chunks (FMD_FLAG_EXACT, -q1, a conditioner stacked under
 * -L, chunk sizes the streaming kernel does not take), staged through shared memory.  A CTA owns
 * FMD_XT consecutive outputs of one chunk: the conditioned lanes it needs are evaluated once into
 * shared memory, then the input-filtered lanes from those, then the discriminator outputs, then
 * the output sweep -- the very same cones (fmd_sweep_cone) with the very same operations, but each
 * intermediate value is computed once per CTA instead of once per consumer per consumer per
 * consumer.  Measured on B200: 4-8x the one-thread-per-output form above.
 */
#define FMD_XT 256                                           /* outputs per CTA                 */
#define FMD_XX (FMD_XT + FMD_CONE_W + 1)                     /* discriminator outputs it needs  */
#define FMD_XZ (2 * FMD_XX)                                  /* complex samples they come from  */
#define FMD_XC (FMD_XZ + FMD_CONE_W + 1)                     /* conditioned samples under those */

struct FmdStagedLane { /* lane values staged in shared memory, from complex index `base` on */
    const float *buf;
    long long base;
    __device__ float operator()(long long c) const { return buf[c - base]; }
};

__global__ void __launch_bounds__(FMD_XT)
fmd_exact_tiled_kernel(const uint8_t *__restrict__ in, float *__restrict__ out, const FmdGeom g,
                       const FmdSections *__restrict__ sec) {
    __shared__ float xl[2][FMD_XC]; /* conditioned lanes I, Q                 */
    __shared__ float zl[2][FMD_XZ]; /* after the input filter                 */
    __shared__ float xd[FMD_XX];    /* discriminator outputs                  */
    const unsigned tiles_per_chunk = (g.chunk_out + FMD_XT - 1u) / FMD_XT;
    const unsigned chunk = blockIdx.x / tiles_per_chunk;
    const unsigned k0 = (blockIdx.x - chunk * tiles_per_chunk) * FMD_XT;
    const uint8_t *cin = in + (unsigned long long) chunk * g.chunk_stride;
    const long long n_out = g.chunk_out, nc = 2ll * g.chunk_out, c0 = 2ll * k0;
    const bool filtered = sec->lin != 0;

    for (unsigned i = threadIdx.x; i < 2u * FMD_XC; i += FMD_XT) {
        const int q = (int) (i & 1u);
        xl[q][i >> 1] = FmdConditionedLane{cin, nc, sec, q}(c0 + (long long) (i >> 1));
    }
    __syncthreads();
    if (filtered) {
        for (unsigned i = threadIdx.x; i < 2u * FMD_XZ; i += FMD_XT) {
            const int q = (int) (i & 1u);
            const FmdStagedLane sx{xl[q], c0};
            zl[q][i >> 1] = fmd_sweep_cone(c0 + (long long) (i >> 1), nc, sec->lin, sec->in, q == 0, sx,
                                           FmdConditionerSpill{FmdConditionedLane{cin, nc, sec, q}});
        }
        __syncthreads();
    }
    const float(*zz)[FMD_XZ] = zl;
    for (unsigned i = threadIdx.x; i < FMD_XX; i += FMD_XT) {
        const long long j = (long long) k0 + i;
        float x = 0.0f; /* zero past the end of the chunk (demodRet slack, matrix.c:184) */
        if (j < n_out) {
            const float a = filtered ? zz[0][2 * i] : xl[0][2 * i], b = filtered ? zz[1][2 * i] : xl[1][2 * i];
            const float c = filtered ? zz[0][2 * i + 1] : xl[0][2 * i + 1], d = filtered ? zz[1][2 * i + 1] : xl[1][2 * i + 1];
            const float re = __fadd_rn(__fmul_rn(a, c), __fmul_rn(b, d));
            const float im = __fadd_rn(__fmul_rn(-a, d), __fmul_rn(b, c));
            x = fmd_atan2(im, re);
        }
        xd[i] = x;
    }
    __syncthreads();
    const unsigned k = k0 + threadIdx.x;
    if (k < g.chunk_out) {
        const FmdStagedLane sx{xd, (long long) k0};
        out[(unsigned long long) chunk * g.chunk_out + k] =
            fmd_sweep_cone((long long) k, n_out, sec->lout, sec->out, true, sx, FmdSpillInit{cin, nc, sec});
    }
}

#endif /* FMD_WITH_EXACT_KERNELS */

/* --------------------------------------------------------------- fast kernel */

#define FMD_FILL 0x7f7f7f7fu /* four bytes that shift to 0.0: the zero extension past a chunk */

/*
 * The arithmetic of one pass once the lane's samples are floats: vi[k] = (I_k, I_{k+4}),
 * vq[k] = (Q_k, Q_{k+4}) for k = 0 .. 3 + (NTI > 1 ? NTI - 1 : 0) (entries from 4 on reach into the
 * next lane's samples).  Input stencil, four discriminator outputs into xcur, output stencil over
 * the PREVIOUS pass's xprev, one 16-byte store.  Shared by the byte-input kernel (fmd_pass unpacks
 * into vi / vq) and the float-input one (-q1: the samples arrive conditioned).
 */
template <int NTI, int NTO, int AT, bool LITE, int NK>
__device__ __forceinline__ void fmd_pass_core(const fmd_f2 (&vi)[NK], const fmd_f2 (&vq)[NK], const float (&xprev)[4],
                                              float (&xcur)[4], const bool zero_x, float *__restrict__ dst,
                                              const bool do_store, const unsigned lane, const unsigned nxt_lane,
                                              const FmdStencil &st) {
    constexpr int NB = NTO - 1; /* discriminator outputs needed from the next lane */
#pragma unroll
    for (int j = 0; j < (LITE ? 1 : 2); ++j) { /* outputs j and j+2 */
        fmd_f2 a, b, c, d;
        if (NTI == 0) {
            a = vi[2 * j]; b = vq[2 * j]; c = vi[2 * j + 1]; d = vq[2 * j + 1];
        } else {
            /* The input stencil is applied up to its (positive) common factor tin[0]^2, which the
             * atan2 of the product does not see: z'[c] = s[c] + sum_{t>=1} (tin[t]/tin[0]) s[c+t],
             * plus dcq/tin[0] on the Q lane (the host normalises; dci is 0 for every design). */
            const fmd_f2 dq = fmd_bc(st.dcq);
            a = vi[2 * j];
            b = fmd_add2(vq[2 * j], dq);
            c = vi[2 * j + 1];
            d = fmd_add2(vq[2 * j + 1], dq);
#pragma unroll
            for (int t = 1; t < NTI; ++t) {
                const fmd_f2 tap = fmd_bc(st.tin[t]);
                a = fmd_fma2(tap, vi[2 * j + t], a);
                b = fmd_fma2(tap, vq[2 * j + t], b);
                c = fmd_fma2(tap, vi[2 * j + 1 + t], c);
                d = fmd_fma2(tap, vq[2 * j + 1 + t], d);
            }
        }
        /* s0 * conj(s1).  The fused forms keep the reference's signed-zero results: a sum of
         * two exact-zero products takes the same sign fused or not. */
        const fmd_f2 re = fmd_fma2(a, c, fmd_mul2(b, d));
        const fmd_f2 im = fmd_fma2(fmd_neg2(a), d, fmd_mul2(b, c));
        fmd_atan2_x2<AT>(im, re, xcur[j], xcur[j + 2]);
    }
    if (LITE) xcur[1] = xcur[3] = 0.0f; /* drain pass: only lane 0's first output is ever read */
    if (zero_x) { /* past the end of the chunk the discriminator output is exactly 0 */
        xcur[0] = 0.0f; xcur[1] = 0.0f; xcur[2] = 0.0f; xcur[3] = 0.0f;
    }

    float xn[4 + (NB > 0 ? NB : 1)];
    xn[0] = xprev[0]; xn[1] = xprev[1]; xn[2] = xprev[2]; xn[3] = xprev[3];
#pragma unroll
    for (int t = 0; t < NB; ++t)
        xn[4 + t] = __shfl_sync(0xffffffffu, lane == 0 ? xcur[t] : xprev[t], nxt_lane);
    float4 y;
    float *yy = reinterpret_cast<float *>(&y);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        float acc = st.dco;
#pragma unroll
        for (int t = 0; t < NTO; ++t) acc = fmaf(st.tout[t], xn[j + t], acc);
        yy[j] = acc;
    }
    if (do_store) *reinterpret_cast<float4 *>(dst) = y;
}

/*
 * One pass of one lane: unpack `cur` (8 complex samples), input lowpass, four
 * discriminator outputs into xcur; then the output lowpass over the PREVIOUS
 * pass's xprev (neighbours: next lane's xprev, or lane 0's xcur for lane 31)
 * and one 16-byte store (predicate `do_store`: the very first pass of a segment has no
 * previous pass, and the head of a chunk belongs to the exact path).  `nxt` is the first
 * 8 bytes of the lane's raw input of the following pass: lane 0's copy is what lane 31 needs
 * as its input-filter halo.  LITE: only outputs 0 and 2 are computed (the drain pass).
 */
#ifndef FMD_HALO_CARRY
#define FMD_HALO_CARRY 1 /* 0: re-bias the shuffled raw halo word every pass (round-1 form) */
#endif

/* Does the pass keep the biased first word of the NEXT pass (`w0c`) from pass to pass?  The input
 * filter's halo is the next lane's first word -- for lane 31 lane 0's first word of the next pass.
 * That word is biased (3 instructions) anyway when its pass comes; biasing it one pass early and
 * handing it on lets the halo be a shuffle of ALREADY biased words: 3 instructions per pass less
 * than shuffling the raw word and biasing the result. */
template <int NTI, int COND>
struct FmdCarry {
    static constexpr bool value = FMD_HALO_CARRY && COND == 0 && NTI == 2; /* measured there; with 3 taps it costs a spilled register */
};

template <int NTI, int NTO, int AT, bool LITE = false, int COND = 0>
__device__ __forceinline__ void fmd_pass(const uint4 &cur, const uint2 &nxt, const float (&xprev)[4],
                                         float (&xcur)[4], const bool zero_x, float *__restrict__ dst,
                                         const bool do_store, const unsigned lane, const unsigned nxt_lane,
                                         const FmdStencil &st, uint32_t &w0c, const bool hz = false) {
    constexpr int HC = NTI > 1 ? NTI - 1 : 0; /* complex samples needed from the next lane */
    constexpr int HW = (HC + 1) / 2;          /* ... as 4-byte words (2 complex each)      */
    constexpr bool CARRY = FmdCarry<NTI, COND>::value;

    /* biased words: own four, then the halo from the next lane (lane 31: lane 0's next pass) */
    static_assert(COND == 0 || COND == 3, "conditioners of the streaming kernel: shiftOrigin (and what folds into its stencil), normalizeInput");
    uint32_t w[6];
    if (COND == 3) { /* normalizeInput works on the raw bytes */
        w[0] = cur.x; w[1] = cur.y; w[2] = cur.z; w[3] = cur.w;
    } else {
        w[0] = CARRY ? w0c : fmd_bias_bytes(cur.x);
        w[1] = fmd_bias_bytes(cur.y);
        w[2] = fmd_bias_bytes(cur.z);
        w[3] = fmd_bias_bytes(cur.w);
    }
    w[4] = w[5] = 0u;
    if (CARRY) {
        w0c = fmd_bias_bytes(nxt.x); /* this lane's first word of the next pass */
        w[4] = __shfl_sync(0xffffffffu, lane == 0 ? w0c : w[0], nxt_lane);
    } else {
        if (HW >= 1) w[4] = __shfl_sync(0xffffffffu, lane == 0 ? nxt.x : cur.x, nxt_lane);
        if (HW >= 2) w[5] = __shfl_sync(0xffffffffu, lane == 0 ? nxt.y : cur.y, nxt_lane);
        if (COND != 3) { /* normalizeInput works on the raw bytes */
            if (HW >= 1) w[4] = fmd_bias_bytes(w[4]);
            if (HW >= 2) w[5] = fmd_bias_bytes(w[5]);
        }
    }

    /* vi[k] = (I_k, I_{k+4}), vq[k] = (Q_k, Q_{k+4}): complex samples k and k+4 of this lane feed
     * outputs j and j+2, so every later step runs two outputs per instruction (FADD2/FFMA2) */
    constexpr int NK = 4 + HC;
    fmd_f2 vi[NK], vq[NK];
#pragma unroll
    for (int k = 0; k < NK; ++k) {
        if (COND == 3) {
            vi[k] = (k & 1) ? fmd_bytes_to_unit_f2<2>(w[k >> 1], w[(k >> 1) + 2])
                            : fmd_bytes_to_unit_f2<0>(w[k >> 1], w[(k >> 1) + 2]);
            vq[k] = (k & 1) ? fmd_bytes_to_unit_f2<3>(w[k >> 1], w[(k >> 1) + 2])
                            : fmd_bytes_to_unit_f2<1>(w[k >> 1], w[(k >> 1) + 2]);
        } else if (k & 1) {
            vi[k] = fmd_bytes_to_f2<2>(w[k >> 1], w[(k >> 1) + 2]);
            vq[k] = fmd_bytes_to_f2<3>(w[k >> 1], w[(k >> 1) + 2]);
        } else {
            vi[k] = fmd_bytes_to_f2<0>(w[k >> 1], w[(k >> 1) + 2]);
            vq[k] = fmd_bytes_to_f2<1>(w[k >> 1], w[(k >> 1) + 2]);
        }
    }
    if (COND == 3 && HC > 0) {
        /* The zero extension past the end of a chunk is the FLOAT 0.0 (the reference's zeroed
         * arena); with shiftOrigin the fill byte 0x7f shifts to exactly that, but no byte
         * normalises to 0.0: where the halo is the zero extension (lane 31 of the chunk's last
         * pass) its samples -- the high halves of entries 4.. -- are cleared after the conversion. */
        const bool clear = hz && lane == 31u;
#pragma unroll
        for (int k = 4; k < NK; ++k) {
            float lo, hi;
            fmd_upk(vi[k], lo, hi);
            vi[k] = fmd_pk(lo, clear ? 0.0f : hi);
            fmd_upk(vq[k], lo, hi);
            vq[k] = fmd_pk(lo, clear ? 0.0f : hi);
        }
    }

    fmd_pass_core<NTI, NTO, AT, LITE, NK>(vi, vq, xprev, xcur, zero_x, dst, do_store, lane, nxt_lane, st);
}

/* ---- cp.async (LDGSTS) ring: global -> shared without staging registers ---- */
__device__ __forceinline__ void fmd_cp_async16(uint32_t smem_dst, const void *gsrc) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_dst), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void fmd_cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void fmd_cp_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ uint4 fmd_lds128(uint32_t a) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a));
    return v;
}
__device__ __forceinline__ uint32_t fmd_lds32(uint32_t a) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ uint2 fmd_lds64(uint32_t a) {
    uint2 v;
    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(a));
    return v;
}

/* ---- 1-D TMA ring: one elected lane per warp moves the warp's 512 bytes of a pass with a single
 * bulk copy (cp.async.bulk, SASS UBLKCP) that signals an mbarrier; the 32 lanes only test the
 * barrier's phase.  Address, predicate, commit and wait bookkeeping of the per-lane cp.async ring
 * leave 31 of the 32 lanes. ---- */
__device__ __forceinline__ void fmd_mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void fmd_mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
/* arm the barrier with the byte count and start the copy (called by ONE lane) */
__device__ __forceinline__ void fmd_bulk_fetch(uint32_t smem_dst, const void *gsrc, uint32_t bytes, uint32_t bar) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_dst),
                 "l"(gsrc), "r"(bytes), "r"(bar)
                 : "memory");
}
/* wait until the barrier's phase with parity `parity` has completed; bounded: a copy that never
 * lands traps (an error the host sees) instead of hanging the GPU */
__device__ __forceinline__ void fmd_mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done = 0u;
#pragma unroll 1
    for (unsigned spin = 0; spin < (1u << 22); ++spin) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done)
                     : "r"(bar), "r"(parity)
                     : "memory");
        if (done) return;
    }
    __trap();
}

/* one lane of the (converged) warp, chosen by the hardware: lets ptxas keep the copy's operands uniform */
__device__ __forceinline__ bool fmd_elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0u;
}

#ifndef FMD_EARLY_LDS
#define FMD_EARLY_LDS 0 /* 1: (halo kernels) read a pass's raw words from the ring one pass early, so that its
                           unpacking may be scheduled into the previous pass's polynomial */
#endif
#ifndef FMD_TMA
#define FMD_TMA 0 /* 1: bulk-copy (TMA) ring with mbarriers; 0: per-lane cp.async ring */
#endif
#ifndef FMD_MINB
#define FMD_MINB(NTI) 4 /* resident CTAs per SM the register budget is cut for */
#endif
#ifndef FMD_OVERREAD
#define FMD_OVERREAD 0 /* 1 = EXPERIMENT ONLY: unpredicated prefetch (reads up to 3.5 KB past a segment) */
#endif
#ifndef FMD_STAGES
#define FMD_STAGES 8 /* ring depth per lane (4 or 8): 16 B x FMD_STAGES in flight or landed */
#endif

/*
 * A warp owns one segment: n_it passes of 128 outputs (512 B in, 512 B out per pass),
 * plus one "drain" pass that only feeds the last real pass its neighbours.  If the
 * segment ends at the end of a chunk the drain pass is the zero extension (no load,
 * x = 0); otherwise it re-reads the first 512 B of what follows.
 *
 * Input reaches the lanes through a per-lane ring in shared memory filled by cp.async
 * (LDGSTS, L2-only): every lane copies its own 16 bytes FMD_STAGES-1 passes ahead and
 * reads them back with one LDS.128, so the loads in flight per SM do not depend on the
 * register budget or on occupancy, and no lane ever reads another lane's slot (no barrier).
 */
template <int NTI, int NTO, int AT, int COND = 0>
__global__ void __launch_bounds__(FMD_BLOCK, FMD_MINB(NTI))
fmd_fast_kernel(const uint8_t *__restrict__ in, float *__restrict__ out, const FmdGeom g, const FmdStencil st,
                const FmdSections *__restrict__ sec) {
    /* The first CTAs do the op-for-op work (chunk heads, ragged remainder): they start with the
     * grid and finish long before the streaming CTAs do, so that work costs no extra launch and
     * hides behind the bandwidth-bound part. */
    if (blockIdx.x < g.exact_blocks) {
        fmd_exact_thread(in, out, g, sec, (unsigned long long) blockIdx.x * FMD_BLOCK + threadIdx.x);
        return;
    }
    constexpr int S = FMD_STAGES;
    constexpr bool HALO = NTI > 1;               /* lane 31 needs lane 0's next pass */
    constexpr int PENDING = HALO ? S - 2 : S - 1; /* groups allowed in flight at the wait */
    constexpr bool TMA = FMD_TMA != 0;
    __shared__ __align__(128) uint4 ring[S * FMD_BLOCK];
    __shared__ __align__(8) unsigned long long bars[TMA ? S * FMD_WARPS : 1]; /* one mbarrier per warp and stage */

    const unsigned lane = threadIdx.x & 31u;
    /* warp-uniform by construction (the shuffle tells ptxas so): segment, source and ring addresses
     * of the warp can then live in uniform registers, which is what the bulk copy takes */
    const unsigned warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
    /* the host keeps nchunks * segs_per_chunk below 2^32 per launch */
    const unsigned seg = (blockIdx.x - g.exact_blocks) * FMD_WARPS + warp;
    if (seg >= g.nchunks * g.segs_per_chunk) return;
    const unsigned chunk = seg / g.segs_per_chunk;
    const unsigned sic = seg - chunk * g.segs_per_chunk;
    const unsigned it0 = sic * g.iters_per_seg;
    const unsigned iters_chunk = g.fast_out >> 7;
    const unsigned n_it = min(g.iters_per_seg, iters_chunk - it0);
    /* does the drain pass fall off the end of the chunk? (fast_out == chunk_out there) */
    const bool end_zero = (it0 + n_it == iters_chunk) && (g.fast_out == g.chunk_out);
    const unsigned n_ld = end_zero ? n_it : n_it + 1u; /* positions that exist in memory */
    const unsigned nxt_lane = (lane + 1u) & 31u;
    const unsigned k0 = it0 * FMD_PASS_OUT + lane * 4u;

    /* the warp's first byte (what the bulk copy reads from), then this lane's 16 bytes in it */
    const uint8_t *__restrict__ src0 = in + (unsigned long long) chunk * g.chunk_stride + 4ull * (it0 * FMD_PASS_OUT);
    const uint8_t *__restrict__ src = src0 + lane * 16u;
    float *__restrict__ dst = out + (unsigned long long) chunk * g.chunk_out + k0;
    const uint4 fill = make_uint4(FMD_FILL, FMD_FILL, FMD_FILL, FMD_FILL);
    /* TMA ring: the warp's 512 bytes of slot q start at lane 0's entry; stage barriers of this warp */
    const uint32_t wbase = (uint32_t) __cvta_generic_to_shared(&ring[0]) + warp * 512u;
    const uint32_t sbase = wbase + lane * 16u;
    constexpr uint32_t SLOT = FMD_BLOCK * 16u;
    static_assert(S == 4 || S == 8, "the main loop is unrolled by the ring depth");
    const uint32_t bbase = (uint32_t) __cvta_generic_to_shared(&bars[0]) + (TMA ? warp * (8u * S) : 0u);
    if (TMA) {
        if (lane == 0u) {
#pragma unroll
            for (int q = 0; q < S; ++q) fmd_mbar_init(bbase + 8u * q, 1u);
            fmd_mbar_fence_init();
        }
        __syncwarp();
    }

    /* position q of the segment lives in ring slot q % S; positions >= n_ld do not exist in memory */
#pragma unroll
    for (int q = 0; q < S - 1; ++q) {
        if (TMA) {
            if ((unsigned) q < n_ld && fmd_elect_one()) fmd_bulk_fetch(wbase + q * SLOT, src0 + 512 * q, 512u, bbase + 8u * q);
        } else {
            if ((unsigned) q < n_ld) fmd_cp_async16(sbase + q * SLOT, src + 512 * q);
            fmd_cp_commit();
        }
    }
    /* stage s holds position base + s during the trip that starts at `base`: its k-th fill, k = base / S,
     * completes the barrier phase of parity k & 1 */
    unsigned ph = 0u;
    if (TMA && HALO) fmd_mbar_wait(bbase, 0u); /* position 0; every pass then waits for its successor's stage */

    float x0[4], x1[4] = {0.f, 0.f, 0.f, 0.f};
    const uint2 fill2 = make_uint2(FMD_FILL, FMD_FILL);
    constexpr bool CARRY = FmdCarry<NTI, COND>::value;
    constexpr int HWORDS = NTI > 1 ? NTI / 2 : 0; /* halo words: (NTI - 1 + 1) / 2 */
    uint32_t w0c = 0u;
    constexpr bool EARLY = FMD_EARLY_LDS && HALO && !TMA;
    uint4 rawc = fill; /* EARLY: the raw words of the pass about to run */
    if (CARRY || EARLY) { /* the first pass takes its (biased) first word from here, every later one from its predecessor */
        if (!TMA) fmd_cp_wait<S - 2>();
        if (EARLY) rawc = fmd_lds128(sbase);
        w0c = fmd_bias_bytes(EARLY ? rawc.x : fmd_lds32(sbase));
    }
    /* the chunk head [0, head) belongs to the exact path: those lanes of the very first store
     * of a chunk's first segment keep quiet */
    const bool head_keep = !(it0 == 0 && k0 < g.head);

    /* pass p = base + U: fetch position p + S - 1 into the slot pass p - 1 just released, wait for
     * position p (and p + 1 when the input filter needs lane 0's next word), discriminate, and
     * finish pass p - 1 (store predicate STP).  Slots, source and destination offsets are
     * immediates because the loop is unrolled by the ring depth. */
#define FMD_STEP(U, XPREV, XCUR, STP)                                                                   \
    {                                                                                                   \
        const bool hz = HALO && (U) == S - 1 && base + S >= n_ld; /* what follows is the zero extension */ \
        if (TMA) {                                                                                      \
            /* the slot pass p - 1 read is free: every lane's LDS of it has been consumed */            \
            __syncwarp();                                                                               \
            if (base + (U) + S - 1 < n_ld && fmd_elect_one())                                           \
                fmd_bulk_fetch(wbase + (((U) + S - 1) & (S - 1)) * SLOT, src0 + 512 * ((U) + S - 1), 512u, \
                               bbase + 8u * (((U) + S - 1) & (S - 1)));                                 \
            if (!HALO) fmd_mbar_wait(bbase + 8u * (U), ph);                                             \
            else if (!hz) fmd_mbar_wait(bbase + 8u * (((U) + 1) & (S - 1)), (U) == S - 1 ? ph ^ 1u : ph); \
        } else {                                                                                        \
            if (FMD_OVERREAD || base + (U) + S - 1 < n_ld)                                              \
                fmd_cp_async16(sbase + (((U) + S - 1) & (S - 1)) * SLOT, src + 512 * ((U) + S - 1));    \
            fmd_cp_commit();                                                                            \
            fmd_cp_wait<PENDING>();                                                                     \
        }                                                                                               \
        uint4 cur;                                                                                      \
        uint2 nx = fill2;                                                                               \
        if (EARLY) {                                                                                    \
            cur = rawc; /* read from the ring during the previous pass */                               \
            rawc = hz ? fill : fmd_lds128(sbase + (((U) + 1) & (S - 1)) * SLOT);                        \
            nx = make_uint2(rawc.x, rawc.y);                                                            \
        } else {                                                                                        \
            cur = fmd_lds128(sbase + (U) * SLOT);                                                       \
            if (HALO && HWORDS >= 2) nx = fmd_lds64(sbase + (((U) + 1) & (S - 1)) * SLOT);              \
            if (HALO && HWORDS < 2) nx.x = fmd_lds32(sbase + (((U) + 1) & (S - 1)) * SLOT);             \
            if (hz) nx = fill2;                                                                         \
        }                                                                                               \
        fmd_pass<NTI, NTO, AT, false, COND>(cur, nx, XPREV, XCUR, false, dst + 128 * ((U) - 1), STP, lane, nxt_lane, st, w0c, hz); \
    }

    unsigned base = 0;
    for (; base + S <= n_it; base += S) {
        const bool later = base != 0;
        FMD_STEP(0, x1, x0, later)
        FMD_STEP(1, x0, x1, later || head_keep)
        FMD_STEP(2, x1, x0, true)
        FMD_STEP(3, x0, x1, true)
        if (S == 8) {
            FMD_STEP(4, x1, x0, true)
            FMD_STEP(5, x0, x1, true)
            FMD_STEP(6, x1, x0, true)
            FMD_STEP(7, x0, x1, true)
        }
        src += 512 * S;
        src0 += 512 * S;
        dst += 128 * S;
        ph ^= 1u;
    }
#undef FMD_STEP
    /* n_it is not a multiple of the ring depth: up to S-1 more real passes, slots by arithmetic.
     * After the loop above the newest x is in x1 and position `base` sits in slot 0. */
    unsigned p = base;
    for (; p < n_it; ++p) {
        const unsigned u = p - base;
        if (TMA) {
            __syncwarp();
            if (p + S - 1 < n_ld && fmd_elect_one())
                fmd_bulk_fetch(wbase + ((u + S - 1) & (S - 1)) * SLOT, src0 + 512 * (u + S - 1), 512u,
                               bbase + 8u * ((u + S - 1) & (S - 1)));
            /* fewer than S passes are left: stages u and u + 1 still belong to this trip's phase */
            if (!HALO) fmd_mbar_wait(bbase + 8u * u, ph);
            else if (p + 1 < n_ld) fmd_mbar_wait(bbase + 8u * (u + 1u), ph);
        } else {
            if (p + S - 1 < n_ld) fmd_cp_async16(sbase + ((u + S - 1) & (S - 1)) * SLOT, src + 512 * (u + S - 1));
            fmd_cp_commit();
            fmd_cp_wait<PENDING>();
        }
        uint4 cur;
        uint2 nx = fill2;
        if (EARLY) {
            cur = rawc;
            rawc = (p + 1 < n_ld) ? fmd_lds128(sbase + ((u + 1) & (S - 1)) * SLOT) : fill;
            nx = make_uint2(rawc.x, rawc.y);
        } else {
            cur = fmd_lds128(sbase + (u & (S - 1)) * SLOT);
            if (HALO && p + 1 < n_ld) nx = fmd_lds64(sbase + ((u + 1) & (S - 1)) * SLOT);
        }
        fmd_pass<NTI, NTO, AT, false, COND>(cur, nx, x1, x0, false, dst + 128 * ((int) u - 1), p >= 2 || (p == 1 && head_keep),
                                     lane, nxt_lane, st, w0c, HALO && !(p + 1 < n_ld));
        x1[0] = x0[0]; x1[1] = x0[1]; x1[2] = x0[2]; x1[3] = x0[3];
    }
    /* the drain pass (position n_it): only lane 0's first NTO-1 outputs of it are consumed (by lane
     * 31, through the shuffle), so with a 2-tap output stencil half a pass is enough */
    {
        const unsigned u = p - base;
        if (TMA) {
            if (!HALO && p < n_ld) fmd_mbar_wait(bbase + 8u * u, ph); /* (with a halo the last pass waited for it) */
        } else {
            fmd_cp_commit();
            fmd_cp_wait<PENDING>();
        }
        uint4 cur = fill;
        if (EARLY) cur = rawc;
        else if (p < n_ld) cur = fmd_lds128(sbase + (u & (S - 1)) * SLOT);
        fmd_pass<NTI, NTO, AT, (NTO <= 2), COND>(cur, fill2, x1, x0, end_zero, dst + 128 * ((int) u - 1),
                                                 p >= 2 || (p == 1 && head_keep), lane, nxt_lane, st, w0c);
    }

    if (!TMA) fmd_cp_wait<0>();
}

/* ------------------------------------------------- float-input streaming kernel (-q1) */

/*
 * The same streaming pass over samples that arrive as FLOATS, one per input byte in the same
 * interleaved layout: -q1 (correctIq, matrix.c:120-140) conditions the stream with offsets that
 * run through the whole stream (fmd_correctiq.cuh), so its samples are not a function of their
 * own bytes.  A lane reads its 8 complex samples (64 bytes, four 16-byte loads) one pass ahead
 * into registers; the input-filter halo is the next lane's first samples by shuffle (lane 31:
 * lane 0's already loaded next pass); the zero extension past a chunk is the float 0.
 * Chunk heads and ragged remainders go to the exact CTAs exactly as in fmd_fast_kernel (their
 * base lane reads the same floats).  HBM-bound: 8 B read + 2 B written per sample.
 */
#define FMD_FBLOCK_MINB 3

template <int NTI, int NTO, int AT>
__global__ void __launch_bounds__(FMD_BLOCK, FMD_FBLOCK_MINB)
fmd_fastf_kernel(const float *__restrict__ in, float *__restrict__ out, const FmdGeom g, const FmdStencil st,
                 const FmdSections *__restrict__ sec) {
    if (blockIdx.x < g.exact_blocks) {
        fmd_exact_thread(reinterpret_cast<const uint8_t *>(in), out, g, sec,
                         (unsigned long long) blockIdx.x * FMD_BLOCK + threadIdx.x);
        return;
    }
    constexpr int HC = NTI > 1 ? NTI - 1 : 0; /* complex samples needed from the next lane */
    constexpr int NK = 4 + HC;
    const unsigned lane = threadIdx.x & 31u;
    const unsigned seg = (blockIdx.x - g.exact_blocks) * FMD_WARPS + (threadIdx.x >> 5);
    if (seg >= g.nchunks * g.segs_per_chunk) return;
    const unsigned chunk = seg / g.segs_per_chunk;
    const unsigned sic = seg - chunk * g.segs_per_chunk;
    const unsigned it0 = sic * g.iters_per_seg;
    const unsigned iters_chunk = g.fast_out >> 7;
    const unsigned n_it = min(g.iters_per_seg, iters_chunk - it0);
    const bool end_zero = (it0 + n_it == iters_chunk) && (g.fast_out == g.chunk_out);
    const unsigned n_ld = end_zero ? n_it : n_it + 1u;
    const unsigned nxt_lane = (lane + 1u) & 31u;
    const unsigned k0 = it0 * FMD_PASS_OUT + lane * 4u;
    /* chunk_stride is in bytes; an output comes from 2 complex samples = 4 floats */
    const float4 *__restrict__ src = reinterpret_cast<const float4 *>(
        reinterpret_cast<const uint8_t *>(in) + (unsigned long long) chunk * g.chunk_stride + 16ull * k0);
    float *__restrict__ dst = out + (unsigned long long) chunk * g.chunk_out + k0;
    const bool head_keep = !(it0 == 0 && k0 < g.head);
    const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);

    /* position p: 128 float4 per warp (lane l: entries 4l .. 4l+3) */
    float4 cur[4], nxt[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) cur[i] = __ldg(src + i);

    float x0[4], x1[4] = {0.f, 0.f, 0.f, 0.f};
    for (unsigned p = 0; p <= n_it; ++p) {
        const bool drain = p == n_it;
        /* the following position, or the zero extension past the chunk */
#pragma unroll
        for (int i = 0; i < 4; ++i) nxt[i] = (p + 1u < n_ld) ? __ldg(src + 128u * (p + 1u) + i) : zero4;
        if (drain && !(p < n_ld)) {
#pragma unroll
            for (int i = 0; i < 4; ++i) cur[i] = zero4;
        }
        /* complex sample k of this lane: (c[2k], c[2k+1]); samples 8.. come from the next lane */
        const float *c = reinterpret_cast<const float *>(cur);
        const float *n = reinterpret_cast<const float *>(nxt);
        float hI[HC > 0 ? HC : 1], hQ[HC > 0 ? HC : 1];
#pragma unroll
        for (int h = 0; h < HC; ++h) {
            hI[h] = __shfl_sync(0xffffffffu, lane == 0 ? n[2 * h] : c[2 * h], nxt_lane);
            hQ[h] = __shfl_sync(0xffffffffu, lane == 0 ? n[2 * h + 1] : c[2 * h + 1], nxt_lane);
        }
        fmd_f2 vi[NK], vq[NK];
#pragma unroll
        for (int k = 0; k < NK; ++k) {
            vi[k] = fmd_pk(c[2 * k], k + 4 < 8 ? c[2 * (k + 4)] : hI[k + 4 - 8 < HC ? k + 4 - 8 : 0]);
            vq[k] = fmd_pk(c[2 * k + 1], k + 4 < 8 ? c[2 * (k + 4) + 1] : hQ[k + 4 - 8 < HC ? k + 4 - 8 : 0]);
        }
        const bool do_store = p >= 2 || (p == 1 && head_keep);
        if (!drain) {
            fmd_pass_core<NTI, NTO, AT, false, NK>(vi, vq, x1, x0, false, dst + 128 * ((int) p - 1), do_store, lane, nxt_lane, st);
        } else { /* the drain pass: only lane 0's first outputs are consumed (by lane 31) */
            fmd_pass_core<NTI, NTO, AT, (NTO <= 2), NK>(vi, vq, x1, x0, end_zero, dst + 128 * ((int) p - 1), do_store, lane,
                                                        nxt_lane, st);
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            cur[i] = nxt[i];
            x1[i] = x0[i];
        }
    }
}

#endif /* FMD_KERNELS_CUH */
