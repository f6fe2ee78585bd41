/*
 * fmd_quality.cuh -- sm_100a kernels of the opt-in quality pipeline (FMD_FLAG_QUALITY).
 *
 *   s[n] = (I - 127.5, Q - 127.5)
 *   z[n] = input lowpass cascade (recursive, per lane; only with -L)
 *   x[n] = atan2(Im, Re)(z[n] * conj(z[n-1]))
 *   v[n] = output lowpass cascade (recursive)
 *   y[m] = v[m * M]                     only these floats are written to HBM
 *
 * The recursions run as a three-level scan with an explicit state carry (fmd_quality.h):
 *
 *   thread  T = 64 (or 128, large decimation factors) consecutive samples, serially, from a ZERO state
 *   CTA     affine scan over its 256 threads (warp shuffles with the matrices A^(T*2^j), then a
 *           short serial pass over the 8 warp totals): state at the start of every thread's span
 *   launch  fmq_scan_tiles_kernel scans the per-tile end states of all CTAs (and takes the carry
 *           of the previous call in, leaves the carry for the next call behind)
 *
 * Input lowpass (a scan whose result feeds a non-linearity, so it has to be resolved first):
 *   fmq_main_kernel          thread- and CTA-level scan, then a decoupled look-back over the end
 *                            states the earlier tiles of the launch have published (FmqLookback):
 *                            the input is read once; re-runs the cascade from the right state and
 *                            goes on to x and v
 * Output lowpass (linear in x, so the carry can be added afterwards, to the kept samples only):
 *   fmq_main_kernel          v from a zero state at the tile start (thread-level carry applied
 *                            in shared memory), decimated; per-tile end state
 *   fmq_scan_tiles_kernel    state at the start of every tile
 *   fmq_zir_kernel           y[m] += C A^k q_tile        (the tile's zero-input response)
 *
 * The input cascade's spans are shifted by one sample (thread t owns samples F-1 .. F+T-2 of it,
 * F = first sample of its x span): re-running it from the state at F-1 yields z[F-1], the
 * neighbour the discriminator needs for x[F], with no exchange between threads.
 *
 * The only inter-CTA dependency is the look-back of the input cascade (tiles in ticket order,
 * bounded waits); every other carry travels between launches.
 */
#ifndef FMD_QUALITY_CUH
#define FMD_QUALITY_CUH

#include <cuda_runtime.h>
#include <stdint.h>

#include "fmd_math.cuh"
#include "fmd_quality.h"

/* the carry between calls (one per context, device memory) */
struct FmqState {
    float in_q[2 * FMQ_MAXD]; /* input cascade, [lane][D] packed: state BEFORE the sample s_prev */
    float out_q[FMQ_MAXD];    /* output cascade: state before the next x                        */
    float s_prev[2];          /* the centred sample that precedes the next call (0,0 at start)  */
    unsigned error;           /* raised by a look-back that gave up waiting (never seen so far)  */
};

struct FmqGeom {
    unsigned n;      /* samples in this launch (< 2^31)                                        */
    unsigned ntiles;
    unsigned r0;     /* launch-relative index of the first kept sample (may be >= n)           */
    unsigned decim;  /* M                                                                      */
    unsigned n_out;  /* outputs of this launch                                                 */
};

struct FmqMat {
    float m[FMQ_MAXD * FMQ_MAXD];
};

/* byte B of a raw word -> exact float u - 127.5:  0x4700uu00 = 32768 + u */
template <int B>
__device__ __forceinline__ float fmq_centre(uint32_t w) {
    return __uint_as_float(__byte_perm(w, 0x47000000u, 0x7504 | (B << 4))) - 32895.5f;
}

__device__ __forceinline__ void fmq_load_sample(const uint8_t *__restrict__ in, unsigned idx, float &si, float &sq) {
    const unsigned short h = *reinterpret_cast<const unsigned short *>(in + 2ull * idx);
    si = (float) (h & 0xffu) - 127.5f;
    sq = (float) (h >> 8) - 127.5f;
}

/* one sample through L sections, direct form II transposed; q = (s1_0, s2_0, s1_1, ...).
 * FO: the last section is first order (odd degree: b2 = a2 = 0, src/filter.c zp2Sos), its second
 * state stays 0 and is not touched -- two instructions less per sample and cascade */
template <int L, bool FO>
__device__ __forceinline__ float fmq_step(const fmq_consts &c, float (&q)[2 * L], float u) {
#pragma unroll
    for (int i = 0; i < L; ++i) {
        const float w = fmaf(c.sos[i][0], u, q[2 * i]);
        if (FO && i == L - 1) {
            q[2 * i] = fmaf(-c.sos[i][3], w, c.sos[i][1] * u);
        } else {
            q[2 * i] = fmaf(-c.sos[i][3], w, fmaf(c.sos[i][1], u, q[2 * i + 1]));
            q[2 * i + 1] = fmaf(-c.sos[i][4], w, c.sos[i][2] * u);
        }
        u = w;
    }
    return u;
}

/* number of kept samples (launch-relative index >= r0, (index - r0) % M == 0) below `pos` */
__device__ __forceinline__ unsigned fmq_kept_below(unsigned pos, unsigned r0, unsigned M) {
    return pos <= r0 ? 0u : (pos - r0 + M - 1u) / M;
}

/* inclusive affine scan over the lanes of a warp: v_l <- sum_{i<=l} P^(l-i) v_i, with the
 * matrices c.lvl[j] = P^(2^j) */
template <int D, int NV>
__device__ __forceinline__ void fmq_warp_scan(float (&v)[NV][D], const fmq_consts &c, const unsigned lane) {
#pragma unroll
    for (int j = 0; j < 5; ++j) {
        float up[NV][D];
#pragma unroll
        for (int a = 0; a < NV; ++a)
#pragma unroll
            for (int i = 0; i < D; ++i) up[a][i] = __shfl_up_sync(0xffffffffu, v[a][i], 1u << j);
        if (lane >= (1u << j)) {
#pragma unroll
            for (int a = 0; a < NV; ++a)
#pragma unroll
                for (int i = 0; i < D; ++i) {
                    float acc = v[a][i];
#pragma unroll
                    for (int k = 0; k < D; ++k) acc = fmaf(c.lvl[j][i * D + k], up[a][k], acc);
                    v[a][i] = acc;
                }
        }
    }
}

/*
 * Exclusive affine scan over the threads of a CTA: in  v[nv] = end state of the thread's span run
 * from a zero state; out v[nv] = state at the START of the thread's span, given `carry` = state at
 * the start of the tile.  Every thread of the CTA must call it (one __syncthreads inside).
 */
template <int D, int NV>
__device__ __forceinline__ void fmq_block_scan(float (&v)[NV][D], const fmq_consts &c,
                                               const float *__restrict__ lane_tab, const float (&carry)[NV][D],
                                               float *tot /* shared, FMQ_WARPS * NV * D */) {
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    fmq_warp_scan<D, NV>(v, c, lane);
    if (lane == 31u) {
#pragma unroll
        for (int a = 0; a < NV; ++a)
#pragma unroll
            for (int i = 0; i < D; ++i) tot[(warp * NV + a) * D + i] = v[a][i];
    }
    __syncthreads();
    float wp[NV][D];
#pragma unroll
    for (int a = 0; a < NV; ++a)
#pragma unroll
        for (int i = 0; i < D; ++i) wp[a][i] = carry[a][i];
    for (unsigned w = 0; w < warp; ++w) {
        float nw[NV][D];
#pragma unroll
        for (int a = 0; a < NV; ++a)
#pragma unroll
            for (int i = 0; i < D; ++i) {
                float acc = tot[(w * NV + a) * D + i];
#pragma unroll
                for (int k = 0; k < D; ++k) acc = fmaf(c.warp[i * D + k], wp[a][k], acc);
                nw[a][i] = acc;
            }
#pragma unroll
        for (int a = 0; a < NV; ++a)
#pragma unroll
            for (int i = 0; i < D; ++i) wp[a][i] = nw[a][i];
    }
    float ex[NV][D];
#pragma unroll
    for (int a = 0; a < NV; ++a)
#pragma unroll
        for (int i = 0; i < D; ++i) {
            const float t = __shfl_up_sync(0xffffffffu, v[a][i], 1u);
            ex[a][i] = lane ? t : 0.0f;
        }
    const float *__restrict__ lt = lane_tab + lane * (D * D);
#pragma unroll
    for (int i = 0; i < D; ++i)
#pragma unroll
        for (int k = 0; k < D; ++k) {
            const float m = __ldg(lt + i * D + k);
#pragma unroll
            for (int a = 0; a < NV; ++a) ex[a][i] = fmaf(m, wp[a][k], ex[a][i]);
        }
#pragma unroll
    for (int a = 0; a < NV; ++a)
#pragma unroll
        for (int i = 0; i < D; ++i) v[a][i] = ex[a][i];
}

/* end state of a span of `cnt` samples: A^cnt q0 + e  (apow = table of A^k, k = 0..T) */
template <int D>
__device__ __forceinline__ void fmq_span_end(const float *__restrict__ apow, unsigned cnt, const float (&q0)[D],
                                             const float (&e)[D], float *__restrict__ dst) {
    const float *__restrict__ m = apow + cnt * (D * D);
#pragma unroll
    for (int i = 0; i < D; ++i) {
        float acc = e[i];
#pragma unroll
        for (int k = 0; k < D; ++k) acc = fmaf(__ldg(m + i * D + k), q0[k], acc);
        dst[i] = acc;
    }
}

/* Full spans are walked in T / 8 groups of 8 samples (one 16-byte load each); only the body of
 * a group is unrolled, so the instruction footprint stays a quarter of the fully unrolled span
 * (the fully unrolled <2,2> kernel stalled mostly on instruction fetch: ncu `no_instruction`). */

__device__ __forceinline__ uint4 fmq_load_group(const uint8_t *__restrict__ in, unsigned F, int grp) {
    return __ldg(reinterpret_cast<const uint4 *>(in + 2ull * F) + grp);
}

/* sample i (0..7) of a group: compile-time i */
#define FMQ_GW(v, i) ((i) < 2 ? (v).x : (i) < 4 ? (v).y : (i) < 6 ? (v).z : (v).w)
#define FMQ_SI(v, i) (((i) & 1) ? fmq_centre<2>(FMQ_GW(v, i)) : fmq_centre<0>(FMQ_GW(v, i)))
#define FMQ_SQ(v, i) (((i) & 1) ? fmq_centre<3>(FMQ_GW(v, i)) : fmq_centre<1>(FMQ_GW(v, i)))

/*
 * Input cascade from a zero state over the thread's SHIFTED span: the halo sample (F-1), then
 * samples F .. F+cnt-2.  cnt = number of samples of the x span (0 for a thread past the end).
 */
template <int LIN, bool FOI, int T>
__device__ __forceinline__ void fmq_in_zero_state(const fmq_consts &cin, const uint8_t *__restrict__ in, unsigned F,
                                                  unsigned cnt, float hi, float hq, float (&e)[2][2 * LIN]) {
#pragma unroll
    for (int i = 0; i < 2 * LIN; ++i) e[0][i] = e[1][i] = 0.0f;
    if (cnt == T) {
        fmq_step<LIN, FOI>(cin, e[0], hi);
        fmq_step<LIN, FOI>(cin, e[1], hq);
        uint4 nx = fmq_load_group(in, F, 0);
#pragma unroll 1
        for (int grp = 0; grp < (T / 8); ++grp) {
            const uint4 v = nx;
            if (grp + 1 < (T / 8)) nx = fmq_load_group(in, F, grp + 1);
#pragma unroll
            for (int i = 0; i < 7; ++i) {
                fmq_step<LIN, FOI>(cin, e[0], FMQ_SI(v, i));
                fmq_step<LIN, FOI>(cin, e[1], FMQ_SQ(v, i));
            }
            if (grp + 1 < (T / 8)) { /* the span's very last sample belongs to the next thread's span */
                fmq_step<LIN, FOI>(cin, e[0], FMQ_SI(v, 7));
                fmq_step<LIN, FOI>(cin, e[1], FMQ_SQ(v, 7));
            }
        }
    } else if (cnt) {
        fmq_step<LIN, FOI>(cin, e[0], hi);
        fmq_step<LIN, FOI>(cin, e[1], hq);
#pragma unroll 1
        for (unsigned i = 0; i + 1 < cnt; ++i) {
            float si, sq;
            fmq_load_sample(in, F + i, si, sq);
            fmq_step<LIN, FOI>(cin, e[0], si);
            fmq_step<LIN, FOI>(cin, e[1], sq);
        }
    }
}

/* per-thread geometry shared by the kernels */
struct FmqSpan {
    unsigned S, F, cnt; /* tile start, thread start (launch-relative samples), samples in the span */
    bool last;          /* this thread's span reaches the end of the tile's valid samples          */
};

template <int T>
__device__ __forceinline__ FmqSpan fmq_span(const FmqGeom &g, const unsigned tile) {
    FmqSpan s;
    s.S = tile * (unsigned) (T * FMQ_THREADS);
    s.F = s.S + threadIdx.x * (unsigned) T;
    s.cnt = s.F >= g.n ? 0u : min((unsigned) T, g.n - s.F);
    s.last = s.cnt != 0u && (threadIdx.x == FMQ_THREADS - 1 || s.F + T >= g.n);
    return s;
}

__device__ __forceinline__ void fmq_halo(const uint8_t *__restrict__ in, const FmqState *__restrict__ st, unsigned F,
                                         unsigned cnt, float &hi, float &hq) {
    hi = hq = 0.0f;
    if (cnt == 0u) return;
    if (F == 0u) {
        hi = st->s_prev[0];
        hq = st->s_prev[1];
    } else {
        fmq_load_sample(in, F - 1u, hi, hq);
    }
}

/* ------------------------------------------ input cascade: carry between the tiles of a launch */

/*
 * The input cascade feeds a non-linearity, so its state at the start of a tile has to be known
 * INSIDE the main kernel.  Decoupled look-back (Merrill & Garland) over affine maps with a constant
 * matrix: every tile publishes its end state from a zero state (E, flag 1) as soon as it has it,
 * then walks back over its predecessors until it meets one that has already published its true
 * end state (S, flag 2), replays  q <- P q + E_m  forward from there, and publishes its own S.
 * Tile numbers are handed out by an atomic ticket, so every predecessor of a running tile is
 * itself running or done: no deadlock.  (Measured and dropped: persistent CTAs that run phase A
 * of their NEXT tile before the look-back of the current one, so that no look-back ever waits for
 * an E -- 1.25 ms vs 1.23 ms for the -L configuration, and 4 % slower without -L.)  The wait is bounded anyway (FMQ_SPIN_LIMIT polls): on
 * expiry the kernel raises FmqState::error and carries on with a zero carry instead of hanging.
 */
struct FmqLookback {
    unsigned *ticket;  /* [1]  zeroed before every launch together with `flag`            */
    unsigned *flag;    /* [ntiles] 0 = nothing, 1 = E published, 2 = S published            */
    float *E;          /* [ntiles][2][D]                                                    */
    float *S;          /* [ntiles][2][D]                                                    */
};

#define FMQ_SPIN_LIMIT (1u << 22)
#ifndef FMQ_SLEEP_NS
#define FMQ_SLEEP_NS 0 /* measured on B200: any __nanosleep between polls (20..500 ns asked for) costs the
                          -L configuration 10 % -- the sleep quantum is far longer than an L2 round trip */
#endif

__device__ __forceinline__ unsigned fmq_poll_flag(const unsigned *flag, unsigned *error) {
    unsigned f = 0;
    for (unsigned spin = 0; spin < FMQ_SPIN_LIMIT; ++spin) {
        f = *reinterpret_cast<const volatile unsigned *>(flag);
        if (f) return f;
        if (FMQ_SLEEP_NS) __nanosleep(FMQ_SLEEP_NS);
    }
    atomicExch(error, 1u);
    return 0u;
}

/*
 * State of the input cascade at the start of `tile` (before sample S-1), lanes I and Q.  Called by
 * ONE WHOLE WARP after E[tile] has been published.  Lane l inspects predecessor tile-1-l (32 per
 * round, one L2 round trip for all of them): records up to the nearest one that already holds its
 * true end state S count, each advanced by (A^TILE)^l (`lane_tile`) and, from the second round
 * on, by A^(32 TILE) per round (`wtile`); a butterfly sum gives every lane the result.
 */
template <int D>
__device__ __forceinline__ void fmq_lookback(const FmqLookback &lb, const unsigned tile, const unsigned lane,
                                             const float *__restrict__ lane_tile, const FmqMat &wtile,
                                             float (&q)[2][D], unsigned *error) {
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int i = 0; i < D; ++i) q[a][i] = 0.0f;
    float lp[D * D];
#pragma unroll
    for (int i = 0; i < D * D; ++i) lp[i] = __ldg(lane_tile + lane * (D * D) + i);
    long long base = (long long) tile - 1;
    for (unsigned round = 0;; ++round) {
        const long long idx = base - (long long) lane;
        unsigned f = 0u;
        if (idx >= 0) f = fmq_poll_flag(lb.flag + idx, error);
        __threadfence();
        const unsigned have_s = __ballot_sync(0xffffffffu, idx >= 0 && f == 2u);
        const unsigned gave_up = __ballot_sync(0xffffffffu, idx >= 0 && f == 0u);
        const unsigned first_s = have_s ? (unsigned) __ffs((int) have_s) - 1u : 32u;
        float v[2][D];
#pragma unroll
        for (int a = 0; a < 2; ++a)
#pragma unroll
            for (int i = 0; i < D; ++i) {
                float x = 0.0f;
                if (idx >= 0 && lane <= first_s && f != 0u) {
                    const float *src = (lane == first_s ? lb.S : lb.E) + ((size_t) idx * 2 + a) * D + i;
                    x = __ldcg(src);
                }
                v[a][i] = x;
            }
        /* (A^TILE)^lane, butterfly sum over the lanes, then A^(32 TILE) once per earlier round */
        float t[2][D];
#pragma unroll
        for (int a = 0; a < 2; ++a)
#pragma unroll
            for (int i = 0; i < D; ++i) {
                float acc = 0.0f;
#pragma unroll
                for (int k = 0; k < D; ++k) acc = fmaf(lp[i * D + k], v[a][k], acc);
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
                t[a][i] = acc;
            }
        for (unsigned r = 0; r < round; ++r) {
            float n2[2][D];
#pragma unroll
            for (int a = 0; a < 2; ++a)
#pragma unroll
                for (int i = 0; i < D; ++i) {
                    float acc = 0.0f;
#pragma unroll
                    for (int k = 0; k < D; ++k) acc = fmaf(wtile.m[i * D + k], t[a][k], acc);
                    n2[a][i] = acc;
                }
#pragma unroll
            for (int a = 0; a < 2; ++a)
#pragma unroll
                for (int i = 0; i < D; ++i) t[a][i] = n2[a][i];
        }
#pragma unroll
        for (int a = 0; a < 2; ++a)
#pragma unroll
            for (int i = 0; i < D; ++i) q[a][i] += t[a][i];
        if (have_s || gave_up || base < 32) break;
        base -= 32;
    }
}

/* ------------------------------------------------------------------------------- the main kernel */

/* discriminator for two consecutive samples: x0 from (z0, zp), x1 from (z1, z0) */
__device__ __forceinline__ void fmq_discriminate2(float pr, float pi, float z0r, float z0i, float z1r, float z1i,
                                                  float &x0, float &x1) {
    const float re0 = fmaf(z0r, pr, z0i * pi), im0 = fmaf(z0i, pr, -(z0r * pi));
    const float re1 = fmaf(z1r, z0r, z1i * z0i), im1 = fmaf(z1i, z0r, -(z1r * z0i));
    fmd_atan2_x2<8>(fmd_pk(im0, im1), fmd_pk(re0, re1), x0, x1);
}

/* Kept samples of thread t start at slot t*T/M and are stored at (slot + t): one word of padding
 * per owning thread keeps the threads of a warp on different shared-memory banks (M = 2 had 80 M
 * bank conflicts per launch without padding) and leaves the writer a plain post-increment. */

#ifndef FMQ_MINB
#define FMQ_MINB (1024 / FMQ_THREADS) /* resident CTAs per SM the register budget of the main kernel is cut for: 32 warps */
#endif

template <int NIN, int NOUT, int T> /* degrees of the input (0 = none) and output cascades; samples per thread */
__global__ void __launch_bounds__(FMQ_THREADS, FMQ_MINB)
fmq_main_kernel(const uint8_t *__restrict__ in, float *__restrict__ out, const FmqGeom g, const fmq_consts cin,
                const fmq_consts cout, const float *__restrict__ lane_in, const float *__restrict__ lane_out,
                const float *__restrict__ apow_out, const float *__restrict__ ck_out,
                const float *__restrict__ apow_in, const float *__restrict__ lane_tile_in, const FmqLookback lb,
                const FmqMat plast_in, const FmqMat wtile_in,
                float *__restrict__ eout_tile /* [ntiles][DOUT] */, FmqState *__restrict__ st) {
    constexpr int LIN = (NIN + 1) / 2, LOUT = (NOUT + 1) / 2; /* sections */
    constexpr int LI1 = LIN > 0 ? LIN : 1;
    constexpr bool FOI = (NIN & 1) != 0, FOO = (NOUT & 1) != 0;
    constexpr int DI = 2 * LI1;
    constexpr int DO = 2 * LOUT;
    constexpr int TILE = T * FMQ_THREADS, GROUPS = T / 8;
    extern __shared__ float ybuf[];                    /* kept v of this tile, zero state at the tile start */
    __shared__ float qs[FMQ_THREADS * DO];             /* output-cascade state at the start of every thread */
    __shared__ float ck[T * DO];
    __shared__ float tot_in[FMQ_WARPS * 2 * DI];
    __shared__ float tot_out[FMQ_WARPS * DO];

    __shared__ unsigned tile_s;
    __shared__ float carry_s[2 * DI];
    /* tiles in ticket order: a tile's predecessors have all started (the look-back relies on it) */
    if (threadIdx.x == 0) tile_s = LIN > 0 ? atomicAdd(lb.ticket, 1u) : blockIdx.x;
    for (unsigned i = threadIdx.x; i < T * DO; i += FMQ_THREADS) ck[i] = __ldg(ck_out + i);
    __syncthreads();
    const unsigned tile = tile_s;
    const FmqSpan sp = fmq_span<T>(g, tile);

    const bool full = sp.cnt == T;
    float hi, hq;
    fmq_halo(in, st, sp.F, sp.cnt, hi, hq);

    /* input cascade: state at F-1 */
    float qi[2][DI];
    if (LIN > 0) {
        float e[2][DI], zero[2][DI];
        fmq_in_zero_state<LI1, FOI, T>(cin, in, sp.F, sp.cnt, hi, hq, e);
#pragma unroll
        for (int a = 0; a < 2; ++a)
#pragma unroll
            for (int i = 0; i < DI; ++i) {
                qi[a][i] = e[a][i];
                zero[a][i] = 0.0f;
            }
        /* relative to a zero state at the tile start ... */
        fmq_block_scan<DI, 2>(qi, cin, lane_in, zero, tot_in);
        /* ... which gives the tile's end state from a zero state (E, held by the last thread with
         * samples).  Its warp publishes E, finds the true state at the tile start by look-back,
         * publishes the true end state S. */
        const unsigned tile_samples = min((unsigned) TILE, g.n - sp.S);
        const unsigned last_t = (tile_samples + T - 1u) / T - 1u;
        if ((threadIdx.x >> 5) == (last_t >> 5)) {
            const unsigned lane = threadIdx.x & 31u, src_lane = last_t & 31u;
            float E[2][DI], q0[2][DI];
#pragma unroll
            for (int a = 0; a < 2; ++a)
#pragma unroll
                for (int i = 0; i < DI; ++i) E[a][i] = 0.0f;
            if (sp.last) {
                fmq_span_end<DI>(apow_in, sp.cnt, qi[0], e[0], E[0]);
                fmq_span_end<DI>(apow_in, sp.cnt, qi[1], e[1], E[1]);
            }
#pragma unroll
            for (int a = 0; a < 2; ++a)
#pragma unroll
                for (int i = 0; i < DI; ++i) E[a][i] = __shfl_sync(0xffffffffu, E[a][i], src_lane);
            if (tile == 0u) {
#pragma unroll
                for (int a = 0; a < 2; ++a)
#pragma unroll
                    for (int i = 0; i < DI; ++i) q0[a][i] = st->in_q[a * DI + i];
            } else {
                if (lane == 0u) {
#pragma unroll
                    for (int a = 0; a < 2; ++a)
#pragma unroll
                        for (int i = 0; i < DI; ++i) lb.E[((size_t) tile * 2 + a) * DI + i] = E[a][i];
                    __threadfence();
                    atomicExch(lb.flag + tile, 1u);
                }
                fmq_lookback<DI>(lb, tile, lane, lane_tile_in, wtile_in, q0, &st->error);
            }
            const bool final_tile = tile + 1u == g.ntiles;
            if (lane == 0u) {
#pragma unroll
                for (int a = 0; a < 2; ++a)
#pragma unroll
                    for (int i = 0; i < DI; ++i) {
                        float acc = E[a][i];
#pragma unroll
                        for (int k = 0; k < DI; ++k)
                            acc = fmaf(final_tile ? plast_in.m[i * DI + k] : cin.tile[i * DI + k], q0[a][k], acc);
                        if (final_tile) st->in_q[a * DI + i] = acc; /* the carry of the next call */
                        else lb.S[((size_t) tile * 2 + a) * DI + i] = acc;
                        carry_s[a * DI + i] = q0[a][i];
                    }
                if (!final_tile) {
                    __threadfence();
                    atomicExch(lb.flag + tile, 2u);
                }
            }
        }
        __syncthreads();
        /* thread state = (relative state) + (A^T)^t * carry, t = thread index in the tile */
        {
            const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
            float wq[2][DI];
#pragma unroll
            for (int a = 0; a < 2; ++a)
#pragma unroll
                for (int i = 0; i < DI; ++i) wq[a][i] = carry_s[a * DI + i];
            for (unsigned w = 0; w < warp; ++w) {
                float nw[2][DI];
#pragma unroll
                for (int a = 0; a < 2; ++a)
#pragma unroll
                    for (int i = 0; i < DI; ++i) {
                        float acc = 0.0f;
#pragma unroll
                        for (int k = 0; k < DI; ++k) acc = fmaf(cin.warp[i * DI + k], wq[a][k], acc);
                        nw[a][i] = acc;
                    }
#pragma unroll
                for (int a = 0; a < 2; ++a)
#pragma unroll
                    for (int i = 0; i < DI; ++i) wq[a][i] = nw[a][i];
            }
            const float *__restrict__ lt = lane_in + lane * (DI * DI);
#pragma unroll
            for (int i = 0; i < DI; ++i)
#pragma unroll
                for (int k = 0; k < DI; ++k) {
                    const float m = __ldg(lt + i * DI + k);
                    qi[0][i] = fmaf(m, wq[0][k], qi[0][i]);
                    qi[1][i] = fmaf(m, wq[1][k], qi[1][i]);
                }
        }
    }

    /* kept samples of this thread: offset of the first one in its span, and its slot in ybuf */
    const unsigned o_tile = fmq_kept_below(sp.S, g.r0, g.decim);
    unsigned next;
    float *yp; /* where this thread's next kept sample goes */
    {
        const unsigned o_first = fmq_kept_below(sp.F, g.r0, g.decim);
        const unsigned long long kfirst = (unsigned long long) g.r0 + (unsigned long long) o_first * g.decim;
        next = kfirst - sp.F >= (unsigned long long) T ? 0xffffffffu : (unsigned) (kfirst - sp.F);
        yp = ybuf + (o_first - o_tile) + threadIdx.x;
    }

    /* second pass: z (from the right state), x, v from a zero state; v kept every M-th sample */
    float pr, pi; /* z[F-1] */
    if (LIN > 0) {
        pr = fmq_step<LI1, FOI>(cin, qi[0], hi);
        pi = fmq_step<LI1, FOI>(cin, qi[1], hq);
    } else {
        pr = hi;
        pi = hq;
    }
    float eo[DO];
#pragma unroll
    for (int i = 0; i < DO; ++i) eo[i] = 0.0f;

    if (full) {
        /* measured on B200: without the input cascade the fully unrolled span is 8 % faster (more
         * freedom for the scheduler); with it the rolled form is 4 % faster (instruction fetch) */
#ifndef FMQ_UNR_MAX
#define FMQ_UNR_MAX 4 /* groups unrolled without the input cascade: 4 measured best for both spans (616 vs 602 GS/s at T = 128) */
#endif
        constexpr int UNR = LIN > 0 ? 1 : (GROUPS > FMQ_UNR_MAX ? FMQ_UNR_MAX : GROUPS);
        uint4 nx = fmq_load_group(in, sp.F, 0);
#pragma unroll(UNR)
        for (int grp = 0; grp < GROUPS; ++grp) {
            const uint4 v = nx;
            if (grp + 1 < GROUPS) nx = fmq_load_group(in, sp.F, grp + 1);
            const unsigned i0 = 8u * (unsigned) grp;
#pragma unroll
            for (int i = 0; i < 8; i += 2) {
                float z0r = FMQ_SI(v, i), z0i = FMQ_SQ(v, i), z1r = FMQ_SI(v, i + 1), z1i = FMQ_SQ(v, i + 1);
                if (LIN > 0) {
                    z0r = fmq_step<LI1, FOI>(cin, qi[0], z0r);
                    z0i = fmq_step<LI1, FOI>(cin, qi[1], z0i);
                    z1r = fmq_step<LI1, FOI>(cin, qi[0], z1r);
                    z1i = fmq_step<LI1, FOI>(cin, qi[1], z1i);
                }
                float x0, x1;
                fmq_discriminate2(pr, pi, z0r, z0i, z1r, z1i, x0, x1);
                pr = z1r;
                pi = z1i;
                const float v0 = fmq_step<LOUT, FOO>(cout, eo, x0);
                if (i0 + (unsigned) i == next) {
                    *yp++ = v0;
                    next += g.decim;
                }
                const float v1 = fmq_step<LOUT, FOO>(cout, eo, x1);
                if (i0 + (unsigned) (i + 1) == next) {
                    *yp++ = v1;
                    next += g.decim;
                }
            }
        }
    } else {
#pragma unroll 1
        for (unsigned i = 0; i < sp.cnt; ++i) {
            float zr, zi;
            fmq_load_sample(in, sp.F + i, zr, zi);
            if (LIN > 0) {
                zr = fmq_step<LI1, FOI>(cin, qi[0], zr);
                zi = fmq_step<LI1, FOI>(cin, qi[1], zi);
            }
            const float re = fmaf(zr, pr, zi * pi), im = fmaf(zi, pr, -(zr * pi));
            pr = zr;
            pi = zi;
            const float v = fmq_step<LOUT, FOO>(cout, eo, fmd_atan2(im, re));
            if (i == next) {
                *yp++ = v;
                next += g.decim;
            }
        }
    }

    /* output cascade: state at the start of every thread's span (zero at the tile start) */
    float qo[1][DO], zero[1][DO];
#pragma unroll
    for (int i = 0; i < DO; ++i) {
        qo[0][i] = eo[i];
        zero[0][i] = 0.0f;
    }
    fmq_block_scan<DO, 1>(qo, cout, lane_out, zero, tot_out);
#pragma unroll
    for (int i = 0; i < DO; ++i) qs[threadIdx.x * DO + i] = qo[0][i];
    if (sp.last) fmq_span_end<DO>(apow_out, sp.cnt, qo[0], eo, eout_tile + (size_t) tile * DO);
    __syncthreads();

    /* kept samples: add the thread-level zero-input response, store coalesced */
    const unsigned tile_end = min(sp.S + (unsigned) TILE, g.n);
    const unsigned count = fmq_kept_below(tile_end, g.r0, g.decim) - o_tile;
    for (unsigned o = threadIdx.x; o < count; o += FMQ_THREADS) {
        const unsigned ks = (unsigned) ((unsigned long long) g.r0 + (unsigned long long) (o_tile + o) * g.decim - sp.S);
        const unsigned owner = ks / T, k = ks % T;
        float y = ybuf[o + owner];
#pragma unroll
        for (int i = 0; i < DO; ++i) y = fmaf(ck[k * DO + i], qs[owner * DO + i], y);
        out[o_tile + o] = y;
    }
}

/* ------------------------------------------- tile-level zero-input response on the kept samples */

/*
 * y[o] += C A^k q_tile for the kept samples at tile offsets k < zir_len (beyond that the rows of
 * C A^k are below FMQ_ZIR_EPS and the host does not even launch threads for them).
 * Grid: x = tile, y = group of 256 kept samples inside the tile.
 */
template <int DO>
__global__ void __launch_bounds__(256)
fmq_zir_kernel(float *__restrict__ out, const FmqGeom g, const unsigned zir_len, const unsigned tile_samples,
               const float *__restrict__ ckn /* [TILE][DO] */, const float *__restrict__ q_tile /* [ntiles][DO] */) {
    const unsigned tile = blockIdx.x, S = tile * tile_samples;
    const unsigned o = fmq_kept_below(S, g.r0, g.decim) + blockIdx.y * 256u + threadIdx.x;
    if (o >= g.n_out) return;
    const unsigned long long ks = (unsigned long long) g.r0 + (unsigned long long) o * g.decim;
    if (ks >= (unsigned long long) S + zir_len || ks >= (unsigned long long) S + tile_samples) return;
    const unsigned k = (unsigned) (ks - S);
    const float *__restrict__ c = ckn + (size_t) k * DO;
    const float *__restrict__ q = q_tile + (size_t) tile * DO;
    float y = out[o];
#pragma unroll
    for (int i = 0; i < DO; ++i) y = fmaf(__ldg(c + i), __ldg(q + i), y);
    out[o] = y;
}

/* --------------------------------------------------------- scan of the per-tile end states */

/*
 * q[j+1] = P q[j] + E[j] over the tiles of a launch, q[0] = *carry, P = A^TILE.  Writes q[j] (the
 * state at the START of tile j) for every tile and leaves the state after the last tile in
 * *carry (its transition is `plast` = A^(samples in the last, possibly ragged, tile)).
 *
 * One CTA of 32 warps.  A warp owns a run of `runlen` consecutive tiles (a multiple of 32) and
 * walks it 32 tiles at a time with the same shuffle scan the CTAs of the main kernel use over
 * their threads (`ct` and `lane_tile` are the tables for the unit TILE instead of T): once from a
 * zero state to get the run's end state, then -- after the 32 run states have been chained with
 * `prun` = P^runlen -- again from the right state, storing.
 * When `last_sample` is given, also records the centred last input sample for the next call.
 */
template <int D, int NV>
__global__ void __launch_bounds__(FMQ_SCAN_THREADS)
fmq_scan_tiles_kernel(const float *__restrict__ E, float *__restrict__ Q, float *__restrict__ carry,
                      const fmq_consts ct, const float *__restrict__ lane_tile, const FmqMat prun, const FmqMat plast,
                      const unsigned ntiles, const unsigned runlen, const uint8_t *__restrict__ last_sample,
                      float *__restrict__ s_prev) {
    __shared__ float run_e[FMQ_SCAN_THREADS / 32][NV * D];
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const unsigned j0 = min(warp * runlen, ntiles), j1 = min(j0 + runlen, ntiles);

    float c[NV][D]; /* read before any thread may overwrite *carry */
#pragma unroll
    for (int a = 0; a < NV; ++a)
#pragma unroll
        for (int i = 0; i < D; ++i) c[a][i] = carry[a * D + i];

    /* pass 1: end state of the run from a zero state */
    float cw[NV][D];
#pragma unroll
    for (int a = 0; a < NV; ++a)
#pragma unroll
        for (int i = 0; i < D; ++i) cw[a][i] = 0.0f;
    for (unsigned base = j0; base < j1; base += 32u) {
        const unsigned j = base + lane;
        float v[NV][D];
#pragma unroll
        for (int a = 0; a < NV; ++a)
#pragma unroll
            for (int i = 0; i < D; ++i) v[a][i] = j < j1 ? E[((size_t) j * NV + a) * D + i] : 0.0f;
        fmq_warp_scan<D, NV>(v, ct, lane);
        float nw[NV][D];
#pragma unroll
        for (int a = 0; a < NV; ++a)
#pragma unroll
            for (int i = 0; i < D; ++i) {
                float acc = __shfl_sync(0xffffffffu, v[a][i], 31);
#pragma unroll
                for (int k = 0; k < D; ++k) acc = fmaf(ct.warp[i * D + k], cw[a][k], acc);
                nw[a][i] = acc;
            }
#pragma unroll
        for (int a = 0; a < NV; ++a)
#pragma unroll
            for (int i = 0; i < D; ++i) cw[a][i] = nw[a][i];
    }
    if (lane == 0u) {
#pragma unroll
        for (int a = 0; a < NV; ++a)
#pragma unroll
            for (int i = 0; i < D; ++i) run_e[warp][a * D + i] = cw[a][i];
    }
    __syncthreads();
    /* state at the start of this warp's run (runs before it are full: runlen tiles each) */
    for (unsigned w = 0; w < warp; ++w) {
        float nw[NV][D];
#pragma unroll
        for (int a = 0; a < NV; ++a)
#pragma unroll
            for (int i = 0; i < D; ++i) {
                float acc = run_e[w][a * D + i];
#pragma unroll
                for (int k = 0; k < D; ++k) acc = fmaf(prun.m[i * D + k], c[a][k], acc);
                nw[a][i] = acc;
            }
#pragma unroll
        for (int a = 0; a < NV; ++a)
#pragma unroll
            for (int i = 0; i < D; ++i) c[a][i] = nw[a][i];
    }
    /* pass 2: the same walk from the right state, storing the state at the start of every tile */
    float lp[D * D];
#pragma unroll
    for (int i = 0; i < D * D; ++i) lp[i] = __ldg(lane_tile + lane * (D * D) + i);
    for (unsigned base = j0; base < j1; base += 32u) {
        const unsigned j = base + lane;
        float e[NV][D], v[NV][D];
#pragma unroll
        for (int a = 0; a < NV; ++a)
#pragma unroll
            for (int i = 0; i < D; ++i) v[a][i] = e[a][i] = j < j1 ? E[((size_t) j * NV + a) * D + i] : 0.0f;
        fmq_warp_scan<D, NV>(v, ct, lane);
        float q[NV][D];
#pragma unroll
        for (int a = 0; a < NV; ++a)
#pragma unroll
            for (int i = 0; i < D; ++i) {
                const float t = __shfl_up_sync(0xffffffffu, v[a][i], 1u);
                float acc = lane ? t : 0.0f;
#pragma unroll
                for (int k = 0; k < D; ++k) acc = fmaf(lp[i * D + k], c[a][k], acc);
                q[a][i] = acc;
            }
        if (j < j1) {
#pragma unroll
            for (int a = 0; a < NV; ++a)
#pragma unroll
                for (int i = 0; i < D; ++i) Q[((size_t) j * NV + a) * D + i] = q[a][i];
            if (j + 1u == ntiles) { /* the state after the last tile: the carry for the next call */
#pragma unroll
                for (int a = 0; a < NV; ++a)
#pragma unroll
                    for (int i = 0; i < D; ++i) {
                        float acc = e[a][i];
#pragma unroll
                        for (int k = 0; k < D; ++k) acc = fmaf(plast.m[i * D + k], q[a][k], acc);
                        carry[a * D + i] = acc;
                    }
                if (last_sample) {
                    float si, sq;
                    fmq_load_sample(last_sample, 0u, si, sq);
                    s_prev[0] = si;
                    s_prev[1] = sq;
                }
            }
        }
        float nw[NV][D];
#pragma unroll
        for (int a = 0; a < NV; ++a)
#pragma unroll
            for (int i = 0; i < D; ++i) {
                float acc = __shfl_sync(0xffffffffu, v[a][i], 31);
#pragma unroll
                for (int k = 0; k < D; ++k) acc = fmaf(ct.warp[i * D + k], c[a][k], acc);
                nw[a][i] = acc;
            }
#pragma unroll
        for (int a = 0; a < NV; ++a)
#pragma unroll
            for (int i = 0; i < D; ++i) c[a][i] = nw[a][i];
    }
}

/* ------------------------------------------------------------- egress: float32 -> signed 16 bit */

/*
 * FMD_FLAG_S16: what the `sox ... -e signed-integer -b16` stage of the reference's pipelines does to
 * the float stream (README.md:22,29) -- full scale 1.0 = 32768, round to nearest, clip -- applied to
 * the decimated outputs, so that `demodulator -Q ... -S48000` feeds dsd / multimon-ng directly.
 */
__global__ void __launch_bounds__(256)
fmq_to_s16_kernel(const float *__restrict__ y, short *__restrict__ o, const unsigned n) {
    const unsigned i = blockIdx.x * 256u + threadIdx.x;
    if (i >= n) return;
    const float v = fminf(fmaxf(rintf(y[i] * 32768.0f), -32768.0f), 32767.0f);
    o[i] = (short) (int) v;
}

#endif /* FMD_QUALITY_CUH */
