/*
 * fmd_q1_fused.cuh -- -q1 (correctIq, /root/reference/src/matrix.c:120-140) conditioning AND
 * demodulation in one kernel: the conditioned samples never go to HBM.
 *
 * The reference conditions a chunk from both ends at once: step k takes complex k from the front
 * and complex nc-1-k from the back, subtracts the running offsets and moves them; the offsets run
 * on from chunk to chunk.  A CTA (tile) of this kernel owns FMQ1_TILE = 2048 consecutive STEPS of
 * one chunk, i.e. two runs of 2048 complex samples -- front run [s0, s0+2048), back run
 * [nc-s0-2048, nc-s0):
 *
 *   1. offsets: every thread reads its 8 steps (16 bytes from each run), the CTA scan gives it the
 *      offsets at its first step (tile start values from fmq_scan_tiles_kernel, as in
 *      fmd_correctiq.cuh), then it walks its steps with the reference's own float operations and
 *      leaves the conditioned samples of both runs in shared memory;
 *   2. a few samples past each run (the stencils' reach): ahead of the front run the offsets are
 *      walked on from the tile's end values (or, in the chunk's last tile, the front run simply
 *      continues into this tile's own back run); ahead of the back run -- steps BEFORE this tile --
 *      they are walked backwards from the tile's start values (the recurrence is invertible);
 *      past the end of the chunk it is the zero extension;
 *   3. input stencil, pair discriminator, atan2 for both runs into shared memory, then the output
 *      stencil and coalesced 16-byte stores.
 * The first outputs of every chunk (where the reference's sweep is not in steady state) belong to
 * the op-for-op path as everywhere else: the kernel leaves the conditioned samples of the chunk's
 * first FMQ1_HEADS complex samples in a side buffer for it.
 *
 * HBM traffic: 2 B (tile-end pass) + 2 B + 2 B per sample instead of the 22 B of the three-pass form.
 */
#ifndef FMD_Q1_FUSED_CUH
#define FMD_Q1_FUSED_CUH

#include "fmd_correctiq.cuh"
#include "fmd_types.h"

#define FMQ1_LA 16u    /* complex samples kept past each run: 2 (NTO - 1) + NTI - 1 <= 10 */
#define FMQ1_HEADS 128u /* conditioned complex samples per chunk handed to the op-for-op head path */

/* shared-memory index of complex sample i of a run: two float2 of padding per 8 samples make the
 * stride between threads 80 bytes, conflict-free for the 16-byte accesses of both phases */
__device__ __forceinline__ unsigned fmq1_idx(unsigned i) { return i + 2u * (i >> 3); }
#define FMQ1_RUN_SLOTS (((FMQ1_TILE + FMQ1_LA) / 8u) * 10u) /* float2 slots per run */
#define FMQ1_X_SLOTS (FMQ1_TILE / 2u + 8u)                  /* discriminator outputs per run */
#define FMQ1_SMEM_BYTES ((2u * FMQ1_RUN_SLOTS * 2u + 2u * FMQ1_X_SLOTS) * 4u)

template <int NTI, int NTO>
__global__ void __launch_bounds__(FMQ_THREADS)
fmd_q1_fused_kernel(const uint8_t *__restrict__ in, float *__restrict__ out, const FmdQ1Geom g, const fmq_consts c,
                    const float *__restrict__ lane_tab, const float *__restrict__ q_tile /* [ntiles][2] */,
                    const FmdStencil st, const unsigned head, float2 *__restrict__ heads /* [nchunks][FMQ1_HEADS] */) {
    extern __shared__ float4 fmq1_smem4[];
    float2 *cf = reinterpret_cast<float2 *>(fmq1_smem4); /* front run, conditioned */
    float2 *cb = cf + FMQ1_RUN_SLOTS;                     /* back run                */
    float *xf = reinterpret_cast<float *>(cb + FMQ1_RUN_SLOTS);
    float *xb = xf + FMQ1_X_SLOTS;
    __shared__ float tot[FMQ_WARPS * 2];

    constexpr unsigned TS = FMQ1_TILE;
    constexpr unsigned LA = 2u * (NTO - 1) + (NTI > 1 ? NTI - 1 : 0); /* reach past a run, complex samples */
    static_assert(LA <= FMQ1_LA, "look-ahead buffer too small");
    const unsigned t = threadIdx.x;
    const unsigned tiles_per_chunk = g.steps_per_chunk / TS;
    const unsigned chunk = blockIdx.x / tiles_per_chunk;
    const unsigned s0 = (blockIdx.x - chunk * tiles_per_chunk) * TS; /* first step of the tile, chunk-relative */
    const unsigned nc = 2u * g.steps_per_chunk;
    const uint8_t *__restrict__ cin = in + (unsigned long long) chunk * g.chunk_bytes;
    const unsigned short *__restrict__ smp = reinterpret_cast<const unsigned short *>(cin); /* complex sample c: I low, Q high */
    const bool last_tile = s0 + TS == g.steps_per_chunk; /* its front run runs straight into its back run */

#define FMQ1_W(v, i) ((i) < 2 ? (v).x : (i) < 4 ? (v).y : (i) < 6 ? (v).z : (v).w)
#define FMQ1_I(v, i) ((FMQ1_W(v, i) >> (16 * ((i) & 1))) & 0xffu)
#define FMQ1_Q(v, i) ((FMQ1_W(v, i) >> (16 * ((i) & 1) + 8)) & 0xffu)
    /* ---- 1. the offsets at this thread's first step, then its 8 steps ---- */
    const unsigned k0 = s0 + 8u * t;
    const uint4 f = __ldg(reinterpret_cast<const uint4 *>(smp + k0));
    const uint4 b = __ldg(reinterpret_cast<const uint4 *>(smp + (nc - 8u - k0)));
    float v[2][1] = {{0.0f}, {0.0f}};
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        v[0][0] = fmaf(g.a, v[0][0], g.esr * (float) (FMQ1_I(f, j) + FMQ1_I(b, 7 - j)));
        v[1][0] = fmaf(g.a, v[1][0], g.esr * (float) (FMQ1_Q(f, j) + FMQ1_Q(b, 7 - j)));
    }
    float carry[2][1];
    carry[0][0] = __ldg(q_tile + 2 * (size_t) blockIdx.x);
    carry[1][0] = __ldg(q_tile + 2 * (size_t) blockIdx.x + 1);
    fmq_block_scan<1, 2>(v, c, lane_tab, carry, tot);
    float off0 = v[0][0], off1 = v[1][0];
    {
        /* the reference's loop body, operation for operation (matrix.c:129-138); two steps per 16-byte store */
        float4 *__restrict__ pf = reinterpret_cast<float4 *>(cf + fmq1_idx(8u * t));
        float4 *__restrict__ pb = reinterpret_cast<float4 *>(cb + fmq1_idx(TS - 8u - 8u * t));
#pragma unroll
        for (int j = 0; j < 8; j += 2) {
            float fi[2], fq[2], bi[2], bq[2];
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                fi[u] = __fsub_rn((float) FMQ1_I(f, j + u), off0);
                bi[u] = __fsub_rn((float) FMQ1_I(b, 7 - (j + u)), off0);
                fq[u] = __fsub_rn((float) FMQ1_Q(f, j + u), off1);
                bq[u] = __fsub_rn((float) FMQ1_Q(b, 7 - (j + u)), off1);
                off0 = __fadd_rn(off0, __fmul_rn(__fadd_rn(fi[u], bi[u]), g.esr));
                off1 = __fadd_rn(off1, __fmul_rn(__fadd_rn(fq[u], bq[u]), g.esr));
            }
            pf[j >> 1] = make_float4(fi[0], fq[0], fi[1], fq[1]);
            pb[3 - (j >> 1)] = make_float4(bi[1], bq[1], bi[0], bq[0]); /* back samples 6-j, 7-j of the group */
        }
    }
    /* ---- 2. what lies past the runs ---- */
    if (t == FMQ_THREADS - 1u && !last_tile) {
        /* ahead of the front run: the front samples of the steps that follow this tile, conditioned
         * with the offsets walked on from the tile's end (off0/off1 of the last thread) */
        float o0 = off0, o1 = off1;
        for (unsigned i = 0; i < LA; ++i) {
            const unsigned k = s0 + TS + i; /* < steps_per_chunk: there is at least one more tile */
            const unsigned uf = smp[k], ub = smp[nc - 1u - k];
            const float fi = __fsub_rn((float) (uf & 0xffu), o0), fq = __fsub_rn((float) (uf >> 8), o1);
            const float bi = __fsub_rn((float) (ub & 0xffu), o0), bq = __fsub_rn((float) (ub >> 8), o1);
            cf[fmq1_idx(TS + i)] = make_float2(fi, fq);
            o0 = __fadd_rn(o0, __fmul_rn(__fadd_rn(fi, bi), g.esr));
            o1 = __fadd_rn(o1, __fmul_rn(__fadd_rn(fq, bq), g.esr));
        }
    }
    if (t == 0u) {
        /* ahead of the back run (higher addresses): the back samples of the steps BEFORE this tile,
         * with the offsets walked backwards from the tile's start values,
         *     off[k] = (off[k+1] - esr (u_front + u_back)) / (1 - 2 esr);
         * for the chunk's first tile it is the zero extension past the chunk */
        float o0 = carry[0][0], o1 = carry[1][0];
        const float ra = 1.0f / g.a;
        for (unsigned i = 0; i < LA; ++i) {
            float2 val = make_float2(0.0f, 0.0f);
            if (s0 > i) {
                const unsigned k = s0 - 1u - i;
                const unsigned uf = smp[k], ub = smp[nc - 1u - k];
                o0 = (o0 - g.esr * (float) ((uf & 0xffu) + (ub & 0xffu))) * ra;
                o1 = (o1 - g.esr * (float) ((uf >> 8) + (ub >> 8))) * ra;
                val = make_float2((float) (ub & 0xffu) - o0, (float) (ub >> 8) - o1);
            }
            cb[fmq1_idx(TS + i)] = val;
        }
    }
    __syncthreads();
    if (last_tile) { /* complex nc/2 + i follows complex nc/2 - 1: the back run IS what lies ahead of the front run */
        if (t < LA) cf[fmq1_idx(TS + t)] = cb[fmq1_idx(t)];
        __syncthreads();
    }
    if (s0 == 0u && heads && t < FMQ1_HEADS) heads[(size_t) chunk * FMQ1_HEADS + t] = cf[fmq1_idx(t)];

    /* ---- 3. stencil, discriminator, atan2: 4 outputs per thread and run (+ NTO - 1 past the run) ---- */
    auto disc = [&](const float2 *__restrict__ run, unsigned kx) -> float {
        /* z[c] = s[c] + sum_{u>=1} tin[u] s[c+u] (+ dcq on the Q lane): the stencil divided by its first tap,
         * as in fmd_pass_core */
        float2 s[NTI > 0 ? NTI + 1 : 2];
#pragma unroll
        for (int u = 0; u < (NTI > 0 ? NTI + 1 : 2); ++u) s[u] = run[fmq1_idx(2u * kx + (unsigned) u)];
        float ar = s[0].x, ai = s[0].y, br = s[1].x, bi = s[1].y;
        if (NTI > 0) {
            ai = __fadd_rn(ai, st.dcq);
            bi = __fadd_rn(bi, st.dcq);
#pragma unroll
            for (int u = 1; u < NTI; ++u) {
                ar = fmaf(st.tin[u], s[u].x, ar);
                ai = fmaf(st.tin[u], s[u].y, ai);
                br = fmaf(st.tin[u], s[u + 1].x, br);
                bi = fmaf(st.tin[u], s[u + 1].y, bi);
            }
        }
        const float re = fmaf(ar, br, ai * bi);
        const float im = fmaf(-ar, bi, ai * br);
        return fmd_atan2(im, re);
    };
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        xf[4u * t + (unsigned) j] = disc(cf, 4u * t + (unsigned) j);
        xb[4u * t + (unsigned) j] = disc(cb, 4u * t + (unsigned) j);
    }
    if (t < (unsigned) (NTO - 1)) {
        xf[TS / 2u + t] = disc(cf, TS / 2u + t);
        /* past the end of the chunk the discriminator output is exactly 0 (demodRet slack, matrix.c:184) */
        xb[TS / 2u + t] = s0 == 0u ? 0.0f : disc(cb, TS / 2u + t);
    }
    __syncthreads();

    /* ---- 4. output stencil, coalesced stores ---- */
    const unsigned chunk_out = g.steps_per_chunk; /* nc / 2 */
    float *__restrict__ oc = out + (unsigned long long) chunk * chunk_out;
    float yf[4], yb[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        float af = st.dco, ab = st.dco;
#pragma unroll
        for (int u = 0; u < NTO; ++u) {
            af = fmaf(st.tout[u], xf[4u * t + (unsigned) (j + u)], af);
            ab = fmaf(st.tout[u], xb[4u * t + (unsigned) (j + u)], ab);
        }
        yf[j] = af;
        yb[j] = ab;
    }
    /* the chunk head [0, head) belongs to the op-for-op path */
    if (!(s0 == 0u && 4u * t < head)) *reinterpret_cast<float4 *>(oc + s0 / 2u + 4u * t) = make_float4(yf[0], yf[1], yf[2], yf[3]);
    *reinterpret_cast<float4 *>(oc + (nc - s0 - TS) / 2u + 4u * t) = make_float4(yb[0], yb[1], yb[2], yb[3]);
#undef FMQ1_W
#undef FMQ1_I
#undef FMQ1_Q
}

#endif /* FMD_Q1_FUSED_CUH */
