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

#define FMQ1_SUBS 8u /* tiles one CTA walks one after the other: the offsets run on from tile to tile inside it */

/* byte B of a raw word as an exact float: 0x4B0000uu = 8388608 + u */
template <int B>
__device__ __forceinline__ float fmq1_byte(uint32_t w) {
    return __uint_as_float(__byte_perm(w, 0x4B000000u, 0x7540 | B)) - 8388608.0f;
}

/* discriminator output from NTI + 1 (or 2) consecutive conditioned samples s[0..]:
 * z[c] = s[c] + sum_{u>=1} tin[u] s[c+u] (+ dcq on the Q lane) -- the stencil divided by its first tap,
 * as in fmd_pass_core -- then atan2 of z[0] conj z[1] */
template <int NTI>
__device__ __forceinline__ float fmq1_disc(const float2 *s, const FmdStencil &st) {
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
    return fmd_atan2(fmaf(-ar, bi, ai * br), fmaf(ar, br, ai * bi));
}

/* complex sample `pos` (0..7, a compile-time value after unrolling) of a 16-byte group: I low byte, Q high byte */
__device__ __forceinline__ unsigned fmq1_sample(const uint4 &v, int pos) {
    const uint32_t w = pos < 2 ? v.x : pos < 4 ? v.y : pos < 6 ? v.z : v.w;
    return (w >> (16 * (pos & 1))) & 0xffffu;
}

/* the thread's 8 steps of tile `tile` as floats, step order: step j = front sample j, back sample 7 - j */
__device__ __forceinline__ void fmq1_load_steps(const uint8_t *__restrict__ in, const FmdQ1Geom &g, unsigned tile, unsigned t,
                                                float (&fI)[8], float (&fQ)[8], float (&bI)[8], float (&bQ)[8]) {
    const unsigned tiles_per_chunk = g.steps_per_chunk / FMQ1_TILE;
    const unsigned chunk = tile / tiles_per_chunk;
    const unsigned k0 = (tile - chunk * tiles_per_chunk) * FMQ1_TILE + 8u * t;
    const unsigned nc = 2u * g.steps_per_chunk;
    const unsigned short *__restrict__ smp = reinterpret_cast<const unsigned short *>(in + (unsigned long long) chunk * g.chunk_bytes);
    const uint4 f = __ldg(reinterpret_cast<const uint4 *>(smp + k0));
    const uint4 b = __ldg(reinterpret_cast<const uint4 *>(smp + (nc - 8u - k0)));
    const uint32_t fw[4] = {f.x, f.y, f.z, f.w}, bw[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
    for (int w = 0; w < 4; ++w) {
        fI[2 * w] = fmq1_byte<0>(fw[w]);
        fQ[2 * w] = fmq1_byte<1>(fw[w]);
        fI[2 * w + 1] = fmq1_byte<2>(fw[w]);
        fQ[2 * w + 1] = fmq1_byte<3>(fw[w]);
        /* back group, address position 2w / 2w+1 = step 7-2w / 6-2w */
        bI[7 - 2 * w] = fmq1_byte<0>(bw[w]);
        bQ[7 - 2 * w] = fmq1_byte<1>(bw[w]);
        bI[6 - 2 * w] = fmq1_byte<2>(bw[w]);
        bQ[6 - 2 * w] = fmq1_byte<3>(bw[w]);
    }
}

/*
 * End values of the offsets, from zero, over every CTA's run of up to FMQ1_SUBS tiles (what
 * fmq_scan_tiles_kernel then chains through the launch).  In exact arithmetic the end value is the
 * weighted sum  esr * sum_k a^(steps after k) (u_front[k] + u_back[k]):  a lane weights its 8 steps,
 * a butterfly sums the warp, every warp carries its own total from tile to tile (x a^2048), and
 * only at the very end the 8 warp totals are combined -- one barrier per CTA, no scan.
 */
__global__ void __launch_bounds__(FMQ_THREADS)
fmd_q1_ends_kernel(const uint8_t *__restrict__ in, const FmdQ1Geom g, const fmq_consts c, const float *__restrict__ lane_tab,
                   const unsigned ntiles, float *__restrict__ e_run /* [nruns][2] */) {
    __shared__ float wt[FMQ_WARPS * 2];
    const unsigned t = threadIdx.x, lane = t & 31u, warp = t >> 5;
    const float wl = __ldg(lane_tab + (31u - lane)); /* (a^8)^(31 - lane): from the end of this lane's steps to the end of the warp's */
    float acc0 = 0.0f, acc1 = 0.0f;
    unsigned tile = blockIdx.x * FMQ1_SUBS;
    for (unsigned sub = 0; sub < FMQ1_SUBS && tile < ntiles; ++sub, ++tile) {
        float fI[8], fQ[8], bI[8], bQ[8];
        fmq1_load_steps(in, g, tile, t, fI, fQ, bI, bQ);
        float v0 = 0.0f, v1 = 0.0f;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            v0 = fmaf(g.a, v0, fI[j] + bI[j]);
            v1 = fmaf(g.a, v1, fQ[j] + bQ[j]);
        }
        v0 *= wl;
        v1 *= wl;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            v0 += __shfl_xor_sync(0xffffffffu, v0, o);
            v1 += __shfl_xor_sync(0xffffffffu, v1, o);
        }
        acc0 = fmaf(c.tile[0], acc0, v0); /* earlier tiles of this warp move a whole tile further from the end */
        acc1 = fmaf(c.tile[0], acc1, v1);
    }
    if (lane == 0u) {
        wt[2 * warp] = acc0;
        wt[2 * warp + 1] = acc1;
    }
    __syncthreads();
    if (t == 0u) {
        float e0 = 0.0f, e1 = 0.0f;
        for (unsigned w = 0; w < FMQ_WARPS; ++w) { /* Horner over the warps: a^256 per warp */
            e0 = fmaf(c.warp[0], e0, wt[2 * w]);
            e1 = fmaf(c.warp[0], e1, wt[2 * w + 1]);
        }
        e_run[2 * (size_t) blockIdx.x] = g.esr * e0;
        e_run[2 * (size_t) blockIdx.x + 1] = g.esr * e1;
    }
}

/*
 * Everything a thread needs of its own 16 samples stays in registers; shared memory only carries
 * what crosses a thread: the first NTI - 1 conditioned samples of every group (the next group's
 * input-stencil halo), the first NTO - 1 discriminator outputs of every thread, and the few samples /
 * outputs past the two runs.  (A first version kept both conditioned runs in shared memory and read
 * them back: 387 GS/s for the whole -q1 pipeline; this one: see profiles/.)
 */
template <int NTI, int NTO>
__global__ void __launch_bounds__(FMQ_THREADS)
fmd_q1_fused_kernel(const uint8_t *__restrict__ in, float *__restrict__ out, const FmdQ1Geom g, const fmq_consts c,
                    const float *__restrict__ lane_tab, const float *__restrict__ q_run /* [nruns][2]: offsets at the start of every CTA's run */,
                    const unsigned ntiles, const FmdStencil st, const unsigned head,
                    float2 *__restrict__ heads /* [nchunks][FMQ1_HEADS] */) {
    constexpr unsigned TS = FMQ1_TILE;
    constexpr int HC = NTI > 1 ? NTI - 1 : 0; /* conditioned samples needed from the following group */
    constexpr int NB = NTO - 1;               /* discriminator outputs needed from the following thread */
    constexpr unsigned LA = 2u * NB + HC;     /* reach past a run, complex samples */
    static_assert(LA <= FMQ1_LA, "look-ahead buffer too small");
    constexpr int HC1 = HC > 0 ? HC : 1, NB1 = NB > 0 ? NB : 1;
    __shared__ float2 hf[FMQ_THREADS * HC1], hb[FMQ_THREADS * HC1]; /* first HC samples (address order) of every group */
    __shared__ float2 laf[FMQ1_LA + 4], lab[FMQ1_LA + 4];           /* samples past the front / back run */
    __shared__ float xhf[(FMQ_THREADS + 1) * NB1], xhb[(FMQ_THREADS + 1) * NB1]; /* first NB outputs of every thread; [256]: past the run */
    __shared__ float tot[FMQ_WARPS * 2];
    __shared__ float carry_s[2]; /* the offsets at the start of the next tile of this CTA's run */

    const unsigned t = threadIdx.x;
    const unsigned tiles_per_chunk = g.steps_per_chunk / TS;
    const unsigned nc = 2u * g.steps_per_chunk;
    unsigned tile = blockIdx.x * FMQ1_SUBS;
#pragma unroll 1
    for (unsigned sub = 0; sub < FMQ1_SUBS && tile < ntiles; ++sub, ++tile) {
    const unsigned chunk = tile / tiles_per_chunk;
    const unsigned s0 = (tile - chunk * tiles_per_chunk) * TS; /* first step of the tile, chunk-relative */
    const uint8_t *__restrict__ cin = in + (unsigned long long) chunk * g.chunk_bytes;
    const unsigned short *__restrict__ smp = reinterpret_cast<const unsigned short *>(cin); /* complex sample c: I low, Q high */
    const bool last_tile = s0 + TS == g.steps_per_chunk; /* its front run runs straight into its back run */

    /* ---- 1. the thread's 8 steps as floats: step j = front sample j, back sample 7 - j of the two 16-byte groups ---- */
    float fI[8], fQ[8], bI[8], bQ[8]; /* step order */
    fmq1_load_steps(in, g, tile, t, fI, fQ, bI, bQ);
    /* The two threads that will evaluate what lies past the runs (below) fetch the bytes for it NOW,
     * with everybody's loads: fetched inside their serial loops they cost the whole CTA -- waiting at
     * the barrier -- LA dependent global-memory latencies per tile (ncu: barrier 42 % of the stalls). */
    const bool walk_on = t == FMQ_THREADS - 1u && !last_tile, walk_back = t == 0u && s0 > 0u;
    uint4 lf0 = make_uint4(0u, 0u, 0u, 0u), lb0 = lf0, lf1 = lf0, lb1 = lf0;
    if (walk_on) { /* steps s0+TS+i: front samples ascending from s0+TS, back samples descending from nc-1-(s0+TS) */
        lf0 = __ldg(reinterpret_cast<const uint4 *>(smp + (s0 + TS)));
        lb0 = __ldg(reinterpret_cast<const uint4 *>(smp + (nc - 8u - (s0 + TS))));
        if (LA > 8u) {
            lf1 = __ldg(reinterpret_cast<const uint4 *>(smp + (s0 + TS + 8u)));
            lb1 = __ldg(reinterpret_cast<const uint4 *>(smp + (nc - 16u - (s0 + TS))));
        }
    }
    if (walk_back) { /* steps s0-1-i: front samples descending from s0-1, back samples ascending from nc-s0 */
        lf0 = __ldg(reinterpret_cast<const uint4 *>(smp + (s0 - 8u)));
        lb0 = __ldg(reinterpret_cast<const uint4 *>(smp + (nc - s0)));
        if (LA > 8u) {
            lf1 = __ldg(reinterpret_cast<const uint4 *>(smp + (s0 - 16u)));
            lb1 = __ldg(reinterpret_cast<const uint4 *>(smp + (nc - s0 + 8u)));
        }
    }
    /* offsets at this thread's first step: 8 steps from zero, CTA scan with the tile's start values */
    float v[2][1] = {{0.0f}, {0.0f}};
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        v[0][0] = fmaf(g.a, v[0][0], g.esr * (fI[j] + bI[j]));
        v[1][0] = fmaf(g.a, v[1][0], g.esr * (fQ[j] + bQ[j]));
    }
    /* the tile's start values: from the scan over the runs for the CTA's first tile, then the offsets the
     * previous tile's last thread ended with (the reference's own float trajectory) */
    float carry[2][1];
    if (sub == 0u) {
        carry[0][0] = __ldg(q_run + 2 * (size_t) blockIdx.x);
        carry[1][0] = __ldg(q_run + 2 * (size_t) blockIdx.x + 1);
    } else {
        carry[0][0] = carry_s[0];
        carry[1][0] = carry_s[1];
    }
    fmq_block_scan<1, 2>(v, c, lane_tab, carry, tot);
    float off0 = v[0][0], off1 = v[1][0];
    /* the reference's loop body, operation for operation (matrix.c:129-138), in place */
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        fI[j] = __fsub_rn(fI[j], off0);
        bI[j] = __fsub_rn(bI[j], off0);
        fQ[j] = __fsub_rn(fQ[j], off1);
        bQ[j] = __fsub_rn(bQ[j], off1);
        off0 = __fadd_rn(off0, __fmul_rn(__fadd_rn(fI[j], bI[j]), g.esr));
        off1 = __fadd_rn(off1, __fmul_rn(__fadd_rn(fQ[j], bQ[j]), g.esr));
    }
    /* ---- 2. what crosses a thread, and what lies past the runs ---- */
    if (t == FMQ_THREADS - 1u) { /* read at the top of the next tile, after this tile's barriers */
        carry_s[0] = off0;
        carry_s[1] = off1;
    }
#pragma unroll
    for (int h = 0; h < HC; ++h) {
        hf[t * HC1 + h] = make_float2(fI[h], fQ[h]);
        hb[t * HC1 + h] = make_float2(bI[7 - h], bQ[7 - h]); /* address position h of the back group */
    }
    if (walk_on) {
        /* ahead of the front run: the front samples of the steps that follow this tile, conditioned
         * with the offsets walked on from the tile's end (off0/off1 of the last thread) */
        float o0 = off0, o1 = off1;
#pragma unroll
        for (int i = 0; i < (int) LA; ++i) {
            const unsigned uf = fmq1_sample(i < 8 ? lf0 : lf1, i & 7), ub = fmq1_sample(i < 8 ? lb0 : lb1, 7 - (i & 7));
            const float fi = __fsub_rn((float) (uf & 0xffu), o0), fq = __fsub_rn((float) (uf >> 8), o1);
            const float bi = __fsub_rn((float) (ub & 0xffu), o0), bq = __fsub_rn((float) (ub >> 8), o1);
            laf[i] = make_float2(fi, fq);
            o0 = __fadd_rn(o0, __fmul_rn(__fadd_rn(fi, bi), g.esr));
            o1 = __fadd_rn(o1, __fmul_rn(__fadd_rn(fq, bq), g.esr));
        }
    }
    if (last_tile && t >= FMQ_THREADS - 2u) {
        /* complex nc/2 + i follows complex nc/2 - 1: the back run IS what lies ahead of the front run
         * (its first 8 samples are the last thread's back group, the next 8 the thread's before it) */
#pragma unroll
        for (int m = 0; m < 8; ++m) {
            const unsigned idx = (FMQ_THREADS - 1u - t) * 8u + (unsigned) m;
            if (idx < LA) laf[idx] = make_float2(bI[7 - m], bQ[7 - m]);
        }
    }
    if (t == 0u) {
        /* ahead of the back run (higher addresses): the back samples of the steps BEFORE this tile,
         * with the offsets walked backwards from the tile's start values,
         *     off[k] = (off[k+1] - esr (u_front + u_back)) / (1 - 2 esr);
         * for the chunk's first tile it is the zero extension past the chunk */
        float o0 = carry[0][0], o1 = carry[1][0];
        const float ra = 1.0f / g.a;
#pragma unroll
        for (int i = 0; i < (int) LA; ++i) {
            float2 val = make_float2(0.0f, 0.0f);
            if (walk_back) {
                const unsigned uf = fmq1_sample(i < 8 ? lf0 : lf1, 7 - (i & 7)), ub = fmq1_sample(i < 8 ? lb0 : lb1, i & 7);
                o0 = (o0 - g.esr * (float) ((uf & 0xffu) + (ub & 0xffu))) * ra;
                o1 = (o1 - g.esr * (float) ((uf >> 8) + (ub >> 8))) * ra;
                val = make_float2((float) (ub & 0xffu) - o0, (float) (ub >> 8) - o1);
            }
            lab[i] = val;
        }
    }
    if (s0 == 0u && heads && t < FMQ1_HEADS / 8u) {
        float4 *__restrict__ hp = reinterpret_cast<float4 *>(heads + (size_t) chunk * FMQ1_HEADS + 8u * t);
#pragma unroll
        for (int j = 0; j < 8; j += 2) hp[j >> 1] = make_float4(fI[j], fQ[j], fI[j + 1], fQ[j + 1]);
    }
    __syncthreads();

    /* ---- 3. stencil, discriminator, atan2: 4 outputs per run from registers + the following group's halo ---- */
    float2 sf[8 + HC1], sb[8 + HC1]; /* address order */
#pragma unroll
    for (int m = 0; m < 8; ++m) {
        sf[m] = make_float2(fI[m], fQ[m]);
        sb[m] = make_float2(bI[7 - m], bQ[7 - m]);
    }
#pragma unroll
    for (int h = 0; h < HC; ++h) {
        sf[8 + h] = t + 1u < FMQ_THREADS ? hf[(t + 1u) * HC1 + h] : laf[h]; /* the next group of the front run */
        sb[8 + h] = t > 0u ? hb[(t - 1u) * HC1 + h] : lab[h];               /* the back run's next group is the PREVIOUS thread's */
    }
    float xF[4 + NB1], xB[4 + NB1];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        xF[j] = fmq1_disc<NTI>(sf + 2 * j, st);
        xB[j] = fmq1_disc<NTI>(sb + 2 * j, st);
    }
#pragma unroll
    for (int u = 0; u < NB; ++u) {
        xhf[t * NB1 + u] = xF[u];
        xhb[t * NB1 + u] = xB[u];
    }
    if (t < (unsigned) NB) {
        /* the outputs past the runs, from the look-ahead samples */
        xhf[FMQ_THREADS * NB1 + t] = fmq1_disc<NTI>(laf + 2u * t, st);
        /* past the end of the chunk the discriminator output is exactly 0 (demodRet slack, matrix.c:184) */
        xhb[FMQ_THREADS * NB1 + t] = s0 == 0u ? 0.0f : fmq1_disc<NTI>(lab + 2u * t, st);
    }
    __syncthreads();
#pragma unroll
    for (int u = 0; u < NB; ++u) {
        xF[4 + u] = xhf[(t + 1u) * NB1 + u];                                   /* thread 255 reads the slot past the run */
        xB[4 + u] = t > 0u ? xhb[(t - 1u) * NB1 + u] : xhb[FMQ_THREADS * NB1 + u]; /* previous thread = higher addresses */
    }

    /* ---- 4. output stencil, coalesced stores ---- */
    const unsigned chunk_out = g.steps_per_chunk; /* nc / 2 */
    float *__restrict__ oc = out + (unsigned long long) chunk * chunk_out;
    float yf[4], yb[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        float af = st.dco, ab = st.dco;
#pragma unroll
        for (int u = 0; u < NTO; ++u) {
            af = fmaf(st.tout[u], xF[j + u], af);
            ab = fmaf(st.tout[u], xB[j + u], ab);
        }
        yf[j] = af;
        yb[j] = ab;
    }
    /* the chunk head [0, head) belongs to the op-for-op path */
    if (!(s0 == 0u && 4u * t < head)) *reinterpret_cast<float4 *>(oc + s0 / 2u + 4u * t) = make_float4(yf[0], yf[1], yf[2], yf[3]);
    /* the thread's back group sits at run position TS - 8 - 8t: the back run is walked towards lower addresses */
    *reinterpret_cast<float4 *>(oc + (nc - s0 - TS) / 2u + (TS / 2u - 4u - 4u * t)) = make_float4(yb[0], yb[1], yb[2], yb[3]);
    } /* next tile of the run: its first writes to shared memory (carry_s aside) come after its own barriers */
}

#endif /* FMD_Q1_FUSED_CUH */
