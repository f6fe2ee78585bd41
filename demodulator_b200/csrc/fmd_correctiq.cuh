/*
 * fmd_correctiq.cuh -- the -q1 input conditioner (correctIq, /root/reference/src/matrix.c:120-140) on
 * the GPU.
 *
 * The reference walks a chunk from both ends: step k (complex k from the front, complex nc-1-k from
 * the back) subtracts the running offsets (one per lane) and then moves them,
 *     off += ((u_front - off) + (u_back - off)) * esr,         esr = (float)(50 / -S),
 * and the offsets live on from chunk to chunk.  In exact arithmetic that is the affine recurrence
 *     off' = (1 - 2 esr) off + esr (u_front + u_back)
 * over all steps of the stream, i.e. a scan with a constant scalar "matrix" -- the D = 1 case of the
 * machinery of fmd_quality.h:
 *   fmd_correctiq_kernel<false>   per tile (256 threads x FMQ1_T steps) the offsets' end value from zero
 *   fmq_scan_tiles_kernel<1, 2>   the offsets at the start of every tile; carry in from the previous
 *                                 call, carry out for the next
 *   fmd_correctiq_kernel<true>    every thread starts from its scanned offsets and walks its steps
 *                                 with the reference's own float operations in the reference's order,
 *                                 writing the conditioned samples as floats
 * The conditioned floats (one per input byte, same layout) then stream through fmd_fastf_kernel, the
 * float-input instantiation of the streaming pass (fmd_kernels.cuh); its exact CTAs -- chunk heads --
 * read the same floats through FmdShiftedLane (conditioner 1).
 *
 * This THREE-PASS form (22 B of HBM traffic per sample) is the fallback: designs with at most 3 input
 * taps and chunks that are a multiple of 8 KiB take the FUSED form of fmd_q1_fused.cuh, which
 * conditions and demodulates in one kernel (6 B per sample) on the same scan machinery.
 *
 * Not bit-exact: the scan associates the additions differently from the serial float32 loop, and the
 * recurrence contracts (factor 1 - 2 esr per step), so the offsets agree to about 1e-4 of ~127 and
 * the demodulated output to about 1e-6 -- inside the 1e-4 bar, like everything else on the fast side.
 */
#ifndef FMD_CORRECTIQ_CUH
#define FMD_CORRECTIQ_CUH

#include "fmd_quality.cuh"

/* Steps per thread.  The quality pipeline gives a thread 64 consecutive samples (its arithmetic is
 * heavy, the scan light); here a step is six instructions, so the memory pattern decides: with 8
 * steps per thread a warp reads 512 contiguous bytes from each end and writes 2 KiB of conditioned
 * floats to each end per instruction group, where 64 steps per thread left every lane walking its
 * own 512-byte stream (32 DRAM pages per store instruction: 65 GS/s for the whole -q1 pipeline). */
#define FMQ1_T 8
#define FMQ1_TILE (FMQ1_T * FMQ_THREADS) /* steps per CTA: 2048 */

struct FmdQ1Geom {
    unsigned long long nsteps;      /* steps in this call: nchunks * steps_per_chunk           */
    unsigned long long chunk_bytes; /* B: bytes per chunk (= floats per chunk of the output)   */
    unsigned steps_per_chunk;       /* B / 4                                                    */
    float esr, a;                   /* (float)(50 / -S), 1 - 2 esr                               */
};

template <bool APPLY>
__global__ void __launch_bounds__(FMQ_THREADS)
fmd_correctiq_kernel(const uint8_t *__restrict__ in, float *__restrict__ sbuf, const FmdQ1Geom g, const fmq_consts c,
                     const float *__restrict__ lane_tab, const float *__restrict__ apow,
                     const float *__restrict__ q_tile /* [ntiles][2], APPLY only */,
                     float *__restrict__ e_tile /* [ntiles][2], !APPLY only */) {
    __shared__ float tot[FMQ_WARPS * 2];
    const unsigned long long first = ((unsigned long long) blockIdx.x * FMQ_THREADS + threadIdx.x) * FMQ1_T;
    const unsigned cnt = first >= g.nsteps ? 0u : (unsigned) min((unsigned long long) FMQ1_T, g.nsteps - first);
    const bool last = cnt != 0u && (threadIdx.x == FMQ_THREADS - 1 || first + FMQ1_T >= g.nsteps);
    /* FMQ1_T divides steps_per_chunk (the host checks): a thread's steps lie in one chunk */
    const unsigned long long chunk = first / g.steps_per_chunk;
    const unsigned k0 = (unsigned) (first - chunk * g.steps_per_chunk);
    const unsigned nc = 2u * g.steps_per_chunk;
    const uint8_t *__restrict__ cin = in + chunk * g.chunk_bytes;
    const unsigned short *__restrict__ front = reinterpret_cast<const unsigned short *>(cin) + k0;
    const unsigned short *__restrict__ back = reinterpret_cast<const unsigned short *>(cin) + (nc - 1u - k0);

    /* A thread's eight steps are one 16-byte load of the front run (complex k0 .. k0+7) and one of
     * the back run (complex nc-8-k0 .. +7, i.e. step j is its sample 7-j). */
#define FMQ1_W(v, i) ((i) < 2 ? (v).x : (i) < 4 ? (v).y : (i) < 6 ? (v).z : (v).w)
#define FMQ1_I(v, i) ((FMQ1_W(v, i) >> (16 * ((i) & 1))) & 0xffu)
#define FMQ1_Q(v, i) ((FMQ1_W(v, i) >> (16 * ((i) & 1) + 8)) & 0xffu)
    const uint4 *__restrict__ front4 = reinterpret_cast<const uint4 *>(front);
    const uint4 *__restrict__ back4 = reinterpret_cast<const uint4 *>(back - 7);
    /* end value of the offsets after this thread's steps, from zero */
    float v[2][1] = {{0.0f}, {0.0f}};
    static_assert(FMQ1_T == 8, "one 16-byte load per end and thread");
    uint4 f = make_uint4(0u, 0u, 0u, 0u), b = f;
    if (cnt == FMQ1_T) {
        f = __ldg(front4);
        b = __ldg(back4);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            v[0][0] = fmaf(g.a, v[0][0], g.esr * (float) (FMQ1_I(f, j) + FMQ1_I(b, 7 - j)));
            v[1][0] = fmaf(g.a, v[1][0], g.esr * (float) (FMQ1_Q(f, j) + FMQ1_Q(b, 7 - j)));
        }
    } else {
#pragma unroll 4
        for (unsigned k = 0; k < cnt; ++k) {
            const unsigned f = front[k], b = back[-(int) k];
            v[0][0] = fmaf(g.a, v[0][0], g.esr * (float) ((f & 0xffu) + (b & 0xffu)));
            v[1][0] = fmaf(g.a, v[1][0], g.esr * (float) ((f >> 8) + (b >> 8)));
        }
    }
    const float e0 = v[0][0], e1 = v[1][0];
    float carry[2][1] = {{0.0f}, {0.0f}};
    if (APPLY) {
        carry[0][0] = __ldg(q_tile + 2 * (size_t) blockIdx.x);
        carry[1][0] = __ldg(q_tile + 2 * (size_t) blockIdx.x + 1);
    }
    fmq_block_scan<1, 2>(v, c, lane_tab, carry, tot); /* -> offsets at the start of this thread's steps */
    if (!APPLY) {
        if (last) {
            const float p = __ldg(apow + cnt); /* a^cnt */
            e_tile[2 * (size_t) blockIdx.x] = fmaf(p, v[0][0], e0);
            e_tile[2 * (size_t) blockIdx.x + 1] = fmaf(p, v[1][0], e1);
        }
        return;
    }
    /* the reference's loop body, operation for operation (matrix.c:129-138) */
    float off0 = v[0][0], off1 = v[1][0];
    /* FMQ1_T divides the steps of a chunk, so cnt is 0 or FMQ1_T: steps go in pairs, 16 bytes per store */
    float4 *__restrict__ sf = reinterpret_cast<float4 *>(reinterpret_cast<float2 *>(sbuf + chunk * g.chunk_bytes) + k0);
    float4 *__restrict__ sb =
        reinterpret_cast<float4 *>(reinterpret_cast<float2 *>(sbuf + chunk * g.chunk_bytes) + (nc - 2u - k0));
    if (cnt == FMQ1_T) {
        {
#pragma unroll
            for (int j = 0; j < 8; j += 2) {
                float fi[2], fq[2], bi[2], bq[2];
#pragma unroll
                for (int t = 0; t < 2; ++t) {
                    fi[t] = __fsub_rn((float) FMQ1_I(f, j + t), off0);
                    bi[t] = __fsub_rn((float) FMQ1_I(b, 7 - (j + t)), off0);
                    fq[t] = __fsub_rn((float) FMQ1_Q(f, j + t), off1);
                    bq[t] = __fsub_rn((float) FMQ1_Q(b, 7 - (j + t)), off1);
                    off0 = __fadd_rn(off0, __fmul_rn(__fadd_rn(fi[t], bi[t]), g.esr));
                    off1 = __fadd_rn(off1, __fmul_rn(__fadd_rn(fq[t], bq[t]), g.esr));
                }
                const unsigned k = (unsigned) j;
                sf[k >> 1] = make_float4(fi[0], fq[0], fi[1], fq[1]);
                sb[-(int) (k >> 1)] = make_float4(bi[1], bq[1], bi[0], bq[0]); /* complex nc-2-k0-k, then nc-1-k0-k */
            }
        }
    }
#undef FMQ1_W
#undef FMQ1_I
#undef FMQ1_Q
}

#endif /* FMD_CORRECTIQ_CUH */
