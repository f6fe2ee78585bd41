/*
 * tools/experiments/fmd_fast2.cuh -- NOT part of the product build.
 *
 * Measured and dropped: a streaming kernel in which a lane owns 32 contiguous input bytes per
 * trip (8 outputs) instead of 16, so that only one halo word and one neighbour value per 8
 * outputs cross lanes.  Compiles against demodulator_b200/csrc/fmd_kernels.cuh (include it
 * after that header); 64 registers, 4 bytes of spill for <2,2,6>.
 *
 * [B200] wired under -DFMD_EXPERIMENT_FAST2 / FMD_FAST2=1 (commit "experiment: 8-outputs-per-lane
 * streaming kernel ..."; the host counts trips of 256 outputs, iters_per_seg halved): all 101
 * parity tests green, but 832 us vs 742 us (cfg2) and 725 us vs 646 us (cfg1) per 1 Gi samples
 * at the best segment length -- every warp-level load and store touches 32 half-used sectors
 * (lanes are 32 bytes apart, accesses are 16 bytes), which costs more than the saved shuffles.
 * tools/gpu_fast2.sh is the session that measured it.
 */
/* ------------------------------------------------ double-pass streaming kernel (experiment) */

/*
 * A lane owns 32 contiguous input bytes per trip (16 complex samples, 8 outputs) instead of 16:
 * the second half's first samples and first output are the first half's look-ahead, so only one
 * halo word and one neighbour value per 8 outputs cross lanes.
 */
template <int NTI>
struct FmdHalo {
    static constexpr int HC = NTI > 1 ? NTI - 1 : 0;
    float i[HC > 0 ? HC : 1], q[HC > 0 ? HC : 1];
};

/* 4 raw words (8 complex samples) + the HC samples that follow them -> 4 discriminator outputs;
 * `own` receives this half's first HC samples (the look-ahead of the half before it) */
template <int NTI, int AT, bool LITE>
__device__ __forceinline__ void fmd_half(const uint4 &cur, const FmdHalo<NTI> &next, float (&x)[4],
                                         FmdHalo<NTI> &own, const FmdStencil &st) {
    constexpr int HC = FmdHalo<NTI>::HC;
    constexpr int NK = 4 + HC;
    uint32_t w[4];
    w[0] = fmd_bias_bytes(cur.x);
    w[1] = fmd_bias_bytes(cur.y);
    w[2] = fmd_bias_bytes(cur.z);
    w[3] = fmd_bias_bytes(cur.w);
    fmd_f2 vi[NK], vq[NK];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        if (k & 1) {
            vi[k] = fmd_bytes_to_f2<2>(w[k >> 1], w[(k >> 1) + 2]);
            vq[k] = fmd_bytes_to_f2<3>(w[k >> 1], w[(k >> 1) + 2]);
        } else {
            vi[k] = fmd_bytes_to_f2<0>(w[k >> 1], w[(k >> 1) + 2]);
            vq[k] = fmd_bytes_to_f2<1>(w[k >> 1], w[(k >> 1) + 2]);
        }
    }
#pragma unroll
    for (int t = 0; t < HC; ++t) { /* (I_{4+t}, I_{8+t}) = (high half of vi[t], look-ahead sample t) */
        float lo, hi;
        fmd_upk(vi[t], lo, hi);
        own.i[t] = lo;
        vi[4 + t] = fmd_pk(hi, next.i[t]);
        fmd_upk(vq[t], lo, hi);
        own.q[t] = lo;
        vq[4 + t] = fmd_pk(hi, next.q[t]);
    }
#pragma unroll
    for (int j = 0; j < (LITE ? 1 : 2); ++j) {
        fmd_f2 a, b, c, d;
        if (NTI == 0) {
            a = vi[2 * j]; b = vq[2 * j]; c = vi[2 * j + 1]; d = vq[2 * j + 1];
        } else {
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
        const fmd_f2 re = fmd_fma2(a, c, fmd_mul2(b, d));
        const fmd_f2 im = fmd_fma2(fmd_neg2(a), d, fmd_mul2(b, c));
        fmd_atan2_x2<AT>(im, re, x[j], x[j + 2]);
    }
    if (LITE) x[1] = x[3] = 0.0f;
}

/* the HC samples that start at raw words (h0, h1), as floats */
template <int NTI>
__device__ __forceinline__ void fmd_halo_from_raw(uint32_t h0, uint32_t h1, FmdHalo<NTI> &o) {
    constexpr int HC = FmdHalo<NTI>::HC;
    const uint32_t b0 = fmd_bias_bytes(h0), b1 = fmd_bias_bytes(h1);
    if (HC >= 1) { o.i[0] = fmd_byte_to_float<0>(b0); o.q[0] = fmd_byte_to_float<1>(b0); }
    if (HC >= 2) { o.i[1 % (HC > 0 ? HC : 1)] = fmd_byte_to_float<2>(b0); o.q[1 % (HC > 0 ? HC : 1)] = fmd_byte_to_float<3>(b0); }
    if (HC >= 3) { o.i[2 % (HC > 0 ? HC : 1)] = fmd_byte_to_float<0>(b1); o.q[2 % (HC > 0 ? HC : 1)] = fmd_byte_to_float<1>(b1); }
    if (HC >= 4) { o.i[3 % (HC > 0 ? HC : 1)] = fmd_byte_to_float<2>(b1); o.q[3 % (HC > 0 ? HC : 1)] = fmd_byte_to_float<3>(b1); }
}

template <int NTO>
__device__ __forceinline__ float4 fmd_out_stage(const float (&xn)[8], const FmdStencil &st) {
    float4 y;
    float *yy = reinterpret_cast<float *>(&y);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        float acc = st.dco;
#pragma unroll
        for (int t = 0; t < NTO; ++t) acc = fmaf(st.tout[t], xn[j + t], acc);
        yy[j] = acc;
    }
    return y;
}

#define FMD_DP_OUT 256u /* outputs per warp per double pass */

template <int NTI, int NTO, int AT>
__global__ void __launch_bounds__(FMD_BLOCK, FMD_MINB(NTI))
fmd_fast2_kernel(const uint8_t *__restrict__ in, float *__restrict__ out, const FmdGeom g, const FmdStencil st,
                 const FmdSections *__restrict__ sec) {
    if (blockIdx.x < g.exact_blocks) {
        fmd_exact_thread(in, out, g, sec, (unsigned long long) blockIdx.x * FMD_BLOCK + threadIdx.x);
        return;
    }
    constexpr int S = 4;
    constexpr int HC = FmdHalo<NTI>::HC;
    constexpr bool HALO = HC > 0;
    constexpr int NB = NTO - 1;
    constexpr int PENDING = HALO ? S - 2 : S - 1;
    constexpr uint32_t SLOT = FMD_BLOCK * 32u;
    __shared__ uint4 ring[S * FMD_BLOCK * 2];

    const unsigned lane = threadIdx.x & 31u;
    const unsigned nxt_lane = (lane + 1u) & 31u;
    const unsigned seg = (blockIdx.x - g.exact_blocks) * FMD_WARPS + (threadIdx.x >> 5);
    if (seg >= g.nchunks * g.segs_per_chunk) return;
    const unsigned chunk = seg / g.segs_per_chunk;
    const unsigned sic = seg - chunk * g.segs_per_chunk;
    const unsigned it0 = sic * g.iters_per_seg;
    const unsigned iters_chunk = g.fast_out >> 8;
    const unsigned n_it = min(g.iters_per_seg, iters_chunk - it0);
    const bool end_zero = (it0 + n_it == iters_chunk) && (g.fast_out == g.chunk_out);
    const unsigned n_ld = end_zero ? n_it : n_it + 1u;
    const unsigned k0 = it0 * FMD_DP_OUT + lane * 8u;
    const bool head_lane = it0 == 0 && k0 < g.head; /* head is a multiple of 8: both halves of the lane */

    const uint8_t *__restrict__ src = in + (unsigned long long) chunk * g.chunk_stride + 4ull * k0;
    float *__restrict__ dst = out + (unsigned long long) chunk * g.chunk_out + k0;
    const uint32_t sbase = (uint32_t) __cvta_generic_to_shared(&ring[2 * threadIdx.x]);

#pragma unroll
    for (int q = 0; q < S - 1; ++q) {
        if ((unsigned) q < n_ld) {
            fmd_cp_async16(sbase + q * SLOT, src + 1024 * q);
            fmd_cp_async16(sbase + q * SLOT + 16, src + 1024 * q + 16);
        }
        fmd_cp_commit();
    }

    float pb[4] = {0.f, 0.f, 0.f, 0.f};           /* second-half outputs of the previous trip */
    float pa[NB > 0 ? NB : 1] = {0.f};              /* first outputs of the previous trip       */

    /* trip p: fetch, discriminate both halves (second half first: its first samples and first
     * output are the first half's look-ahead), store this trip's first half and the previous
     * trip's second half */
#define FMD_TRIP(SLOTR, SLOTN, SLOTW, P, STORE_A, STORE_B)                                            \
    {                                                                                                 \
        if ((P) + S - 1 < n_ld) {                                                                     \
            fmd_cp_async16(sbase + (SLOTW) * SLOT, src + 1024 * (S - 1));                             \
            fmd_cp_async16(sbase + (SLOTW) * SLOT + 16, src + 1024 * (S - 1) + 16);                   \
        }                                                                                             \
        fmd_cp_commit();                                                                              \
        fmd_cp_wait<PENDING>();                                                                       \
        const uint4 ra = fmd_lds128(sbase + (SLOTR) * SLOT);                                          \
        const uint4 rb = fmd_lds128(sbase + (SLOTR) * SLOT + 16);                                     \
        FmdHalo<NTI> hb, ob, oa;                                                                      \
        if (HALO) {                                                                                   \
            uint2 nx = make_uint2(FMD_FILL, FMD_FILL);                                                \
            if ((P) + 1 < n_ld) nx = fmd_lds64(sbase + (SLOTN) * SLOT);                               \
            const uint32_t h0 = __shfl_sync(0xffffffffu, lane == 0 ? nx.x : ra.x, nxt_lane);          \
            const uint32_t h1 = HC > 2 ? __shfl_sync(0xffffffffu, lane == 0 ? nx.y : ra.y, nxt_lane) : 0u; \
            fmd_halo_from_raw<NTI>(h0, h1, hb);                                                       \
        }                                                                                             \
        float xa[4], xb[4];                                                                           \
        fmd_half<NTI, AT, false>(rb, hb, xb, ob, st);                                                 \
        fmd_half<NTI, AT, false>(ra, ob, xa, oa, st);                                                 \
        float xn[8];                                                                                  \
        xn[0] = xa[0]; xn[1] = xa[1]; xn[2] = xa[2]; xn[3] = xa[3];                                   \
        xn[4] = xb[0]; xn[5] = xb[1]; xn[6] = xb[2]; xn[7] = xb[3];                                   \
        const float4 ya = fmd_out_stage<NTO>(xn, st);                                                 \
        if (STORE_A) *reinterpret_cast<float4 *>(dst) = ya;                                           \
        xn[0] = pb[0]; xn[1] = pb[1]; xn[2] = pb[2]; xn[3] = pb[3];                                   \
        _Pragma("unroll") for (int t = 0; t < NB; ++t)                                                \
            xn[4 + t] = __shfl_sync(0xffffffffu, lane == 0 ? xa[t] : pa[t], nxt_lane);                \
        const float4 yb = fmd_out_stage<NTO>(xn, st);                                                 \
        if (STORE_B) *reinterpret_cast<float4 *>(dst - FMD_DP_OUT + 4) = yb;                          \
        pb[0] = xb[0]; pb[1] = xb[1]; pb[2] = xb[2]; pb[3] = xb[3];                                   \
        _Pragma("unroll") for (int t = 0; t < NB; ++t) pa[t] = xa[t];                                 \
        src += 1024;                                                                                  \
        dst += FMD_DP_OUT;                                                                            \
    }

    unsigned p = 0;
    for (; p + S <= n_it; p += S) {
        const bool later = p != 0;
        FMD_TRIP(0, 1, 3, p, later || !head_lane, later)
        FMD_TRIP(1, 2, 0, p + 1, true, later || !head_lane)
        FMD_TRIP(2, 3, 1, p + 2, true, true)
        FMD_TRIP(3, 0, 2, p + 3, true, true)
    }
    for (; p < n_it; ++p) { /* n_it not a multiple of the ring depth: slots by arithmetic */
        const unsigned u = p & (S - 1);
        FMD_TRIP(u, (u + 1) & (S - 1), (u + S - 1) & (S - 1), p, p != 0 || !head_lane, p >= 2 || (p == 1 && !head_lane))
    }
#undef FMD_TRIP
    /* the look-ahead trip: only lane 0's first outputs of it matter (to lane 31's last ones) */
    {
        const unsigned u = p & (S - 1);
        fmd_cp_commit();
        fmd_cp_wait<PENDING>();
        float xa[4] = {0.f, 0.f, 0.f, 0.f};
        if (!end_zero) {
            const uint4 ra = fmd_lds128(sbase + u * SLOT);
            FmdHalo<NTI> hb, oa;
            if (HALO) {
                const uint2 rbx = fmd_lds64(sbase + u * SLOT + 16);
                fmd_halo_from_raw<NTI>(rbx.x, rbx.y, hb);
            }
            fmd_half<NTI, AT, (NTO <= 2)>(ra, hb, xa, oa, st);
        }
        float xn[8];
        xn[0] = pb[0]; xn[1] = pb[1]; xn[2] = pb[2]; xn[3] = pb[3];
        xn[4] = xn[5] = xn[6] = xn[7] = 0.f;
#pragma unroll
        for (int t = 0; t < NB; ++t) xn[4 + t] = __shfl_sync(0xffffffffu, lane == 0 ? xa[t] : pa[t], nxt_lane);
        const float4 yb = fmd_out_stage<NTO>(xn, st);
        if (p >= 2 || (p == 1 && !head_lane)) *reinterpret_cast<float4 *>(dst - FMD_DP_OUT + 4) = yb;
    }
    fmd_cp_wait<0>();
}

