/*
 * fmd_quality.h -- opt-in "quality" pipeline (FMD_FLAG_QUALITY): shared between the host
 * table builder (fmd_quality.c, plain C, long double) and the kernels (fmd_quality.cuh).
 *
 * What it computes is NOT what the reference computes: it is what the
 * project's north star describes -- recursive lowpass sections run as a chunked parallel scan
 * with an explicit state carry between threads, blocks and calls, an adjacent-sample
 * discriminator, and decimation to the -S rate fused into the output filter so that only the
 * decimated floats reach HBM.  The coefficient rows are still the reference's own design
 * (processFilterOption, src/matrix.c:25-80), designed at the input rate.
 *
 * A cascade of L sections {b0,b1,b2,1,a1,a2} in direct form II transposed is the linear system
 *     q[n+1] = A q[n] + B u[n],   w[n] = C q[n] + Dd u[n],        q in R^D, D = 2L
 * (q = (s1_0, s2_0, s1_1, s2_1, ...)).  A thread runs FMQ_T consecutive samples from a ZERO
 * state; the true state at the start of its span then follows from an affine scan whose
 * matrices are powers of A -- all of them built here on the host in long double.
 */
#ifndef FMD_QUALITY_H
#define FMD_QUALITY_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#ifndef FMQ_T
#define FMQ_T 64                          /* samples per thread (128 input bytes); measured on
                                             B200: 64 is 25 % faster than 32 (the CTA scan and the
                                             per-thread set-up are paid half as often)         */
#endif
#ifndef FMQ_THREADS
#define FMQ_THREADS 256                   /* threads per CTA                                  */
#endif
#define FMQ_T_BIG 128                     /* thread span the quality pipeline picks when the
                                             decimation factor is large (measured: +11 % at M = 50,
                                             -25 % at M = 1, where a tile's kept samples then take
                                             128 KiB of shared memory)                          */
#define FMQ_WARPS (FMQ_THREADS / 32)
#define FMQ_TILE (FMQ_T * FMQ_THREADS)    /* samples per CTA: 16384 (32 KiB of input)         */
#define FMQ_MAXL 3                        /* sections per cascade: degree <= 6                */
#define FMQ_MAXD (2 * FMQ_MAXL)
#define FMQ_SCAN_THREADS 1024             /* the tile-carry scan is one CTA of this size      */
#define FMQ_ZIR_EPS 1.0e-15L              /* rows of C A^k below this are not applied at all   */

/* per-filter constants that travel as kernel parameters (constant bank); matrices are packed
 * row-major D x D for the filter's own D */
typedef struct fmq_consts {
    int L, D;
    float sos[FMQ_MAXL][5];                 /* b0 b1 b2 a1 a2                                 */
    float lvl[5][FMQ_MAXD * FMQ_MAXD];      /* (A^T)^(2^j), j = 0..4: warp-level scan steps   */
    float warp[FMQ_MAXD * FMQ_MAXD];        /* A^(32 T): from one warp to the next            */
    float tile[FMQ_MAXD * FMQ_MAXD];        /* A^TILE: from one tile to the next              */
} fmq_consts;
/* The same struct describes the scan over TILES (fmq_scan_tiles_kernel): built with unit = TILE,
 * lvl[j] = A^(TILE 2^j), warp = A^(32 TILE), tile = A^TILE. */

/* per-filter tables in device global memory (floats), one allocation laid out in this order */
typedef struct fmq_table_sizes {
    unsigned lane;   /* [32][D*D]      (A^T)^l                                               */
    unsigned apow;   /* [T+1][D*D]     A^k, k = 0..T                                         */
    unsigned ck;     /* [T][D]         C A^k: zero-input response seen at offset k of a thread */
    unsigned ckn;    /* [TILE][D]      C A^k: the same for a whole tile                       */
    unsigned lane_tile; /* [32][D*D]   (A^TILE)^l: the lane table of the scan over tiles      */
} fmq_table_sizes;

typedef struct fmq_design {
    int L, D;
    long double A[FMQ_MAXD][FMQ_MAXD], B[FMQ_MAXD], C[FMQ_MAXD], Dd;
    float sos[FMQ_MAXL][5];
} fmq_design;

/* sections x {b0,b1,b2,a0,a1,a2} (fmd_design_sos rows) -> state-space form; 0 or FMD_EINVAL */
int fmq_design_from_sos(const float *sos6, unsigned sections, fmq_design *d);
/* out = A^k, packed row-major D x D */
void fmq_matpow(const fmq_design *d, unsigned long long k, long double *out);
/* unit = T for the scan over threads, T * FMQ_THREADS for the scan over tiles; `tile` = samples per tile */
void fmq_build_consts(const fmq_design *d, unsigned unit, unsigned tile, fmq_consts *c);
/* T = samples per thread (FMQ_T or FMQ_T_BIG); a tile is T * FMQ_THREADS samples */
void fmq_table_layout(const fmq_design *d, unsigned T, fmq_table_sizes *sz);
/* fills the tables (sizes from fmq_table_layout); returns the number of leading rows of ckn that
 * are not negligible (|C A^k| >= FMQ_ZIR_EPS somewhere in the row): the tile-level zero-input
 * response is applied to those offsets only */
unsigned fmq_build_tables(const fmq_design *d, unsigned T, float *lane, float *apow, float *ck, float *ckn,
                          float *lane_tile);
/* float copy of A^k for the launch-dependent transition of a ragged last tile */
void fmq_matpow_f(const fmq_design *d, unsigned long long k, float *out);

#ifdef __cplusplus
}
#endif
#endif
