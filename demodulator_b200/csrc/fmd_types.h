/* fmd_types.h -- plain structs shared by the host code and the kernels of the
 * reference-compatible path (fmd_kernels.cuh). */
#ifndef FMD_TYPES_H
#define FMD_TYPES_H

#include "fmdemod.h"

#ifndef FMD_BLOCK
#define FMD_BLOCK 256               /* threads per CTA: 8 warps, one segment each */
#endif
#define FMD_WARPS (FMD_BLOCK / 32)
#define FMD_PASS_OUT 128u           /* outputs per warp per pass (32 lanes x 4)   */

/* folded filters (host: fmd_fold_taps); tin[] is divided by tin[0], and so is dcq */
struct FmdStencil {
    float tin[FMD_MAX_TAPS];
    float dci, dcq;
    float tout[FMD_MAX_TAPS];
    float dco;
};

struct FmdGeom {
    unsigned long long chunk_stride; /* bytes between chunk starts                        */
    unsigned chunk_out;              /* outputs per chunk (edge rule: zero beyond)        */
    unsigned fast_out;               /* outputs per chunk the fast kernel covers (x128)   */
    unsigned iters_per_seg;          /* passes per warp segment                           */
    unsigned segs_per_chunk;
    unsigned nchunks;
    unsigned head;                   /* outputs [0, head) of every chunk: exact path (x4)   */
    unsigned exact_blocks;           /* leading CTAs of the fast kernel that run that path  */
};

struct FmdSections {
    float in[FMD_MAX_SECTIONS][6];
    float out[FMD_MAX_SECTIONS][6];
    int lin, lout;
    int conditioner; /* consumerArgs.mode bits 2-3 (-q): 0 shiftOrigin, 2 highpassDc, 3 normalizeInput */
    float hp[2][6];  /* -q2: the degree-3 Butterworth highpass at 1 Hz (matrix.c:151) */
};

#endif /* FMD_TYPES_H */
