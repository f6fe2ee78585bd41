/*
 * fmd_fast_inst.cu -- the instantiations of fmd_fast_kernel<NTI, NTO, AT, COND> (fmd_kernels.cuh),
 * compiled once per group (-DFMD_INST_GROUP=0..3) so that the groups build in parallel:
 *   0: no conditioner, 0..1 input taps      2: no conditioner, 4..5 input taps
 *   1: no conditioner, 2..3 input taps      3: -q3 (normalizeInput), 0..3 input taps
 *   4: float input (-q1: samples conditioned by fmd_correctiq.cuh), 0..3 input taps
 * (-q2 needs no instantiation of its own: its highpass folds into the input stencil.)
 */
#include "fmd_ctx.h"
#include "fmd_kernels.cuh"

#ifndef FMD_INST_GROUP
#error "compile with -DFMD_INST_GROUP=0..4 (9: the headline instantiation only)"
#endif

template <int NTI, int AT, int COND>
static fmd_fast_fn pick_out(unsigned nto) {
    switch (nto) {
        case 1: return fmd_fast_kernel<NTI, 1, AT, COND>;
        case 2: return fmd_fast_kernel<NTI, 2, AT, COND>;
        case 3: return fmd_fast_kernel<NTI, 3, AT, COND>;
        case 4: return fmd_fast_kernel<NTI, 4, AT, COND>;
        case 5: return fmd_fast_kernel<NTI, 5, AT, COND>;
        default: return nullptr;
    }
}

template <int NTI, int COND>
static fmd_fast_fn pick_at(unsigned nto, unsigned a8) {
    return a8 ? pick_out<NTI, 8, COND>(nto) : pick_out<NTI, 6, COND>(nto);
}

#if FMD_INST_GROUP == 9
/* the headline instantiation alone (cfg2: 2 input taps, 2 output taps, 6-term atan): what the build-option
 * tests and SASS inspections compile when they do not need the other 139 */
template __global__ void fmd_fast_kernel<2, 2, 6, 0>(const uint8_t *, float *, const FmdGeom, const FmdStencil, const FmdSections *);
#elif FMD_INST_GROUP == 0
fmd_fast_fn fmd_pick_n01(unsigned nti, unsigned nto, unsigned a8) {
    return nti == 0 ? pick_at<0, 0>(nto, a8) : nti == 1 ? pick_at<1, 0>(nto, a8) : nullptr;
}
#elif FMD_INST_GROUP == 1
fmd_fast_fn fmd_pick_n23(unsigned nti, unsigned nto, unsigned a8) {
    return nti == 2 ? pick_at<2, 0>(nto, a8) : nti == 3 ? pick_at<3, 0>(nto, a8) : nullptr;
}
#elif FMD_INST_GROUP == 2
fmd_fast_fn fmd_pick_n45(unsigned nti, unsigned nto, unsigned a8) {
    return nti == 4 ? pick_at<4, 0>(nto, a8) : nti == 5 ? pick_at<5, 0>(nto, a8) : nullptr;
}
#elif FMD_INST_GROUP == 4
template <int NTI, int AT>
static fmd_fast_fn pick_out_f(unsigned nto) {
    /* same launch signature up to the element type of the input pointer */
    switch (nto) {
        case 1: return reinterpret_cast<fmd_fast_fn>(fmd_fastf_kernel<NTI, 1, AT>);
        case 2: return reinterpret_cast<fmd_fast_fn>(fmd_fastf_kernel<NTI, 2, AT>);
        case 3: return reinterpret_cast<fmd_fast_fn>(fmd_fastf_kernel<NTI, 3, AT>);
        case 4: return reinterpret_cast<fmd_fast_fn>(fmd_fastf_kernel<NTI, 4, AT>);
        case 5: return reinterpret_cast<fmd_fast_fn>(fmd_fastf_kernel<NTI, 5, AT>);
        default: return nullptr;
    }
}
template <int NTI>
static fmd_fast_fn pick_at_f(unsigned nto, unsigned a8) {
    return a8 ? pick_out_f<NTI, 8>(nto) : pick_out_f<NTI, 6>(nto);
}
fmd_fast_fn fmd_pick_float(unsigned nti, unsigned nto, unsigned a8) {
    switch (nti) {
        case 0: return pick_at_f<0>(nto, a8);
        case 1: return pick_at_f<1>(nto, a8);
        case 2: return pick_at_f<2>(nto, a8);
        case 3: return pick_at_f<3>(nto, a8);
        default: return nullptr;
    }
}
#else
fmd_fast_fn fmd_pick_q3(unsigned nti, unsigned nto, unsigned a8) {
    switch (nti) {
        case 0: return pick_at<0, 3>(nto, a8);
        case 1: return pick_at<1, 3>(nto, a8);
        case 2: return pick_at<2, 3>(nto, a8);
        case 3: return pick_at<3, 3>(nto, a8);
        default: return nullptr;
    }
}
#endif
