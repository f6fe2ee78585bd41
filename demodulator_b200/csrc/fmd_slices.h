/*
 * fmd_slices.h -- how fmd_demod_host / fmd_copy_host cut one call into slices for the ring
 * (plain C, no CUDA: tests/test_host_slices.py compiles this header alone).
 *
 * Slices are whole chunks (multiples of B) except the call's last one, never larger than a ring slot.
 * Graded (`ramp`): they grow 4, 8, 16 ... MiB from the start of the call and halve towards its end, so
 * that the two stretches in which only one direction of the link is busy -- the first copy-in, the
 * last copy-out -- stay short, while the slices in between are as large as the slot
 * (profiles/r02_e2e_probe.txt).
 */
#ifndef FMD_SLICES_H
#define FMD_SLICES_H

#include <stddef.h>

/* the smallest graded slice: 4 MiB in whole chunks, at least one chunk, at most a slot */
static inline size_t fmd_slice_first(size_t B, size_t slot_bytes) {
    size_t first = ((size_t) 4 << 20) / B * B;
    if (first == 0) first = B;
    if (first > slot_bytes) first = slot_bytes;
    return first;
}

/* length of slice number n of a call that still has `remaining` (> 0) bytes to submit */
static inline size_t fmd_slice_len(unsigned n, size_t remaining, size_t B, size_t slot_bytes, int ramp) {
    size_t cap = slot_bytes;
    if (ramp) {
        const size_t first = fmd_slice_first(B, slot_bytes);
        const size_t up = first << (n < 16u ? n : 16u);
        if (up < cap) cap = up;
        if (remaining > first) {
            size_t half = remaining / 2 / B * B;
            if (half < first) half = first;
            if (half < cap) cap = half;
        }
    }
    return remaining < cap ? remaining : cap;
}

#endif
