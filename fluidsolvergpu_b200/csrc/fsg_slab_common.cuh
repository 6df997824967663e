// fsg_slab_common.cuh — helpers shared by the two slab pipelines (fsg_slab.cu: ghosts travel through the sort; fsg_slab2.cu:
// sorted ghosts in their own zones).
#pragma once
#include "fsg_device.cuh"

#include <stdio.h>

#define CUS(ctx, call)                                                                                  \
    do {                                                                                                \
        cudaError_t e_ = (call);                                                                        \
        if (e_ != cudaSuccess) {                                                                        \
            char b_[512];                                                                               \
            snprintf(b_, sizeof b_, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
            (ctx)->err = b_;                                                                            \
            return e_ == cudaErrorMemoryAllocation ? FSG_E_NOMEM : FSG_E_CUDA;                          \
        }                                                                                               \
    } while (0)

// After the first step the particles are in bin-sorted order, and only slots within two layers of a face that has
// a neighbour can have become migrants or ghosts (a particle moves less than one bin per step): the sorted slots
// [0, region[0]) and [region[1], n_keep), found by k_reorder (n_keep: the slots in use; dead slots sort last).  The pack
// kernels are launched over THAT index space only — compact index t -> slot t (head) or region[1] + t - region[0] (tail) —
// with a fixed-size grid that strides over it: nothing is launched, read or written for the slots in between, which is
// also what lets overlap mode pack while the interior particles are still being updated.
struct SlabRegion {
    int64_t r0, r1, total;            // head = [0, r0), tail = [r1, r1 + total - r0)
};
__device__ __forceinline__ SlabRegion slab_region(const int *__restrict__ region, const int *__restrict__ n_keep, int64_t n)
{
    SlabRegion R;
    if (!region) { R.r0 = n; R.r1 = n; R.total = n; return R; }      // before the first step: every slot
    const int64_t keep = min((int64_t)*n_keep, n);
    R.r0 = min((int64_t)region[0], keep);
    R.r1 = max(min((int64_t)region[1], keep), R.r0);
    R.total = R.r0 + (keep - R.r1);
    return R;
}
__device__ __forceinline__ int64_t slab_slot(const SlabRegion &R, int64_t t) { return t < R.r0 ? t : R.r1 + (t - R.r0); }

// fsg_slab.cu
int fsg_slab_ensure_counts(fsg_ctx *c, int64_t nw);        // per-warp count / offset arrays + scan workspace + diagnostics
long long *fsg_slab_diag(fsg_ctx *c);                      // [0..3] sent (migrants / ghosts left, right), [4..7] received
// one-thread device-side wait until both tails hold a stamp >= expected (bounded by wall-clock time; raises flag 4)
cudaError_t fsg_launch_slab_wait(fsg_ctx *c, const long long *tail_left, const long long *tail_right, long long expected, cudaStream_t s);
