// fsg_unidyn.cuh — declarations shared by the unidyn kernels (fsg_unidyn.cu: pure-fluid pair sums, update, AoS, stage API;
// fsg_unidyn_mixed.cu: the two-pass mixed-phase / granular pair sums).
#pragma once
#include "fsg_device.cuh"

#define UNI_WARPS 2
#define UNI_TILE 1024                   // the reference's threads per block = neighbour particles per bin (solver-unidyn.cu:363)

struct UniWarpSmem {
    float4 sp[UNI_TILE];                // x, y, z, +-dens (sign = boundary)
    float4 sv[UNI_TILE];                // vx, vy, vz, press / dens^2
    float sf[UNI_TILE];                 // fluid
    unsigned short q[UNI_TILE];
    unsigned char tag[UNI_TILE];        // neighbour slot 0..26 of the candidate's bin
};
#define UNI_SMEM (sizeof(UniWarpSmem) * UNI_WARPS)

struct UniArgs {
    FsgDev d;
    int n;
    const int *start, *end, *binlist, *nocc;
    int *work;
    FsgState A;
    float4 *sums, *sums2;               // (newdens, newdelpress xyz), (diffusion xyz, delfluid)
    unsigned long long *stats;
    // mixed-phase / granular scenes (fsg_unidyn_mixed.cu): per sorted slot
    float *mixA;                        // [n][UNI_MIXA]: solid drift xyz, fluid drift xyz, vel_grad[3][3], stress_accel xyz   (pass A)
    float *mixB;                        // [n][UNI_MIXB]: mixture_accel xyz, delsolid, delfluid                                   (pass B)
};
#define UNI_MIXA 18
#define UNI_MIXB 8
#define UNI_STRESS 18                   // FsgState::stress row: stress_tensor[3][3] then stress_rate[3][3] (FluidGPU-unidyn.cuh)

// fsg_unidyn_mixed.cu: pass A (pass == 0) / pass B (pass == 1) of the mixed-phase pair sums; a.work must be zero
cudaError_t fsg_launch_unidyn_mixed(const UniArgs &a, int pass, int sm_count, cudaStream_t s);

// octant of a particle inside its bin, FluidGPU-unidyn.cu:182-184
__device__ __forceinline__ int uni_subindex(const FsgDev &d, float x, float y, float z)
{
    float fx = x - d.origin, fy = y - d.origin, fz = z - d.origin;
    const double cs = d.cellsize;
    int sx = (int)((double)fx / cs) == (int)(((double)fx + cs / 2) / cs);
    int sy = (int)((double)fy / cs) == (int)(((double)fy + cs / 2) / cs);
    int sz = (int)((double)fz / cs) == (int)(((double)fz + cs / 2) / cs);
    return 1 - sx + 2 - 2 * sy + 4 * sz;
}
// the 8 neighbour slots (of the 27) mykernel3 visits for an octant, :579-583
__device__ __forceinline__ unsigned uni_octant_mask(int oct)
{
    const int ax = (oct & 1) ? 1 : -1, ay = (oct & 2) ? 1 : -1, az = (oct & 4) ? -1 : 1;
    unsigned m = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) {
        int a = (k & 1) ? ax : 0, b = (k & 2) ? ay : 0, c = (k & 4) ? az : 0;
        m |= 1u << ((a + 1) * 9 + (b + 1) * 3 + (c + 1));
    }
    return m;
}

