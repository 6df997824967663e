// fsg_device.cuh — device helpers shared by the kernels of the base particle step.
#pragma once
#include "fsg_internal.cuh"

#include <math.h>

#define FULL 0xffffffffu

// Bin id of a position — the expression of FluidGPU.cu:419 / solver.cu:119:
//   int((x - XMIN)/CELLSIZE)*G*G + int((y - YMIN)/CELLSIZE)*G + int((z - ZMIN)/CELLSIZE)
// (float subtraction, double division, truncation).  Where the reference's linear id would leave
// [0, numcells) — it then writes start[]/end[] out of bounds, FluidGPU.cu:110 — the particle is
// parked with id == numcells: it sorts last and is never touched again.
__device__ __forceinline__ int bin_id(const FsgDev &d, float x, float y, float z)
{
    float fx = x - d.origin, fy = y - d.origin, fz = z - d.origin;
    double qx = (double)fx / d.cellsize, qy = (double)fy / d.cellsize, qz = (double)fz / d.cellsize;
    if (!(fabs(qx) < 1e6 && fabs(qy) < 1e6 && fabs(qz) < 1e6)) return d.numcells;
    long long l = (long long)(int)qx * d.G2 + (long long)(int)qy * d.G + (int)qz;
    return (l < 0 || l >= d.numcells) ? d.numcells : (int)l;
}

// Squared distance exactly as Particle::distance forms it (FluidGPU.cuh:193-195): three rounded
// squares added left to right, no FMA contraction — the in/out decision at ds == 2h must not depend
// on how the compiler fuses the sum.
__device__ __forceinline__ float dist2(float rx, float ry, float rz)
{
    return __fadd_rn(__fadd_rn(__fmul_rn(rx, rx), __fmul_rn(ry, ry)), __fmul_rn(rz, rz));
}

// Particle::update (FluidGPU.cuh:270-304) + the tail of mykernel2 (FluidGPU.cu:419-425) for one
// particle.  Follows the reference's promotions expression by expression — it runs once per
// particle, so the double arithmetic is free next to the pair loop.
template <bool WITH_KEY = true>
__device__ __forceinline__ void particle_update(const FsgDev &d, float4 &pd, float4 &vp, float4 &af, float4 &dpi,
                                                float newdens, float ndx, float ndy, float ndz, int &key)
{
    bool bnd = pd.w < 0.f;
    // set_dens  cuh:165-167
    float dens = (float)((double)(newdens + d.w0) / 23.0 * (double)(1 + (float)bnd * 1.5) + 9250);
    // calculate_pressure  cuh:256-257
    float press = (float)((double)(1000 * powf((float)d.sound, 0.f) * 9550) / 7.0 * (double)(powf(dens / 9550, 7.f) - 1));
    dpi.x = ndx;   // set_delpress  cuh:276
    dpi.y = ndy;
    dpi.z = ndz;
    if (!bnd) {
        const double DT = d.dt;
        float x = (float)((double)pd.x + DT * (double)vp.x);   // cuh:286-288 (DIFF == 0)
        float y = (float)((double)pd.y + DT * (double)vp.y);
        float z = (float)((double)pd.z + DT * (double)vp.z);
        double tx = ((double)vp.x + DT * (double)af.x + DT * 0.0);          // cuh:290-295
        float vx = (float)(tx - (tx > 0) * 0.003 + (tx < 0) * 0.003);
        vx *= ((double)fabsf(vx) > 0.003);
        double ty = ((double)vp.y + DT * (double)af.y + DT * 0.0);
        float vy = (float)(ty - (ty > 0) * 0.003 + (ty < 0) * 0.003);
        vy *= ((double)fabsf(vy) > 0.003);
        float vz = (float)((double)vp.z + DT * (double)af.z + DT * 0.0);
        vz *= ((double)fabsf(vz) > 0.003);
        af.x = (float)(-(150.0 / (double)dens) * (double)ndx);               // cuh:298-300
        af.y = (float)(-(150.0 / (double)dens) * (double)ndy);
        af.z = (float)(d.gravity + (-150.0 / (double)dens) * (double)ndz);
        pd.x = x; pd.y = y; pd.z = z;
        vp.x = vx; vp.y = vy; vp.z = vz;
    }
    pd.w = bnd ? -dens : dens;
    vp.w = press;
    if (WITH_KEY) key = bin_id(d, pd.x, pd.y, pd.z);   // FluidGPU.cu:419
}

// The bin id a particle will have AFTER its next Particle::update: the new position is pos + DT*vel of the state the pair sums are
// taken over (FluidGPU.cuh:286-288: the position step uses the velocity from before the update, and no pair sum) — so the NEXT
// step's sort keys exist before this step's pair kernel has even started.  Same expressions, same bits as particle_update.
__device__ __forceinline__ int predicted_key(const FsgDev &d, const float4 &pd, const float4 &vp)
{
    if (pd.w < 0.f) return bin_id(d, pd.x, pd.y, pd.z);                  // boundary particles do not move (cuh:285)
    const double DT = d.dt;
    const float x = (float)((double)pd.x + DT * (double)vp.x);
    const float y = (float)((double)pd.y + DT * (double)vp.y);
    const float z = (float)((double)pd.z + DT * (double)vp.z);
    return bin_id(d, x, y, z);
}

struct PairArgs {
    FsgDev d;
    int n;
    const int *keysA;
    const int *start, *end;
    const int *binlist, *nocc;
    int *work;
    FsgState A, B;
    int *keysB;
    const float4 *carry;
    unsigned long long *stats;
    float4 *sums;       // non-null: write the pair sums here instead of updating the particles (stage API)
};


// fsg_pair_fast.cu
cudaError_t fsg_launch_pair_fast(const PairArgs &a, bool stats, int sm_count, cudaStream_t s);
// fsg_pair_v2.cu
cudaError_t fsg_launch_pair_v2(const PairArgs &a, float4 *sums, bool stats, bool has_boundary, int sm_count, int blocks_per_sm,
                               cudaStream_t s);
// fsg_pair_v3.cu — symmetric pair sums (each pair of particles in different bins is evaluated once); clears `sums` first
cudaError_t fsg_launch_pair_v3(const PairArgs &a, float4 *sums, bool has_boundary, int sm_count, cudaStream_t s);
// pair sums only (no update launch): the deferred-update schedule of fsg_step applies them in the next step's reorder
cudaError_t fsg_launch_pair_sums(const fsg_ctx *c, int64_t n, const int *binlist, const int *nocc, int *work, int *launches, cudaStream_t s);
// part: 0 every slot, 1 the slab's boundary slots (boundary bins, ghosts, parked, dead), 2 the interior slots
cudaError_t fsg_launch_update(const FsgDev &d, int64_t n, const int *keysA, FsgState A, FsgState B, int *keysB,
                              const float4 *sums, const float4 *carry, int part, int *violation, cudaStream_t s);
