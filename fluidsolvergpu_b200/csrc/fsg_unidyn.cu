// fsg_unidyn.cu — the "unidyn" particle step (reference: FluidGPU-unidyn.cu / FluidGPU-unidyn.cuh, driven
// by solver-unidyn.cu:313-573): coarse-bin pair sums (mykernel, :124-449), octant pair sums of the split
// bins (mykernel3, :569-870), update (mykernel2 :451-497 + Particle::update(t) cuh:296-423) and re-binning
// (cell_calc :544-551).
//
// Scope (checked at upload, FSG_E_UNSUPPORTED otherwise): every non-boundary particle is pure fluid
// (solid == 0) with mass 1 — the default scene of solver-unidyn.cu:127-184 and scenes built like it.  For
// such scenes the mixed-phase block (:317-357), mixfactor / vel_grad / stress_accel / mixture_accel /
// delsolid (:368-400) and the granular stress update (:410-446) are exactly zero, merging is unreachable
// (:261) and splitting needs mass > 3 (:278).  Live sums per pair: newdens, newdelpress, diffusion, delfluid.
//
// The reference's dynamic bin splitting is reproduced as what it does to the numbers: a home particle in
// a bin with more than 6 particles (:181) only sees the 8 bins on the side of its octant (:579-583,
// z polarity inverted as in :184) — here a 27-bit mask over the staged neighbourhood, not a second kernel.
// Gather form, one warp per home bin, in-range candidates compacted into a queue, no atomics.
#include "fsg_device.cuh"

#include "fsg_unidyn.cuh"

// Pure-fluid pair sums, one warp per home bin.  The neighbourhood (27 bins by linear offset, the first UNI_TILE = 1024 neighbour
// particles like the reference's 1024 threads) is staged in tiles of UP_TILE candidates — position + density and the sorted slot,
// 22 bytes each, so that 16 warps are resident per SM (the first version staged 39 bytes x 1024 per warp and ran 4 warps per
// SM) — and every home particle sweeps only the candidate ranges it may see: everything for an unsplit bin, the four (dx, dy)
// columns x two z-adjacent bins of its octant for a split one (contiguous in the staged order).  In-range candidates are
// compacted into a queue; their velocity / pressure / volume fraction come from global memory (L2) when the pair is evaluated.
#define UP_WARPS 4
#define UP_TILE 512
struct UpWarpSmem {
    float4 sp[UP_TILE];                 // x, y, z, +-dens (sign = boundary)
    int sj[UP_TILE];                    // sorted slot of the candidate
    unsigned short q[UP_TILE];
};
#define UP_SMEM (sizeof(UpWarpSmem) * UP_WARPS)

template <bool STATS>
__global__ void __launch_bounds__(UP_WARPS * 32)
k_pair_unidyn(UniArgs a)
{
    extern __shared__ __align__(16) unsigned char s_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    UpWarpSmem &S = reinterpret_cast<UpWarpSmem *>(s_raw)[warp];
    const FsgDev &d = a.d;
    const int nocc = *a.nocc;
    const unsigned lt_mask = (1u << lane) - 1u;
    const float alpha_sb = (float)d.alpha_boundary;     // ALPHA__SAND_BOUNDARY for this model
    unsigned long long st_tested = 0, st_in = 0, st_drop = 0;

    for (;;) {
        int m = 0;
        if (lane == 0) m = atomicAdd(a.work, 1);
        m = __shfl_sync(FULL, m, 0);
        if (m >= nocc) break;
        const int b = a.binlist[m];
        int p = 0, st = 0;
        if (lane < 27) {
            int off = (lane / 9 - 1) * d.G2 + ((lane / 3) % 3 - 1) * d.G + (lane % 3 - 1);   // cu:130-132
            int c = b + off;
            if (c >= 0 && c < d.numcells) {
                int s0 = a.start[c], e0 = a.end[c];
                if (s0 >= 0 && e0 >= 0 && s0 < a.n && 1 + e0 - s0 > 0) { p = 1 + e0 - s0; st = s0; }
            }
        }
        int incl = p;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int t = __shfl_up_sync(FULL, incl, o);
            if (lane >= o) incl += t;
        }
        const int excl = incl - p;
        int C = __shfl_sync(FULL, incl, 31);
        if (C > UNI_TILE) {                         // beyond the reference's 1024 threads per block: not reproduced, reported
            if (lane == 0) st_drop += C - UNI_TILE;
            C = UNI_TILE;
        }
        const int hs = __shfl_sync(FULL, st, 13), hn = __shfl_sync(FULL, p, 13);
        const bool split = hn > 6;                  // cu:181

        for (int ig = 0; ig < hn; ig += 32) {
            const int gcount = min(32, hn - ig);
#pragma unroll 1
            for (int t0 = 0; t0 < C; t0 += UP_TILE) {
                const int t1 = min(t0 + UP_TILE, C);
                // ---- stage candidates [t0, t1) of the concatenated neighbourhood ----
                __syncwarp();
#pragma unroll 1
                for (int t = 0; t < 27; t++) {
                    const int pt = __shfl_sync(FULL, p, t);
                    if (pt == 0) continue;
                    const int ex = __shfl_sync(FULL, excl, t), stt = __shfl_sync(FULL, st, t);
                    const int lo = max(ex, t0), hi = min(ex + pt, t1);
                    for (int k = lo + lane; k < hi; k += 32) {
                        const int j = stt + (k - ex);
                        // (asynchronous 16-byte copies: the loop does not wait for one bin's particles before it asks for the next bin's)
                        asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(&S.sp[k - t0])),
                                     "l"(a.A.posd + j) : "memory");
                        S.sj[k - t0] = j;
                    }
                }
                // octants of this group's home particles (cu:182-184), one per lane instead of one warp-wide evaluation per particle
                int my_oct = 0;
                if (split && lane < gcount) { const float4 ph = a.A.posd[hs + ig + lane]; my_oct = uni_subindex(d, ph.x, ph.y, ph.z); }
                asm volatile("cp.async.wait_all;" ::: "memory");
                __syncwarp();
                // the home particle of the NEXT iteration is requested while this one is being processed
                float4 n_pi = a.A.posd[hs + ig], n_vi = a.A.velp[hs + ig], n_mi = a.A.mix[hs + ig];
#pragma unroll 1
                for (int il = 0; il < gcount; il++) {
                    const int i = hs + ig + il;
                    const float4 pi = n_pi, vi = n_vi, mi = n_mi;
                    if (il + 1 < gcount) { n_pi = a.A.posd[i + 1]; n_vi = a.A.velp[i + 1]; n_mi = a.A.mix[i + 1]; }
                    const float densi = fabsf(pi.w);
                    const bool bi = pi.w < 0.f;
                    const float pod2i = __fdividef(vi.w, densi * densi);
                    const float solid_i = mi.x, fluid_i = mi.y, mass_i = 1.f + mi.z;
                    // candidate ranges this particle sees: [0, C) or, in a split bin, 4 columns x 2 z-adjacent bins of its octant
                    int nr = 1, zlo = 0;
                    int ax = 0, ay = 0;
                    if (split) {
                        const int oct = __shfl_sync(FULL, my_oct, il);                // cu:182-184, neighbourhood cu:579-583
                        ax = (oct & 1) ? 1 : -1; ay = (oct & 2) ? 1 : -1;
                        zlo = (oct & 4) ? 0 : 1;                                       // z offsets {-1, 0} (tags +0, +1) or {0, +1} (tags +1, +2)
                        nr = 4;
                    }
                    int qn = 0, ntest = 0;
#pragma unroll 1
                    for (int r = 0; r < nr; r++) {
                        int lo = 0, hi = C;
                        if (split) {
                            const int a_ = (r & 1) ? ax : 0, b_ = (r & 2) ? ay : 0;
                            const int tl = (a_ + 1) * 9 + (b_ + 1) * 3 + zlo;
                            lo = __shfl_sync(FULL, excl, tl);
                            hi = __shfl_sync(FULL, excl, tl + 1) + __shfl_sync(FULL, p, tl + 1);
                            hi = min(hi, C);
                        }
                        if (STATS) ntest += max(hi - lo, 0) * (t0 == 0);               // (counted once, not per tile)
                        lo = max(lo, t0);
                        hi = min(hi, t1);
                        for (int c0 = lo; c0 < hi; c0 += 32) {
                            const int c = c0 + lane;
                            bool in = false;
                            if (c < hi) {
                                const float4 pj = S.sp[c - t0];
                                const float d2 = dist2(pi.x - pj.x, pi.y - pj.y, pi.z - pj.z);
                                in = (d2 <= d.d2_max) && (d2 > 0.f);                   // cu:287
                            }
                            const unsigned mk = __ballot_sync(FULL, in);
                            if (in) S.q[qn + __popc(mk & lt_mask)] = (unsigned short)(c - t0);
                            qn += __popc(mk);
                        }
                    }
                    if (STATS && lane == 0) { st_tested += ntest; st_in += qn; }
                    __syncwarp();
                    float t_nd = 0.f, t_x = 0.f, t_y = 0.f, t_z = 0.f, t_dx = 0.f, t_dy = 0.f, t_dz = 0.f, t_df = 0.f;
                    for (int q = lane; q < qn; q += 32) {
                        const int c = S.q[q];
                        const int j = S.sj[c];
                        const float4 pj = S.sp[c];
                        float4 vj = a.A.velp[j];
                        const float4 mjx = a.A.mix[j];
                        const float fluid_j = mjx.y, mass_j = 1.f + mjx.z;             // Particle::mass of the candidate, cu:358-366
                        const float rx = pi.x - pj.x, ry = pi.y - pj.y, rz = pi.z - pj.z;
                        const float ds = sqrtf(dist2(rx, ry, rz));
                        const float densj = fabsf(pj.w);
                        const bool bj = pj.w < 0.f;
                        vj.w = __fdividef(vj.w, densj * densj);                         // press / powf(dens, 2), cu:310 (approximate reciprocals: <= 2 ulp, the parity bar is 1e-5)
                        // W(ds), FluidGPU-unidyn.cu:11-21
                        float w;
                        const float qq = ds * d.inv_h;
                        if (ds <= d.h_le) w = d.w_c * (1.f - 1.5f * qq * qq + 0.75f * qq * qq * qq);
                        else if (ds <= d.twoh_lt) { float tt = 2.f - qq; w = d.w_c * 0.25f * tt * tt * tt; }
                        else w = 0.f;
                        t_nd += w * ((!bi && bj) ? 2.5f : 1.f) * mass_j;                // cu:362
                        if (ds <= d.h_lt) {                                             // support of dW, cu:35-43
                            const float tt = d.hf - ds;
                            const float g = __fdividef(d.dw_c * tt * tt, ds);
                            const float dkx = g * rx, dky = g * ry, dkz = g * rz;       // cu:296-298
                            const float vabx = vi.x - vj.x, vaby = vi.y - vj.y, vabz = vi.z - vj.z;
                            const float dd = vabx * rx + vaby * ry + vabz * rz;         // cu:304
                            float s = 0.f;
                            if (dd < 0.f) {                                             // cu:307
                                const float mu = __fdividef(dd, ds * ds + d.eps);
                                const float hm = d.hf * mu;
                                const float bf = (!bi && bj) ? 1.f + (1.f + 3.f * fluid_i * fluid_i) * alpha_sb : 1.f;
                                // (cu:307 multiplies the linear term by `SPptr[i].mass` with the raw loop index: the home particle's mass is meant)
                                s = __fdividef(((solid_i * 9.f + 1.f) * (float)d.alpha_fluid) * (float)d.sound * (mass_i * hm + d.visc_q * hm * hm),
                                               (densi + densj) * 0.5f) * bf;
                            }
                            const float pp = vj.w + pod2i + s;                          // cu:310-312
                            t_x += pp * dkx * mass_j;                                   // cu:358-360
                            t_y += pp * dky * mass_j;
                            t_z += pp * dkz * mass_j;
                            if (!bi && !bj) {
                                const float inv = __fdividef(1.f, densj);
                                t_dx += inv * mass_j * dkx;                             // cu:364-366
                                t_dy += inv * mass_j * dky;
                                t_dz += inv * mass_j * dkz;
                                t_df += -0.5f * inv * (fluid_i + fluid_j) * (dkx * vabx + dky * vaby + dkz * vabz);   // cu:401
                            }
                        }
                    }
                    // The eight sums of this home particle over the 32 lanes by recursive halving (9 shuffles instead of 40): after the
                    // rounds with 16, 8 and 4 every lane holds ONE of the eight values summed over the 8 lanes that differ from it in
                    // those bits; two more rounds complete it.  Lane 4 k (k = 0..7) then holds sum number (k&4 ? 4 : 0) + (k&2) + (k&1)
                    // read off its own bits 4, 3, 2 — and stores it (first tile) or adds it to what the earlier tiles stored.
                    float r4[4], r2[2], r1;
                    {
                        const bool h = lane & 16;
                        r4[0] = (h ? t_dx : t_nd) + __shfl_xor_sync(FULL, h ? t_nd : t_dx, 16);
                        r4[1] = (h ? t_dy : t_x) + __shfl_xor_sync(FULL, h ? t_x : t_dy, 16);
                        r4[2] = (h ? t_dz : t_y) + __shfl_xor_sync(FULL, h ? t_y : t_dz, 16);
                        r4[3] = (h ? t_df : t_z) + __shfl_xor_sync(FULL, h ? t_z : t_df, 16);
                        const bool g = lane & 8;
                        r2[0] = (g ? r4[2] : r4[0]) + __shfl_xor_sync(FULL, g ? r4[0] : r4[2], 8);
                        r2[1] = (g ? r4[3] : r4[1]) + __shfl_xor_sync(FULL, g ? r4[1] : r4[3], 8);
                        const bool f = lane & 4;
                        r1 = (f ? r2[1] : r2[0]) + __shfl_xor_sync(FULL, f ? r2[0] : r2[1], 4);
                        r1 += __shfl_xor_sync(FULL, r1, 2);
                        r1 += __shfl_xor_sync(FULL, r1, 1);
                    }
                    if ((lane & 3) == 0) {
                        // component held by this lane: bit 4 -> sums2 (else sums), bit 3 -> +2, bit 2 -> +1   (nd, x, y, z | dx, dy, dz, df)
                        float *dst = reinterpret_cast<float *>((lane & 16) ? a.sums2 + i : a.sums + i) + ((lane >> 2) & 3);
                        *dst = t0 == 0 ? r1 : *dst + r1;
                    }
                    __syncwarp();
                }
            }
        }
        __syncwarp();
    }
    if (STATS && lane == 0) {
        atomicAdd(a.stats + 0, st_tested);
        atomicAdd(a.stats + 1, st_in);
    }
    if (lane == 0 && st_drop) atomicAdd(a.stats + 2, st_drop);
}

// ------------------------------------------------------------------------------------------------
// mykernel2 (cu:451-497) + Particle::update(t) (cuh:296-423) + cell_calc (cu:544-551), per sorted slot.
// Follows the reference's float / double promotions expression by expression.
// ------------------------------------------------------------------------------------------------
// mykernel2's per-particle work after the export: Particle::update(t) (cuh:296-423) for one particle of a live bin,
// then cell_calc's bin id (cu:547).  s = (newdens, newdelpress xyz), s2 = (diffusion xyz, delfluid).
// The granular stress update of one particle, FluidGPU-unidyn.cu:410-446, with the COMPLETED vel_grad sums (race-free reading,
// fsg_unidyn_mixed.cu).  st: stress_tensor[9] then stress_rate[9]; press: the pressure before this step's update.
__device__ __forceinline__ void unidyn_stress_update(float *st, const float *vg, float solid, float press)
{
    if (!(solid != 0.f)) return;                                          // cu:411
    float *sr = st + 9;
    float strain[9];
    float tr = 0, tr3 = 0, tr4 = 0, tr5 = 0;
#pragma unroll
    for (int pq = 0; pq < 9; pq++) strain[pq] = (float)(0.5 * (double)(vg[pq] + vg[3 * (pq % 3) + pq / 3]));                     // :419
#pragma unroll
    for (int p = 0; p < 3; p++) {
#pragma unroll
        for (int q = 0; q < 3; q++) {
            tr3 = (float)((double)tr3 + 0.5 * (double)st[3 * p + q] * (double)st[3 * p + q]);                                   // :421
            tr5 += strain[3 * p + q] * strain[3 * p + q];                                                                        // :423
            tr4 += st[3 * p + q] * strain[3 * q + p];                                                                            // :424
        }
        tr += strain[3 * p + p];                                                                                                 // :426
    }
    const double tp = tan(1.23), root = sqrt(9 + 12 * pow(tp, 2));                                                               // PHI
    const double yield = 3 * tp / root * (double)press * (double)(press > 0) + 1e9 / root;                                       // KC
    const bool scale = yield < (double)tr3 && tr3 != 0;                                                                          // :435
#pragma unroll
    for (int pq = 0; pq < 9; pq++) {
        if (scale) st[pq] = (float)((double)st[pq] * (yield / (double)tr3));                                                     // :436
        // :438 (C1 = 1.5e1, C2 = 0e6, C3 = 5e1)
        sr[pq] = (float)(3 * 1.5e1 * (double)press * ((double)strain[pq] - 1. / 3. * (double)tr * (double)(pq % 4 == 0)) +
                         1.5e1 * 0e6 * ((double)tr4 + (double)(tr * press * (float)(press > 0))) / (pow((double)press, 2) + 1e8) * (double)st[pq] -
                         1.5e1 * 5e1 * sqrt((double)tr5) * (double)st[pq]);
    }
}

// mixed: (stress_accel xyz, mixture_accel xyz, delsolid) of the step, all zero for pure-fluid scenes
struct UniMixedTerms { float sa[3], ma[3], delsolid; };

__device__ __forceinline__ void unidyn_particle_update(const FsgDev &d, float4 &pd, float4 &vp, float4 &af, float4 &dpi, float4 &mx,
                                                       const float4 s, const float4 s2, int &key, const UniMixedTerms &mt)
{
    const float diffx = s2.x, diffy = s2.y, diffz = s2.z;
    float delfluid = s2.w;
    const float delsolid = mt.delsolid;
    const bool bnd = pd.w < 0.f;
    float solid = mx.x, fluid = mx.y;
    const double DT = d.dt;
    // set_dens cuh:183-185, calculate_pressure cuh:282-284 (double pow; RHO_0_SAND == RHO_0)
    const float dens = (float)((double)(s.x + d.w0) / 23.0 * (double)(1 + (float)bnd * 1.5) + 9250);
    const double eos = pow((double)(dens / 9550), 7.0) - 1;
    const float press = (float)((double)((1 - solid) * 1000) * 1.0 * 9550 / 7.0 * eos + (double)(solid * 1000) * 1.0 * 9550 / 7.0 * eos);
    dpi.x = s.y; dpi.y = s.z; dpi.z = s.w;                                // set_delpress cuh:302
    if (!bnd) {
        const float friction = fabsf(diffx) + fabsf(diffy) + fabsf(diffz);   // cuh:311
        solid = (float)((double)solid + DT * (double)delsolid);           // :312-313
        solid *= (solid >= 0.0);
        if ((double)(fluid + delfluid) < 0.2) delfluid = 0;               // :315
        fluid = (float)((double)fluid + DT * (double)delfluid);           // :316-317
        fluid *= (fluid >= 0);
        fluid *= 1 / (fluid + solid);                                     // :319-320
        solid *= 1 / (fluid + solid);
        float x = (float)((double)pd.x + DT * (double)vp.x + 0.5 * DT * DT * (double)af.x + (double)(0 * diffx));   // :328-330
        float y = (float)((double)pd.y + DT * (double)vp.y + 0.5 * DT * DT * (double)af.y + (double)(0 * diffy));
        float z = (float)((double)pd.z + DT * (double)vp.z + 0.5 * DT * DT * (double)af.z + (double)(0 * diffz));
        float vx = vp.x, vy = vp.y, vz = vp.z;
        if (!d.uni_open && (double)z < -0.89) { vx = 0; vy = 0; }         // :332-341
        // :351-353 — the y and z lines test the NEW xvel with xacc (sic), each with its own stress_accel / mixture_accel component
        const double fr = (double)friction * 0.0000002 * (double)solid;
        double tx = (double)vx + DT * (double)af.x + DT * (double)mt.sa[0] + DT * DT * (double)mt.ma[0];
        vx = (float)(((double)vx + 0.5 * DT * (double)af.x + DT * (double)mt.sa[0] + 5 * DT * DT * (double)mt.ma[0]) - (tx > 0) * fr + (tx < 0) * fr);
        tx = (double)vx + DT * (double)af.x + DT * (double)mt.sa[1] + DT * DT * (double)mt.ma[1];
        vy = (float)(((double)vy + 0.5 * DT * (double)af.y + DT * (double)mt.sa[1] + 5 * DT * DT * (double)mt.ma[1]) - (tx > 0) * fr + (tx < 0) * fr);
        tx = (double)vx + DT * (double)af.x + DT * (double)mt.sa[2] + DT * DT * (double)mt.ma[2];
        vz = (float)(((double)vz + 0.5 * DT * (double)af.z + DT * (double)mt.sa[2] + 5 * DT * DT * (double)mt.ma[2]) - (tx > 0) * fr + (tx < 0) * fr);
        af.x = (float)(-((220.0 - 70.0 * (double)solid) / (double)dens) * (double)dpi.x);     // :357-359
        af.y = (float)(-((220.0 - 70.0 * (double)solid) / (double)dens) * (double)dpi.y);
        af.z = (float)(d.gravity + ((-220.0 + 70.0 * (double)solid) / (double)dens) * (double)dpi.z);
        vx = (float)((double)vx + 0.5 * (double)af.x * DT);                // :390-392
        vy = (float)((double)vy + 0.5 * (double)af.y * DT);
        vz = (float)((double)vz + 0.5 * (double)af.z * DT);
        if (!d.uni_open) {          // the reference's unit-box floor and walls are literals of Particle::update (fsg_config.unidyn_open_box)
            if ((double)fabsf(z) > 0.98) { z = (float)(0.97 / (double)z); vz = 0; }           // :404-413
            if ((double)fabsf(y) > 0.98) vy = -vy;
            if ((double)fabsf(x) > 0.98) vx = -vx;
        }
        pd.x = x; pd.y = y; pd.z = z;
        vp.x = vx; vp.y = vy; vp.z = vz;
        mx.x = solid; mx.y = fluid;
    }
    pd.w = bnd ? -dens : dens;
    vp.w = press;
    key = bin_id(d, pd.x, pd.y, pd.z);                                    // cell_calc, cu:547
}

__global__ void __launch_bounds__(256)
k_update_unidyn(FsgDev d, int n, const int *__restrict__ keysA, FsgState A, FsgState B, int *__restrict__ keysB,
                const float4 *__restrict__ sums, const float4 *__restrict__ sums2, const float4 *__restrict__ carry, float *__restrict__ vizb,
                int *violation, const float *__restrict__ mixA, const float *__restrict__ mixB)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float4 pd = A.posd[i], vp = A.velp[i], af = A.accf[i], dpi = A.dpi[i], mx = A.mix[i];
    int key = keysA[i];
    float b3 = 0.f;
    float st[UNI_STRESS];
    if (A.stress) {
#pragma unroll
        for (int k = 0; k < UNI_STRESS; k++) st[k] = A.stress[(size_t)i * UNI_STRESS + k];
    }
    if (key < d.numcells) {
        const int ix = key / d.G2;
        if (ix < d.x0 || ix >= d.x1) {      // slab contexts: ghost copy of a neighbour slab's particle — drop it
            keysB[i] = d.dead;
            vizb[i] = 0.f;
            return;
        }
        float4 s = sums[i], s2 = sums2[i];
        if (carry) { float4 cy = carry[i]; s.x += cy.x; s.y += cy.y; s.z += cy.z; s.w += cy.w; }
        b3 = s2.x * s2.x + s2.y * s2.y + s2.z * s2.z;                         // cu:466
        UniMixedTerms mt = {{0.f, 0.f, 0.f}, {0.f, 0.f, 0.f}, 0.f};
        if (mixA) {                                                           // mixed-phase / granular scene (fsg_unidyn_mixed.cu)
            const float *ma = mixA + (size_t)i * UNI_MIXA, *mb = mixB + (size_t)i * UNI_MIXB;
            unidyn_stress_update(st, ma + 6, mx.x, vp.w);                     // cu:410-446, before mykernel2's update
            mt.sa[0] = ma[15]; mt.sa[1] = ma[16]; mt.sa[2] = ma[17];
            mt.ma[0] = mb[0]; mt.ma[1] = mb[1]; mt.ma[2] = mb[2];
            mt.delsolid = mb[3];
            s2.w = mb[4];                                                     // delfluid comes from pass B
        }
        if (A.stress) {
#pragma unroll
            for (int k = 0; k < 9; k++) st[k] = (float)(d.dt * (double)st[9 + k]);     // stress_tensor = DT * stress_rate, cuh:304-308
        }
        unidyn_particle_update(d, pd, vp, af, dpi, mx, s, s2, key, mt);
        // the one-layer ghost band assumes less than one bin layer per step
        if (violation && key < d.numcells && abs(key / d.G2 - ix) > 1) atomicOr(violation, 1);
    }
    B.posd[i] = pd;
    B.velp[i] = vp;
    B.accf[i] = af;
    B.dpi[i] = dpi;
    B.mix[i] = mx;
    if (A.stress) {
#pragma unroll
        for (int k = 0; k < UNI_STRESS; k++) B.stress[(size_t)i * UNI_STRESS + k] = st[k];
    }
    keysB[i] = key;
    vizb[i] = b3;
}

// split[] as mykernel leaves it (cu:181-190): bin id for bins with more than 6 particles, else -1
__global__ void k_split_table(int numcells, const int *__restrict__ start, const int *__restrict__ end, int *split)
{
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= numcells) return;
    int s0 = start[c];
    split[c] = (s0 >= 0 && end[c] - s0 + 1 > 6) ? c : -1;
}

cudaError_t fsg_launch_unidyn(fsg_ctx *c, int64_t n, const int *binlist, const int *nocc, int *work, const float4 *carry,
                              int *launches, cudaStream_t s)
{
    if (n <= 0) return cudaSuccess;
    static FsgAttrOnce attr_once;
    if (attr_once.need()) {
        cudaFuncSetAttribute(k_pair_unidyn<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)UP_SMEM);
        cudaFuncSetAttribute(k_pair_unidyn<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)UP_SMEM);
    }
    UniArgs a;
    a.d = c->dev;
    a.n = (int)n;
    a.start = c->start;
    a.end = c->end;
    a.binlist = binlist;
    a.nocc = nocc;
    a.work = work;
    a.A = c->A;
    a.sums = c->sums;
    a.sums2 = c->sums2;
    a.stats = c->dstats;
    a.mixA = c->mixA;
    a.mixB = c->mixB;
    int64_t blocks = (n + UP_WARPS - 1) / UP_WARPS;
    int64_t maxb = (int64_t)c->sm_count * 5;           // 5 x 45 KB of shared memory per SM
    if (blocks > maxb) blocks = maxb;
    cudaError_t e;
    if (c->mixed) {
        // mixed-phase / granular scene: pass A, then pass B over the completed drift sums (fsg_unidyn_mixed.cu); the pair counters
        // of collect_stats come from the pure-fluid kernel's bookkeeping and are not kept here
        e = fsg_launch_unidyn_mixed(a, 0, c->sm_count, s);
        if (e != cudaSuccess) return e;
        e = cudaMemsetAsync(work, 0, sizeof(int), s);
        if (e != cudaSuccess) return e;
        e = fsg_launch_unidyn_mixed(a, 1, c->sm_count, s);
        if (e != cudaSuccess) return e;
        *launches += 1;
    } else {
        if (c->cfg.collect_stats) k_pair_unidyn<true><<<(unsigned)blocks, UP_WARPS * 32, UP_SMEM, s>>>(a);
        else k_pair_unidyn<false><<<(unsigned)blocks, UP_WARPS * 32, UP_SMEM, s>>>(a);
        e = cudaGetLastError();
        if (e != cudaSuccess) return e;
        // particle merging / splitting (fsg_unidyn_adapt.cu): on the sorted state, between the pair sums and the update
        if (c->cfg.unidyn_adapt) { e = fsg_unidyn_adapt_pre(c, n, s); if (e != cudaSuccess) return e; }
    }
    k_update_unidyn<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(c->dev, (int)n, c->keysA, c->A, c->B, c->keysB, c->sums, c->sums2, carry,
                                                              c->vizb, c->cfg.world > 1 ? c->counters + 6 : nullptr,
                                                              c->mixed ? c->mixA : nullptr, c->mixed ? c->mixB : nullptr);
    *launches += 2;
    return cudaGetLastError();
}

cudaError_t fsg_launch_split_table(const fsg_ctx *c, int *split, cudaStream_t s)
{
    const int nc = c->dev.numcells;
    k_split_table<<<(unsigned)((nc + 255) / 256), 256, 0, s>>>(nc, c->start, c->end, split);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// AoS <-> SoA for the unidyn Particle record (offsets: FluidGPU-unidyn.cuh:119-181 as laid out by nvcc,
// sizeof == 340)
// ------------------------------------------------------------------------------------------------
namespace aos_uni {
enum { POS = 0, VEL = 12, ACC = 36, INDEX = 60, CELL = 64, SUBINDEX = 68, MASS = 72, DENS = 76, PRESS = 80, DELP_Z = 84, DELP_Y = 88,
       DELP_X = 92, DIFFUSION = 96, NEWDENS = 108, NDELP_Z = 112, NDELP_Y = 116, NDELP_X = 120, BOUNDARY = 316, SOLID = 320, FLUID = 324,
       DELSOLID = 328, DELFLUID = 332, FLAG = 336, SPLIT = 337, STRESS_RATE = 220, STRESS_TENSOR = 256 };
}
__device__ __forceinline__ float uldf(const unsigned char *r, int off) { return *reinterpret_cast<const float *>(r + off); }
__device__ __forceinline__ void ustf(unsigned char *r, int off, float v) { *reinterpret_cast<float *>(r + off) = v; }

__global__ void k_unpack_aos_unidyn(const unsigned char *__restrict__ aos, int64_t n, FsgState st, float4 *carry, int *bad)
{
    using namespace aos_uni;
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const unsigned char *r = aos + i * FSG_AOS_STRIDE;
    bool bnd = r[BOUNDARY] != 0;
    float dens = uldf(r, DENS), solid = uldf(r, SOLID);
    st.posd[i] = make_float4(uldf(r, POS), uldf(r, POS + 4), uldf(r, POS + 8), bnd ? -dens : dens);
    st.velp[i] = make_float4(uldf(r, VEL), uldf(r, VEL + 4), uldf(r, VEL + 8), uldf(r, PRESS));
    st.accf[i] = make_float4(uldf(r, ACC), uldf(r, ACC + 4), uldf(r, ACC + 8), __int_as_float(bnd ? 1 : 0));
    st.dpi[i] = make_float4(uldf(r, DELP_X), uldf(r, DELP_Y), uldf(r, DELP_Z), __int_as_float(*reinterpret_cast<const int *>(r + INDEX)));
    st.mix[i] = make_float4(solid, uldf(r, FLUID), uldf(r, MASS) - 1.f, 0.f);          // z = mass - 1 (fsg_unidyn_adapt.cu)
    carry[i] = make_float4(uldf(r, NEWDENS), uldf(r, NDELP_X), uldf(r, NDELP_Y), uldf(r, NDELP_Z));
    if (st.stress) {
        for (int k = 0; k < 9; k++) {
            st.stress[(size_t)i * UNI_STRESS + k] = uldf(r, STRESS_TENSOR + 4 * k);
            st.stress[(size_t)i * UNI_STRESS + 9 + k] = uldf(r, STRESS_RATE + 4 * k);
        }
    }
    if (uldf(r, MASS) != 1.f) atomicOr(bad, 1);           // outside the scope (merged / split particles)
    if (!bnd && solid != 0.f) atomicOr(bad, 2);           // a mixed-phase / granular scene
}

__global__ void k_pack_aos_unidyn(unsigned char *__restrict__ aos, int64_t n, FsgState st, const float4 *carry, const int *keys, FsgDev d)
{
    using namespace aos_uni;
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    unsigned char *r = aos + i * FSG_AOS_STRIDE;
    for (int o = 0; o < FSG_AOS_STRIDE; o += 4) *reinterpret_cast<int *>(r + o) = 0;
    float4 pd = st.posd[i], vp = st.velp[i], af = st.accf[i], dpi = st.dpi[i], mx = st.mix[i];
    ustf(r, POS, pd.x); ustf(r, POS + 4, pd.y); ustf(r, POS + 8, pd.z);
    ustf(r, VEL, vp.x); ustf(r, VEL + 4, vp.y); ustf(r, VEL + 8, vp.z);
    ustf(r, ACC, af.x); ustf(r, ACC + 4, af.y); ustf(r, ACC + 8, af.z);
    *reinterpret_cast<int *>(r + INDEX) = __float_as_int(dpi.w);
    *reinterpret_cast<int *>(r + CELL) = keys[i];
    *reinterpret_cast<int *>(r + SUBINDEX) = uni_subindex(d, pd.x, pd.y, pd.z);
    ustf(r, MASS, 1.f + mx.z);
    ustf(r, DENS, fabsf(pd.w));
    ustf(r, PRESS, vp.w);
    ustf(r, DELP_X, dpi.x); ustf(r, DELP_Y, dpi.y); ustf(r, DELP_Z, dpi.z);
    float4 cy = carry ? carry[i] : make_float4(0.f, 0.f, 0.f, 0.f);
    ustf(r, NEWDENS, cy.x);
    ustf(r, NDELP_X, cy.y); ustf(r, NDELP_Y, cy.z); ustf(r, NDELP_Z, cy.w);
    r[BOUNDARY] = pd.w < 0.f ? 1 : 0;
    ustf(r, SOLID, mx.x);
    ustf(r, FLUID, mx.y);
    if (st.stress) {
        for (int k = 0; k < 9; k++) {
            ustf(r, STRESS_TENSOR + 4 * k, st.stress[(size_t)i * UNI_STRESS + k]);
            ustf(r, STRESS_RATE + 4 * k, st.stress[(size_t)i * UNI_STRESS + 9 + k]);
        }
    }
    r[FLAG] = 1;                                     // update() leaves flag = true, cuh:422
}

cudaError_t fsg_launch_unpack_aos_unidyn(const unsigned char *aos, int64_t n, FsgState st, float4 *carry, int *bad, cudaStream_t s)
{
    if (n <= 0) return cudaSuccess;
    k_unpack_aos_unidyn<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(aos, n, st, carry, bad);
    return cudaGetLastError();
}
cudaError_t fsg_launch_pack_aos_unidyn(unsigned char *aos, int64_t n, FsgState st, const float4 *carry, const int *keys, const FsgDev &d,
                                       cudaStream_t s)
{
    if (n <= 0) return cudaSuccess;
    k_pack_aos_unidyn<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(aos, n, st, carry, keys, d);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// Stage API for the unidyn model: one call per reference launch of solver-unidyn.cu:341-548, on caller-owned
// DEVICE buffers in the reference's layout (340-byte unidyn Particle records, int tables).  The context (model
// FSG_MODEL_UNIDYN) supplies constants and scratch memory.  Single-device form of the loop (x origin 0, buffer 0 —
// what the shipped driver runs, solver-unidyn.cu:192-195).  The inter-kernel protocol is kept: mykernel marks the
// split bins, sets subindex and ADDS the pair sums of the particles in unsplit bins onto the records' accumulators,
// mykernel3 adds those of the split bins, mykernel2 exports / updates / zeroes / resets, cell_calc re-bins.
// ------------------------------------------------------------------------------------------------
#include <stdio.h>
#define CUU(ctx, call)                                                                                  \
    do {                                                                                                \
        cudaError_t e_ = (call);                                                                        \
        if (e_ != cudaSuccess) {                                                                        \
            char b_[512];                                                                               \
            snprintf(b_, sizeof b_, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
            (ctx)->err = b_;                                                                            \
            return e_ == cudaErrorMemoryAllocation ? FSG_E_NOMEM : FSG_E_CUDA;                          \
        }                                                                                               \
    } while (0)

static int uni_stage_ok(fsg_ctx *c, int64_t n, const char *what)
{
    if (c->cfg.model != FSG_MODEL_UNIDYN) { c->err = std::string(what) + ": the context is not a unidyn context"; return FSG_E_STATE; }
    if (n > c->cap) { c->err = std::string(what) + ": n exceeds the context capacity"; return FSG_E_INVALID; }
    return FSG_OK;
}

// count_after_merge, FluidGPU-unidyn.cu:554-562: the first sorted slot whose bin id is outside the grid
__global__ void k_stage_uni_count(const int *__restrict__ cells, int64_t n, int numcells, int *newsize)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i > 0 && i < n && cells[i] >= numcells && cells[i - 1] < numcells) *newsize = (int)i;
}
extern "C" int fsg_stage_unidyn_count_after_merge(fsg_ctx *c, const int32_t *d_cells, int64_t n, int32_t *d_newsize)
{
    if (!c || n < 0 || !d_newsize || (n > 0 && !d_cells)) return FSG_E_INVALID;
    if (n == 0) return FSG_OK;
    CUU(c, cudaSetDevice(c->device));
    k_stage_uni_count<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(d_cells, n, c->dev.numcells, d_newsize);
    CUU(c, cudaGetLastError());
    c->launches++;
    return FSG_OK;
}

// findneighbours, FluidGPU-unidyn.cu:106-122 (x = id of the slab's first bin)
__global__ void k_stage_uni_findneighbours(const int *__restrict__ cell, int *start, int *start_copy, int *end, int64_t n, int x, int numcells)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int key = cell[i] - x;
    if (key < 0 || key >= numcells) return;       // the reference writes out of bounds here
    if (i == 0 || cell[i - 1] != cell[i]) { start[key] = (int)i; start_copy[key] = (int)i; }
    if (i == n - 1 || cell[i + 1] != cell[i]) end[key] = (int)i;
}
extern "C" int fsg_stage_unidyn_findneighbours(fsg_ctx *c, const int32_t *d_cells, int32_t *d_start, int32_t *d_start_copy, int32_t *d_end,
                                               int64_t n, int32_t x)
{
    if (!c || n < 0 || (n > 0 && (!d_cells || !d_start || !d_start_copy || !d_end))) return FSG_E_INVALID;
    if (n == 0) return FSG_OK;
    CUU(c, cudaSetDevice(c->device));
    k_stage_uni_findneighbours<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(d_cells, d_start, d_start_copy, d_end, n, x, c->dev.numcells);
    CUU(c, cudaGetLastError());
    c->launches++;
    return FSG_OK;
}

// records -> the SoA streams the pair kernel reads + the list of occupied bins of one kind (which = 0: at most 6
// particles, mykernel's bins; 1: more than 6, mykernel3's).  mark: also do mykernel's split marking (cu:181-191).
__global__ void __launch_bounds__(256)
k_stage_uni_unpack(FsgDev d, const unsigned char *aos_c, unsigned char *aos, const int *__restrict__ cell,
                   const int *__restrict__ start, const int *__restrict__ end, int64_t n, FsgState st, int *binlist, int *nocc, int which,
                   int *split, int *numsplit)
{
    using namespace aos_uni;
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    bool head = false;
    int key = -1;
    if (i < n) {
        const unsigned char *r = aos_c + i * FSG_AOS_STRIDE;
        const bool bnd = r[BOUNDARY] != 0;
        const float dens = uldf(r, DENS);
        const float4 pd = make_float4(uldf(r, POS), uldf(r, POS + 4), uldf(r, POS + 8), bnd ? -dens : dens);
        st.posd[i] = pd;
        st.velp[i] = make_float4(uldf(r, VEL), uldf(r, VEL + 4), uldf(r, VEL + 8), uldf(r, PRESS));
        st.mix[i] = make_float4(uldf(r, SOLID), uldf(r, FLUID), 0.f, 0.f);
        key = cell[i];
        if (key >= 0 && key < d.numcells) {
            const int pop = end[key] - start[key] + 1;
            const bool is_split = pop > 6;
            const bool first = i == 0 || cell[i - 1] != key;
            head = first && (is_split == (which == 1));
            if (split && is_split) {
                *reinterpret_cast<int *>(aos + i * FSG_AOS_STRIDE + SUBINDEX) = uni_subindex(d, pd.x, pd.y, pd.z);   // cu:182-184
                if (first) { split[key] = key; atomicAdd(numsplit, 1); }                                              // cu:186-189
            }
        }
    }
    __shared__ int s_cnt[8], s_base;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned m = __ballot_sync(FULL, head);
    if (lane == 0) s_cnt[warp] = __popc(m);
    __syncthreads();
    if (threadIdx.x == 0) {
        int tot = 0;
#pragma unroll
        for (int w = 0; w < 8; w++) { int cc = s_cnt[w]; s_cnt[w] = tot; tot += cc; }
        s_base = tot ? atomicAdd(nocc, tot) : 0;
    }
    __syncthreads();
    if (head) binlist[s_base + s_cnt[warp] + __popc(m & ((1u << lane) - 1))] = key;
}

// the atomicAdd targets of the pair loops (cu:358-366, 401): sums are ADDED to the accumulators of the particles
// whose bin was in this launch's list
__global__ void k_stage_uni_add_sums(unsigned char *__restrict__ aos, const int *__restrict__ cell, const int *__restrict__ start,
                                     const int *__restrict__ end, const float4 *__restrict__ sums, const float4 *__restrict__ sums2,
                                     int64_t n, int numcells, int which)
{
    using namespace aos_uni;
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int key = cell[i];
    if (key < 0 || key >= numcells) return;
    if ((end[key] - start[key] + 1 > 6) != (which == 1)) return;
    unsigned char *r = aos + i * FSG_AOS_STRIDE;
    const float4 s = sums[i], s2 = sums2[i];
    ustf(r, NEWDENS, uldf(r, NEWDENS) + s.x);
    ustf(r, NDELP_X, uldf(r, NDELP_X) + s.y);
    ustf(r, NDELP_Y, uldf(r, NDELP_Y) + s.z);
    ustf(r, NDELP_Z, uldf(r, NDELP_Z) + s.w);
    ustf(r, DIFFUSION, uldf(r, DIFFUSION) + s2.x);
    ustf(r, DIFFUSION + 4, uldf(r, DIFFUSION + 4) + s2.y);
    ustf(r, DIFFUSION + 8, uldf(r, DIFFUSION + 8) + s2.z);
    ustf(r, DELFLUID, uldf(r, DELFLUID) + s2.w);
}

static int uni_stage_pairs(fsg_ctx *c, void *d_particles, const int32_t *d_cells, const int32_t *d_start, const int32_t *d_end,
                           int32_t *d_split, int32_t *d_numsplit, int64_t n, int which)
{
    CUU(c, cudaSetDevice(c->device));
    static FsgAttrOnce attr_once;
    if (attr_once.need()) {
        cudaFuncSetAttribute(k_pair_unidyn<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)UP_SMEM);
        cudaFuncSetAttribute(k_pair_unidyn<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)UP_SMEM);
    }
    int *binlist = c->binlist[0], *nocc = c->counters + 7, *work = c->counters + 2;
    CUU(c, cudaMemsetAsync(nocc, 0, sizeof(int), c->stream));
    CUU(c, cudaMemsetAsync(work, 0, sizeof(int), c->stream));
    k_stage_uni_unpack<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(c->dev, (const unsigned char *)d_particles, (unsigned char *)d_particles,
                                                                         d_cells, d_start, d_end, n, c->A, binlist, nocc, which, d_split, d_numsplit);
    CUU(c, cudaGetLastError());
    UniArgs a;
    a.d = c->dev;
    a.n = (int)n;
    a.start = d_start;
    a.end = d_end;
    a.binlist = binlist;
    a.nocc = nocc;
    a.work = work;
    a.A = c->A;
    a.sums = c->sums;
    a.sums2 = c->sums2;
    a.stats = c->dstats;
    int64_t blocks = (n + UP_WARPS - 1) / UP_WARPS;
    const int64_t maxb = (int64_t)c->sm_count * 2;
    if (blocks > maxb) blocks = maxb;
    k_pair_unidyn<false><<<(unsigned)blocks, UP_WARPS * 32, UP_SMEM, c->stream>>>(a);
    CUU(c, cudaGetLastError());
    k_stage_uni_add_sums<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>((unsigned char *)d_particles, d_cells, d_start, d_end, c->sums,
                                                                           c->sums2, n, c->dev.numcells, which);
    CUU(c, cudaGetLastError());
    c->launches += 3;
    return FSG_OK;
}

extern "C" int fsg_stage_unidyn_mykernel(fsg_ctx *c, void *d_particles, const int32_t *d_cells, const int32_t *d_start, const int32_t *d_end,
                                         int32_t *d_split, int32_t *d_numsplit, int64_t n)
{
    if (!c || n < 0 || (n > 0 && (!d_particles || !d_cells || !d_start || !d_end || !d_split || !d_numsplit))) return FSG_E_INVALID;
    int rc = uni_stage_ok(c, n, "fsg_stage_unidyn_mykernel");
    if (rc != FSG_OK || n == 0) return rc;
    return uni_stage_pairs(c, d_particles, d_cells, d_start, d_end, d_split, d_numsplit, n, 0);
}
extern "C" int fsg_stage_unidyn_mykernel3(fsg_ctx *c, void *d_particles, const int32_t *d_cells, const int32_t *d_start, const int32_t *d_end,
                                          int64_t n)
{
    if (!c || n < 0 || (n > 0 && (!d_particles || !d_cells || !d_start || !d_end))) return FSG_E_INVALID;
    int rc = uni_stage_ok(c, n, "fsg_stage_unidyn_mykernel3");
    if (rc != FSG_OK || n == 0) return rc;
    return uni_stage_pairs(c, d_particles, d_cells, d_start, d_end, nullptr, nullptr, n, 1);
}

// mykernel2 (cu:451-497) on the records: export of the pre-update state, Particle::update(t), stale cells[], accumulators
// zeroed, tables / split / numsplit reset.  The new bin id is NOT stored here — that is cell_calc's job (cu:544-551).
__global__ void __launch_bounds__(256)
k_stage_uni_mykernel2(FsgDev d, unsigned char *__restrict__ aos, int *__restrict__ cells, int *start_copy, int *start, int *end, int *split,
                      int *numsplit, int64_t n, int x, float *spts, float *a3, float *b3)
{
    using namespace aos_uni;
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        unsigned char *r = aos + i * FSG_AOS_STRIDE;
        const bool bnd = r[BOUNDARY] != 0;
        const int cellnumber = *reinterpret_cast<const int *>(r + CELL);
        const float dens0 = uldf(r, DENS);
        float4 pd = make_float4(uldf(r, POS), uldf(r, POS + 4), uldf(r, POS + 8), bnd ? -dens0 : dens0);
        const float4 s = make_float4(uldf(r, NEWDENS), uldf(r, NDELP_X), uldf(r, NDELP_Y), uldf(r, NDELP_Z));
        const float4 s2 = make_float4(uldf(r, DIFFUSION), uldf(r, DIFFUSION + 4), uldf(r, DIFFUSION + 8), uldf(r, DELFLUID));
        if (cellnumber >= 0 && cellnumber < x) {                                         // cu:459-466 (lb = 0, hb = x)
            if (spts) { spts[3 * i] = pd.x; spts[3 * i + 1] = pd.y; spts[3 * i + 2] = pd.z; }
            if (a3) a3[i] = uldf(r, MASS);
            if (b3) b3[i] = s2.x * s2.x + s2.y * s2.y + s2.z * s2.z;
        }
        float4 vp = make_float4(uldf(r, VEL), uldf(r, VEL + 4), uldf(r, VEL + 8), uldf(r, PRESS));
        float4 af = make_float4(uldf(r, ACC), uldf(r, ACC + 4), uldf(r, ACC + 8), 0.f);
        float4 dpi = make_float4(0.f, 0.f, 0.f, 0.f);
        float4 mx = make_float4(uldf(r, SOLID), uldf(r, FLUID), 0.f, 0.f);
        int key;
        const UniMixedTerms none = {{0.f, 0.f, 0.f}, {0.f, 0.f, 0.f}, 0.f};              // (the stage API keeps the pure-fluid scope)
        unidyn_particle_update(d, pd, vp, af, dpi, mx, s, s2, key, none);                // cu:469
        ustf(r, POS, pd.x); ustf(r, POS + 4, pd.y); ustf(r, POS + 8, pd.z);
        ustf(r, VEL, vp.x); ustf(r, VEL + 4, vp.y); ustf(r, VEL + 8, vp.z);
        ustf(r, ACC, af.x); ustf(r, ACC + 4, af.y); ustf(r, ACC + 8, af.z);
        ustf(r, DENS, fabsf(pd.w));
        ustf(r, PRESS, vp.w);
        ustf(r, DELP_X, dpi.x); ustf(r, DELP_Y, dpi.y); ustf(r, DELP_Z, dpi.z);
        ustf(r, SOLID, mx.x);
        ustf(r, FLUID, mx.y);
        for (int q = 0; q < 9; q++) ustf(r, 256 + 4 * q, (float)(d.dt * (double)uldf(r, 220 + 4 * q)));   // stress_tensor = DT * stress_rate, cuh:304-308
        r[FLAG] = 1;                                                                     // cuh:422
        cells[i] = cellnumber;                                                           // cu:474 (stale; cell_calc follows)
        ustf(r, NEWDENS, 0.f); ustf(r, NDELP_X, 0.f); ustf(r, NDELP_Y, 0.f); ustf(r, NDELP_Z, 0.f);      // cu:475-478
        for (int o = 124; o < 184; o += 4) ustf(r, o, 0.f);                              // drift velocities, vel_grad  cu:479,481
        for (int o = 292; o < 316; o += 4) ustf(r, o, 0.f);                              // stress_accel, mixture_accel  cu:480,482
        ustf(r, DELSOLID, 0.f); ustf(r, DELFLUID, 0.f);                                  // cu:483
        ustf(r, DIFFUSION, 0.f); ustf(r, DIFFUSION + 4, 0.f); ustf(r, DIFFUSION + 8, 0.f);
    }
    if (i < x) { start[i] = -1; start_copy[i] = -1; end[i] = -1; split[i] = -1; }        // cu:486-491
    if (i == 0) numsplit[0] = 0;                                                         // cu:493
}
extern "C" int fsg_stage_unidyn_mykernel2(fsg_ctx *c, void *d_particles, int32_t *d_cells, int32_t *d_start_copy, int32_t *d_start, int32_t *d_end,
                                          int32_t *d_split, int32_t *d_numsplit, int64_t n, int32_t x, int32_t t, float *spts, float *a3, float *b3)
{
    (void)t;      // Particle::update(int t) never reads t (its Runge-Kutta branch is commented out, cuh:361-387)
    if (!c || n < 0 || x < 0 || !d_start_copy || !d_start || !d_end || !d_split || !d_numsplit || (n > 0 && (!d_particles || !d_cells)))
        return FSG_E_INVALID;
    int rc = uni_stage_ok(c, n, "fsg_stage_unidyn_mykernel2");
    if (rc != FSG_OK) return rc;
    CUU(c, cudaSetDevice(c->device));
    const int64_t threads = n > x ? n : x;
    if (threads == 0) return FSG_OK;
    k_stage_uni_mykernel2<<<(unsigned)((threads + 255) / 256), 256, 0, c->stream>>>(c->dev, (unsigned char *)d_particles, d_cells, d_start_copy,
                                                                                  d_start, d_end, d_split, d_numsplit, n, x, spts, a3, b3);
    CUU(c, cudaGetLastError());
    c->launches++;
    return FSG_OK;
}

// cell_calc, cu:544-551
__global__ void k_stage_uni_cell_calc(FsgDev d, unsigned char *__restrict__ aos, int *__restrict__ cells, int64_t n)
{
    using namespace aos_uni;
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    unsigned char *r = aos + i * FSG_AOS_STRIDE;
    const int key = bin_id(d, uldf(r, POS), uldf(r, POS + 4), uldf(r, POS + 8));
    *reinterpret_cast<int *>(r + CELL) = key;
    cells[i] = key;
}
extern "C" int fsg_stage_unidyn_cell_calc(fsg_ctx *c, void *d_particles, int32_t *d_cells, int64_t n)
{
    if (!c || n < 0 || (n > 0 && (!d_particles || !d_cells))) return FSG_E_INVALID;
    int rc = uni_stage_ok(c, n, "fsg_stage_unidyn_cell_calc");
    if (rc != FSG_OK || n == 0) return rc;
    CUU(c, cudaSetDevice(c->device));
    k_stage_uni_cell_calc<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(c->dev, (unsigned char *)d_particles, d_cells, n);
    CUU(c, cudaGetLastError());
    c->launches++;
    return FSG_OK;
}
