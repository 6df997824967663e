// fsg_unidyn_mixed.cu — the unidyn pair sums for MIXED-PHASE / GRANULAR scenes (some non-boundary particle with solid != 0):
// the terms of FluidGPU-unidyn.cu:314-401 that are exactly zero for pure-fluid scenes (fsg_unidyn.cu).
//
// The reference evaluates them in one launch in which mixture_accel / delsolid / delfluid read the drift velocities of both
// particles of a pair (:385-401) while other blocks are still accumulating those (:351-357), and its stress update reads
// vel_grad sums that are not complete (:410-446) — its result is not a function of its input.  libfsg runs the
// race-free reading, the one oracle/fsg_oracle_unidyn.c restates and these kernels are pinned against:
//
//   pass A  per home particle: newdens, newdelpress, diffusion (as in fsg_unidyn.cu) + solid / fluid drift velocities (:317-357),
//           vel_grad (:368-377), stress_accel (:379-381)                                   -> sums, sums2.xyz, mixA
//   pass B  mixture_accel (:383-398), delsolid, delfluid (:400-401) from the COMPLETED drift velocities of both particles -> mixB
//   then    one stress update per particle with the completed vel_grad and Particle::update (k_update_unidyn, fsg_unidyn.cu).
//
// Same neighbourhood rules as the pure-fluid kernel: 27 bins by linear offset, the first 1024 neighbour particles, the 8-bin
// octant neighbourhood for home bins with more than 6 particles.  Gather form, one warp per home bin, no atomics: deterministic.
// The frame is the pure-fluid kernel's (fsg_unidyn.cu): the neighbourhood is staged in tiles of UM_TILE candidates with asynchronous
// 16-byte copies (22 bytes each, 4 warps x 5 blocks per SM), a home particle sweeps only the candidate ranges it may see (its
// octant's four columns x two z-adjacent bins in a split bin), in-range candidates are queued, the next home particle is requested
// while this one is processed, and the 25 (pass A) / 5 (pass B) sums of a home particle are reduced over the warp by recursive
// halving — 31 shuffles instead of 125, lane k ends up holding sum k and stores it.  The pair bodies follow the reference's
// expression order and promote where it does.
#include "fsg_unidyn.cuh"

#define UM_WARPS 4
#define UM_TILE 512
#define UM_MIXPRESSURE 1e-12
#define UM_MIXBROWNIAN 5e-9

struct UmWarpSmem {
    float4 sp[UM_TILE];                 // x, y, z, +-dens
    int sj[UM_TILE];                    // sorted slot of the candidate
    unsigned short q[UM_TILE];
};
#define UM_SMEM (sizeof(UmWarpSmem) * UM_WARPS)

// v[0 .. 31] summed over the 32 lanes: afterwards v[0] of lane k is the total of value k.  Five halving rounds (16 + 8 + 4 + 2 + 1 = 31
// shuffles): a lane keeps the half of the values that matches its lane bit and hands the other half to its partner.
__device__ __forceinline__ float um_reduce_halving(float (&v)[32], int lane)
{
#pragma unroll
    for (int w = 16; w >= 1; w >>= 1) {
        const bool hi = lane & w;
#pragma unroll
        for (int k = 0; k < w; k++) {
            const float keep = hi ? v[k + w] : v[k], send = hi ? v[k] : v[k + w];
            v[k] = keep + __shfl_xor_sync(FULL, send, w);
        }
    }
    return v[0];
}

template <int PASS>
__global__ void __launch_bounds__(UM_WARPS * 32)
k_pair_unidyn_mixed(UniArgs a)
{
    extern __shared__ __align__(16) unsigned char s_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    UmWarpSmem &S = reinterpret_cast<UmWarpSmem *>(s_raw)[warp];
    const FsgDev &d = a.d;
    const int nocc = *a.nocc;
    const unsigned lt_mask = (1u << lane) - 1u;
    const float alpha_sb = (float)d.alpha_boundary;

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
        if (C > UNI_TILE) C = UNI_TILE;             // (counted as dropped by the pure-fluid bookkeeping; same truncation here)
        const int hs = __shfl_sync(FULL, st, 13), hn = __shfl_sync(FULL, p, 13);
        const bool split = hn > 6;                  // cu:181

        for (int ig = 0; ig < hn; ig += 32) {
        const int gcount = min(32, hn - ig);
#pragma unroll 1
        for (int t0 = 0; t0 < C; t0 += UM_TILE) {
        const int t1 = min(t0 + UM_TILE, C);
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
                asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(&S.sp[k - t0])),
                             "l"(a.A.posd + j) : "memory");
                S.sj[k - t0] = j;
            }
        }
        int my_oct = 0;                              // octants of this group's home particles (cu:182-184), one per lane
        if (split && lane < gcount) { const float4 ph = a.A.posd[hs + ig + lane]; my_oct = uni_subindex(d, ph.x, ph.y, ph.z); }
        asm volatile("cp.async.wait_all;" ::: "memory");
        __syncwarp();

        float4 n_pi = a.A.posd[hs + ig], n_vi = a.A.velp[hs + ig], n_mi = a.A.mix[hs + ig];
#pragma unroll 1
        for (int il = 0; il < gcount; il++) {
            const int i = hs + ig + il;
            const float4 pi = n_pi, vi = n_vi, mi = n_mi;
            if (il + 1 < gcount) { n_pi = a.A.posd[i + 1]; n_vi = a.A.velp[i + 1]; n_mi = a.A.mix[i + 1]; }
            const float densi = fabsf(pi.w);
            const bool bi = pi.w < 0.f;
            const float solid_i = mi.x, fluid_i = mi.y, press_i = vi.w;
            // candidate ranges this particle sees: [0, C) or, in a split bin, 4 columns x 2 z-adjacent bins of its octant (cu:579-583)
            int nr = 1, zlo = 0, ax = 0, ay = 0;
            if (split) {
                const int oct = __shfl_sync(FULL, my_oct, il);
                ax = (oct & 1) ? 1 : -1; ay = (oct & 2) ? 1 : -1;
                zlo = (oct & 4) ? 0 : 1;
                nr = 4;
            }
            int qn = 0;
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
                lo = max(lo, t0);
                hi = min(hi, t1);
                for (int c0 = lo; c0 < hi; c0 += 32) {
                    const int c = c0 + lane;
                    bool in = false;
                    if (c < hi) {
                        const float4 pj = S.sp[c - t0];
                        const float d2 = dist2(pi.x - pj.x, pi.y - pj.y, pi.z - pj.z);
                        in = (d2 <= d.d2_max) && (d2 > 0.f);             // cu:287
                        if (PASS == 1) in = in && sqrtf(d2) <= d.h_lt;   // every pass-B term carries dW, whose support is h (cu:35-43)
                    }
                    const unsigned mk = __ballot_sync(FULL, in);
                    if (in) S.q[qn + __popc(mk & lt_mask)] = (unsigned short)(c - t0);
                    qn += __popc(mk);
                }
            }
            __syncwarp();

            if (PASS == 0) {
                // acc: 0 newdens, 1-3 newdelpress, 4-6 diffusion, 7-9 solid drift, 10-12 fluid drift, 13-21 vel_grad, 22-24 stress_accel
                float acc[32];
#pragma unroll
                for (int k = 0; k < 32; k++) acc[k] = 0.f;
                const float4 dpi = a.A.dpi[i];                                          // delpress of the home particle (:342-348)
                float sti[9];
#pragma unroll
                for (int k = 0; k < 9; k++) sti[k] = a.A.stress[(size_t)i * UNI_STRESS + k];
                const float pod2i = press_i / (densi * densi);
                // mass fractions and the mixed-phase guard's home half, cu:314-317 (RHO_0_SAND == RHO_0 == 9550)
                const float msf = solid_i * 9550 / (9550 * solid_i + 9550 * (fluid_i));
                const float mff = fluid_i * 9550 / (9550 * solid_i + 9550 * (fluid_i));
                const bool guard_i = (double)msf > 0.001 && (double)msf < 0.999 && (double)mff > 0.001 && (double)mff < 0.999 && !bi;
                for (int q = lane; q < qn; q += 32) {
                    const int c = S.q[q];
                    const int j = S.sj[c];
                    const float4 pj = S.sp[c], vj = a.A.velp[j], mj = a.A.mix[j];
                    const float rx = pi.x - pj.x, ry = pi.y - pj.y, rz = pi.z - pj.z;
                    const float ds = sqrtf(dist2(rx, ry, rz));
                    const float densj = fabsf(pj.w);
                    const bool bj = pj.w < 0.f;
                    const float solid_j = mj.x, fluid_j = mj.y, press_j = vj.w;
                    float w;                                                            // W(ds), cu:11-21
                    const float qq = ds * d.inv_h;
                    if (ds <= d.h_le) w = d.w_c * (1.f - 1.5f * qq * qq + 0.75f * qq * qq * qq);
                    else if (ds <= d.twoh_lt) { float tt = 2.f - qq; w = d.w_c * 0.25f * tt * tt * tt; }
                    else w = 0.f;
                    acc[0] += w * ((!bi && bj) ? 2.5f : 1.f);                           // cu:362 (mass == 1)
                    float dkx = 0.f, dky = 0.f, dkz = 0.f;
                    if (ds <= d.h_lt) {                                                 // support of dW
                        const float tt = d.hf - ds;
                        const float g = d.dw_c * tt * tt / ds;
                        dkx = g * rx; dky = g * ry; dkz = g * rz;                       // cu:296-298
                    }
                    const float vabx = vi.x - vj.x, vaby = vi.y - vj.y, vabz = vi.z - vj.z;
                    if (ds <= d.h_lt) {
                        const float dd = vabx * rx + vaby * ry + vabz * rz;             // cu:304
                        float s = 0.f;
                        if (dd < 0.f) {                                                 // cu:307
                            const float mu = dd / (ds * ds + d.eps);
                            const float hm = d.hf * mu;
                            const float bf = (!bi && bj) ? 1.f + (1.f + 3.f * fluid_i * fluid_i) * alpha_sb : 1.f;
                            s = ((solid_i * 9.f + 1.f) * (float)d.alpha_fluid) * (float)d.sound * (hm + d.visc_q * hm * hm) /
                                ((densi + densj) * 0.5f) * bf;
                        }
                        const float pp = press_j / (densj * densj) + pod2i + s;         // cu:310-312
                        acc[1] += pp * dkx; acc[2] += pp * dky; acc[3] += pp * dkz;
                        if (!bi && !bj) {
                            const float inv = 1.f / densj;
                            acc[4] += inv * dkx; acc[5] += inv * dky; acc[6] += inv * dkz;   // cu:364-366
                        }
                    }
                    // ---- drift velocities, cu:317-357 (the body term has a part without dW: every in-range pair counts) ----
                    if (guard_i && !bj) {
                        const float dk[3] = {dkx, dky, dkz}, vab[3] = {vabx, vaby, vabz}, dp[3] = {dpi.x, dpi.y, dpi.z};
                        const float coefs = solid_i * densi - (msf * solid_i * densi + mff * fluid_i * densi);
                        const float coeff = fluid_i * densi - (msf * solid_i * densi + mff * fluid_i * densi);
                        const float sps = solid_i * press_i - solid_j * press_j, fps = fluid_i * press_i - fluid_j * press_j;
#pragma unroll
                        for (int k = 0; k < 3; k++) {
                            const float sg = (solid_j - solid_i) * dk[k], fg = (fluid_j - fluid_i) * dk[k];
                            const float sbr = sg / solid_i - (msf * sg / solid_i + mff * fg / fluid_i);
                            const float fbr = fg / fluid_i - (mff * fg / fluid_i + msf * sg / solid_i);
                            const float slip_s = sps * dk[k] - msf * sps * dk[k] - mff * fps * dk[k];
                            const float slip_f = fps * dk[k] - msf * sps * dk[k] - mff * fps * dk[k];
                            const double second = (k == 2 ? d.gravity : 0.0) + (150.0 / (double)densi) * (double)dp[k] -
                                                  (double)(vi.x * dkx * vab[k]) - (double)(vi.y * dky * vab[k]) - (double)(vi.z * dkz * vab[k]);
                            const float sbody = (float)((double)coefs * second), fbody = (float)((double)coeff * second);
                            acc[7 + k] += (float)(UM_MIXPRESSURE * (double)(sbody + slip_s) - UM_MIXBROWNIAN * (double)sbr);
                            acc[10 + k] += (float)(UM_MIXPRESSURE * (double)(fbody + slip_f) - UM_MIXBROWNIAN * (double)fbr);
                        }
                    }
                    // ---- vel_grad and stress_accel, cu:368-381 (all carry dW) ----
                    if (ds <= d.h_lt) {
                        const float mixfactor = (!bj && !bi && solid_i > 0.f && solid_j > 0.f)
                                                    ? (float)(2.0 * (double)solid_i * (double)solid_j / ((double)solid_i + (double)solid_j + 0.01)) : 0.f;
                        const float dk[3] = {dkx, dky, dkz}, vab[3] = {vabx, vaby, vabz};
#pragma unroll
                        for (int pq = 0; pq < 9; pq++) acc[13 + pq] += (float)((double)(-mixfactor * vab[pq % 3] * dk[pq / 3]) * 1. / (double)densi);
                        const double d2i = (double)densi * (double)densi;
#pragma unroll
                        for (int pr = 0; pr < 3; pr++) {
                            const float sdk = sti[3 * pr] * dkx + sti[3 * pr + 1] * dky + sti[3 * pr + 2] * dkz;
                            acc[22 + pr] += (float)((double)(mixfactor * sdk) / d2i + (double)sdk / d2i);
                        }
                    }
                }
                const float tot = um_reduce_halving(acc, lane);          // lane k: sum number k
                {
                    // 0-3 -> sums, 4-6 -> sums2.xyz, 7-24 -> mixA[0..17]; lane 25 clears sums2.w (delfluid comes from pass B)
                    float *dst = lane < 4 ? reinterpret_cast<float *>(a.sums + i) + lane
                                 : lane < 7 ? reinterpret_cast<float *>(a.sums2 + i) + (lane - 4)
                                 : lane < 25 ? a.mixA + (size_t)i * UNI_MIXA + (lane - 7) : reinterpret_cast<float *>(a.sums2 + i) + 3;
                    if (lane < 25) *dst = t0 == 0 ? tot : *dst + tot;
                    else if (lane == 25 && t0 == 0) *dst = 0.f;
                }
            } else {
                // pass B, cu:383-401: 0-2 mixture_accel, 3 delsolid, 4 delfluid
                float acc[32];
#pragma unroll
                for (int k = 0; k < 32; k++) acc[k] = 0.f;
                const float *di = a.mixA + (size_t)i * UNI_MIXA;
                const float sdi[3] = {di[0], di[1], di[2]}, fdi[3] = {di[3], di[4], di[5]};
                for (int q = lane; q < qn; q += 32) {
                    const int c = S.q[q];
                    const int j = S.sj[c];
                    const float4 pj = S.sp[c], vj = a.A.velp[j], mj = a.A.mix[j];
                    const float rx = pi.x - pj.x, ry = pi.y - pj.y, rz = pi.z - pj.z;
                    const float ds = sqrtf(dist2(rx, ry, rz));
                    const float densj = fabsf(pj.w);
                    const bool bj = pj.w < 0.f;
                    const float solid_j = mj.x, fluid_j = mj.y;
                    const float tt = d.hf - ds;
                    const float g = d.dw_c * tt * tt / ds;
                    const float dkx = g * rx, dky = g * ry, dkz = g * rz;
                    const float vabx = vi.x - vj.x, vaby = vi.y - vj.y, vabz = vi.z - vj.z;
                    const float *dj = a.mixA + (size_t)j * UNI_MIXA;
                    const float sdj[3] = {dj[0], dj[1], dj[2]}, fdj[3] = {dj[3], dj[4], dj[5]};
                    const float ds2 = sdj[0] * dkx + sdj[1] * dky + sdj[2] * dkz, dsi = sdi[0] * dkx + sdi[1] * dky + sdi[2] * dkz;
                    const float df2 = fdj[0] * dkx + fdj[1] * dky + fdj[2] * dkz, dfi = fdi[0] * dkx + fdi[1] * dky + fdi[2] * dkz;
#pragma unroll
                    for (int k = 0; k < 3; k++)                                         // cu:391-398
                        acc[k] += -1 / densi / densj * (solid_j * densj * (solid_j * sdj[k] * ds2 + solid_i * sdi[k] * dsi) +
                                                         fluid_j * densj * (fluid_j * fdj[k] * df2 + fluid_i * fdi[k] * dfi));
                    const double nb = (!bj && !bi) ? 1.0 : 0.0;
                    const float dv = dkx * vabx + dky * vaby + dkz * vabz;
                    acc[3] += (float)(nb * -0.5 / (double)densj * (double)(solid_i + solid_j) * (double)dv +
                                      (double)((-(solid_i * sdi[0] + solid_j * sdj[0]) * dkx - (solid_i * sdi[1] + solid_j * sdj[1]) * dky -
                                                (solid_i * sdi[2] + solid_j * sdj[2]) * dkz) / densj));                   // cu:400
                    acc[4] += (float)(nb * -0.5 / (double)densj * (double)(fluid_i + fluid_j) * (double)dv +
                                      (double)((-(fluid_i * fdi[0] + fluid_j * fdj[0]) * dkx - (fluid_i * fdi[1] + fluid_j * fdj[1]) * dky -
                                                (fluid_i * fdi[2] + fluid_j * fdj[2]) * dkz) / densj));                   // cu:401
                }
                const float tot = um_reduce_halving(acc, lane);
                if (lane < 5) {
                    float *dst = a.mixB + (size_t)i * UNI_MIXB + lane;
                    *dst = t0 == 0 ? tot : *dst + tot;
                }
            }
            __syncwarp();
        }
        }       // tiles
        }       // groups of 32 home particles
        __syncwarp();
    }
}

cudaError_t fsg_launch_unidyn_mixed(const UniArgs &a, int pass, int sm_count, cudaStream_t s)
{
    if (a.n <= 0) return cudaSuccess;
    static FsgAttrOnce attr_once;
    if (attr_once.need()) {
        cudaFuncSetAttribute(k_pair_unidyn_mixed<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)UM_SMEM);
        cudaFuncSetAttribute(k_pair_unidyn_mixed<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)UM_SMEM);
    }
    int64_t blocks = ((int64_t)a.n + UM_WARPS - 1) / UM_WARPS;
    const int64_t maxb = (int64_t)sm_count * 5;           // 5 x 45 KB of shared memory per SM
    if (blocks > maxb) blocks = maxb;
    if (pass == 0) k_pair_unidyn_mixed<0><<<(unsigned)blocks, UM_WARPS * 32, UM_SMEM, s>>>(a);
    else k_pair_unidyn_mixed<1><<<(unsigned)blocks, UM_WARPS * 32, UM_SMEM, s>>>(a);
    return cudaGetLastError();
}
