// fsg_pair_v2.cu — pair sums of the base particle step for the uncapped configuration
// (neighbour_cap == 0 && bin_cap == 0: every particle of the 27 linear-offset bins is visited).
//
// Reference work it replaces: mykernel (FluidGPU.cu:119-285).  Particle::update / mykernel2
// (FluidGPU.cuh:270-304, FluidGPU.cu:404-432) run afterwards as the streaming kernel k_update.
//
// Structure (DESIGN.md "Kernels"): persistent blocks of 4 consumer warps + 1 producer warp.
//   * the PRODUCER walks a dynamic queue of occupied home bins.  Because bin id = ix*G^2 + iy*G + iz
//     (FluidGPU.cu:419) and the particles are sorted by it, the three z-adjacent bins of every
//     (dx,dy) column of the 27-bin neighbourhood are ONE contiguous particle run: a neighbourhood
//     is 9 runs.  Each run is moved into shared memory with one 1-D bulk async copy
//     (cp.async.bulk, the TMA engine) that signals an mbarrier; three stages are in flight, so the
//     latency of gathering a neighbourhood is hidden behind the sweep of the previous ones.
//   * the CONSUMER warps split the home particles of the bin (two per pass).  Lanes are
//     candidates: every 32-candidate chunk is read once from shared memory and tested against both
//     home particles; the W(r) term of the outer kernel branch (h < r <= 2h, FluidGPU.cu:15-16) is
//     evaluated branch-free in the sweep, candidates with r <= h (the support of dW,
//     FluidGPU.cu:35-43) are only marked in a per-lane bit mask.  After the sweep the marked
//     candidates are compacted into a per-warp queue and processed 32 at a time with all lanes busy
//     (pressure gradient + viscosity, FluidGPU.cu:238-279); per-particle sums come from a segmented
//     warp scan, so the result is deterministic (no atomics, fixed order).
//   * sums (newdens, newdelpress x/y/z) go to a float4 array that k_update consumes.
#include "fsg_device.cuh"
#include "fsg_pair_common.cuh"

#include <stdlib.h>

#define V2_CWARPS_MAX 8                 // consumer warps per block: template parameter CW (4, 6 or 8)
#define V2_TILE 512                     // staged candidates per stage
#define V2_NST 3                        // pipeline stages
#define V2_QCAP 160                     // near-pair queue entries per warp (drained at >= 32)
#define V2_GROUP 32                     // home particles per item
#define V2_BINS_PER_GRAB 16
#define V2_DEFAULT_CW 4
#ifndef V2_BPS4
#define V2_BPS4 7                       // resident blocks per SM for CW == 4
#endif

struct V2Stage {
    float4 sp[V2_TILE];                 // candidate (x, y, z, +-dens)
    float4 hp[V2_GROUP];                // home particles of the group
    int run_lo[12];                     // first tile slot of each staged run (V2_TILE where unused) ...
    int run_j[12];                      // ... and the global slot that tile slot holds
    int hs, gcount, ct, first, last, pad0, pad1, pad2;
};
struct V2Warp {
    unsigned q[V2_QCAP];
    float4 acc[8];                      // the warp's home particles of the group: dens, delpress x/y/z
};
template <int CW>
struct V2Smem {
    V2Stage st[V2_NST];
    V2Warp w[CW];
    unsigned long long full[V2_NST];
};

// "stage is free again" goes through hardware named barriers (ids 1..V2_NST): the consumer warps
// arrive without blocking, the producer warp blocks in bar.sync — no polling, no issue slots taken
// from the consumers while the producer is stages ahead.
template <int THREADS>
__device__ __forceinline__ void stage_free_arrive(int stage)
{
    // constant barrier ids, so that the kernel reserves V2_NST + 1 barriers and not all 16
    if (stage == 0) asm volatile("bar.arrive 1, %0;" ::"n"(THREADS) : "memory");
    else if (stage == 1) asm volatile("bar.arrive 2, %0;" ::"n"(THREADS) : "memory");
    else asm volatile("bar.arrive 3, %0;" ::"n"(THREADS) : "memory");
}
template <int THREADS>
__device__ __forceinline__ void stage_free_wait(int stage)
{
    if (stage == 0) asm volatile("bar.sync 1, %0;" ::"n"(THREADS) : "memory");
    else if (stage == 1) asm volatile("bar.sync 2, %0;" ::"n"(THREADS) : "memory");
    else asm volatile("bar.sync 3, %0;" ::"n"(THREADS) : "memory");
}
static_assert(V2_NST == 3, "stage_free_* name one barrier per stage");
// pressure / viscosity / inner-W terms of one r <= h pair (FluidGPU.cu:238-279)
__device__ __forceinline__ float4 v2_near_pair(const FsgDev &d, const float4 &pi, const float4 &vi, const float4 &pj,
                                               const float4 &vj)
{
    // approximate reciprocals (rsqrt.approx / div.approx, <= 2 ulp): the reference forms these terms in mixed
    // float / double, the 1e-5 parity bar is three orders of magnitude above either rounding
    float rx = pi.x - pj.x, ry = pi.y - pj.y, rz = pi.z - pj.z;
    float d2 = fmaf(rz, rz, fmaf(ry, ry, rx * rx));
    float inv = rsqrt_fast(d2);
    float ds = d2 * inv;
    float densi = fabsf(pi.w), densj = fabsf(pj.w);
    bool bi = pi.w < 0.f, bj = pj.w < 0.f;
    float q = ds * d.inv_h;
    float to = 2.f - q;
    // inner branch (FluidGPU.cu:13) minus the outer-branch value the sweep has already added for this pair
    float w = d.w_c * ((1.f - 1.5f * q * q + 0.75f * q * q * q) - 0.25f * to * to * to);
    float t = d.hf - ds;
    float g = d.dw_c * t * t * inv;                                            // FluidGPU.cu:37 (0 at r == h), / ds
    float vabx = vi.x - vj.x, vaby = vi.y - vj.y, vabz = vi.z - vj.z;
    float dd = vabx * rx + vaby * ry + vabz * rz;                              // :253
    float s = 0.f;
    const bool ib = !bi && bj;
    if (dd < 0.f) {                                                            // :255
        float hm = d.hf * __fdividef(dd, d2 + d.eps);
        float bf = ib ? 2.f * (1.f + (float)d.alpha_boundary) : 2.f;
        s = d.visc_c * (hm + d.visc_q * hm * hm) * __fdividef(bf, densi + densj);
    }
    float pp = __fdividef(vj.w, densj * densj) + __fdividef(vi.w, densi * densi) + s;   // :258-260
    float pg = pp * g;
    return make_float4(w * (ib ? 2.5f : 1.f), pg * rx, pg * ry, pg * rz);
}

struct V2Args {
    PairArgs a;
    float4 *sums;       // [n] newdens, newdelpress x, y, z
};

// one batch of <= 32 queued near pairs: lanes = pairs, segmented scan keyed by the home slot
template <int CW>
__device__ __forceinline__ void v2_drain_batch(const FsgDev &d, const V2Stage &S, V2Warp &W, const float4 *__restrict__ velp,
                                               int qh, int qn, int lane, int warp)
{
    const int e = qh + lane;
    const bool valid = e < qn;
    int key = -1 - lane;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (valid) {
        unsigned ent = W.q[e];
        key = (int)(ent >> 16);                                   // slot within the warp: pass * 2 + which
        int c = (int)(ent & 0xffffu);
        int k = 2 * CW * (key >> 1) + 2 * warp + (key & 1);       // home particle within the group (warp = rotated warp slot)
        int i = S.hs + k, j = 0;
#pragma unroll
        for (int r = 0; r < 9; r++) {                             // runs are staged in ascending slot order
            int lo = S.run_lo[r];
            if (c >= lo) j = S.run_j[r] + (c - lo);
        }
        v = v2_near_pair(d, S.hp[k], velp[i], S.sp[c], velp[j]);
    }
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int ku = __shfl_up_sync(FULL, key, o);
        float x0 = __shfl_up_sync(FULL, v.x, o), x1 = __shfl_up_sync(FULL, v.y, o);
        float x2 = __shfl_up_sync(FULL, v.z, o), x3 = __shfl_up_sync(FULL, v.w, o);
        if (lane >= o && ku == key) { v.x += x0; v.y += x1; v.z += x2; v.w += x3; }
    }
    int kn = __shfl_down_sync(FULL, key, 1);
    if (valid && (lane == 31 || kn != key)) {
        float4 c4 = W.acc[key];
        c4.x += v.x; c4.y += v.y; c4.z += v.z; c4.w += v.w;
        W.acc[key] = c4;
    }
    __syncwarp();
}

template <bool STATS, bool HASB, int CW, bool PK>
__global__ void __launch_bounds__((CW + 1) * 32, (CW == 4 ? V2_BPS4 : CW == 6 ? 5 : 4))
k_pair_v2(V2Args va)
{
    constexpr int PASSES = (V2_GROUP + 2 * CW - 1) / (2 * CW);
    extern __shared__ __align__(128) unsigned char s_raw[];
    V2Smem<CW> &SM = *reinterpret_cast<V2Smem<CW> *>(s_raw);
    const PairArgs &a = va.a;
    const FsgDev &d = a.d;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int nocc = *a.nocc;

    if (threadIdx.x == 0) {
#pragma unroll
        for (int s = 0; s < V2_NST; s++) {
            mbar_init(&SM.full[s], 1);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if (warp == CW) {
        // =========================== producer ===========================
        int it = 0;
        int grab = 0, grab_end = 0, mybin = 0;
        // table entries of the NEXT bin are requested one item ahead (pf_*), so that the only exposed
        // global latency is the queue grab every V2_BINS_PER_GRAB bins
        int pf_s0 = -1, pf_s1 = -1, pf_s2 = -1, pf_e0 = -1, pf_e1 = -1, pf_e2 = -1;
        bool pf_valid = false;
        auto prefetch = [&](int b) {
            pf_s0 = pf_s1 = pf_s2 = pf_e0 = pf_e1 = pf_e2 = -1;
            if (lane < 9) {
                const int c0 = b + (lane / 3 - 1) * d.G2 + (lane % 3 - 1) * d.G;
                if (c0 - 1 >= 0 && c0 - 1 < d.numcells && c0 != d.kx0 && c0 != d.kx1) { pf_s0 = a.start[c0 - 1]; pf_e0 = a.end[c0 - 1]; }
                if (c0 >= 0 && c0 < d.numcells) { pf_s1 = a.start[c0]; pf_e1 = a.end[c0]; }
                if (c0 + 1 >= 0 && c0 + 1 < d.numcells && c0 + 1 != d.kx0 && c0 + 1 != d.kx1) { pf_s2 = a.start[c0 + 1]; pf_e2 = a.end[c0 + 1]; }
            }
        };
        for (;;) {
            if (grab >= grab_end) {
                int m = 0;
                if (lane == 0) m = atomicAdd(a.work, V2_BINS_PER_GRAB);
                grab = __shfl_sync(FULL, m, 0);
                grab_end = min(grab + V2_BINS_PER_GRAB, nocc);
                if (grab + lane < grab_end) mybin = a.binlist[grab + lane];
                pf_valid = false;
            }
            if (grab >= nocc) {
                const int stage = it % V2_NST;
                if (it >= V2_NST) stage_free_wait<(CW + 1) * 32>(stage);
                if (lane == 0) {
                    SM.st[stage].gcount = -1;
                    mbar_arrive(&SM.full[stage]);
                }
                break;
            }
            const int slot = grab & (V2_BINS_PER_GRAB - 1);     // grabs are aligned to V2_BINS_PER_GRAB
            const int b = __shfl_sync(FULL, mybin, slot);
            if (!pf_valid) prefetch(b);
            // runs: lane r < 9 owns column (dx, dy) = (r / 3 - 1, r % 3 - 1), bins c0 - 1 .. c0 + 1
            int rs = 0, rp = 0;
            {
                int s = pf_s0 >= 0 ? pf_s0 : (pf_s1 >= 0 ? pf_s1 : pf_s2);
                int e = pf_s2 >= 0 ? pf_e2 : (pf_s1 >= 0 ? pf_e1 : pf_e0);
                if (s >= 0) { rs = s; rp = e - s + 1; }
            }
            const int hs = __shfl_sync(FULL, pf_s1, 4);
            const int hn = __shfl_sync(FULL, pf_e1, 4) - hs + 1;
            grab++;
            pf_valid = grab < grab_end;
            if (pf_valid) prefetch(__shfl_sync(FULL, mybin, grab & (V2_BINS_PER_GRAB - 1)));
            int incl = rp;
#pragma unroll
            for (int o = 1; o < 16; o <<= 1) {
                int t = __shfl_up_sync(FULL, incl, o);
                if (lane >= o) incl += t;
            }
            const int excl = incl - rp;
            const int C = __shfl_sync(FULL, incl, 8);
            for (int ig = 0; ig < hn; ig += V2_GROUP) {
                const int gcount = min(V2_GROUP, hn - ig);
                for (int t0 = 0; t0 < C; t0 += V2_TILE) {
                    const int ct = min(V2_TILE, C - t0);
                    const int stage = it % V2_NST;
                    V2Stage &S = SM.st[stage];
                    if (it >= V2_NST) stage_free_wait<(CW + 1) * 32>(stage);
                    // run table + padding + header (generic-proxy writes, published by the arrive below)
                    if (lane < 12) {
                        int lo = max(excl, t0), hi = min(excl + rp, t0 + ct);
                        bool used = lane < 9 && rp > 0 && hi > lo;
                        S.run_lo[lane] = used ? lo - t0 : V2_TILE;
                        S.run_j[lane] = used ? rs + (lo - excl) : 0;
                    }
                    const int cpad = (ct + 31) & ~31;
                    if (ct + lane < cpad) S.sp[ct + lane] = make_float4(1e30f, 1e30f, 1e30f, 0.f);
                    if (lane == 0) {
                        S.hs = hs + ig;
                        S.gcount = gcount;
                        S.ct = ct;
                        S.first = (t0 == 0);
                        S.last = (t0 + V2_TILE >= C);
                    }
                    __syncwarp();
                    if (lane == 0) mbar_arrive_expect_tx(&SM.full[stage], (unsigned)(ct + gcount) * 16u);
                    __syncwarp();
                    if (lane < 9 && rp > 0) {
                        int lo = max(excl, t0), hi = min(excl + rp, t0 + ct);
                        if (hi > lo) bulk_g2s(&S.sp[lo - t0], a.A.posd + rs + (lo - excl), (unsigned)(hi - lo) * 16u, &SM.full[stage]);
                    }
                    if (lane == 9) bulk_g2s(&S.hp[0], a.A.posd + hs + ig, (unsigned)gcount * 16u, &SM.full[stage]);
                    it++;
                }
            }
        }
        return;
    }

    // =========================== consumers ===========================
    V2Warp &W = SM.w[warp];
    const unsigned d2max_bits = __float_as_uint(d.d2_max), d2h_bits = __float_as_uint(d.d2_h);
    const unsigned lt_mask = (1u << lane) - 1u;
    const float w_outer = d.w_c * 0.25f;
    const float inv_h = d.inv_h;
    unsigned long long st_tested = 0, st_in = 0;
    int rot = 0;

    for (int it = 0;; it++) {
        const int stage = it % V2_NST;
        V2Stage &S = SM.st[stage];
        mbar_wait(&SM.full[stage], (it / V2_NST) & 1);
        const int gcount = S.gcount;
        if (gcount < 0) break;
        const int ct = S.ct;
        const int cpad = (ct + 31) & ~31;
        // the pairs of home particles are dealt to the warps round-robin; rotating the deal from item to item
        // keeps the warps of a block level over time (without it warp 0 always gets the odd pair out and the
        // others end up waiting for it at the stage release)
        if (S.first) rot = (rot + 1 == CW) ? 0 : rot + 1;       // (constant over the tiles of one group: acc[] persists across them)
        const int wr = (warp + rot) % CW;
        if (S.first && lane < 8) W.acc[lane] = make_float4(0.f, 0.f, 0.f, 0.f);
        __syncwarp();
        int qn = 0;
#pragma unroll 1
        for (int pass = 0; pass < PASSES; pass++) {
            const int k0 = 2 * CW * pass + 2 * wr;
            if (k0 >= gcount) break;
            const bool has1 = k0 + 1 < gcount;
            const float4 pi0 = S.hp[k0];
            float4 pi1 = S.hp[has1 ? k0 + 1 : k0];
            if (!has1) pi1.x = -1e30f;                               // nothing is in range of it
            const float ci0 = pi0.w < 0.f ? 0.f : 1.5f;              // float(!b_i)*BDENSFACTOR, FluidGPU.cu:276
            const float ci1 = pi1.w < 0.f ? 0.f : 1.5f;
            float w0 = 0.f, w1 = 0.f;
            unsigned m0 = 0, m1 = 0, bit = 1;
            int nin = 0;
            if (!STATS && PK) {
                // packed sweep: both home particles of the pass in the two halves of every FP32 instruction
                const f32x2 nhx = pk2(-pi0.x, -pi1.x), nhy = pk2(-pi0.y, -pi1.y), nhz = pk2(-pi0.z, -pi1.z);
                const f32x2 ninvh = pk2(-inv_h, -inv_h), two = pk2(2.f, 2.f), one = pk2(1.f, 1.f), ci = pk2(ci0, ci1);
                f32x2 wacc = pk2(0.f, 0.f);
#pragma unroll 4
                for (int c0 = 0; c0 < cpad; c0 += 32, bit <<= 1) {
                    const float4 pj = S.sp[c0 + lane];
                    const f32x2 rx = add2(pk2(pj.x, pj.x), nhx), ry = add2(pk2(pj.y, pj.y), nhy), rz = add2(pk2(pj.z, pj.z), nhz);
                    const f32x2 d2p = fma2(rz, rz, fma2(ry, ry, mul2(rx, rx)));
                    float d2a, d2b;
                    upk2(d2p, d2a, d2b);
                    const f32x2 r = mul2(d2p, pk2(rsqrt_fast(d2a), rsqrt_fast(d2b)));
                    float ta, tb;
                    upk2(fma2(r, ninvh, two), ta, tb);
                    ta = fmaxf(ta, 0.f);                                     // NaN (d2 == 0: the particle itself; padding) -> 0
                    tb = fmaxf(tb, 0.f);
                    // 2 - r/h >= 1  <=>  0 < r <= h: the near test costs one compare on a value the sweep has anyway.  A pair
                    // within rounding of r == h may land on either side: its near terms are O((h - r)^2) ~ 0 there.
                    if (ta >= 1.f) m0 |= bit;
                    if (tb >= 1.f) m1 |= bit;
                    const f32x2 tt = pk2(ta, tb);
                    f32x2 t3 = mul2(mul2(tt, tt), tt);
                    if (HASB) {
                        const float bjf = pj.w < 0.f ? 1.f : 0.f;
                        t3 = mul2(t3, fma2(ci, pk2(bjf, bjf), one));
                    }
                    wacc = add2(wacc, t3);
                }
                upk2(wacc, w0, w1);
            } else {
#pragma unroll 4
            for (int c0 = 0; c0 < cpad; c0 += 32, bit <<= 1) {
                const float4 pj = S.sp[c0 + lane];
                float bjf = 0.f;
                if (HASB) bjf = pj.w < 0.f ? 1.f : 0.f;
                {
                    float rx = pi0.x - pj.x, ry = pi0.y - pj.y, rz = pi0.z - pj.z;
                    float d2 = STATS ? dist2(rx, ry, rz) : fmaf(rz, rz, fmaf(ry, ry, rx * rx));
                    unsigned u = __float_as_uint(d2) - 1u;       // 0 < d2 <= thr  <=>  bits(d2) - 1 < bits(thr)
                    bool nearp = u < d2h_bits;                   // r <= h
                    float inv = rsqrt_fast(d2);
                    // (2 - r/h)^3 clamped at 0: zero beyond 2h (FluidGPU.cu:236), and zero for d2 == 0 (the particle
                    // itself) and for the padding sentinels, where 0 * inf = NaN and fmaxf returns the other operand.
                    // Pairs with r <= h also get this OUTER-branch value here; near_pair adds the difference to the inner
                    // branch (FluidGPU.cu:13), so the sweep needs no range test and no select.
                    float tt = fmaxf(fmaf(-d2 * inv, inv_h, 2.f), 0.f);
                    float t3 = tt * tt * tt;
                    if (HASB) t3 *= fmaf(ci0, bjf, 1.f);
                    w0 += t3;
                    if (nearp) m0 |= bit;
                    if (STATS) nin += __popc(__ballot_sync(FULL, u < d2max_bits));
                }
                {
                    float rx = pi1.x - pj.x, ry = pi1.y - pj.y, rz = pi1.z - pj.z;
                    float d2 = STATS ? dist2(rx, ry, rz) : fmaf(rz, rz, fmaf(ry, ry, rx * rx));
                    unsigned u = __float_as_uint(d2) - 1u;
                    bool nearp = u < d2h_bits;
                    float inv = rsqrt_fast(d2);
                    float tt = fmaxf(fmaf(-d2 * inv, inv_h, 2.f), 0.f);
                    float t3 = tt * tt * tt;
                    if (HASB) t3 *= fmaf(ci1, bjf, 1.f);
                    w1 += t3;
                    if (nearp) m1 |= bit;
                    if (STATS) nin += __popc(__ballot_sync(FULL, u < d2max_bits));
                }
            }
            }
            if (STATS && lane == 0) { st_tested += (unsigned long long)ct * (has1 ? 2 : 1); st_in += nin; }
#pragma unroll
            for (int o = 16; o; o >>= 1) {
                w0 += __shfl_xor_sync(FULL, w0, o);
                w1 += __shfl_xor_sync(FULL, w1, o);
            }
            if (lane < 2) W.acc[2 * pass + lane].x += (lane ? w1 : w0) * w_outer;
            // ---- compact the marked near candidates into the queue ----
#pragma unroll
            for (int which = 0; which < 2; which++) {
                unsigned m = which ? m1 : m0;
                const unsigned tag = (unsigned)(2 * pass + which) << 16;
                for (;;) {
                    unsigned any = __ballot_sync(FULL, m != 0);
                    if (!any) break;
                    if (m) {
                        int c = (__ffs(m) - 1) * 32 + lane;
                        W.q[qn + __popc(any & lt_mask)] = tag | (unsigned)c;
                        m &= m - 1;
                    }
                    qn += __popc(any);
                    if (qn >= V2_QCAP - 32) {                      // keep room for one more round
                        __syncwarp();
                        int qh = 0;
                        while (qn - qh >= 32) { v2_drain_batch<CW>(d, S, W, a.A.velp, qh, qn, lane, wr); qh += 32; }
                        int left = qn - qh;
                        unsigned ent = 0;
                        if (lane < left) ent = W.q[qh + lane];
                        __syncwarp();
                        if (lane < left) W.q[lane] = ent;
                        qn = left;
                        __syncwarp();
                    }
                }
            }
        }
        // ---- drain the queue (entries refer to this stage's candidates) ----
        __syncwarp();
        for (int qh = 0; qh < qn; qh += 32) v2_drain_batch<CW>(d, S, W, a.A.velp, qh, qn, lane, wr);
        const int last = S.last, hs = S.hs;
        __syncwarp();
        stage_free_arrive<(CW + 1) * 32>(stage);
        if (last && lane < 2 * PASSES) {
            int k = 2 * CW * (lane >> 1) + 2 * wr + (lane & 1);
            if (k < gcount) va.sums[hs + k] = W.acc[lane];
        }
        __syncwarp();
    }
    if (STATS && lane == 0) {
        atomicAdd(a.stats + 0, st_tested);
        atomicAdd(a.stats + 1, st_in);
    }
}

// ------------------------------------------------------------------------------------------------
// k_update — Particle::update + the tail of mykernel2 (FluidGPU.cuh:270-304, FluidGPU.cu:419-425)
// for every sorted slot: sums (+ accumulators carried in from the upload) -> EOS, integration, new
// bin id.  Streaming: reads 64 + 16 B, writes 64 + 4 B per particle.  Particles parked outside the
// bin grid (key == numcells) are copied through unchanged; ghosts of a slab context are dropped.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_update(FsgDev d, int n, const int *__restrict__ keysA, FsgState A, FsgState B, int *__restrict__ keysB,
         const float4 *__restrict__ sums, const float4 *__restrict__ carry, int part, int *violation)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int key = keysA[i];
    if (part) {      // 1: the slab's boundary slots (boundary bins, ghosts, parked, dead), 2: the interior slots
        const int ixp = key / d.G2;
        const bool interior = key < d.numcells && ixp >= d.bx0 && ixp < d.bx1;
        if (interior != (part == 2)) return;
    }
    float4 pd = A.posd[i], vp = A.velp[i], af = A.accf[i], dpi = A.dpi[i];
    if (key < d.numcells) {
        const int ix = key / d.G2;
        if (ix < d.x0 || ix >= d.x1) {      // ghost copy of a neighbour slab's particle: drop it
            keysB[i] = d.dead;
            return;
        }
        float4 s = sums[i];
        if (carry) { float4 cy = carry[i]; s.x += cy.x; s.y += cy.y; s.z += cy.z; s.w += cy.w; }
        particle_update(d, pd, vp, af, dpi, s.x, s.y, s.z, s.w, key);
        // slab contexts: the one-layer ghost band (and the pack's two-layer region) assume less than one bin layer per step
        if (violation && key < d.numcells && abs(key / d.G2 - ix) > 1) atomicOr(violation, 1);
    }
    B.posd[i] = pd;
    B.velp[i] = vp;
    B.accf[i] = af;
    B.dpi[i] = dpi;
    keysB[i] = key;
}

template <int CW, bool PK>
static cudaError_t launch_pair_v2_cw(const V2Args &va, bool stats, bool has_boundary, int sm_count, int blocks_per_sm, cudaStream_t s)
{
    static FsgAttrOnce attr_once;
    const int smem = (int)sizeof(V2Smem<CW>);
    if (attr_once.need()) {
        cudaFuncSetAttribute(k_pair_v2<false, false, CW, PK>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        cudaFuncSetAttribute(k_pair_v2<false, true, CW, PK>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        cudaFuncSetAttribute(k_pair_v2<true, false, CW, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        cudaFuncSetAttribute(k_pair_v2<true, true, CW, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    }
    int per_sm = CW == 4 ? V2_BPS4 : CW == 6 ? 5 : 4;
    if (blocks_per_sm > 0 && blocks_per_sm < per_sm) per_sm = blocks_per_sm;
    int64_t blocks = ((int64_t)va.a.n + 2 * CW - 1) / (2 * CW);
    int64_t maxb = (int64_t)sm_count * per_sm;
    if (blocks > maxb) blocks = maxb;
    if (blocks < 1) blocks = 1;
    const unsigned threads = (CW + 1) * 32;
    if (stats) {     // the counting build keeps the scalar sweep (its range test uses the reference's unfused distance)
        if (has_boundary) k_pair_v2<true, true, CW, false><<<(unsigned)blocks, threads, smem, s>>>(va);
        else k_pair_v2<true, false, CW, false><<<(unsigned)blocks, threads, smem, s>>>(va);
    } else {
        if (has_boundary) k_pair_v2<false, true, CW, PK><<<(unsigned)blocks, threads, smem, s>>>(va);
        else k_pair_v2<false, false, CW, PK><<<(unsigned)blocks, threads, smem, s>>>(va);
    }
    return cudaGetLastError();
}

cudaError_t fsg_launch_pair_v2(const PairArgs &a, float4 *sums, bool stats, bool has_boundary, int sm_count, int blocks_per_sm,
                               cudaStream_t s)
{
    V2Args va;
    va.a = a;
    va.sums = sums;
    static int cw = 0;
    if (!cw) {                                        // FSG_PAIR_CW = 4 | 6 | 8: consumer warps per block (tuning knob)
        const char *e = getenv("FSG_PAIR_CW");
        cw = e ? atoi(e) : V2_DEFAULT_CW;
        if (cw != 4 && cw != 6 && cw != 8) cw = V2_DEFAULT_CW;
    }
    static int pk = -1;
    if (pk < 0) {                                     // FSG_PAIR_PK = 0: scalar sweep (A/B knob; default packed FFMA2 sweep)
        const char *e = getenv("FSG_PAIR_PK");
        pk = e ? (atoi(e) != 0) : 1;
    }
    if (cw == 6) return launch_pair_v2_cw<6, true>(va, stats, has_boundary, sm_count, blocks_per_sm, s);
    if (cw == 8) return launch_pair_v2_cw<8, true>(va, stats, has_boundary, sm_count, blocks_per_sm, s);
    if (!pk) return launch_pair_v2_cw<4, false>(va, stats, has_boundary, sm_count, blocks_per_sm, s);
    return launch_pair_v2_cw<4, true>(va, stats, has_boundary, sm_count, blocks_per_sm, s);
}

cudaError_t fsg_launch_update(const FsgDev &d, int64_t n, const int *keysA, FsgState A, FsgState B, int *keysB,
                              const float4 *sums, const float4 *carry, int part, int *violation, cudaStream_t s)
{
    if (n <= 0) return cudaSuccess;
    k_update<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(d, (int)n, keysA, A, B, keysB, sums, carry, part, violation);
    return cudaGetLastError();
}
