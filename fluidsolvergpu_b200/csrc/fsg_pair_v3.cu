// fsg_pair_v3.cu — SYMMETRIC pair sums of the base particle step for the uncapped configuration.
//
// Reference work it replaces: mykernel (FluidGPU.cu:119-285), whose 27-bin sweep evaluates every pair twice
// (once from each side).  Here a pair of particles in two different bins is evaluated ONCE, by the bin with
// the lower id, and added to both particles; only pairs inside one bin are still seen from both sides.  A
// home particle therefore meets ~14 bins of candidates instead of 27.
//
//   * the neighbourhood of bin b by linear offsets (FluidGPU.cu:124-126) is symmetric (offset <-> -offset), so
//     "the 13 bins with a larger id" + the bin itself is exactly one half of every pair relation, wrap-around
//     of the linear offsets included.  With bin id = ix*G^2 + iy*G + iz the forward half is 5 contiguous runs of
//     sorted particles: [b, b+1], [b+G-1, b+G+1] and the three runs of layer ix+1.
//   * every WARP is an independent worker: it takes home bins from the dynamic queue, stages the runs of the
//     next work item with 1-D bulk async copies (TMA engine) into its own two-stage shared-memory ring while it
//     computes the current one — no producer warp, no block-level barrier, nothing shared between warps.
//   * lanes are candidates.  The sweep (packed FP32, both home particles of a pass per instruction) adds the
//     outer-branch W(r) (FluidGPU.cu:15-16) to the home particles (warp reduction per pass) AND to the lane's
//     candidate (a register per 32-candidate chunk, no communication at all).  Pairs with r <= h are queued and
//     processed 32 at a time as in fsg_pair_v2.cu (FluidGPU.cu:238-279); both sides of such a pair are added with
//     red.global.add.v4.f32.
//   * candidate-side and home-side sums reach `sums[]` through float reductions in L2 (measured 350 G float4
//     reductions/s, tools/micro/red_rate.cu), so `sums[]` is cleared before the launch and the order of the
//     additions is not fixed: results agree with the gather kernel to rounding (~1e-7), not bit for bit.
//     fsg_config.pair_mode = 1 selects the deterministic gather kernel (fsg_pair_v2.cu) instead.
//   * slab contexts: the ghost layer below the slab (ix == x0 - 1) is walked as home bins too, restricted to its
//     runs in layer x0 — those are the pairs (ghost, owned) whose lower bin is the ghost's.
#include "fsg_device.cuh"
#include "fsg_pair_common.cuh"

#ifndef V3_WARPS
#define V3_WARPS 4                      // independent warps per block
#define V3_BPS 5                        // resident blocks per SM
#endif
#define V3_TILE 256                     // staged candidates per stage
#define V3_CH (V3_TILE / 32)            // 32-candidate chunks per tile = candidate accumulators per lane
#define V3_GROUP 32                     // home particles per item (lane k owns the row sums of home particle k)
#define V3_QCAP 160                     // near-pair queue entries
#define V3_GRAB 8                       // home bins per queue grab

struct V3Stage {
    float4 sp[V3_TILE];                 // candidate (x, y, z, +-dens)
    float4 hp[V3_GROUP];                // home particles of the group
};
struct V3Warp {
    V3Stage st[2];
    unsigned q[V3_QCAP];
    int run_lo[2][8];                   // first tile slot of each staged run (V3_TILE where unused) ...
    int run_j[2][8];                    // ... and the global slot that tile slot holds
    unsigned long long full[2];
};
struct V3Sub {                          // one work item: a group of home particles x a tile of candidates
    int hs, gcount, t0, ct, flags;      // flags: 1 = first tile of the group, 2 = last tile of the group
    int hnlim;                          // candidates at position t0 + c < hnlim are the home bin's own particles: home side only
};

__device__ __forceinline__ void red_add_v4(float4 *p, float4 v)
{
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void red_add_f32(float *p, float v)
{
    asm volatile("red.global.add.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory");
}

struct V3Near {                         // the constants of the near-pair terms, passed by value to the out-of-line batch
    float inv_h, w_c, hf, dw_c, eps, visc_c, visc_q, ab;
};
// both sides of one r <= h pair (FluidGPU.cu:238-279): hv for the home particle i, cv for the candidate j
__device__ __forceinline__ void v3_near_pair(const V3Near &d, const float4 &pi, const float4 &vi, const float4 &pj, const float4 &vj,
                                             float4 &hv, float4 &cv)
{
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
    float g = d.dw_c * t * t * inv;                                            // FluidGPU.cu:37, / ds
    float vabx = vi.x - vj.x, vaby = vi.y - vj.y, vabz = vi.z - vj.z;
    float dd = vabx * rx + vaby * ry + vabz * rz;                              // :253 (the same from either side)
    float s = 0.f;
    if (dd < 0.f) {                                                            // :255
        float hm = d.hf * __fdividef(dd, d2 + d.eps);
        s = d.visc_c * (hm + d.visc_q * hm * hm) * __fdividef(2.f, densi + densj);
    }
    const bool ib = !bi && bj, jb = !bj && bi;
    float pp = __fdividef(vj.w, densj * densj) + __fdividef(vi.w, densi * densi);   // :258-260
    float pgi = (pp + (ib ? s * d.ab : s)) * g;
    float pgj = -(pp + (jb ? s * d.ab : s)) * g;
    hv = make_float4(w * (ib ? 2.5f : 1.f), pgi * rx, pgi * ry, pgi * rz);
    cv = make_float4(w * (jb ? 2.5f : 1.f), pgj * rx, pgj * ry, pgj * rz);
}

struct V3Args {
    PairArgs a;
    float4 *sums;       // [n] newdens, newdelpress x, y, z — cleared by the launcher
};

__device__ __forceinline__ int v3_global_slot(const V3Warp &W, int stage, int c)
{
    int j = 0;
#pragma unroll
    for (int r = 0; r < 5; r++) {                               // runs are staged in ascending slot order
        int lo = W.run_lo[stage][r];
        if (c >= lo) j = W.run_j[stage][r] + (c - lo);
    }
    return j;
}

// one batch of <= 32 queued near pairs: lanes = pairs, both sides reduced straight into sums[]
// (queue entries are RAW marks — bits 0-4: 8 * pass + position in the mark word, bit 5: which home particle of the pass, bits 8-12:
// the lane that held the mark, bits 16+: first home particle of the batch — decoded here, 32 at a time, instead of once per mark in
// the append loop, where only the lanes that hold a mark are active)
__device__ __forceinline__ void v3_drain_batch(const V3Near &d, const V3Warp &W, int stage, int hs, int t0, int hnlim, int nch,
                                               const float4 *__restrict__ velp, float4 *__restrict__ sums, int qh, int qn, int lane)
{
    const int e = qh + lane;
    if (e < qn) {
        const V3Stage &S = W.st[stage];
        const unsigned ent = W.q[e];
        const int p5 = (int)(ent & 31u);
        const int k = (int)(ent >> 16) + 2 * (p5 >> 3) + (int)((ent >> 5) & 1u);
        const int c = (nch - 1 - (p5 & 7)) * 32 + (int)((ent >> 8) & 31u);
        const int i = hs + k, j = v3_global_slot(W, stage, c);
        const float *P = reinterpret_cast<const float *>(&S.hp[0]) + (k >> 1) * 8 + (k & 1);     // packed homes, see the main loop
        const float4 pi = make_float4(-P[0], -P[2], -P[4], P[6]);
        float4 hv, cv;
        v3_near_pair(d, pi, velp[i], S.sp[c], velp[j], hv, cv);
        red_add_v4(sums + i, hv);
        if (t0 + c >= hnlim) red_add_v4(sums + j, cv);
    }
}

// One batch of up to four passes (two home particles each, packed) over the NCH 32-candidate chunks of the tile.  The chunk
// loop is straight-line code per chunk count: the candidate accumulators cw[] stay in fixed registers across the passes and
// the chunks overlap in the pipeline.  M0 / M1 collect the near marks of the batch: bit 8 * pass + (NCH - 1 - chunk).
template <bool HASB> __device__ __forceinline__ constexpr float V3_FIX() { return HASB ? 4194304.f : 8388608.f; }      // 2^22, 2^23

template <int NCH, bool HASB>
__device__ __forceinline__ void v3_batch(const float4 *__restrict__ hp, const float4 *__restrict__ sp, const int kb, const int gcount,
                                         const int lane, const f32x2 ninvh, f32x2 (&cw)[V3_CH], float &hrow, unsigned &M0, unsigned &M1)
{
    const f32x2 one = pk2(1.f, 1.f), mhalf = pk2(-0.5f, -0.5f);
    float nih0, nih1;
    upk2(ninvh, nih0, nih1);                                 // -1 / (2h)
#pragma unroll 1
    for (int pp = 0; pp < 4; pp++) {
        const int k0 = kb + 2 * pp;
        if (k0 >= gcount) break;
        const float4 ha = hp[k0], hb = hp[k0 + 1];             // packed by the caller: (-x0, -x1, -y0, -y1), (-z0, -z1, w0, w1)
        const f32x2 nhx = pk2(ha.x, ha.y), nhy = pk2(ha.z, ha.w), nhz = pk2(hb.x, hb.y);
        // float(!b_i)*BDENSFACTOR (FluidGPU.cu:276) for the home side, the homes' boundary flags for the candidate side
        const f32x2 ci = pk2(hb.z < 0.f ? 0.f : 1.5f, hb.w < 0.f ? 0.f : 1.5f);
        const f32x2 bi = pk2(hb.z < 0.f ? 1.f : 0.f, hb.w < 0.f ? 1.f : 0.f);
        f32x2 wacc = pk2(0.f, 0.f);
        unsigned m0 = 0, m1 = 0;                             // bit (NCH - 1 - chunk) SET = the candidate is NOT within h
#pragma unroll
        for (int k = 0; k < NCH; k++) {
            const float4 pj = sp[k * 32];
            const f32x2 rx = add2(pk2(pj.x, pj.x), nhx), ry = add2(pk2(pj.y, pj.y), nhy), rz = add2(pk2(pj.z, pj.z), nhz);
            const f32x2 d2p = fma2(rz, rz, fma2(ry, ry, mul2(rx, rx)));
            float d2a, d2b;
            upk2(d2p, d2a, d2b);
            const f32x2 r = mul2(d2p, pk2(rsqrt_fast(d2a), rsqrt_fast(d2b)));
            float ra, rb;
            upk2(r, ra, rb);
            // u = 1 - r/(2h) saturated to [0, 1] by the FMA itself: 0 beyond 2h (FluidGPU.cu:236) and for d2 == 0 (r = 0 * inf = NaN
            // -> +0, the particle itself); the padding sits 1e15 away.  (2 - r/h)^3 = 8 u^3 — the 8 is folded into w_outer.
            // Pairs with r <= h also get this OUTER-branch value; the near-pair pass adds the difference to the inner
            // branch (FluidGPU.cu:13).  No compare, no min/max: the sign of u - 1/2 is shifted into the near marks.
            const float ua = fma_sat(ra, nih0, 1.f), ub = fma_sat(rb, nih1, 1.f);
            const f32x2 uu = pk2(ua, ub);
            float va, vb;
            upk2(add2(uu, mhalf), va, vb);                  // < 0  <=>  r > h
            m0 = __funnelshift_l(__float_as_uint(va), m0, 1);
            m1 = __funnelshift_l(__float_as_uint(vb), m1, 1);
            const f32x2 t3 = mul2(mul2(uu, uu), uu);
            if (HASB) {
                const float bjf = pj.w < 0.f ? 1.f : 0.f, cjf = pj.w < 0.f ? 0.f : 1.5f;
                wacc = fma2(t3, fma2(ci, pk2(bjf, bjf), one), wacc);
                cw[k] = fma2(t3, fma2(bi, pk2(cjf, cjf), one), cw[k]);
            } else {
                wacc = add2(wacc, t3);
                cw[k] = add2(cw[k], t3);
            }
        }
        m0 = ~m0 & ((1u << NCH) - 1u);                       // bit (NCH - 1 - chunk) set = within h
        m1 = ~m1 & ((1u << NCH) - 1u);
        M0 |= m0 << (8 * pp);
        M1 |= m1 << (8 * pp);
        // Home side: the 32 lanes' partial sums of the two home particles are added with the warp-wide INTEGER reduction (one REDUX
        // each instead of five shuffle + add rounds): fixed point with 23 (22 with boundary factors) fraction bits.  A pass adds at
        // most 256 terms u^3 <= 1 (x 2.5 with boundary factors), so the 32-bit sum cannot overflow: 256 * 2^23 = 2^31,
        // 640 * 2^22 < 2^32; the rounding (<= 6e-8 per lane and pass, absolute) is below the float rounding of sums of this size.
        // The running total over passes and tiles stays a float.  Order independent: this side of the sums is deterministic.
        {
            float w0, w1;
            upk2(mul2(wacc, pk2(V3_FIX<HASB>(), V3_FIX<HASB>())), w0, w1);
            const unsigned s0 = __reduce_add_sync(FULL, __float2uint_rn(w0)), s1 = __reduce_add_sync(FULL, __float2uint_rn(w1));
            if (lane == k0) hrow += (float)s0;
            if (lane == k0 + 1) hrow += (float)s1;
        }
    }
}

// (the variant that evaluates the boundary factors needs a few more registers: one resident block less instead of spills)
template <bool HASB>
__global__ void __launch_bounds__(V3_WARPS * 32, HASB ? V3_BPS - 1 : V3_BPS)
k_pair_v3(V3Args va)
{
    extern __shared__ __align__(128) unsigned char s_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    V3Warp &W = reinterpret_cast<V3Warp *>(s_raw)[warp];
    const PairArgs &a = va.a;
    const FsgDev &d = a.d;
    const int nocc = *a.nocc;
    float4 *__restrict__ sums = va.sums;

    if (lane == 0) {
        mbar_init(&W.full[0], 1);
        mbar_init(&W.full[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();

    // ---------------- work-item iterator (warp-uniform unless noted) ----------------
    int grab = 0, grab_end = 0, mybin = 0;                    // mybin: lane l holds bin grab0 + l
    int pf_s0 = -1, pf_s1 = -1, pf_s2 = -1, pf_e0 = -1, pf_e1 = -1, pf_e2 = -1;   // lane r < 5: table entries of run r of the NEXT bin
    bool pf_valid = false;
    int rs = 0, rp = 0, excl = 0;                             // lane r < 5: run r of the current bin (first slot, length, prefix)
    int C = 0, hn = 0, hs0 = 0, hnlim = 0, ig = 0, t0 = 0;
    bool have = false;
    const int ghost_below = d.x0 * d.G2;                      // bins with a smaller id lie in the ghost layer x0 - 1
    // run r of home bin b: lane 0 -> bins [b, b+1]; lane 1 -> [b+G-1, b+G+1]; lanes 2..4 -> the three runs of layer ix+1
    auto prefetch = [&](int b) {
        pf_s0 = pf_s1 = pf_s2 = pf_e0 = pf_e1 = pf_e2 = -1;
        if (lane < 5) {
            const int c0 = b + (lane == 0 ? 0 : lane == 1 ? d.G : d.G2 + (lane - 3) * d.G);
            if (lane != 0 && c0 - 1 >= 0 && c0 - 1 < d.numcells && c0 != d.kx0 && c0 != d.kx1) { pf_s0 = a.start[c0 - 1]; pf_e0 = a.end[c0 - 1]; }
            if (c0 >= 0 && c0 < d.numcells) { pf_s1 = a.start[c0]; pf_e1 = a.end[c0]; }
            if (c0 + 1 >= 0 && c0 + 1 < d.numcells && c0 + 1 != d.kx0 && c0 + 1 != d.kx1) { pf_s2 = a.start[c0 + 1]; pf_e2 = a.end[c0 + 1]; }
        }
    };
    auto next_sub = [&](V3Sub &sub) -> bool {
        for (;;) {
            if (!have) {
                if (grab >= grab_end) {
                    int m = 0;
                    if (lane == 0) m = atomicAdd(a.work, V3_GRAB);
                    grab = __shfl_sync(FULL, m, 0);
                    grab_end = min(grab + V3_GRAB, nocc);
                    if (grab >= nocc) return false;
                    if (grab + lane < grab_end) mybin = a.binlist[grab + lane];
                    pf_valid = false;
                }
                const int b = __shfl_sync(FULL, mybin, grab & (V3_GRAB - 1));     // grabs are aligned to V3_GRAB
                if (!pf_valid) prefetch(b);
                rs = 0;
                rp = 0;
                {
                    int s = pf_s0 >= 0 ? pf_s0 : (pf_s1 >= 0 ? pf_s1 : pf_s2);
                    int e = pf_s2 >= 0 ? pf_e2 : (pf_s1 >= 0 ? pf_e1 : pf_e0);
                    if (s >= 0) { rs = s; rp = e - s + 1; }
                }
                hs0 = __shfl_sync(FULL, pf_s1, 0);
                hn = __shfl_sync(FULL, pf_e1, 0) - hs0 + 1;
                const bool ghost = b < ghost_below;             // ghost layer below the slab: only its pairs with layer x0
                if (ghost && lane < 2) rp = 0;
                hnlim = ghost ? 0 : hn;
                grab++;
                pf_valid = grab < grab_end;
                if (pf_valid) prefetch(__shfl_sync(FULL, mybin, grab & (V3_GRAB - 1)));
                int incl = rp;
#pragma unroll
                for (int o = 1; o < 8; o <<= 1) {
                    int t = __shfl_up_sync(FULL, incl, o);
                    if (lane >= o) incl += t;
                }
                excl = incl - rp;
                C = __shfl_sync(FULL, incl, 4);
                ig = 0;
                t0 = 0;
                have = C > 0 && hn > 0;
                if (!have) continue;
            }
            sub.hs = hs0 + ig;
            sub.gcount = min(V3_GROUP, hn - ig);
            sub.t0 = t0;
            sub.ct = min(V3_TILE, C - t0);
            sub.flags = (t0 == 0 ? 1 : 0) | (t0 + V3_TILE >= C ? 2 : 0);
            sub.hnlim = hnlim;
            t0 += V3_TILE;
            if (t0 >= C) {
                t0 = 0;
                ig += V3_GROUP;
                if (ig >= hn) have = false;
            }
            return true;
        }
    };
    // stage the candidates and home particles of `sub` (must be called right after next_sub produced it)
    auto issue = [&](const V3Sub &sub, int stage) {
        V3Stage &S = W.st[stage];
        const int lo = max(excl, sub.t0), hi = min(excl + rp, sub.t0 + sub.ct);
        const bool used = lane < 5 && rp > 0 && hi > lo;
        // the runs this tile holds, COMPACTED (ascending tile slot): entry q = (first tile slot, global slot of it); unused entries
        // hold V3_TILE, so "the last entry with run_lo <= c" is candidate c's run and the table can be walked monotonically
        const unsigned um = __ballot_sync(FULL, used);
        if (lane < 8) { W.run_lo[stage][lane] = V3_TILE; W.run_j[stage][lane] = 0; }
        __syncwarp();
        if (used) {
            const int q = __popc(um & ((1u << lane) - 1u));
            W.run_lo[stage][q] = lo - sub.t0;
            W.run_j[stage][q] = rs + (lo - excl);
        }
        // padding up to the chunk count the sweep runs (whole chunks, at least 4: the sweep has no variant below)
        const int cpad = max((sub.ct + 31) & ~31, 128);
#pragma unroll 1
        for (int c = sub.ct + lane; c < cpad; c += 32) S.sp[c] = make_float4(1e15f, 1e15f, 1e15f, 0.f);      // (finite d2: no NaN in the sweep)
        // the stage was read through the generic proxy two items ago; order those reads before the async writes
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) mbar_arrive_expect_tx(&W.full[stage], (unsigned)(sub.ct + sub.gcount) * 16u);
        __syncwarp();
        if (used) bulk_g2s(&S.sp[lo - sub.t0], a.A.posd + rs + (lo - excl), (unsigned)(hi - lo) * 16u, &W.full[stage]);
        if (lane == 5) bulk_g2s(&S.hp[0], a.A.posd + sub.hs, (unsigned)sub.gcount * 16u, &W.full[stage]);
    };

    const float w_outer = d.w_c * 2.f;                         // w_c / 4 * 8: the sweep sums u^3 with u = (2 - r/h) / 2
    const float inv_h = d.inv_h;
    const f32x2 ninvh = pk2(-0.5f * inv_h, -0.5f * inv_h);
    V3Near nc;
    nc.inv_h = d.inv_h; nc.w_c = d.w_c; nc.hf = d.hf; nc.dw_c = d.dw_c; nc.eps = d.eps; nc.visc_c = d.visc_c; nc.visc_q = d.visc_q;
    nc.ab = 1.f + (float)d.alpha_boundary;
    float hrow = 0.f;                                          // lane k: sum of outer-branch terms of home particle k of the group (x 2^23 / 2^22)

    V3Sub cur, nxt;
    cur.hs = cur.gcount = cur.t0 = cur.ct = cur.flags = cur.hnlim = 0;
    // software pipeline with ONE copy of the iterator / staging code: iteration n stages item n + 1, then computes item n
    // (iteration -1 only stages)
    for (int n = -1;; n++) {
        const int stage = n & 1;
        const bool hnx = next_sub(nxt);
        if (hnx) issue(nxt, stage ^ 1);
        if (n < 0) {
            if (!hnx) break;
            cur = nxt;
            continue;
        }
        mbar_wait(&W.full[stage], (n >> 1) & 1);

        V3Stage &S = W.st[stage];
        const int gcount = cur.gcount, ct = cur.ct;
        {
            // the home particles, re-packed in place the way a pass reads them: pass p (home particles 2p, 2p + 1) =
            // {(-x0, -x1, -y0, -y1), (-z0, -z1, w0, w1)} — two 16-byte loads give the packed operands, no register shuffling per pass.
            // A home particle that does not exist sits 1e15 away (nothing is in range of it).
            float4 me = S.hp[lane];
            __syncwarp();
            if (lane >= gcount) me = make_float4(-1e15f, -1e15f, -1e15f, 0.f);
            float *P = reinterpret_cast<float *>(&S.hp[0]) + (lane >> 1) * 8 + (lane & 1);
            P[0] = -me.x; P[2] = -me.y; P[4] = -me.z; P[6] = me.w;
            __syncwarp();
        }
        const int nch = max((ct + 31) >> 5, 4);
        if (cur.flags & 1) hrow = 0.f;
        f32x2 cw[V3_CH];                                       // candidate of chunk k of this lane: its sum over the home particles
#pragma unroll
        for (int k = 0; k < V3_CH; k++) cw[k] = pk2(0.f, 0.f);
        int qn = 0;
        const float4 *__restrict__ spl = &S.sp[lane], *__restrict__ hpp = &S.hp[0];
#pragma unroll 1
        for (int kb = 0; kb < gcount; kb += 8) {                     // batches of four passes: one compaction per batch
            unsigned M0 = 0, M1 = 0;                                 // bit 8 * pass + chunk: candidate (chunk, lane) is within h
            switch (nch) {
            case 4: v3_batch<4, HASB>(hpp, spl, kb, gcount, lane, ninvh, cw, hrow, M0, M1); break;
            case 5: v3_batch<5, HASB>(hpp, spl, kb, gcount, lane, ninvh, cw, hrow, M0, M1); break;
            case 6: v3_batch<6, HASB>(hpp, spl, kb, gcount, lane, ninvh, cw, hrow, M0, M1); break;
#if V3_CH > 7
            case 7: v3_batch<7, HASB>(hpp, spl, kb, gcount, lane, ninvh, cw, hrow, M0, M1); break;
#endif
            default: v3_batch<V3_CH, HASB>(hpp, spl, kb, gcount, lane, ninvh, cw, hrow, M0, M1); break;
            }
            // ---- queue the marked near candidates of the batch: every lane appends its own marks at the offset a warp
            //      scan of the counts gives it.  Queued pairs are processed 32 at a time when the next append would not fit,
            //      and completely after the last batch.  A batch with more marks than the queue can take at once (a dense
            //      neighbourhood) goes through in 32 slices of one (pass, chunk) each — at most 64 marks. ----
            const bool last_batch = kb + 8 >= gcount;
            const int nsl = __reduce_add_sync(FULL, __popc(M0) + __popc(M1)) > V3_QCAP - 32 ? 32 : 1;
            // processes queued pairs 32 at a time: whole batches only (`all` false: make room), or everything
            auto drain = [&](const bool all) {
                __syncwarp();
                int qh = 0;
                while (qn - qh >= 32 || (all && qh < qn)) {
                    v3_drain_batch(nc, W, stage, cur.hs, cur.t0, cur.hnlim, nch, a.A.velp, sums, qh, qn, lane);
                    qh += 32;
                }
                const int left = max(qn - qh, 0);
                unsigned ent = 0;
                if (lane < left) ent = W.q[qh + lane];
                __syncwarp();
                if (lane < left) W.q[lane] = ent;
                qn = left;
                __syncwarp();
            };
#pragma unroll 1
            for (int sl = 0; sl < nsl; sl++) {
                const unsigned smask = nsl == 1 ? 0xffffffffu : 1u << sl;
                const unsigned a0 = M0 & smask, a1 = M1 & smask;
                const int mine = __popc(a0) + __popc(a1);
                // exclusive prefix of the lanes' mark counts: three ballots while every lane holds fewer than 8 marks (nearly always),
                // the shuffle scan otherwise
                int excl_m, total;
                if (__ballot_sync(FULL, mine >= 8) == 0u) {
                    const unsigned ltm = (1u << lane) - 1u;
                    const unsigned b0 = __ballot_sync(FULL, mine & 1), b1 = __ballot_sync(FULL, mine & 2), b2 = __ballot_sync(FULL, mine & 4);
                    excl_m = __popc(b0 & ltm) + 2 * __popc(b1 & ltm) + 4 * __popc(b2 & ltm);
                    total = __popc(b0) + 2 * __popc(b1) + 4 * __popc(b2);
                } else {
                    int incl = mine;
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) {
                        const int t = __shfl_up_sync(FULL, incl, o);
                        if (lane >= o) incl += t;
                    }
                    total = __shfl_sync(FULL, incl, 31);
                    excl_m = incl - mine;
                }
                if (qn + total > V3_QCAP) drain(false);
                int at = qn + excl_m;
                // raw marks, one 32-bit loop per home particle of the passes (decoded by v3_drain_batch)
                const unsigned ebase = ((unsigned)kb << 16) | ((unsigned)lane << 8);
                for (unsigned m = a0; m; m &= m - 1) W.q[at++] = ebase | (unsigned)(__ffs((int)m) - 1);
                for (unsigned m = a1; m; m &= m - 1) W.q[at++] = ebase | 32u | (unsigned)(__ffs((int)m) - 1);
                qn += total;
            }
            if (last_batch && qn > 0) drain(true);
        }
        // ---- candidate side of the sweep: one float reduction per candidate of the other bins ----
        {
            // this lane's candidates lane, lane + 32, ... ascend, and so do the runs of the table: walk it instead of searching per chunk
            int fq = 0, fnext = W.run_lo[stage][1], fbase = W.run_j[stage][0] - W.run_lo[stage][0];
#pragma unroll
            for (int k = 0; k < V3_CH; k++) {
                if (k < nch) {
                    const int c = k * 32 + lane;
                    while (c >= fnext) {                       // (entries 5..7 hold V3_TILE: the walk stops by itself)
                        fq++;
                        fbase = W.run_j[stage][fq] - W.run_lo[stage][fq];
                        fnext = W.run_lo[stage][fq + 1];
                    }
                    float lo, hi;
                    upk2(cw[k], lo, hi);
                    const float v = lo + hi;
                    if (c < ct && cur.t0 + c >= cur.hnlim && v != 0.f) red_add_f32(&sums[fbase + c].x, v * w_outer);
                }
            }
        }
        // ---- home side of the sweep, once per group ----
        if ((cur.flags & 2) && lane < gcount) red_add_f32(&sums[cur.hs + lane].x, hrow * (w_outer / V3_FIX<HASB>()));
        __syncwarp();
        if (!hnx) break;
        cur = nxt;
    }
}

cudaError_t fsg_launch_pair_v3(const PairArgs &a, float4 *sums, bool has_boundary, int sm_count, cudaStream_t s)
{
    V3Args va;
    va.a = a;
    va.sums = sums;
    static FsgAttrOnce attr_once;
    const int smem = (int)sizeof(V3Warp) * V3_WARPS;
    if (attr_once.need()) {
        cudaFuncSetAttribute(k_pair_v3<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        cudaFuncSetAttribute(k_pair_v3<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    }
    cudaError_t e = cudaMemsetAsync(sums, 0, sizeof(float4) * (size_t)a.n, s);
    if (e != cudaSuccess) return e;
    int64_t blocks = ((int64_t)a.n + V3_WARPS - 1) / V3_WARPS;      // upper bound on useful warps: one per occupied bin
    int64_t maxb = (int64_t)sm_count * (has_boundary ? V3_BPS - 1 : V3_BPS);
    if (blocks > maxb) blocks = maxb;
    if (blocks < 1) blocks = 1;
    if (has_boundary) k_pair_v3<true><<<(unsigned)blocks, V3_WARPS * 32, smem, s>>>(va);
    else k_pair_v3<false><<<(unsigned)blocks, V3_WARPS * 32, smem, s>>>(va);
    return cudaGetLastError();
}
