// fsg_pair_fast.cu — the production pair-sum + update kernel (fp32) of the base particle step.
//
// Reference work it replaces: mykernel (FluidGPU.cu:119-285) + mykernel2 (FluidGPU.cu:404-432).
//
// One warp per occupied home bin, taken from a dynamic queue.  The kernel is CUDA-core bound
// (DESIGN.md "Roofline"), so its structure is about instruction count and lane utilisation:
//
//   * candidates of the 27 linear-offset neighbour bins are staged ONCE per home bin in shared
//     memory as (x, y, z, boundary) + their global slot — 20 B each — padded to a multiple of 32
//     with far-away sentinels so the sweep needs no bounds checks;
//   * SWEEP (per home particle i, lanes = candidates): the distance test and the W(r) density term
//     of the outer kernel branch (h < r <= 2h, 7/8 of all in-range pairs, FluidGPU.cu:15-16) are
//     evaluated branch-free for every candidate and accumulated in a lane-private register;
//   * only candidates with r <= h — the support of dW (FluidGPU.cu:35-43), 1/8 of the in-range
//     pairs — carry the pressure/viscosity term.  They are compacted (ballot + popc) into a
//     per-warp queue ACROSS home particles and processed 32 at a time with every lane busy;
//     per-particle sums come out of a segmented warp scan keyed by the home particle, so the
//     result is deterministic (no atomics);
//   * lanes then own one home particle each for EOS / integration / re-binning and write the other
//     state buffer.
#include "fsg_device.cuh"

#define FAST_WARPS 4
#define FAST_TILE 512                    // staged candidates per warp
#define FAST_QCAP (FAST_TILE + 32)       // queue: < 32 left over + at most one sweep's worth

struct FastWarpSmem {
    float4 sp[FAST_TILE];                // x, y, z, boundary (0/1)
    int sj[FAST_TILE];                   // global slot of the candidate
    unsigned q[FAST_QCAP];               // (home particle within group << 16) | candidate
    float4 acc[32];                      // per home particle: dens, delpress x/y/z  (queue part)
};
#define FAST_SMEM (sizeof(FastWarpSmem) * FAST_WARPS)

// pressure / viscosity / inner-W terms of one r <= h pair (FluidGPU.cu:238-279)
__device__ __forceinline__ float4 near_pair(const FsgDev &d, const float4 &pi, const float4 &vi, const float4 &pj,
                                            const float4 &vj)
{
    float rx = pi.x - pj.x, ry = pi.y - pj.y, rz = pi.z - pj.z;
    float d2 = dist2(rx, ry, rz);
    float ds = sqrtf(d2);
    float densi = fabsf(pi.w), densj = fabsf(pj.w);
    bool bi = pi.w < 0.f, bj = pj.w < 0.f;
    float q = ds * d.inv_h;
    float w = d.w_c * (1.f - 1.5f * q * q + 0.75f * q * q * q);                // FluidGPU.cu:13
    float t = d.hf - ds;
    float dwv = d.dw_c * t * t;                                                // FluidGPU.cu:37 (0 at r == h)
    float g = dwv / ds;
    float vabx = vi.x - vj.x, vaby = vi.y - vj.y, vabz = vi.z - vj.z;
    float dd = vabx * rx + vaby * ry + vabz * rz;                              // :253
    float s = 0.f;
    if (dd < 0.f) {                                                            // :255
        float mu = dd / (ds * ds + d.eps);
        float hm = d.hf * mu;
        float bf = (!bi && bj) ? 1.f + (float)d.alpha_boundary : 1.f;
        s = d.visc_c * (hm + d.visc_q * hm * hm) / ((densi + densj) * 0.5f) * bf;
    }
    float pp = vj.w / (densj * densj) + vi.w / (densi * densi) + s;            // :258-260
    float pg = pp * g;
    return make_float4(w * ((!bi && bj) ? 2.5f : 1.f), pg * rx, pg * ry, pg * rz);
}

template <bool STATS, bool FUSED>
__global__ void __launch_bounds__(FAST_WARPS * 32, 4)
k_pair_update_fast(PairArgs a)
{
    extern __shared__ __align__(16) unsigned char s_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    FastWarpSmem &S = reinterpret_cast<FastWarpSmem *>(s_raw)[warp];
    const FsgDev &d = a.d;
    const int nocc = *a.nocc;
    const unsigned d2max_bits = __float_as_uint(d.d2_max), d2h_bits = __float_as_uint(d.d2_h);
    const unsigned lt_mask = (1u << lane) - 1u;
    const float w_outer = d.w_c * 0.25f;
    unsigned long long st_tested = 0, st_in = 0, st_drop = 0;

    for (;;) {
        int m = 0;
        if (lane == 0) m = atomicAdd(a.work, 1);
        m = __shfl_sync(FULL, m, 0);
        if (m >= nocc) break;
        const int b = a.binlist[m];

        // ---- neighbour-bin populations (FluidGPU.cu:150-183) ----
        int p = 0, st = 0;
        if (lane < 27) {
            int off = (lane / 9 - 1) * d.G2 + ((lane / 3) % 3 - 1) * d.G + (lane % 3 - 1);
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
        int pc = (d.bin_cap <= 0 || p < d.bin_cap) ? p : 0;
#pragma unroll
        for (int o = 16; o; o >>= 1) pc += __shfl_xor_sync(FULL, pc, o);
        int C = pc;                                   // `total`
        if (d.cap > 0 && C > d.cap) C = d.cap;        // threads that exist in the reference launch
        if (STATS) { int all = __shfl_sync(FULL, incl, 31); if (lane == 0) st_drop += all - C; }
        const int hs = __shfl_sync(FULL, st, 13), hn = __shfl_sync(FULL, p, 13);

        for (int ig = 0; ig < hn; ig += 32) {
            const int gcount = min(32, hn - ig);
            float aw = 0.f;                                       // lane l: sweep part of the density of particle ig+l
            S.acc[lane] = make_float4(0.f, 0.f, 0.f, 0.f);
            for (int t0 = 0; t0 < C; t0 += FAST_TILE) {
                const int Ct = min(FAST_TILE, C - t0);
                const int Cpad = (Ct + 31) & ~31;
                __syncwarp();
                // ---- stage candidates [t0, t0+Ct) of the concatenated neighbour list ----
#pragma unroll 1
                for (int t = 0; t < 27; t++) {
                    int pt = __shfl_sync(FULL, p, t);
                    if (pt == 0) continue;
                    int ex = __shfl_sync(FULL, excl, t), stt = __shfl_sync(FULL, st, t);
                    int lo = max(ex, t0), hi = min(ex + pt, t0 + Ct);
                    for (int k = lo + lane; k < hi; k += 32) {
                        int j = stt + pt - 1 - (k - ex);          // reversed inside the bin, FluidGPU.cu:228
                        float4 pj = a.A.posd[j];
                        pj.w = pj.w < 0.f ? 1.f : 0.f;
                        S.sp[k - t0] = pj;
                        S.sj[k - t0] = j;
                    }
                }
                if (Ct + lane < Cpad) S.sp[Ct + lane] = make_float4(1e30f, 1e30f, 1e30f, 0.f);
                __syncwarp();

                int qn = 0;
#pragma unroll 1
                for (int il = 0; il < gcount; il++) {
                    const float4 pi = a.A.posd[hs + ig + il];
                    const float ci = pi.w < 0.f ? 0.f : 1.5f;     // float(!b_i)*BDENSFACTOR, FluidGPU.cu:276
                    float wsum = 0.f;
                    int nin = 0;
                    // ---- sweep ----
#pragma unroll 2
                    for (int c0 = 0; c0 < Cpad; c0 += 32) {
                        const float4 pj = S.sp[c0 + lane];
                        float rx = pi.x - pj.x, ry = pi.y - pj.y, rz = pi.z - pj.z;
                        float d2 = dist2(rx, ry, rz);
                        unsigned u = __float_as_uint(d2) - 1u;    // 0 < d2 <= thr  <=>  bits(d2) - 1 < bits(thr)
                        bool inr = u < d2max_bits;                // FluidGPU.cu:236
                        bool nearp = u < d2h_bits;                // r <= h
                        float inv = rsqrtf(d2);
                        float tt = fmaf(-d2 * inv, d.inv_h, 2.f); // 2 - r/h
                        float t3 = tt * tt * tt;
                        float fac = fmaf(ci, pj.w, 1.f);
                        if (inr && !nearp) wsum = fmaf(t3, fac, wsum);
                        unsigned mk = __ballot_sync(FULL, nearp);
                        if (nearp) S.q[qn + __popc(mk & lt_mask)] = ((unsigned)il << 16) | (unsigned)(c0 + lane);
                        qn += __popc(mk);
                        if (STATS) nin += __popc(__ballot_sync(FULL, inr));
                    }
                    if (STATS && lane == 0) { st_tested += Ct; st_in += nin; }
#pragma unroll
                    for (int o = 16; o; o >>= 1) wsum += __shfl_xor_sync(FULL, wsum, o);
                    if (lane == il) aw += wsum * w_outer;
                    // ---- drain full batches of near pairs ----
                    __syncwarp();
                    int qh = 0;
                    while (qn - qh >= 32 || (il == gcount - 1 && qh < qn)) {
                        const int e = qh + lane;
                        const bool valid = e < qn;
                        int key = -1 - lane;
                        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (valid) {
                            unsigned ent = S.q[e];
                            key = (int)(ent >> 16);
                            int c = (int)(ent & 0xffffu);
                            int i = hs + ig + key, j = S.sj[c];
                            v = near_pair(d, a.A.posd[i], a.A.velp[i], a.A.posd[j], a.A.velp[j]);
                        }
                        // segmented inclusive scan keyed by the home particle (keys are non-decreasing)
#pragma unroll
                        for (int o = 1; o < 32; o <<= 1) {
                            int ku = __shfl_up_sync(FULL, key, o);
                            float x0 = __shfl_up_sync(FULL, v.x, o), x1 = __shfl_up_sync(FULL, v.y, o);
                            float x2 = __shfl_up_sync(FULL, v.z, o), x3 = __shfl_up_sync(FULL, v.w, o);
                            if (lane >= o && ku == key) { v.x += x0; v.y += x1; v.z += x2; v.w += x3; }
                        }
                        int kn = __shfl_down_sync(FULL, key, 1);
                        if (valid && (lane == 31 || kn != key)) {
                            float4 c4 = S.acc[key];
                            c4.x += v.x; c4.y += v.y; c4.z += v.z; c4.w += v.w;
                            S.acc[key] = c4;
                        }
                        __syncwarp();
                        qh += 32;
                    }
                    if (qh > 0) {                                   // keep the < 32 leftovers at the front
                        int left = qn - qh;
                        unsigned ent = 0;
                        if (lane < left) ent = S.q[qh + lane];
                        __syncwarp();
                        if (lane < left) S.q[lane] = ent;
                        qn = left > 0 ? left : 0;
                        __syncwarp();
                    }
                }
            }
            // ---- EOS, integration, new bin id (Particle::update + FluidGPU.cu:419-425) ----
            __syncwarp();
            if (lane < gcount) {
                const int i = hs + ig + lane;
                float4 q4 = S.acc[lane];
                float nd = aw + q4.x, nx = q4.y, ny = q4.z, nz = q4.w;
                if (!FUSED) {                     // stage API: the sums only, the caller's mykernel2 stage consumes them
                    a.sums[i] = make_float4(nd, nx, ny, nz);
                } else {
                    float4 pd = a.A.posd[i], vp = a.A.velp[i], af = a.A.accf[i], dpi = a.A.dpi[i];
                    if (a.carry) { float4 cy = a.carry[i]; nd += cy.x; nx += cy.y; ny += cy.z; nz += cy.w; }
                    int key;
                    particle_update(d, pd, vp, af, dpi, nd, nx, ny, nz, key);
                    a.B.posd[i] = pd;
                    a.B.velp[i] = vp;
                    a.B.accf[i] = af;
                    a.B.dpi[i] = dpi;
                    a.keysB[i] = key;
                }
            }
            __syncwarp();
        }
    }
    if (STATS && lane == 0) {
        atomicAdd(a.stats + 0, st_tested);
        atomicAdd(a.stats + 1, st_in);
        atomicAdd(a.stats + 2, st_drop);
    }
}

cudaError_t fsg_launch_pair_fast(const PairArgs &a, bool stats, int sm_count, cudaStream_t s)
{
    static FsgAttrOnce attr_once;
    if (attr_once.need()) {
        cudaFuncSetAttribute(k_pair_update_fast<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FAST_SMEM);
        cudaFuncSetAttribute(k_pair_update_fast<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FAST_SMEM);
        cudaFuncSetAttribute(k_pair_update_fast<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FAST_SMEM);
    }
    int64_t blocks = ((int64_t)a.n + FAST_WARPS - 1) / FAST_WARPS;
    int64_t maxb = (int64_t)sm_count * 4;      // persistent: 4 resident blocks (16 warps) per SM
    if (blocks > maxb) blocks = maxb;
    if (blocks < 1) blocks = 1;
    if (a.sums) k_pair_update_fast<false, false><<<(unsigned)blocks, FAST_WARPS * 32, FAST_SMEM, s>>>(a);   // sums only
    else if (stats) k_pair_update_fast<true, true><<<(unsigned)blocks, FAST_WARPS * 32, FAST_SMEM, s>>>(a);
    else k_pair_update_fast<false, true><<<(unsigned)blocks, FAST_WARPS * 32, FAST_SMEM, s>>>(a);
    return cudaGetLastError();
}
