// fsg_base_kernels.cu — sm_100a kernels of the base particle step (reference: FluidGPU.cu / FluidGPU.cuh).
//
// Step = [radix sort of (bin id, slot)]  ->  k_reorder  ->  k_pair_update   (DESIGN.md "Kernels").
//   k_reorder      : gathers the 64-B SoA records into bin order and derives the bin tables
//                    (the reference's value-carrying sort + findneighbours, solver.cu:181-182).
//   k_pair_update  : gather-form pair sums over the 27-bin neighbourhood, fused with
//                    Particle::update and re-binning (the reference's mykernel + mykernel2).
#include "fsg_device.cuh"

// ------------------------------------------------------------------------------------------------
// small utilities
// ------------------------------------------------------------------------------------------------
__global__ void k_iota(int *p, int64_t n)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = (int)i;
}
__global__ void k_fill(int *p, int v, int64_t n)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}
cudaError_t fsg_launch_iota(int *p, int64_t n, cudaStream_t s)
{
    if (n <= 0) return cudaSuccess;
    k_iota<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(p, n);
    return cudaGetLastError();
}
// the per-step resets of the device-side counters in ONE launch (they were five memsets and a fill: on small scenes the step is
// bound by its launch count): nocc of the next list, work counter + n_live, n_keep, the boundary list's pair, the slab bounds
// (preset to n), the pair statistics
__global__ void k_step_counters(int *counters, int nxt, int n, bool slab, unsigned long long *dstats)
{
    const int t = threadIdx.x;
    if (t == 0) counters[nxt] = 0;
    if (t == 1) { counters[2] = 0; counters[3] = 0; }
    if (t == 2) counters[5] = 0;
    if (t == 3) { counters[10] = 0; counters[11] = 0; }
    if (slab && t >= 4 && t < 8) counters[16 + (t - 4)] = n;
    if (dstats && t >= 8 && t < 12) dstats[t - 8] = 0ull;
}
cudaError_t fsg_launch_step_counters(int *counters, int nxt, int n, bool slab, unsigned long long *dstats, cudaStream_t s)
{
    k_step_counters<<<1, 32, 0, s>>>(counters, nxt, n, slab, dstats);
    return cudaGetLastError();
}

cudaError_t fsg_launch_fill(int *p, int v, int64_t n, cudaStream_t s)
{
    if (n <= 0) return cudaSuccess;
    k_fill<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(p, v, n);
    return cudaGetLastError();
}

// slab_filter: a freshly uploaded particle that lies outside this context's slab is not this context's
// (every rank is handed the same scene): its slot is marked dead and disappears at the next sort.
__global__ void k_keys(FsgDev d, const float4 *__restrict__ posd, int *__restrict__ keys, int64_t n, int *any_boundary,
                       bool slab_filter, const int *__restrict__ slot_state)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    bool b = false;
    if (i < n) {
        float4 p = posd[i];
        int key = bin_id(d, p.x, p.y, p.z);
        if (slab_filter) {
            int ix = key / d.G2;
            if (key >= d.numcells) key = (d.x1 == d.G) ? key : d.dead;     // parked particles stay with the last slab
            else if (ix < d.x0 || ix >= d.x1) key = d.dead;
        }
        if (slot_state && slot_state[i] > d.numcells) key = d.dead;      // an empty slot of a downloaded slab state
        keys[i] = key;
        b = p.w < 0.f;      // (slab contexts: any boundary particle of the scene may arrive later as a ghost)
    }
    if (__any_sync(FULL, b) && (threadIdx.x & 31) == 0) atomicOr(any_boundary, 1);
}
cudaError_t fsg_launch_keys(const FsgDev &d, const float4 *posd, int *keys, int64_t n, int *any_boundary, bool slab_filter,
                            const int *slot_state, cudaStream_t s)
{
    if (n <= 0) return cudaSuccess;
    k_keys<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(d, posd, keys, n, any_boundary, slab_filter, slot_state);
    return cudaGetLastError();
}

// the same reset driven by the sorted key array (slab contexts: ghost bins have table entries but
// are not in the home-bin list)
__global__ void k_reset_tables_keys(int numcells, const int *__restrict__ keysA, int *start, int *end, int64_t n)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int key = keysA[i];
    if (key < numcells && (i == 0 || keysA[i - 1] != key)) {
        start[key] = -1;
        end[key] = -1;
    }
}
cudaError_t fsg_launch_reset_tables_keys(const FsgDev &d, const int *keysA, int *start, int *end, int64_t n, cudaStream_t s)
{
    if (n <= 0) return cudaSuccess;
    k_reset_tables_keys<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(d.numcells, keysA, start, end, n);
    return cudaGetLastError();
}

// start/end = -1 for the bins the previous step occupied (mykernel2's reset, FluidGPU.cu:427-430,
// restricted to the entries that are not already -1).
__global__ void k_reset_tables(const int *__restrict__ binlist, const int *__restrict__ nocc,
                               const int *__restrict__ keysA, int *start, int *end)
{
    int m = blockIdx.x * blockDim.x + threadIdx.x;
    int stride = gridDim.x * blockDim.x;
    int cnt = *nocc;
    for (; m < cnt; m += stride) {
        int b = binlist[m];
        start[b] = -1;
        end[b] = -1;
    }
}
cudaError_t fsg_launch_reset_tables(const int *binlist, const int *nocc, const int *keysA, int *start, int *end,
                                    int64_t n, cudaStream_t s)
{
    if (n <= 0) return cudaSuccess;
    int64_t blocks = (n + 255) / 256;
    int dev = 0, sms = 0;                                           // grid-stride kernel: 8 blocks per SM of the current device
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms < 1) sms = 1;
    if (blocks > (int64_t)sms * 8) blocks = (int64_t)sms * 8;
    k_reset_tables<<<(unsigned)blocks, 256, 0, s>>>(binlist, nocc, keysA, start, end);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// k_reorder: dst[k] = src[perm[k]] for the four float4 streams (the value half of
// thrust::sort_by_key, solver.cu:181) + findneighbours (FluidGPU.cu:106-117) + the list of occupied
// bins that k_pair_update walks.  Streaming, HBM-bound: 4+4 B keys/perm + 64 B in + 64 B out.
// ------------------------------------------------------------------------------------------------
// UPD: the source records are the sorted PRE-update state of the previous step: Particle::update (with that step's pair sums,
// `sums_src`, + accumulators carried in from an upload) is applied on the way through — the deferred-update schedule of fsg_step,
// which saves the separate update pass (read 80 B + write 72 B per particle) and one trip of the state through HBM.
// keys_next != nullptr: the bin id each particle will have after ITS next update goes to keys_next[k] (predicted_key) — the next
// step's sort input, produced before this step's pair kernel runs.
template <bool UPD>
__global__ void __launch_bounds__(256)
k_reorder(FsgDev d, int64_t n, const int *__restrict__ perm, const int *__restrict__ keysA, FsgState src,
          FsgState dst, const float4 *__restrict__ carry_src, float4 *__restrict__ carry_dst, const float4 *__restrict__ sums_src,
          int *__restrict__ keys_next, int *start, int *end,
          int *binlist, int *nocc, int *binlistB, int *noccB, int *nlive, int *nkeep, int *ranges, int *order_flag)
{
    const int numcells = d.numcells;
    int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    bool head = false, headB = false;
    if (k < n) {
        int sidx = perm[k];
        float4 a = src.posd[sidx], b = src.velp[sidx], c = src.accf[sidx], e = src.dpi[sidx];
        int key = keysA[k];
        if (UPD) {
            // particles that were parked BEFORE the update pass through unchanged, like k_update; one that leaves the grid in this
            // update (key == numcells now, position still inside) gets it, and is parked from then on
            if (key < numcells || bin_id(d, a.x, a.y, a.z) < numcells) {
                float4 sm = sums_src[sidx];
                if (carry_src) { const float4 cy = carry_src[sidx]; sm.x += cy.x; sm.y += cy.y; sm.z += cy.z; sm.w += cy.w; }
                // (the bin id of the new position is `key`: it was predicted a step ago from the same bits, predicted_key; the three double
                // divisions of a second bin_id would make this streaming kernel compute bound)
                int unused;
                particle_update<false>(d, a, b, c, e, sm.x, sm.y, sm.z, sm.w, unused);
            }
        } else if (carry_src) carry_dst[k] = carry_src[sidx];
        dst.posd[k] = a;
        dst.velp[k] = b;
        dst.accf[k] = c;
        dst.dpi[k] = e;
        if (src.mix) dst.mix[k] = src.mix[sidx];
        if (src.stress) {
            for (int q = 0; q < 18; q++) dst.stress[(size_t)k * 18 + q] = src.stress[(size_t)sidx * 18 + q];
        }
        if (keys_next) keys_next[k] = key < numcells ? predicted_key(d, a, b) : key;
        int next = k + 1 < n ? keysA[k + 1] : d.dead;
        if (k + 1 < n && next < key) atomicOr(order_flag, 1);          // the key sort's result, verified where it is consumed
        if (key < numcells) {
            int prev = k > 0 ? keysA[k - 1] : -1;
            if (key != prev) {
                start[key] = (int)k;
                int ix = key / d.G2;
                head = (ix >= d.bx0 && ix < d.bx1) || (d.sym && ix == d.x0 - 1);   // interior home bins (+ the lower ghost layer for the symmetric kernel)
                headB = ix >= d.x0 && ix < d.x1 && !head;                 // boundary home bins (empty list unless the exchange overlaps)
            }
            if (key != next) end[key] = (int)k;
            if (next >= numcells) *nlive = (int)k + 1;
        }
        if (key <= numcells && next > numcells) *nkeep = (int)k + 1;   // live + parked; dead slots are trimmed
        if (ranges) {
            // slab contexts: [0, ranges[0]) and [ranges[1], n) are the sorted slots within two layers of a face that has a
            // neighbour (plus ghosts, parked, dead): the only ones the next pack has to look at.  Preset to n by the caller.
            const int prevk = k > 0 ? keysA[k - 1] : -1;
            const int KL = d.rl * d.G2, KR = d.rr * d.G2;
            if (key >= KL && prevk < KL) ranges[0] = (int)k;
            if (key >= KR && prevk < KR) ranges[1] = (int)k;
            // the face layers themselves (sorted-ghost pipeline): first slot beyond layer x0, first slot of layer x1 - 1
            const int KL1 = (d.x0 + 1) * d.G2, KR1 = (d.x1 - 1) * d.G2;
            if (key >= KL1 && prevk < KL1) ranges[2] = (int)k;
            if (key >= KR1 && prevk < KR1) ranges[3] = (int)k;
        }
    }
    // block-aggregated append of the home-bin heads: ONE atomic per block and list (a per-warp atomic on the
    // single counter serialises in L2 and was the bottleneck of this kernel)
    __shared__ int s_cnt[2][8], s_base[2];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned m = __ballot_sync(FULL, head), mB = __ballot_sync(FULL, headB);
    if (lane == 0) { s_cnt[0][warp] = __popc(m); s_cnt[1][warp] = __popc(mB); }
    __syncthreads();
    if (threadIdx.x < 2) {
        const int L = threadIdx.x;
        int tot = 0;
#pragma unroll
        for (int w = 0; w < 8; w++) { int c = s_cnt[L][w]; s_cnt[L][w] = tot; tot += c; }
        s_base[L] = tot ? atomicAdd(L ? noccB : nocc, tot) : 0;
    }
    __syncthreads();
    if (head) binlist[s_base[0] + s_cnt[0][warp] + __popc(m & ((1u << lane) - 1))] = keysA[k];
    if (headB) binlistB[s_base[1] + s_cnt[1][warp] + __popc(mB & ((1u << lane) - 1))] = keysA[k];
}
cudaError_t fsg_launch_reorder(const FsgDev &d, int64_t n, const int *perm, const int *keysA, FsgState src,
                               FsgState dst, const float4 *carry_src, float4 *carry_dst, const float4 *sums_src, int *keys_next, int *start,
                               int *end, int *binlist, int *nocc, int *binlistB, int *noccB, int *nlive, int *nkeep, int *ranges, int *order_flag,
                               cudaStream_t s)
{
    if (n <= 0) return cudaSuccess;
    const unsigned blocks = (unsigned)((n + 255) / 256);
    if (sums_src)
        k_reorder<true><<<blocks, 256, 0, s>>>(d, n, perm, keysA, src, dst, carry_src, carry_dst, sums_src, keys_next, start, end, binlist, nocc,
                                               binlistB, noccB, nlive, nkeep, ranges, order_flag);
    else
        k_reorder<false><<<blocks, 256, 0, s>>>(d, n, perm, keysA, src, dst, carry_src, carry_dst, nullptr, keys_next, start, end, binlist, nocc,
                                                binlistB, noccB, nlive, nkeep, ranges, order_flag);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// Pair physics (FluidGPU.cu:235-279).  R is the type of the sub-expressions the reference
// evaluates in double because of its unsuffixed literals (App. A.3): float for the fast path,
// double for the promotion-faithful path (cfg.pair_fp64).
// Candidate j is staged in shared memory as  pj = (x, y, z, ±dens)  vj = (vx, vy, vz, press/dens^2).
// ------------------------------------------------------------------------------------------------
template <typename R>
__device__ __forceinline__ void pair_body(const FsgDev &d, float rabx, float raby, float rabz, float d2in,
                                          const float4 &vi, float densi, bool bi, float pod2i, const float4 &pj,
                                          const float4 &vj, float &ad, float &ax, float &ay, float &az)
{
    float ds = sqrtf(d2in);
    float densj = fabsf(pj.w);
    bool bj = pj.w < 0.f;
    // W(ds)  FluidGPU.cu:11-21
    float w;
    if (sizeof(R) == 8) {
        double q = (double)ds / d.h;
        double c = 1. / 3.14159 / (double)powf((float)d.h, 3.f);
        if (ds <= d.h_le) w = (float)(c * (1 - 3. / 2. * (double)powf((float)q, 2.f) + 3. / 4. * (double)powf((float)q, 3.f)));
        else if (ds <= d.twoh_lt) w = (float)(c * 1 / 4. * (double)powf((float)(2 - q), 3.f));
        else w = 0.f;
    } else {
        float q = ds * d.inv_h;
        if (ds <= d.h_le) w = d.w_c * (1.f - 1.5f * q * q + 0.75f * q * q * q);
        else if (ds <= d.twoh_lt) { float t = 2.f - q; w = d.w_c * 0.25f * t * t * t; }
        else w = 0.f;
    }
    float bfac_d = (!bi && bj) ? 2.5f : 1.f;    // 1 + float(!b_i)*float(b_j)*BDENSFACTOR   FluidGPU.cu:276
    ad += w * bfac_d;
    // dW(ds) has support h (FluidGPU.cu:35-43): beyond it the pressure/viscosity term is exactly zero
    if (ds <= d.h_lt) {
        float dwv;
        if (sizeof(R) == 8) dwv = (float)(-45.0 / 3.14159 / (double)powf((float)d.h, 6.f) * (double)powf((float)(d.h - (double)ds), 2.f));
        else { float t = d.hf - ds; dwv = d.dw_c * t * t; }
        float dkx = dwv * rabx / ds, dky = dwv * raby / ds, dkz = dwv * rabz / ds;          // :245-247
        float vabx = vi.x - vj.x, vaby = vi.y - vj.y, vabz = vi.z - vj.z;                   // :242-244
        float dd = vabx * rabx + vaby * raby + vabz * rabz;                                 // :253
        float s = 0.f;
        if (dd < 0.f) {                                                                     // (d < 0) factor of :255
            float d2 = ds * ds;                                                             // :254
            if (sizeof(R) == 8) {
                double mu = (double)dd / ((double)d2 + 0.01 * (double)powf((float)d.h, 2.f));
                double hm = d.h * mu;
                double bf = 1 + ((!bi && bj) ? d.alpha_boundary : 0.0);
                s = (float)((d.alpha_fluid * d.sound * (hm + 50 * 1.0 / d.sound * (double)powf((float)hm, 2.f)) /
                             (((double)(densi + densj)) / 2.0)) * bf);
            } else {
                float mu = dd / (d2 + d.eps);
                float hm = d.hf * mu;
                float bf = (!bi && bj) ? 1.f + (float)d.alpha_boundary : 1.f;
                s = d.visc_c * (hm + d.visc_q * hm * hm) / ((densi + densj) * 0.5f) * bf;
            }
        }
        float pp = vj.w + pod2i + s;                                                        // :258-260
        ax += pp * dkx;
        ay += pp * dky;
        az += pp * dkz;
    }
}

// ------------------------------------------------------------------------------------------------
// k_pair_update — one warp per occupied home bin (dynamic queue over the occupied-bin list).
//
//  phase 1  lanes 0..26 read start/end of the 27 linear-offset neighbour bins (FluidGPU.cu:124-126,
//           155-156: no per-axis clamp, wrap-around candidates are kept and die in the distance
//           test) and prefix-sum the populations.  With neighbour_cap > 0 only the first `cap`
//           neighbour particles in the reference's thread order are visited (:174, :204-231).
//  phase 2  the neighbour particles are staged once per home bin in shared memory
//           (32 B each: pos+dens, vel+press/dens^2) and reused by every home particle.
//  phase 3  for each home particle i, lanes sweep the staged candidates 32 at a time; in-range
//           candidates are compacted (ballot + popc) into a per-warp queue so that the expensive
//           pair body runs with (nearly) all lanes busy; partial sums are butterfly-reduced.
//  phase 4  lanes own one home particle each: EOS, integration, new bin id; results go to the
//           other state buffer (other warps still read this one).
// ------------------------------------------------------------------------------------------------
#define PAIR_WARPS 4
#define PAIR_TILE 512   // staged candidates per warp (32 B each)
#define PAIR_SMEM ((size_t)PAIR_WARPS * PAIR_TILE * (2 * sizeof(float4) + sizeof(unsigned short)))

template <typename R, bool STATS>
__global__ void __launch_bounds__(PAIR_WARPS * 32)
k_pair_update(PairArgs a)
{
    extern __shared__ float4 s_dyn[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float4 *sp = s_dyn + (size_t)warp * 2 * PAIR_TILE, *sv = sp + PAIR_TILE;
    unsigned short *sq = reinterpret_cast<unsigned short *>(s_dyn + (size_t)PAIR_WARPS * 2 * PAIR_TILE) + warp * PAIR_TILE;
    const FsgDev &d = a.d;
    const int nocc = *a.nocc;
    unsigned long long st_tested = 0, st_in = 0, st_drop = 0;

    for (;;) {
        int m = 0;
        if (lane == 0) m = atomicAdd(a.work, 1);
        m = __shfl_sync(FULL, m, 0);
        if (m >= nocc) break;
        const int b = a.binlist[m];

        // ---- phase 1: neighbour-bin populations ----
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
        const int all = __shfl_sync(FULL, incl, 31);
        int pc = (d.bin_cap <= 0 || p < d.bin_cap) ? p : 0;
#pragma unroll
        for (int o = 16; o; o >>= 1) pc += __shfl_xor_sync(FULL, pc, o);
        int C = pc;                                   // `total`, FluidGPU.cu:170-177
        if (d.cap > 0 && C > d.cap) C = d.cap;        // threads that exist, solver.cu:187
        if (STATS && lane == 0) st_drop += all - C;
        const int hs = __shfl_sync(FULL, st, 13), hn = __shfl_sync(FULL, p, 13);

        for (int ig = 0; ig < hn; ig += 32) {
            const int gcount = min(32, hn - ig);
            float ad = 0.f, ax = 0.f, ay = 0.f, az = 0.f;     // lane l accumulates home particle hs+ig+l
            for (int t0 = 0; t0 < C; t0 += PAIR_TILE) {
                const int Ct = min(PAIR_TILE, C - t0);
                __syncwarp();
                // ---- phase 2: stage candidates [t0, t0+Ct) of the concatenated neighbour list ----
#pragma unroll 1
                for (int t = 0; t < 27; t++) {
                    int pt = __shfl_sync(FULL, p, t);
                    if (pt == 0) continue;
                    int ex = __shfl_sync(FULL, excl, t), stt = __shfl_sync(FULL, st, t);
                    int lo = max(ex, t0), hi = min(ex + pt, t0 + Ct);
                    for (int k = lo + lane; k < hi; k += 32) {
                        int j = stt + pt - 1 - (k - ex);      // reversed inside the bin, FluidGPU.cu:228
                        float4 pj = a.A.posd[j], vj = a.A.velp[j];
                        float dj = fabsf(pj.w);
                        vj.w = vj.w / (dj * dj);              // press / powf(dens,2), :258
                        sp[k - t0] = pj;
                        sv[k - t0] = vj;
                    }
                }
                __syncwarp();
                // ---- phase 3 ----
#pragma unroll 1
                for (int il = 0; il < gcount; il++) {
                    const int i = hs + ig + il;
                    const float4 pi = a.A.posd[i], vi = a.A.velp[i];
                    const float densi = fabsf(pi.w);
                    const bool bi = pi.w < 0.f;
                    const float pod2i = vi.w / (densi * densi);
                    int qn = 0;
                    for (int c0 = 0; c0 < Ct; c0 += 32) {
                        int c = c0 + lane;
                        bool in = false;
                        if (c < Ct) {
                            float4 pj = sp[c];
                            float rx = pi.x - pj.x, ry = pi.y - pj.y, rz = pi.z - pj.z;
                            float d2 = dist2(rx, ry, rz);
                            in = (d2 <= d.d2_max) && (d2 > 0.f);                 // FluidGPU.cu:236
                        }
                        unsigned mk = __ballot_sync(FULL, in);
                        if (in) sq[qn + __popc(mk & ((1u << lane) - 1))] = (unsigned short)c;
                        qn += __popc(mk);
                    }
                    if (STATS) { st_tested += (lane == 0) ? Ct : 0; st_in += (lane == 0) ? qn : 0; }
                    __syncwarp();
                    float td = 0.f, tx = 0.f, ty = 0.f, tz = 0.f;
                    for (int q = lane; q < qn; q += 32) {
                        int c = sq[q];
                        float4 pj = sp[c], vj = sv[c];
                        float rx = pi.x - pj.x, ry = pi.y - pj.y, rz = pi.z - pj.z;
                        float d2 = dist2(rx, ry, rz);
                        pair_body<R>(d, rx, ry, rz, d2, vi, densi, bi, pod2i, pj, vj, td, tx, ty, tz);
                    }
#pragma unroll
                    for (int o = 16; o; o >>= 1) {
                        td += __shfl_xor_sync(FULL, td, o);
                        tx += __shfl_xor_sync(FULL, tx, o);
                        ty += __shfl_xor_sync(FULL, ty, o);
                        tz += __shfl_xor_sync(FULL, tz, o);
                    }
                    if (lane == il) { ad += td; ax += tx; ay += ty; az += tz; }
                    __syncwarp();
                }
            }
            // ---- phase 4 ----
            if (lane < gcount) {
                const int i = hs + ig + lane;
                float4 pd = a.A.posd[i], vp = a.A.velp[i], af = a.A.accf[i], dpi = a.A.dpi[i];
                float nd = ad, nx = ax, ny = ay, nz = az;
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
    }
    if (STATS && lane == 0) {
        atomicAdd(a.stats + 0, st_tested);
        atomicAdd(a.stats + 1, st_in);
        atomicAdd(a.stats + 2, st_drop);
    }
}

// parked particles (bin id == numcells) are copied through unchanged
__global__ void k_copy_parked(int64_t from, int64_t n, FsgState A, FsgState B, int *keysB, int numcells)
{
    int64_t i = from + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        B.posd[i] = A.posd[i];
        B.velp[i] = A.velp[i];
        B.accf[i] = A.accf[i];
        B.dpi[i] = A.dpi[i];
        keysB[i] = numcells;
    }
}
__global__ void k_copy_parked_dyn(const int *nlive, int64_t n, FsgState A, FsgState B, int *keysB, int numcells)
{
    int64_t from = *nlive;
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = from + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        B.posd[i] = A.posd[i];
        B.velp[i] = A.velp[i];
        B.accf[i] = A.accf[i];
        B.dpi[i] = A.dpi[i];
        keysB[i] = numcells;
    }
}

static PairArgs pair_args(const fsg_ctx *c, int64_t n, const int *binlist, const int *nocc, int *work, const float4 *carry)
{
    PairArgs a;
    a.d = c->dev;
    a.n = (int)n;
    a.keysA = c->keysA;
    a.start = c->start;
    a.end = c->end;
    a.binlist = binlist;
    a.nocc = nocc;
    a.work = work;
    a.A = c->A;
    a.B = c->B;
    a.keysB = c->keysB;
    a.carry = carry;
    a.stats = c->dstats;
    a.sums = nullptr;
    return a;
}

// the pair sums of the uncapped fp32 configuration into c->sums, nothing else (deferred-update schedule)
cudaError_t fsg_launch_pair_sums(const fsg_ctx *c, int64_t n, const int *binlist, const int *nocc, int *work, int *launches, cudaStream_t s)
{
    if (n <= 0) return cudaSuccess;
    const PairArgs a = pair_args(c, n, binlist, nocc, work, nullptr);
    *launches += 1;
    if (c->dev.sym) return fsg_launch_pair_v3(a, c->sums, c->has_boundary, c->sm_count, s);
    return fsg_launch_pair_v2(a, c->sums, c->cfg.collect_stats != 0, c->has_boundary, c->sm_count, 0, s);
}

cudaError_t fsg_launch_pair_update(const fsg_ctx *c, int64_t n, const int *binlist, const int *nocc, int *work,
                                   const float4 *carry, int *launches, cudaStream_t s)
{
    if (n <= 0) return cudaSuccess;
    PairArgs a;
    a.d = c->dev;
    a.n = (int)n;
    a.keysA = c->keysA;
    a.start = c->start;
    a.end = c->end;
    a.binlist = binlist;
    a.nocc = nocc;
    a.work = work;
    a.A = c->A;
    a.B = c->B;
    a.keysB = c->keysB;
    a.carry = carry;
    a.stats = c->dstats;
    a.sums = nullptr;
    // persistent grid: resident blocks per SM x SM count (static smem 4*(512*32+1024) = 68 KB -> 3 blocks/SM)
    int64_t warps_needed = n;   // upper bound on occupied bins
    int64_t blocks = (warps_needed + PAIR_WARPS - 1) / PAIR_WARPS;
    int64_t maxb = (int64_t)c->sm_count * 3;
    if (blocks > maxb) blocks = maxb;
    if (blocks < 1) blocks = 1;
    const bool stats = c->cfg.collect_stats != 0;
    const size_t smem = PAIR_SMEM;
    static FsgAttrOnce attr_once;
    if (attr_once.need()) {
        cudaFuncSetAttribute(k_pair_update<double, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cudaFuncSetAttribute(k_pair_update<double, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cudaFuncSetAttribute(k_pair_update<float, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cudaFuncSetAttribute(k_pair_update<float, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    }
    // Uncapped configuration (every particle of the 27 bins is visited): the pipelined pair-sum kernel
    // + the streaming update kernel (fsg_pair_v2.cu).  pair_fp64 == 3 keeps the fused kernel instead.
    if (c->cfg.pair_fp64 == 0 && c->dev.cap <= 0 && c->dev.bin_cap <= 0) {
        cudaError_t e2;
        if (c->overlap && c->cfg.world > 1) {
            // boundary bins first; once their particles are updated the next step's messages are packed and copied
            // on the communication stream while the interior bins are still being computed on this one
            PairArgs b = a;
            b.binlist = c->binlistB;
            b.nocc = c->counters + 10;
            b.work = c->counters + 11;
            e2 = fsg_launch_pair_v2(b, c->sums, stats, c->has_boundary, c->sm_count, 0, s);
            if (e2 != cudaSuccess) return e2;
            e2 = fsg_launch_update(c->dev, n, c->keysA, c->A, c->B, c->keysB, c->sums, carry, 1, c->counters + 6, s);
            if (e2 != cudaSuccess) return e2;
            if (fsg_slab_send_next(const_cast<fsg_ctx *>(c)) != FSG_OK) return cudaErrorUnknown;
            e2 = fsg_launch_pair_v2(a, c->sums, stats, c->has_boundary, c->sm_count, 5, s);      // leave room for the pack kernels
            if (e2 != cudaSuccess) return e2;
            e2 = fsg_launch_update(c->dev, n, c->keysA, c->A, c->B, c->keysB, c->sums, carry, 2, c->counters + 6, s);
            *launches += 4;
            return e2;
        }
        if (c->dev.sym) e2 = fsg_launch_pair_v3(a, c->sums, c->has_boundary, c->sm_count, s);
        else e2 = fsg_launch_pair_v2(a, c->sums, stats, c->has_boundary, c->sm_count, 0, s);
        if (e2 != cudaSuccess) return e2;
        e2 = fsg_launch_update(c->dev, n, c->keysA, c->A, c->B, c->keysB, c->sums, carry, 0, c->cfg.world > 1 ? c->counters + 6 : nullptr, s);
        *launches += 2;
        return e2;
    }
    // pair_fp64: 1 = promotion-faithful double path, 2 = the same queue-everything kernel in fp32
    // (kept as a cross-check of the fast kernel), 0 / 3 = fused fp32 kernel (fsg_pair_fast.cu)
    if (c->cfg.pair_fp64 == 1) {
        if (stats) k_pair_update<double, true><<<(unsigned)blocks, PAIR_WARPS * 32, smem, s>>>(a);
        else k_pair_update<double, false><<<(unsigned)blocks, PAIR_WARPS * 32, smem, s>>>(a);
    } else if (c->cfg.pair_fp64 == 2) {
        if (stats) k_pair_update<float, true><<<(unsigned)blocks, PAIR_WARPS * 32, smem, s>>>(a);
        else k_pair_update<float, false><<<(unsigned)blocks, PAIR_WARPS * 32, smem, s>>>(a);
    } else {
        cudaError_t e2 = fsg_launch_pair_fast(a, stats, c->sm_count, s);
        if (e2 != cudaSuccess) return e2;
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    k_copy_parked_dyn<<<32, 256, 0, s>>>(c->counters + 3, n, c->A, c->B, c->keysB, c->dev.numcells);
    *launches += 2;
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// AoS <-> SoA (the reference's 340-byte Particle record; offsets from FluidGPU.cuh:112-162, pinned
// by tests/golden/kat_base.json "offsets").
// ------------------------------------------------------------------------------------------------
namespace aos_base {
enum { POS = 0, VEL = 12, ACC = 24, INDEX = 36, CELL = 40, MASS = 44, DENS = 48, PRESS = 52, DELP_Z = 56, DELP_Y = 60,
       DELP_X = 64, NEWDENS = 84, NEWPRESS = 88, NDELP_Z = 92, NDELP_Y = 96, NDELP_X = 100, BOUNDARY = 336, SOLID = 337,
       FLAG = 338 };
}

__device__ __forceinline__ float ldf(const unsigned char *r, int off) { return *reinterpret_cast<const float *>(r + off); }
__device__ __forceinline__ void stf(unsigned char *r, int off, float v) { *reinterpret_cast<float *>(r + off) = v; }

__global__ void k_unpack_aos_base(const unsigned char *__restrict__ aos, int64_t n, FsgState st, float4 *carry)
{
    using namespace aos_base;
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const unsigned char *r = aos + i * FSG_AOS_STRIDE;
    bool bnd = r[BOUNDARY] != 0, solid = r[SOLID] != 0;
    float dens = ldf(r, DENS);
    st.posd[i] = make_float4(ldf(r, POS), ldf(r, POS + 4), ldf(r, POS + 8), bnd ? -dens : dens);
    st.velp[i] = make_float4(ldf(r, VEL), ldf(r, VEL + 4), ldf(r, VEL + 8), ldf(r, PRESS));
    st.accf[i] = make_float4(ldf(r, ACC), ldf(r, ACC + 4), ldf(r, ACC + 8), __int_as_float((bnd ? 1 : 0) | (solid ? 2 : 0)));
    st.dpi[i] = make_float4(ldf(r, DELP_X), ldf(r, DELP_Y), ldf(r, DELP_Z), __int_as_float(*reinterpret_cast<const int *>(r + INDEX)));
    carry[i] = make_float4(ldf(r, NEWDENS), ldf(r, NDELP_X), ldf(r, NDELP_Y), ldf(r, NDELP_Z));
}

__global__ void k_pack_aos_base(unsigned char *__restrict__ aos, int64_t n, FsgState st, const float4 *carry,
                                const int *keys, float p0)
{
    using namespace aos_base;
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    unsigned char *r = aos + i * FSG_AOS_STRIDE;
    for (int o = 0; o < FSG_AOS_STRIDE; o += 4) *reinterpret_cast<int *>(r + o) = 0;
    float4 pd = st.posd[i], vp = st.velp[i], af = st.accf[i], dpi = st.dpi[i];
    int fl = __float_as_int(af.w);
    stf(r, POS, pd.x); stf(r, POS + 4, pd.y); stf(r, POS + 8, pd.z);
    stf(r, VEL, vp.x); stf(r, VEL + 4, vp.y); stf(r, VEL + 8, vp.z);
    stf(r, ACC, af.x); stf(r, ACC + 4, af.y); stf(r, ACC + 8, af.z);
    *reinterpret_cast<int *>(r + INDEX) = __float_as_int(dpi.w);
    *reinterpret_cast<int *>(r + CELL) = keys[i];
    stf(r, MASS, 1.f);                      // FluidGPU.cuh:132
    stf(r, DENS, fabsf(pd.w));
    stf(r, PRESS, vp.w);
    stf(r, DELP_X, dpi.x); stf(r, DELP_Y, dpi.y); stf(r, DELP_Z, dpi.z);
    float4 cy = carry ? carry[i] : make_float4(0.f, 0.f, 0.f, 0.f);   // zeroed by mykernel2, FluidGPU.cu:422-425
    stf(r, NEWDENS, cy.x);
    stf(r, NEWPRESS, p0);                   // FluidGPU.cuh:145
    stf(r, NDELP_X, cy.y); stf(r, NDELP_Y, cy.z); stf(r, NDELP_Z, cy.w);
    r[BOUNDARY] = (fl & 1) ? 1 : 0;
    r[SOLID] = (fl & 2) ? 1 : 0;
    r[FLAG] = 0;
}

cudaError_t fsg_launch_unpack_aos(int model, const unsigned char *aos, int64_t n, FsgState st, float4 *carry,
                                  cudaStream_t s)
{
    if (n <= 0) return cudaSuccess;
    if (model != FSG_MODEL_BASE) return cudaErrorNotSupported;
    k_unpack_aos_base<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(aos, n, st, carry);
    return cudaGetLastError();
}
cudaError_t fsg_launch_pack_aos(int model, unsigned char *aos, int64_t n, FsgState st, const float4 *carry,
                                const int *keys, float p0, cudaStream_t s)
{
    if (n <= 0) return cudaSuccess;
    if (model != FSG_MODEL_BASE) return cudaErrorNotSupported;
    k_pack_aos_base<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(aos, n, st, carry, keys, p0);
    return cudaGetLastError();
}

// mykernel2's export (FluidGPU.cu:410-414) from the sorted pre-update state
__global__ void k_export_viz(int64_t n, const float4 *__restrict__ posd, const int *__restrict__ keys, float *spts,
                             float *a3, float *b3)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float4 p = posd[i];
    if (spts) { spts[3 * i] = p.x; spts[3 * i + 1] = p.y; spts[3 * i + 2] = p.z; }
    if (a3) a3[i] = fabsf(p.w);
    if (b3) b3[i] = (float)keys[i];
}
cudaError_t fsg_launch_export_viz(int64_t n, const float4 *posd, const int *keys, float *spts, float *a3, float *b3,
                                  cudaStream_t s)
{
    if (n <= 0) return cudaSuccess;
    k_export_viz<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(n, posd, keys, spts, a3, b3);
    return cudaGetLastError();
}
