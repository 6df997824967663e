// fsg_unidyn_adapt.cu — particle merging / splitting of the unidyn model made live: the blocks of
// FluidGPU-unidyn.cu:260-285 (== :680-700) and the host loop of solver-unidyn.cu:495-542.
//
// In the reference nothing of this ever happens: the merge test is `ds <= (-10.00) && ds > 0` (:261), a merge would set mass 2.75
// (:262) while a split needs mass > 3 (:278), and the host loop that creates the second particle is commented out.  Here the two
// thresholds are configuration (fsg_config::unidyn_merge_distance, unidyn_split_mass_min; the reference's literals are the defaults,
// with which the pass is a no-op) and the blocks — which in the reference rewrite particles inside the pair loop while every other
// thread of the launch is reading them, and test diffusion sums that launch is still accumulating — get the race-free reading that
// oracle/fsg_oracle_unidyn.c states and these kernels are pinned against (PARITY UNPINNED against the reference: it has no live
// behaviour here):
//   1. the pair sums of the step are taken over the unmodified state (mass-weighted: cu:358-366);
//   2. k_adapt_nearest / k_adapt_merge: a particle merges with its nearest candidate (0 < ds <= merge distance, both masses in
//      (0, 2), neither a boundary particle, |diffusion|^2 < 20 for both with the COMPLETED sums of this step; ties: the smaller
//      Particle::index) if that choice is mutual; the smaller index survives with mass 2.75 and the pair's mean velocity and
//      position, the other gets mass 0, boundary = true, position 90.99 (:262-272) — it leaves the grid and is parked;
//   3. k_adapt_split: mass > split_mass_min, inside the grid, not a boundary particle, |diffusion|^2 > 35000 or dens < 9400 (:278):
//      mass 1, split flag, y += 0.015 (:279-282);
//   4. mykernel2 / Particle::update as always (k_update_unidyn);
//   5. k_adapt_children: for the split particles in DESCENDING slot order (solver-unidyn.cu:499) a child is appended at
//      (x, y - 0.03, z) of the parent's updated position with its velocity, mass 1, boundary = false (:501-520), every other field
//      the class default, while the capacity lasts.  The particle count grows on the host (one read-back per step while the pass is on).
// Mass lives in FsgState::mix.z as (mass - 1), so that every writer that knows nothing about it (z = 0) means mass 1.
// Single-device contexts, pure-fluid scenes.
#include "fsg_device.cuh"

#include "fsg_unidyn.cuh"

__device__ __forceinline__ float adapt_diff2(const float4 s2)          // powf(dx,2) + powf(dy,2) + powf(dz,2), no contraction
{
    return __fadd_rn(__fadd_rn(__fmul_rn(s2.x, s2.x), __fmul_rn(s2.y, s2.y)), __fmul_rn(s2.z, s2.z));
}
__device__ __forceinline__ bool adapt_mergeable(float4 pd, float4 mx, float4 s2)
{
    const float mass = 1.f + mx.z;
    return !(pd.w < 0.f) && mass > 0.f && mass < 2.f && adapt_diff2(s2) < 20.f;      // :261
}

__global__ void __launch_bounds__(128)
k_adapt_nearest(FsgDev d, int n, const int *__restrict__ keysA, const int *__restrict__ start, const int *__restrict__ end, FsgState A,
                const float4 *__restrict__ sums2, double merge_distance, int *__restrict__ nn)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int best_j = -1;
    const int key = keysA[i];
    if (key < d.numcells) {
        const float4 pi = A.posd[i];
        if (adapt_mergeable(pi, A.mix[i], sums2[i])) {
            float best = 0.f;
            int best_idx = 0;
            for (int t = 0; t < 27; t++) {
                const int c = key + (t / 9 - 1) * d.G2 + ((t / 3) % 3 - 1) * d.G + (t % 3 - 1);     // cu:130-132
                if (c < 0 || c >= d.numcells) continue;
                const int s0 = start[c], e0 = end[c];
                if (s0 < 0 || e0 < 0) continue;
                for (int j = s0; j <= e0 && j < n; j++) {
                    if (j == i) continue;
                    const float4 pj = A.posd[j];
                    if (!adapt_mergeable(pj, A.mix[j], sums2[j])) continue;
                    const float ds = sqrtf(dist2(pi.x - pj.x, pi.y - pj.y, pi.z - pj.z));
                    if (!((double)ds <= merge_distance && ds > 0.f)) continue;
                    const int idx = __float_as_int(A.dpi[j].w);
                    if (best_j < 0 || ds < best || (ds == best && idx < best_idx)) { best_j = j; best = ds; best_idx = idx; }
                }
            }
        }
    }
    nn[i] = best_j;
}

// the survivor of a mutual pair rewrites both particles (pairs are disjoint: no two threads touch the same record)
__global__ void __launch_bounds__(128)
k_adapt_merge(int n, const int *__restrict__ nn, FsgState A, int *__restrict__ counts)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int j = nn[i];
    if (j < 0 || nn[j] != i) return;
    if (!(__float_as_int(A.dpi[i].w) < __float_as_int(A.dpi[j].w))) return;
    float4 pi = A.posd[i], pj = A.posd[j], vi = A.velp[i], vj = A.velp[j], mi = A.mix[i], mj = A.mix[j], aj = A.accf[j];
    mi.z = 2.75f - 1.f;                                                    // :262
    mj.z = 0.f - 1.f;                                                      // :263
    vi.x = (float)((double)(vi.x + vj.x) / 2.0);                           // :266-268
    vi.y = (float)((double)(vi.y + vj.y) / 2.0);
    vi.z = (float)((double)(vi.z + vj.z) / 2.0);
    pi.x = (float)((double)(pi.x + pj.x) / 2.0);                           // :269-271
    pi.y = (float)((double)(pi.y + pj.y) / 2.0);
    pi.z = (float)((double)(pi.z + pj.z) / 2.0);
    pj.x = pj.y = pj.z = (float)90.99;                                     // :272
    pj.w = -fabsf(pj.w);                                                   // boundary = true, :265
    aj.w = __int_as_float(__float_as_int(aj.w) | 1);
    A.posd[i] = pi; A.velp[i] = vi; A.mix[i] = mi;
    A.posd[j] = pj; A.mix[j] = mj; A.accf[j] = aj;
    atomicAdd(counts + 0, 1);
}

__global__ void __launch_bounds__(256)
k_adapt_split(FsgDev d, int n, const int *__restrict__ keysA, FsgState A, const float4 *__restrict__ sums2, double split_mass_min,
              int *__restrict__ flag, int *__restrict__ counts)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int f = 0;
    if (keysA[i] < d.numcells) {
        float4 pd = A.posd[i], mx = A.mix[i];
        const float mass = 1.f + mx.z;
        if ((double)mass > split_mass_min && !(pd.w < 0.f) && (adapt_diff2(sums2[i]) > 35000.f || fabsf(pd.w) < 9400.f)) {     // :278
            mx.z = 0.f;                                                    // mass = 1, :279
            pd.y = (float)((double)pd.y + 0.015);                          // :282
            A.posd[i] = pd;
            A.mix[i] = mx;
            f = 1;
            atomicAdd(counts + 1, 1);
        }
    }
    flag[i] = f;
}

// off = exclusive scan of flag over [0, n); parents in descending slot order get child ranks 0, 1, ...
__global__ void __launch_bounds__(256)
k_adapt_children(FsgDev d, int n, int64_t cap, const int *__restrict__ flag, const int *__restrict__ off, FsgState A, FsgState B, int *__restrict__ keysB,
                 float4 *__restrict__ sums, float4 *__restrict__ sums2, float *__restrict__ vizb, int next_index, float gravity, int *__restrict__ counts)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n || !flag[j]) return;
    const int total = off[n - 1] + flag[n - 1];
    const int r = total - 1 - off[j];
    const int64_t k = (int64_t)n + r;
    if (k >= cap) return;                                                  // no room: the parent stays split, no child (solver-unidyn.cu:512)
    const float4 pp = B.posd[j], pv = B.velp[j];                           // (a split particle is never a boundary particle, :278 / :500)
    const float4 cp = make_float4(pp.x, (float)((double)pp.y - 0.03), pp.z, 9550.f);        // :501-503, dens = RHO_0
    const float4 cv = make_float4(pv.x, pv.y, pv.z, 0.f);                                   // :504-506, press = 0
    const float4 ca = make_float4(0.f, 0.f, gravity, __int_as_float(0));
    const float4 cd = make_float4(0.f, 0.f, 0.f, __int_as_float(next_index + r));
    const float4 cm = make_float4(0.f, 1.f, 0.f, 0.f);                                      // solid 0, fluid 1, mass 1 (:519)
    B.posd[k] = cp; B.velp[k] = cv; B.accf[k] = ca; B.dpi[k] = cd; B.mix[k] = cm;
    A.posd[k] = cp; A.velp[k] = cv; A.accf[k] = ca; A.dpi[k] = cd; A.mix[k] = cm;           // (what an export of this step shows for the new slot)
    sums[k] = make_float4(0.f, 0.f, 0.f, 0.f);
    sums2[k] = make_float4(0.f, 0.f, 0.f, 0.f);
    vizb[k] = 0.f;
    keysB[k] = bin_id(d, cp.x, cp.y, cp.z);                                                 // :522
    atomicAdd(counts + 2, 1);
}

__global__ void k_adapt_export_mass(int64_t n, const float4 *__restrict__ mix, float *__restrict__ a3)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) a3[i] = 1.f + mix[i].z;
}
cudaError_t fsg_launch_export_mass(int64_t n, const float4 *mix, float *a3, cudaStream_t s)
{
    if (n <= 0) return cudaSuccess;
    k_adapt_export_mass<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(n, mix, a3);
    return cudaGetLastError();
}

static cudaError_t adapt_scratch(fsg_ctx *c)
{
    if (c->adapt_ws) return cudaSuccess;
    // nn | flag | off : three int arrays of `cap`, then 4 counters
    cudaError_t e = cudaMalloc(&c->adapt_ws, sizeof(int) * (3 * (size_t)c->cap + 8));
    if (e != cudaSuccess) return e;
    c->adapt_scan_bytes = fsg_scan_temp_bytes(c->cap);
    return cudaMalloc(&c->adapt_scan, c->adapt_scan_bytes ? c->adapt_scan_bytes : 16);
}

// between the pair sums and the update: merge + split marks on the sorted pre-update state A
cudaError_t fsg_unidyn_adapt_pre(fsg_ctx *c, int64_t n, cudaStream_t s)
{
    cudaError_t e = adapt_scratch(c);
    if (e != cudaSuccess) return e;
    int *nn = c->adapt_ws, *flag = nn + c->cap, *counts = nn + 3 * c->cap;
    e = cudaMemsetAsync(counts, 0, sizeof(int) * 4, s);
    if (e != cudaSuccess) return e;
    k_adapt_nearest<<<(unsigned)((n + 127) / 128), 128, 0, s>>>(c->dev, (int)n, c->keysA, c->start, c->end, c->A, c->sums2,
                                                               c->cfg.unidyn_merge_distance, nn);
    k_adapt_merge<<<(unsigned)((n + 127) / 128), 128, 0, s>>>((int)n, nn, c->A, counts);
    k_adapt_split<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(c->dev, (int)n, c->keysA, c->A, c->sums2, c->cfg.unidyn_split_mass_min, flag, counts);
    c->launches += 3;
    return cudaGetLastError();
}

// after the update: children appended behind the n particles; the new count comes back to the host
int fsg_unidyn_adapt_post(fsg_ctx *c, int64_t n)
{
    int *nn = c->adapt_ws, *flag = nn + c->cap, *off = nn + 2 * c->cap, *counts = nn + 3 * c->cap;
    cudaStream_t s = c->stream;
    cudaError_t e = fsg_scan_exclusive(c->adapt_scan, c->adapt_scan_bytes, flag, off, n, s);
    if (e == cudaSuccess) {
        k_adapt_children<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(c->dev, (int)n, c->cap, flag, off, c->A, c->B, c->keysB, c->sums, c->sums2, c->vizb,
                                                                    c->adapt_next_index, (float)c->cfg.gravity, counts);
        e = cudaGetLastError();
    }
    int h[4] = {0, 0, 0, 0};
    if (e == cudaSuccess) e = cudaMemcpyAsync(h, counts, sizeof h, cudaMemcpyDeviceToHost, s);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    if (e != cudaSuccess) { c->err = std::string("unidyn merge / split pass: ") + cudaGetErrorString(e); return FSG_E_CUDA; }
    c->launches += 2;
    c->adapt_counts[0] = h[0]; c->adapt_counts[1] = h[1]; c->adapt_counts[2] = h[2];
    c->adapt_totals[0] += h[0]; c->adapt_totals[1] += h[1]; c->adapt_totals[2] += h[2];
    if (h[2] > 0) {
        c->n += h[2];
        c->adapt_next_index += h[2];
        c->keys_prev_valid = false;            // the arrays grew: the next step sorts from scratch
    }
    return FSG_OK;
}

// merges / splits / children of the last step and since the upload
extern "C" int fsg_unidyn_adapt_counts(fsg_ctx *c, int64_t last[3], int64_t total[3])
{
    if (!c) return FSG_E_INVALID;
    for (int k = 0; k < 3; k++) { if (last) last[k] = c->adapt_counts[k]; if (total) total[k] = c->adapt_totals[k]; }
    return FSG_OK;
}
