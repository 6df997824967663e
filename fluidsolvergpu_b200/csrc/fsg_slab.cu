// fsg_slab.cu — slab decomposition along x, the slowest bin axis: the multi-device hand-off of
// solver-unidyn.cu:396-470 (find_idx / host-staged cudaMemcpy / mem_shift, FluidGPU-unidyn.cu:499-542)
// re-designed for N slabs.  Because bin id = ix*G^2 + iy*G + iz, a slab is a contiguous range of
// bin ids; the ghost band is one bin layer on each side (the reference's `buffer`, solver-unidyn.cu:187).
//
// Per step and slab:  pack (this file) -> exchange with the two x-neighbours (caller: NCCL
// send/recv) -> unpack (this file) -> sort / reorder / pair sums / update (fsg_step).
//   migrants : particles whose new bin layer lies outside [x0, x1): full state; the sender keeps them one
//              more step as ghosts (they land in the neighbour's outermost layer)
//   ghosts   : particles in layer x0 (for the left neighbour) or x1-1 (for the right one): read state
// Message order = current particle order (two-phase count / scan / scatter, no atomics), so runs are
// reproducible.
#include "fsg_device.cuh"

#include <cub/device/device_scan.cuh>
#include <stdio.h>

size_t fsg_scan_temp_bytes(int64_t n)
{
    size_t bytes = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, bytes, (const int *)nullptr, (int *)nullptr, n);
    return bytes;
}
cudaError_t fsg_scan_exclusive(void *tmp, size_t tmp_bytes, const int *in, int *out, int64_t n, cudaStream_t s)
{
    return cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, in, out, n, s);
}

// category of slot i: bit 0 migrate left, bit 1 ghost for left, bit 2 migrate right, bit 3 ghost for right
__device__ __forceinline__ int slab_category(const FsgDev &d, int key, int rank, int world)
{
    if (key >= d.numcells) return 0;              // parked / dead
    int ix = key / d.G2;
    int c = 0;
    if (ix < d.x0) c = 1;
    else if (ix >= d.x1) c = 4;
    else {
        if (ix == d.x0 && rank > 0) c |= 2;
        if (ix == d.x1 - 1 && rank < world - 1) c |= 8;
    }
    return c;
}

// counts per warp: cnt[cat * nw + warp]
__global__ void __launch_bounds__(256)
k_slab_count(FsgDev d, int rank, int world, int64_t n, const int *__restrict__ keys, int *__restrict__ cnt, int64_t nw,
             int *violation)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int c = 0;
    if (i < n) {
        int key = keys[i];
        c = slab_category(d, key, rank, world);
        // a particle that moved more than one bin layer in a step has left the one-layer ghost band
        if (key < d.numcells && (key / d.G2 < d.x0 - 1 || key / d.G2 > d.x1)) atomicOr(violation, 1);
    }
    int64_t w = i >> 5;
    int lane = threadIdx.x & 31;
#pragma unroll
    for (int k = 0; k < 4; k++) {
        unsigned m = __ballot_sync(FULL, (c >> k) & 1);
        if (lane == 0 && w < nw) cnt[k * nw + w] = __popc(m);
    }
    if (i == 0) cnt[4 * nw] = 0;
}

struct SlabMsg {
    float4 *m_posd, *m_velp, *m_accf, *m_dpi, *g_posd, *g_velp;
};
__host__ __device__ inline SlabMsg slab_msg(void *base, int64_t m, int64_t g)
{
    SlabMsg r;
    float4 *p = (float4 *)base;
    r.m_posd = p; r.m_velp = p + m; r.m_accf = p + 2 * m; r.m_dpi = p + 3 * m;
    r.g_posd = p + 4 * m; r.g_velp = p + 4 * m + g;
    return r;
}

// off = exclusive scan of cnt (length 4*nw + 1).  tot[k] = off[(k+1)*nw] - off[k*nw].
__global__ void k_slab_totals(const int *__restrict__ off, int64_t nw, const int *__restrict__ nkeep_dev, int have_nkeep,
                              int64_t n_host, const int *violation, int64_t *out)
{
    if (threadIdx.x < 4) out[threadIdx.x] = off[(threadIdx.x + 1) * nw] - off[threadIdx.x * nw];
    if (threadIdx.x == 4) out[4] = have_nkeep ? (int64_t)*nkeep_dev : n_host;
    if (threadIdx.x == 5) out[5] = *violation;
}

__global__ void __launch_bounds__(256)
k_slab_scatter(FsgDev d, int rank, int world, int64_t n, const int *__restrict__ keys, FsgState B, const int *__restrict__ off,
               int64_t nw, void *to_left, void *to_right)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int c = i < n ? slab_category(d, keys[i], rank, world) : 0;
    int64_t w = i >> 5;
    int lane = threadIdx.x & 31;
    unsigned lt = (1u << lane) - 1u;
    int pos[4];
#pragma unroll
    for (int k = 0; k < 4; k++) {
        unsigned m = __ballot_sync(FULL, (c >> k) & 1);
        pos[k] = (w < nw ? off[k * nw + w] - off[k * nw] : 0) + __popc(m & lt);
    }
    if (!c) return;
    const int64_t mL = off[1 * nw] - off[0], gL = off[2 * nw] - off[1 * nw], mR = off[3 * nw] - off[2 * nw], gR = off[4 * nw] - off[3 * nw];
    float4 pd = B.posd[i], vp = B.velp[i];
    if (c & 5) {
        // migrant: full state goes to the neighbour.  The slot is NOT freed here: a particle moves less
        // than one bin per step, so it lands in the neighbour's outermost layer, where this slab still
        // needs it as a candidate for one more step.  Its bin is outside [x0, x1), so it is treated as
        // a ghost (never a home particle) and k_update drops it.
        SlabMsg M = (c & 1) ? slab_msg(to_left, mL, gL) : slab_msg(to_right, mR, gR);
        int q = (c & 1) ? pos[0] : pos[2];
        M.m_posd[q] = pd;
        M.m_velp[q] = vp;
        M.m_accf[q] = B.accf[i];
        M.m_dpi[q] = B.dpi[i];
    }
    if (c & 2) { SlabMsg M = slab_msg(to_left, mL, gL); M.g_posd[pos[1]] = pd; M.g_velp[pos[1]] = vp; }
    if (c & 8) { SlabMsg M = slab_msg(to_right, mR, gR); M.g_posd[pos[3]] = pd; M.g_velp[pos[3]] = vp; }
}

// appended slots: migrants (full state) then ghosts (read state); keys from the positions, exactly
// as the owner computed them (bin_id is the same function on both sides)
__global__ void __launch_bounds__(256)
k_slab_unpack(FsgDev d, const void *msg, int64_t m, int64_t g, int64_t at, FsgState B, float4 *carry, int *keys)
{
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= m + g) return;
    SlabMsg M = slab_msg(const_cast<void *>(msg), m, g);
    int64_t i = at + t;
    float4 pd, vp, af, dp;
    if (t < m) { pd = M.m_posd[t]; vp = M.m_velp[t]; af = M.m_accf[t]; dp = M.m_dpi[t]; }
    else {
        pd = M.g_posd[t - m]; vp = M.g_velp[t - m];
        af = make_float4(0.f, 0.f, 0.f, __int_as_float(pd.w < 0.f ? 1 : 0));
        dp = make_float4(0.f, 0.f, 0.f, __int_as_float(-1));
    }
    B.posd[i] = pd; B.velp[i] = vp; B.accf[i] = af; B.dpi[i] = dp;
    carry[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    keys[i] = bin_id(d, pd.x, pd.y, pd.z);
}

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

extern "C" int64_t fsg_slab_message_bytes(int64_t m, int64_t g) { return (4 * m + 2 * g) * (int64_t)sizeof(float4); }

extern "C" int fsg_slab_pack(fsg_ctx *c, void *d_to_left, void *d_to_right, int64_t cap_bytes, int64_t counts[5])
{
    if (!c || !counts) return FSG_E_INVALID;
    if (c->cfg.world <= 1) { c->err = "fsg_slab_pack: not a slab context (world == 1)"; return FSG_E_STATE; }
    CUS(c, cudaSetDevice(c->device));
    const int64_t n = c->n;
    const int64_t nw = (n + 31) / 32 > 0 ? (n + 31) / 32 : 1;
    if (nw > c->slab_warps) {
        CUS(c, cudaStreamSynchronize(c->stream));
        cudaFree(c->slab_cnt); cudaFree(c->scan_tmp);
        c->slab_cnt = nullptr; c->scan_tmp = nullptr;
        const int64_t capw = (c->cap + 31) / 32 + 1;
        CUS(c, cudaMalloc(&c->slab_cnt, sizeof(int) * 2 * (4 * capw + 8) + 128));
        c->scan_tmp_bytes = fsg_scan_temp_bytes(4 * capw + 1);
        CUS(c, cudaMalloc(&c->scan_tmp, c->scan_tmp_bytes ? c->scan_tmp_bytes : 16));
        c->slab_warps = capw;
    }
    int *cnt = c->slab_cnt, *off = c->slab_cnt + (4 * c->slab_warps + 8);
    // the 5 totals live in the 8-byte aligned tail of the same allocation
    int64_t *d_out = reinterpret_cast<int64_t *>(reinterpret_cast<char *>(c->slab_cnt) + sizeof(int) * 2 * (4 * c->slab_warps + 8));
    const unsigned blocks = (unsigned)((nw * 32 + 255) / 256);
    if (n > 0) {
        k_slab_count<<<blocks, 256, 0, c->stream>>>(c->dev, c->cfg.rank, c->cfg.world, n, c->keysB, cnt, nw, c->counters + 6);
        CUS(c, cudaGetLastError());
    } else {
        CUS(c, cudaMemsetAsync(cnt, 0, sizeof(int) * (4 * nw + 1), c->stream));
    }
    CUS(c, fsg_scan_exclusive(c->scan_tmp, c->scan_tmp_bytes, cnt, off, 4 * nw + 1, c->stream));
    k_slab_totals<<<1, 32, 0, c->stream>>>(off, nw, c->counters + 5, c->steps > 0 ? 1 : 0, n, c->counters + 6, d_out);
    CUS(c, cudaGetLastError());
    c->launches += 2;
    int64_t h[6];
    CUS(c, cudaMemcpyAsync(h, d_out, sizeof h, cudaMemcpyDeviceToHost, c->stream));
    CUS(c, cudaStreamSynchronize(c->stream));
    for (int k = 0; k < 5; k++) counts[k] = h[k];
    if (h[5]) {
        c->err = "fsg_slab_pack: a particle moved more than one bin layer in a step and left the one-layer ghost band "
                 "(solver-unidyn.cu:187 makes the same assumption); reduce dt or use fewer slabs";
        return FSG_E_STATE;
    }
    if (fsg_slab_message_bytes(h[0], h[1]) > cap_bytes || fsg_slab_message_bytes(h[2], h[3]) > cap_bytes) {
        c->err = "fsg_slab_pack: message buffer too small";
        return FSG_E_NOMEM;
    }
    if ((h[0] + h[1] > 0 && !d_to_left) || (h[2] + h[3] > 0 && !d_to_right)) return FSG_E_INVALID;
    if (n > 0 && h[0] + h[1] + h[2] + h[3] > 0) {
        k_slab_scatter<<<blocks, 256, 0, c->stream>>>(c->dev, c->cfg.rank, c->cfg.world, n, c->keysB, c->B, off, nw, d_to_left,
                                                      d_to_right);
        CUS(c, cudaGetLastError());
        c->launches++;
    }
    // slots in use: after a step the sort has moved the dead slots behind n_keep
    if (c->steps > 0 && h[4] <= c->n) c->n = h[4];
    counts[4] = c->n;
    return FSG_OK;
}

extern "C" int fsg_slab_unpack(fsg_ctx *c, const void *d_from_left, int64_t mig_left, int64_t ghost_left,
                               const void *d_from_right, int64_t mig_right, int64_t ghost_right)
{
    if (!c || mig_left < 0 || ghost_left < 0 || mig_right < 0 || ghost_right < 0) return FSG_E_INVALID;
    if (c->cfg.world <= 1) { c->err = "fsg_slab_unpack: not a slab context (world == 1)"; return FSG_E_STATE; }
    CUS(c, cudaSetDevice(c->device));
    const int64_t nl = mig_left + ghost_left, nr = mig_right + ghost_right;
    if (c->n + nl + nr > c->cap) { c->err = "fsg_slab_unpack: received particles exceed the context capacity"; return FSG_E_NOMEM; }
    if ((nl > 0 && !d_from_left) || (nr > 0 && !d_from_right)) return FSG_E_INVALID;
    if (nl > 0) {
        k_slab_unpack<<<(unsigned)((nl + 255) / 256), 256, 0, c->stream>>>(c->dev, d_from_left, mig_left, ghost_left, c->n, c->B,
                                                                         c->carryB, c->keysB);
        CUS(c, cudaGetLastError());
        c->launches++;
        c->n += nl;
    }
    if (nr > 0) {
        k_slab_unpack<<<(unsigned)((nr + 255) / 256), 256, 0, c->stream>>>(c->dev, d_from_right, mig_right, ghost_right, c->n, c->B,
                                                                         c->carryB, c->keysB);
        CUS(c, cudaGetLastError());
        c->launches++;
        c->n += nr;
    }
    return FSG_OK;
}
