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
#include "fsg_slab_common.cuh"

#include <cub/device/device_scan.cuh>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <limits.h>

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

// counts per warp of the compact index space: cnt[cat * nw + warp] (cnt is cleared by the caller; only non-zero counts are written)
__global__ void __launch_bounds__(256)
k_slab_count(FsgDev d, int rank, int world, int64_t n, const int *__restrict__ keys, const int *__restrict__ region,
             const int *__restrict__ n_keep, int *__restrict__ cnt, int64_t nw, int *violation)
{
    const SlabRegion R = slab_region(region, n_keep, n);
    const int lane = threadIdx.x & 31;
    const int64_t warps = ((int64_t)gridDim.x * blockDim.x) >> 5, nwc = (R.total + 31) >> 5;
    for (int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; w < nwc; w += warps) {
        const int64_t t = w * 32 + lane;
        int c = 0;
        if (t < R.total) {
            const int key = keys[slab_slot(R, t)];
            c = slab_category(d, key, rank, world);
            // a particle that moved more than one bin layer in a step has left the one-layer ghost band
            if (key < d.numcells && (key / d.G2 < d.x0 - 1 || key / d.G2 > d.x1)) atomicOr(violation, 1);
        }
        if (!__any_sync(FULL, c != 0)) continue;
#pragma unroll
        for (int k = 0; k < 4; k++) {
            unsigned m = __ballot_sync(FULL, (c >> k) & 1);
            if (lane == 0 && m) cnt[k * nw + w] = __popc(m);
        }
    }
}

// Fixed-layout message (device memory), the same size on every rank so that nothing on the host depends
// on how many particles cross a face this step:
//   [header 64 B: int64 m, int64 g][posd cap_m][velp cap_m][accf cap_m][dpi cap_m]{mix cap_m}[posd cap_g][velp cap_g]{mix cap_g}[tail 64 B: int64 stamp]
// ({..}: unidyn model only — the volume fractions, FluidGPU-unidyn.cuh:180-181, which the pair terms of a neighbour read).
// The stamp (exchange sequence number) is copied AFTER the rest of the message, so a receiver that sees it
// has the whole message.
struct SlabMsg {
    long long *hdr;
    float4 *m_posd, *m_velp, *m_accf, *m_dpi, *m_mix, *g_posd, *g_velp, *g_mix;
};
__host__ __device__ inline int64_t slab_body_float4(bool mix, int64_t cap_m, int64_t cap_g) { return (mix ? 5 : 4) * cap_m + (mix ? 3 : 2) * cap_g; }
__host__ __device__ inline SlabMsg slab_msg(void *base, int64_t cap_m, int64_t cap_g, bool mix)
{
    SlabMsg r;
    r.hdr = (long long *)base;
    float4 *p = (float4 *)((char *)base + 64);
    r.m_posd = p; r.m_velp = p + cap_m; r.m_accf = p + 2 * cap_m; r.m_dpi = p + 3 * cap_m;
    r.m_mix = mix ? p + 4 * cap_m : nullptr;
    float4 *g = p + (mix ? 5 : 4) * cap_m;
    r.g_posd = g; r.g_velp = g + cap_g;
    r.g_mix = mix ? g + 2 * cap_g : nullptr;
    return r;
}

// off = exclusive scan of cnt (length 4*nw + 1): totals of the four categories -> message headers (clamped to
// the message capacities; an overflow is flagged, the excess is not sent) and the diagnostics array
__global__ void k_slab_headers(const int *__restrict__ off, int64_t nw, void *to_left, void *to_right, int64_t cap_m, int64_t cap_g,
                               int *overflow, long long *diag, long long stamp, bool mix)
{
    if (threadIdx.x != 0) return;
    const int64_t tail = (64 + slab_body_float4(mix, cap_m, cap_g) * 16) / 8;
    if (to_left) ((long long *)to_left)[tail] = stamp;
    if (to_right) ((long long *)to_right)[tail] = stamp;
    long long t[4];
    for (int k = 0; k < 4; k++) t[k] = off[(k + 1) * nw] - off[k * nw];
    for (int k = 0; k < 4; k++) diag[k] = t[k];
    if (t[0] > cap_m || t[2] > cap_m || t[1] > cap_g || t[3] > cap_g) atomicOr(overflow, 1);
    if (to_left) { long long *h = (long long *)to_left; h[0] = t[0] < cap_m ? t[0] : cap_m; h[1] = t[1] < cap_g ? t[1] : cap_g; }
    if (to_right) { long long *h = (long long *)to_right; h[0] = t[2] < cap_m ? t[2] : cap_m; h[1] = t[3] < cap_g ? t[3] : cap_g; }
}

__global__ void __launch_bounds__(256)
k_slab_scatter(FsgDev d, int rank, int world, int64_t n, const int *__restrict__ keys, const int *__restrict__ region,
               const int *__restrict__ n_keep, FsgState B, const int *__restrict__ off, int64_t nw, void *to_left, void *to_right,
               int64_t cap_m, int64_t cap_g)
{
    const SlabRegion R = slab_region(region, n_keep, n);
    const int lane = threadIdx.x & 31;
    const unsigned lt = (1u << lane) - 1u;
    const bool mix = B.mix != nullptr;
    const int64_t warps = ((int64_t)gridDim.x * blockDim.x) >> 5, nwc = (R.total + 31) >> 5;
    for (int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; w < nwc; w += warps) {
        const int64_t t = w * 32 + lane;
        const int64_t i = t < R.total ? slab_slot(R, t) : 0;
        const int c = t < R.total ? slab_category(d, keys[i], rank, world) : 0;
        if (!__any_sync(FULL, c != 0)) continue;
        int pos[4];
#pragma unroll
        for (int k = 0; k < 4; k++) {
            unsigned m = __ballot_sync(FULL, (c >> k) & 1);
            pos[k] = off[k * nw + w] - off[k * nw] + __popc(m & lt);
        }
        if (!c) continue;
        float4 pd = B.posd[i], vp = B.velp[i], mx = mix ? B.mix[i] : make_float4(0.f, 0.f, 0.f, 0.f);
        if (c & 5) {
            // migrant: full state goes to the neighbour.  The slot is NOT freed here: a particle moves less
            // than one bin per step, so it lands in the neighbour's outermost layer, where this slab still
            // needs it as a candidate for one more step.  Its bin is outside [x0, x1), so it is treated as
            // a ghost (never a home particle) and k_update drops it.
            SlabMsg M = slab_msg((c & 1) ? to_left : to_right, cap_m, cap_g, mix);
            int q = (c & 1) ? pos[0] : pos[2];
            if (q < cap_m) {
                M.m_posd[q] = pd;
                M.m_velp[q] = vp;
                M.m_accf[q] = B.accf[i];
                M.m_dpi[q] = B.dpi[i];
                if (mix) M.m_mix[q] = mx;
            }
        }
        if ((c & 2) && pos[1] < cap_g) {
            SlabMsg M = slab_msg(to_left, cap_m, cap_g, mix);
            M.g_posd[pos[1]] = pd; M.g_velp[pos[1]] = vp;
            if (mix) M.g_mix[pos[1]] = mx;
        }
        if ((c & 8) && pos[3] < cap_g) {
            SlabMsg M = slab_msg(to_right, cap_m, cap_g, mix);
            M.g_posd[pos[3]] = pd; M.g_velp[pos[3]] = vp;
            if (mix) M.g_mix[pos[3]] = mx;
        }
    }
}

// Appends both received messages behind the slots in use (*n_used, a device-side count: nothing here needs
// the host to know how many particles arrived): migrants (full state) then ghosts (read state); keys from
// the positions, exactly as the owner computed them (bin_id is the same function on both sides).
__global__ void __launch_bounds__(256)
k_slab_unpack(FsgDev d, const void *from_left, const void *from_right, int64_t cap_m, int64_t cap_g, const int *n_used, int64_t cap,
              FsgState B, float4 *carry, int *keys, int *overflow, long long *diag)
{
    const int64_t per = cap_m + cap_g;
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= 2 * per) return;
    const int side = t >= per;
    const void *msg = side ? from_right : from_left;
    const long long *hl = (const long long *)from_left, *hr = (const long long *)from_right;
    const long long nl = from_left ? hl[0] + hl[1] : 0;
    if (t == 0) { diag[4] = from_left ? hl[0] : 0; diag[5] = from_left ? hl[1] : 0; diag[6] = from_right ? hr[0] : 0; diag[7] = from_right ? hr[1] : 0; }
    if (!msg) return;
    if (*overflow & 4) return;        // the wait for this message timed out: what the inbox holds is two steps old
    const bool mix = B.mix != nullptr;
    SlabMsg M = slab_msg(const_cast<void *>(msg), cap_m, cap_g, mix);
    const long long m = M.hdr[0], g = M.hdr[1];
    int64_t u = t - (side ? per : 0);
    if (u >= m + g) return;
    int64_t i = (int64_t)*n_used + (side ? nl : 0) + u;
    if (i >= cap) { atomicOr(overflow, 2); return; }
    float4 pd, vp, af, dp;
    if (mix) B.mix[i] = u < m ? M.m_mix[u] : M.g_mix[u - m];
    if (u < m) { pd = M.m_posd[u]; vp = M.m_velp[u]; af = M.m_accf[u]; dp = M.m_dpi[u]; }
    else {
        pd = M.g_posd[u - m]; vp = M.g_velp[u - m];
        af = make_float4(0.f, 0.f, 0.f, __int_as_float(pd.w < 0.f ? 1 : 0));
        dp = make_float4(0.f, 0.f, 0.f, __int_as_float(-1));
    }
    B.posd[i] = pd; B.velp[i] = vp; B.accf[i] = af; B.dpi[i] = dp;
    carry[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    keys[i] = bin_id(d, pd.x, pd.y, pd.z);
}

extern "C" int64_t fsg_slab_message_bytes(int64_t cap_m, int64_t cap_g) { return 64 + slab_body_float4(false, cap_m, cap_g) * (int64_t)sizeof(float4) + 64; }
extern "C" int64_t fsg_slab_message_bytes_model(int model, int64_t cap_m, int64_t cap_g)
{
    return 64 + slab_body_float4(model == FSG_MODEL_UNIDYN, cap_m, cap_g) * (int64_t)sizeof(float4) + 64;
}
static size_t slab_bytes(const fsg_ctx *c, int64_t cap_m, int64_t cap_g) { return (size_t)fsg_slab_message_bytes_model(c->cfg.model, cap_m, cap_g); }

// Waits (on the device) until both neighbours' messages number `expected` have landed in this rank's inboxes:
// the stamp is the last thing a sender copies.  One thread; bounded by WALL-CLOCK time (%globaltimer, nanoseconds;
// fsg_ctx::slab_timeout_ns, FSG_SLAB_TIMEOUT_MS in the environment, default 30 s — rank skew under a profiler's
// kernel replay or a first-touch IPC mapping is seconds, not microseconds).  A missing neighbour raises flag 4 in the
// device-side flags (k_slab_unpack then appends NOTHING: the stale message of two steps ago is not integrated) and in
// the host-mapped flag word, which makes every following fsg_slab_* / fsg_step call return FSG_E_STATE.
__device__ __forceinline__ unsigned long long fsg_globaltimer()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__global__ void k_slab_wait(const volatile long long *tail_left, const volatile long long *tail_right, long long expected, int *flags,
                            volatile int *host_flag, unsigned long long timeout_ns)
{
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const unsigned long long t0 = fsg_globaltimer();
    for (;;) {
        bool ok = (!tail_left || *tail_left >= expected) && (!tail_right || *tail_right >= expected);
        if (ok) break;
        if (fsg_globaltimer() - t0 > timeout_ns) {
            atomicOr(flags, 4);
            if (host_flag) *host_flag = 4;
            break;
        }
        __nanosleep(200);
    }
    __threadfence_system();
}

int fsg_slab_ensure_counts(fsg_ctx *c, int64_t nw)
{
    if (nw <= c->slab_warps) return FSG_OK;
    CUS(c, cudaStreamSynchronize(c->stream));
    cudaFree(c->slab_cnt); cudaFree(c->scan_tmp);
    c->slab_cnt = nullptr; c->scan_tmp = nullptr;
    const int64_t capw = (c->cap + 31) / 32 + 1;
    CUS(c, cudaMalloc(&c->slab_cnt, sizeof(int) * 2 * (4 * capw + 8) + 128));
    CUS(c, cudaMemsetAsync(reinterpret_cast<char *>(c->slab_cnt) + sizeof(int) * 2 * (4 * capw + 8), 0, 128, c->stream));
    c->scan_tmp_bytes = fsg_scan_temp_bytes(4 * capw + 1);
    CUS(c, cudaMalloc(&c->scan_tmp, c->scan_tmp_bytes ? c->scan_tmp_bytes : 16));
    c->slab_warps = capw;
    return FSG_OK;
}
long long *fsg_slab_diag(fsg_ctx *c)
{
    return reinterpret_cast<long long *>(reinterpret_cast<char *>(c->slab_cnt) + sizeof(int) * 2 * (4 * c->slab_warps + 8));
}
cudaError_t fsg_launch_slab_wait(fsg_ctx *c, const long long *tail_left, const long long *tail_right, long long expected, cudaStream_t s);

cudaError_t fsg_launch_slab_wait(fsg_ctx *c, const long long *tail_left, const long long *tail_right, long long expected, cudaStream_t s)
{
    k_slab_wait<<<1, 32, 0, s>>>(tail_left, tail_right, expected, c->counters + 9, c->host_flag_dev, c->slab_timeout_ns);
    return cudaGetLastError();
}

// counters: [5] slots in use (device-side), [6] ghost-band violation, [9] message / capacity overflow
static int slab_pack_on(fsg_ctx *c, void *d_to_left, void *d_to_right, int64_t cap_m, int64_t cap_g, const int *region, long long stamp,
                        cudaStream_t st)
{
    if (!c || cap_m < 0 || cap_g < 0) return FSG_E_INVALID;
    if (c->cfg.world <= 1) { c->err = "fsg_slab_pack: not a slab context (world == 1)"; return FSG_E_STATE; }
    if ((c->cfg.rank > 0 && !d_to_left) || (c->cfg.rank < c->cfg.world - 1 && !d_to_right)) return FSG_E_INVALID;
    CUS(c, cudaSetDevice(c->device));
    const int64_t n = c->n;                           // == capacity for a slab context: unused slots hold the dead key
    const int64_t nw = (n + 31) / 32 > 0 ? (n + 31) / 32 : 1;
    if (int rc = fsg_slab_ensure_counts(c, nw)) return rc;
    int *cnt = c->slab_cnt, *off = c->slab_cnt + (4 * c->slab_warps + 8);
    long long *diag = fsg_slab_diag(c);
    // a fixed-size grid strides over the compact index space (its length is only known on the device)
    int64_t want = (nw * 32 + 255) / 256;
    const int64_t cap_blocks = (int64_t)c->sm_count * 16;
    const unsigned blocks = (unsigned)(want < cap_blocks ? (want > 0 ? want : 1) : cap_blocks);
    const int *n_keep = c->counters + 5;
    CUS(c, cudaMemsetAsync(cnt, 0, sizeof(int) * (size_t)(4 * nw + 1), st));
    k_slab_count<<<blocks, 256, 0, st>>>(c->dev, c->cfg.rank, c->cfg.world, n, c->keysB, region, n_keep, cnt, nw, c->counters + 6);
    CUS(c, cudaGetLastError());
    CUS(c, fsg_scan_exclusive(c->scan_tmp, c->scan_tmp_bytes, cnt, off, 4 * nw + 1, st));
    k_slab_headers<<<1, 32, 0, st>>>(off, nw, c->cfg.rank > 0 ? d_to_left : nullptr, c->cfg.rank < c->cfg.world - 1 ? d_to_right : nullptr,
                                     cap_m, cap_g, c->counters + 9, diag, stamp, c->B.mix != nullptr);
    CUS(c, cudaGetLastError());
    k_slab_scatter<<<blocks, 256, 0, st>>>(c->dev, c->cfg.rank, c->cfg.world, n, c->keysB, region, n_keep, c->B, off, nw, d_to_left, d_to_right,
                                           cap_m, cap_g);
    CUS(c, cudaGetLastError());
    c->launches += 3;
    return FSG_OK;
}

extern "C" int fsg_slab_pack(fsg_ctx *c, void *d_to_left, void *d_to_right, int64_t cap_m, int64_t cap_g)
{
    if (!c) return FSG_E_INVALID;
    return slab_pack_on(c, d_to_left, d_to_right, cap_m, cap_g, nullptr, 0, c->stream);
}

extern "C" int fsg_slab_unpack(fsg_ctx *c, const void *d_from_left, const void *d_from_right, int64_t cap_m, int64_t cap_g)
{
    if (!c || cap_m < 0 || cap_g < 0) return FSG_E_INVALID;
    if (c->cfg.world <= 1) { c->err = "fsg_slab_unpack: not a slab context (world == 1)"; return FSG_E_STATE; }
    CUS(c, cudaSetDevice(c->device));
    if (c->cfg.rank == 0) d_from_left = nullptr;
    if (c->cfg.rank == c->cfg.world - 1) d_from_right = nullptr;
    if (!c->slab_cnt) { c->err = "fsg_slab_unpack: call fsg_slab_pack first"; return FSG_E_STATE; }
    long long *diag = fsg_slab_diag(c);
    const int64_t threads = 2 * (cap_m + cap_g);
    if (threads > 0) {
        k_slab_unpack<<<(unsigned)((threads + 255) / 256), 256, 0, c->stream>>>(c->dev, d_from_left, d_from_right, cap_m, cap_g,
                                                                              c->counters + 5, c->cap, c->B, c->carryB, c->keysB,
                                                                              c->counters + 9, diag);
        CUS(c, cudaGetLastError());
        c->launches++;
    }
    return FSG_OK;
}

// Synchronises and reports: info[0..3] particles sent (migrants / ghosts to the left, to the right) and
// info[4..7] received in the last round, info[8] slots in use after the last sort.  FSG_E_STATE if a particle
// left the one-layer ghost band, FSG_E_NOMEM if a message or the particle capacity overflowed.
extern "C" int fsg_slab_check(fsg_ctx *c, int64_t info[9])
{
    if (!c) return FSG_E_INVALID;
    if (c->cfg.world <= 1) { c->err = "fsg_slab_check: not a slab context (world == 1)"; return FSG_E_STATE; }
    CUS(c, cudaSetDevice(c->device));
    int cnt[16];
    long long diag[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (c->comm) CUS(c, cudaStreamSynchronize(c->comm));
    CUS(c, cudaMemcpyAsync(cnt, c->counters, sizeof cnt, cudaMemcpyDeviceToHost, c->stream));
    if (c->slab_cnt) {
        CUS(c, cudaMemcpyAsync(diag, fsg_slab_diag(c), sizeof diag, cudaMemcpyDeviceToHost, c->stream));
    }
    CUS(c, cudaStreamSynchronize(c->stream));
    if (info) {
        for (int k = 0; k < 8; k++) info[k] = diag[k];
        info[8] = cnt[5];
    }
    if (cnt[6]) {
        c->err = "slab exchange: a particle moved more than one bin layer in a step and left the one-layer ghost band "
                 "(solver-unidyn.cu:187 makes the same assumption); reduce dt or use fewer slabs";
        return FSG_E_STATE;
    }
    if (cnt[9]) {
        c->err = (cnt[9] & 8) ? "slab exchange: a neighbour sent migrants in a different update state (every rank has to make the same calls)"
                 : (cnt[9] & 4) ? "slab exchange: timed out waiting for a neighbour's message"
                 : (cnt[9] & 2) ? "slab exchange: received particles exceed the context capacity"
                              : "slab exchange: a message exceeded its capacity (cap_m / cap_g)";
        return FSG_E_NOMEM;
    }
    return FSG_OK;
}


// ------------------------------------------------------------------------------------------------
// Peer-memory exchange: the neighbour's inboxes are mapped into this process (CUDA IPC), and the
// packed messages are copied straight into them over NVLink by the copy engines (no staging through a
// communication library, no SMs taken from the pair kernel).  The caller only has to order the
// neighbour's stream behind the copy — a few-byte NCCL send/recv on the same stream does that.
// Inboxes are double-buffered by step parity: a message for step s+1 never lands in memory the
// neighbour may still be unpacking for step s.
// ------------------------------------------------------------------------------------------------
static int slab_alloc_messages(fsg_ctx *c, int64_t cap_m, int64_t cap_g, int flags)
{
    if (!c || cap_m < 0 || cap_g < 0) return FSG_E_INVALID;
    if (c->cfg.world <= 1) { c->err = "fsg_slab_alloc_messages: not a slab context (world == 1)"; return FSG_E_STATE; }
    CUS(c, cudaSetDevice(c->device));
    CUS(c, cudaStreamSynchronize(c->stream));
    for (int k = 0; k < 2; k++) { cudaFree(c->outbox[k]); c->outbox[k] = nullptr; }
    for (int k = 0; k < 4; k++) { cudaFree(c->inbox[k]); c->inbox[k] = nullptr; }
    size_t bytes = slab_bytes(c, cap_m, cap_g);
    // the sorted-ghost pipeline (fsg_slab2.cu) where it applies, unless the caller asked for the classic one (flags bit 0)
    c->slab2 = false;
    c->dev.kx0 = c->dev.kx1 = INT_MIN;
    if (!(flags & 1)) { int rc = fsg_slab2_engage(c, cap_m, cap_g, &bytes); if (rc != FSG_OK) return rc; }
    for (int k = 0; k < 2; k++) { CUS(c, cudaMalloc(&c->outbox[k], bytes)); CUS(c, cudaMemsetAsync(c->outbox[k], 0, bytes, c->stream)); }
    for (int k = 0; k < 4; k++) { CUS(c, cudaMalloc(&c->inbox[k], bytes)); CUS(c, cudaMemsetAsync(c->inbox[k], 0, bytes, c->stream)); }
    CUS(c, cudaStreamSynchronize(c->stream));
    c->msg_cap_m = cap_m;
    c->msg_cap_g = cap_g;
    if (!c->host_flag) {              // host-visible flag word the device-side wait raises on a timeout
        CUS(c, cudaHostAlloc((void **)&c->host_flag, 64, cudaHostAllocMapped));
        *c->host_flag = 0;
        CUS(c, cudaHostGetDevicePointer((void **)&c->host_flag_dev, (void *)c->host_flag, 0));
    }
    const char *te = getenv("FSG_SLAB_TIMEOUT_MS");
    const double tms = te ? atof(te) : 30000.0;
    c->slab_timeout_ns = (unsigned long long)((tms > 1.0 ? tms : 1.0) * 1e6);
    return FSG_OK;
}

extern "C" int fsg_slab_alloc_messages(fsg_ctx *c, int64_t cap_m, int64_t cap_g) { return slab_alloc_messages(c, cap_m, cap_g, 0); }
extern "C" int fsg_slab_alloc_messages2(fsg_ctx *c, int64_t cap_m, int64_t cap_g, int flags) { return slab_alloc_messages(c, cap_m, cap_g, flags); }
extern "C" int fsg_slab_mode(fsg_ctx *c) { return !c || c->cfg.world <= 1 ? 0 : c->slab2 ? 2 : 1; }

// side 0: the inbox that receives from the LEFT neighbour, side 1: from the RIGHT one; parity 0/1
extern "C" int fsg_slab_inbox_handle(fsg_ctx *c, int side, int parity, void *handle64)
{
    if (!c || !handle64 || side < 0 || side > 1 || parity < 0 || parity > 1) return FSG_E_INVALID;
    if (!c->inbox[2 * side + parity]) { c->err = "fsg_slab_inbox_handle: call fsg_slab_alloc_messages first"; return FSG_E_STATE; }
    CUS(c, cudaSetDevice(c->device));
    cudaIpcMemHandle_t h;
    CUS(c, cudaIpcGetMemHandle(&h, c->inbox[2 * side + parity]));
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handles are 64 bytes");
    memcpy(handle64, &h, 64);
    return FSG_OK;
}

// side 0: `handle64` is the LEFT neighbour's from-right inbox, side 1: the RIGHT neighbour's from-left inbox
extern "C" int fsg_slab_open_peer(fsg_ctx *c, int side, int parity, const void *handle64)
{
    if (!c || !handle64 || side < 0 || side > 1 || parity < 0 || parity > 1) return FSG_E_INVALID;
    CUS(c, cudaSetDevice(c->device));
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, 64);
    void *p = nullptr;
    CUS(c, cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    c->peer_inbox[2 * side + parity] = p;
    return FSG_OK;
}

// pack on `st`, copy both messages into the neighbours' inboxes (stamp last)
static int slab_send_on(fsg_ctx *c, const int *region, cudaStream_t st)
{
    if (!c->outbox[0]) { c->err = "slab send: call fsg_slab_alloc_messages first"; return FSG_E_STATE; }
    const long long seq = ++c->seq_send;
    const int par = (int)(seq & 1);
    const bool left = c->cfg.rank > 0, right = c->cfg.rank < c->cfg.world - 1;
    if ((left && !c->peer_inbox[par]) || (right && !c->peer_inbox[2 + par])) {
        c->err = "slab send: the neighbours' inboxes are not mapped (fsg_slab_open_peer)";
        return FSG_E_STATE;
    }
    int rc = slab_pack_on(c, c->outbox[0], c->outbox[1], c->msg_cap_m, c->msg_cap_g, region, seq, st);
    if (rc != FSG_OK) return rc;
    const size_t bytes = slab_bytes(c, c->msg_cap_m, c->msg_cap_g), body = bytes - 64;
    for (int side = 0; side < 2; side++) {
        if (side == 0 ? !left : !right) continue;
        char *dst = (char *)c->peer_inbox[2 * side + par], *src = (char *)c->outbox[side];
        CUS(c, cudaMemcpyAsync(dst, src, body, cudaMemcpyDefault, st));
        CUS(c, cudaMemcpyAsync(dst + body, src + body, 8, cudaMemcpyDefault, st));     // the stamp, after the body
    }
    return FSG_OK;
}

// A wait that timed out (k_slab_wait) is sticky: the state has been integrated without its neighbours.
int fsg_slab_sticky_error(fsg_ctx *c)
{
    if (c->host_flag && *c->host_flag) {
        c->err = "slab exchange: timed out waiting for a neighbour's message; the state of this context is no longer valid "
                 "(upload again; FSG_SLAB_TIMEOUT_MS sets the limit)";
        return FSG_E_STATE;
    }
    return FSG_OK;
}

extern "C" int fsg_slab_pack_send(fsg_ctx *c)
{
    if (!c) return FSG_E_INVALID;
    if (int rc = fsg_slab_sticky_error(c)) return rc;
    if (c->sent_ahead) { c->sent_ahead = false; return FSG_OK; }      // fsg_step already issued this exchange (overlap mode)
    // (before the first step the particles are in upload order: every slot is looked at)
    if (c->slab2) return fsg_slab2_pack_send(c);
    return slab_send_on(c, c->steps > 0 ? c->counters + 16 : nullptr, c->stream);
}

// overlap mode, called by fsg_step between the boundary and the interior bins: the next step's messages are
// packed and copied on the communication stream, behind the update of the boundary particles
int fsg_slab_send_next(fsg_ctx *c)
{
    CUS(c, cudaEventRecord(c->ev_boundary, c->stream));
    CUS(c, cudaStreamWaitEvent(c->comm, c->ev_boundary, 0));
    int rc = slab_send_on(c, c->counters + 16, c->comm);
    if (rc != FSG_OK) return rc;
    CUS(c, cudaEventRecord(c->ev_sent, c->comm));
    c->sent_ahead = true;
    return FSG_OK;
}

extern "C" int fsg_slab_unpack_recv(fsg_ctx *c)
{
    if (!c) return FSG_E_INVALID;
    if (!c->inbox[0]) { c->err = "fsg_slab_unpack_recv: call fsg_slab_alloc_messages first"; return FSG_E_STATE; }
    if (int rc = fsg_slab_sticky_error(c)) return rc;
    CUS(c, cudaSetDevice(c->device));
    if (c->slab2) return fsg_slab2_unpack_recv(c);
    const long long seq = ++c->seq_recv;
    const int par = (int)(seq & 1);
    const bool left = c->cfg.rank > 0, right = c->cfg.rank < c->cfg.world - 1;
    if (c->overlap) CUS(c, cudaStreamWaitEvent(c->stream, c->ev_sent, 0));   // my own pack (other stream) reads the slots unpack writes
    const size_t tail = (slab_bytes(c, c->msg_cap_m, c->msg_cap_g) - 64);
    CUS(c, fsg_launch_slab_wait(c, left ? (const long long *)((char *)c->inbox[par] + tail) : nullptr,
                                right ? (const long long *)((char *)c->inbox[2 + par] + tail) : nullptr, seq, c->stream));
    c->launches++;
    return fsg_slab_unpack(c, c->inbox[par], c->inbox[2 + par], c->msg_cap_m, c->msg_cap_g);
}

// In-process wiring (several slab contexts in ONE process, e.g. tests on one GPU): the neighbour's inbox is a
// plain device pointer, no IPC mapping.  side / parity as fsg_slab_open_peer.
extern "C" int fsg_slab_set_peer(fsg_ctx *c, int side, int parity, void *neighbour_inbox)
{
    if (!c || side < 0 || side > 1 || parity < 0 || parity > 1) return FSG_E_INVALID;
    c->peer_inbox[2 * side + parity] = neighbour_inbox;
    c->peer_local = true;
    return FSG_OK;
}
extern "C" void *fsg_slab_inbox_ptr(fsg_ctx *c, int side, int parity)
{
    if (!c || side < 0 || side > 1 || parity < 0 || parity > 1) return nullptr;
    return c->inbox[2 * side + parity];
}

// on: fsg_step computes the slab's boundary bins first and issues the NEXT step's pack + copies on a second
// stream, beside the interior bins (needs the peer-memory exchange)
extern "C" int fsg_slab_set_overlap(fsg_ctx *c, int on)
{
    if (!c) return FSG_E_INVALID;
    if (c->cfg.world <= 1) { c->err = "fsg_slab_set_overlap: not a slab context (world == 1)"; return FSG_E_STATE; }
    if (on && !c->outbox[0]) { c->err = "fsg_slab_set_overlap: needs the peer-memory exchange (fsg_slab_alloc_messages)"; return FSG_E_STATE; }
    if (on && c->cfg.model != FSG_MODEL_BASE) { c->err = "fsg_slab_set_overlap: base model only"; return FSG_E_UNSUPPORTED; }
    if (on && c->slab2) {
        c->err = "fsg_slab_set_overlap: the messages were allocated for the sorted-ghost pipeline; use fsg_slab_alloc_messages2(.., 1)";
        return FSG_E_STATE;
    }
    CUS(c, cudaSetDevice(c->device));
    CUS(c, cudaStreamSynchronize(c->stream));
    if (on && !c->comm) {
        int lo = 0, hi = 0;
        cudaDeviceGetStreamPriorityRange(&lo, &hi);
        CUS(c, cudaStreamCreateWithPriority(&c->comm, cudaStreamNonBlocking, hi));
        CUS(c, cudaEventCreateWithFlags(&c->ev_boundary, cudaEventDisableTiming));
        CUS(c, cudaEventCreateWithFlags(&c->ev_sent, cudaEventDisableTiming));
        CUS(c, cudaEventRecord(c->ev_sent, c->comm));
        const int64_t nb = c->cap < c->dev.numcells ? c->cap : c->dev.numcells;
        CUS(c, cudaMalloc(&c->binlistB, sizeof(int) * (nb > 0 ? nb : 1)));
    }
    if (c->comm) CUS(c, cudaStreamSynchronize(c->comm));
    c->overlap = on != 0;
    // interior layers: everything further than two layers from a face that has a neighbour
    c->dev.bx0 = c->dev.x0 + ((on && c->cfg.rank > 0) ? 2 : 0);
    c->dev.bx1 = c->dev.x1 - ((on && c->cfg.rank < c->cfg.world - 1) ? 2 : 0);
    if (c->dev.bx1 < c->dev.bx0) c->dev.bx1 = c->dev.bx0;
    fsg_update_pair_mode(c);
    return FSG_OK;
}

extern "C" int fsg_slab_close_peers(fsg_ctx *c)
{
    if (!c) return FSG_E_INVALID;
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    if (c->comm) cudaStreamSynchronize(c->comm);
    for (int k = 0; k < 4; k++) if (c->peer_inbox[k]) { if (!c->peer_local) cudaIpcCloseMemHandle(c->peer_inbox[k]); c->peer_inbox[k] = nullptr; }
    return FSG_OK;
}
