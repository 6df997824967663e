// fsg_slab2.cu — the sorted-ghost slab pipeline: the multi-device hand-off of solver-unidyn.cu:396-470 (find_idx / host-staged
// cudaMemcpy / mem_shift, FluidGPU-unidyn.cu:499-542) for N slabs, in the order that keeps ghosts OUT of the sort.
//
// The classic pipeline (fsg_slab.cu) appends the neighbours' face particles behind the slab's own every step, so that a third of the
// key array are movers for the sort and every ghost travels through the gather twice.  But a neighbour's face layer is a contiguous
// range of ITS sorted state, already in bin order.  Per step and slab:
//
//   fsg_slab_pack_send    migrants only: particles whose NEXT bin (known a step ahead, predicted_key) lies outside the slab leave with
//                         their pre-update state and their pending pair sums; the slot is freed at once
//   fsg_slab_unpack_recv  device-side wait for the neighbours' migrants, appended behind the slots in use
//   fsg_step              key sort (movers: bin changers + migrants, a fraction of a percent) -> reorder, applying the deferred update
//                         -> k_s2_ghost_send: the sorted face layers [first slot, first slot of layer x0 + 1) and [first slot of
//                            layer x1 - 1, n_live) go straight into the neighbours' memory (posd, velp, keys), stamp last
//                         -> wait for theirs -> k_s2_install: ghosts into the two zones at the top of the particle arrays, their bins
//                            into the tables (and, for the symmetric kernel, the lower ghost layer into the home-bin list)
//                         -> pair sums
//
// The top 2 * cap_g slots of every particle array are the ghost zones; the context works on n_own = capacity - 2 * cap_g slots.
// The particle array is therefore not contiguous across the bin ids x0 * G^2 and x1 * G^2 (FsgDev::kx0 / kx1): the pair kernels do not
// join a run across them (the bins on either side of those ids are wrap-around neighbours, FluidGPU.cu:124-126, never in range).
// Nothing waits for the host; every count is read on the device.
#include "fsg_slab_common.cuh"

#include <limits.h>
#include <stdlib.h>

// ---- message layouts (inside one library-owned buffer per neighbour, direction and parity: [migrants | ghosts]) ----
//   migrants: [header 64 B: int64 m, int64 deferred][posd cap_m][velp cap_m][accf cap_m][dpi cap_m][acc cap_m][tail 64 B: int64 stamp]
//   ghosts:   [header 64 B: int64 g][posd cap_g][velp cap_g][keys cap_g (int), padded to 64 B][tail 64 B: int64 stamp]
struct S2Mig {
    long long *hdr, *tail;
    float4 *posd, *velp, *accf, *dpi, *acc;
};
struct S2Gh {
    long long *hdr, *tail;
    float4 *posd, *velp;
    int *keys;
};
__host__ __device__ inline size_t s2_mig_bytes(int64_t cap_m) { return 64 + (size_t)cap_m * 5 * sizeof(float4) + 64; }
__host__ __device__ inline size_t s2_gh_keys_bytes(int64_t cap_g) { return ((size_t)cap_g * sizeof(int) + 63) & ~(size_t)63; }
__host__ __device__ inline size_t s2_gh_bytes(int64_t cap_g) { return 64 + (size_t)cap_g * 2 * sizeof(float4) + s2_gh_keys_bytes(cap_g) + 64; }
__host__ __device__ inline S2Mig s2_mig(void *base, int64_t cap_m)
{
    S2Mig r;
    r.hdr = (long long *)base;
    float4 *p = (float4 *)((char *)base + 64);
    r.posd = p; r.velp = p + cap_m; r.accf = p + 2 * cap_m; r.dpi = p + 3 * cap_m; r.acc = p + 4 * cap_m;
    r.tail = (long long *)(p + 5 * cap_m);
    return r;
}
__host__ __device__ inline S2Gh s2_gh(void *base, int64_t cap_g)
{
    S2Gh r;
    r.hdr = (long long *)base;
    float4 *p = (float4 *)((char *)base + 64);
    r.posd = p; r.velp = p + cap_g;
    r.keys = (int *)(p + 2 * cap_g);
    r.tail = (long long *)((char *)r.keys + s2_gh_keys_bytes(cap_g));
    return r;
}

// bit 0: leaves to the left, bit 1: leaves to the right
__device__ __forceinline__ int s2_category(const FsgDev &d, int key)
{
    if (key >= d.numcells) return 0;              // parked / dead
    const int ix = key / d.G2;
    return ix < d.x0 ? 1 : ix >= d.x1 ? 2 : 0;
}

__global__ void __launch_bounds__(256)
k_s2_count(FsgDev d, int rank, int world, int64_t n, const int *__restrict__ keys, const int *__restrict__ region,
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
            c = s2_category(d, key);
            // more than one bin layer in a step, or out through a face that has no neighbour: the one-layer ghost band is broken
            if (key < d.numcells && (key / d.G2 < d.x0 - 1 || key / d.G2 > d.x1)) atomicOr(violation, 1);
            if ((c == 1 && rank == 0) || (c == 2 && rank == world - 1)) atomicOr(violation, 1);
        }
        if (!__any_sync(FULL, c != 0)) continue;
#pragma unroll
        for (int k = 0; k < 2; k++) {
            const unsigned m = __ballot_sync(FULL, (c >> k) & 1);
            if (lane == 0 && m) cnt[k * nw + w] = __popc(m);
        }
    }
}

__global__ void k_s2_headers(const int *__restrict__ off, int64_t nw, void *to_left, void *to_right, int64_t cap_m, int deferred, int *overflow,
                             long long *diag, long long stamp)
{
    if (threadIdx.x != 0) return;
    long long t[2];
    for (int k = 0; k < 2; k++) t[k] = off[(k + 1) * nw] - off[k * nw];
    diag[0] = t[0];
    diag[2] = t[1];
    if (t[0] > cap_m || t[1] > cap_m) atomicOr(overflow, 1);
    if (to_left) { S2Mig M = s2_mig(to_left, cap_m); M.hdr[0] = t[0] < cap_m ? t[0] : cap_m; M.hdr[1] = deferred; M.tail[0] = stamp; }
    if (to_right) { S2Mig M = s2_mig(to_right, cap_m); M.hdr[0] = t[1] < cap_m ? t[1] : cap_m; M.hdr[1] = deferred; M.tail[0] = stamp; }
}

// acc: the accumulators that are still to be applied to the particle — the pair sums of the last step when its update is deferred
// (+ acc2, the accumulators carried in by an upload, while those are pending), the uploaded accumulators before the first step
__global__ void __launch_bounds__(256)
k_s2_scatter(FsgDev d, int64_t n, int *__restrict__ keys, const int *__restrict__ region, const int *__restrict__ n_keep, FsgState S,
             const float4 *__restrict__ acc, const float4 *__restrict__ acc2, const int *__restrict__ off, int64_t nw, void *to_left,
             void *to_right, int64_t cap_m)
{
    const SlabRegion R = slab_region(region, n_keep, n);
    const int lane = threadIdx.x & 31;
    const unsigned lt = (1u << lane) - 1u;
    const int64_t warps = ((int64_t)gridDim.x * blockDim.x) >> 5, nwc = (R.total + 31) >> 5;
    for (int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; w < nwc; w += warps) {
        const int64_t t = w * 32 + lane;
        const int64_t i = t < R.total ? slab_slot(R, t) : 0;
        const int c = t < R.total ? s2_category(d, keys[i]) : 0;
        if (!__any_sync(FULL, c != 0)) continue;
        int pos[2];
#pragma unroll
        for (int k = 0; k < 2; k++) {
            const unsigned m = __ballot_sync(FULL, (c >> k) & 1);
            pos[k] = off[k * nw + w] - off[k * nw] + __popc(m & lt);
        }
        if (!c) continue;
        void *msg = c == 1 ? to_left : to_right;
        const int q = c == 1 ? pos[0] : pos[1];
        if (!msg || q >= cap_m) continue;          // (flagged by k_s2_count / k_s2_headers; the particle stays where it is)
        S2Mig M = s2_mig(msg, cap_m);
        M.posd[q] = S.posd[i];
        M.velp[q] = S.velp[i];
        M.accf[q] = S.accf[i];
        M.dpi[q] = S.dpi[i];
        float4 a = acc ? acc[i] : make_float4(0.f, 0.f, 0.f, 0.f);
        if (acc2) { const float4 b = acc2[i]; a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w; }
        M.acc[q] = a;
        keys[i] = d.dead;                          // the slot is free: the particle lives at the neighbour from now on
    }
}

// Appends the migrants of both neighbours behind the slots in use.  deferred: S is the pre-update state, acc_dst the pair sums;
// the key is the bin after the pending update (predicted_key, the same bits the owner had).
__global__ void __launch_bounds__(256)
k_s2_unpack(FsgDev d, const void *from_left, const void *from_right, int64_t cap_m, const int *n_used, int64_t n_own, FsgState S,
            float4 *acc_dst, float4 *zero_dst, int *keys, int deferred, int *overflow, long long *diag)
{
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= 2 * cap_m) return;
    const int side = t >= cap_m;
    const void *msg = side ? from_right : from_left;
    const long long nl = from_left ? ((const long long *)from_left)[0] : 0;
    if (t == 0) { diag[4] = nl; diag[6] = from_right ? ((const long long *)from_right)[0] : 0; }
    if (!msg) return;
    if (*overflow & 4) return;        // the wait for this message timed out: what the inbox holds is two steps old
    S2Mig M = s2_mig(const_cast<void *>(msg), cap_m);
    const long long m = M.hdr[0];
    const int64_t u = t - (side ? cap_m : 0);
    if (u >= m) return;
    if (u == 0 && M.hdr[1] != deferred) atomicOr(overflow, 8);
    const int64_t i = (int64_t)*n_used + (side ? nl : 0) + u;
    if (i >= n_own) { atomicOr(overflow, 2); return; }
    const float4 pd = M.posd[u], vp = M.velp[u];
    S.posd[i] = pd; S.velp[i] = vp; S.accf[i] = M.accf[u]; S.dpi[i] = M.dpi[u];
    if (acc_dst) acc_dst[i] = M.acc[u];
    if (zero_dst) zero_dst[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    keys[i] = deferred ? predicted_key(d, pd, vp) : bin_id(d, pd.x, pd.y, pd.z);
}

// The face layers of the sorted state -> the neighbours' ghost messages (remote stores over NVLink / plain stores in one process).
// blockIdx.y = side (0: layer x0 to the left neighbour, 1: layer x1 - 1 to the right one).  The last block to finish writes the
// count and, after a system-wide fence, the stamp the receiver waits for.
__global__ void __launch_bounds__(256)
k_s2_ghost_send(FsgState A, const int *__restrict__ keysA, const int *__restrict__ ranges, const int *__restrict__ nlive, void *peer_left,
                void *peer_right, int64_t cap_g, long long seq, int *done, int *overflow, long long *diag)
{
    const int side = blockIdx.y;
    void *peer = side ? peer_right : peer_left;
    if (!peer) return;
    const int64_t live = *nlive;
    const int64_t lo = side ? min((int64_t)ranges[3], live) : 0;
    const int64_t hi = side ? live : min((int64_t)ranges[2], live);
    int64_t g = hi - lo;
    if (g > cap_g) g = cap_g;
    const S2Gh M = s2_gh(peer, cap_g);
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < g; t += (int64_t)gridDim.x * blockDim.x) {
        M.posd[t] = A.posd[lo + t];
        M.velp[t] = A.velp[lo + t];
        M.keys[t] = keysA[lo + t];
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        const int before = atomicAdd(&done[side], 1);
        if (before == (int)gridDim.x - 1) {
            done[side] = 0;
            if (hi - lo > cap_g) atomicOr(overflow, 1);
            M.hdr[0] = g;
            diag[side ? 3 : 1] = g;
            __threadfence_system();
            *(volatile long long *)M.tail = seq;
            __threadfence_system();
        }
    }
}

// Ghost messages -> the ghost zones of the sorted state + their bins in the tables; the lower ghost layer joins the home-bin list
// when the symmetric kernel runs (fsg_pair_v3.cu walks it restricted to its runs in layer x0).
__global__ void __launch_bounds__(256)
k_s2_install(FsgDev d, const void *from_left, const void *from_right, int64_t cap_g, int64_t n_own, FsgState A, int *start, int *end,
             int *binlist, int *nocc, int *overflow, int *violation, long long *diag)
{
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int side = t >= cap_g;
    const void *msg = side ? from_right : from_left;
    bool head = false;
    int key = 0;
    if (t < 2 * cap_g && msg && !(*overflow & 4)) {
        const S2Gh M = s2_gh(const_cast<void *>(msg), cap_g);
        const long long g = M.hdr[0];
        const int64_t u = t - (side ? cap_g : 0);
        if (u == 0) diag[side ? 7 : 5] = g;
        if (u < g) {
            const int64_t z = n_own + (side ? cap_g : 0) + u;
            A.posd[z] = M.posd[u];
            A.velp[z] = M.velp[u];
            key = M.keys[u];
            const int want = side ? d.x1 : d.x0 - 1;
            if (key < 0 || key >= d.numcells || key / d.G2 != want) atomicOr(violation, 1);
            else {
                const int prev = u > 0 ? M.keys[u - 1] : -1;
                const int next = u + 1 < g ? M.keys[u + 1] : -1;
                if (key != prev) { start[key] = (int)z; head = d.sym && side == 0; }
                if (key != next) end[key] = (int)z;
            }
        }
    }
    const unsigned m = __ballot_sync(FULL, head);
    if (m) {
        const int lane = threadIdx.x & 31;
        int base = 0;
        if (lane == 0) base = atomicAdd(nocc, __popc(m));
        base = __shfl_sync(FULL, base, 0);
        if (head) binlist[base + __popc(m & ((1u << lane) - 1u))] = key;
    }
}

__global__ void __launch_bounds__(256)
k_s2_reset_ghost(const void *from_left, const void *from_right, int64_t cap_g, int numcells, int *start, int *end)
{
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= 2 * cap_g) return;
    const int side = t >= cap_g;
    const void *msg = side ? from_right : from_left;
    if (!msg) return;
    const S2Gh M = s2_gh(const_cast<void *>(msg), cap_g);
    const int64_t u = t - (side ? cap_g : 0);
    if (u >= M.hdr[0]) return;
    const int key = M.keys[u];
    if (key >= 0 && key < numcells) { start[key] = -1; end[key] = -1; }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
int fsg_slab2_engage(fsg_ctx *c, int64_t cap_m, int64_t cap_g, size_t *bytes)
{
    const fsg_config &f = c->cfg;
    const char *e = getenv("FSG_SLAB2"), *ed = getenv("FSG_DEFER_UPDATE");
    const bool want = (!e || atoi(e) != 0) && (!ed || atoi(ed) != 0);
    const bool can = f.model == FSG_MODEL_BASE && f.pair_fp64 == 0 && f.neighbour_cap == 0 && f.bin_cap == 0 && f.grid >= 4 && !c->overlap &&
                     c->n == 0 && cap_g > 0 && c->cap - 2 * cap_g >= 64;
    if (!want || !can) return FSG_OK;
    c->slab2 = true;
    c->slab2_mid = false;
    c->n_own = c->cap - 2 * cap_g;
    c->gh_par_last = -1;
    c->dev.kx0 = c->dev.x0 * c->dev.G2;
    c->dev.kx1 = c->dev.x1 * c->dev.G2;
    c->mig_bytes = s2_mig_bytes(cap_m);
    c->gh_bytes = s2_gh_bytes(cap_g);
    *bytes = c->mig_bytes + c->gh_bytes;
    return FSG_OK;
}

static void *s2_gh_part(const fsg_ctx *c, void *buf) { return buf ? (char *)buf + c->mig_bytes : nullptr; }

int fsg_slab2_pack_send(fsg_ctx *c)
{
    if (!c->outbox[0]) { c->err = "slab send: call fsg_slab_alloc_messages first"; return FSG_E_STATE; }
    CUS(c, cudaSetDevice(c->device));
    const long long seq = ++c->seq_send;
    const int par = (int)(seq & 1);
    const bool left = c->cfg.rank > 0, right = c->cfg.rank < c->cfg.world - 1;
    if ((left && !c->peer_inbox[par]) || (right && !c->peer_inbox[2 + par])) {
        c->err = "slab send: the neighbours' inboxes are not mapped (fsg_slab_open_peer)";
        return FSG_E_STATE;
    }
    cudaStream_t st = c->stream;
    const int64_t n = c->n, cap_m = c->msg_cap_m;
    const int64_t nw = (n + 31) / 32 > 0 ? (n + 31) / 32 : 1;
    if (int rc = fsg_slab_ensure_counts(c, nw)) return rc;
    int *cnt = c->slab_cnt, *off = c->slab_cnt + (4 * c->slab_warps + 8);
    long long *diag = fsg_slab_diag(c);
    int64_t want = (nw * 32 + 255) / 256;
    const int64_t cap_blocks = (int64_t)c->sm_count * 16;
    const unsigned blocks = (unsigned)(want < cap_blocks ? (want > 0 ? want : 1) : cap_blocks);
    // (before the first step the particles are in upload order: every slot is looked at; afterwards only the two face layers —
    // a particle moves less than one bin layer per step, so a migrant was in layer x0 or x1 - 1)
    const int *region = c->steps > 0 ? c->counters + 18 : nullptr;
    const int *n_keep = c->counters + 5;
    // the state a migrant carries: pre-update + pending sums while the update is deferred, else the materialised state
    const FsgState S = c->deferred ? c->A : c->B;
    const float4 *acc = c->deferred ? c->sums : (c->carry_live ? c->carryB : nullptr);
    const float4 *acc2 = (c->deferred && c->carry_pending) ? c->carryA : nullptr;
    void *to_left = left ? c->outbox[0] : nullptr, *to_right = right ? c->outbox[1] : nullptr;
    CUS(c, cudaMemsetAsync(cnt, 0, sizeof(int) * (size_t)(2 * nw + 1), st));
    k_s2_count<<<blocks, 256, 0, st>>>(c->dev, c->cfg.rank, c->cfg.world, n, c->keysB, region, n_keep, cnt, nw, c->counters + 6);
    CUS(c, cudaGetLastError());
    CUS(c, fsg_scan_exclusive(c->scan_tmp, c->scan_tmp_bytes, cnt, off, 2 * nw + 1, st));
    k_s2_headers<<<1, 32, 0, st>>>(off, nw, to_left, to_right, cap_m, c->deferred ? 1 : 0, c->counters + 9, diag, seq);
    CUS(c, cudaGetLastError());
    k_s2_scatter<<<blocks, 256, 0, st>>>(c->dev, n, c->keysB, region, n_keep, S, acc, acc2, off, nw, to_left, to_right, cap_m);
    CUS(c, cudaGetLastError());
    c->launches += 3;
    const size_t body = c->mig_bytes - 64;
    for (int side = 0; side < 2; side++) {
        if (side == 0 ? !left : !right) continue;
        char *dst = (char *)c->peer_inbox[2 * side + par], *src = (char *)c->outbox[side];
        CUS(c, cudaMemcpyAsync(dst, src, body, cudaMemcpyDefault, st));
        CUS(c, cudaMemcpyAsync(dst + body, src + body, 8, cudaMemcpyDefault, st));     // the stamp, after the body
    }
    c->slab2_mid = true;
    return FSG_OK;
}

int fsg_slab2_unpack_recv(fsg_ctx *c)
{
    const long long seq = ++c->seq_recv;
    const int par = (int)(seq & 1);
    const bool left = c->cfg.rank > 0, right = c->cfg.rank < c->cfg.world - 1;
    if (!c->slab_cnt) { c->err = "fsg_slab_unpack_recv: call fsg_slab_pack_send first"; return FSG_E_STATE; }
    const size_t tail = c->mig_bytes - 64;
    CUS(c, fsg_launch_slab_wait(c, left ? (const long long *)((char *)c->inbox[par] + tail) : nullptr,
                                right ? (const long long *)((char *)c->inbox[2 + par] + tail) : nullptr, seq, c->stream));
    c->launches++;
    const int64_t cap_m = c->msg_cap_m;
    if (cap_m > 0) {
        const FsgState S = c->deferred ? c->A : c->B;
        float4 *acc_dst = c->deferred ? c->sums : (c->carry_live ? c->carryB : nullptr);
        float4 *zero_dst = (c->deferred && c->carry_pending) ? c->carryA : nullptr;
        k_s2_unpack<<<(unsigned)((2 * cap_m + 255) / 256), 256, 0, c->stream>>>(c->dev, left ? c->inbox[par] : nullptr, right ? c->inbox[2 + par] : nullptr,
                                                                             cap_m, c->counters + 5, c->n_own, S, acc_dst, zero_dst, c->keysB,
                                                                             c->deferred ? 1 : 0, c->counters + 9, fsg_slab_diag(c));
        CUS(c, cudaGetLastError());
        c->launches++;
    }
    return FSG_OK;
}

int fsg_slab2_reset_ghost_tables(fsg_ctx *c)
{
    if (c->gh_par_last < 0 || !c->inbox[0]) return FSG_OK;
    const int par = c->gh_par_last;
    const bool left = c->cfg.rank > 0, right = c->cfg.rank < c->cfg.world - 1;
    const int64_t cap_g = c->msg_cap_g;
    k_s2_reset_ghost<<<(unsigned)((2 * cap_g + 255) / 256), 256, 0, c->stream>>>(left ? s2_gh_part(c, c->inbox[par]) : nullptr,
                                                                              right ? s2_gh_part(c, c->inbox[2 + par]) : nullptr, cap_g,
                                                                              c->dev.numcells, c->start, c->end);
    CUS(c, cudaGetLastError());
    c->launches++;
    c->gh_par_last = -1;
    return FSG_OK;
}

static cudaEvent_t s2_event(fsg_ctx *c)
{
    cudaEvent_t e = nullptr;
    if (cudaEventCreate(&e) != cudaSuccess) return nullptr;
    cudaEventRecord(e, c->stream);
    c->ev_ghost.push_back(e);
    return e;
}

// first half: the face layers of the sorted state -> the neighbours' ghost messages
int fsg_slab2_ghost_send(fsg_ctx *c)
{
    if (!c->inbox[0]) { c->err = "fsg_step: the sorted-ghost pipeline needs its messages (fsg_slab_alloc_messages)"; return FSG_E_STATE; }
    const long long seq = ++c->seq_ghost;
    const int par = (int)(seq & 1);
    const bool left = c->cfg.rank > 0, right = c->cfg.rank < c->cfg.world - 1;
    if ((left && !c->peer_inbox[par]) || (right && !c->peer_inbox[2 + par])) {
        c->err = "fsg_step: the neighbours' inboxes are not mapped (fsg_slab_open_peer)";
        return FSG_E_STATE;
    }
    const bool prof = c->profiling && c->ev_ghost.size() < 2 * 4096;
    c->ghost_prof = prof;
    if (prof) s2_event(c);
    if (int rc = fsg_slab_ensure_counts(c, 1)) return rc;
    // remote stores: enough blocks to keep the link busy, few enough to start at once
    const unsigned gb = (unsigned)(c->sm_count > 0 ? c->sm_count : 1);
    k_s2_ghost_send<<<dim3(gb, 2), 256, 0, c->stream>>>(c->A, c->keysA, c->counters + 16, c->counters + 3,
                                                        left ? s2_gh_part(c, c->peer_inbox[par]) : nullptr,
                                                        right ? s2_gh_part(c, c->peer_inbox[2 + par]) : nullptr, c->msg_cap_g, seq, c->counters + 20,
                                                        c->counters + 9, fsg_slab_diag(c));
    CUS(c, cudaGetLastError());
    c->launches++;
    return FSG_OK;
}

// second half: device-side wait for the neighbours' ghosts, then into the zones / tables / home-bin list
int fsg_slab2_ghost_recv(fsg_ctx *c, int nxt)
{
    const long long seq = c->seq_ghost;
    const int par = (int)(seq & 1);
    const bool left = c->cfg.rank > 0, right = c->cfg.rank < c->cfg.world - 1;
    const int64_t cap_g = c->msg_cap_g;
    const size_t tail = c->mig_bytes + c->gh_bytes - 64;
    CUS(c, fsg_launch_slab_wait(c, left ? (const long long *)((char *)c->inbox[par] + tail) : nullptr,
                                right ? (const long long *)((char *)c->inbox[2 + par] + tail) : nullptr, seq, c->stream));
    k_s2_install<<<(unsigned)((2 * cap_g + 255) / 256), 256, 0, c->stream>>>(c->dev, left ? s2_gh_part(c, c->inbox[par]) : nullptr,
                                                                          right ? s2_gh_part(c, c->inbox[2 + par]) : nullptr, cap_g, c->n_own, c->A,
                                                                          c->start, c->end, c->binlist[nxt], c->counters + nxt, c->counters + 9,
                                                                          c->counters + 6, fsg_slab_diag(c));
    CUS(c, cudaGetLastError());
    c->launches += 2;
    c->gh_par_last = par;
    if (c->ghost_prof) s2_event(c);
    return FSG_OK;
}

// mean device milliseconds per step of the ghost exchange (send + wait for the neighbours + install) since the last call
extern "C" int fsg_slab_get_ghost_ms(fsg_ctx *c, double *ms, int64_t *steps)
{
    if (!c || !ms) return FSG_E_INVALID;
    CUS(c, cudaSetDevice(c->device));
    CUS(c, cudaStreamSynchronize(c->stream));
    double tot = 0;
    int64_t k = 0;
    for (size_t g = 0; g + 2 <= c->ev_ghost.size(); g += 2) {
        float t;
        if (cudaEventElapsedTime(&t, c->ev_ghost[g], c->ev_ghost[g + 1]) == cudaSuccess) { tot += t; k++; }
    }
    for (cudaEvent_t e : c->ev_ghost) cudaEventDestroy(e);
    c->ev_ghost.clear();
    *ms = k ? tot / (double)k : 0.0;
    if (steps) *steps = k;
    return FSG_OK;
}
