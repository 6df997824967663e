// fsg_nsort.cu — the key half of thrust::sort_by_key (solver.cu:181, solver-unidyn.cu:331) for a key array that is ALMOST
// sorted already, hand-written, with NO host synchronisation and no library call.
//
// After a step the slots are in the order of the previous step's sorted keys (`prev`), and only the particles that changed
// bin — a fraction of a percent per step at the reference's time step; in a slab context also the ghosts that were dropped
// or appended — are out of place.  The stable sort by bin id is the order of the composites (bin id, slot).  Slots whose key
// did not change ("stayers") are a sorted subsequence; the others ("movers") are few:
//
//   k_ns_count    movers per tile of 2048 slots                                   (8 B read per slot)
//   k_ns_scan     exclusive scan of the tile counts, one block -> tile offsets, M = number of movers (device side)
//   k_ns_compact  movers (key, slot) compacted IN SLOT ORDER + per 32-slot group: mover bit mask and mover prefix
//   k_rs_*        stable LSD radix sort of the M movers by key, 8 bits a pass (histogram / column scan / scatter); M is read
//                 on the device: fixed grids stride over ceil(M / 2048) tiles, so nothing is launched "per mover count"
//   k_ns_place    every element computes its own final position (a merge by ranking, no merge-path partition):
//                   stayer at slot k :  k - movers_before_slot(k) + #{sorted movers < (key, k)}     (block-level search
//                                       of the tile's key range in the mover list, then a search in shared memory)
//                   mover  q (key,s) :  q + #{stayers < (key, s)}    (two binary searches in `prev` + the group prefix)
//                 and writes (key, slot) there: the sorted key array and the gather permutation k_reorder reads.
//
// Exact for ANY input as long as `prev` is sorted (it is the previous sort's output): when every slot is a mover the radix
// sort does all the work — slow (its scan is one block) but correct, so there is no fallback path and nothing to check on the
// host.  k_reorder verifies the order of the result for free (it reads neighbouring keys anyway) and raises a device flag.
// Traffic at 68 M particles: 8 + 8 + 16 B per slot ≈ 2.2 GB ≈ 0.4 ms; the mover-side kernels work on ~1.6 MB.
#include "fsg_device.cuh"

#define NS_TILE 2048          // slots per tile (256 threads x 8)
#define NS_THREADS 256
#define NS_SM_MOVERS 1024     // movers of a tile's key range held in shared memory by k_ns_place

struct NsBufs {
    int *tile_cnt;            // [ntiles + 1] movers per tile -> exclusive offsets; [ntiles] = M
    int *tile_lo;             // [ntiles + 1] sorted movers below the tile's first composite
    int *grp_off;             // [n / 32 + 1] movers before the 32-slot group
    unsigned *grp_mask;       // [n / 32 + 1] which slots of the group are movers
    int *mk[2], *ms[2];       // mover keys / slots, ping-pong for the radix passes
    int *hist;                // [mover tiles][256]
    int *base;                // [256] digit totals, [256] digit bases of the current pass
    int *M;                   // device-side mover count (== tile_cnt[ntiles])
};

__device__ __forceinline__ bool ns_is_mover(const int *__restrict__ knew, const int *__restrict__ prev, int64_t k) { return knew[k] != prev[k]; }

__global__ void __launch_bounds__(NS_THREADS)
k_ns_count(const int *__restrict__ knew, const int *__restrict__ prev, int64_t n, int *__restrict__ tile_cnt)
{
    const int64_t base = (int64_t)blockIdx.x * NS_TILE;
    int c = 0;
#pragma unroll
    for (int r = 0; r < NS_TILE / NS_THREADS; r++) {
        const int64_t k = base + r * NS_THREADS + threadIdx.x;
        if (k < n) c += ns_is_mover(knew, prev, k);
    }
    c = __reduce_add_sync(FULL, c);
    __shared__ int s[NS_THREADS / 32];
    if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        int t = 0;
#pragma unroll
        for (int w = 0; w < NS_THREADS / 32; w++) t += s[w];
        tile_cnt[blockIdx.x] = t;
    }
}

// exclusive scan of cnt[0..m) in place, cnt[m] = total; one block of 1024 threads, 8 consecutive entries per thread and round
__global__ void __launch_bounds__(1024)
k_ns_scan(int *__restrict__ cnt, int64_t m, int *__restrict__ total_out)
{
    __shared__ int s_w[32];
    __shared__ int s_run;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) s_run = 0;
    __syncthreads();
    for (int64_t b = 0; b < m; b += 8192) {
        const int64_t i0 = b + (int64_t)threadIdx.x * 8;
        int v[8], sum = 0;
#pragma unroll
        for (int r = 0; r < 8; r++) { v[r] = i0 + r < m ? cnt[i0 + r] : 0; sum += v[r]; }
        int incl = sum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(FULL, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) s_w[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            int w = s_w[lane], wi = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(FULL, wi, o);
                if (lane >= o) wi += t;
            }
            s_w[lane] = wi - w;
        }
        __syncthreads();
        const int run = s_run;
        int ex = run + s_w[warp] + incl - sum;
#pragma unroll
        for (int r = 0; r < 8; r++) { if (i0 + r < m) cnt[i0 + r] = ex; ex += v[r]; }
        __syncthreads();
        if (threadIdx.x == 1023) s_run = run + s_w[31] + incl;
        __syncthreads();
    }
    if (threadIdx.x == 0) { cnt[m] = s_run; if (total_out) *total_out = s_run; }
}

// movers compacted in slot order; group masks / prefixes.  Thread t of a tile handles the 8 CONSECUTIVE slots 8t .. 8t+7 of
// it, so a warp covers 8 whole 32-slot groups: lanes 4g .. 4g+3 hold group g.
__global__ void __launch_bounds__(NS_THREADS)
k_ns_compact(const int *__restrict__ knew, const int *__restrict__ prev, int64_t n, const int *__restrict__ tile_off, int *__restrict__ mk,
             int *__restrict__ ms, int *__restrict__ grp_off, unsigned *__restrict__ grp_mask)
{
    const int64_t base = (int64_t)blockIdx.x * NS_TILE + (int64_t)threadIdx.x * 8;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int key[8];
    unsigned bits = 0;
    if (base + 8 <= n) {                                 // (the arrays are 256-byte aligned, base is a multiple of 8: two 16-byte loads each)
        const int4 a0 = *reinterpret_cast<const int4 *>(knew + base), a1 = *reinterpret_cast<const int4 *>(knew + base + 4);
        const int4 p0 = *reinterpret_cast<const int4 *>(prev + base), p1 = *reinterpret_cast<const int4 *>(prev + base + 4);
        key[0] = a0.x; key[1] = a0.y; key[2] = a0.z; key[3] = a0.w; key[4] = a1.x; key[5] = a1.y; key[6] = a1.z; key[7] = a1.w;
        bits = (a0.x != p0.x) | ((a0.y != p0.y) << 1) | ((a0.z != p0.z) << 2) | ((a0.w != p0.w) << 3) | ((a1.x != p1.x) << 4) |
               ((a1.y != p1.y) << 5) | ((a1.z != p1.z) << 6) | ((a1.w != p1.w) << 7);
    } else {
#pragma unroll
        for (int r = 0; r < 8; r++) {
            const int64_t k = base + r;
            key[r] = 0;
            if (k < n) {
                key[r] = knew[k];
                if (key[r] != prev[k]) bits |= 1u << r;
            }
        }
    }
    const int mine = __popc(bits);
    int incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(FULL, incl, o);
        if (lane >= o) incl += t;
    }
    __shared__ int s_w[NS_THREADS / 32];
    if (lane == 31) s_w[warp] = incl;
    __syncthreads();
    int wbase = tile_off[blockIdx.x];
#pragma unroll
    for (int w = 0; w < NS_THREADS / 32; w++) if (w < warp) wbase += s_w[w];
    int at = wbase + incl - mine;                       // movers before this thread's first slot
    // the 32-slot group of this thread: its 4 threads' bytes -> one mask; prefix = `at` of the group's first thread
    const unsigned b0 = __shfl_sync(FULL, bits, lane & ~3), b1 = __shfl_sync(FULL, bits, (lane & ~3) + 1),
                   b2 = __shfl_sync(FULL, bits, (lane & ~3) + 2), b3 = __shfl_sync(FULL, bits, (lane & ~3) + 3);
    if ((lane & 3) == 0 && base < n) {
        grp_off[base >> 5] = at;
        grp_mask[base >> 5] = b0 | (b1 << 8) | (b2 << 16) | (b3 << 24);
    }
#pragma unroll
    for (int r = 0; r < 8; r++) {
        if (bits & (1u << r)) {
            mk[at] = key[r];
            ms[at] = (int)(base + r);
            at++;
        }
    }
}

// ---- stable LSD radix sort of the movers: 8 bits a pass, M read on the device ----
// hist[d * tcap + t] = movers of digit d in mover tile t (digit-major, so that the scan of one digit over the tiles is contiguous)
__global__ void __launch_bounds__(NS_THREADS)
k_rs_hist(const int *__restrict__ mk, const int *__restrict__ Mp, int shift, int *__restrict__ hist, int64_t tcap)
{
    __shared__ int s_h[256];
    const int M = *Mp;
    const int ntiles = (M + NS_TILE - 1) / NS_TILE;
    for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
        s_h[threadIdx.x] = 0;
        __syncthreads();
        const int b = t * NS_TILE;
#pragma unroll
        for (int r = 0; r < NS_TILE / NS_THREADS; r++) {
            const int i = b + r * NS_THREADS + threadIdx.x;
            if (i < M) atomicAdd(&s_h[((unsigned)mk[i] >> shift) & 255u], 1);
        }
        __syncthreads();
        hist[(int64_t)threadIdx.x * tcap + t] = s_h[threadIdx.x];
        __syncthreads();
    }
}

// block d: exclusive scan of hist[d][0 .. ntiles) in place (movers of digit d in the tiles before t), total[d] = their number
__global__ void __launch_bounds__(256)
k_rs_colscan(int *__restrict__ hist, const int *__restrict__ Mp, int64_t tcap, int *__restrict__ total)
{
    __shared__ int s_w[8];
    __shared__ int s_run;
    const int M = *Mp;
    const int ntiles = (M + NS_TILE - 1) / NS_TILE;
    int *col = hist + (int64_t)blockIdx.x * tcap;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) s_run = 0;
    __syncthreads();
    for (int b = 0; b < ntiles; b += 1024) {
        const int i0 = b + threadIdx.x * 4;
        int v[4], sum = 0;
#pragma unroll
        for (int r = 0; r < 4; r++) { v[r] = i0 + r < ntiles ? col[i0 + r] : 0; sum += v[r]; }
        int incl = sum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(FULL, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) s_w[warp] = incl;
        __syncthreads();
        int wb = 0;
#pragma unroll
        for (int w = 0; w < 8; w++) if (w < warp) wb += s_w[w];
        const int run = s_run;
        int ex = run + wb + incl - sum;
#pragma unroll
        for (int r = 0; r < 4; r++) { if (i0 + r < ntiles) col[i0 + r] = ex; ex += v[r]; }
        __syncthreads();
        if (threadIdx.x == 255) s_run = run + wb + incl;
        __syncthreads();
    }
    if (threadIdx.x == 0) total[blockIdx.x] = s_run;
}

// base[d] = movers of smaller digits (exclusive scan of the 256 totals)
__global__ void __launch_bounds__(256)
k_rs_base(const int *__restrict__ total, int *__restrict__ base)
{
    __shared__ int s_w[8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int v = total[threadIdx.x];
    int incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(FULL, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) s_w[warp] = incl;
    __syncthreads();
    int wb = 0;
#pragma unroll
    for (int w = 0; w < 8; w++) if (w < warp) wb += s_w[w];
    base[threadIdx.x] = wb + incl - v;
}

// warp w of a tile takes the 256 consecutive movers [256 w, 256 w + 256) in 8 rounds of 32: tile order = (warp, round, lane)
__global__ void __launch_bounds__(NS_THREADS)
k_rs_scatter(const int *__restrict__ mk, const int *__restrict__ ms, int *__restrict__ ok, int *__restrict__ os, const int *__restrict__ Mp,
             int shift, const int *__restrict__ hist, int64_t tcap, const int *__restrict__ base)
{
    __shared__ int s_cnt[NS_THREADS / 32][256];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned lt = (1u << lane) - 1u;
    const int M = *Mp;
    const int ntiles = (M + NS_TILE - 1) / NS_TILE;
    for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
#pragma unroll
        for (int w = 0; w < NS_THREADS / 32; w++) s_cnt[w][threadIdx.x] = 0;
        __syncthreads();
        int key[8], val[8], lrank[8];
        const int b = t * NS_TILE + warp * 256;
#pragma unroll
        for (int r = 0; r < 8; r++) {
            const int i = b + r * 32 + lane;
            const bool in = i < M;
            key[r] = in ? mk[i] : 0;
            val[r] = in ? ms[i] : 0;
            const int d = in ? (int)(((unsigned)key[r] >> shift) & 255u) : 256 + lane;      // padding lanes match nobody
            const unsigned peers = __match_any_sync(FULL, d);
            const int leader = __ffs(peers) - 1;
            int old = 0;
            if (in && lane == leader) { old = s_cnt[warp][d]; s_cnt[warp][d] = old + __popc(peers); }
            old = __shfl_sync(FULL, old, leader);
            lrank[r] = old + __popc(peers & lt);
            __syncwarp();
        }
        __syncthreads();
        {   // exclusive prefix over the warps, per digit (thread = digit)
            int run = 0;
#pragma unroll
            for (int w = 0; w < NS_THREADS / 32; w++) { const int v = s_cnt[w][threadIdx.x]; s_cnt[w][threadIdx.x] = run; run += v; }
        }
        __syncthreads();
#pragma unroll
        for (int r = 0; r < 8; r++) {
            const int i = b + r * 32 + lane;
            if (i < M) {
                const int d = (int)(((unsigned)key[r] >> shift) & 255u);
                const int at = base[d] + hist[(int64_t)d * tcap + t] + s_cnt[warp][d] + lrank[r];
                ok[at] = key[r];
                os[at] = val[r];
            }
        }
        __syncthreads();
    }
}

// (key a, slot sa) < (key b, slot sb)
__device__ __forceinline__ bool ns_less(int ka, int sa, int kb, int sb) { return ka < kb || (ka == kb && sa < sb); }

// number of sorted movers in [lo, hi) that are < (key, slot)
__device__ __forceinline__ int ns_lower_bound(const int *__restrict__ mk, const int *__restrict__ ms, int lo, int hi, int key, int slot)
{
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (ns_less(mk[mid], ms[mid], key, slot)) lo = mid + 1; else hi = mid;
    }
    return lo;
}

// tile_lo[t] = sorted movers below the first composite of tile t, (prev[2048 t], 2048 t); tile_lo[ntiles] = M.  One thread per
// tile: the searches of all tiles run side by side instead of at the head of every placing block.
__global__ void __launch_bounds__(NS_THREADS)
k_ns_tile_bounds(const int *__restrict__ prev, int64_t ntiles, const int *__restrict__ mk, const int *__restrict__ ms, const int *__restrict__ Mp,
                 int *__restrict__ tile_lo)
{
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t > ntiles) return;
    const int M = *Mp;
    tile_lo[t] = t == ntiles ? M : ns_lower_bound(mk, ms, 0, M, prev[t * NS_TILE], (int)(t * NS_TILE));
}

__global__ void __launch_bounds__(NS_THREADS)
k_ns_place_stayers(const int *__restrict__ knew, const int *__restrict__ prev, int64_t n, const int *__restrict__ mk, const int *__restrict__ ms,
                   const int *__restrict__ tile_lo, const int *__restrict__ grp_off, const unsigned *__restrict__ grp_mask, int *__restrict__ keys_out,
                   int *__restrict__ perm)
{
    __shared__ int s_mk[NS_SM_MOVERS], s_ms[NS_SM_MOVERS];
    const int64_t base = (int64_t)blockIdx.x * NS_TILE;
    // the tile's stayers have composites in [(prev[base], base), first composite of the next tile): movers below that range are
    // counted by tile_lo, movers above it are after all of them
    const int s_lo = tile_lo[blockIdx.x], s_hi = tile_lo[blockIdx.x + 1];
    const int lo = s_lo, hi = s_hi, cnt = hi - lo;
    const bool in_smem = cnt <= NS_SM_MOVERS;
    if (in_smem)
        for (int i = threadIdx.x; i < cnt; i += NS_THREADS) { s_mk[i] = mk[lo + i]; s_ms[i] = ms[lo + i]; }
    __syncthreads();
    // all of this thread's loads first (8 slots x new key, previous key, group prefix, group mask): the placing loop below has
    // searches and early exits that would otherwise keep one slot's loads from overlapping the next one's
    int kn[NS_TILE / NS_THREADS], kp[NS_TILE / NS_THREADS], go[NS_TILE / NS_THREADS];
    unsigned gm[NS_TILE / NS_THREADS];
#pragma unroll
    for (int r = 0; r < NS_TILE / NS_THREADS; r++) {
        const int64_t k = base + r * NS_THREADS + threadIdx.x;
        const bool in = k < n;
        kn[r] = in ? knew[k] : 0;
        kp[r] = in ? prev[k] : 1;                                       // (out of range: looks like a mover, is skipped)
        go[r] = in ? grp_off[k >> 5] : 0;
        gm[r] = in ? grp_mask[k >> 5] : 0u;
    }
#pragma unroll
    for (int r = 0; r < NS_TILE / NS_THREADS; r++) {
        const int64_t k = base + r * NS_THREADS + threadIdx.x;
        const int key = kn[r];
        if (key != kp[r]) continue;                                     // movers place themselves
        const int before = go[r] + __popc(gm[r] & ((1u << (k & 31)) - 1u));
        int rank;
        if (in_smem) {
            int a = 0, b = cnt;
            while (a < b) {
                const int mid = (a + b) >> 1;
                if (ns_less(s_mk[mid], s_ms[mid], key, (int)k)) a = mid + 1; else b = mid;
            }
            rank = lo + a;
        } else
            rank = ns_lower_bound(mk, ms, lo, hi, key, (int)k);
        const int64_t pos = k - before + rank;
        keys_out[pos] = key;
        perm[pos] = (int)k;
    }
}

__global__ void __launch_bounds__(NS_THREADS)
k_ns_place_movers(const int *__restrict__ prev, int64_t n, const int *__restrict__ mk, const int *__restrict__ ms, const int *__restrict__ Mp,
                  const int *__restrict__ grp_off, const unsigned *__restrict__ grp_mask, int *__restrict__ keys_out, int *__restrict__ perm)
{
    const int M = *Mp;
    for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < M; q += gridDim.x * blockDim.x) {
        const int key = mk[q], slot = ms[q];
        // slots j with (prev[j], j) < (key, slot): all slots whose prev key is smaller + those of the same prev key below `slot`
        int64_t a = 0, b = n;
        while (a < b) { const int64_t mid = (a + b) >> 1; if (prev[mid] < key) a = mid + 1; else b = mid; }
        int64_t e = a, f = n;
        while (e < f) { const int64_t mid = (e + f) >> 1; if (prev[mid] <= key) e = mid + 1; else f = mid; }
        int64_t p = slot;                                  // clamp(slot, a, e)
        if (p < a) p = a;
        if (p > e) p = e;
        // stayers among the slots [0, p)
        int64_t movers_before;
        if (p >= n) movers_before = M;
        else movers_before = grp_off[p >> 5] + __popc(grp_mask[p >> 5] & ((1u << (p & 31)) - 1u));
        const int64_t pos = q + (p - movers_before);
        keys_out[pos] = key;
        perm[pos] = slot;
    }
}

size_t fsg_nsort_bytes(int64_t n)
{
    const int64_t ntiles = (n + NS_TILE - 1) / NS_TILE;
    auto al = [](size_t b) { return (b + 255) & ~(size_t)255; };
    return 2 * al(sizeof(int) * (ntiles + 2)) + 2 * al(sizeof(int) * (n / 32 + 2)) + 4 * al(sizeof(int) * (size_t)n) + al(sizeof(int) * (size_t)ntiles * 256) +
           al(sizeof(int) * 512) + 256;
}

static NsBufs ns_layout(void *ws, int64_t n)
{
    const int64_t ntiles = (n + NS_TILE - 1) / NS_TILE;
    auto al = [](size_t b) { return (b + 255) & ~(size_t)255; };
    char *p = (char *)ws;
    NsBufs B;
    B.tile_cnt = (int *)p; p += al(sizeof(int) * (ntiles + 2));
    B.tile_lo = (int *)p; p += al(sizeof(int) * (ntiles + 2));
    B.grp_off = (int *)p; p += al(sizeof(int) * (n / 32 + 2));
    B.grp_mask = (unsigned *)p; p += al(sizeof(int) * (n / 32 + 2));
    for (int k = 0; k < 2; k++) { B.mk[k] = (int *)p; p += al(sizeof(int) * (size_t)n); B.ms[k] = (int *)p; p += al(sizeof(int) * (size_t)n); }
    B.hist = (int *)p; p += al(sizeof(int) * (size_t)ntiles * 256);
    B.base = (int *)p; p += al(sizeof(int) * 512);
    B.M = B.tile_cnt + ntiles;
    return B;
}

// keys_out / perm_out must not alias keys_new / keys_prev.  Asynchronous on `s`; nothing is read back.
cudaError_t fsg_nsort(void *ws, const int *keys_new, const int *keys_prev, int *keys_out, int *perm_out, int64_t n, int bits, int sm_count,
                      cudaStream_t s, int *launches)
{
    if (n <= 0) return cudaSuccess;
    if (n > 0x7fffffffll - NS_TILE) return cudaErrorInvalidValue;
    const NsBufs B = ns_layout(ws, n);
    const int64_t ntiles = (n + NS_TILE - 1) / NS_TILE;
    k_ns_count<<<(unsigned)ntiles, NS_THREADS, 0, s>>>(keys_new, keys_prev, n, B.tile_cnt);
    k_ns_scan<<<1, 1024, 0, s>>>(B.tile_cnt, ntiles, nullptr);
    k_ns_compact<<<(unsigned)ntiles, NS_THREADS, 0, s>>>(keys_new, keys_prev, n, B.tile_cnt, B.mk[0], B.ms[0], B.grp_off, B.grp_mask);
    int64_t g = ntiles < (int64_t)sm_count * 8 ? ntiles : (int64_t)sm_count * 8;
    if (g < 1) g = 1;
    int cur = 0, nl = 3;
    for (int shift = 0; shift < bits; shift += 8) {
        k_rs_hist<<<(unsigned)g, NS_THREADS, 0, s>>>(B.mk[cur], B.M, shift, B.hist, ntiles);
        k_rs_colscan<<<256, 256, 0, s>>>(B.hist, B.M, ntiles, B.base);
        k_rs_base<<<1, 256, 0, s>>>(B.base, B.base + 256);
        k_rs_scatter<<<(unsigned)g, NS_THREADS, 0, s>>>(B.mk[cur], B.ms[cur], B.mk[cur ^ 1], B.ms[cur ^ 1], B.M, shift, B.hist, ntiles, B.base + 256);
        cur ^= 1;
        nl += 4;
    }
    k_ns_tile_bounds<<<(unsigned)((ntiles + 1 + NS_THREADS - 1) / NS_THREADS), NS_THREADS, 0, s>>>(keys_prev, ntiles, B.mk[cur], B.ms[cur], B.M, B.tile_lo);
    k_ns_place_stayers<<<(unsigned)ntiles, NS_THREADS, 0, s>>>(keys_new, keys_prev, n, B.mk[cur], B.ms[cur], B.tile_lo, B.grp_off, B.grp_mask, keys_out,
                                                             perm_out);
    k_ns_place_movers<<<(unsigned)g, NS_THREADS, 0, s>>>(keys_prev, n, B.mk[cur], B.ms[cur], B.M, B.grp_off, B.grp_mask, keys_out, perm_out);
    if (launches) *launches += nl + 3;
    return cudaGetLastError();
}
