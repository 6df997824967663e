// fsg_stage.cu — stage API: one call per reference launch, on caller-owned DEVICE buffers in the
// reference's own layout (340-byte Particle AoS, int key / start / end arrays).  See include/fsg.h (2)
// and INTEGRATION.md §2.  The context supplies constants and scratch memory only.
//
//   fsg_stage_sort            thrust::sort_by_key(t_v, t_v + n, t_a)                       solver.cu:181
//   fsg_stage_findneighbours  findneighbours<<<NUMCELLS,1024>>>      FluidGPU.cu:106-117,  solver.cu:182
//   fsg_stage_mykernel        mykernel<<<NUMCELLS,64>>>              FluidGPU.cu:119-285,  solver.cu:187
//   fsg_stage_mykernel2       mykernel2<<<NUMCELLS,1024>>>           FluidGPU.cu:404-432,  solver.cu:198
#include "fsg_device.cuh"

#include <stdio.h>

#define CUG(ctx, call)                                                                                  \
    do {                                                                                                \
        cudaError_t e_ = (call);                                                                        \
        if (e_ != cudaSuccess) {                                                                        \
            char b_[512];                                                                               \
            snprintf(b_, sizeof b_, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
            (ctx)->err = b_;                                                                            \
            return e_ == cudaErrorMemoryAllocation ? FSG_E_NOMEM : FSG_E_CUDA;                          \
        }                                                                                               \
    } while (0)

namespace aosb {   // FluidGPU.cuh:112-162 (pinned by tests/golden/kat_base.json)
enum { POS = 0, VEL = 12, ACC = 24, INDEX = 36, CELL = 40, DENS = 48, PRESS = 52, DELP_Z = 56, DELP_Y = 60, DELP_X = 64,
       NEWDENS = 84, NDELP_Z = 92, NDELP_Y = 96, NDELP_X = 100, STRESS_RATE = 180, STRESS_TENSOR = 252, BOUNDARY = 336, FLAG = 338 };
}
__device__ __forceinline__ float ldf(const unsigned char *r, int off) { return *reinterpret_cast<const float *>(r + off); }
__device__ __forceinline__ void stf(unsigned char *r, int off, float v) { *reinterpret_cast<float *>(r + off) = v; }

static int stage_scratch(fsg_ctx *c, size_t bytes)
{
    if (c->stage_bytes >= bytes) return FSG_OK;
    if (c->stage) { CUG(c, cudaStreamSynchronize(c->stream)); cudaFree(c->stage); c->stage = nullptr; c->stage_bytes = 0; }
    CUG(c, cudaMalloc(&c->stage, bytes));
    c->stage_bytes = bytes;
    return FSG_OK;
}

// dst[k] = src[perm[k]] for 340-byte records, one 4-byte word per thread (coalesced on both sides)
__global__ void __launch_bounds__(256)
k_gather_records(const unsigned *__restrict__ src, unsigned *__restrict__ dst, const int *__restrict__ perm, int64_t n)
{
    const int W = FSG_AOS_STRIDE / 4;
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n * W) return;
    int64_t k = t / W;
    int w = (int)(t % W);
    dst[t] = src[(int64_t)perm[k] * W + w];
}

extern "C" int fsg_stage_sort(fsg_ctx *c, int32_t *d_cells, void *d_particles, int64_t n)
{
    if (!c || n < 0 || (n > 0 && (!d_cells || !d_particles))) return FSG_E_INVALID;
    if (n > c->cap) { c->err = "fsg_stage_sort: n exceeds the context capacity"; return FSG_E_INVALID; }
    if (n == 0) return FSG_OK;
    CUG(c, cudaSetDevice(c->device));
    const size_t rec = (size_t)n * FSG_AOS_STRIDE, tb = fsg_sort_int_temp_bytes(n);
    int rc = stage_scratch(c, rec + tb + 512);
    if (rc != FSG_OK) return rc;
    unsigned char *tmp_rec = (unsigned char *)c->stage;
    void *tmp_sort = tmp_rec + ((rec + 255) & ~(size_t)255);
    c->keys_prev_valid = false;       // keysA / perm are used as scratch here
    // stable LSD radix sort of (key, slot): the permutation thrust::sort_by_key applies to the records
    CUG(c, fsg_sort_pairs_int(tmp_sort, tb, d_cells, c->keysA, c->iota, c->perm, n, c->stream));
    CUG(c, cudaMemcpyAsync(d_cells, c->keysA, sizeof(int) * (size_t)n, cudaMemcpyDeviceToDevice, c->stream));
    const int64_t words = n * (FSG_AOS_STRIDE / 4);
    k_gather_records<<<(unsigned)((words + 255) / 256), 256, 0, c->stream>>>((const unsigned *)d_particles, (unsigned *)tmp_rec, c->perm, n);
    CUG(c, cudaGetLastError());
    CUG(c, cudaMemcpyAsync(d_particles, tmp_rec, rec, cudaMemcpyDeviceToDevice, c->stream));
    c->launches++;
    return FSG_OK;
}

// findneighbours, FluidGPU.cu:106-117 (without its reads of cell[-1] and cell[n])
__global__ void k_stage_findneighbours(const int *__restrict__ cell, int *start, int *end, int64_t n, int numcells)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int key = cell[i];
    if (key < 0 || key >= numcells) return;       // the reference writes out of bounds here
    if (i == 0 || cell[i - 1] != key) start[key] = (int)i;
    if (i == n - 1 || cell[i + 1] != key) end[key] = (int)i;
}

extern "C" int fsg_stage_findneighbours(fsg_ctx *c, const int32_t *d_cells, int32_t *d_start, int32_t *d_end, int64_t n)
{
    if (!c || n < 0 || (n > 0 && (!d_cells || !d_start || !d_end))) return FSG_E_INVALID;
    if (n == 0) return FSG_OK;
    CUG(c, cudaSetDevice(c->device));
    k_stage_findneighbours<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(d_cells, d_start, d_end, n, c->dev.numcells);
    CUG(c, cudaGetLastError());
    c->launches++;
    return FSG_OK;
}

// read state of the records -> SoA streams the pair kernels read, + the list of occupied bins
__global__ void __launch_bounds__(256)
k_stage_unpack(const unsigned char *__restrict__ aos, const int *__restrict__ cell, int64_t n, int numcells, float4 *posd,
               float4 *velp, int *binlist, int *nocc)
{
    using namespace aosb;
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    bool head = false;
    if (i < n) {
        const unsigned char *r = aos + i * FSG_AOS_STRIDE;
        float dens = ldf(r, DENS);
        posd[i] = make_float4(ldf(r, POS), ldf(r, POS + 4), ldf(r, POS + 8), r[BOUNDARY] ? -dens : dens);
        velp[i] = make_float4(ldf(r, VEL), ldf(r, VEL + 4), ldf(r, VEL + 8), ldf(r, PRESS));
        int key = cell[i];
        head = key >= 0 && key < numcells && (i == 0 || cell[i - 1] != key);
    }
    __shared__ int s_cnt[8], s_base;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned m = __ballot_sync(FULL, head);
    if (lane == 0) s_cnt[warp] = __popc(m);
    __syncthreads();
    if (threadIdx.x == 0) {
        int tot = 0;
#pragma unroll
        for (int w = 0; w < 8; w++) { int cc = s_cnt[w]; s_cnt[w] = tot; tot += cc; }
        s_base = tot ? atomicAdd(nocc, tot) : 0;
    }
    __syncthreads();
    if (head) binlist[s_base + s_cnt[warp] + __popc(m & ((1u << lane) - 1))] = cell[i];
}

// the atomicAdd targets of mykernel (FluidGPU.cu:276-279): sums are ADDED to the record's accumulators
__global__ void k_stage_add_sums(unsigned char *__restrict__ aos, const int *__restrict__ cell, const float4 *__restrict__ sums,
                                 int64_t n, int numcells)
{
    using namespace aosb;
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int key = cell[i];
    if (key < 0 || key >= numcells) return;
    unsigned char *r = aos + i * FSG_AOS_STRIDE;
    float4 s = sums[i];
    stf(r, NEWDENS, ldf(r, NEWDENS) + s.x);
    stf(r, NDELP_X, ldf(r, NDELP_X) + s.y);
    stf(r, NDELP_Y, ldf(r, NDELP_Y) + s.z);
    stf(r, NDELP_Z, ldf(r, NDELP_Z) + s.w);
}

extern "C" int fsg_stage_mykernel(fsg_ctx *c, void *d_particles, const int32_t *d_cells, const int32_t *d_start,
                                  const int32_t *d_end, int64_t n)
{
    if (!c || n < 0 || (n > 0 && (!d_particles || !d_cells || !d_start || !d_end))) return FSG_E_INVALID;
    if (n > c->cap) { c->err = "fsg_stage_mykernel: n exceeds the context capacity"; return FSG_E_INVALID; }
    if (n == 0) return FSG_OK;
    CUG(c, cudaSetDevice(c->device));
    int *binlist = c->binlist[0], *nocc = c->counters + 7, *work = c->counters + 2;
    CUG(c, cudaMemsetAsync(nocc, 0, sizeof(int), c->stream));
    CUG(c, cudaMemsetAsync(work, 0, sizeof(int), c->stream));
    k_stage_unpack<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>((const unsigned char *)d_particles, d_cells, n, c->dev.numcells,
                                                                     c->A.posd, c->A.velp, binlist, nocc);
    CUG(c, cudaGetLastError());
    PairArgs a;
    a.d = c->dev;
    a.n = (int)n;
    a.keysA = d_cells;
    a.start = d_start;
    a.end = d_end;
    a.binlist = binlist;
    a.nocc = nocc;
    a.work = work;
    a.A = c->A;
    a.B = c->B;
    a.keysB = c->keysB;
    a.carry = nullptr;
    a.stats = c->dstats;
    a.sums = c->sums;
    if (c->dev.cap <= 0 && c->dev.bin_cap <= 0) {
        bool hasb = true;                         // unknown scene: evaluate the boundary factors
        CUG(c, fsg_launch_pair_v2(a, c->sums, false, hasb, c->sm_count, 0, c->stream));
    } else {
        CUG(c, fsg_launch_pair_fast(a, false, c->sm_count, c->stream));
    }
    k_stage_add_sums<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>((unsigned char *)d_particles, d_cells, c->sums, n, c->dev.numcells);
    CUG(c, cudaGetLastError());
    c->launches += 3;
    return FSG_OK;
}

// mykernel2 (FluidGPU.cu:404-432) on the records: viz export of the pre-update state, Particle::update,
// new bin id into the record and the key array, accumulators zeroed, bin tables reset for index < NUMCELLS
__global__ void __launch_bounds__(256)
k_stage_mykernel2(FsgDev d, unsigned char *__restrict__ aos, int *__restrict__ cells, int *start, int *end, int64_t n, float *spts,
                  float *a3, float *b3)
{
    using namespace aosb;
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        unsigned char *r = aos + i * FSG_AOS_STRIDE;
        const bool bnd = r[BOUNDARY] != 0;
        const int cellnumber = *reinterpret_cast<const int *>(r + CELL);
        float4 pd = make_float4(ldf(r, POS), ldf(r, POS + 4), ldf(r, POS + 8), ldf(r, DENS));
        if (spts) { spts[3 * i] = pd.x; spts[3 * i + 1] = pd.y; spts[3 * i + 2] = pd.z; }     // :410-414
        if (a3) a3[i] = pd.w;
        if (b3) b3[i] = (float)cellnumber;
        if (bnd) pd.w = -pd.w;
        float4 vp = make_float4(ldf(r, VEL), ldf(r, VEL + 4), ldf(r, VEL + 8), ldf(r, PRESS));
        float4 af = make_float4(ldf(r, ACC), ldf(r, ACC + 4), ldf(r, ACC + 8), 0.f);
        float4 dpi = make_float4(0.f, 0.f, 0.f, 0.f);
        int key;
        particle_update(d, pd, vp, af, dpi, ldf(r, NEWDENS), ldf(r, NDELP_X), ldf(r, NDELP_Y), ldf(r, NDELP_Z), key);
        stf(r, POS, pd.x); stf(r, POS + 4, pd.y); stf(r, POS + 8, pd.z);
        stf(r, VEL, vp.x); stf(r, VEL + 4, vp.y); stf(r, VEL + 8, vp.z);
        stf(r, ACC, af.x); stf(r, ACC + 4, af.y); stf(r, ACC + 8, af.z);
        stf(r, DENS, fabsf(pd.w));
        stf(r, PRESS, vp.w);
        stf(r, DELP_X, dpi.x); stf(r, DELP_Y, dpi.y); stf(r, DELP_Z, dpi.z);
        // stress_tensor = DT * stress_rate (FluidGPU.cuh:278-282)
        for (int q = 0; q < 9; q++) stf(r, STRESS_TENSOR + 4 * q, (float)(d.dt * (double)ldf(r, STRESS_RATE + 4 * q)));
        *reinterpret_cast<int *>(r + CELL) = key;          // :419
        cells[i] = key;                                    // :420
        stf(r, NEWDENS, 0.f); stf(r, NDELP_X, 0.f); stf(r, NDELP_Y, 0.f); stf(r, NDELP_Z, 0.f);   // :422-425
        r[FLAG] = 0;                                       // update() sets flag = false, FluidGPU.cuh:303
    }
    if (i < d.numcells) { start[i] = -1; end[i] = -1; }    // :427-430
}

extern "C" int fsg_stage_mykernel2(fsg_ctx *c, void *d_particles, int32_t *d_cells, int32_t *d_start, int32_t *d_end,
                                   int64_t n, float *spts, float *a3, float *b3)
{
    if (!c || n < 0 || !d_start || !d_end || (n > 0 && (!d_particles || !d_cells))) return FSG_E_INVALID;
    CUG(c, cudaSetDevice(c->device));
    const int64_t threads = n > c->dev.numcells ? n : c->dev.numcells;
    k_stage_mykernel2<<<(unsigned)((threads + 255) / 256), 256, 0, c->stream>>>(c->dev, (unsigned char *)d_particles, d_cells, d_start,
                                                                              d_end, n, spts, a3, b3);
    CUG(c, cudaGetLastError());
    c->launches++;
    return FSG_OK;
}
