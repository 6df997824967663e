// fsg_compat_unidyn.cu — link-compatible entry points of the reference's unidyn kernels (FluidGPU-unidyn.cuh:52-56, 537-544).
//
// Same mechanism as fsg_compat_base.cu: a driver compiled against FluidGPU-unidyn.cuh launches
//   mykernel<<<NUMCELLS,1024>>>(d_SPptr, d_particleindex, v_d, d_start, d_end, d_split, dsz, x, dev, buffer, d_numsplit)
// (solver-unidyn.cu:363); nvcc turns that into __cudaPushCallConfiguration(...) + a plain call of the host function
// mykernel(Particle*, int*, int*, int*, int*, int*, int, int, int, int, int*).  This file defines host functions with
// exactly the C++ signatures of FluidGPU-unidyn.cuh:537-544: each pops the launch configuration and runs the matching
// fsg_stage_unidyn_* call on the SAME stream, on the caller's own device buffers.  An object built from
// solver-unidyn.cu links against  fsg_compat_unidyn.o + libfsg.so  instead of FluidGPU-unidyn.o.
//
// Scope: the single-device loop the shipped driver runs (deviceCount forced to 1, buffer 0, solver-unidyn.cu:192-195) and
// the scenes the context API accepts (non-boundary particles pure fluid, mass 1).  particleindex is the identity there
// (solver-unidyn.cu:233-238) and is not read.  find_idx / mem_shift (the 2-device hand-off, never launched by the shipped
// driver) are provided so that the driver links: find_idx follows FluidGPU-unidyn.cu:499-529, mem_shift does the shift
// of :531-542 with a real ordering between its two phases (the reference uses __syncthreads() as if it were a grid barrier).
#include <cuda_runtime.h>
#include <limits.h>
#include <math.h>
#include <stdio.h>

#include "../../include/fsg.h"

class Particle;      // only ever passed by pointer here; its 340-byte layout is fixed by FSG_AOS_STRIDE

extern "C" cudaError_t CUDARTAPI __cudaPopCallConfiguration(dim3 *gridDim, dim3 *blockDim, size_t *sharedMem, void *stream);

namespace {
fsg_ctx *g_ctx = nullptr;
int64_t g_cap = 0;

// one process-wide context with the reference's compile-time constants (FluidGPU-unidyn.cuh:1-36), grown on demand
fsg_ctx *compat_ctx(int64_t n, cudaStream_t stream)
{
    if (!g_ctx || n > g_cap) {
        if (g_ctx) { fsg_destroy(g_ctx); g_ctx = nullptr; }
        fsg_config cfg;
        fsg_config_default(&cfg, FSG_MODEL_UNIDYN);
        int dev = 0;
        cudaGetDevice(&dev);
        cfg.device = dev;
        cfg.capacity = n > 14040 ? n : 14040;
        if (fsg_create(&cfg, &g_ctx) != FSG_OK) {
            fprintf(stderr, "libfsg compat (unidyn): %s\n", fsg_last_error(nullptr));
            g_ctx = nullptr;
            return nullptr;
        }
        g_cap = cfg.capacity;
    }
    fsg_set_stream(g_ctx, (void *)stream);
    return g_ctx;
}

cudaStream_t pop_config()
{
    dim3 g, b;
    size_t sh = 0;
    cudaStream_t st = nullptr;
    __cudaPopCallConfiguration(&g, &b, &sh, &st);
    return st;
}

void report(const char *what, int rc)
{
    static bool said = false;
    if (rc != FSG_OK && !said) {
        fprintf(stderr, "libfsg compat (unidyn): %s failed (%d): %s\n", what, rc, g_ctx ? fsg_last_error(g_ctx) : fsg_last_error(nullptr));
        said = true;
    }
}

// find_idx, FluidGPU-unidyn.cu:499-529 (NUMCELLS = 4913; the read of SPptr[npts] past the end is treated as "beyond every bin")
__global__ void k_find_idx(const int *cells, int dev, int npts, int buffer, int *xleft, int *xright, int *sleft, int *sright, int numcells)
{
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    const int lb = dev * numcells / 2, hb = lb + numcells / 2;
    if (idx >= npts) return;
    const int a = cells[idx], b = idx + 1 < npts ? cells[idx + 1] : INT_MAX;
    if (dev == 0 && b >= hb && a < hb) { xleft[0] = 0; xright[0] = idx; sleft[0] = 0; }
    if (dev == 0 && b >= hb - buffer && a < hb - buffer) sright[0] = idx + 1;
    if (dev == 1 && a < lb && b >= lb) { xleft[0] = idx + 1; xright[0] = npts - 1; sright[0] = npts; }
    if (dev == 1 && a < lb + buffer && b >= lb + buffer) sleft[0] = idx + 1;
}
// mem_shift, FluidGPU-unidyn.cu:531-542, as two launches: records [l, r] -> buff, then buff -> records [l - shifts, r - shifts]
__global__ void k_copy_words(const unsigned *src, unsigned *dst, long long words)
{
    long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < words) dst[t] = src[t];
}
}  // namespace

// FluidGPU-unidyn.cuh:537 / FluidGPU-unidyn.cu:106-122
void findneighbours(int *cell, int *start, int *start_copy, int *end, int nspts, int x)
{
    cudaStream_t st = pop_config();
    fsg_ctx *c = compat_ctx(nspts, st);
    if (c) report("findneighbours", fsg_stage_unidyn_findneighbours(c, cell, start, start_copy, end, nspts, x));
}

// FluidGPU-unidyn.cuh:538 / FluidGPU-unidyn.cu:124-449
void mykernel(Particle *SPptr, int *particleindex, int *cell, int *start, int *end, int *split, int nspts, int x, int dev, int buffer,
              int *numsplit)
{
    (void)particleindex; (void)x; (void)dev; (void)buffer;
    cudaStream_t st = pop_config();
    fsg_ctx *c = compat_ctx(nspts, st);
    if (c) report("mykernel", fsg_stage_unidyn_mykernel(c, SPptr, cell, start, end, split, numsplit, nspts));
}

// FluidGPU-unidyn.cuh:539 / FluidGPU-unidyn.cu:569-870
void mykernel3(Particle *SPptr, int *particleindex, int *cell, int *start, int *end, int *split, int nspts, int x, int dev, int buffer,
               int *numsplit)
{
    (void)particleindex; (void)split; (void)x; (void)dev; (void)buffer; (void)numsplit;
    cudaStream_t st = pop_config();
    fsg_ctx *c = compat_ctx(nspts, st);
    if (c) report("mykernel3", fsg_stage_unidyn_mykernel3(c, SPptr, cell, start, end, nspts));
}

// FluidGPU-unidyn.cuh:540 / FluidGPU-unidyn.cu:451-497
void mykernel2(Particle *SPptr, int *particleindex, int *cell, int *start_copy, int *start, int *end, int *split, int *numsplit, int nspts,
               int x, int dev, int buffer, int t, float *spts, float *a3, float *b3)
{
    (void)particleindex; (void)dev; (void)buffer;
    cudaStream_t st = pop_config();
    fsg_ctx *c = compat_ctx(nspts, st);
    if (c) report("mykernel2", fsg_stage_unidyn_mykernel2(c, SPptr, cell, start_copy, start, end, split, numsplit, nspts, x, t, spts, a3, b3));
}

// FluidGPU-unidyn.cuh:541 / FluidGPU-unidyn.cu:499-529
void find_idx(int *SPptr, int dev, int npts, int buffer, int *xleft, int *xright, int *sleft, int *sright)
{
    cudaStream_t st = pop_config();
    if (npts > 0) k_find_idx<<<(npts + 255) / 256, 256, 0, st>>>(SPptr, dev, npts, buffer, xleft, xright, sleft, sright, 4913);
}

// FluidGPU-unidyn.cuh:542 / FluidGPU-unidyn.cu:531-542
void mem_shift(Particle *SPptr, Particle *buff, int *cells, int *ibuff, int dev, int shifts, int indexleft, int indexright)
{
    (void)cells; (void)ibuff; (void)dev;
    cudaStream_t st = pop_config();
    if (shifts == 0 || indexright < indexleft) return;
    const long long W = FSG_AOS_STRIDE / 4, words = (long long)(indexright - indexleft + 1) * W;
    const unsigned *src = (const unsigned *)SPptr + (long long)indexleft * W;
    unsigned *dst = (unsigned *)SPptr + (long long)(indexleft - shifts) * W;
    k_copy_words<<<(unsigned)((words + 255) / 256), 256, 0, st>>>(src, (unsigned *)buff, words);
    k_copy_words<<<(unsigned)((words + 255) / 256), 256, 0, st>>>((const unsigned *)buff, dst, words);
}

// FluidGPU-unidyn.cuh:543 / FluidGPU-unidyn.cu:544-551
void cell_calc(Particle *SPptr, int *particleindex, int *cells, int size, int dev)
{
    (void)particleindex; (void)dev;
    cudaStream_t st = pop_config();
    fsg_ctx *c = compat_ctx(size, st);
    if (c) report("cell_calc", fsg_stage_unidyn_cell_calc(c, SPptr, cells, size));
}

// FluidGPU-unidyn.cuh:544 / FluidGPU-unidyn.cu:554-562
void count_after_merge(int *cells, int *particleindex, int size, int *newsize)
{
    (void)particleindex;
    cudaStream_t st = pop_config();
    fsg_ctx *c = compat_ctx(size, st);
    if (c) report("count_after_merge", fsg_stage_unidyn_count_after_merge(c, cells, size, newsize));
}

// The smoothing kernels are host-callable in the reference (FluidGPU-unidyn.cuh:52-56; Particle::set_dens, inline in the
// header, calls kernel(0) on the host).  cutoff = 0.06 (FluidGPU-unidyn.cuh:35); unsuffixed literals are double as there.
static const double kCutoff = 0.06;
float kernel(float r)               // FluidGPU-unidyn.cu:11-21
{
    if (r >= 0 && r <= kCutoff) return 1. / 3.14159 / (powf(kCutoff, 3)) * (1 - 3. / 2. * powf((r / kCutoff), 2) + 3. / 4. * powf((r / kCutoff), 3));
    else if (r > kCutoff && r < (2 * kCutoff)) return 1. / 3.14159 / (powf(kCutoff, 3)) * 1 / 4. * powf(2 - (r / kCutoff), 3);
    return 0;
}
float kernel_test(float r)          // FluidGPU-unidyn.cu:23-33
{
    if (r >= 0 && r <= kCutoff) return 1. / 3.14159 / (powf(kCutoff, 4)) * (1 - 3. * powf((r / kCutoff), 1) + 9. / 4. * powf((r / kCutoff), 2));
    else if (r > kCutoff && r < (2 * kCutoff)) return -1. / 3.14159 / (powf(kCutoff, 4)) * 1 / 2. * powf(2 - (r / kCutoff), 2);
    return 0;
}
float kernel_derivative(float r)    // FluidGPU-unidyn.cu:35-43
{
    if (r < kCutoff) return -45.0 / 3.14159 / powf(kCutoff, 6) * powf((kCutoff - r), 2);
    return 0;
}
