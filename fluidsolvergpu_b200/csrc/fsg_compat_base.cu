// fsg_compat_base.cu — link-compatible entry points of the reference's base kernels (FluidGPU.cuh:43-47, 417-419).
//
// A driver compiled against the reference header launches   mykernel<<<NUMCELLS,64>>>(d_SPptr, v_d, d_start, d_end, n)
// (solver.cu:187).  nvcc turns that into  __cudaPushCallConfiguration(grid, block, shmem, stream)  followed by a plain
// call of the host function  mykernel(Particle*, int*, int*, int*, int)  — normally the stub nvcc generates next to the
// __global__ definition.  This file defines host functions with exactly those C++ signatures (the mangled names
// solver.o needs: _Z14findneighboursPiS_S_i, _Z8mykernelP8ParticlePiS1_S1_i, _Z9mykernel2P8ParticlePiS1_S1_iPfS2_S2_,
// SURVEY.md §8b): each pops the launch configuration the caller pushed — the driver's grid shape is irrelevant to
// how libfsg schedules the work — and runs the corresponding fsg_stage_* call on the SAME stream, on the caller's
// own device buffers (340-byte Particle records, int tables).  So an object file built from the reference's
// solver.cu links against  fsg_compat_base.o + libfsg.so  instead of FluidGPU.o, unchanged.
//
// Errors cannot be returned (the reference's kernels are void): they are printed once to stderr and the call
// becomes a no-op, which the driver's own cudaGetLastError checks will not see — use the C API for anything new.
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>

#include "../../include/fsg.h"

class Particle;      // only ever passed by pointer here; its 340-byte layout is fixed by FSG_AOS_STRIDE

extern "C" cudaError_t CUDARTAPI __cudaPopCallConfiguration(dim3 *gridDim, dim3 *blockDim, size_t *sharedMem, void *stream);

namespace {
fsg_ctx *g_ctx = nullptr;
int64_t g_cap = 0;

// one process-wide context with the reference's compile-time constants (FluidGPU.cuh:1-31), grown on demand
fsg_ctx *compat_ctx(int64_t n, cudaStream_t stream)
{
    if (!g_ctx || n > g_cap) {
        if (g_ctx) { fsg_destroy(g_ctx); g_ctx = nullptr; }
        fsg_config cfg;
        fsg_config_default(&cfg, FSG_MODEL_BASE);
        int dev = 0;
        cudaGetDevice(&dev);
        cfg.device = dev;
        cfg.capacity = n > 8000 ? n : 8000;
        if (fsg_create(&cfg, &g_ctx) != FSG_OK) {
            fprintf(stderr, "libfsg compat: %s\n", fsg_last_error(nullptr));
            g_ctx = nullptr;
            return nullptr;
        }
        g_cap = cfg.capacity;
    }
    fsg_set_stream(g_ctx, (void *)stream);       // the stream of the <<<>>> launch (the drivers use the legacy default stream)
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
        fprintf(stderr, "libfsg compat: %s failed (%d): %s\n", what, rc, g_ctx ? fsg_last_error(g_ctx) : fsg_last_error(nullptr));
        said = true;
    }
}
}  // namespace

// FluidGPU.cuh:417 / FluidGPU.cu:106-117
void findneighbours(int *cell, int *start, int *end, int nspts)
{
    cudaStream_t st = pop_config();
    fsg_ctx *c = compat_ctx(nspts, st);
    if (c) report("findneighbours", fsg_stage_findneighbours(c, cell, start, end, nspts));
}

// FluidGPU.cuh:418 / FluidGPU.cu:119-285
void mykernel(Particle *SPptr, int *cell, int *start, int *end, int nspts)
{
    cudaStream_t st = pop_config();
    fsg_ctx *c = compat_ctx(nspts, st);
    if (c) report("mykernel", fsg_stage_mykernel(c, SPptr, cell, start, end, nspts));
}

// FluidGPU.cuh:419 / FluidGPU.cu:404-432
void mykernel2(Particle *SPptr, int *cells, int *start, int *end, int nspts, float *spts, float *a3, float *b3)
{
    cudaStream_t st = pop_config();
    fsg_ctx *c = compat_ctx(nspts, st);
    if (c) report("mykernel2", fsg_stage_mykernel2(c, SPptr, cells, start, end, nspts, spts, a3, b3));
}

// The smoothing kernels are host-callable in the reference (FluidGPU.cuh:43-47; Particle::set_dens, inline in the
// header, calls kernel(0) on the host).  cutoff = 0.06 (FluidGPU.cuh:30); unsuffixed literals are double as there.
static const double kCutoff = 0.06;
float kernel(float r)               // FluidGPU.cu:11-21
{
    if (r >= 0 && r <= kCutoff) return 1. / 3.14159 / (powf(kCutoff, 3)) * (1 - 3. / 2. * powf((r / kCutoff), 2) + 3. / 4. * powf((r / kCutoff), 3));
    else if (r > kCutoff && r < (2 * kCutoff)) return 1. / 3.14159 / (powf(kCutoff, 3)) * 1 / 4. * powf(2 - (r / kCutoff), 3);
    return 0;
}
float kernel_test(float r)          // FluidGPU.cu:23-33
{
    if (r >= 0 && r <= kCutoff) return 1. / 3.14159 / (powf(kCutoff, 4)) * (1 - 3. * powf((r / kCutoff), 1) + 9. / 4. * powf((r / kCutoff), 2));
    else if (r > kCutoff && r < (2 * kCutoff)) return -1. / 3.14159 / (powf(kCutoff, 4)) * 1 / 2. * powf(2 - (r / kCutoff), 2);
    return 0;
}
float kernel_derivative(float r)    // FluidGPU.cu:35-43
{
    if (r < kCutoff) return -45.0 / 3.14159 / powf(kCutoff, 6) * powf((kCutoff - r), 2);
    return 0;
}
