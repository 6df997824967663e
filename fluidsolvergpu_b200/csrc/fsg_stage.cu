// fsg_stage.cu — stage API: one call per reference kernel on caller-owned device buffers in the
// reference's own layout (340-byte Particle AoS).  See include/fsg.h (2) and INTEGRATION.md.
#include "fsg_internal.cuh"

extern "C" int fsg_stage_sort(fsg_ctx *c, int32_t *d_cells, void *d_particles, int64_t n)
{
    if (!c) return FSG_E_INVALID;
    c->err = "fsg_stage_sort: not implemented yet";
    return FSG_E_UNSUPPORTED;
}
extern "C" int fsg_stage_findneighbours(fsg_ctx *c, const int32_t *d_cells, int32_t *d_start, int32_t *d_end, int64_t n)
{
    if (!c) return FSG_E_INVALID;
    c->err = "fsg_stage_findneighbours: not implemented yet";
    return FSG_E_UNSUPPORTED;
}
extern "C" int fsg_stage_mykernel(fsg_ctx *c, void *d_particles, const int32_t *d_cells, const int32_t *d_start,
                                  const int32_t *d_end, int64_t n)
{
    if (!c) return FSG_E_INVALID;
    c->err = "fsg_stage_mykernel: not implemented yet";
    return FSG_E_UNSUPPORTED;
}
extern "C" int fsg_stage_mykernel2(fsg_ctx *c, void *d_particles, int32_t *d_cells, int32_t *d_start, int32_t *d_end,
                                   int64_t n, float *spts, float *a3, float *b3)
{
    if (!c) return FSG_E_INVALID;
    c->err = "fsg_stage_mykernel2: not implemented yet";
    return FSG_E_UNSUPPORTED;
}
