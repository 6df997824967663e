// fsg_scene.cu — synthetic "plume" scene for the throughput configs: a column of
// fluid particles about the z axis of the bin domain.  The same generator runs on the host
// (fsg_scene_plume_host, used by tests and the end-to-end path) and on the device
// (fsg_scene_plume, used by the resident benchmark); both give identical bits.
//
//   domain      [origin, origin + G*cellsize)^3
//   lattice     spacing `spacing`, first point at origin + spacing/2, columns (ix, iy) whose centre is
//               within R = G*cellsize/8 of the axis, kz over the lower 3/4 of the domain height
//   particle id column_rank * nz + kz   (columns ranked in ix-major order)  == Particle::index
//   jitter      U(-jitter, jitter) per coordinate from splitmix64(seed + 3*id + axis)
//   velocity    (0, 0, 0.5*exp(-(r/R)^2)),  acc = (0,0,GRAVITY), dens = RHO_0, press = 0, newdens = RHO_0
#include "fsg_internal.cuh"

#include <math.h>
#include <vector>

__host__ __device__ static inline uint64_t splitmix64(uint64_t x)
{
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}
__host__ __device__ static inline double u01(uint64_t r) { return (double)(r >> 11) * (1.0 / 9007199254740992.0); }

struct PlumeGeom {
    double x0, spacing, jitter, R, cx;
    int nxy, nz;
    uint64_t seed;
};

static PlumeGeom plume_geom(int G, float origin, double cellsize, double spacing, double jitter, uint64_t seed)
{
    PlumeGeom g;
    double L = G * cellsize;
    g.x0 = (double)origin + 0.5 * spacing;
    g.spacing = spacing;
    g.jitter = jitter;
    g.R = L / 8.0;
    g.cx = (double)origin + 0.5 * L;
    g.nxy = (int)floor(L / spacing);
    g.nz = (int)floor(0.75 * L / spacing);
    if (g.nz < 1) g.nz = 1;
    g.seed = seed;
    return g;
}

// included columns in ix-major order, packed as ix * nxy + iy
static void plume_columns(const PlumeGeom &g, std::vector<int> &cols)
{
    cols.clear();
    for (int ix = 0; ix < g.nxy; ix++) {
        double x = g.x0 + ix * g.spacing - g.cx;
        if (fabs(x) > g.R) continue;
        for (int iy = 0; iy < g.nxy; iy++) {
            double y = g.x0 + iy * g.spacing - g.cx;
            if (x * x + y * y <= g.R * g.R) cols.push_back(ix * g.nxy + iy);
        }
    }
}

__host__ __device__ static inline void plume_particle(const PlumeGeom &g, int col, int kz, int64_t id, float *pos, float *vel)
{
    int ix = col / g.nxy, iy = col % g.nxy;
    double x = g.x0 + ix * g.spacing, y = g.x0 + iy * g.spacing, z = g.x0 + kz * g.spacing;
    double jx = (2.0 * u01(splitmix64(g.seed + 3ull * (uint64_t)id + 0)) - 1.0) * g.jitter;
    double jy = (2.0 * u01(splitmix64(g.seed + 3ull * (uint64_t)id + 1)) - 1.0) * g.jitter;
    double jz = (2.0 * u01(splitmix64(g.seed + 3ull * (uint64_t)id + 2)) - 1.0) * g.jitter;
    double rx = x - g.cx, ry = y - g.cx;
    double r2 = (rx * rx + ry * ry) / (g.R * g.R);
    pos[0] = (float)(x + jx);
    pos[1] = (float)(y + jy);
    pos[2] = (float)(z + jz);
    vel[0] = 0.f;
    vel[1] = 0.f;
    // exp(-r2) by a fixed series on [0,1] so that host and device agree bit for bit
    double t = -r2, e = 1.0, term = 1.0;
    for (int k = 1; k <= 20; k++) { term *= t / k; e += term; }
    vel[2] = (float)(0.5 * e);
}

__global__ void k_plume(PlumeGeom g, const int *__restrict__ cols, int64_t n, float gravity, FsgState st, float4 *carry)
{
    int64_t id = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= n) return;
    int c = (int)(id / g.nz), kz = (int)(id % g.nz);
    float p[3], v[3];
    plume_particle(g, cols[c], kz, id, p, v);
    st.posd[id] = make_float4(p[0], p[1], p[2], 9550.f);
    st.velp[id] = make_float4(v[0], v[1], v[2], 0.f);
    st.accf[id] = make_float4(0.f, 0.f, gravity, __int_as_float(0));
    st.dpi[id] = make_float4(0.f, 0.f, 0.f, __int_as_float((int)id));
    carry[id] = make_float4(9550.f, 0.f, 0.f, 0.f);
}

extern "C" int fsg_scene_plume_host(const fsg_config *cfg, double spacing, double jitter, uint64_t seed, float *pos,
                                    float *vel, int64_t capacity, int64_t *n_out)
{
    if (!cfg || !n_out || spacing <= 0) return FSG_E_INVALID;
    PlumeGeom g = plume_geom(cfg->grid, cfg->origin, cfg->cellsize, spacing, jitter, seed);
    std::vector<int> cols;
    plume_columns(g, cols);
    int64_t n = (int64_t)cols.size() * g.nz;
    *n_out = n;
    if (!pos) return FSG_OK;
    if (!vel || capacity < n) return FSG_E_INVALID;
    for (int64_t id = 0; id < n; id++) plume_particle(g, cols[id / g.nz], (int)(id % g.nz), id, pos + 3 * id, vel + 3 * id);
    return FSG_OK;
}

// device variant; returns the particle count or a negative error through *count_out
int fsg_scene_plume_device(fsg_ctx *c, double spacing, double jitter, uint64_t seed, int64_t *n_out)
{
    PlumeGeom g = plume_geom(c->cfg.grid, c->cfg.origin, c->cfg.cellsize, spacing, jitter, seed);
    std::vector<int> cols;
    plume_columns(g, cols);
    int64_t n = (int64_t)cols.size() * g.nz;
    *n_out = n;
    if (n > c->cap) return FSG_E_NOMEM;
    if (n == 0) return FSG_OK;
    int *dcols = nullptr;
    if (cudaMalloc(&dcols, sizeof(int) * cols.size()) != cudaSuccess) return FSG_E_NOMEM;
    cudaMemcpyAsync(dcols, cols.data(), sizeof(int) * cols.size(), cudaMemcpyHostToDevice, c->stream);
    k_plume<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(g, dcols, n, (float)c->cfg.gravity, c->B, c->carryB);
    cudaError_t e = cudaGetLastError();
    cudaStreamSynchronize(c->stream);
    cudaFree(dcols);
    c->launches++;
    return e == cudaSuccess ? FSG_OK : FSG_E_CUDA;
}
