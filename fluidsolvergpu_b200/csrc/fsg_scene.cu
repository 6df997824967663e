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

__global__ void k_plume(PlumeGeom g, const int *__restrict__ cols, int64_t n, int64_t id0, float gravity, FsgState st, float4 *carry)
{
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    int c = (int)(t / g.nz), kz = (int)(t % g.nz);
    int64_t id = id0 + t;                         // global particle id == Particle::index
    float p[3], v[3];
    plume_particle(g, cols[c], kz, id, p, v);
    st.posd[t] = make_float4(p[0], p[1], p[2], 9550.f);
    st.velp[t] = make_float4(v[0], v[1], v[2], 0.f);
    st.accf[t] = make_float4(0.f, 0.f, gravity, __int_as_float(0));
    st.dpi[t] = make_float4(0.f, 0.f, 0.f, __int_as_float((int)id));
    carry[t] = make_float4(9550.f, 0.f, 0.f, 0.f);
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

// particles per bin layer ix (lattice positions, jitter ignored): what the host uses to cut the domain
// into slabs of equal particle count
extern "C" int fsg_scene_plume_hist(const fsg_config *cfg, double spacing, int64_t *hist)
{
    if (!cfg || !hist || spacing <= 0) return FSG_E_INVALID;
    PlumeGeom g = plume_geom(cfg->grid, cfg->origin, cfg->cellsize, spacing, 0.0, 0);
    std::vector<int> cols;
    plume_columns(g, cols);
    for (int i = 0; i < cfg->grid; i++) hist[i] = 0;
    for (int c : cols) {
        double x = g.x0 + (c / g.nxy) * g.spacing;
        int ix = (int)(((float)x - cfg->origin) / cfg->cellsize);
        if (ix < 0) ix = 0;
        if (ix >= cfg->grid) ix = cfg->grid - 1;
        hist[ix] += g.nz;
    }
    return FSG_OK;
}

// device variant; returns the particle count or a negative error through *count_out.
// Slab contexts generate only the lattice columns that can fall into their slab (ids stay global);
// strays are dropped by the slab filter of the key kernel.
int fsg_scene_plume_device(fsg_ctx *c, double spacing, double jitter, uint64_t seed, int64_t *n_out)
{
    PlumeGeom g = plume_geom(c->cfg.grid, c->cfg.origin, c->cfg.cellsize, spacing, jitter, seed);
    std::vector<int> cols;
    plume_columns(g, cols);
    size_t c_lo = 0, c_hi = cols.size();
    if (c->cfg.world > 1) {
        const double xlo = (double)c->cfg.origin + c->cfg.slab_x0 * c->cfg.cellsize - jitter - 1e-3 * spacing;
        const double xhi = (double)c->cfg.origin + c->cfg.slab_x1 * c->cfg.cellsize + jitter + 1e-3 * spacing;
        while (c_lo < cols.size() && g.x0 + (cols[c_lo] / g.nxy) * g.spacing < xlo) c_lo++;
        c_hi = c_lo;
        while (c_hi < cols.size() && g.x0 + (cols[c_hi] / g.nxy) * g.spacing <= xhi) c_hi++;
    }
    const size_t ncols = c_hi - c_lo;
    int64_t n = (int64_t)ncols * g.nz;
    *n_out = n;
    if (n > c->cap) return FSG_E_NOMEM;
    if (n == 0) return FSG_OK;
    int *dcols = nullptr;
    if (cudaMalloc(&dcols, sizeof(int) * ncols) != cudaSuccess) return FSG_E_NOMEM;
    cudaMemcpyAsync(dcols, cols.data() + c_lo, sizeof(int) * ncols, cudaMemcpyHostToDevice, c->stream);
    k_plume<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(g, dcols, n, (int64_t)c_lo * g.nz, (float)c->cfg.gravity, c->B, c->carryB);
    cudaError_t e = cudaGetLastError();
    cudaStreamSynchronize(c->stream);
    cudaFree(dcols);
    c->launches++;
    return e == cudaSuccess ? FSG_OK : FSG_E_CUDA;
}
