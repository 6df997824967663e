// fsg_api.cu — the extern "C" boundary of libfsg (include/fsg.h): context lifetime, host<->device
// movement and the step schedule.  No torch types, no exceptions across the boundary, no CPU path.
#include <limits.h>
#include "fsg_device.cuh"
#ifndef FSG_SORT_MERGE_MIN_CAP
#define FSG_SORT_MERGE_MIN_CAP (1 << 20)
#endif
#include <stdlib.h>

#include <math.h>
#include <stdio.h>
#include <string.h>
#include <new>

static std::string g_create_err;

#define CU(ctx, call)                                                                                   \
    do {                                                                                                \
        cudaError_t e_ = (call);                                                                        \
        if (e_ != cudaSuccess) {                                                                        \
            char b_[512];                                                                               \
            snprintf(b_, sizeof b_, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
            (ctx)->err = b_;                                                                            \
            return e_ == cudaErrorMemoryAllocation ? FSG_E_NOMEM : FSG_E_CUDA;                          \
        }                                                                                               \
    } while (0)

int fsg_scene_plume_device(fsg_ctx *c, double spacing, double jitter, uint64_t seed, int64_t *n_out);

// ---- host-derived fp32 thresholds (see FsgDev) ----
static float largest_float_le(double v) { float f = (float)v; while ((double)f > v) f = nextafterf(f, -INFINITY); return f; }
static float largest_float_lt(double v) { float f = (float)v; while ((double)f >= v) f = nextafterf(f, -INFINITY); return f; }

// Which pair kernel the uncapped fp32 configuration runs: the symmetric one (fsg_pair_v3.cu) unless the caller asked for
// the deterministic gather kernel (pair_mode = 1, or FSG_PAIR_SYM=0 in the environment), pair counts are being collected
// (the counting build is the gather kernel) or the exchange overlaps the interior bins (two launches per step).
void fsg_update_pair_mode(fsg_ctx *c)
{
    static int env = -1;
    if (env < 0) {
        const char *e = getenv("FSG_PAIR_SYM");
        env = e ? (atoi(e) != 0) : 1;
    }
    const fsg_config &f = c->cfg;
    c->dev.sym = env && f.model == FSG_MODEL_BASE && f.pair_mode == 0 && f.pair_fp64 == 0 && f.neighbour_cap == 0 && f.bin_cap == 0 &&
                 f.collect_stats == 0 && !c->overlap && f.grid >= 4;
}

void fsg_derive_constants(const fsg_config &cfg, FsgDev &d)
{
    d.sym = 0;
    d.uni_open = cfg.unidyn_open_box != 0;
    d.G = cfg.grid;
    d.G2 = cfg.grid * cfg.grid;
    d.numcells = cfg.grid * cfg.grid * cfg.grid;
    d.x0 = cfg.world > 1 ? cfg.slab_x0 : 0;
    d.x1 = cfg.world > 1 ? cfg.slab_x1 : cfg.grid;
    d.bx0 = d.x0;
    d.bx1 = d.x1;
    d.rl = d.x0 + ((cfg.world > 1 && cfg.rank > 0) ? 2 : 0);
    d.rr = d.x1 - ((cfg.world > 1 && cfg.rank < cfg.world - 1) ? 2 : 0);
    if (d.rr < d.rl) d.rr = d.rl;
    d.dead = d.numcells + 1;
    d.kx0 = d.kx1 = INT_MIN;
    d.cap = cfg.neighbour_cap;
    d.bin_cap = cfg.bin_cap;
    d.origin = cfg.origin;
    d.cellsize = cfg.cellsize;
    d.h = cfg.h;
    d.dt = cfg.dt;
    d.gravity = cfg.gravity;
    d.sound = cfg.sound;
    d.alpha_fluid = cfg.alpha_fluid;
    d.alpha_boundary = cfg.alpha_boundary;
    const double h = cfg.h;
    // gate of FluidGPU.cu:236: (double)ds <= 2*cutoff with ds = sqrtf(d2)  <=>  d2 <= d2_max
    float ds_max = largest_float_le(2 * h);
    float d2 = ds_max * ds_max;
    while (sqrtf(d2) > ds_max) d2 = nextafterf(d2, 0.f);
    while (sqrtf(nextafterf(d2, INFINITY)) <= ds_max) d2 = nextafterf(d2, INFINITY);
    d.d2_max = d2;
    d.h_le = largest_float_le(h);
    d2 = d.h_le * d.h_le;
    while (sqrtf(d2) > d.h_le) d2 = nextafterf(d2, 0.f);
    while (sqrtf(nextafterf(d2, INFINITY)) <= d.h_le) d2 = nextafterf(d2, INFINITY);
    d.d2_h = d2;
    d.h_lt = largest_float_lt(h);
    d.twoh_lt = largest_float_lt(2 * h);
    d.hf = (float)h;
    d.inv_h = (float)(1.0 / h);
    d.w_c = (float)(1. / 3.14159 / (double)powf((float)h, 3.f));
    d.dw_c = (float)(-45.0 / 3.14159 / (double)powf((float)h, 6.f));
    d.w0 = (float)(1. / 3.14159 / (double)powf((float)h, 3.f) * (1 - 3. / 2. * 0.0 + 3. / 4. * 0.0));   // kernel(0)
    d.eps = (float)(0.01 * (double)powf((float)h, 2.f));
    d.visc_c = (float)(cfg.alpha_fluid * cfg.sound);
    d.visc_q = (float)(50 * 1.0 / cfg.sound);
}

extern "C" int fsg_version(void) { return FSG_VERSION; }

extern "C" int fsg_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

extern "C" int fsg_config_default(fsg_config *cfg, int model)
{
    if (!cfg) return FSG_E_INVALID;
    memset(cfg, 0, sizeof *cfg);
    cfg->model = model;
    cfg->origin = -1.0f;
    cfg->h = 0.06;
    cfg->gravity = -9.8;
    cfg->sound = 1450.0;
    cfg->world = 1;
    if (model == FSG_MODEL_BASE) {            // FluidGPU.cuh:1-31, solver.cu:17-19,187
        cfg->grid = 40;
        cfg->cellsize = 0.05;
        cfg->dt = 0.0005;
        cfg->alpha_fluid = -0.01e2;
        cfg->alpha_boundary = 2000e-1;
        cfg->neighbour_cap = 64;
        cfg->bin_cap = 64;
        cfg->capacity = 8000;
    } else if (model == FSG_MODEL_UNIDYN) {   // FluidGPU-unidyn.cuh:1-36, solver-unidyn.cu:21-23,363
        cfg->grid = 17;
        cfg->cellsize = 0.12;
        cfg->dt = 0.0018;
        cfg->alpha_fluid = -0.0155e1;
        cfg->alpha_boundary = 100e-1;     // ALPHA__SAND_BOUNDARY: the boundary factor the unidyn pair term uses (FluidGPU-unidyn.cu:307)
        cfg->neighbour_cap = 1024;
        cfg->bin_cap = 0;
        cfg->capacity = 14040;
        cfg->unidyn_merge_distance = -10.0;   // FluidGPU-unidyn.cu:261
        cfg->unidyn_split_mass_min = 3.0;     // :278
    } else
        return FSG_E_INVALID;
    return FSG_OK;
}

static void free_state(FsgState &s)
{
    cudaFree(s.posd); cudaFree(s.velp); cudaFree(s.accf); cudaFree(s.dpi); cudaFree(s.mix); cudaFree(s.stress);
    s.posd = s.velp = s.accf = s.dpi = s.mix = nullptr;
    s.stress = nullptr;
}

extern "C" int fsg_destroy(fsg_ctx *c)
{
    if (!c) return FSG_E_INVALID;
    cudaSetDevice(c->device);
    fsg_frame_writer_destroy(c);
    if (c->stream) cudaStreamSynchronize(c->stream);
    free_state(c->A);
    free_state(c->B);
    cudaFree(c->carryA); cudaFree(c->carryB); cudaFree(c->sums); cudaFree(c->sums2); cudaFree(c->vizb); cudaFree(c->mixA); cudaFree(c->mixB);
    cudaFree(c->keysA); cudaFree(c->keysB); cudaFree(c->perm); cudaFree(c->iota);
    cudaFree(c->start); cudaFree(c->end);
    cudaFree(c->binlist[0]); cudaFree(c->binlist[1]);
    cudaFree(c->counters); cudaFree(c->dstats); cudaFree(c->sort_tmp);
    cudaFree(c->keysC); cudaFree(c->ns_ws);
    cudaFree(c->adapt_ws); cudaFree(c->adapt_scan);
    cudaFree(c->slab_cnt); cudaFree(c->scan_tmp);
    if (c->comm) cudaStreamSynchronize(c->comm);
    for (int k = 0; k < 4; k++) if (c->peer_inbox[k] && !c->peer_local) cudaIpcCloseMemHandle(c->peer_inbox[k]);
    if (c->ev_boundary) cudaEventDestroy(c->ev_boundary);
    if (c->ev_sent) cudaEventDestroy(c->ev_sent);
    if (c->comm) cudaStreamDestroy(c->comm);
    cudaFree(c->binlistB);
    for (int k = 0; k < 2; k++) cudaFree(c->outbox[k]);
    for (int k = 0; k < 4; k++) cudaFree(c->inbox[k]);
    cudaFree(c->stage);
    if (c->host_flag) cudaFreeHost((void *)c->host_flag);
    for (cudaEvent_t e : c->ev_pool) cudaEventDestroy(e);
    for (cudaEvent_t e : c->ev_used) cudaEventDestroy(e);
    for (cudaEvent_t e : c->ev_ghost) cudaEventDestroy(e);
    if (c->own_stream && c->stream) cudaStreamDestroy(c->stream);
    delete c;
    return FSG_OK;
}

static int alloc_state(fsg_ctx *c, FsgState &s, int64_t cap)
{
    CU(c, cudaMalloc(&s.posd, sizeof(float4) * cap));
    CU(c, cudaMalloc(&s.velp, sizeof(float4) * cap));
    CU(c, cudaMalloc(&s.accf, sizeof(float4) * cap));
    CU(c, cudaMalloc(&s.dpi, sizeof(float4) * cap));
    if (c->cfg.model == FSG_MODEL_UNIDYN) {
        CU(c, cudaMalloc(&s.mix, sizeof(float4) * cap));
        if (c->cfg.world == 1) {        // granular state (stress_tensor, stress_rate): carried by single-device contexts only
            CU(c, cudaMalloc(&s.stress, sizeof(float) * 18 * cap));
            CU(c, cudaMemsetAsync(s.stress, 0, sizeof(float) * 18 * cap, c->stream));
        }
    }
    return FSG_OK;
}

static int create_impl(fsg_ctx *c)
{
    const fsg_config &cfg = c->cfg;
    CU(c, cudaSetDevice(cfg.device));
    cudaDeviceProp prop;
    CU(c, cudaGetDeviceProperties(&prop, cfg.device));
    c->sm_count = prop.multiProcessorCount;
    CU(c, cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    c->own_stream = true;
    const int64_t cap = cfg.capacity > 0 ? cfg.capacity : 1;
    c->cap = cap;
    int rc;
    if ((rc = alloc_state(c, c->A, cap)) != FSG_OK) return rc;
    if ((rc = alloc_state(c, c->B, cap)) != FSG_OK) return rc;
    CU(c, cudaMalloc(&c->carryA, sizeof(float4) * cap));
    CU(c, cudaMalloc(&c->carryB, sizeof(float4) * cap));
    CU(c, cudaMalloc(&c->sums, sizeof(float4) * cap));
    if (cfg.model == FSG_MODEL_UNIDYN) {
        CU(c, cudaMalloc(&c->sums2, sizeof(float4) * cap));
        CU(c, cudaMalloc(&c->vizb, sizeof(float) * cap));
        CU(c, cudaMemsetAsync(c->vizb, 0, sizeof(float) * cap, c->stream));
        if (cfg.world == 1) {
            CU(c, cudaMalloc(&c->mixA, sizeof(float) * 18 * cap));
            CU(c, cudaMalloc(&c->mixB, sizeof(float) * 8 * cap));
        }
    }
    CU(c, cudaMalloc(&c->keysA, sizeof(int) * cap));
    CU(c, cudaMalloc(&c->keysB, sizeof(int) * cap));
    CU(c, cudaMalloc(&c->perm, sizeof(int) * cap));
    CU(c, cudaMalloc(&c->iota, sizeof(int) * cap));
    const int64_t nc = c->dev.numcells;
    CU(c, cudaMalloc(&c->start, sizeof(int) * nc));
    CU(c, cudaMalloc(&c->end, sizeof(int) * nc));
    const int64_t nb = cap < nc ? cap : nc;
    CU(c, cudaMalloc(&c->binlist[0], sizeof(int) * (nb > 0 ? nb : 1)));
    CU(c, cudaMalloc(&c->binlist[1], sizeof(int) * (nb > 0 ? nb : 1)));
    CU(c, cudaMalloc(&c->counters, sizeof(int) * 32));
    CU(c, cudaMalloc(&c->dstats, sizeof(unsigned long long) * 4));
    CU(c, cudaMemsetAsync(c->counters, 0, sizeof(int) * 32, c->stream));
    CU(c, cudaMemsetAsync(c->dstats, 0, sizeof(unsigned long long) * 4, c->stream));
    CU(c, fsg_launch_iota(c->iota, cap, c->stream));
    CU(c, fsg_launch_fill(c->start, -1, nc, c->stream));    // solver.cu:163-169
    CU(c, fsg_launch_fill(c->end, -1, nc, c->stream));
    c->launches += 3;
    c->sort_bits = 1;
    while ((1ll << c->sort_bits) <= nc + 1) c->sort_bits++;  // keys are in [0, numcells + 1] (parked, dead)
    c->sort_tmp_bytes = fsg_sort_temp_bytes(cap, c->sort_bits);
    CU(c, cudaMalloc(&c->sort_tmp, c->sort_tmp_bytes ? c->sort_tmp_bytes : 16));
    CU(c, cudaStreamSynchronize(c->stream));
    return FSG_OK;
}

extern "C" int fsg_create(const fsg_config *cfg, fsg_ctx **out)
{
    if (!cfg || !out) return FSG_E_INVALID;
    *out = nullptr;
    if (cfg->grid < 1 || cfg->grid > 1290 || cfg->cellsize <= 0 || cfg->h <= 0 || cfg->capacity < 0 ||
        cfg->capacity > 2000000000ll || (cfg->model != FSG_MODEL_BASE && cfg->model != FSG_MODEL_UNIDYN)) {
        g_create_err = "fsg_create: invalid configuration";
        return FSG_E_INVALID;
    }
    if (cfg->world < 1 || cfg->rank < 0 || cfg->rank >= cfg->world ||
        (cfg->world > 1 && (cfg->slab_x0 < 0 || cfg->slab_x1 > cfg->grid || cfg->slab_x1 - cfg->slab_x0 < 2))) {
        g_create_err = "fsg_create: invalid slab configuration (rank/world; a slab needs at least 2 bin layers)";
        return FSG_E_INVALID;
    }
    if (cfg->model == FSG_MODEL_BASE && cfg->world > 1 && (cfg->neighbour_cap != 0 || cfg->bin_cap != 0 || cfg->pair_fp64 != 0)) {
        g_create_err = "fsg_create: slab decomposition needs the uncapped configuration (neighbour_cap = bin_cap = 0) "
                       "and a one-layer ghost band, i.e. cellsize >= 2h";
        return FSG_E_UNSUPPORTED;
    }
    if (cfg->unidyn_adapt && (cfg->model != FSG_MODEL_UNIDYN || cfg->world > 1)) {
        g_create_err = "fsg_create: unidyn_adapt (particle merging / splitting) needs the unidyn model on a single-device context";
        return FSG_E_UNSUPPORTED;
    }
    int ndev = fsg_device_count();
    if (ndev <= 0 || cfg->device < 0 || cfg->device >= ndev) {
        g_create_err = "fsg_create: no usable CUDA device (libfsg has no CPU path)";
        return FSG_E_NO_DEVICE;
    }
    fsg_ctx *c = new (std::nothrow) fsg_ctx();
    if (!c) return FSG_E_NOMEM;
    c->cfg = *cfg;
    c->device = cfg->device;
    c->ns_mode = -1;
    c->defer_mode = -1;
    c->gh_par_last = -1;
    fsg_derive_constants(*cfg, c->dev);
    fsg_update_pair_mode(c);
    int rc = create_impl(c);
    if (rc != FSG_OK) {
        g_create_err = c->err;
        fsg_destroy(c);
        return rc;
    }
    *out = c;
    return FSG_OK;
}

extern "C" const char *fsg_last_error(const fsg_ctx *c) { return c ? c->err.c_str() : g_create_err.c_str(); }

extern "C" int fsg_set_stream(fsg_ctx *c, void *s)
{
    if (!c) return FSG_E_INVALID;
    if (!c->own_stream && c->stream == (cudaStream_t)s) return FSG_OK;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    if (c->own_stream) cudaStreamDestroy(c->stream);
    c->stream = (cudaStream_t)s;
    c->own_stream = false;
    return FSG_OK;
}
extern "C" void *fsg_get_stream(fsg_ctx *c) { return c ? (void *)c->stream : nullptr; }

extern "C" int fsg_sync(fsg_ctx *c)
{
    if (!c) return FSG_E_INVALID;
    CU(c, cudaSetDevice(c->device));
    CU(c, cudaStreamSynchronize(c->stream));
    if (c->comm) CU(c, cudaStreamSynchronize(c->comm));
    return FSG_OK;
}

// after new particles have been written to B / carryB
static int after_upload(fsg_ctx *c, int64_t n, const int *slot_state = nullptr)
{
    // the bin tables may hold the entries of a previous run
    if (c->tables_dirty) {
        if (c->cfg.world > 1 && !c->slab2) CU(c, fsg_launch_reset_tables_keys(c->dev, c->keysA, c->start, c->end, c->n_sorted, c->stream));
        else CU(c, fsg_launch_reset_tables(c->binlist[c->cur], c->counters + c->cur, c->keysA, c->start, c->end, c->n, c->stream));
        c->launches++;
        c->tables_dirty = false;
        if (c->slab2) { int rc = fsg_slab2_reset_ghost_tables(c); if (rc != FSG_OK) return rc; }
    }
    if (c->slab2 && n > c->n_own) {
        c->err = "upload: n exceeds the context capacity minus the two ghost zones (capacity - 2 * cap_g)";
        return FSG_E_INVALID;
    }
    c->slab2_mid = false;
    c->n = n;
    c->keys_prev_valid = false;
    c->deferred = false;
    c->carry_pending = false;
    c->sent_ahead = false;
    if (c->host_flag) *c->host_flag = 0;      // a fresh state: a previous exchange time-out no longer applies
    if (c->comm) CU(c, cudaStreamSynchronize(c->comm));
    if (c->cfg.world > 1) {
        // a slab context always works on all `cap` slots (nothing on the host depends on how many are in use):
        // the tail holds the dead key, the device-side count of slots in use starts at n
        CU(c, fsg_launch_fill(c->keysB + n, c->dev.dead, c->cap - n, c->stream));
        if (!c->keep_foreign) {      // (a re-uploaded download keeps the device-side count of the slots in use)
            const int n32 = (int)n;
            CU(c, cudaMemcpyAsync(c->counters + 5, &n32, sizeof(int), cudaMemcpyHostToDevice, c->stream));
        }
        CU(c, cudaMemsetAsync(c->counters + 9, 0, sizeof(int), c->stream));
    }
    CU(c, cudaMemsetAsync(c->counters + 4, 0, sizeof(int), c->stream));
    CU(c, cudaMemsetAsync(c->counters + 6, 0, sizeof(int), c->stream));
    CU(c, fsg_launch_keys(c->dev, c->B.posd, c->keysB, n, c->counters + 4, c->cfg.world > 1 && !c->keep_foreign, slot_state,
                          c->stream));   // solver.cu:119
    c->launches++;
    int flag[5] = {0, 0, 0, 0, 0};
    CU(c, cudaMemcpyAsync(flag, c->counters + 4, sizeof(flag), cudaMemcpyDeviceToHost, c->stream));
    CU(c, cudaStreamSynchronize(c->stream));
    c->has_boundary = flag[0] != 0;
    if (c->cfg.world > 1) c->n = c->slab2 ? c->n_own : c->cap;
    CU(c, cudaMemsetAsync(c->counters + 8, 0, sizeof(int), c->stream));
    c->mixed = false;
    if (c->cfg.model == FSG_MODEL_UNIDYN && flag[4]) {
        // bit 1: a particle with mass != 1 (merged / split particles are outside the built scope); bit 2: a non-boundary particle
        // with solid != 0 — a mixed-phase / granular scene: the two-pass kernels of fsg_unidyn_mixed.cu, single-device contexts
        const bool bad_mass = (flag[4] & 1) && !c->cfg.unidyn_adapt, mixed_scene = (flag[4] & 2) != 0;
        if (bad_mass || (mixed_scene && (c->cfg.world > 1 || c->cfg.unidyn_adapt))) {
            c->n = 0;
            c->err = bad_mass ? "upload: the unidyn path needs mass == 1 for every particle unless fsg_config.unidyn_adapt is set (solver-unidyn.cu:127-184)"
                     : c->cfg.unidyn_adapt ? "upload: particle merging / splitting (unidyn_adapt) runs on pure-fluid scenes only"
                                   : "upload: mixed-phase / granular unidyn scenes (a non-boundary particle with solid != 0) run on single-device "
                                     "contexts only (the slab messages do not carry the granular stress state)";
            return FSG_E_UNSUPPORTED;
        }
        c->mixed = mixed_scene;
    }
    c->adapt_next_index = (int)n;             // children continue the numbering 0 .. n-1 of the uploaded particles
    for (int k = 0; k < 3; k++) c->adapt_counts[k] = c->adapt_totals[k] = 0;
    c->carry_live = true;
    c->steps = 0;
    return FSG_OK;
}

static int ensure_stage(fsg_ctx *c, size_t bytes)
{
    if (c->stage_bytes >= bytes) return FSG_OK;
    if (c->stage) { CU(c, cudaStreamSynchronize(c->stream)); cudaFree(c->stage); c->stage = nullptr; c->stage_bytes = 0; }
    CU(c, cudaMalloc(&c->stage, bytes));
    c->stage_bytes = bytes;
    return FSG_OK;
}

extern "C" int fsg_upload_aos(fsg_ctx *c, const void *particles, int64_t n)
{
    if (!c || (!particles && n > 0) || n < 0) return FSG_E_INVALID;
    if (n > c->cap) { c->err = "fsg_upload_aos: n exceeds the context capacity"; return FSG_E_INVALID; }
    CU(c, cudaSetDevice(c->device));
    const int64_t chunk = 1 << 20;
    int rc = ensure_stage(c, (size_t)(n < chunk ? n : chunk) * FSG_AOS_STRIDE + 16);
    if (rc != FSG_OK) return rc;
    for (int64_t o = 0; o < n; o += chunk) {
        int64_t m = n - o < chunk ? n - o : chunk;
        CU(c, cudaMemcpyAsync(c->stage, (const unsigned char *)particles + o * FSG_AOS_STRIDE, (size_t)m * FSG_AOS_STRIDE,
                              cudaMemcpyHostToDevice, c->stream));
        FsgState st = {c->B.posd + o, c->B.velp + o, c->B.accf + o, c->B.dpi + o, nullptr};
        if (c->cfg.model == FSG_MODEL_UNIDYN) {
            st.mix = c->B.mix + o;
            st.stress = c->B.stress ? c->B.stress + 18 * o : nullptr;
            CU(c, fsg_launch_unpack_aos_unidyn((const unsigned char *)c->stage, m, st, c->carryB + o, c->counters + 8, c->stream));
        } else
            CU(c, fsg_launch_unpack_aos(c->cfg.model, (const unsigned char *)c->stage, m, st, c->carryB + o, c->stream));
        c->launches++;
        CU(c, cudaStreamSynchronize(c->stream));
    }
    return after_upload(c, n);
}

extern "C" int fsg_download_aos(fsg_ctx *c, void *particles, int64_t n)
{
    if (!c || (!particles && n > 0) || n < 0 || n > c->n) return FSG_E_INVALID;
    CU(c, cudaSetDevice(c->device));
    if (int rm = fsg_materialize(c)) return rm;
    const int64_t chunk = 1 << 20;
    int rc = ensure_stage(c, (size_t)(n < chunk ? n : chunk) * FSG_AOS_STRIDE + 16);
    if (rc != FSG_OK) return rc;
    for (int64_t o = 0; o < n; o += chunk) {
        int64_t m = n - o < chunk ? n - o : chunk;
        FsgState st = {c->B.posd + o, c->B.velp + o, c->B.accf + o, c->B.dpi + o, nullptr};
        if (c->cfg.model == FSG_MODEL_UNIDYN) {
            st.mix = c->B.mix + o;
            st.stress = c->B.stress ? c->B.stress + 18 * o : nullptr;
            CU(c, fsg_launch_pack_aos_unidyn((unsigned char *)c->stage, m, st, c->carry_live ? c->carryB + o : nullptr, c->keysB + o, c->dev,
                                             c->stream));
        } else
            CU(c, fsg_launch_pack_aos(c->cfg.model, (unsigned char *)c->stage, m, st, c->carry_live ? c->carryB + o : nullptr,
                                      c->keysB + o, 101325.f, c->stream));
        c->launches++;
        CU(c, cudaMemcpyAsync((unsigned char *)particles + o * FSG_AOS_STRIDE, c->stage, (size_t)m * FSG_AOS_STRIDE,
                              cudaMemcpyDeviceToHost, c->stream));
        CU(c, cudaStreamSynchronize(c->stream));
    }
    return FSG_OK;
}

// ---- SoA host interface: raw arrays are copied into a staging area and (un)packed on the device ----
__global__ void k_pack_soa(int64_t n, const float *pos, const float *vel, const float *acc, const float *dens,
                           const float *press, const float *delp, const float *nd, const float *ndp, const int *index,
                           const unsigned char *bnd, const float *solid, const float *fluid, const float *stress_tensor, const float *stress_rate,
                           const float *mass, float gravity, FsgState st, float4 *carry, int *bad)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    bool b = bnd ? bnd[i] != 0 : false;
    float de = dens ? dens[i] : 9550.f;
    st.posd[i] = make_float4(pos[3 * i], pos[3 * i + 1], pos[3 * i + 2], b ? -de : de);
    st.velp[i] = make_float4(vel ? vel[3 * i] : 0.f, vel ? vel[3 * i + 1] : 0.f, vel ? vel[3 * i + 2] : 0.f, press ? press[i] : 0.f);
    st.accf[i] = make_float4(acc ? acc[3 * i] : 0.f, acc ? acc[3 * i + 1] : 0.f, acc ? acc[3 * i + 2] : (b ? 0.f : gravity),
                             __int_as_float(b ? 1 : 0));
    st.dpi[i] = make_float4(delp ? delp[3 * i] : 0.f, delp ? delp[3 * i + 1] : 0.f, delp ? delp[3 * i + 2] : 0.f,
                            __int_as_float(index ? index[i] : (int)i));
    carry[i] = make_float4(nd ? nd[i] : 9550.f, ndp ? ndp[3 * i] : 0.f, ndp ? ndp[3 * i + 1] : 0.f, ndp ? ndp[3 * i + 2] : 0.f);
    if (st.mix) {      // unidyn: solid / fluid default to 0/1 for fluid, 1/0 for boundary particles (solver-unidyn.cu:129-143)
        float so = solid ? solid[i] : (b ? 1.f : 0.f), fl = fluid ? fluid[i] : (b ? 0.f : 1.f);
        const float ms = mass ? mass[i] : 1.f;
        st.mix[i] = make_float4(so, fl, ms - 1.f, 0.f);                                   // z = mass - 1 (fsg_unidyn_adapt.cu)
        if (ms != 1.f) atomicOr(bad, 1);              // merged / split particles: fsg_config.unidyn_adapt
        if (!b && so != 0.f) atomicOr(bad, 2);        // a mixed-phase / granular scene
    }
    if (st.stress) {
        for (int k = 0; k < 9; k++) {
            st.stress[i * 18 + k] = stress_tensor ? stress_tensor[9 * i + k] : 0.f;
            st.stress[i * 18 + 9 + k] = stress_rate ? stress_rate[9 * i + k] : 0.f;
        }
    }
}

__global__ void k_unpack_soa(int64_t n, FsgState st, const float4 *carry, const int *keys, float *pos, float *vel, float *acc,
                             float *dens, float *press, float *delp, float *nd, float *ndp, int *index, int *cell,
                             unsigned char *bnd, float *solid, float *fluid, float *stress_tensor, float *stress_rate, float *mass)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float4 pd = st.posd[i], vp = st.velp[i], af = st.accf[i], dp = st.dpi[i];
    float4 cy = carry ? carry[i] : make_float4(0.f, 0.f, 0.f, 0.f);
    if (pos) { pos[3 * i] = pd.x; pos[3 * i + 1] = pd.y; pos[3 * i + 2] = pd.z; }
    if (vel) { vel[3 * i] = vp.x; vel[3 * i + 1] = vp.y; vel[3 * i + 2] = vp.z; }
    if (acc) { acc[3 * i] = af.x; acc[3 * i + 1] = af.y; acc[3 * i + 2] = af.z; }
    if (dens) dens[i] = fabsf(pd.w);
    if (press) press[i] = vp.w;
    if (delp) { delp[3 * i] = dp.x; delp[3 * i + 1] = dp.y; delp[3 * i + 2] = dp.z; }
    if (nd) nd[i] = cy.x;
    if (ndp) { ndp[3 * i] = cy.y; ndp[3 * i + 1] = cy.z; ndp[3 * i + 2] = cy.w; }
    if (index) index[i] = __float_as_int(dp.w);
    if (cell) cell[i] = keys[i];
    if (bnd) bnd[i] = pd.w < 0.f ? 1 : 0;
    if (st.mix && (solid || fluid || mass)) {
        float4 mx = st.mix[i];
        if (solid) solid[i] = mx.x;
        if (fluid) fluid[i] = mx.y;
        if (mass) mass[i] = 1.f + mx.z;
    }
    for (int k = 0; k < 9; k++) {
        if (stress_tensor) stress_tensor[9 * i + k] = st.stress ? st.stress[i * 18 + k] : 0.f;
        if (stress_rate) stress_rate[9 * i + k] = st.stress ? st.stress[i * 18 + 9 + k] : 0.f;
    }
}

struct SoaStage {
    float *pos, *vel, *acc, *dens, *press, *delp, *nd, *ndp;
    int *index, *cell;
    unsigned char *bnd;
    float *solid, *fluid, *stress_tensor, *stress_rate, *mass;
};
static size_t soa_stage_layout(char *base, int64_t n, const fsg_soa *h, SoaStage &s)
{
    size_t o = 0;
    auto take = [&](const void *hp, size_t bytes) -> char * {
        if (!hp) return nullptr;
        char *p = base ? base + o : (char *)1;
        o += (bytes + 255) & ~(size_t)255;
        return p;
    };
    s.pos = (float *)take(h->pos, 12 * n);
    s.vel = (float *)take(h->vel, 12 * n);
    s.acc = (float *)take(h->acc, 12 * n);
    s.dens = (float *)take(h->dens, 4 * n);
    s.press = (float *)take(h->press, 4 * n);
    s.delp = (float *)take(h->delpress, 12 * n);
    s.nd = (float *)take(h->newdens, 4 * n);
    s.ndp = (float *)take(h->newdelpress, 12 * n);
    s.index = (int *)take(h->index, 4 * n);
    s.cell = (int *)take(h->cell, 4 * n);
    s.bnd = (unsigned char *)take(h->boundary, n);
    s.solid = (float *)take(h->solid, 4 * n);
    s.fluid = (float *)take(h->fluid, 4 * n);
    s.stress_tensor = (float *)take(h->stress_tensor, 36 * n);
    s.stress_rate = (float *)take(h->stress_rate, 36 * n);
    s.mass = (float *)take(h->mass, 4 * n);
    return o + 256;
}

extern "C" int fsg_upload_soa(fsg_ctx *c, const fsg_soa *h)
{
    if (!c || !h || h->n < 0 || (h->n > 0 && !h->pos)) return FSG_E_INVALID;
    const int64_t n = h->n;
    if (n > c->cap) { c->err = "fsg_upload_soa: n exceeds the context capacity"; return FSG_E_INVALID; }
    CU(c, cudaSetDevice(c->device));
    SoaStage s;
    fsg_soa hh = *h;
    // bin ids are recomputed (solver.cu:119).  Slab contexts that re-upload what they downloaded (keep_foreign) use the
    // cell array only to tell empty slots (cell == numcells + 1) from particles.
    if (!(c->cfg.world > 1 && c->keep_foreign)) hh.cell = nullptr;
    size_t bytes = soa_stage_layout(nullptr, n, &hh, s);
    int rc = ensure_stage(c, bytes);
    if (rc != FSG_OK) return rc;
    soa_stage_layout((char *)c->stage, n, &hh, s);
#define H2D(dst, src, bytes) if (src) CU(c, cudaMemcpyAsync(dst, src, (size_t)(bytes), cudaMemcpyHostToDevice, c->stream))
    H2D(s.pos, h->pos, 12 * n); H2D(s.vel, h->vel, 12 * n); H2D(s.acc, h->acc, 12 * n); H2D(s.dens, h->dens, 4 * n);
    H2D(s.press, h->press, 4 * n); H2D(s.delp, h->delpress, 12 * n); H2D(s.nd, h->newdens, 4 * n);
    H2D(s.ndp, h->newdelpress, 12 * n); H2D(s.index, h->index, 4 * n); H2D(s.bnd, h->boundary, n);
    H2D(s.solid, h->solid, 4 * n); H2D(s.fluid, h->fluid, 4 * n);
    H2D(s.stress_tensor, h->stress_tensor, 36 * n); H2D(s.stress_rate, h->stress_rate, 36 * n);
    H2D(s.mass, h->mass, 4 * n);
    if (hh.cell) CU(c, cudaMemcpyAsync(s.cell, h->cell, (size_t)(4 * n), cudaMemcpyHostToDevice, c->stream));
#undef H2D
    if (n > 0) {
        k_pack_soa<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(n, s.pos, s.vel, s.acc, s.dens, s.press, s.delp, s.nd,
                                                                       s.ndp, s.index, s.bnd, s.solid, s.fluid, s.stress_tensor, s.stress_rate,
                                                                       c->B.mix ? s.mass : nullptr, (float)c->cfg.gravity, c->B, c->carryB, c->counters + 8);
        CU(c, cudaGetLastError());
        c->launches++;
    }
    return after_upload(c, n, hh.cell ? s.cell : nullptr);
}

extern "C" int fsg_download_soa(fsg_ctx *c, fsg_soa *h)
{
    if (!c || !h) return FSG_E_INVALID;
    const int64_t n = c->n;
    h->n = n;
    CU(c, cudaSetDevice(c->device));
    if (int rm = fsg_materialize(c)) return rm;
    SoaStage s;
    size_t bytes = soa_stage_layout(nullptr, n, h, s);
    int rc = ensure_stage(c, bytes);
    if (rc != FSG_OK) return rc;
    soa_stage_layout((char *)c->stage, n, h, s);
    if (n > 0) {
        k_unpack_soa<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(n, c->B, c->carry_live ? c->carryB : nullptr, c->keysB,
                                                                         s.pos, s.vel, s.acc, s.dens, s.press, s.delp, s.nd,
                                                                         s.ndp, s.index, s.cell, s.bnd, s.solid, s.fluid, s.stress_tensor, s.stress_rate,
                                                                         c->B.mix ? s.mass : nullptr);
        CU(c, cudaGetLastError());
        c->launches++;
    }
#define D2H(dst, src, bytes) if (dst) CU(c, cudaMemcpyAsync(dst, src, (size_t)(bytes), cudaMemcpyDeviceToHost, c->stream))
    D2H(h->pos, s.pos, 12 * n); D2H(h->vel, s.vel, 12 * n); D2H(h->acc, s.acc, 12 * n); D2H(h->dens, s.dens, 4 * n);
    D2H(h->press, s.press, 4 * n); D2H(h->delpress, s.delp, 12 * n); D2H(h->newdens, s.nd, 4 * n);
    D2H(h->newdelpress, s.ndp, 12 * n); D2H(h->index, s.index, 4 * n); D2H(h->cell, s.cell, 4 * n); D2H(h->boundary, s.bnd, n);
    if (c->B.mix) { D2H(h->solid, s.solid, 4 * n); D2H(h->fluid, s.fluid, 4 * n); D2H(h->mass, s.mass, 4 * n); }
    D2H(h->stress_tensor, s.stress_tensor, 36 * n); D2H(h->stress_rate, s.stress_rate, 36 * n);
#undef D2H
    int order_bad = 0;
    CU(c, cudaMemcpyAsync(&order_bad, c->counters + 14, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CU(c, cudaStreamSynchronize(c->stream));
    if (order_bad) {
        c->err = "the key sort left the bin ids out of order (found by k_reorder): the state of this context is not valid";
        return FSG_E_STATE;
    }
    return FSG_OK;
}

// ---- per-phase device timing (fsg_set_profiling / fsg_get_phase_ms) ----
static cudaEvent_t prof_mark(fsg_ctx *c)
{
    cudaEvent_t e = nullptr;
    if (!c->ev_pool.empty()) { e = c->ev_pool.back(); c->ev_pool.pop_back(); }
    else if (cudaEventCreate(&e) != cudaSuccess) return nullptr;
    cudaEventRecord(e, c->stream);
    c->ev_used.push_back(e);
    return e;
}

static void prof_collect(fsg_ctx *c)
{
    // ev_used holds groups of 5 marks: t0 | sort | reorder | pair+update | end
    for (size_t g = 0; g + 5 <= c->ev_used.size(); g += 5) {
        float ms;
        cudaEvent_t *e = &c->ev_used[g];
        if (cudaEventElapsedTime(&ms, e[0], e[1]) == cudaSuccess) c->phase_ms[3] += ms;
        if (cudaEventElapsedTime(&ms, e[1], e[2]) == cudaSuccess) c->phase_ms[0] += ms;
        if (cudaEventElapsedTime(&ms, e[2], e[3]) == cudaSuccess) c->phase_ms[1] += ms;
        if (cudaEventElapsedTime(&ms, e[3], e[4]) == cudaSuccess) c->phase_ms[2] += ms;
        c->phase_steps++;
    }
    for (cudaEvent_t e : c->ev_used) c->ev_pool.push_back(e);
    c->ev_used.clear();
}

extern "C" int fsg_set_profiling(fsg_ctx *c, int on)
{
    if (!c) return FSG_E_INVALID;
    CU(c, cudaSetDevice(c->device));
    CU(c, cudaStreamSynchronize(c->stream));
    prof_collect(c);
    c->profiling = on != 0;
    for (double &m : c->phase_ms) m = 0;
    c->phase_steps = 0;
    return FSG_OK;
}

extern "C" int fsg_get_phase_ms(fsg_ctx *c, double ms[4], int64_t *steps)
{
    if (!c || !ms) return FSG_E_INVALID;
    CU(c, cudaSetDevice(c->device));
    CU(c, cudaStreamSynchronize(c->stream));
    prof_collect(c);
    for (int i = 0; i < 4; i++) { ms[i] = c->phase_ms[i]; c->phase_ms[i] = 0; }
    if (steps) *steps = c->phase_steps;
    c->phase_steps = 0;
    return FSG_OK;
}

// The nearly-sorted key sort (fsg_nsort.cu) — every step but the first after an upload, single-device and slab contexts alike; its
// workspace (~17 B per slot) and the third key buffer are allocated on first use.
static bool nearly_sorted_enabled(fsg_ctx *c)
{
    if (c->ns_mode < 0) {
        // FSG_SORT_MERGE = 1 / 0 forces it on / off; by default it is used where the sort's memory traffic matters (contexts of a
        // million slots and more) — below that its ~17 small launches cost more than the library sort's three
        const char *e = getenv("FSG_SORT_MERGE");
        const bool on = e ? atoi(e) != 0 : c->cap >= FSG_SORT_MERGE_MIN_CAP;
        c->ns_mode = on && c->cap >= 64;
        if (c->ns_mode) {
            c->ns_ws_bytes = fsg_nsort_bytes(c->cap);
            bool ok = cudaMalloc(&c->keysC, sizeof(int) * (size_t)c->cap) == cudaSuccess && cudaMalloc(&c->ns_ws, c->ns_ws_bytes) == cudaSuccess;
            if (!ok) {                  // no room: stay with the radix sort
                cudaGetLastError();
                cudaFree(c->keysC); cudaFree(c->ns_ws);
                c->keysC = nullptr; c->ns_ws = nullptr;
                c->ns_mode = 0;
            }
        }
    }
    return c->ns_mode == 1;
}

// Deferred update: single-device base contexts (and slab contexts on the sorted-ghost pipeline, fsg_slab2.cu) running the pipelined fp32 pair kernels (they write `sums`); FSG_DEFER_UPDATE=0
// keeps the separate k_update pass after every step.
static bool defer_enabled(fsg_ctx *c)
{
    if (c->defer_mode < 0) {
        const char *e = getenv("FSG_DEFER_UPDATE");
        c->defer_mode = e ? atoi(e) != 0 : 1;
    }
    const fsg_config &f = c->cfg;
    return c->defer_mode == 1 && (f.world == 1 || c->slab2) && f.model == FSG_MODEL_BASE && f.pair_fp64 == 0 && f.neighbour_cap == 0 && f.bin_cap == 0 &&
           !c->overlap;
}

// Makes B / keysB the post-update state of the last step (what downloads, fsg_device_ptr and a slab pack read).
int fsg_materialize(fsg_ctx *c)
{
    if (!c->deferred) return FSG_OK;
    if (c->slab2_mid) {
        c->err = "the state cannot be read between fsg_slab_pack_send and fsg_step (migrants have left, their update is pending at the neighbours)";
        return FSG_E_STATE;
    }
    CU(c, cudaSetDevice(c->device));
    CU(c, fsg_launch_update(c->dev, c->n, c->keysA, c->A, c->B, c->keysB, c->sums, c->carry_pending ? c->carryA : nullptr, 0, nullptr, c->stream));
    c->launches++;
    c->deferred = false;
    c->carry_pending = false;
    return FSG_OK;
}

// second half of a step: (sorted-ghost slabs: the neighbours' ghosts) + mykernel + mykernel2 (solver.cu:187,198)
static int step_second_half(fsg_ctx *c, int64_t n, int nxt, bool prof)
{
    if (c->slab2) { int rc = fsg_slab2_ghost_recv(c, nxt); if (rc != FSG_OK) return rc; }
    if (prof) prof_mark(c);
    const bool defer = defer_enabled(c);
    int l = 0;
    if (c->cfg.model == FSG_MODEL_UNIDYN)      // mykernel + mykernel3 + mykernel2 + cell_calc (solver-unidyn.cu:363-548)
        CU(c, fsg_launch_unidyn(c, n, c->binlist[nxt], c->counters + nxt, c->counters + 2, c->carry_live ? c->carryA : nullptr, &l,
                                c->stream));
    else if (defer) {
        // pair sums only: the update they feed runs inside the next step's reorder (or in fsg_materialize)
        // (slab2: the candidates include the ghost zones at the top of the arrays, the sums of every slot are cleared)
        CU(c, fsg_launch_pair_sums(c, c->slab2 ? c->cap : n, c->binlist[nxt], c->counters + nxt, c->counters + 2, &l, c->stream));
        c->deferred = true;
        c->carry_pending = c->carry_live;
    } else
        CU(c, fsg_launch_pair_update(c, n, c->binlist[nxt], c->counters + nxt, c->counters + 2,
                                     c->carry_live ? c->carryA : nullptr, &l, c->stream));
    c->launches += l;
    if (c->cfg.model == FSG_MODEL_UNIDYN && c->cfg.unidyn_adapt && !c->mixed) {      // children of the split particles (fsg_unidyn_adapt.cu)
        int rc = fsg_unidyn_adapt_post(c, n);
        if (rc != FSG_OK) return rc;
    }
    if (prof) prof_mark(c);
    c->carry_live = false;
    c->slab2_mid = false;
    c->cur = nxt;
    c->tables_dirty = true;
    c->keys_prev_valid = true;      // keysA = this step's sorted keys, keysB = the new keys of the same slots
    c->steps++;
    return FSG_OK;
}

// ---- the step: solver.cu:181-198 ----
extern "C" int fsg_step(fsg_ctx *c, int nsteps)
{
    if (!c || nsteps < 0) return FSG_E_INVALID;
    CU(c, cudaSetDevice(c->device));
    if (c->cfg.world > 1 && nsteps > 1) {
        c->err = "fsg_step: a slab context advances one step per pack / exchange / unpack round";
        return FSG_E_STATE;
    }
    if (c->cfg.world > 1) { int rc = fsg_slab_sticky_error(c); if (rc != FSG_OK) return rc; }
    if (c->step_pending) { c->err = "fsg_step: the previous step is half done (fsg_slab_step_finish)"; return FSG_E_STATE; }
    const int64_t n = c->n;
    if (n <= 0) { c->steps += nsteps; return FSG_OK; }
    for (int t = 0; t < nsteps; t++) {
        const bool prof = c->profiling && c->ev_used.size() < 5 * 4096;
        if (prof) prof_mark(c);
        if (c->tables_dirty) {
            // (classic slab contexts: ghost bins have table entries but are not in the home-bin list, so the reset walks the sorted
            // keys; on the sorted-ghost pipeline every own bin is a home bin and the ghost bins are reset from their messages)
            if (c->cfg.world > 1 && !c->slab2) CU(c, fsg_launch_reset_tables_keys(c->dev, c->keysA, c->start, c->end, c->n_sorted, c->stream));
            else CU(c, fsg_launch_reset_tables(c->binlist[c->cur], c->counters + c->cur, c->keysA, c->start, c->end, n, c->stream));
            c->launches++;
            if (c->slab2) { int rc = fsg_slab2_reset_ghost_tables(c); if (rc != FSG_OK) return rc; }
        }
        const int nxt = c->cur ^ 1;
        CU(c, fsg_launch_step_counters(c->counters, nxt, (int)n, c->cfg.world > 1, c->cfg.collect_stats ? c->dstats : nullptr, c->stream));
        if (prof) prof_mark(c);
        // thrust::sort_by_key, key half (solver.cu:181)
        if (c->keys_prev_valid && nearly_sorted_enabled(c)) {
            // after a step only the particles that changed bin (and, in a slab, the ghosts that came or went) are out of place:
            // flag the movers, radix-sort them, merge by ranking — hand-written, nothing read back (fsg_nsort.cu)
            int l = 0;
            CU(c, fsg_nsort(c->ns_ws, c->keysB, c->keysA, c->keysC, c->perm, n, c->sort_bits, c->sm_count, c->stream, &l));
            int *t = c->keysA; c->keysA = c->keysC; c->keysC = t;
            c->launches += l;
            c->ns_used++;
        } else {
            CU(c, fsg_sort_pairs(c->sort_tmp, c->sort_tmp_bytes, c->keysB, c->keysA, c->iota, c->perm, n, c->sort_bits, c->stream));
            c->launches++;
        }
        if (prof) prof_mark(c);
        // value half + findneighbours (solver.cu:181-182); with a deferred update also mykernel2's update of the PREVIOUS step
        const bool defer = defer_enabled(c);
        if (c->deferred) {
            CU(c, fsg_launch_reorder(c->dev, n, c->perm, c->keysA, c->A, c->B, c->carry_pending ? c->carryA : nullptr, nullptr, c->sums,
                                     defer ? c->keysB : nullptr, c->start, c->end, c->binlist[nxt], c->counters + nxt, c->binlist[nxt],
                                     c->counters + 10, c->counters + 3, c->counters + 5, c->cfg.world > 1 ? c->counters + 16 : nullptr,
                                     c->counters + 14, c->stream));
            FsgState t = c->A; c->A = c->B; c->B = t;      // A: the sorted pre-update state of THIS step; B: scratch until materialised
            c->carry_pending = false;
            c->deferred = false;
        } else
            CU(c, fsg_launch_reorder(c->dev, n, c->perm, c->keysA, c->B, c->A, c->carry_live ? c->carryB : nullptr, c->carryA, nullptr,
                                     defer ? c->keysB : nullptr, c->start, c->end, c->binlist[nxt], c->counters + nxt,
                                     c->binlistB ? c->binlistB : c->binlist[nxt], c->counters + 10, c->counters + 3, c->counters + 5,
                                     c->cfg.world > 1 ? c->counters + 16 : nullptr, c->counters + 14, c->stream));
        c->n_sorted = n;
        c->launches++;
        // sorted-ghost slab pipeline: the face layers of the sorted state go to the neighbours' ghost zones, theirs come in
        if (c->slab2) {
            int rc = fsg_slab2_ghost_send(c);
            if (rc != FSG_OK) return rc;
            if (c->slab2_split) {           // in-process slab groups: the second half runs once every slab has sent (fsg_slab_step_finish)
                c->step_pending = true;
                c->pending_nxt = nxt;
                c->pending_prof = prof;
                return FSG_OK;
            }
        }
        int rc = step_second_half(c, n, nxt, prof);
        if (rc != FSG_OK) return rc;
    }
    return FSG_OK;
}

// In-process slab groups on the sorted-ghost pipeline (several slab contexts of ONE process sharing a device, fluidsolvergpu_b200.slab.
// SlabGroup): with split steps on, fsg_step returns after the ghost send and fsg_slab_step_finish does the rest — the caller
// finishes the first half of every slab before any second half, so no kernel ever spins on a device that has yet to run (or
// lazily load) the kernels it is waiting for.  One process per device (the production layout) needs neither call.
extern "C" int fsg_slab_set_split_step(fsg_ctx *c, int on)
{
    if (!c) return FSG_E_INVALID;
    if (c->step_pending) { c->err = "fsg_slab_set_split_step: a step is half done"; return FSG_E_STATE; }
    c->slab2_split = on != 0;
    return FSG_OK;
}
extern "C" int fsg_slab_step_finish(fsg_ctx *c)
{
    if (!c) return FSG_E_INVALID;
    if (!c->step_pending) return FSG_OK;
    CU(c, cudaSetDevice(c->device));
    c->step_pending = false;
    return step_second_half(c, c->n, c->pending_nxt, c->pending_prof);
}

extern "C" int fsg_export_viz(fsg_ctx *c, float *spts, float *a3, float *b3)
{
    if (!c) return FSG_E_INVALID;
    if (c->steps < 1) { c->err = "fsg_export_viz: no step taken since upload"; return FSG_E_STATE; }
    CU(c, cudaSetDevice(c->device));
    const int64_t n = c->n;
    int rc = ensure_stage(c, (size_t)n * 20 + 1024);
    if (rc != FSG_OK) return rc;
    float *ds = (float *)c->stage, *da = ds + 3 * n, *db = da + n;
    CU(c, fsg_launch_export_viz(n, c->A.posd, c->keysA, ds, da, db, c->stream));
    c->launches++;
    if (c->cfg.model == FSG_MODEL_UNIDYN) {        // a3 = mass, b3 = |diffusion|^2 (FluidGPU-unidyn.cu:465-466)
        if (c->cfg.unidyn_adapt) CU(c, fsg_launch_export_mass(n, c->A.mix, da, c->stream));
        else CU(c, fsg_launch_fill((int *)da, 0x3f800000, n, c->stream));
        CU(c, cudaMemcpyAsync(db, c->vizb, sizeof(float) * (size_t)n, cudaMemcpyDeviceToDevice, c->stream));
        c->launches++;
    }
    if (spts) CU(c, cudaMemcpyAsync(spts, ds, 12 * (size_t)n, cudaMemcpyDeviceToHost, c->stream));
    if (a3) CU(c, cudaMemcpyAsync(a3, da, 4 * (size_t)n, cudaMemcpyDeviceToHost, c->stream));
    if (b3) CU(c, cudaMemcpyAsync(b3, db, 4 * (size_t)n, cudaMemcpyDeviceToHost, c->stream));
    CU(c, cudaStreamSynchronize(c->stream));
    return FSG_OK;
}

extern "C" int fsg_get_tables(fsg_ctx *c, int32_t *cells, int32_t *start, int32_t *end)
{
    if (!c) return FSG_E_INVALID;
    if (c->steps < 1) { c->err = "fsg_get_tables: no step taken since upload"; return FSG_E_STATE; }
    CU(c, cudaSetDevice(c->device));
    if (cells) CU(c, cudaMemcpyAsync(cells, c->keysA, sizeof(int) * (size_t)c->n, cudaMemcpyDeviceToHost, c->stream));
    if (start) CU(c, cudaMemcpyAsync(start, c->start, sizeof(int) * (size_t)c->dev.numcells, cudaMemcpyDeviceToHost, c->stream));
    if (end) CU(c, cudaMemcpyAsync(end, c->end, sizeof(int) * (size_t)c->dev.numcells, cudaMemcpyDeviceToHost, c->stream));
    CU(c, cudaStreamSynchronize(c->stream));
    return FSG_OK;
}

extern "C" int fsg_get_split(fsg_ctx *c, int32_t *split)
{
    if (!c || !split) return FSG_E_INVALID;
    if (c->cfg.model != FSG_MODEL_UNIDYN) { c->err = "fsg_get_split: unidyn model only"; return FSG_E_STATE; }
    if (c->steps < 1) { c->err = "fsg_get_split: no step taken since upload"; return FSG_E_STATE; }
    CU(c, cudaSetDevice(c->device));
    const size_t bytes = sizeof(int) * (size_t)c->dev.numcells;
    int rc = ensure_stage(c, bytes + 256);
    if (rc != FSG_OK) return rc;
    CU(c, fsg_launch_split_table(c, (int *)c->stage, c->stream));
    c->launches++;
    CU(c, cudaMemcpyAsync(split, c->stage, bytes, cudaMemcpyDeviceToHost, c->stream));
    CU(c, cudaStreamSynchronize(c->stream));
    return FSG_OK;
}

extern "C" int fsg_get_stats(fsg_ctx *c, fsg_stats *out)
{
    if (!c || !out) return FSG_E_INVALID;
    CU(c, cudaSetDevice(c->device));
    int cnt[16] = {0};
    unsigned long long ds[4] = {0, 0, 0, 0};
    CU(c, cudaMemcpyAsync(cnt, c->counters, sizeof cnt, cudaMemcpyDeviceToHost, c->stream));
    CU(c, cudaMemcpyAsync(ds, c->dstats, sizeof ds, cudaMemcpyDeviceToHost, c->stream));
    CU(c, cudaStreamSynchronize(c->stream));
    memset(out, 0, sizeof *out);
    out->n = c->n;
    out->n_live = c->steps > 0 ? cnt[3] : c->n;
    out->occupied_bins = c->steps > 0 ? cnt[c->cur] : 0;
    out->pairs_tested = (int64_t)ds[0];
    out->pairs_in_range = (int64_t)ds[1];
    out->dropped = (int64_t)ds[2];
    out->steps = c->steps;
    out->kernel_launches = c->launches;
    if (cnt[14]) {
        c->err = "the key sort left the bin ids out of order (found by k_reorder): the state of this context is not valid";
        return FSG_E_STATE;
    }
    return FSG_OK;
}

// Slab contexts: 0 (default) — an upload hands every rank the whole scene and each keeps the particles of its own slab;
// 1 — an upload returns what fsg_download_soa gave (the particles this rank holds, including ones that have just
// crossed a face and migrate at the next pack; `cell` marks the empty slots): nothing is filtered by position.
extern "C" int fsg_slab_keep_foreign(fsg_ctx *c, int on)
{
    if (!c) return FSG_E_INVALID;
    c->keep_foreign = on != 0;
    return FSG_OK;
}

extern "C" int fsg_set_collect_stats(fsg_ctx *c, int on)
{
    if (!c) return FSG_E_INVALID;
    c->cfg.collect_stats = on != 0;
    fsg_update_pair_mode(c);
    return FSG_OK;
}

extern "C" int fsg_scene_plume(fsg_ctx *c, double spacing, double jitter, uint64_t seed, int64_t *n_out)
{
    if (!c || !n_out || spacing <= 0) return FSG_E_INVALID;
    CU(c, cudaSetDevice(c->device));
    int64_t n = 0;
    int rc = fsg_scene_plume_device(c, spacing, jitter, seed, &n);
    *n_out = n;
    if (rc != FSG_OK) { c->err = rc == FSG_E_NOMEM ? "fsg_scene_plume: scene exceeds the context capacity" : "fsg_scene_plume: launch failed"; return rc; }
    return after_upload(c, n);
}

extern "C" int fsg_device_ptr(fsg_ctx *c, int which, void **ptr)
{
    if (!c || !ptr) return FSG_E_INVALID;
    if (int rm = fsg_materialize(c)) return rm;
    switch (which) {
    case 0: *ptr = c->B.posd; break;
    case 1: *ptr = c->B.velp; break;
    case 2: *ptr = c->B.accf; break;
    case 3: *ptr = c->B.dpi; break;
    case 4: *ptr = c->keysB; break;
    default: return FSG_E_INVALID;
    }
    return FSG_OK;
}

// ---- stage API (implemented in fsg_stage.cu) ----
