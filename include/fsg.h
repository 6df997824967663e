/*
 * fsg.h — C ABI of libfsg, the B200-native (sm_100a) implementation of FluidSolverGPU's
 * per-timestep particle update.
 *
 * What it replaces (paths relative to the reference repository root):
 *   the body of the time loops solver.cu:171-216 and solver-unidyn.cu:313-573, i.e. the launches of
 *   thrust::sort_by_key (solver.cu:181), findneighbours (FluidGPU.cuh:417 / FluidGPU.cu:106),
 *   mykernel (FluidGPU.cuh:418 / FluidGPU.cu:119), mykernel2 (FluidGPU.cuh:419 / FluidGPU.cu:404)
 *   and the unidyn counterparts (FluidGPU-unidyn.cuh:537-544).
 *
 * Two faces:
 *   (1) context API  — fsg_create / fsg_upload_* / fsg_step / fsg_download_* : the library owns SoA
 *       device state and runs the whole step; this is what drivers, benchmarks and multi-GPU use.
 *   (2) stage API    — fsg_stage_* : one call per reference kernel, on caller-owned DEVICE buffers
 *       laid out exactly as the reference's (340-byte `Particle` AoS, int key/start/end arrays), so
 *       the reference drivers can swap each `<<<>>>` launch for one call.
 *
 * Conventions: every function returns FSG_OK (0) or a negative FSG_E_* code and never exits or
 * throws; fsg_last_error() gives a message.  Plain pointers and sizes only.  One host thread per
 * context (the reference drives everything from one thread, solver.cu:171).  There is NO CPU
 * fallback: without a CUDA device every compute entry point returns FSG_E_NO_DEVICE.
 */
#ifndef FSG_H
#define FSG_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FSG_VERSION 100

#if defined(__GNUC__)
#define FSG_API __attribute__((visibility("default")))
#else
#define FSG_API
#endif

enum {
    FSG_OK = 0,
    FSG_E_INVALID = -1,    /* bad argument */
    FSG_E_NO_DEVICE = -2,  /* no usable CUDA device (there is no CPU path) */
    FSG_E_CUDA = -3,       /* a CUDA call failed; see fsg_last_error */
    FSG_E_NOMEM = -4,
    FSG_E_STATE = -5,      /* call order (e.g. step before upload) */
    FSG_E_UNSUPPORTED = -6
};

enum { FSG_MODEL_BASE = 0 /* FluidGPU.cu */, FSG_MODEL_UNIDYN = 1 /* FluidGPU-unidyn.cu */ };

/* sizeof(Particle) in both reference headers (FluidGPU.cuh:59-305, FluidGPU-unidyn.cuh:68-424). */
#define FSG_AOS_STRIDE 340

/* Run-time form of the reference's compile-time constants (FluidGPU.cuh:1-31,
 * FluidGPU-unidyn.cuh:1-36).  fsg_config_default() fills in the reference's values. */
typedef struct fsg_config {
    int32_t model;          /* FSG_MODEL_*                                                   */
    int32_t grid;           /* GRIDSIZE: bins per axis                        FluidGPU.cuh:8  */
    float   origin;         /* XMIN = YMIN = ZMIN                             FluidGPU.cuh:1  */
    double  cellsize;       /* CELLSIZE                                       FluidGPU.cuh:7  */
    double  h;              /* cutoff (smoothing length)                      FluidGPU.cuh:30 */
    double  dt;             /* DT                                             FluidGPU.cuh:31 */
    double  gravity;        /* GRAVITY                                        FluidGPU.cuh:10 */
    double  sound;          /* SOUND                                          FluidGPU.cuh:11 */
    double  alpha_fluid;    /* ALPHA_FLUID                                    FluidGPU.cuh:16 */
    double  alpha_boundary; /* ALPHA_BOUNDARY                                 FluidGPU.cuh:17 */
    int32_t neighbour_cap;  /* threads per bin block in the reference launch = max neighbour
                               particles visited per bin (64: solver.cu:187; 1024: solver-unidyn.cu:363);
                               0 = visit every particle of the 27 bins                        */
    int32_t bin_cap;        /* bins with >= this many particles are left out of the thread
                               count (64: FluidGPU.cu:174); 0 = off                           */
    int64_t capacity;       /* particles this context can hold                                */
    int32_t device;         /* CUDA device ordinal                                            */
    int32_t pair_fp64;      /* 0: fp32 production kernel (default); 1: evaluate the pair sub-expressions
                               the reference evaluates in double in double;
                               2: the kernel of mode 1 in fp32 (cross-check)                  */
    int32_t collect_stats;  /* 1: count candidate / in-range pairs each step                   */
    /* slab decomposition along x, the slowest bin axis (solver-unidyn.cu:187-193) */
    int32_t rank, world;    /* this context's slab and the number of slabs (1 = no decomposition) */
    int32_t slab_x0, slab_x1; /* world > 1: this slab owns the bin layers slab_x0 <= ix < slab_x1 (ix = bin id / grid^2) */
    int32_t pair_mode;      /* uncapped fp32 configuration only.  0 (default): symmetric pair kernel — a pair of particles in
                               two bins is evaluated once and added to both through float reductions, so sums are
                               reproducible to rounding (~1e-7), not bit for bit; 1: deterministic gather kernel
                               (every particle sums its own 27 bins in a fixed order)                         */
    int32_t unidyn_open_box; /* unidyn model.  0 (default): Particle::update keeps the reference's literal unit-box floor and walls
                               (z < -0.89, |x|,|y|,|z| > 0.98, FluidGPU-unidyn.cuh:332,404-411); 1: they are switched off, for domains
                               other than [-1,1]^3 (the reference has to be rebuilt for that: oracle/Makefile, ref_harness_unidyn_g128) */
    int32_t unidyn_adapt;    /* unidyn model, single-device contexts, pure-fluid scenes.  1: the particle merging / splitting blocks of
                               FluidGPU-unidyn.cu:260-285 + solver-unidyn.cu:495-542 run every step (race-free reading); uploads may
                               then carry mass != 1 and the particle count can grow up to `capacity`.  0 (default): mass must be 1 */
    double  unidyn_merge_distance;  /* the literal -10.00 of FluidGPU-unidyn.cu:261 (default: never merges, like the reference) */
    double  unidyn_split_mass_min;  /* the literal 3 of FluidGPU-unidyn.cu:278 (default: a merged particle of mass 2.75 never splits, like the reference) */
} fsg_config;

/* Host-side structure-of-arrays view used by fsg_upload_soa / fsg_download_soa: the live fields
 * of `class Particle`.  Any pointer may be NULL on upload (class defaults are used: dens = RHO_0,
 * press = 0, acc = (0,0,GRAVITY), newdens = RHO_0, FluidGPU.cuh:64-71,132-148) or on download. */
typedef struct fsg_soa {
    int64_t n;
    float  *pos;         /* [n][3] xcoord,ycoord,zcoord   */
    float  *vel;         /* [n][3]                         */
    float  *acc;         /* [n][3]                         */
    float  *dens;        /* [n]                            */
    float  *press;       /* [n]                            */
    float  *delpress;    /* [n][3] x,y,z                   */
    float  *newdens;     /* [n] accumulator carried into the next step */
    float  *newdelpress; /* [n][3] x,y,z                   */
    int32_t *index;      /* [n] Particle::index (NULL on upload: 0..n-1) */
    int32_t *cell;       /* [n] Particle::cellnumber (download only; recomputed on upload, solver.cu:119) */
    uint8_t *boundary;   /* [n]                            */
    /* unidyn model only (FluidGPU-unidyn.cuh:180-181); NULL on upload: 0/1 for fluid, 1/0 for boundary particles */
    float  *solid;       /* [n] */
    float  *fluid;       /* [n] */
    /* unidyn, granular state of mixed-phase scenes (FluidGPU-unidyn.cuh stress_tensor[3][3], stress_rate[3][3]); NULL on upload: zeros */
    float  *stress_tensor; /* [n][9] row major */
    float  *stress_rate;   /* [n][9] */
    float  *mass;          /* [n] unidyn, Particle::mass (FluidGPU-unidyn.cuh:150); NULL on upload: 1.  Anything but 1 needs unidyn_adapt */
} fsg_soa;

typedef struct fsg_ctx fsg_ctx;

typedef struct fsg_stats {
    int64_t n;               /* particles held                                   */
    int64_t n_live;          /* particles inside the bin grid                    */
    int64_t occupied_bins;   /* bins with at least one particle, last step       */
    int64_t pairs_tested;    /* candidate pairs distance-tested, last step (collect_stats) */
    int64_t pairs_in_range;  /* pairs with 0 < ds <= 2h, last step (collect_stats)         */
    int64_t dropped;         /* neighbour particles left unvisited by neighbour_cap, last step (collect_stats) */
    int64_t steps;           /* steps taken since upload                         */
    int64_t kernel_launches; /* launches of libfsg's own kernels since create    */
} fsg_stats;

/* ---- (1) context API ---- */
FSG_API int  fsg_version(void);
FSG_API int  fsg_device_count(void);                       /* 0 when there is no usable device */
FSG_API int  fsg_config_default(fsg_config *cfg, int model);
FSG_API int  fsg_create(const fsg_config *cfg, fsg_ctx **out);
FSG_API int  fsg_destroy(fsg_ctx *ctx);
FSG_API const char *fsg_last_error(const fsg_ctx *ctx);    /* ctx may be NULL: last create error */
FSG_API int  fsg_set_stream(fsg_ctx *ctx, void *cuda_stream);   /* cudaStream_t; default: a stream the ctx owns */
FSG_API void *fsg_get_stream(fsg_ctx *ctx);

/* Host -> device.  `particles` is an array of n reference `Particle` records (FSG_AOS_STRIDE bytes
 * each), as solver.cu:131 copies to the device.  Bin ids are recomputed from the positions with the
 * expression of solver.cu:119. */
FSG_API int  fsg_upload_aos(fsg_ctx *ctx, const void *particles, int64_t n);
FSG_API int  fsg_upload_soa(fsg_ctx *ctx, const fsg_soa *host);
/* Device -> host, in the device's current (bin-sorted) order, like copying d_SPptr back
 * (solver-unidyn.cu:475).  Fields the path never touches get the class defaults. */
FSG_API int  fsg_download_aos(fsg_ctx *ctx, void *particles, int64_t n);
FSG_API int  fsg_download_soa(fsg_ctx *ctx, fsg_soa *host);

/* nsteps passes of the loop body solver.cu:181-198: sort by bin -> bin ranges -> pair sums ->
 * EOS / integrate / re-bin.  Asynchronous on the context's stream; fsg_sync waits. */
FSG_API int  fsg_step(fsg_ctx *ctx, int nsteps);
FSG_API int  fsg_sync(fsg_ctx *ctx);

/* mykernel2's visualisation export of the LAST step (FluidGPU.cu:410-414): positions, dens and
 * float(cellnumber) before update(), in that step's sorted order.  Host pointers, any may be NULL. */
FSG_API int  fsg_export_viz(fsg_ctx *ctx, float *spts, float *a3, float *b3);
/* The integer tables of the LAST step as the pair kernel saw them (host pointers, may be NULL):
 * cells[n] sorted keys, start/end[numcells] (FluidGPU.cu:106-117; -1 = empty bin). */
FSG_API int  fsg_get_tables(fsg_ctx *ctx, int32_t *cells, int32_t *start, int32_t *end);
/* unidyn: split[numcells] as mykernel leaves it (FluidGPU-unidyn.cu:181-190) — the bin id for bins with more
 * than 6 particles in the LAST step (their particles only see the 8 bins of their octant), else -1. */
FSG_API int  fsg_get_split(fsg_ctx *ctx, int32_t *split);
FSG_API int  fsg_get_stats(fsg_ctx *ctx, fsg_stats *out);
/* Switches the pair counters (fsg_config.collect_stats) on or off for the following steps. */
FSG_API int  fsg_set_collect_stats(fsg_ctx *ctx, int on);
/* Device-side phase timing (CUDA events on the context's stream around each phase of every step;
 * the reference prints the same kind of figure, solver.cu:175-197).  fsg_get_phase_ms synchronises,
 * returns the milliseconds accumulated since the last call — ms[0] key sort, ms[1] reorder + bin
 * tables, ms[2] pair sums + update, ms[3] table reset / bookkeeping — and clears them. */
FSG_API int  fsg_set_profiling(fsg_ctx *ctx, int on);
FSG_API int  fsg_get_phase_ms(fsg_ctx *ctx, double ms[4], int64_t *steps);

/* Device-side scene generation for the throughput configs: a column of fluid
 * particles about the z axis of a grid^3 bin domain, lattice spacing `spacing`, jitter from
 * splitmix64(seed + lattice id).  Returns the particle count through *n_out. */
FSG_API int  fsg_scene_plume(fsg_ctx *ctx, double spacing, double jitter, uint64_t seed, int64_t *n_out);
/* Same generator on the host into caller arrays (for the end-to-end path and the tests);
 * pass pos == NULL to get the count only.  Pure host code, no device needed. */
FSG_API int  fsg_scene_plume_host(const fsg_config *cfg, double spacing, double jitter, uint64_t seed,
                          float *pos, float *vel, int64_t capacity, int64_t *n_out);

/* Particles per bin layer ix of that scene (hist[grid]; lattice positions): the host cuts the domain
 * into slabs of equal particle count with it.  Pure host code. */
FSG_API int  fsg_scene_plume_hist(const fsg_config *cfg, double spacing, int64_t *hist);

/* Device pointers to the SoA state (float4 arrays, see DESIGN.md "Data layout"), for zero-copy
 * consumers on the same device (bench, halo exchange).  which: 0 = posd, 1 = velp, 2 = accf, 3 = dpi,
 * 4 = keys (int32).  Valid until the next fsg_* call that changes state. */
FSG_API int  fsg_device_ptr(fsg_ctx *ctx, int which, void **ptr);

/* ---- frame output: the file write_point_mesh (visit_writer.h:94-96, visit_writer.cpp:673-719) produces ----
 * Byte-identical to the reference's VisIt writer for the same arrays (legacy VTK UNSTRUCTURED_GRID point cloud,
 * ASCII "%20.12e" or big-endian binary), same argument list; returns FSG_E_INVALID when the file cannot be
 * opened (the reference crashes).  Pure host code.  fsg_write_frame = fsg_export_viz + the call the drivers make
 * (solver-unidyn.cu:487: "mass", "surface_level"; solver.cu:213: "dens", "cellnumber"). */
FSG_API int  fsg_write_point_mesh(const char *filename, int use_binary, int npts, const float *pts, int nvars, const int *vardim,
                                  const char *const *varnames, const float *const *vars);
FSG_API int  fsg_write_frame(fsg_ctx *ctx, const char *filename, int use_binary);
/* The same frame OFF the step loop's critical path (the reference dumps synchronously inside its time loop, solver-unidyn.cu:472-493):
 * one small export kernel on the context's stream, the device-to-host copy on a second stream behind an event into one of two pinned
 * staging slots, formatting and file I/O on a writer thread.  Returns at once; a third frame in flight waits for the oldest one.
 * fsg_frame_wait blocks until every frame is on disk, gives the number written so far and returns (and clears) the first error. */
FSG_API int  fsg_write_frame_async(fsg_ctx *ctx, const char *filename, int use_binary);
FSG_API int  fsg_frame_wait(fsg_ctx *ctx, int64_t *frames_written);

/* ---- slab decomposition along x (world > 1): the multi-device hand-off of solver-unidyn.cu:396-470 ----
 * Every step of a slab context is   fsg_slab_pack -> (caller moves the two messages to the x-neighbours,
 * e.g. NCCL send/recv) -> fsg_slab_unpack -> fsg_step(ctx, 1).   ALL of it is asynchronous on the
 * context's stream: no call reads anything back, so the host can run steps ahead of the device.
 * fsg_slab_pack classifies the particles by their new bin layer: those that left the slab are sent with
 * their full 64-byte state (migrants; the sender keeps them one more step as ghosts), those in the slab's
 * outermost layers are copied with the 32-byte read state the neighbour's pair sums need (ghosts; the
 * reference ships a one-layer `buffer` of whole Particle records instead, solver-unidyn.cu:187,421-462).
 * Message = device memory of fsg_slab_message_bytes(cap_m, cap_g) bytes, the SAME size on every rank:
 *   [64-byte header: int64 migrants, int64 ghosts][posd cap_m][velp cap_m][accf cap_m][dpi cap_m][posd cap_g][velp cap_g]
 *   [64-byte tail: int64 stamp — the exchange sequence number, copied after the rest (peer-memory variant)]
 * The receiver reads the counts from the header on the device.  Order inside the messages is the
 * particles' current order (two-phase count / scan / scatter: deterministic).  A slab context always
 * works on `capacity` slots; unused slots hold a dead key and sort last.
 * fsg_slab_check synchronises and reports overflow / ghost-band violations (see fsg_slab.cu). */
FSG_API int  fsg_slab_pack(fsg_ctx *ctx, void *d_to_left, void *d_to_right, int64_t cap_m, int64_t cap_g);
FSG_API int  fsg_slab_unpack(fsg_ctx *ctx, const void *d_from_left, const void *d_from_right, int64_t cap_m, int64_t cap_g);
FSG_API int  fsg_slab_check(fsg_ctx *ctx, int64_t info[9]);
/* Uploads into a slab context: 0 (default) every rank is handed the whole scene and keeps the particles of its slab;
 * 1 the upload returns what fsg_download_soa gave — all `capacity` slots, `cell` == grid^3 + 1 marking the empty ones —
 * and nothing is filtered by position (particles that have just crossed a face migrate at the next pack). */
FSG_API int  fsg_slab_keep_foreign(fsg_ctx *ctx, int on);
FSG_API int64_t fsg_slab_message_bytes(int64_t cap_m, int64_t cap_g);                 /* base model */
/* unidyn messages also carry the volume fractions (16 bytes more per migrant and per ghost) */
FSG_API int64_t fsg_slab_message_bytes_model(int model, int64_t cap_m, int64_t cap_g);
/* Peer-memory variant (one process per GPU on one node): the library owns the message buffers, the
 * neighbours' inboxes are mapped through CUDA IPC and fsg_slab_pack_send copies the packed messages straight
 * into them over NVLink (copy engines).  The caller exchanges the 64-byte handles once
 * (fsg_slab_inbox_handle -> neighbour -> fsg_slab_open_peer) and, every step, orders the neighbour's stream
 * behind the copy with any stream-ordered signal (a few-byte NCCL send/recv).  side 0 = left, 1 = right;
 * inboxes are double-buffered by step parity. */
FSG_API int  fsg_slab_alloc_messages(fsg_ctx *ctx, int64_t cap_m, int64_t cap_g);
/* For the base model the library-owned messages run the SORTED-GHOST pipeline (fsg_slab2.cu) when they are allocated before the
 * first upload: ghosts never pass through the sort.  fsg_slab_pack_send moves only the migrants (pre-update state + pending pair
 * sums; the update is deferred like on one device), fsg_slab_unpack_recv appends them, and fsg_step — after its reorder — copies
 * the face layers of the SORTED state straight into the neighbours' ghost zones (the top 2 * cap_g slots of the particle arrays:
 * the context works on capacity - 2 * cap_g slots), waits for theirs on the device and adds their bins to the tables.  Every rank
 * has to make the same calls; the state cannot be read between fsg_slab_pack_send and fsg_step.
 * fsg_slab_alloc_messages2: flags bit 0 = keep the classic pipeline (ghosts appended and sorted; needed by fsg_slab_set_overlap).
 * fsg_slab_mode: 0 not a slab context, 1 classic, 2 sorted ghosts.  fsg_slab_get_ghost_ms: mean device time per step of the
 * ghost exchange inside fsg_step (send + wait for the neighbours + install) while profiling is on. */
FSG_API int  fsg_slab_alloc_messages2(fsg_ctx *ctx, int64_t cap_m, int64_t cap_g, int flags);
FSG_API int  fsg_slab_mode(fsg_ctx *ctx);
FSG_API int  fsg_slab_get_ghost_ms(fsg_ctx *ctx, double *ms, int64_t *steps);
/* Several slab contexts of ONE process on one device (tests): with split steps on, fsg_step returns after its ghost send and
 * fsg_slab_step_finish does the rest (wait, install, pair sums) — finish the first half of every slab before any second half, so
 * that no kernel spins on a device that still has to run (or lazily load) what it waits for.  One process per device needs neither. */
FSG_API int  fsg_slab_set_split_step(fsg_ctx *ctx, int on);
FSG_API int  fsg_slab_step_finish(fsg_ctx *ctx);
FSG_API int  fsg_slab_inbox_handle(fsg_ctx *ctx, int side, int parity, void *handle64);
FSG_API int  fsg_slab_open_peer(fsg_ctx *ctx, int side, int parity, const void *handle64);
FSG_API int  fsg_slab_pack_send(fsg_ctx *ctx);
FSG_API int  fsg_slab_unpack_recv(fsg_ctx *ctx);
FSG_API int  fsg_slab_close_peers(fsg_ctx *ctx);
/* Several slab contexts inside ONE process (tests on one GPU): wire the neighbour's inbox as a plain device pointer. */
FSG_API int  fsg_slab_set_peer(fsg_ctx *ctx, int side, int parity, void *neighbour_inbox);
FSG_API void *fsg_slab_inbox_ptr(fsg_ctx *ctx, int side, int parity);
/* on: fsg_step computes the slab's boundary bins first and issues the NEXT step's pack + peer copies on a second
 * stream, beside the interior bins; the following fsg_slab_pack_send is then a no-op.  Needs the peer-memory exchange. */
FSG_API int  fsg_slab_set_overlap(fsg_ctx *ctx, int on);

/* ---- (2) stage API: caller-owned DEVICE buffers in the reference's own layout ---- */
/* replaces thrust::sort_by_key(t_v, t_v + n, t_a)            solver.cu:181 */
FSG_API int  fsg_stage_sort(fsg_ctx *ctx, int32_t *d_cells, void *d_particles, int64_t n);
/* replaces findneighbours<<<NUMCELLS,1024>>>(v_d, d_start, d_end, n)   solver.cu:182 */
FSG_API int  fsg_stage_findneighbours(fsg_ctx *ctx, const int32_t *d_cells, int32_t *d_start, int32_t *d_end, int64_t n);
/* replaces mykernel<<<NUMCELLS,64>>>(d_SPptr, v_d, d_start, d_end, n)  solver.cu:187 */
FSG_API int  fsg_stage_mykernel(fsg_ctx *ctx, void *d_particles, const int32_t *d_cells, const int32_t *d_start,
                        const int32_t *d_end, int64_t n);
/* replaces mykernel2<<<NUMCELLS,1024>>>(d_SPptr, v_d, d_start, d_end, n, spts, a3, b3)  solver.cu:198 */
FSG_API int  fsg_stage_mykernel2(fsg_ctx *ctx, void *d_particles, int32_t *d_cells, int32_t *d_start, int32_t *d_end,
                         int64_t n, float *spts, float *a3, float *b3);


/* unidyn_adapt contexts: pairs merged / particles split / children created — in the last step and since the upload */
FSG_API int  fsg_unidyn_adapt_counts(fsg_ctx *ctx, int64_t last[3], int64_t total[3]);

/* unidyn model (context created with FSG_MODEL_UNIDYN): the launches of the single-device loop solver-unidyn.cu:341-548
 * on the reference's own buffers (340-byte unidyn Particle records).  Same scope as the context API: every non-boundary
 * particle pure fluid with mass 1; the caller's thrust sorts (:331, :378) stay where they are
 * (fsg_stage_sort does the first one for either model). */
/* replaces count_after_merge<<<NUMCELLS,1024>>>(v_d, d_particleindex, dsz, newsize)   solver-unidyn.cu:341, FluidGPU-unidyn.cuh:544 */
FSG_API int  fsg_stage_unidyn_count_after_merge(fsg_ctx *ctx, const int32_t *d_cells, int64_t n, int32_t *d_newsize);
/* replaces findneighbours<<<NUMCELLS,1024>>>(v_d, d_start, d_start_copy, d_end, dsz, x)   solver-unidyn.cu:354, FluidGPU-unidyn.cuh:537 */
FSG_API int  fsg_stage_unidyn_findneighbours(fsg_ctx *ctx, const int32_t *d_cells, int32_t *d_start, int32_t *d_start_copy,
                                             int32_t *d_end, int64_t n, int32_t x);
/* replaces mykernel<<<NUMCELLS,1024>>>(d_SPptr, pidx, v_d, d_start, d_end, d_split, dsz, x, dev, buffer, d_numsplit)
 * solver-unidyn.cu:363, FluidGPU-unidyn.cuh:538: split marking + subindex + pair sums of the unsplit bins */
FSG_API int  fsg_stage_unidyn_mykernel(fsg_ctx *ctx, void *d_particles, const int32_t *d_cells, const int32_t *d_start,
                                       const int32_t *d_end, int32_t *d_split, int32_t *d_numsplit, int64_t n);
/* replaces mykernel3<<<numsplit*8,1024>>>(same arguments)   solver-unidyn.cu:379, FluidGPU-unidyn.cuh:539: pair sums of the
 * split bins (each particle sees the 8 bins of its octant); d_split is not needed (bins with more than 6 particles) */
FSG_API int  fsg_stage_unidyn_mykernel3(fsg_ctx *ctx, void *d_particles, const int32_t *d_cells, const int32_t *d_start,
                                        const int32_t *d_end, int64_t n);
/* replaces mykernel2<<<NUMCELLS,1024>>>(d_SPptr, pidx, v_d, d_start_copy, d_start, d_end, d_split, d_numsplit, dsz, x, dev,
 * buffer, t, spts, a3, b3)   solver-unidyn.cu:389, FluidGPU-unidyn.cuh:540 */
FSG_API int  fsg_stage_unidyn_mykernel2(fsg_ctx *ctx, void *d_particles, int32_t *d_cells, int32_t *d_start_copy, int32_t *d_start,
                                        int32_t *d_end, int32_t *d_split, int32_t *d_numsplit, int64_t n, int32_t x, int32_t t,
                                        float *spts, float *a3, float *b3);
/* replaces cell_calc<<<NUMCELLS,1024>>>(d_SPptr, pidx, v_d, dsz, dev)   solver-unidyn.cu:548, FluidGPU-unidyn.cuh:543 */
FSG_API int  fsg_stage_unidyn_cell_calc(fsg_ctx *ctx, void *d_particles, int32_t *d_cells, int64_t n);

#ifdef __cplusplus
}
#endif
#endif /* FSG_H */
