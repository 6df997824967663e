/*
 * fsg_oracle.h — CPU restatement of FluidSolverGPU's per-timestep particle update.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may build,
 * load or call it, and only as the checker / the reported CPU baseline.
 *
 * Parity status: pinned by (a) the known-answer values of the reference's own host-compiled
 * kernel()/kernel_derivative()/set_dens() (tests/test_oracle_kat.py) and
 * (b) golden state dumps produced by the reference's UNMODIFIED CUDA kernels driven by
 * oracle/ref_harness_base.cu on a B200 (tests/golden/, see tests/golden/README.md).
 *
 * Every function cites the reference file:line it restates (paths relative to the reference
 * repository root).  Floating-point types of every sub-expression follow the reference's C++
 * promotion rules (unsuffixed literals are double), see SURVEY.md App. A.
 */
#ifndef FSG_ORACLE_H
#define FSG_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

/* Scene / discretisation constants; defaults = FluidGPU.cuh:1-31. */
typedef struct fsgo_params {
    int    grid;           /* GRIDSIZE            FluidGPU.cuh:8  */
    float  origin;         /* XMIN=YMIN=ZMIN      FluidGPU.cuh:1-3 (int -1 → float) */
    double cellsize;       /* CELLSIZE            FluidGPU.cuh:7  */
    double h;              /* cutoff              FluidGPU.cuh:30 */
    double dt;             /* DT                  FluidGPU.cuh:31 */
    double alpha_fluid;    /* ALPHA_FLUID         FluidGPU.cuh:16 */
    double alpha_boundary; /* ALPHA_BOUNDARY      FluidGPU.cuh:17 */
    double sound;          /* SOUND               FluidGPU.cuh:11 */
    double gravity;        /* GRAVITY             FluidGPU.cuh:10 */
    int    block_threads;  /* threads per bin block = neighbour cap (64: solver.cu:187); 0 = no cap */
    int    bin_cap;        /* populations >= this are left out of `total` (64: FluidGPU.cu:174); 0 = off */
    int    threads;        /* OpenMP threads for the pair loop (0 = library default) */
} fsgo_params;

void fsgo_params_base(fsgo_params *p);

/* SoA view of the live fields of `class Particle` (FluidGPU.cuh:59-305). */
typedef struct fsgo_state {
    int    n;
    float *pos;        /* [n][3] xcoord,ycoord,zcoord */
    float *vel;        /* [n][3] */
    float *acc;        /* [n][3] */
    float *dens;       /* [n] */
    float *press;      /* [n] */
    float *delpress;   /* [n][3] x,y,z */
    float *newdens;    /* [n] accumulators (FluidGPU.cuh:144-148) */
    float *newdelpress;/* [n][3] x,y,z */
    int   *index;      /* [n] Particle::index */
    int   *cell;       /* [n] Particle::cellnumber == cells[] key array (solver.cu:146) */
    unsigned char *boundary; /* [n] */
} fsgo_state;

/* FluidGPU.cu:11-43 — smoothing kernels, host-callable in the reference too. */
float fsgo_kernel(float r);
float fsgo_kernel_test(float r);
float fsgo_kernel_derivative(float r);
/* Same with a runtime h (h = 0.06 gives identical bits). */
float fsgo_kernel_h(float r, double h);
float fsgo_kernel_derivative_h(float r, double h);

/* FluidGPU.cuh:165-167 / :256-257 */
float fsgo_set_dens(float newdens, int boundary);
float fsgo_pressure(float dens);

/* FluidGPU.cu:419 — bin id from a position. */
int fsgo_cell_id(const fsgo_params *p, float x, float y, float z);

/* One pass of the solver.cu:171-216 loop body:
 *   stable sort by cell (solver.cu:181) → findneighbours (FluidGPU.cu:106-117)
 *   → mykernel (FluidGPU.cu:119-285) → mykernel2 (FluidGPU.cu:404-432).
 * State arrays are permuted in place into the sorted order, as the reference does.
 * Optional outputs (may be NULL):
 *   cells_sorted[n], start[numcells], end[numcells] : as they are between findneighbours and mykernel2
 *   spts[3n], a3[n], b3[n]                          : mykernel2's viz export
 *   stats[4]: {pairs tested, pairs in range, candidates dropped by the thread cap, occupied bins}
 * Particles whose new bin id falls outside [0,numcells) would make the reference write out of
 * bounds (FluidGPU.cu:110); the restatement parks them (cell = numcells) and never updates them again.
 * Returns 0, or -1 on allocation failure. */
int fsgo_base_step(const fsgo_params *p, fsgo_state *s,
                   int *cells_sorted, int *start, int *end,
                   float *spts, float *a3, float *b3, long long *stats);

/* ---- unidyn model (FluidGPU-unidyn.cu / .cuh), see fsg_oracle_unidyn.c ---- */
void fsgo_params_unidyn(fsgo_params *p);   /* FluidGPU-unidyn.cuh:1-36; alpha_boundary = ALPHA__SAND_BOUNDARY */

typedef struct fsgo_ustate {
    int    n;
    float *pos, *vel, *acc, *dens, *press, *delpress, *newdens, *newdelpress;   /* as fsgo_state */
    int   *index, *cell;
    unsigned char *boundary;
    float *solid;      /* [n] FluidGPU-unidyn.cuh:180 */
    float *fluid;      /* [n] :181 */
    int   *subindex;   /* [n] octant of the particle inside a split bin (:119, FluidGPU-unidyn.cu:182-184) */
    /* mixed-phase / granular scenes (may be NULL for pure-fluid scenes): FluidGPU-unidyn.cuh stress_tensor[3][3], stress_rate[3][3] */
    float *stress_tensor; /* [n][9] row major */
    float *stress_rate;   /* [n][9] */
    float *mass;          /* [n] Particle::mass (FluidGPU-unidyn.cuh:150); NULL = 1 for every particle */
} fsgo_ustate;

/* One pass of the solver-unidyn.cu:313-573 loop body on one device: sort (:331) -> count_after_merge (:341)
 * -> findneighbours (:354) -> mykernel (:363) -> mykernel3 (:379) -> mykernel2 (:389) -> cell_calc (:548).
 * split_out[numcells]: split[] as mykernel leaves it (bin id for bins with more than 6 particles, else -1).
 * viz: spts = pre-update positions, a3 = mass, b3 = |diffusion|^2 (FluidGPU-unidyn.cu:462-466).
 * Scenes with a non-boundary particle of solid != 0 take the race-free two-pass reading of the mixed-phase / granular terms
 * (see fsg_oracle_unidyn.c) and need stress_tensor / stress_rate.
 * Returns 0; -1 allocation failure; -2 scene outside the restated scope (see fsg_oracle_unidyn.c). */
int fsgo_unidyn_step(const fsgo_params *p, fsgo_ustate *s, int t, int *cells_sorted, int *start, int *end, int *split,
                     float *spts, float *a3, float *b3, long long *stats);

/* Particle merging / splitting made live (FluidGPU-unidyn.cu:260-285 == :680-700, host half solver-unidyn.cu:495-542).
 * In the reference the merge test is `ds <= (-10.00) && ds > 0` (never true), a merge would set mass 2.75 while a split needs
 * mass > 3, and the host loop that creates the second particle is commented out: with the reference's literals nothing happens.
 * Here the two thresholds are parameters (the literals are the defaults) and the blocks get a race-free reading, see
 * fsg_oracle_unidyn.c.  PARITY UNPINNED against the reference (it has no live behaviour to compare with); pinned oracle <-> CUDA. */
typedef struct fsgo_adapt {
    double merge_distance;   /* :261  literal -10.00 */
    double split_mass_min;   /* :278  literal 3 */
    int    capacity;         /* slots the state arrays (and mass) hold: children are appended behind s->n while there is room */
    int    next_index;       /* Particle::index given to the next child (in/out) */
    int    merged, split, added;   /* out: pairs merged, particles split, children created in this step */
} fsgo_adapt;
/* fsgo_unidyn_step with the merge / split blocks (pure-fluid scenes, s->mass required).  s->n grows by ad->added. */
int fsgo_unidyn_step_adapt(const fsgo_params *p, fsgo_ustate *s, fsgo_adapt *ad, int t, int *cells_sorted, int *start, int *end, int *split,
                           float *spts, float *a3, float *b3, long long *stats);

#ifdef __cplusplus
}
#endif
#endif
