/*
 * fsg_oracle_unidyn.c — CPU restatement of FluidSolverGPU's "unidyn" per-timestep particle update
 * (FluidGPU-unidyn.cu / FluidGPU-unidyn.cuh, driven by solver-unidyn.cu:313-573).
 *
 * TEST INFRASTRUCTURE ONLY (see fsg_oracle.h).  Every function cites the reference file:line it follows.
 *
 * Scope of the restatement: mass == 1 (merging is unreachable, `ds <= -10 && ds > 0`, :261, and splitting needs mass > 3, :278).
 * For scenes in which every non-boundary particle is pure fluid (solid == 0) — the default scene of solver-unidyn.cu:127-184 —
 * the mixed-phase block (FluidGPU-unidyn.cu:317-357), mixfactor (:368), vel_grad (:369-377), stress_accel (:379-381),
 * mixture_accel (:391-398), delsolid (:400) and the granular stress update (:410-446) contribute exactly zero; that part is
 * PINNED by dumps of the reference's own kernels (tests/golden/ref_config2_*, ref_unidyn_random_*).
 *
 * MIXED-PHASE / GRANULAR scenes (some non-boundary particle with solid != 0; the caller passes stress_tensor / stress_rate): the
 * reference's result is NOT a function of its input there — mixture_accel / delsolid / delfluid read the drift velocities of both
 * particles (:385-401) while other blocks of the same launch are still accumulating them (:351-357), and the stress update (:410-446,
 * repeated at the end of mykernel3, :832-868) reads vel_grad sums that other blocks are still adding to.  This restatement fixes the
 * RACE-FREE reading the author evidently meant:  pass A accumulates newdens / newdelpress / diffusion / drift velocities / vel_grad
 * / stress_accel over all pairs;  pass B evaluates mixture_accel / delsolid / delfluid with the COMPLETED drift velocities of both
 * particles;  then ONE stress update per particle with the completed vel_grad;  then mykernel2 + update.  Every expression follows
 * the reference text (types included).  PARITY UNPINNED against the reference for these terms (it cannot be pinned: two runs of
 * the reference differ at O(1) in them); the CUDA path is pinned against THIS restatement.
 */
#include "fsg_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

void fsgo_params_unidyn(fsgo_params *p)
{
    p->grid = 17;               /* GRIDSIZE   FluidGPU-unidyn.cuh:8  */
    p->origin = -1.0f;          /* XMIN       :1 */
    p->cellsize = 0.12;         /* CELLSIZE   :7 */
    p->h = 0.06;                /* cutoff     :35 */
    p->dt = 0.0018;             /* DT         :36 */
    p->alpha_fluid = -0.0155e1; /* :17 */
    p->alpha_boundary = 100e-1; /* ALPHA__SAND_BOUNDARY :21 — the factor the unidyn pair term uses (FluidGPU-unidyn.cu:307) */
    p->sound = 1450.0;          /* :11 */
    p->gravity = -9.8;          /* :10 */
    p->block_threads = 1024;    /* solver-unidyn.cu:363 */
    p->bin_cap = 0;
    p->threads = 0;
}

/* FluidGPU-unidyn.cuh:183-185 */
static float u_set_dens(float x, int boundary, double h)
{
    return (x + fsgo_kernel_h(0, h)) / 23.0 * (1 + (float)(boundary != 0) * 1.5) + 9250;
}
/* FluidGPU-unidyn.cuh:282-284: double pow, blended by `solid` (RHO_0_SAND == RHO_0) */
static float u_pressure(float dens, float solid, double sound)
{
    return (1 - solid) * 1000 * pow(sound, 0) * 9550 / 7.0 * (pow(dens / 9550, 7) - 1) +
           (solid) * 1000 * pow(sound, 0) * 9550 / 7.0 * (pow(dens / 9550, 7) - 1);
}

typedef struct {
    float dens, px, py, pz, dx, dy, dz, delfluid;
    /* mixed-phase / granular accumulators */
    float sdrift[3], fdrift[3], vel_grad[9], stress_accel[3], mix[3], delsolid;
} upair_acc;

/* reference constants, FluidGPU-unidyn.cuh:24-33 */
#define U_C1 1.5e1
#define U_C2 0e6
#define U_C3 5e1
#define U_PHI 1.23
#define U_KC 1e9
#define U_MIXPRESSURE 1e-12
#define U_MIXBROWNIAN 5e-9
#define U_RHO_0 9550
#define U_RHO_0_SAND 9550

/* Pass A, mixed-phase / granular terms of one in-range pair (FluidGPU-unidyn.cu:314-357, 368-381), added to the live terms of
 * upair_body.  i = home particle (ii in the reference), j = neighbour. */
static inline void upair_mixed_a(const fsgo_params *P, const fsgo_ustate *s, int i, int j, float dkx, float dky, float dkz,
                                 float vabx, float vaby, float vabz, upair_acc *a)
{
    const int bi = s->boundary[i] != 0, bj = s->boundary[j] != 0;
    const float solid_i = s->solid[i], fluid_i = s->fluid[i], solid_j = s->solid[j], fluid_j = s->fluid[j];
    const float dens_i = s->dens[i], press_i = s->press[i], press_j = s->press[j];
    const float *vi = s->vel + 3 * (size_t)i, *dpi = s->delpress + 3 * (size_t)i;
    float mass_solid_frac = solid_i * U_RHO_0_SAND / (U_RHO_0_SAND * solid_i + U_RHO_0 * (fluid_i));      /* :314 */
    float mass_fluid_frac = fluid_i * U_RHO_0 / (U_RHO_0_SAND * solid_i + U_RHO_0 * (fluid_i));           /* :315 */
    if (mass_solid_frac > 0.001 && mass_solid_frac < 0.999 && mass_fluid_frac > 0.001 && mass_fluid_frac < 0.999 && !bi && !bj) {   /* :317 */
        const float dk[3] = {dkx, dky, dkz}, vab[3] = {vabx, vaby, vabz};
        for (int c = 0; c < 3; c++) {
            float solidgrad = (solid_j - solid_i) * dk[c];                                                 /* :318-320 */
            float fluidgrad = (fluid_j - fluid_i) * dk[c];                                                 /* :322-324 */
            float solidbrownian = (solidgrad / (solid_i) - (mass_solid_frac * solidgrad / (solid_i) + mass_fluid_frac * fluidgrad / (fluid_i)));   /* :326-328 */
            float fluidbrownian = (fluidgrad / (fluid_i) - (mass_fluid_frac * fluidgrad / (fluid_i) + mass_solid_frac * solidgrad / (solid_i)));   /* :330-332 */
            float solidpressureslip = (solid_i * press_i - solid_j * press_j) * dk[c] - mass_solid_frac * (solid_i * press_i - solid_j * press_j) * dk[c] -
                                      mass_fluid_frac * ((fluid_i) * press_i - (fluid_j) * press_j) * dk[c];                                         /* :334-336 */
            float fluidpressureslip = (fluid_i * press_i - fluid_j * press_j) * dk[c] - mass_solid_frac * (solid_i * press_i - solid_j * press_j) * dk[c] -
                                      mass_fluid_frac * ((fluid_i) * press_i - (fluid_j) * press_j) * dk[c];                                         /* :338-340 */
            /* :342-348 — the second factor is (150.0 / dens) * delpress_c - xvel*dkx*vab_c - yvel*dky*vab_c - zvel*dkz*vab_c, + GRAVITY for z */
            double second = (c == 2 ? P->gravity : 0.0) + (150.0 / dens_i) * dpi[c] - vi[0] * dkx * vab[c] - vi[1] * dky * vab[c] - vi[2] * dkz * vab[c];
            if (c != 2) second = (150.0 / dens_i) * dpi[c] - vi[0] * dkx * vab[c] - vi[1] * dky * vab[c] - vi[2] * dkz * vab[c];
            float solidbody = (solid_i * dens_i - (mass_solid_frac * solid_i * dens_i + mass_fluid_frac * fluid_i * dens_i)) * second;
            float fluidbody = (fluid_i * dens_i - (mass_solid_frac * solid_i * dens_i + mass_fluid_frac * fluid_i * dens_i)) * second;
            a->sdrift[c] += (float)(U_MIXPRESSURE * (solidbody + solidpressureslip) - U_MIXBROWNIAN * solidbrownian);                                  /* :350-352 */
            a->fdrift[c] += (float)(U_MIXPRESSURE * (fluidbody + fluidpressureslip) - U_MIXBROWNIAN * fluidbrownian);                                  /* :354-356 */
        }
    }
    /* :368 */
    float mixfactor = (!bj) * (!bi) * (solid_i > 0.0) * (solid_j > 0.0) * 2 * (solid_i - 0.0) * (solid_j - 0.0) / (solid_i - 0.0 + solid_j - 0.0 + 0.01);
    {
        const float dk[3] = {dkx, dky, dkz}, vab[3] = {vabx, vaby, vabz};
        for (int p = 0; p < 3; p++)
            for (int q = 0; q < 3; q++) a->vel_grad[3 * p + q] += (float)(-mixfactor * vab[q] * dk[p] * 1. / dens_i);                               /* :369-377 */
        const float *st = s->stress_tensor + 9 * (size_t)i;
        for (int p = 0; p < 3; p++) {                                                                                                              /* :379-381 */
            a->stress_accel[p] += (float)(mixfactor * (st[3 * p + 0] * dkx + st[3 * p + 1] * dky + st[3 * p + 2] * dkz) / pow(dens_i, 2) +
                                          (st[3 * p + 0] * dkx + st[3 * p + 1] * dky + st[3 * p + 2] * dkz) / pow(dens_i, 2));
        }
    }
}

/* Pass B of one in-range pair (FluidGPU-unidyn.cu:383-401) with the completed drift velocities (race-free reading, see the header). */
static inline void upair_mixed_b(const fsgo_ustate *s, const upair_acc *acc, int i, int j, float dkx, float dky, float dkz,
                                 float vabx, float vaby, float vabz, upair_acc *a)
{
    const int bi = s->boundary[i] != 0, bj = s->boundary[j] != 0;
    const float solid_i = s->solid[i], fluid_i = s->fluid[i], solid_j = s->solid[j], fluid_j = s->fluid[j];
    const float dens_i = s->dens[i], dens_j = s->dens[j];
    const float *sdi = acc[i].sdrift, *sdj = acc[j].sdrift, *fdi = acc[i].fdrift, *fdj = acc[j].fdrift;
    float ds2 = sdj[0] * dkx + sdj[1] * dky + sdj[2] * dkz;          /* :383-387 (dot_prod, FluidGPU-unidyn.cu:46) */
    float ds = sdi[0] * dkx + sdi[1] * dky + sdi[2] * dkz;
    float df2 = fdj[0] * dkx + fdj[1] * dky + fdj[2] * dkz;
    float df = fdi[0] * dkx + fdi[1] * dky + fdi[2] * dkz;
    for (int c = 0; c < 3; c++)                                       /* :391-398 */
        a->mix[c] += -1 / dens_i / dens_j * (solid_j * dens_j * (solid_j * sdj[c] * ds2 + solid_i * sdi[c] * ds) +
                                              fluid_j * dens_j * (fluid_j * fdj[c] * df2 + fluid_i * fdi[c] * df));
    /* :400-401 */
    a->delsolid += (float)((!bj) * (!bi) * -0.5 / dens_j * (solid_i + solid_j) * (dkx * vabx + dky * vaby + dkz * vabz) +
                           (-(solid_i * sdi[0] + solid_j * sdj[0]) * dkx - (solid_i * sdi[1] + solid_j * sdj[1]) * dky -
                            (solid_i * sdi[2] + solid_j * sdj[2]) * dkz) / dens_j);
    a->delfluid += (float)((!bj) * (!bi) * -0.5 / dens_j * (fluid_i + fluid_j) * (dkx * vabx + dky * vaby + dkz * vabz) +
                           (-(fluid_i * fdi[0] + fluid_j * fdj[0]) * dkx - (fluid_i * fdi[1] + fluid_j * fdj[1]) * dky -
                            (fluid_i * fdi[2] + fluid_j * fdj[2]) * dkz) / dens_j);
}

/* The granular stress update of one particle (FluidGPU-unidyn.cu:410-446), once per step with the completed vel_grad. */
static void ustress_update(const fsgo_ustate *s, const upair_acc *a, int i)
{
    if (!(s->solid[i])) return;                                       /* :411 */
    float *st = s->stress_tensor + 9 * (size_t)i, *sr = s->stress_rate + 9 * (size_t)i;
    const float press = s->press[i];
    float strain[9];
    float tr = 0, tr2 = 0, tr3 = 0, tr4 = 0, tr5 = 0;
    /* strain_rate is complete before the traces that read its transpose are formed (the reference forms tr4 from strain_rate[q][p]
       inside the same loop, i.e. partly from the previous step's values — an artefact the race-free reading drops) */
    for (int p = 0; p < 3; p++)
        for (int q = 0; q < 3; q++) strain[3 * p + q] = 0.5 * (a[i].vel_grad[3 * p + q] + a[i].vel_grad[3 * q + p]);        /* :419 */
    for (int p = 0; p < 3; p++) {
        for (int q = 0; q < 3; q++) {
            tr3 += 0.5 * st[3 * p + q] * st[3 * p + q];                                                                  /* :421 */
            tr5 += strain[3 * p + q] * strain[3 * p + q];                                                               /* :423 */
            tr4 += st[3 * p + q] * strain[3 * q + p];                                                                   /* :424 */
        }
        tr += strain[3 * p + p];                                                                                        /* :426 */
        tr2 += st[3 * p + p];
    }
    (void)tr2;
    for (int p = 0; p < 3; p++) {
        for (int q = 0; q < 3; q++) {
            /* :435-437 */
            if (3 * tan(U_PHI) / (sqrt(9 + 12 * pow(tan(U_PHI), 2))) * press * (press > 0) + U_KC / (sqrt(9 + 12 * pow(tan(U_PHI), 2))) < tr3 && tr3 != 0) {
                st[3 * p + q] *= (3 * tan(U_PHI) / (sqrt(9 + 12 * pow(tan(U_PHI), 2))) * press * (press > 0) + U_KC / (sqrt(9 + 12 * pow(tan(U_PHI), 2)))) / tr3;
            }
            /* :438 */
            sr[3 * p + q] = 3 * U_C1 * (press) * (strain[3 * p + q] - 1. / 3. * tr * (p == q)) +
                            U_C1 * U_C2 * (tr4 + tr * press * (press > 0)) / (pow(press, 2) + 1e8) * st[3 * p + q] - U_C1 * U_C3 * sqrt(tr5) * st[3 * p + q];
        }
    }
}

static inline void upair_mixed_a(const fsgo_params *P, const fsgo_ustate *s, int i, int j, float dkx, float dky, float dkz,
                                 float vabx, float vaby, float vabz, upair_acc *a);
static inline void upair_mixed_b(const fsgo_ustate *s, const upair_acc *acc, int i, int j, float dkx, float dky, float dkz,
                                 float vabx, float vaby, float vabz, upair_acc *a);

/* pair body, FluidGPU-unidyn.cu:258-401 (== :679-821 of mykernel3).  pass 0: the live terms (+ the mixed-phase / granular pass-A
 * terms when the scene is mixed); pass 1 (mixed scenes only): mixture_accel / delsolid / delfluid from the completed drift sums `acc`. */
static inline void upair_body(const fsgo_params *P, const fsgo_ustate *s, int i, int j, upair_acc *a, long long *st, int mixed, int pass,
                              const upair_acc *acc)
{
    const double cutoff = P->h;
    const float *pi = s->pos + 3 * (size_t)i, *pj = s->pos + 3 * (size_t)j;
    float rabx = pi[0] - pj[0], raby = pi[1] - pj[1], rabz = pi[2] - pj[2];
    float ds = sqrt(powf(rabx, 2) + powf(raby, 2) + powf(rabz, 2));   /* cuh:211-213: double sqrt of a float sum, narrowed */
    if (pass == 0) st[0]++;
    if (ds <= (2 * cutoff) && ds > 0) {                                /* cu:287 */
        if (pass == 0) st[1]++;
        float k = fsgo_kernel_h(ds, cutoff);
        const float *vi = s->vel + 3 * (size_t)i, *vj = s->vel + 3 * (size_t)j;
        float vabx = vi[0] - vj[0], vaby = vi[1] - vj[1], vabz = vi[2] - vj[2];
        float dkx = fsgo_kernel_derivative_h(ds, cutoff) * rabx / ds;  /* cu:296-298 */
        float dky = fsgo_kernel_derivative_h(ds, cutoff) * raby / ds;
        float dkz = fsgo_kernel_derivative_h(ds, cutoff) * rabz / ds;
        if (pass == 1) { upair_mixed_b(s, acc, i, j, dkx, dky, dkz, vabx, vaby, vabz, a); return; }
        float d = vabx * rabx + vaby * raby + vabz * rabz;             /* cu:304 */
        float d2 = powf(ds, 2);                                        /* cu:305 */
        int bi = s->boundary[i] != 0, bj = s->boundary[j] != 0;
        float di = s->dens[i], dj = s->dens[j];
        float solid_i = s->solid[i], fluid_i = s->fluid[i];
        /* Particle::mass of the candidate (cu:358-366); cu:307 reads `SPptr[i].mass` with the raw loop index i — the mass of whatever
           record sits at that sorted slot number, an indexing slip: the home particle's own mass is used here */
        const float mass = s->mass ? s->mass[j] : 1.0f, mass_i = s->mass ? s->mass[i] : 1.0f;
        /* cu:307 */
        float sv = (((solid_i * 9 + 1) * P->alpha_fluid) * P->sound *
                    (powf(mass_i, 1) * cutoff * (d / (d2 + 0.01 * powf(cutoff, 2))) +
                     50 * 1.0 / P->sound * powf(cutoff * (d / (d2 + 0.01 * powf(cutoff, 2))), 2)) /
                    ((di + dj) / 2.0)) *
                   (d < 0) * (1 + (!bi) * (bj) * ((1 + 3 * fluid_i * fluid_i) * P->alpha_boundary));
        float pp = s->press[j] / powf(dj, 2) + s->press[i] / powf(di, 2) + sv;   /* cu:310-312 */
        float dpx = pp * dkx, dpy = pp * dky, dpz = pp * dkz;
        a->px += dpx * mass;                                           /* cu:358-360 */
        a->py += dpy * mass;
        a->pz += dpz * mass;
        a->dens += (float)(k * (1 + (float)(!bi) * (float)(bj) * 1.5) * mass);   /* cu:362 */
        a->dx += mass / dj * dkx * !bj * !bi;                          /* cu:364-366 */
        a->dy += mass / dj * dky * !bj * !bi;
        a->dz += mass / dj * dkz * !bj * !bi;
        if (mixed) { upair_mixed_a(P, s, i, j, dkx, dky, dkz, vabx, vaby, vabz, a); return; }     /* delfluid comes from pass B there */
        /* cu:401 with zero drift velocities: (!bj)(!bi) * -0.5/dens_j * (fluid_i+fluid_j) * (dk . vab) + (-0...)/dens_j */
        a->delfluid += (float)((!bj) * (!bi) * -0.5 / dj * (fluid_i + s->fluid[j]) * (dkx * vabx + dky * vaby + dkz * vabz) +
                               (-(fluid_i * 0.0f + s->fluid[j] * 0.0f) * dkx - (fluid_i * 0.0f + s->fluid[j] * 0.0f) * dky -
                                (fluid_i * 0.0f + s->fluid[j] * 0.0f) * dkz) / dj);
    }
}

/* thread -> neighbour particle map of mykernel / mykernel3 (cu:196-241, :606-666) over `nn` bin offsets */
static int ucandidates(const fsgo_params *P, int bidx, const int *nb, int nn, const int *start, const int *end, int n,
                       int *cand, int cand_max, long long *dropped)
{
    const int G = P->grid, numcells = G * G * G;
    int p[27], pidx[27];
    long long all = 0;
    int total = 0;
    for (int t = 0; t < nn; t++) {
        p[t] = 0;
        pidx[t] = 0;
        int c = bidx + nb[t];
        if (c >= 0 && c < numcells && start[c] >= 0 && end[c] >= 0 && start[c] < n && 1 + end[c] - start[c] > 0) {
            p[t] = 1 + end[c] - start[c];
            pidx[t] = t;
            all += p[t];
            total += p[t];
        }
    }
    int count = 0;
    for (int t = 0; t < nn; t++)
        if (p[t] != 0) {
            p[count] = p[t];
            pidx[count] = pidx[t];
            count++;
        }
    for (int t = count; t < nn; t++) p[t] = pidx[t] = 0;
    int nthreads = total;
    if (P->block_threads > 0 && nthreads > P->block_threads) nthreads = P->block_threads;
    if (nthreads > cand_max) nthreads = cand_max;
    int used = 0;
    for (int tidx = 0; tidx < nthreads; tidx++) {
        int sum = 0, jj = 0;
        while (tidx + 1 > sum && jj < nn) {
            sum += p[jj];
            jj++;
        }
        int c = bidx + nb[pidx[jj - 1]];
        int j = -1;
        if (c >= 0 && c < numcells) {
            j = start[c] + sum - (tidx + 1);
            if (!(start[c] >= 0 && j < n && j >= 0)) j = -1;
        }
        cand[used++] = j;
    }
    if (dropped) *dropped += all - used;
    return used;
}

static void permute_f(float *a, const int *perm, int n, int w, float *tmp)
{
    for (int i = 0; i < n; i++)
        for (int c = 0; c < w; c++) tmp[(size_t)i * w + c] = a[(size_t)perm[i] * w + c];
    memcpy(a, tmp, sizeof(float) * (size_t)n * w);
}

static int ustep(const fsgo_params *P, fsgo_ustate *s, fsgo_adapt *ad, int t, int *cells_sorted, int *start_out, int *end_out,
                 int *split_out, float *spts, float *a3, float *b3, long long *stats);

int fsgo_unidyn_step(const fsgo_params *P, fsgo_ustate *s, int t, int *cells_sorted, int *start_out, int *end_out,
                     int *split_out, float *spts, float *a3, float *b3, long long *stats)
{
    return ustep(P, s, NULL, t, cells_sorted, start_out, end_out, split_out, spts, a3, b3, stats);
}

int fsgo_unidyn_step_adapt(const fsgo_params *P, fsgo_ustate *s, fsgo_adapt *ad, int t, int *cells_sorted, int *start_out, int *end_out,
                           int *split_out, float *spts, float *a3, float *b3, long long *stats)
{
    if (!ad || !s->mass) return -2;
    return ustep(P, s, ad, t, cells_sorted, start_out, end_out, split_out, spts, a3, b3, stats);
}

/* ---- particle merging / splitting (FluidGPU-unidyn.cu:260-285, solver-unidyn.cu:495-542): the race-free reading ----
 * The reference's blocks sit inside the pair loop of mykernel: thread (bin, neighbour j) walks the bin's particles ii and, for a pair
 * closer than the merge distance, rewrites BOTH particles while every other thread of the launch is reading them — and tests
 * `diffusion*`, which that same launch is still accumulating.  Reading implemented here (and in fsg_unidyn_adapt.cu):
 *   1. the pair sums of the step are taken over the unmodified state;
 *   2. merge candidates are the pairs the reference tests (:261): 0 < ds <= merge_distance, both masses in (0, 2), neither a boundary
 *      particle, |diffusion|^2 < 20 for both — with the COMPLETED diffusion sums of this step; a particle merges with its nearest
 *      candidate (ties: the smaller Particle::index) and only if that choice is mutual; the particle with the smaller index survives:
 *      mass 2.75, velocity and position the pair's mean (:262-270); the other gets mass 0, boundary = true, position 90.99 (:263-271);
 *   3. a particle with mass > split_mass_min inside the grid, not a boundary particle, with |diffusion|^2 > 35000 or dens < 9400 (:278)
 *      gets mass 1, split = true and y += 0.015 (:279-282);
 *   4. mykernel2 / Particle::update as always (the absorbed particle is a boundary particle far outside the grid: it is parked);
 *   5. the host loop of solver-unidyn.cu:499-531 (commented out there): for the split particles in DESCENDING slot order a child is
 *      appended at (x, y - 0.03, z) of the parent's updated position with the parent's velocity, mass 1, boundary = false (:503-520),
 *      every other field the class default (the reference would reuse whatever dead record sits in that slot), while capacity lasts. */
static float u_diff2(const upair_acc *a) { return powf(a->dx, 2) + powf(a->dy, 2) + powf(a->dz, 2); }

static void uadapt_merge_split(const fsgo_params *P, fsgo_ustate *s, fsgo_adapt *ad, const int *start, const int *end, int nlive,
                               const upair_acc *acc, unsigned char *splitflag)
{
    const int G = P->grid, numcells = G * G * G;
    int *nn = (int *)malloc(sizeof(int) * (size_t)(nlive > 0 ? nlive : 1));
    ad->merged = ad->split = 0;
    if (!nn) return;
    for (int i = 0; i < nlive; i++) {
        nn[i] = -1;
        if (s->boundary[i] || !(s->mass[i] > 0 && s->mass[i] < 2) || !(u_diff2(&acc[i]) < 20)) continue;
        float best = 0;
        const int b = s->cell[i];
        for (int a = -1; a <= 1; a++)
            for (int bb = -1; bb <= 1; bb++)
                for (int c = -1; c <= 1; c++) {
                    const int cb = b + a * G * G + bb * G + c;                    /* cu:130-132 */
                    if (cb < 0 || cb >= numcells || start[cb] < 0 || end[cb] < 0) continue;
                    for (int j = start[cb]; j <= end[cb] && j < nlive; j++) {
                        if (j == i || s->boundary[j] || !(s->mass[j] > 0 && s->mass[j] < 2) || !(u_diff2(&acc[j]) < 20)) continue;
                        const float *pi = s->pos + 3 * (size_t)i, *pj = s->pos + 3 * (size_t)j;
                        float ds = sqrt(powf(pi[0] - pj[0], 2) + powf(pi[1] - pj[1], 2) + powf(pi[2] - pj[2], 2));   /* cuh:211-213 */
                        if (!(ds <= ad->merge_distance && ds > 0)) continue;      /* :261 */
                        if (nn[i] < 0 || ds < best || (ds == best && s->index[j] < s->index[nn[i]])) { nn[i] = j; best = ds; }
                    }
                }
    }
    for (int i = 0; i < nlive; i++) {
        const int j = nn[i];
        if (j < 0 || nn[j] != i || !(s->index[i] < s->index[j])) continue;       /* the survivor does the work */
        float *vi = s->vel + 3 * (size_t)i, *vj = s->vel + 3 * (size_t)j, *xi = s->pos + 3 * (size_t)i, *xj = s->pos + 3 * (size_t)j;
        s->mass[i] = 2.75;                                                         /* :262 */
        s->mass[j] = 0;                                                            /* :263 */
        s->boundary[j] = 1;                                                        /* :265 */
        for (int c = 0; c < 3; c++) vi[c] = (vi[c] + vj[c]) / 2.0;                 /* :266-268 */
        for (int c = 0; c < 3; c++) xi[c] = (xi[c] + xj[c]) / 2.0;                 /* :269-271 */
        xj[0] = xj[1] = xj[2] = 90.99;                                             /* :272 */
        ad->merged++;
    }
    for (int i = 0; i < nlive; i++) {
        splitflag[i] = 0;
        if (s->mass[i] > ad->split_mass_min && s->cell[i] < numcells && !s->boundary[i] &&
            (u_diff2(&acc[i]) > 35000 || (s->dens[i] < 9400))) {                   /* :278 */
            s->mass[i] = 1;                                                        /* :279 */
            splitflag[i] = 1;                                                      /* :281 */
            s->pos[3 * (size_t)i + 1] += 0.015;                                    /* :282 */
            ad->split++;
        }
    }
    free(nn);
}

/* solver-unidyn.cu:499-531, after the update */
static void uadapt_children(const fsgo_params *P, fsgo_ustate *s, fsgo_adapt *ad, int nscan, const unsigned char *splitflag)
{
    const int G = P->grid, numcells = G * G * G;
    ad->added = 0;
    for (int j = nscan - 1; j >= 0; j--) {                                         /* :499 */
        if (!splitflag[j] || s->boundary[j]) continue;                             /* :500 */
        if (s->n >= ad->capacity) continue;                                        /* :512 (room left) */
        const int k = s->n;
        const float x = s->pos[3 * (size_t)j], y = s->pos[3 * (size_t)j + 1] - 0.03, z = s->pos[3 * (size_t)j + 2];   /* :501-503 */
        s->pos[3 * (size_t)k] = x; s->pos[3 * (size_t)k + 1] = y; s->pos[3 * (size_t)k + 2] = z;
        for (int c = 0; c < 3; c++) s->vel[3 * (size_t)k + c] = s->vel[3 * (size_t)j + c];                          /* :504-506, :516-518 */
        s->mass[k] = 1;                                                            /* :519 */
        s->boundary[k] = 0;                                                        /* :520 */
        /* class defaults for the rest (FluidGPU-unidyn.cuh:132-187) */
        s->acc[3 * (size_t)k] = s->acc[3 * (size_t)k + 1] = 0; s->acc[3 * (size_t)k + 2] = (float)P->gravity;
        s->dens[k] = 9550; s->press[k] = 0;
        for (int c = 0; c < 3; c++) s->delpress[3 * (size_t)k + c] = s->newdelpress[3 * (size_t)k + c] = 0;
        s->newdens[k] = 0;
        s->solid[k] = 0; s->fluid[k] = 1;
        s->subindex[k] = 0;
        s->index[k] = ad->next_index++;
        {                                                                          /* :522 */
            float fx = x - P->origin, fy = y - P->origin, fz = z - P->origin;
            double qx = fx / P->cellsize, qy = fy / P->cellsize, qz = fz / P->cellsize;
            long long l = (long long)(int)qx * G * G + (long long)(int)qy * G + (int)qz;
            s->cell[k] = (!(fabs(qx) < 1e6 && fabs(qy) < 1e6 && fabs(qz) < 1e6) || l < 0 || l >= numcells) ? numcells : (int)l;
        }
        s->n++;
        ad->added++;
    }
}

static int ustep(const fsgo_params *P, fsgo_ustate *s, fsgo_adapt *ad, int t, int *cells_sorted, int *start_out, int *end_out,
                 int *split_out, float *spts, float *a3, float *b3, long long *stats)
{
    (void)t;
    const int n = s->n, G = P->grid, numcells = G * G * G;
    int rc = -1;
    /* scope: a mixed-phase / granular scene needs the stress arrays */
    int mixed = 0;
    for (int i = 0; i < n; i++)
        if (!s->boundary[i] && s->solid[i] != 0.0f) mixed = 1;
    if (mixed && (!s->stress_tensor || !s->stress_rate)) return -2;
    if (mixed && ad) return -2;                     /* merging / splitting: pure-fluid scenes */
    unsigned char *splitflag = ad ? (unsigned char *)calloc((size_t)(n > 0 ? n : 1), 1) : NULL;
    if (ad && !splitflag) return -1;

    int *perm = (int *)malloc(sizeof(int) * (size_t)(n > 0 ? n : 1));
    float *tmp = (float *)malloc(sizeof(float) * 3 * (size_t)(n > 0 ? n : 1));
    int *start = (int *)malloc(sizeof(int) * (size_t)numcells);
    int *end = (int *)malloc(sizeof(int) * (size_t)numcells);
    int *split = (int *)malloc(sizeof(int) * (size_t)numcells);
    int *cnt = (int *)calloc((size_t)numcells + 2, sizeof(int));
    upair_acc *acc = (upair_acc *)calloc((size_t)(n > 0 ? n : 1), sizeof(upair_acc));
    upair_acc *accb = (upair_acc *)calloc((size_t)(n > 0 ? n : 1), sizeof(upair_acc));
    int *occ = (int *)malloc(sizeof(int) * (size_t)(n > 0 ? n : 1));
    long long st[4] = {0, 0, 0, 0};
    if (!perm || !tmp || !start || !end || !split || !cnt || !acc || !accb || !occ) goto done;

    /* ---- solver-unidyn.cu:331 stable sort by cell id (out-of-grid ids last) ---- */
    for (int i = 0; i < n; i++) {
        int k = s->cell[i];
        if (k < 0 || k >= numcells) k = numcells;
        cnt[k + 1]++;
    }
    for (int c = 0; c <= numcells; c++) cnt[c + 1] += cnt[c];
    for (int i = 0; i < n; i++) {
        int k = s->cell[i];
        if (k < 0 || k >= numcells) k = numcells;
        perm[cnt[k]++] = i;
    }
    permute_f(s->pos, perm, n, 3, tmp);
    permute_f(s->vel, perm, n, 3, tmp);
    permute_f(s->acc, perm, n, 3, tmp);
    permute_f(s->dens, perm, n, 1, tmp);
    permute_f(s->press, perm, n, 1, tmp);
    permute_f(s->delpress, perm, n, 3, tmp);
    permute_f(s->newdens, perm, n, 1, tmp);
    permute_f(s->newdelpress, perm, n, 3, tmp);
    permute_f(s->solid, perm, n, 1, tmp);
    permute_f(s->fluid, perm, n, 1, tmp);
    if (s->mass) permute_f(s->mass, perm, n, 1, tmp);
    if (s->stress_tensor && s->stress_rate) {
        float *t9 = (float *)malloc(sizeof(float) * 9 * (size_t)(n > 0 ? n : 1));
        if (!t9) goto done;
        permute_f(s->stress_tensor, perm, n, 9, t9);
        permute_f(s->stress_rate, perm, n, 9, t9);
        free(t9);
    }
    {
        int *ti = (int *)tmp;
        for (int i = 0; i < n; i++) ti[i] = s->index[perm[i]];
        memcpy(s->index, ti, sizeof(int) * (size_t)n);
        for (int i = 0; i < n; i++) ti[i] = s->cell[perm[i]];
        memcpy(s->cell, ti, sizeof(int) * (size_t)n);
        for (int i = 0; i < n; i++) ti[i] = s->subindex[perm[i]];
        memcpy(s->subindex, ti, sizeof(int) * (size_t)n);
        unsigned char *tb = (unsigned char *)tmp;
        for (int i = 0; i < n; i++) tb[i] = s->boundary[perm[i]];
        memcpy(s->boundary, tb, (size_t)n);
    }
    /* count_after_merge, FluidGPU-unidyn.cu:554-562: particles with cell >= NUMCELLS drop out */
    int nlive = n;
    while (nlive > 0 && (s->cell[nlive - 1] < 0 || s->cell[nlive - 1] >= numcells)) nlive--;

    /* ---- findneighbours, FluidGPU-unidyn.cu:104-122 ---- */
    for (int c = 0; c < numcells; c++) start[c] = end[c] = split[c] = -1;
    int nocc = 0;
    for (int i = 0; i < nlive; i++) {
        if (i == 0 || s->cell[i] != s->cell[i - 1]) {
            start[s->cell[i]] = i;
            occ[nocc++] = s->cell[i];
        }
        if (i == nlive - 1 || s->cell[i] != s->cell[i + 1]) end[s->cell[i]] = i;
    }
    st[3] = nocc;
    /* ---- split marking + octant subindex, FluidGPU-unidyn.cu:181-192 ---- */
    for (int o = 0; o < nocc; o++) {
        int b = occ[o];
        int pop = 1 + end[b] - start[b];
        if (pop > 6) {
            split[b] = b;
            for (int i = start[b]; i <= end[b]; i++) {
                const float *x = s->pos + 3 * (size_t)i;
                float fx = x[0] - P->origin, fy = x[1] - P->origin, fz = x[2] - P->origin;
                /* int((x-XMIN)/CELLSIZE) == int((x-XMIN+CELLSIZE/2)/CELLSIZE): (x - XMIN) float, + 0.06 double */
                int sx = (int)(fx / P->cellsize) == (int)((fx + P->cellsize / 2) / P->cellsize);
                int sy = (int)(fy / P->cellsize) == (int)((fy + P->cellsize / 2) / P->cellsize);
                int sz = (int)(fz / P->cellsize) == (int)((fz + P->cellsize / 2) / P->cellsize);
                s->subindex[i] = 1 - sx + 2 - 2 * sy + 4 * sz;
            }
        }
    }
    if (cells_sorted) memcpy(cells_sorted, s->cell, sizeof(int) * (size_t)n);
    if (start_out) memcpy(start_out, start, sizeof(int) * (size_t)numcells);
    if (end_out) memcpy(end_out, end, sizeof(int) * (size_t)numcells);
    if (split_out) memcpy(split_out, split, sizeof(int) * (size_t)numcells);

    /* ---- mykernel (coarse bins) + mykernel3 (split bins, per octant); mixed scenes: pass A, then pass B with the completed drift sums ---- */
    for (int pass = 0; pass < (mixed ? 2 : 1); pass++) {
        int nthreads = P->threads;
        (void)nthreads;
        long long t0 = 0, t1 = 0, t2 = 0;
#ifdef _OPENMP
        if (nthreads <= 0) nthreads = omp_get_max_threads();
#pragma omp parallel num_threads(nthreads) reduction(+ : t0, t1, t2)
#endif
        {
            int cand_max = 27 * 1024;
            int *cand = (int *)malloc(sizeof(int) * (size_t)cand_max);
            int nb27[27], k = 0;
            for (int a = -1; a <= 1; a++)
                for (int b = -1; b <= 1; b++)
                    for (int c = -1; c <= 1; c++) nb27[k++] = a * G * G + b * G + c;     /* cu:130-132 */
#ifdef _OPENMP
#pragma omp for schedule(dynamic, 16)
#endif
            for (int o = 0; o < nocc; o++) {
                int bidx = occ[o];
                long long lst[2] = {0, 0}, drop = 0;
                if (split[bidx] == -1) {
                    int nc = ucandidates(P, bidx, nb27, 27, start, end, nlive, cand, cand_max, &drop);
                    for (int i = start[bidx]; i <= end[bidx]; i++) {
                        upair_acc a;
                        memset(&a, 0, sizeof a);
                        for (int q = 0; q < nc; q++)
                            if (cand[q] >= 0) upair_body(P, s, i, cand[q], &a, lst, mixed, pass, acc);
                        if (pass == 0) acc[i] = a;
                        else { memcpy(accb[i].mix, a.mix, sizeof a.mix); accb[i].delsolid = a.delsolid; accb[i].delfluid = a.delfluid; }
                    }
                } else {
                    for (int oct = 0; oct < 8; oct++) {
                        /* cu:579-583: pow(-1,1+dirx)*G*G etc.; z polarity is opposite (pow(-1,dirz)) */
                        int dirx = oct & 1, diry = (oct & 2) >> 1, dirz = (oct & 4) >> 2;
                        int ox = (dirx ? 1 : -1) * G * G, oy = (diry ? 1 : -1) * G, oz = dirz ? -1 : 1;
                        int nb8[8] = {0, ox, oy, oz, ox + oy, ox + oz, oy + oz, ox + oy + oz};
                        int any = 0;
                        for (int i = start[bidx]; i <= end[bidx]; i++) any |= (s->subindex[i] == oct);
                        if (!any) continue;
                        int nc = ucandidates(P, bidx, nb8, 8, start, end, nlive, cand, cand_max, &drop);
                        for (int i = start[bidx]; i <= end[bidx]; i++) {
                            if (s->subindex[i] != oct) continue;
                            upair_acc a;
                            memset(&a, 0, sizeof a);
                            for (int q = 0; q < nc; q++)
                                if (cand[q] >= 0) upair_body(P, s, i, cand[q], &a, lst, mixed, pass, acc);
                            if (pass == 0) acc[i] = a;
                            else { memcpy(accb[i].mix, a.mix, sizeof a.mix); accb[i].delsolid = a.delsolid; accb[i].delfluid = a.delfluid; }
                        }
                    }
                }
                t0 += lst[0];
                t1 += lst[1];
                t2 += drop;
            }
            free(cand);
        }
        if (pass == 0) {
            st[0] = t0;
            st[1] = t1;
            st[2] = t2;
        }
    }
    if (mixed) {
        for (int i = 0; i < nlive; i++) {           /* pass B results join the sums; then the stress update (cu:410-446), once per particle */
            memcpy(acc[i].mix, accb[i].mix, sizeof accb[i].mix);
            acc[i].delsolid = accb[i].delsolid;
            acc[i].delfluid = accb[i].delfluid;
        }
        for (int i = 0; i < nlive; i++) ustress_update(s, acc, i);
    }

    /* ---- particle merging / splitting, between the pair sums and the update (see above) ---- */
    if (ad) uadapt_merge_split(P, s, ad, start, end, nlive, acc, splitflag);

    /* ---- mykernel2 (cu:451-497) + Particle::update(t) (cuh:296-423) + cell_calc (cu:544-551) ---- */
    {
        const double DT = P->dt;
        for (int i = 0; i < nlive; i++) {
            float *x = s->pos + 3 * (size_t)i, *v = s->vel + 3 * (size_t)i, *a = s->acc + 3 * (size_t)i;
            float newdens = s->newdens[i] + acc[i].dens;
            float ndx = s->newdelpress[3 * (size_t)i + 0] + acc[i].px, ndy = s->newdelpress[3 * (size_t)i + 1] + acc[i].py,
                  ndz = s->newdelpress[3 * (size_t)i + 2] + acc[i].pz;
            float diffx = acc[i].dx, diffy = acc[i].dy, diffz = acc[i].dz;
            float delfluid = acc[i].delfluid, delsolid = mixed ? acc[i].delsolid : 0.0f;
            const float sa0 = mixed ? acc[i].stress_accel[0] : 0.0f, sa1 = mixed ? acc[i].stress_accel[1] : 0.0f, sa2 = mixed ? acc[i].stress_accel[2] : 0.0f;
            const float ma0 = mixed ? acc[i].mix[0] : 0.0f, ma1 = mixed ? acc[i].mix[1] : 0.0f, ma2 = mixed ? acc[i].mix[2] : 0.0f;
            if (spts) { spts[3 * (size_t)i] = x[0]; spts[3 * (size_t)i + 1] = x[1]; spts[3 * (size_t)i + 2] = x[2]; }   /* cu:462-464 */
            if (a3) a3[i] = s->mass ? s->mass[i] : 1.0f;                                                                /* mass, :465 */
            if (b3) b3[i] = powf(diffx, 2) + powf(diffy, 2) + powf(diffz, 2);                                          /* :466 */
            int bnd = s->boundary[i] != 0;
            float solid = s->solid[i], fluid = s->fluid[i];
            s->dens[i] = u_set_dens(newdens, bnd, P->h);                 /* cuh:300 */
            s->press[i] = u_pressure(s->dens[i], solid, P->sound);       /* cuh:301 */
            float *dp = s->delpress + 3 * (size_t)i;
            dp[0] = ndx; dp[1] = ndy; dp[2] = ndz;                      /* cuh:302 */
            if (s->stress_tensor && s->stress_rate)                      /* cuh:304-308 */
                for (int pq = 0; pq < 9; pq++) s->stress_tensor[9 * (size_t)i + pq] = DT * s->stress_rate[9 * (size_t)i + pq];
            if (!bnd) {
                volatile float friction = fabsf(diffx) + fabsf(diffy) + fabsf(diffz);   /* cuh:311 */
                solid += DT * delsolid;                                   /* :312-313 */
                solid *= (solid >= 0.0);
                if (fluid + delfluid < 0.2) delfluid = 0;                 /* :315 */
                fluid += DT * delfluid;                                   /* :316-317 */
                fluid *= (fluid >= 0);
                fluid *= 1 / (fluid + solid);                             /* :319-320 */
                solid *= 1 / (fluid + solid);
                /* leapfrog :328-330 (DIFF == 0: + 0*diffusion) */
                x[0] = x[0] + DT * v[0] + 0.5 * DT * DT * a[0] + 0 * diffx;
                x[1] = x[1] + DT * v[1] + 0.5 * DT * DT * a[1] + 0 * diffy;
                x[2] = x[2] + DT * v[2] + 0.5 * DT * DT * a[2] + 0 * diffz;
                if (x[2] < -0.89) { v[0] = 0; v[1] = 0; }                 /* :332-341 */
                /* :351-353 — the y and z lines test the NEW xvel and xacc (sic), with their own stress_accel / mixture_accel component */
                v[0] = (v[0] + 0.5 * DT * a[0] + DT * (sa0) + 5 * DT * DT * (ma0)) -
                       ((v[0] + DT * a[0] + DT * (sa0) + DT * DT * (ma0)) > 0) * friction * 0.0000002 * solid +
                       ((v[0] + DT * a[0] + DT * (sa0) + DT * DT * (ma0)) < 0) * friction * 0.0000002 * solid;
                v[1] = (v[1] + 0.5 * DT * a[1] + DT * (sa1) + 5 * DT * DT * (ma1)) -
                       ((v[0] + DT * a[0] + DT * (sa1) + DT * DT * (ma1)) > 0) * friction * 0.0000002 * solid +
                       ((v[0] + DT * a[0] + DT * (sa1) + DT * DT * (ma1)) < 0) * friction * 0.0000002 * solid;
                v[2] = (v[2] + 0.5 * DT * a[2] + DT * (sa2) + 5 * DT * DT * (ma2)) -
                       ((v[0] + DT * a[0] + DT * (sa2) + DT * DT * (ma2)) > 0) * friction * 0.0000002 * solid +
                       ((v[0] + DT * a[0] + DT * (sa2) + DT * DT * (ma2)) < 0) * friction * 0.0000002 * solid;
                /* :357-359 */
                a[0] = -((220.0 - 70.0 * solid) / s->dens[i]) * dp[0];
                a[1] = -((220.0 - 70.0 * solid) / s->dens[i]) * dp[1];
                a[2] = P->gravity + ((-220.0 + 70.0 * solid) / s->dens[i]) * dp[2];
                /* :390-392 */
                v[0] += 0.5 * a[0] * DT;
                v[1] += 0.5 * a[1] * DT;
                v[2] += 0.5 * a[2] * DT;
                /* walls :404-413 */
                if (fabsf(x[2]) > 0.98) { x[2] = 0.97 / x[2]; v[2] = 0; }
                if (fabsf(x[1]) > 0.98) v[1] = -v[1];
                if (fabsf(x[0]) > 0.98) v[0] = -v[0];
                s->solid[i] = solid;
                s->fluid[i] = fluid;
            }
            /* cell_calc, FluidGPU-unidyn.cu:547 */
            {
                float fx = x[0] - P->origin, fy = x[1] - P->origin, fz = x[2] - P->origin;
                double qx = fx / P->cellsize, qy = fy / P->cellsize, qz = fz / P->cellsize;
                int cid;
                if (!(fabs(qx) < 1e6 && fabs(qy) < 1e6 && fabs(qz) < 1e6)) cid = numcells;
                else {
                    long long l = (long long)(int)qx * G * G + (long long)(int)qy * G + (int)qz;
                    cid = (l < 0 || l >= numcells) ? numcells : (int)l;
                }
                s->cell[i] = cid;
            }
            s->newdens[i] = 0;                                            /* cu:475-478 */
            s->newdelpress[3 * (size_t)i] = s->newdelpress[3 * (size_t)i + 1] = s->newdelpress[3 * (size_t)i + 2] = 0;
        }
        if (spts || a3 || b3)
            for (int i = nlive; i < n; i++) {
                if (spts) { spts[3 * (size_t)i] = s->pos[3 * (size_t)i]; spts[3 * (size_t)i + 1] = s->pos[3 * (size_t)i + 1]; spts[3 * (size_t)i + 2] = s->pos[3 * (size_t)i + 2]; }
                if (a3) a3[i] = s->mass ? s->mass[i] : 1.0f;
                if (b3) b3[i] = 0.0f;
            }
    }
    if (ad) uadapt_children(P, s, ad, nlive, splitflag);
    if (stats) memcpy(stats, st, sizeof(st));
    rc = 0;
done:
    free(splitflag);
    free(perm); free(tmp); free(start); free(end); free(split); free(cnt); free(acc); free(accb); free(occ);
    return rc;
}
