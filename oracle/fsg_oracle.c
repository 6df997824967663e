/*
 * fsg_oracle.c — CPU restatement of FluidSolverGPU's base per-timestep particle update.
 *
 * TEST INFRASTRUCTURE ONLY (see fsg_oracle.h).  Plain C, gcc -O2 -ffp-contract=off, OpenMP optional.
 * Every function cites the reference file:line it follows.  Expression types follow the
 * reference's C++ promotion rules: unsuffixed literals are double, `cutoff`, `DT`, `CELLSIZE`,
 * `SOUND`, `ALPHA_*`, `BDENSFACTOR` are double macros, `RHO_0`, `XMIN`, `GRIDSIZE` are int macros.
 */
#include "fsg_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ---- constants: FluidGPU.cuh:1-31 ---- */
void fsgo_params_base(fsgo_params *p)
{
    p->grid = 40;              /* GRIDSIZE  :8  */
    p->origin = -1.0f;         /* XMIN      :1  */
    p->cellsize = 0.05;        /* CELLSIZE  :7  */
    p->h = 0.06;               /* cutoff    :30 */
    p->dt = 0.0005;            /* DT        :31 */
    p->alpha_fluid = -0.01e2;  /* :16 */
    p->alpha_boundary = 2000e-1; /* :17 */
    p->sound = 1450.0;         /* :11 */
    p->gravity = -9.8;         /* :10 */
    p->block_threads = 64;     /* solver.cu:187 */
    p->bin_cap = 64;           /* FluidGPU.cu:174 */
    p->threads = 0;
}

/* ---- smoothing kernels: FluidGPU.cu:11-43 ---- */
float fsgo_kernel_h(float r, double h)
{
    if (r >= 0 && r <= h) {                     /* double compare, FluidGPU.cu:12 */
        return 1. / 3.14159 / (powf(h, 3)) * (1 - 3. / 2. * powf((r / h), 2) + 3. / 4. * powf((r / h), 3));
    } else if (r > h && r < (2 * h)) {          /* FluidGPU.cu:15 */
        return 1. / 3.14159 / (powf(h, 3)) * 1 / 4. * powf(2 - (r / h), 3);
    }
    return 0;
}

static float kernel_test_h(float r, double h)   /* FluidGPU.cu:23-33 */
{
    if (r >= 0 && r <= h) {
        return 1. / 3.14159 / (powf(h, 4)) * (1 - 3. * powf((r / h), 1) + 9. / 4. * powf((r / h), 2));
    } else if (r > h && r < (2 * h)) {
        return -1. / 3.14159 / (powf(h, 4)) * 1 / 2. * powf(2 - (r / h), 2);
    }
    return 0;
}

float fsgo_kernel_derivative_h(float r, double h)  /* FluidGPU.cu:35-43 */
{
    if (r < h) {
        return -45.0 / 3.14159 / powf(h, 6) * powf((h - r), 2);
    }
    return 0;
}

float fsgo_kernel(float r) { return fsgo_kernel_h(r, 0.06); }
float fsgo_kernel_test(float r) { return kernel_test_h(r, 0.06); }
float fsgo_kernel_derivative(float r) { return fsgo_kernel_derivative_h(r, 0.06); }

/* FluidGPU.cuh:165-167: dens = (x + kernel(0)) / 23.0 *(1 + float(boundary)*BDENSFACTOR) + 9250 */
static float set_dens_h(float x, int boundary, double h)
{
    return (x + fsgo_kernel_h(0, h)) / 23.0 * (1 + (float)(boundary != 0) * 1.5) + 9250;
}
float fsgo_set_dens(float newdens, int boundary) { return set_dens_h(newdens, boundary, 0.06); }

/* FluidGPU.cuh:256-257: press = 1000 * powf(SOUND, 0)*RHO_0 / 7.0*(powf(dens / RHO_0, 7) - 1) */
static float pressure_p(float dens, double sound)
{
    return 1000 * powf(sound, 0) * 9550 / 7.0 * (powf(dens / 9550, 7) - 1);
}
float fsgo_pressure(float dens) { return pressure_p(dens, 1450.0); }

/* FluidGPU.cu:419 / solver.cu:119: int((x - XMIN)/CELLSIZE)*G*G + int((y - YMIN)/CELLSIZE)*G + int((z - ZMIN)/CELLSIZE).
 * (x - XMIN) is a float subtraction of an int; the division by the double CELLSIZE promotes. */
int fsgo_cell_id(const fsgo_params *p, float x, float y, float z)
{
    int g = p->grid;
    float fx = x - p->origin, fy = y - p->origin, fz = z - p->origin;
    return (int)(fx / p->cellsize) * g * g + (int)(fy / p->cellsize) * g + (int)(fz / p->cellsize);
}

/* ---- helpers ---- */
static void permute_f(float *a, const int *perm, int n, int w, float *tmp)
{
    for (int i = 0; i < n; i++)
        for (int c = 0; c < w; c++) tmp[(size_t)i * w + c] = a[(size_t)perm[i] * w + c];
    memcpy(a, tmp, sizeof(float) * (size_t)n * w);
}

/* solver.cu:181 — thrust::sort_by_key is a stable LSD radix sort; any stable sort by key gives the
 * same permutation.  Counting sort over the key range; keys outside [0,numcells) are parked last. */
static int stable_sort_perm(const int *key, int n, int numcells, int *perm)
{
    int *cnt = (int *)calloc((size_t)numcells + 2, sizeof(int));
    if (!cnt) return -1;
    for (int i = 0; i < n; i++) {
        int k = key[i];
        if (k < 0 || k >= numcells) k = numcells;
        cnt[k + 1]++;
    }
    for (int c = 0; c <= numcells; c++) cnt[c + 1] += cnt[c];
    for (int i = 0; i < n; i++) {
        int k = key[i];
        if (k < 0 || k >= numcells) k = numcells;
        perm[cnt[k]++] = i;
    }
    free(cnt);
    return 0;
}

/* One in-range test + pair body, FluidGPU.cu:235-279, accumulating into locals of particle i. */
typedef struct { float dens, px, py, pz; } pair_acc;

static inline void pair_body(const fsgo_params *P, const fsgo_state *s, int i, int j, pair_acc *a, long long *st)
{
    const double cutoff = P->h;
    const float *pi = s->pos + 3 * (size_t)i, *pj = s->pos + 3 * (size_t)j;
    /* Particle::distance FluidGPU.cuh:193-195 (float arithmetic, powf(.,2)) */
    float rabx = pi[0] - pj[0], raby = pi[1] - pj[1], rabz = pi[2] - pj[2];
    float ds = sqrtf(powf(rabx, 2) + powf(raby, 2) + powf(rabz, 2));
    st[0]++;
    if (ds <= (2 * cutoff) && ds > 0) {                               /* :236 (double compare) */
        st[1]++;
        float k = fsgo_kernel_h(ds, cutoff);                          /* :238 */
        const float *vi = s->vel + 3 * (size_t)i, *vj = s->vel + 3 * (size_t)j;
        float vabx = vi[0] - vj[0], vaby = vi[1] - vj[1], vabz = vi[2] - vj[2];   /* :242-244 */
        float dkx = fsgo_kernel_derivative_h(ds, cutoff) * rabx / ds; /* :245-247 */
        float dky = fsgo_kernel_derivative_h(ds, cutoff) * raby / ds;
        float dkz = fsgo_kernel_derivative_h(ds, cutoff) * rabz / ds;
        float d = vabx * rabx + vaby * raby + vabz * rabz;            /* :253, dot_prod :46 */
        float d2 = powf(ds, 2);                                       /* :254 */
        int bi = s->boundary[i] != 0, bj = s->boundary[j] != 0;
        float di = s->dens[i], dj = s->dens[j];
        /* :255 — double expression narrowed to float */
        float sv = (P->alpha_fluid * P->sound *
                    (cutoff * (d / (d2 + 0.01 * powf(cutoff, 2))) +
                     50 * 1.0 / P->sound * powf(cutoff * (d / (d2 + 0.01 * powf(cutoff, 2))), 2)) /
                    ((di + dj) / 2.0)) *
                   (d < 0) * (1 + (!bi) * (bj) * P->alpha_boundary);
        float pp = s->press[j] / powf(dj, 2) + s->press[i] / powf(di, 2) + sv;   /* :258-260 */
        float dpx = pp * dkx, dpy = pp * dky, dpz = pp * dkz;
        a->dens += (float)(k * (1 + (float)(!bi) * (float)(bj) * 1.5));          /* :276 */
        a->px += dpx;                                                            /* :277-279 */
        a->py += dpy;
        a->pz += dpz;
    }
}

/* Phases 1-4 of mykernel (FluidGPU.cu:150-231): which sorted slots j the block's threads visit.
 * Fills cand[] in thread order, returns how many threads do pair work.  *dropped counts neighbour
 * particles that exist in the 27 bins but get no thread. */
static int candidates(const fsgo_params *P, int bidx, const int *start, const int *end, int n,
                      int *cand, int cand_max, long long *dropped)
{
    const int G = P->grid, numcells = G * G * G;
    int nb[27], p[27], pidx[27];
    int t = 0;
    for (int a = -1; a <= 1; a++)               /* :124-126: a*G*G + b*G + c, c fastest */
        for (int b = -1; b <= 1; b++)
            for (int c = -1; c <= 1; c++) nb[t++] = a * G * G + b * G + c;
    long long all = 0;
    for (t = 0; t < 27; t++) {                  /* :150-159 */
        p[t] = 0;
        pidx[t] = 0;
        int c = bidx + nb[t];
        if (c >= 0 && c < numcells && start[c] >= 0 && end[c] >= 0 && start[c] < n && 1 + end[c] - start[c] > 0) {
            p[t] = 1 + end[c] - start[c];
            pidx[t] = t;
            all += p[t];
        }
    }
    int total = 0;                              /* :170-177 */
    for (t = 0; t < 27; t++)
        if ((P->bin_cap <= 0 || p[t] < P->bin_cap) && p[t] > 0) total += p[t];
    int count = 0;                              /* :186-198 compaction */
    for (t = 0; t < 27; t++)
        if (p[t] != 0) {
            p[count] = p[t];
            pidx[count] = pidx[t];
            count++;
        }
    for (t = count; t < 27; t++) p[t] = pidx[t] = 0;
    int nthreads = total;                       /* tidx < total, and tidx < blockDim.x */
    if (P->block_threads > 0 && nthreads > P->block_threads) nthreads = P->block_threads;
    if (nthreads > cand_max) nthreads = cand_max;
    int used = 0;
    for (int tidx = 0; tidx < nthreads; tidx++) {   /* :204-231 */
        int sum = 0, jj = 0;
        while (tidx + 1 > sum && jj < 27) {
            sum += p[jj];
            jj++;
        }
        int c = bidx + nb[pidx[jj - 1]];
        int j = -1;
        if (c >= 0 && c < numcells) {
            j = start[c] + sum - (tidx + 1);        /* :228 — reversed order inside each bin */
            if (!(start[c] >= 0 && j < n && j >= 0)) j = -1;
        }
        cand[used++] = j;
    }
    if (dropped) *dropped += all - used;
    return used;
}

int fsgo_base_step(const fsgo_params *P, fsgo_state *s, int *cells_sorted, int *start_out, int *end_out,
                   float *spts, float *a3, float *b3, long long *stats)
{
    const int n = s->n, G = P->grid, numcells = G * G * G;
    int rc = -1;
    int *perm = (int *)malloc(sizeof(int) * (size_t)(n > 0 ? n : 1));
    float *tmp = (float *)malloc(sizeof(float) * 3 * (size_t)(n > 0 ? n : 1));
    int *start = (int *)malloc(sizeof(int) * (size_t)numcells);
    int *end = (int *)malloc(sizeof(int) * (size_t)numcells);
    pair_acc *acc = (pair_acc *)calloc((size_t)(n > 0 ? n : 1), sizeof(pair_acc));
    int *occ = (int *)malloc(sizeof(int) * (size_t)(n > 0 ? n : 1));
    long long st[4] = {0, 0, 0, 0};
    if (!perm || !tmp || !start || !end || !acc || !occ) goto done;

    /* ---- solver.cu:181 stable sort of (cells, particles) ---- */
    if (stable_sort_perm(s->cell, n, numcells, perm)) goto done;
    permute_f(s->pos, perm, n, 3, tmp);
    permute_f(s->vel, perm, n, 3, tmp);
    permute_f(s->acc, perm, n, 3, tmp);
    permute_f(s->dens, perm, n, 1, tmp);
    permute_f(s->press, perm, n, 1, tmp);
    permute_f(s->delpress, perm, n, 3, tmp);
    permute_f(s->newdens, perm, n, 1, tmp);
    permute_f(s->newdelpress, perm, n, 3, tmp);
    {
        int *ti = (int *)tmp;
        for (int i = 0; i < n; i++) ti[i] = s->index[perm[i]];
        memcpy(s->index, ti, sizeof(int) * (size_t)n);
        for (int i = 0; i < n; i++) ti[i] = s->cell[perm[i]];
        memcpy(s->cell, ti, sizeof(int) * (size_t)n);
        unsigned char *tb = (unsigned char *)tmp;
        for (int i = 0; i < n; i++) tb[i] = s->boundary[perm[i]];
        memcpy(s->boundary, tb, (size_t)n);
    }
    /* live = particles still inside the bin grid (the reference would write out of bounds otherwise) */
    int nlive = n;
    while (nlive > 0 && (s->cell[nlive - 1] < 0 || s->cell[nlive - 1] >= numcells)) nlive--;

    /* ---- findneighbours FluidGPU.cu:106-117 ---- */
    for (int c = 0; c < numcells; c++) start[c] = end[c] = -1;
    int nocc = 0;
    for (int i = 0; i < nlive; i++) {
        if (i == 0 || s->cell[i] != s->cell[i - 1]) {
            start[s->cell[i]] = i;
            occ[nocc++] = s->cell[i];
        }
        if (i == nlive - 1 || s->cell[i] != s->cell[i + 1]) end[s->cell[i]] = i;
    }
    st[3] = nocc;
    if (cells_sorted) memcpy(cells_sorted, s->cell, sizeof(int) * (size_t)n);
    if (start_out) memcpy(start_out, start, sizeof(int) * (size_t)numcells);
    if (end_out) memcpy(end_out, end, sizeof(int) * (size_t)numcells);

    /* ---- mykernel FluidGPU.cu:119-285: one block per occupied bin ---- */
    {
        int nthreads = P->threads;
        (void)nthreads;
        long long t0 = 0, t1 = 0, t2 = 0;
#ifdef _OPENMP
        if (nthreads <= 0) nthreads = omp_get_max_threads();
#pragma omp parallel num_threads(nthreads) reduction(+ : t0, t1, t2)
#endif
        {
            int cand_max = 27 * 64;
            int *cand = (int *)malloc(sizeof(int) * (size_t)cand_max);
#ifdef _OPENMP
#pragma omp for schedule(dynamic, 64)
#endif
            for (int o = 0; o < nocc; o++) {
                int bidx = occ[o];
                /* grow the candidate buffer for uncapped runs */
                int need = 0;
                for (int a = -1; a <= 1; a++)
                    for (int b = -1; b <= 1; b++)
                        for (int c = -1; c <= 1; c++) {
                            int cc = bidx + a * G * G + b * G + c;
                            if (cc >= 0 && cc < numcells && start[cc] >= 0) need += 1 + end[cc] - start[cc];
                        }
                if (need > cand_max) {
                    free(cand);
                    cand_max = need * 2;
                    cand = (int *)malloc(sizeof(int) * (size_t)cand_max);
                }
                long long lst[2] = {0, 0}, drop = 0;
                int nc = candidates(P, bidx, start, end, nlive, cand, cand_max, &drop);
                for (int i = start[bidx]; i <= end[bidx]; i++) {      /* :234 */
                    pair_acc a = {0, 0, 0, 0};
                    for (int t = 0; t < nc; t++)
                        if (cand[t] >= 0) pair_body(P, s, i, cand[t], &a, lst);
                    acc[i] = a;
                }
                t0 += lst[0];
                t1 += lst[1];
                t2 += drop;
            }
            free(cand);
        }
        st[0] = t0;
        st[1] = t1;
        st[2] = t2;
    }
    /* atomicAdd targets: newdens / newdelpress (:276-279) */
    for (int i = 0; i < nlive; i++) {
        s->newdens[i] += acc[i].dens;
        s->newdelpress[3 * (size_t)i + 0] += acc[i].px;
        s->newdelpress[3 * (size_t)i + 1] += acc[i].py;
        s->newdelpress[3 * (size_t)i + 2] += acc[i].pz;
    }

    /* ---- mykernel2 FluidGPU.cu:404-432 + Particle::update FluidGPU.cuh:270-304 ---- */
    {
        const double DT = P->dt;
        for (int i = 0; i < nlive; i++) {
            float *x = s->pos + 3 * (size_t)i, *v = s->vel + 3 * (size_t)i, *a = s->acc + 3 * (size_t)i;
            if (spts) {                                   /* :410-414 — pre-update state */
                spts[3 * (size_t)i] = x[0];
                spts[3 * (size_t)i + 1] = x[1];
                spts[3 * (size_t)i + 2] = x[2];
            }
            if (a3) a3[i] = s->dens[i];
            if (b3) b3[i] = (float)s->cell[i];
            int bnd = s->boundary[i] != 0;
            s->dens[i] = set_dens_h(s->newdens[i], bnd, P->h);          /* cuh:274 */
            s->press[i] = pressure_p(s->dens[i], P->sound);              /* cuh:275 */
            float *dp = s->delpress + 3 * (size_t)i, *ndp = s->newdelpress + 3 * (size_t)i;
            dp[0] = ndp[0]; dp[1] = ndp[1]; dp[2] = ndp[2];            /* cuh:276 */
            if (!bnd) {
                /* cuh:286-288 (DIFF == 0: + 0*diffusion, a float zero) */
                x[0] = x[0] + DT * v[0] + 0.0f;
                x[1] = x[1] + DT * v[1] + 0.0f;
                x[2] = x[2] + DT * v[2] + 0.0f;
                /* cuh:290-295 (stress_accel == 0 on this path) */
                double tx = (v[0] + DT * a[0] + DT * 0.0f);
                v[0] = tx - (tx > 0) * 0.003 + (tx < 0) * 0.003;
                v[0] *= (fabsf(v[0]) > 0.003);
                double ty = (v[1] + DT * a[1] + DT * 0.0f);
                v[1] = ty - (ty > 0) * 0.003 + (ty < 0) * 0.003;
                v[1] *= (fabsf(v[1]) > 0.003);
                v[2] = (v[2] + DT * a[2] + DT * 0.0f);
                v[2] *= (fabsf(v[2]) > 0.003);
                /* cuh:298-300 */
                a[0] = -(150.0 / s->dens[i]) * dp[0];
                a[1] = -(150.0 / s->dens[i]) * dp[1];
                a[2] = P->gravity + (-150.0 / s->dens[i]) * dp[2];
            }
            /* FluidGPU.cu:419-425 */
            {
                int g = P->grid;
                float fx = x[0] - P->origin, fy = x[1] - P->origin, fz = x[2] - P->origin;
                double qx = fx / P->cellsize, qy = fy / P->cellsize, qz = fz / P->cellsize;
                int cid;
                /* the reference would index start[]/end[] out of bounds (FluidGPU.cu:110) once the
                 * linear id leaves [0,numcells); the restatement parks such particles instead */
                if (!(fabs(qx) < 1e6 && fabs(qy) < 1e6 && fabs(qz) < 1e6)) cid = numcells;
                else {
                    long long l = (long long)(int)qx * g * g + (long long)(int)qy * g + (int)qz;
                    cid = (l < 0 || l >= numcells) ? numcells : (int)l;
                }
                s->cell[i] = cid;
            }
            s->newdens[i] = 0;
            ndp[0] = ndp[1] = ndp[2] = 0;
        }
        if (spts || a3 || b3)
            for (int i = nlive; i < n; i++) {             /* parked particles: export frozen state */
                if (spts) { spts[3 * (size_t)i] = s->pos[3 * (size_t)i]; spts[3 * (size_t)i + 1] = s->pos[3 * (size_t)i + 1]; spts[3 * (size_t)i + 2] = s->pos[3 * (size_t)i + 2]; }
                if (a3) a3[i] = s->dens[i];
                if (b3) b3[i] = (float)s->cell[i];
            }
    }
    if (stats) memcpy(stats, st, sizeof(st));
    rc = 0;
done:
    free(perm); free(tmp); free(start); free(end); free(acc); free(occ);
    return rc;
}
