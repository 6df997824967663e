/*
 * ref_harness_unidyn.cu — drives the reference's UNMODIFIED unidyn kernels (FluidGPU-unidyn.cu, compiled
 * from /root/reference at build time into oracle/_ref/FluidGPU-unidyn.o) through the single-device loop
 * of solver-unidyn.cu:313-573 and dumps the particle state.  TEST INFRASTRUCTURE ONLY; this file is ours
 * and contains no reference source (it includes the reference header with -I/root/reference).
 *
 * Differences from solver-unidyn.cu, all outside the kernels:
 *   - thrust::sort_by_key<int,Particle> (:331) does not compile with CUB 2.8; replaced by a
 *     stable key+index sort and a gather of the records (same permutation).
 *   - single device only (the driver forces deviceCount = 1, :192-195); the multi-device block (:396-470)
 *     and the VTK dump (:472-493) are left out.
 *   - scene from a section file (--in), step count / dump steps from the command line.
 */
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include <thrust/device_ptr.h>
#include <thrust/functional.h>
#include <thrust/gather.h>
#include <thrust/sequence.h>
#include <thrust/sort.h>

#include "FluidGPU-unidyn.cuh"   /* reference header, found via -I at build time */

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(2); } } while (0)

struct Section { std::string name; int dtype; std::vector<char> data; long long count; };

static void put(FILE *f, const char *name, int dtype, long long count, const void *p)
{
    char nm[16] = {0};
    strncpy(nm, name, 15);
    fwrite(nm, 1, 16, f);
    fwrite(&dtype, 4, 1, f);
    fwrite(&count, 8, 1, f);
    fwrite(p, dtype == 2 ? 1 : 4, (size_t)count, f);
}
static std::vector<Section> read_sections(const char *path)
{
    std::vector<Section> out;
    FILE *f = fopen(path, "rb");
    if (!f) { fprintf(stderr, "cannot open %s\n", path); exit(2); }
    for (;;) {
        char nm[17] = {0};
        if (fread(nm, 1, 16, f) != 16) break;
        Section s;
        s.name = nm;
        if (fread(&s.dtype, 4, 1, f) != 1 || fread(&s.count, 8, 1, f) != 1) break;
        s.data.resize((size_t)s.count * (s.dtype == 2 ? 1 : 4));
        if (fread(s.data.data(), 1, s.data.size(), f) != s.data.size()) { fprintf(stderr, "short read\n"); exit(2); }
        out.push_back(s);
    }
    fclose(f);
    return out;
}
static const Section *find(const std::vector<Section> &v, const char *name)
{
    for (auto &s : v) if (s.name == name) return &s;
    return nullptr;
}

static void dump_state(const std::string &path, int step, const std::vector<Particle> &P, int nlive, const std::vector<int> &cells,
                       const std::vector<int> &start, const std::vector<int> &end, const std::vector<int> &split,
                       const float *spts, const float *a3, const float *b3)
{
    FILE *f = fopen(path.c_str(), "wb");
    if (!f) { fprintf(stderr, "cannot write %s\n", path.c_str()); exit(2); }
    int n = (int)P.size();
    std::vector<float> f3(3 * (size_t)n), f1(n);
    std::vector<int> i1(n);
    std::vector<unsigned char> b1(n);
    int hdr[5] = {n, NUMCELLS, step, (int)sizeof(Particle), nlive};
    put(f, "header", 1, 5, hdr);
#define V3(name, a, b, c) for (int i = 0; i < n; i++) { f3[3*(size_t)i] = P[i].a; f3[3*(size_t)i+1] = P[i].b; f3[3*(size_t)i+2] = P[i].c; } put(f, name, 0, 3LL*n, f3.data());
#define V1(name, a) for (int i = 0; i < n; i++) f1[i] = P[i].a; put(f, name, 0, n, f1.data());
    V3("pos", xcoord, ycoord, zcoord)
    V3("vel", xvel, yvel, zvel)
    V3("acc", xacc, yacc, zacc)
    V1("dens", dens)
    V1("press", press)
    V3("delpress", delpressx, delpressy, delpressz)
    V1("newdens", newdens)
    V3("newdelpress", newdelpressx, newdelpressy, newdelpressz)
    V3("diffusion", diffusionx, diffusiony, diffusionz)
    V1("solid", solid)
    V1("fluid", fluid)
    V1("delfluid", delfluid)
    V1("delsolid", delsolid)
    V1("mass", mass)
    for (int i = 0; i < n; i++) i1[i] = P[i].index;
    put(f, "index", 1, n, i1.data());
    for (int i = 0; i < n; i++) i1[i] = P[i].cellnumber;
    put(f, "cell", 1, n, i1.data());
    for (int i = 0; i < n; i++) i1[i] = P[i].subindex;
    put(f, "subindex", 1, n, i1.data());
    for (int i = 0; i < n; i++) b1[i] = P[i].boundary ? 1 : 0;
    put(f, "boundary", 2, n, b1.data());
    put(f, "cells_sorted", 1, (long long)cells.size(), cells.data());
    put(f, "start", 1, (long long)start.size(), start.data());
    put(f, "end", 1, (long long)end.size(), end.data());
    put(f, "split", 1, (long long)split.size(), split.data());
    put(f, "spts", 0, 3LL * n, spts);
    put(f, "a3", 0, n, a3);
    put(f, "b3", 0, n, b3);
    fclose(f);
}

int main(int argc, char **argv)
{
    const char *in = nullptr;
    std::string out = "ref_unidyn";
    int steps = 100;
    std::vector<int> dumps;
    for (int i = 1; i < argc; i++) {
        if (!strcmp(argv[i], "--in") && i + 1 < argc) in = argv[++i];
        else if (!strcmp(argv[i], "--out") && i + 1 < argc) out = argv[++i];
        else if (!strcmp(argv[i], "--steps") && i + 1 < argc) steps = atoi(argv[++i]);
        else if (!strcmp(argv[i], "--dump") && i + 1 < argc) {
            char *s = argv[++i];
            for (char *t = strtok(s, ","); t; t = strtok(nullptr, ",")) dumps.push_back(atoi(t));
        } else { fprintf(stderr, "usage: %s --in scene.bin [--out prefix] [--steps N] [--dump a,b,c]\n", argv[0]); return 2; }
    }
    if (!in) { fprintf(stderr, "--in is required\n"); return 2; }

    std::vector<Particle> SP;
    {
        auto sec = read_sections(in);
        const Section *pos = find(sec, "pos"), *vel = find(sec, "vel"), *acc = find(sec, "acc"), *dens = find(sec, "dens"),
                      *press = find(sec, "press"), *nd = find(sec, "newdens"), *idx = find(sec, "index"), *bnd = find(sec, "boundary"),
                      *sol = find(sec, "solid"), *flu = find(sec, "fluid");
        if (!pos) { fprintf(stderr, "scene has no pos section\n"); return 2; }
        int n = (int)(pos->count / 3);
        SP.resize(n);
        for (int j = 0; j < n; j++) {
            const float *p = (const float *)pos->data.data() + 3 * (size_t)j;
            bool b = bnd && bnd->data[j] != 0;
            Particle q = b ? Particle(p[0], p[1], p[2], true) : Particle(p[0], p[1], p[2], 0.f, 0.f, 0.f);
            if (vel) { const float *v = (const float *)vel->data.data() + 3 * (size_t)j; q.xvel = v[0]; q.yvel = v[1]; q.zvel = v[2]; }
            if (acc) { const float *a = (const float *)acc->data.data() + 3 * (size_t)j; q.xacc = a[0]; q.yacc = a[1]; q.zacc = a[2]; }
            if (dens) q.dens = ((const float *)dens->data.data())[j];
            if (press) q.press = ((const float *)press->data.data())[j];
            if (nd) q.newdens = ((const float *)nd->data.data())[j];
            q.index = idx ? ((const int *)idx->data.data())[j] : j;
            q.solid = sol ? ((const float *)sol->data.data())[j] : (b ? 1.f : 0.f);
            q.fluid = flu ? ((const float *)flu->data.data())[j] : (b ? 0.f : 1.f);
            q.subindex = 0;
            SP[j] = q;
        }
    }
    const int N = (int)SP.size();
    std::vector<int> keys(N), pidx(N);
    for (int j = 0; j < N; j++) {   /* the expression of solver-unidyn.cu:132 */
        SP[j].cellnumber = int((SP[j].xcoord - XMIN) / CELLSIZE) * GRIDSIZE * GRIDSIZE + int((SP[j].ycoord - YMIN) / CELLSIZE) * GRIDSIZE + int((SP[j].zcoord - ZMIN) / CELLSIZE);
        keys[j] = SP[j].cellnumber;
        pidx[j] = j;                /* identity, solver-unidyn.cu:233-238 */
    }

    Particle *d_SP, *d_tmp;
    int *v_d, *d_perm, *d_pidx, *d_start, *d_start_copy, *d_end, *d_split, *d_numsplit, *newsize;
    float *spts, *a3, *b3;
    CK(cudaMalloc(&d_SP, sizeof(Particle) * (size_t)N));
    CK(cudaMalloc(&d_tmp, sizeof(Particle) * (size_t)N));
    CK(cudaMalloc(&v_d, sizeof(int) * ((size_t)N + 2)));
    CK(cudaMemset(v_d, 0x7f, sizeof(int) * ((size_t)N + 2)));
    v_d += 1;
    CK(cudaMalloc(&d_perm, sizeof(int) * (size_t)N));
    CK(cudaMalloc(&d_pidx, sizeof(int) * (size_t)N));
    CK(cudaMalloc(&d_start, sizeof(int) * NUMCELLS));
    CK(cudaMalloc(&d_start_copy, sizeof(int) * NUMCELLS));
    CK(cudaMalloc(&d_end, sizeof(int) * NUMCELLS));
    CK(cudaMalloc(&d_split, sizeof(int) * NUMCELLS));
    CK(cudaMalloc(&d_numsplit, sizeof(int)));
    CK(cudaMallocManaged(&newsize, sizeof(int)));
    CK(cudaMallocManaged(&spts, sizeof(float) * 3 * (size_t)N));
    CK(cudaMallocManaged(&a3, sizeof(float) * (size_t)N));
    CK(cudaMallocManaged(&b3, sizeof(float) * (size_t)N));
    memset(spts, 0, sizeof(float) * 3 * (size_t)N);
    memset(a3, 0, sizeof(float) * (size_t)N);
    memset(b3, 0, sizeof(float) * (size_t)N);
    CK(cudaMemcpy(d_SP, SP.data(), sizeof(Particle) * (size_t)N, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(v_d, keys.data(), sizeof(int) * (size_t)N, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_pidx, pidx.data(), sizeof(int) * (size_t)N, cudaMemcpyHostToDevice));
    CK(cudaMemset(d_start, 0xff, sizeof(int) * NUMCELLS));        /* -1, solver-unidyn.cu:241-262 */
    CK(cudaMemset(d_start_copy, 0xff, sizeof(int) * NUMCELLS));
    CK(cudaMemset(d_end, 0xff, sizeof(int) * NUMCELLS));
    CK(cudaMemset(d_split, 0xff, sizeof(int) * NUMCELLS));
    CK(cudaMemset(d_numsplit, 0, sizeof(int)));
    int dsz = N;
    *newsize = dsz;                                                  /* solver-unidyn.cu:308-310 */

    thrust::device_ptr<Particle> t_a(d_SP), t_tmp(d_tmp);
    thrust::device_ptr<int> t_v(v_d), t_p(d_perm), t_1(d_start_copy), t_2(d_split);
    cudaEvent_t ev[2];
    for (auto &e : ev) CK(cudaEventCreate(&e));
    double ms = 0;
    std::vector<int> h_cells(N), h_start(NUMCELLS), h_end(NUMCELLS), h_split(NUMCELLS);

    for (int t = 0; t < steps; t++) {
        bool dump = false;
        for (int d : dumps) if (d == t + 1) dump = true;
        CK(cudaEventRecord(ev[0]));
        /* solver-unidyn.cu:331 (substituted, see header) */
        thrust::sequence(t_p, t_p + dsz);
        thrust::stable_sort_by_key(t_v, t_v + dsz, t_p);
        thrust::gather(t_p, t_p + dsz, t_a, t_tmp);
        CK(cudaMemcpyAsync(d_SP, d_tmp, sizeof(Particle) * (size_t)dsz, cudaMemcpyDeviceToDevice));
        count_after_merge<<<NUMCELLS, 1024>>>(v_d, d_pidx, dsz, newsize);                      /* :341 */
        CK(cudaDeviceSynchronize());
        dsz = *newsize;                                                                         /* :346 */
        findneighbours<<<NUMCELLS, 1024>>>(v_d, d_start, d_start_copy, d_end, dsz, 0);          /* :354 */
        mykernel<<<NUMCELLS, 1024>>>(d_SP, d_pidx, v_d, d_start, d_end, d_split, dsz, NUMCELLS, 0, 0, d_numsplit);   /* :363 */
        int numsplit = 0;
        CK(cudaMemcpy(&numsplit, d_numsplit, sizeof(int), cudaMemcpyDeviceToHost));            /* :368 */
        CK(cudaDeviceSynchronize());
        if (dump) {
            CK(cudaMemcpy(h_cells.data(), v_d, sizeof(int) * (size_t)N, cudaMemcpyDeviceToHost));
            CK(cudaMemcpy(h_start.data(), d_start, sizeof(int) * NUMCELLS, cudaMemcpyDeviceToHost));
            CK(cudaMemcpy(h_end.data(), d_end, sizeof(int) * NUMCELLS, cudaMemcpyDeviceToHost));
            CK(cudaMemcpy(h_split.data(), d_split, sizeof(int) * NUMCELLS, cudaMemcpyDeviceToHost));
        }
        thrust::sort_by_key(t_2, t_2 + NUMCELLS, t_1, thrust::greater<int>());                  /* :378 */
        if (numsplit > 0)
            mykernel3<<<numsplit * 8, 1024>>>(d_SP, d_pidx, v_d, d_start, d_end, d_split, dsz, NUMCELLS, 0, 0, d_numsplit);   /* :379 */
        CK(cudaDeviceSynchronize());
        mykernel2<<<NUMCELLS, 1024>>>(d_SP, d_pidx, v_d, d_start_copy, d_start, d_end, d_split, d_numsplit, dsz, NUMCELLS, 0, 0, t, spts, a3, b3);   /* :389 */
        cell_calc<<<NUMCELLS, 1024>>>(d_SP, d_pidx, v_d, dsz, 0);                               /* :548 */
        CK(cudaEventRecord(ev[1]));
        CK(cudaDeviceSynchronize());
        CK(cudaGetLastError());
        float e;
        CK(cudaEventElapsedTime(&e, ev[0], ev[1]));
        if (!dump) ms += e;
        if (dump) {
            CK(cudaMemcpy(SP.data(), d_SP, sizeof(Particle) * (size_t)N, cudaMemcpyDeviceToHost));
            dump_state(out + "_step" + std::to_string(t + 1) + ".bin", t + 1, SP, dsz, h_cells, h_start, h_end, h_split, spts, a3, b3);
        }
    }
    int timed = steps - (int)dumps.size();
    printf("{\"impl\": \"reference-gpu\", \"path\": \"unidyn\", \"n\": %d, \"numcells\": %d, \"steps\": %d, \"ms_per_step\": %.6f}\n",
           N, NUMCELLS, steps, timed > 0 ? ms / timed : 0.0);
    return 0;
}
