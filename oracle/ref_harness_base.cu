/*
 * ref_harness_base.cu — drives the reference's UNMODIFIED base kernels (FluidGPU.cu, compiled from
 * /root/reference at build time into oracle/_ref/FluidGPU.o) through the loop of solver.cu:171-216
 * and dumps the full particle state, so that tests can pin the CPU oracle and the B200 library
 * against what the reference's own CUDA code computes.
 *
 * TEST INFRASTRUCTURE ONLY.  This file is ours; it contains no reference source.  It includes the
 * reference header (for `class Particle` and the kernel prototypes) with -I/root/reference.
 *
 * Differences from solver.cu, all outside the kernels:
 *   - thrust::sort_by_key<int,Particle> (solver.cu:181) does not compile with CUB 2.8 (static smem
 *     overflow); replaced by a stable key+index sort and a gather of the Particle
 *     records, which yields the same permutation (both are stable sorts by bin id).
 *   - the never-read 2 GB `neighbours` array (solver.cu:89-95) is not allocated.
 *   - scene may come from a section file (--in) instead of the built-in lattice (solver.cu:115-121).
 *   - step count / dump steps are command-line arguments (tpts is a compile-time const, solver.cu:19).
 *
 * Section file format (little endian), repeated until EOF:
 *   char name[16]; int32 dtype (0=f32,1=i32,2=u8); int64 count; payload.
 */
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include <thrust/device_ptr.h>
#include <thrust/gather.h>
#include <thrust/sequence.h>
#include <thrust/sort.h>

#include "FluidGPU.cuh"   /* reference header, found via -I at build time */

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

static void dump_state(const std::string &path, int step, const std::vector<Particle> &P, const std::vector<int> &cells,
                       const std::vector<int> &start, const std::vector<int> &end,
                       const float *spts, const float *a3, const float *b3)
{
    FILE *f = fopen(path.c_str(), "wb");
    if (!f) { fprintf(stderr, "cannot write %s\n", path.c_str()); exit(2); }
    int n = (int)P.size();
    std::vector<float> f3(3 * (size_t)n), f1(n);
    std::vector<int> i1(n);
    std::vector<unsigned char> b1(n);
    int hdr[4] = {n, NUMCELLS, step, (int)sizeof(Particle)};
    put(f, "header", 1, 4, hdr);
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
    for (int i = 0; i < n; i++) i1[i] = P[i].index;
    put(f, "index", 1, n, i1.data());
    for (int i = 0; i < n; i++) i1[i] = P[i].cellnumber;
    put(f, "cell", 1, n, i1.data());
    for (int i = 0; i < n; i++) b1[i] = P[i].boundary ? 1 : 0;
    put(f, "boundary", 2, n, b1.data());
    put(f, "cells_sorted", 1, (long long)cells.size(), cells.data());   /* key array as mykernel saw it */
    put(f, "start", 1, (long long)start.size(), start.data());
    put(f, "end", 1, (long long)end.size(), end.data());
    put(f, "spts", 0, 3LL * n, spts);
    put(f, "a3", 0, n, a3);
    put(f, "b3", 0, n, b3);
    fclose(f);
}

int main(int argc, char **argv)
{
    const char *in = nullptr;
    std::string out = "ref_base";
    int steps = 100;
    std::vector<int> dumps;
    for (int i = 1; i < argc; i++) {
        if (!strcmp(argv[i], "--in") && i + 1 < argc) in = argv[++i];
        else if (!strcmp(argv[i], "--out") && i + 1 < argc) out = argv[++i];
        else if (!strcmp(argv[i], "--steps") && i + 1 < argc) steps = atoi(argv[++i]);
        else if (!strcmp(argv[i], "--dump") && i + 1 < argc) {
            char *s = argv[++i];
            for (char *t = strtok(s, ","); t; t = strtok(nullptr, ",")) dumps.push_back(atoi(t));
        } else { fprintf(stderr, "usage: %s [--in scene.bin] [--out prefix] [--steps N] [--dump a,b,c]\n", argv[0]); return 2; }
    }

    std::vector<Particle> SP;
    if (in) {
        auto sec = read_sections(in);
        const Section *pos = find(sec, "pos"), *vel = find(sec, "vel"), *acc = find(sec, "acc"), *dens = find(sec, "dens"),
                      *press = find(sec, "press"), *nd = find(sec, "newdens"), *ndp = find(sec, "newdelpress"),
                      *idx = find(sec, "index"), *bnd = find(sec, "boundary");
        if (!pos) { fprintf(stderr, "scene has no pos section\n"); return 2; }
        int n = (int)(pos->count / 3);
        SP.resize(n);
        for (int j = 0; j < n; j++) {
            const float *p = (const float *)pos->data.data() + 3 * (size_t)j;
            Particle q(p[0], p[1], p[2]);
            if (vel) { const float *v = (const float *)vel->data.data() + 3 * (size_t)j; q.xvel = v[0]; q.yvel = v[1]; q.zvel = v[2]; }
            if (acc) { const float *a = (const float *)acc->data.data() + 3 * (size_t)j; q.xacc = a[0]; q.yacc = a[1]; q.zacc = a[2]; }
            if (dens) q.dens = ((const float *)dens->data.data())[j];
            if (press) q.press = ((const float *)press->data.data())[j];
            if (nd) q.newdens = ((const float *)nd->data.data())[j];
            if (ndp) { const float *a = (const float *)ndp->data.data() + 3 * (size_t)j; q.newdelpressx = a[0]; q.newdelpressy = a[1]; q.newdelpressz = a[2]; }
            q.index = idx ? ((const int *)idx->data.data())[j] : j;
            if (bnd) q.boundary = bnd->data[j] != 0;
            SP[j] = q;
        }
    } else {
        /* the scene of solver.cu:115-121 (nspts = 8000, nbpts = 0, solver.cu:17-18) */
        const int n = 8000;
        SP.resize(n);
        for (int j = 0; j < n; j++) {
            SP[j] = Particle(-.16 + 0.04 * ((j / 15) % 15), -0.76 + 0.04 * (j / 15 / 15), -0.20 + (j % 15) * 0.04, 0., 0., 0.);
            SP[j].index = j;
            SP[j].solid = true;
        }
    }
    const int N = (int)SP.size();
    std::vector<int> keys(N);
    for (int j = 0; j < N; j++) {   /* the expression of solver.cu:119 */
        SP[j].cellnumber = int((SP[j].xcoord - XMIN) / CELLSIZE) * GRIDSIZE * GRIDSIZE + int((SP[j].ycoord - YMIN) / CELLSIZE) * GRIDSIZE + int((SP[j].zcoord - ZMIN) / CELLSIZE);
        keys[j] = SP[j].cellnumber;
    }

    Particle *d_SP, *d_tmp;
    int *v_d, *d_perm, *d_start, *d_end;
    float *spts, *a3, *b3;
    CK(cudaMalloc(&d_SP, sizeof(Particle) * (size_t)N));
    CK(cudaMalloc(&d_tmp, sizeof(Particle) * (size_t)N));
    CK(cudaMalloc(&v_d, sizeof(int) * ((size_t)N + 2)));
    v_d += 1;   /* findneighbours reads cell[-1] and cell[n] (FluidGPU.cu:109,112): keep them inside the allocation */
    CK(cudaMalloc(&d_perm, sizeof(int) * (size_t)N));
    CK(cudaMalloc(&d_start, sizeof(int) * NUMCELLS));
    CK(cudaMalloc(&d_end, sizeof(int) * NUMCELLS));
    CK(cudaMallocManaged(&spts, sizeof(float) * 3 * (size_t)N));
    CK(cudaMallocManaged(&a3, sizeof(float) * (size_t)N));
    CK(cudaMallocManaged(&b3, sizeof(float) * (size_t)N));
    memset(spts, 0, sizeof(float) * 3 * (size_t)N);
    memset(a3, 0, sizeof(float) * (size_t)N);
    memset(b3, 0, sizeof(float) * (size_t)N);
    CK(cudaMemcpy(d_SP, SP.data(), sizeof(Particle) * (size_t)N, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(v_d, keys.data(), sizeof(int) * (size_t)N, cudaMemcpyHostToDevice));
    CK(cudaMemset(d_start, 0xff, sizeof(int) * NUMCELLS));   /* -1, solver.cu:163-169 */
    CK(cudaMemset(d_end, 0xff, sizeof(int) * NUMCELLS));

    thrust::device_ptr<Particle> t_a(d_SP), t_tmp(d_tmp);
    thrust::device_ptr<int> t_v(v_d), t_p(d_perm);
    cudaEvent_t ev[5];
    for (auto &e : ev) CK(cudaEventCreate(&e));
    double ms[4] = {0, 0, 0, 0};
    std::vector<int> h_cells(N), h_start(NUMCELLS), h_end(NUMCELLS);

    for (int t = 0; t < steps; t++) {
        bool dump = false;
        for (int d : dumps) if (d == t + 1) dump = true;
        CK(cudaEventRecord(ev[0]));
        /* solver.cu:181 (substituted, see header) */
        thrust::sequence(t_p, t_p + N);
        thrust::stable_sort_by_key(t_v, t_v + N, t_p);
        thrust::gather(t_p, t_p + N, t_a, t_tmp);
        CK(cudaMemcpyAsync(d_SP, d_tmp, sizeof(Particle) * (size_t)N, cudaMemcpyDeviceToDevice));
        CK(cudaEventRecord(ev[1]));
        findneighbours<<<NUMCELLS, 1024>>>(v_d, d_start, d_end, N);           /* solver.cu:182 */
        CK(cudaEventRecord(ev[2]));
        if (dump) {
            CK(cudaMemcpy(h_cells.data(), v_d, sizeof(int) * (size_t)N, cudaMemcpyDeviceToHost));
            CK(cudaMemcpy(h_start.data(), d_start, sizeof(int) * NUMCELLS, cudaMemcpyDeviceToHost));
            CK(cudaMemcpy(h_end.data(), d_end, sizeof(int) * NUMCELLS, cudaMemcpyDeviceToHost));
            CK(cudaEventRecord(ev[2]));
        }
        mykernel<<<NUMCELLS, 64>>>(d_SP, v_d, d_start, d_end, N);             /* solver.cu:187 */
        CK(cudaEventRecord(ev[3]));
        mykernel2<<<NUMCELLS, 1024>>>(d_SP, v_d, d_start, d_end, N, spts, a3, b3);   /* solver.cu:198 */
        CK(cudaEventRecord(ev[4]));
        CK(cudaDeviceSynchronize());
        CK(cudaGetLastError());
        float e;
        CK(cudaEventElapsedTime(&e, ev[0], ev[1])); ms[0] += e;
        if (!dump) { CK(cudaEventElapsedTime(&e, ev[1], ev[2])); ms[1] += e; }
        CK(cudaEventElapsedTime(&e, ev[2], ev[3])); ms[2] += e;
        CK(cudaEventElapsedTime(&e, ev[3], ev[4])); ms[3] += e;
        if (dump) {
            CK(cudaMemcpy(SP.data(), d_SP, sizeof(Particle) * (size_t)N, cudaMemcpyDeviceToHost));
            dump_state(out + "_step" + std::to_string(t + 1) + ".bin", t + 1, SP, h_cells, h_start, h_end, spts, a3, b3);
        }
    }
    printf("{\"impl\": \"reference-gpu\", \"path\": \"base\", \"n\": %d, \"numcells\": %d, \"steps\": %d, "
           "\"ms_sort\": %.6f, \"ms_findneighbours\": %.6f, \"ms_mykernel\": %.6f, \"ms_mykernel2\": %.6f, \"ms_per_step\": %.6f}\n",
           N, NUMCELLS, steps, ms[0] / steps, ms[1] / steps, ms[2] / steps, ms[3] / steps, (ms[0] + ms[1] + ms[2] + ms[3]) / steps);
    return 0;
}
