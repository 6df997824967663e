/*
 * ref_bookkeeping.cu — drives the unidyn bookkeeping kernels find_idx / mem_shift / count_after_merge (FluidGPU-unidyn.cuh:541-544,
 * FluidGPU-unidyn.cu:499-562) the way the 2-device block of solver-unidyn.cu:396-470 does, on a synthetic sorted key array, and
 * prints what they leave behind.  TEST INFRASTRUCTURE ONLY; this file is ours (it includes the reference header with
 * -I/root/reference at build time).  Built twice by oracle/Makefile: against the reference's own FluidGPU-unidyn.o
 * (_ref/ref_bookkeeping) and against libfsg's link-compatible object (_ref/compat_bookkeeping); tests/test_parity_gpu.py compares
 * the two outputs byte for byte.
 *
 * Two things keep the REFERENCE side deterministic: the key array has one sentinel element behind its end (find_idx reads
 * SPptr[idx + 1] for idx = npts - 1, FluidGPU-unidyn.cu:505-523), and mem_shift is launched as ONE block (its __syncthreads()
 * orders the copy out and the copy back only inside a block, :531-542).
 */
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "FluidGPU-unidyn.cuh"

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(2); } } while (0)

static unsigned long long fnv(const void *p, size_t n)
{
    unsigned long long h = 1469598103934665603ull;
    const unsigned char *b = (const unsigned char *)p;
    for (size_t i = 0; i < n; i++) { h ^= b[i]; h *= 1099511628211ull; }
    return h;
}

int main(int argc, char **argv)
{
    const int seed = argc > 1 ? atoi(argv[1]) : 1;
    srand(seed);
    const int buffer = GRIDSIZE * GRIDSIZE;                 /* solver-unidyn.cu:187 */
    /* a sorted key array covering both halves of the bin range, with a parked tail (keys >= NUMCELLS) */
    const int npts = 900, parked = 37;
    std::vector<int> keys(npts + 1);
    for (int i = 0; i < npts - parked; i++) keys[i] = rand() % NUMCELLS;
    for (int a = 0; a < npts - parked; a++)                 /* insertion sort: tiny */
        for (int b = a; b > 0 && keys[b - 1] > keys[b]; b--) { int t = keys[b]; keys[b] = keys[b - 1]; keys[b - 1] = t; }
    for (int i = npts - parked; i < npts; i++) keys[i] = NUMCELLS + (i % 2);
    keys[npts] = 0x7fffffff;                                /* the element find_idx reads behind the end */
    int *d_keys, *d_out;
    CK(cudaMalloc(&d_keys, sizeof(int) * (npts + 1)));
    CK(cudaMalloc(&d_out, sizeof(int) * 16));
    CK(cudaMemcpy(d_keys, keys.data(), sizeof(int) * (npts + 1), cudaMemcpyHostToDevice));
    printf("{\"seed\": %d, \"npts\": %d", seed, npts);
    /* count_after_merge, solver-unidyn.cu:341 */
    {
        int h = -1;
        CK(cudaMemcpy(d_out, &h, sizeof(int), cudaMemcpyHostToDevice));
        count_after_merge<<<NUMCELLS, 1024>>>(d_keys, d_keys, npts, d_out);
        CK(cudaDeviceSynchronize());
        CK(cudaMemcpy(&h, d_out, sizeof(int), cudaMemcpyDeviceToHost));
        printf(", \"newsize\": %d", h);
    }
    const int live = npts - parked;
    /* find_idx for both devices, solver-unidyn.cu:404,432 */
    for (int dev = 0; dev < 2; dev++) {
        int h[4] = {-7, -7, -7, -7};
        CK(cudaMemcpy(d_out, h, sizeof h, cudaMemcpyHostToDevice));
        find_idx<<<(live + 1023) / 1024, 1024>>>(d_keys, dev, live, buffer, d_out, d_out + 1, d_out + 2, d_out + 3);
        CK(cudaDeviceSynchronize());
        CK(cudaMemcpy(h, d_out, sizeof h, cudaMemcpyDeviceToHost));
        printf(", \"find_idx_dev%d\": [%d, %d, %d, %d]", dev, h[0], h[1], h[2], h[3]);
    }
    /* mem_shift: records [left, right] move down by `shifts` through the staging buffer, solver-unidyn.cu:447-466 */
    {
        const int nrec = 1000, left = 300, right = 811, shifts = 123;
        std::vector<unsigned char> rec((size_t)nrec * sizeof(Particle));
        for (size_t i = 0; i < rec.size(); i++) rec[i] = (unsigned char)(rand() & 0xff);
        /* the reference moves records by Particle assignment: padding bytes (317-319, 338-339 of the unidyn record, FluidGPU-unidyn.cuh
           offsets in SURVEY.md §8 a1) need not travel and bools are only defined for 0 / 1 — keep both out of the comparison */
        static const int pad[] = {317, 318, 319, 338, 339}, bools[] = {316, 336, 337};
        for (int r = 0; r < nrec; r++) {
            for (int p : pad) rec[(size_t)r * sizeof(Particle) + p] = 0;
            for (int b : bools) rec[(size_t)r * sizeof(Particle) + b] &= 1;
        }
        Particle *d_p, *d_b;
        CK(cudaMalloc(&d_p, rec.size()));
        CK(cudaMalloc(&d_b, rec.size()));
        CK(cudaMemset(d_b, 0, rec.size()));
        CK(cudaMemcpy(d_p, rec.data(), rec.size(), cudaMemcpyHostToDevice));
        mem_shift<<<1, 1024>>>(d_p, d_b, d_keys, d_keys, 0, shifts, left, right);
        CK(cudaDeviceSynchronize());
        std::vector<unsigned char> out(rec.size());
        CK(cudaMemcpy(out.data(), d_p, rec.size(), cudaMemcpyDeviceToHost));
        for (int r = 0; r < nrec; r++) for (int p : pad) out[(size_t)r * sizeof(Particle) + p] = 0;
        /* expectation independent of either implementation: [left - shifts, right - shifts] holds the old [left, right], the rest is untouched */
        std::vector<unsigned char> want(rec);
        memmove(&want[(size_t)(left - shifts) * sizeof(Particle)], &rec[(size_t)left * sizeof(Particle)], (size_t)(right - left + 1) * sizeof(Particle));
        printf(", \"sizeof_particle\": %d, \"mem_shift_hash\": \"%016llx\", \"mem_shift_as_expected\": %s", (int)sizeof(Particle), fnv(out.data(), out.size()),
               memcmp(out.data(), want.data(), out.size()) == 0 ? "true" : "false");
        /* shifts == 0 is a no-op (:533) */
        mem_shift<<<1, 1024>>>(d_p, d_b, d_keys, d_keys, 0, 0, left, right);
        CK(cudaDeviceSynchronize());
        std::vector<unsigned char> out2(rec.size());
        CK(cudaMemcpy(out2.data(), d_p, rec.size(), cudaMemcpyDeviceToHost));
        for (int r = 0; r < nrec; r++) for (int p : pad) out2[(size_t)r * sizeof(Particle) + p] = 0;
        printf(", \"mem_shift_zero_is_noop\": %s", out2 == out ? "true" : "false");
    }
    printf("}\n");
    return 0;
}
