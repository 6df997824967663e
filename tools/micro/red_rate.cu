// Microbenchmark: throughput of red.global.add.v4.f32 (vector float reduction, sm_90+) in the access pattern a
// symmetric pair kernel would produce: every warp adds 32 consecutive float4 (512 B) to an array, every element
// is hit HITS times per pass by different warps at different times.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o red_rate red_rate.cu
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void red4(float4 *p, float4 v)
{
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
template <int MODE>
__global__ void __launch_bounds__(256) k(float4 *a, long long nchunk, int hits, long long shift)
{
    const int lane = threadIdx.x & 31;
    long long w = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nw = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long c = w; c < nchunk; c += nw)
        for (int h = 0; h < hits; h++) {
            long long t = (c + h * shift) % nchunk;
            float4 v = make_float4(1.f, 2.f, 3.f, (float)h);
            if (MODE == 0) red4(a + t * 32 + lane, v);
            if (MODE == 1) { float *f = (float *)(a + t * 32 + lane); atomicAdd(f, v.x); atomicAdd(f + 1, v.y); atomicAdd(f + 2, v.z); atomicAdd(f + 3, v.w); }
            if (MODE == 2) a[t * 32 + lane] = v;
        }
}
template <int MODE>
void run(const char *name, float4 *a, long long n, int hits, long long shift, int sms)
{
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<sms * 8, 256>>>(a, n / 32, hits, shift);
    cudaEventRecord(e0);
    k<MODE><<<sms * 8, 256>>>(a, n / 32, hits, shift);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    double ops = (double)n * hits;
    printf("%-34s n=%9lld hits=%2d shift=%7lld  %8.3f ms  %7.2f G float4-ops/s  %7.1f GB/s payload  (%s)\n", name, n, hits, shift, ms,
           ops / ms / 1e6, ops * 16 / ms / 1e6, cudaGetErrorString(cudaGetLastError()));
}
int main()
{
    cudaDeviceProp pr;
    cudaGetDeviceProperties(&pr, 0);
    int sms = pr.multiProcessorCount;
    const long long NMAX = 64ll << 20;     // 64 Mi float4 = 1 GiB
    float4 *a;
    cudaMalloc(&a, NMAX * 16);
    cudaMemset(a, 0, NMAX * 16);
    for (long long n : {4ll << 20, 64ll << 20}) {             // 64 MiB (L2 resident) and 1 GiB
        for (long long shift : {1ll, 37ll, 4099ll}) {
            run<0>("red.global.add.v4.f32", a, n, 14, shift, sms);
        }
        run<1>("4 x atomicAdd(float)", a, n, 14, 37, sms);
        run<2>("plain st.v4 (same pattern)", a, n, 14, 37, sms);
    }
    return 0;
}
