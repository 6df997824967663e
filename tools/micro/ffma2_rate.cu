// Microbenchmark: issue / pipe rate of scalar FFMA vs packed FFMA2 (fma.rn.f32x2) on sm_100a, alone and
// interleaved with ALU-pipe work.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma2_rate ffma2_rate.cu
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
#define ITER 4096
template <int MODE>
__global__ void __launch_bounds__(256) k(float *out, float a, float b)
{
    float x[8];
    u64 p[8];
    unsigned m[4] = {threadIdx.x, 1, 2, 3};
    for (int i = 0; i < 8; i++) { x[i] = threadIdx.x + i; asm("mov.b64 %0, {%1,%2};" : "=l"(p[i]) : "f"(x[i]), "f"(x[i] + 1.f)); }
    u64 pa, pb;
    asm("mov.b64 %0, {%1,%2};" : "=l"(pa) : "f"(a), "f"(a));
    asm("mov.b64 %0, {%1,%2};" : "=l"(pb) : "f"(b), "f"(b));
    for (int it = 0; it < ITER; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            if (MODE == 0 || MODE == 2) x[i] = fmaf(x[i], a, b);
            if (MODE == 1 || MODE == 3) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[i]) : "l"(pa), "l"(pb));
            if (MODE == 2 || MODE == 3) m[i & 3] = (m[i & 3] ^ (m[(i + 1) & 3] >> 3)) + 0x9e37u;   // ALU pipe filler (LOP3/SHF/IADD)
        }
    }
    float s = 0;
    for (int i = 0; i < 8; i++) { float lo, hi; asm("mov.b64 {%0,%1}, %2;" : "=f"(lo), "=f"(hi) : "l"(p[i])); s += x[i] + lo + hi; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s + m[0] + m[1] + m[2] + m[3];
}
template <int MODE>
void run(const char *name, float *out, int sms, double flop_per_op)
{
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<sms * 8, 256>>>(out, 0.999f, 0.001f);
    cudaEventRecord(e0);
    k<MODE><<<sms * 8, 256>>>(out, 0.999f, 0.001f);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    double ops = (double)sms * 8 * 256 * ITER * 8;          // thread-level FMA instructions
    printf("%-28s %8.3f ms  %7.2f G thread-instr/s/SM  %7.2f TFLOP/s\n", name, ms, ops / ms / 1e6 / sms, ops * flop_per_op / ms / 1e9);
}
int main()
{
    cudaDeviceProp pr;
    cudaGetDeviceProperties(&pr, 0);
    int sms = pr.multiProcessorCount;
    float *out;
    cudaMalloc(&out, (size_t)sms * 8 * 256 * 4);
    printf("%s, %d SMs, %d MHz\n", pr.name, sms, pr.clockRate / 1000);
    run<0>("FFMA", out, sms, 2);
    run<1>("FFMA2", out, sms, 4);
    run<2>("FFMA + 3 ALU", out, sms, 2);
    run<3>("FFMA2 + 3 ALU", out, sms, 4);
    return 0;
}
