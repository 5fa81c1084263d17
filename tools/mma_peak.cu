// tools/mma_peak.cu -- measures the FP64 ceilings the f64 GEMM roofline is quoted against on THIS box:
// DFMA (vector pipe) and DMMA mma.sync m8n8k4 / m16n8k4 / m16n8k8 / m16n8k16, registers only (no memory traffic).
// build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o build/mma_peak tools/mma_peak.cu
#include <cstdio>
#include <cuda_runtime.h>

constexpr int ITERS = 2048;
constexpr int CHAINS = 4;

__global__ void k_dfma(double *out, double s) {
    double acc[8];
    for (int i = 0; i < 8; i++) acc[i] = threadIdx.x + i;
    for (int it = 0; it < ITERS; it++)
#pragma unroll
        for (int i = 0; i < 8; i++) acc[i] = fma(acc[i], s, 1e-9);
    double r = 0;
    for (int i = 0; i < 8; i++) r += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

template <int SHAPE>  // 0: m8n8k4, 1: m16n8k4, 2: m16n8k8, 3: m16n8k16
__global__ void k_dmma(double *out, double s) {
    double c[CHAINS][4];
    double a[8], b[4];
    for (int i = 0; i < 8; i++) a[i] = s * (threadIdx.x + i);
    for (int i = 0; i < 4; i++) b[i] = s * (i + 1);
    for (int j = 0; j < CHAINS; j++)
        for (int i = 0; i < 4; i++) c[j][i] = j + i;
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int j = 0; j < CHAINS; j++) {
            if (SHAPE == 0)
                asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                             : "+d"(c[j][0]), "+d"(c[j][1]) : "d"(a[0]), "d"(b[0]));
            else if (SHAPE == 1)
                asm volatile("mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
                             : "+d"(c[j][0]), "+d"(c[j][1]), "+d"(c[j][2]), "+d"(c[j][3]) : "d"(a[0]), "d"(a[1]), "d"(b[0]));
            else if (SHAPE == 2)
                asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                             : "+d"(c[j][0]), "+d"(c[j][1]), "+d"(c[j][2]), "+d"(c[j][3])
                             : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
            else
                asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};"
                             : "+d"(c[j][0]), "+d"(c[j][1]), "+d"(c[j][2]), "+d"(c[j][3])
                             : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]),
                               "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
        }
    }
    double r = 0;
    for (int j = 0; j < CHAINS; j++)
        for (int i = 0; i < 4; i++) r += c[j][i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

template <typename F> float time_it(F f) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    f();
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    for (int i = 0; i < 5; i++) f();
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    return ms / 5;
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    const int sms = p.multiProcessorCount;
    double *out;
    cudaMalloc(&out, sizeof(double) * sms * 8 * 512);
    for (int warps = 4; warps <= 16; warps *= 2) {
        const int threads = warps * 32, blocks = sms * 2;
        float ms = time_it([&] { k_dfma<<<blocks, threads>>>(out, 0.999); });
        printf("DFMA            warps/CTA %2d x2 CTA/SM: %7.2f TFLOP/s\n", warps, 2.0 * blocks * threads * 8.0 * ITERS / ms / 1e9);
        const double flops[4] = {2.0 * 8 * 8 * 4, 2.0 * 16 * 8 * 4, 2.0 * 16 * 8 * 8, 2.0 * 16 * 8 * 16};
        const char *names[4] = {"DMMA m8n8k4  ", "DMMA m16n8k4 ", "DMMA m16n8k8 ", "DMMA m16n8k16"};
        for (int s = 0; s < 4; s++) {
            float t;
            if (s == 0) t = time_it([&] { k_dmma<0><<<blocks, threads>>>(out, 0.5); });
            else if (s == 1) t = time_it([&] { k_dmma<1><<<blocks, threads>>>(out, 0.5); });
            else if (s == 2) t = time_it([&] { k_dmma<2><<<blocks, threads>>>(out, 0.5); });
            else t = time_it([&] { k_dmma<3><<<blocks, threads>>>(out, 0.5); });
            printf("%s   warps/CTA %2d x2 CTA/SM: %7.2f TFLOP/s\n", names[s], warps,
                   flops[s] * CHAINS * ITERS * (double)blocks * warps / t / 1e9);
        }
    }
    cudaError_t e = cudaDeviceSynchronize();
    printf("status: %s\n", cudaGetErrorString(e));
    return 0;
}
