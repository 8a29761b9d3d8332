// FP64 peak probes for B200 (sm_100a): register-resident DMMA (mma.sync m8n8k4 f64) and DFMA loops.
// Yardstick only (never on the product path). Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_peak fp64_peak.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

template <int NACC>
__global__ void __launch_bounds__(256) dmma_loop(double* out, int iters, double a0, double b0) {
    double c[NACC][2];
#pragma unroll
    for (int i = 0; i < NACC; i++) { c[i][0] = 0.0; c[i][1] = 0.0; }
    double a = a0 + threadIdx.x * 1e-9, b = b0;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < NACC; i++) {
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                         : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NACC; i++) s += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int NACC>
__global__ void __launch_bounds__(256) dfma_loop(double* out, int iters, double a0, double b0) {
    double c[NACC];
#pragma unroll
    for (int i = 0; i < NACC; i++) c[i] = i;
    double a = a0 + threadIdx.x * 1e-9, b = b0;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < NACC; i++) c[i] = fma(a, c[i], b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NACC; i++) s += c[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    printf("device %s SMs %d clock %d kHz\n", p.name, p.multiProcessorCount, p.clockRate);
    double* out; CK(cudaMalloc(&out, sizeof(double) * 148 * 8 * 1024));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    int nsm = p.multiProcessorCount;
    for (int wps = 4; wps <= 32; wps *= 2) {       // warps per SM
        int threads = 256, blocks = nsm * wps * 32 / threads; if (blocks < nsm) { threads = wps * 32; blocks = nsm; }
        int iters = 20000;
        for (int rep = 0; rep < 3; rep++) {
            CK(cudaEventRecord(e0));
            dmma_loop<16><<<blocks, threads>>>(out, iters, 1.0, 1e-3);
            CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
            float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
            double flops = (double)blocks * (threads / 32) * iters * 16 * 512.0;
            if (rep == 2) printf("DMMA m8n8k4 warps/SM %2d: %.3f ms  %.2f TFLOP/s\n", wps, ms, flops / ms * 1e-9);
        }
    }
    for (int wps = 8; wps <= 32; wps *= 2) {
        int threads = 256, blocks = nsm * wps * 32 / threads;
        int iters = 20000;
        for (int rep = 0; rep < 3; rep++) {
            CK(cudaEventRecord(e0));
            dfma_loop<16><<<blocks, threads>>>(out, iters, 1.0000001, 1e-3);
            CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
            float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
            double flops = (double)blocks * threads * iters * 16 * 2.0;
            if (rep == 2) printf("DFMA warps/SM %2d: %.3f ms  %.2f TFLOP/s\n", wps, ms, flops / ms * 1e-9);
        }
    }
    // sustained DMMA for ~3 s with 8 warps/SM
    {
        int threads = 256, blocks = nsm; int iters = 200000;
        CK(cudaEventRecord(e0));
        for (int r = 0; r < 10; r++) dmma_loop<16><<<blocks, threads>>>(out, iters, 1.0, 1e-3);
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        double flops = 10.0 * blocks * (threads / 32) * iters * 16 * 512.0;
        printf("DMMA sustained (8 warps/SM, %.0f ms): %.2f TFLOP/s\n", ms, flops / ms * 1e-9);
    }
    return 0;
}
