// Micro-test: are generic-proxy global stores of kernel N visible to cp.async.bulk (async proxy) reads of kernel N+1?
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)
__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void writer(double* buf, size_t n, double val, int fence) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) buf[i] = val + (double)(i & 1023);
    if (fence) { __threadfence(); asm volatile("fence.proxy.async.global;" ::: "memory"); }
}
// read-modify-write like the GEMM epilogue: load with ld.cg, store
__global__ void rmw(double* buf, size_t n, double delta) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) buf[i] = __ldcg(buf + i) + delta;
}
__global__ void reader(const double* buf, size_t nchunks, double val, unsigned long long* bad, int mode) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ uint64_t bar;
    double* s = reinterpret_cast<double*>(smem);
    if (threadIdx.x == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&bar))); asm volatile("fence.mbarrier_init.release.cluster;"); }
    __syncthreads();
    uint32_t phase = 0;
    for (size_t c = blockIdx.x; c < nchunks; c += gridDim.x) {
        const double* src = buf + c * 2048;
        if (mode == 0) {
            if (threadIdx.x == 0) {
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(&bar)), "r"(16384));
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(s32(s)), "l"(src), "r"(16384), "r"(s32(&bar)) : "memory");
            }
            asm volatile("{\n.reg .pred p;\nW1:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D1;\nbra W1;\nD1:\n}\n" ::"r"(s32(&bar)), "r"(phase) : "memory");
            phase ^= 1;
        } else {
            for (int e = threadIdx.x; e < 2048; e += blockDim.x) s[e] = __ldcg(src + e);
            __syncthreads();
        }
        unsigned long long nb = 0;
        for (int e = threadIdx.x; e < 2048; e += blockDim.x) {
            size_t i = c * 2048 + e;
            if (s[e] != val + (double)(i & 1023)) nb++;
        }
        if (nb) atomicAdd(bad, nb);
        __syncthreads();
    }
}
int main(int argc, char** argv) {
    size_t mb = argc > 1 ? atol(argv[1]) : 512;
    size_t n = mb * 1024 * 1024 / 8, nchunks = n / 2048;
    double* buf; CK(cudaMalloc(&buf, n * 8));
    unsigned long long* bad; CK(cudaMallocManaged(&bad, 8));
    for (int variant = 0; variant < 4; variant++) {
        int fence = variant & 1, use_rmw = variant >> 1;
        unsigned long long total_bulk = 0, total_gen = 0;
        for (int rep = 0; rep < 30; rep++) {
            double val = 0;
            if (!use_rmw) { val = rep * 7.0 + 1; writer<<<148 * 4, 256>>>(buf, n, val, fence); }
            else { if (rep == 0) writer<<<148 * 4, 256>>>(buf, n, 0.0, 1); rmw<<<148 * 4, 256>>>(buf, n, 3.0); val = 3.0 * (rep + 1); }
            *bad = 0;
            reader<<<148, 256, 16384>>>(buf, nchunks, val, bad, 0);
            CK(cudaDeviceSynchronize()); total_bulk += *bad;
            *bad = 0;
            reader<<<148, 256, 16384>>>(buf, nchunks, val, bad, 1);
            CK(cudaDeviceSynchronize()); total_gen += *bad;
        }
        printf("buffer %zu MB variant fence=%d rmw=%d: mismatching doubles via bulk copy %llu, via ld.global.cg %llu\n", mb, fence, use_rmw, total_bulk, total_gen);
    }
    return 0;
}
