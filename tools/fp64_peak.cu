// FP64 peak micro-benchmark for B200 (sm_100a).
//
// Measures the two FP64 issue rates that bound the Cholesky path:
//   (1) DMMA.8x8x4  via mma.sync.aligned.m8n8k4.row.col.f64 (the only native FP64 MMA on
//       sm_100a; m16n8k{4,8,16} lower to sequences of it),
//   (2) plain DFMA.
// Prints one JSON line; bench.py / DESIGN.md quote it as the "FP64 tensor peak" denominator.
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/fp64_peak tools/fp64_peak.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { \
    fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1);} } while (0)

__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

template <int NACC>
__global__ void __launch_bounds__(256) dmma_loop(double* out, int iters, double a0, double b0) {
    double acc[NACC][2];
#pragma unroll
    for (int i = 0; i < NACC; ++i) { acc[i][0] = 0.0; acc[i][1] = 0.0; }
    double a = a0 + threadIdx.x * 1e-9, b = b0;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i) dmma884(acc[i][0], acc[i][1], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NACC; ++i) s += acc[i][0] + acc[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int NACC>
__global__ void __launch_bounds__(256) dfma_loop(double* out, int iters, double a0, double b0) {
    double acc[NACC];
#pragma unroll
    for (int i = 0; i < NACC; ++i) acc[i] = i;
    double a = a0 + threadIdx.x * 1e-9, b = b0;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i) acc[i] = fma(a, acc[i], b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NACC; ++i) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main(int argc, char** argv) {
    int dev = 0;
    CK(cudaSetDevice(dev));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, dev));
    int sms = prop.multiProcessorCount;
    int ctas_per_sm = 2, threads = 256;
    int grid = sms * ctas_per_sm;
    double* out;
    CK(cudaMalloc(&out, sizeof(double) * grid * threads));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    const int NACC = 8;
    int iters = 20000;

    // warm-up
    dmma_loop<NACC><<<grid, threads>>>(out, 1000, 1.0, 1.0);
    dfma_loop<NACC><<<grid, threads>>>(out, 1000, 1.0, 1.0);
    CK(cudaDeviceSynchronize());

    double best_dmma = 0, best_dfma = 0, sust_dmma = 0;
    for (int rep = 0; rep < 5; ++rep) {
        CK(cudaEventRecord(e0));
        dmma_loop<NACC><<<grid, threads>>>(out, iters, 1.0, 1.0);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        double flops = 2.0 * 256.0 * NACC * (double)iters * (threads / 32) * grid;
        double tf = flops / (ms * 1e-3) / 1e12;
        if (tf > best_dmma) best_dmma = tf;
    }
    {   // sustained: ~2 s of back-to-back launches
        int launches = 0;
        CK(cudaEventRecord(e0));
        for (int rep = 0; rep < 60; ++rep) { dmma_loop<NACC><<<grid, threads>>>(out, iters * 4, 1.0, 1.0); ++launches; }
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        double flops = 2.0 * 256.0 * NACC * (double)iters * 4 * (threads / 32) * grid * launches;
        sust_dmma = flops / (ms * 1e-3) / 1e12;
    }
    for (int rep = 0; rep < 5; ++rep) {
        CK(cudaEventRecord(e0));
        dfma_loop<NACC><<<grid, threads>>>(out, iters * 4, 1.0000001, 1e-9);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        double flops = 2.0 * NACC * (double)iters * 4 * threads * grid;
        double tf = flops / (ms * 1e-3) / 1e12;
        if (tf > best_dfma) best_dfma = tf;
    }
    printf("{\"gpu\": \"%s\", \"sms\": %d, \"dmma884_tflops_burst\": %.3f, \"dmma884_tflops_sustained\": %.3f, "
           "\"dfma_tflops_burst\": %.3f}\n", prop.name, sms, best_dmma, sust_dmma, best_dfma);
    return 0;
}
