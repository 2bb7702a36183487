// Micro-benchmark: latency and throughput of mma.sync.m8n8k4.f64 (SASS DMMA.8x8x4) and of plain
// DFMA on one SM / on the whole chip.  Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double &c0, double &c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <int CHAINS>
__global__ void k_dmma(double *out, int iters, long long *cycles) {
    double c0[CHAINS], c1[CHAINS];
    double a = threadIdx.x * 1e-3 + 1.0, b = 1.0 - threadIdx.x * 1e-4;
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) { c0[i] = i; c1[i] = -i; }
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < CHAINS; ++i) dmma(c0[i], c1[i], a, b);
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) s += c0[i] + c1[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}

template <int CHAINS>
__global__ void k_dfma(double *out, int iters, long long *cycles) {
    double c[CHAINS];
    double a = threadIdx.x * 1e-9 + 1.0, b = 1e-9 * threadIdx.x;
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) c[i] = i;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < CHAINS; ++i) c[i] = fma(c[i], a, b);
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) s += c[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}

template <class K>
void run(const char *name, K kern, int chains, int blocks, int threads, int iters, double flop_per_inst_thread) {
    double *out; long long *cyc, h;
    cudaMalloc(&out, sizeof(double) * blocks * threads);
    cudaMalloc(&cyc, sizeof(long long));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    kern<<<blocks, threads>>>(out, 10, cyc);
    cudaEventRecord(e0);
    kern<<<blocks, threads>>>(out, iters, cyc);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    cudaMemcpy(&h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double insts_per_warp = (double)iters * chains;
    double tflops = flop_per_inst_thread * insts_per_warp * blocks * threads / (ms * 1e-3) / 1e12;
    printf("%-10s chains=%d blocks=%4d threads=%4d : %8.2f cycles/inst/warp  %8.3f ms  %8.2f TFLOP/s\n", name, chains, blocks,
           threads, (double)h / insts_per_warp, ms, tflops);
    cudaFree(out); cudaFree(cyc);
}

int main() {
    const int it = 20000;
    // m8n8k4: 8*8*4 FMAs per warp-instruction = 256*2 flop / 32 threads = 16 flop per thread-instruction
    run("dmma", k_dmma<1>, 1, 1, 32, it, 16);     // latency (dependent chain, one warp)
    run("dmma", k_dmma<2>, 2, 1, 32, it, 16);
    run("dmma", k_dmma<4>, 4, 1, 32, it, 16);
    run("dmma", k_dmma<8>, 8, 1, 32, it, 16);
    run("dmma", k_dmma<4>, 4, 1, 128, it, 16);    // one warp per SMSP
    run("dmma", k_dmma<4>, 4, 1, 512, it, 16);
    run("dmma", k_dmma<4>, 4, 148, 512, it, 16);  // whole chip
    run("dmma", k_dmma<4>, 4, 148 * 2, 1024, it, 16);
    run("dfma", k_dfma<1>, 1, 1, 32, it, 2);
    run("dfma", k_dfma<8>, 8, 1, 32, it, 2);
    run("dfma", k_dfma<8>, 8, 1, 128, it, 2);
    run("dfma", k_dfma<8>, 8, 148, 512, it, 2);
    run("dfma", k_dfma<8>, 8, 148 * 2, 1024, it, 2);
    return 0;
}
