// Dev micro-benchmark: FP64 DFMA dependent-issue latency and throughput on the current GPU.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dfma_lat dfma_lat.cu && ./dfma_lat
#include <cstdio>
#include <cuda_runtime.h>
template <int CH>
__global__ void k(double* out, int iters, double b, double c) {
    double a[CH];
#pragma unroll
    for (int i = 0; i < CH; ++i) a[i] = threadIdx.x * 1e-9 + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 16; ++u)
#pragma unroll
            for (int i = 0; i < CH; ++i) a[i] = fma(a[i], b, c);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < CH; ++i) s += a[i];
    if (s == 123.456) out[0] = s;
}
template <int CH>
void run(int warps_per_sm, int sms, double clk_ghz) {
    double* d;
    cudaMalloc(&d, 8);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    int iters = 20000;
    int threads = warps_per_sm * 32;
    int blocks = sms;
    if (threads > 1024) { blocks = sms * (threads / 1024); threads = 1024; }
    k<CH><<<blocks, threads>>>(d, 100, 1.0000001, 1e-7);
    cudaEventRecord(e0);
    k<CH><<<blocks, threads>>>(d, iters, 1.0000001, 1e-7);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    double inst_per_warp = (double)iters * 16 * CH;
    double cycles = ms * 1e-3 * clk_ghz * 1e9;
    double tf = 2.0 * inst_per_warp * 32 * warps_per_sm * sms / (ms * 1e-3) / 1e12;
    printf("chains %d warps/SM %2d : %.2f cycles per DFMA per warp, %.2f cycles per dependent step, %.1f TFLOP/s\n", CH,
           warps_per_sm, cycles / inst_per_warp, cycles / (iters * 16.0), tf);
    cudaFree(d);
}
int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    int sms = p.multiProcessorCount;
    double clk = 1.965;
    for (int w : {4, 8, 12, 16, 32, 64}) { run<1>(w, sms, clk); }
    for (int w : {4, 8, 12, 16}) { run<2>(w, sms, clk); }
    for (int w : {4, 8, 12, 16}) { run<4>(w, sms, clk); }
    for (int w : {4, 12}) { run<8>(w, sms, clk); }
    return 0;
}
