// FP64 / FP32 FMA throughput and dependent-issue latency on one GPU (no library needed):
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o dfma_bench scripts/dfma_bench.cu && ./dfma_bench
// B200: 64 DFMA / clk / SM (36.8 TFLOP/s); a single warp with 8 independent chains reaches 15 / clk, i.e. a 17-cycle dependent-issue latency.
#include <cstdio>
#include <cuda_runtime.h>
template <typename T>
__global__ void k(T* out, int iters, T a, T b) {
    T x[8];
    for (int i = 0; i < 8; ++i) x[i] = a * (T)(threadIdx.x + i);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) x[i] = fma(x[i], b, a);
    }
    T s = 0;
    for (int i = 0; i < 8; ++i) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <typename T>
void run(const char* name, int threads, int blocks_per_sm) {
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    T* out; cudaMalloc(&out, sizeof(T) * sms * blocks_per_sm * threads);
    int iters = 20000;
    k<T><<<sms * blocks_per_sm, threads>>>(out, iters, (T)1.0001, (T)0.9999);
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<T><<<sms * blocks_per_sm, threads>>>(out, iters, (T)1.0001, (T)0.9999);
    cudaEventRecord(e1); cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double fma_total = (double)sms * blocks_per_sm * threads * iters * 8;
    printf("%s threads=%d blocks/SM=%d: %.2f T FMA/s = %.1f TFLOP/s  (%.1f FMA/clk/SM at 1.9 GHz)\n", name, threads, blocks_per_sm,
           fma_total / ms / 1e9, 2 * fma_total / ms / 1e9, fma_total / (ms * 1e-3) / sms / 1.9e9);
    cudaFree(out);
}
int main() {
    run<double>("fp64", 256, 4);
    run<double>("fp64", 512, 1);
    run<double>("fp64", 32, 4);
    run<double>("fp64", 32, 1);
    run<float>("fp32", 256, 4);
    return 0;
}
