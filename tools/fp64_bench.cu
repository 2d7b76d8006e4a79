// Microbenchmark: FP64 pipe throughput and dependent-chain latency on the device the chain
// kernels run on.  Output feeds DESIGN.md's FP64 roofline (the sweep is FP64-issue bound, not HBM bound).
#include <cstdio>
#include <cuda_runtime.h>

template <int ILP>
__global__ void k_dfma(double* out, double a, double b, int iters) {
    double x[ILP];
    for (int j = 0; j < ILP; ++j) x[j] = a + j + threadIdx.x;
    for (int i = 0; i < iters; ++i)
#pragma unroll
        for (int j = 0; j < ILP; ++j) x[j] = fma(x[j], b, a);
    double s = 0;
    for (int j = 0; j < ILP; ++j) s += x[j];
    if (s == 12345.678) out[0] = s;
}
__global__ void k_ddiv(double* out, double a, double b, int iters) {
    double x = a + threadIdx.x, y = b + threadIdx.x * 1e-9;
    for (int i = 0; i < iters; ++i) { x = (x * 0.999) / y + 1.0; }
    if (x == 12345.678) out[0] = x;
}
__global__ void k_chain_latency(double* out, long long* cycles, double a, double b, int iters) {
    double x = a;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) x = fma(x, b, a);
    long long t1 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) cycles[0] = t1 - t0;
    if (x == 12345.678) out[0] = x;
}

template <class F>
float timeit(F f) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    f(); cudaDeviceSynchronize();
    cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); return ms;
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int sms = p.multiProcessorCount;
    double* out; long long* cyc; cudaMalloc(&out, 8); cudaMalloc(&cyc, 8);
    printf("{\"device\": \"%s\", \"sms\": %d, \"clock_khz\": %d", p.name, sms, p.clockRate);
    const int iters = 4096;
    for (int warps = 1; warps <= 16; warps *= 2) {
        int threads = 128 * warps > 1024 ? 1024 : 128 * warps;   // warps per SMSP = warps (4 SMSPs)
        int blocks = sms * ((128 * warps + threads - 1) / threads);
        float ms = timeit([&] { k_dfma<1><<<blocks, threads>>>(out, 1.0, 0.999, iters); });
        double n = (double)blocks * threads * iters;
        printf(", \"dfma_ilp1_w%d_gops\": %.1f", warps, n / ms / 1e6);
    }
    {
        int blocks = sms * 2, threads = 1024;
        float ms = timeit([&] { k_dfma<4><<<blocks, threads>>>(out, 1.0, 0.999, iters); });
        printf(", \"dfma_ilp4_gops\": %.1f", (double)blocks * threads * iters * 4 / ms / 1e6);
        ms = timeit([&] { k_ddiv<<<blocks, threads>>>(out, 1.0, 1.001, iters); });
        printf(", \"ddiv_gops\": %.1f", (double)blocks * threads * iters / ms / 1e6);
    }
    k_chain_latency<<<1, 32>>>(out, cyc, 1.0, 0.999, 4096);
    long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    printf(", \"dfma_dependent_latency_cycles\": %.2f}\n", (double)c / 4096);
    return 0;
}
