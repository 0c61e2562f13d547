// SFU (MUFU.EX2 / MUFU.LG2 / MUFU.RCP) throughput on the device at hand -- BASELINE.md section 3 listed it as "assumed 16
// results / clk / SM, must be micro-benchmarked"; the register-resident Sinkhorn kernels spend their first iteration on it and
// the streaming kernel one ex2 per element per iteration.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/microbench/sfu_bench tools/microbench/sfu_bench.cu
//   tools/microbench/sfu_bench            # prints one JSON object
//
// Each thread runs ILP independent dependency chains of the instruction; all warp slots of every SM are filled (2048 threads per SM).
#include <cstdio>
#include <cuda_runtime.h>

template <int OP>
__device__ __forceinline__ float op(float x) {
    float r;
    if (OP == 0) asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    else if (OP == 1) asm volatile("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    else if (OP == 2) asm volatile("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    else asm volatile("fma.rn.f32 %0, %1, %1, %1;" : "=f"(r) : "f"(x));  // FFMA reference
    return r;
}

template <int OP, int ILP>
__global__ void __launch_bounds__(1024) chain(float *out, int iters, float seed) {
    float v[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) v[i] = seed + 1e-3f * (threadIdx.x + i);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) v[i] = op<OP>(v[i]);
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += v[i];
    if (s == 123.456f) out[0] = s;  // keeps the chains alive
}

template <int OP>
double run(int sms, float clock_ghz, float *d_out, double *per_clk_sm) {
    constexpr int ILP = 8;
    const int iters = 4096;
    chain<OP, ILP><<<sms * 2, 1024>>>(d_out, 64, 0.5f);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0), cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 5; ++rep) {
        cudaEventRecord(e0);
        chain<OP, ILP><<<sms * 2, 1024>>>(d_out, iters, 0.5f);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    const double n = (double)sms * 2 * 1024 * ILP * iters;
    const double per_s = n / (best * 1e-3);
    *per_clk_sm = per_s / (clock_ghz * 1e9) / sms;
    return per_s;
}

int main() {
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, 0);
    int clk_khz = 0;
    cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    const float ghz = clk_khz * 1e-6f;
    float *d;
    cudaMalloc(&d, 4);
    const char *names[4] = {"ex2_approx_ftz", "lg2_approx_ftz", "rcp_approx_ftz", "ffma"};
    double rate[4], pcs[4];
    rate[0] = run<0>(prop.multiProcessorCount, ghz, d, &pcs[0]);
    rate[1] = run<1>(prop.multiProcessorCount, ghz, d, &pcs[1]);
    rate[2] = run<2>(prop.multiProcessorCount, ghz, d, &pcs[2]);
    rate[3] = run<3>(prop.multiProcessorCount, ghz, d, &pcs[3]);
    printf("{\"device\": \"%s\", \"sms\": %d, \"clock_ghz_nominal\": %.3f", prop.name, prop.multiProcessorCount, ghz);
    for (int i = 0; i < 4; ++i) printf(", \"%s\": {\"results_per_s\": %.4g, \"results_per_clk_per_sm_at_nominal_clock\": %.2f}", names[i], rate[i], pcs[i]);
    printf("}\n");
    return cudaGetLastError() != cudaSuccess;
}
