// Shared host/device helpers for the pats_b200 CUDA library (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#ifdef __cplusplus
#include <atomic>
#endif
#include <math.h>
#include <stdint.h>

#include "../../include/pats_b200.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "pats_b200 kernels are written for sm_100a (B200); build with -gencode arch=compute_100a,code=sm_100a"
#endif

#define PATS_API extern "C" __attribute__((visibility("default")))

namespace pats {

// ---- host side -------------------------------------------------------------------------------
void set_error(const char *fmt, ...);
int cuda_fail(cudaError_t e, const char *what);
int invalid(const char *fmt, ...);
int sm_count();                 // of the CURRENT device (cached per device ordinal)
// Every piece of cached state in the library is per DEVICE: kernel attributes (cudaFuncSetAttribute applies to the current
// device only), scratch buffers, flag pools and counters live in memory of the device they were made on.  The ordinal of the
// current device, or -1 (error set) beyond kMaxDevices.
constexpr int kMaxDevices = 64;
int current_device();
// "has this (kernel, attribute) been configured on the current device?"  One 64-bit mask per call site: bit = device ordinal.
struct PerDeviceOnce {
    unsigned long long mask = 0ull;
    bool done(int dev) const { return dev >= 0 && ((__atomic_load_n(&mask, __ATOMIC_ACQUIRE) >> dev) & 1ull); }
    void mark(int dev) {
        if (dev >= 0) __atomic_fetch_or(&mask, 1ull << dev, __ATOMIC_RELEASE);
    }
};

#define PATS_CUDA_TRY(expr)                                            \
    do {                                                               \
        cudaError_t _e = (expr);                                       \
        if (_e != cudaSuccess) return ::pats::cuda_fail(_e, #expr);    \
    } while (0)

#define PATS_LAUNCH_CHECK(name)                                        \
    do {                                                               \
        cudaError_t _e = cudaGetLastError();                           \
        if (_e != cudaSuccess) return ::pats::cuda_fail(_e, name);     \
    } while (0)

static inline cudaStream_t as_stream(void *s) { return reinterpret_cast<cudaStream_t>(s); }

// ---- device side -----------------------------------------------------------------------------
#ifdef __CUDACC__

constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;

// exp / log on the SFU (MUFU.EX2 / MUFU.LG2): 2 ulp class, ample for the 1e-4 log-domain tolerance.  The .ftz forms
// are spelled out: exp2f() / __log2f() without them wrap the MUFU in a denormal range test (FSETP + two predicated
// FMULs per call), which tripled the cost of the first (log-domain) iteration.  Results below 2^-126 flush to zero.
__device__ __forceinline__ float fast_exp2(float x) {
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float fast_log2(float x) {
    float r;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
#ifdef PATS_AB_LEGACY_EXP
__device__ __forceinline__ float fast_exp(float x) { return exp2f(x * kLog2e); }
__device__ __forceinline__ float fast_log(float x) { return __log2f(x) * kLn2; }
#else
__device__ __forceinline__ float fast_exp(float x) { return fast_exp2(x * kLog2e); }
__device__ __forceinline__ float fast_log(float x) { return fast_log2(x) * kLn2; }
#endif
__device__ __forceinline__ float fast_rcp(float x) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

// Packed FP32 pair arithmetic (Blackwell FFMA2 / FADD2 / FMUL2): one issue slot for two lanes of work.
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
    unsigned long long ra = *reinterpret_cast<unsigned long long *>(&a);
    unsigned long long rb = *reinterpret_cast<unsigned long long *>(&b);
    unsigned long long rc = *reinterpret_cast<unsigned long long *>(&c);
    unsigned long long rd;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(ra), "l"(rb), "l"(rc));
    return *reinterpret_cast<float2 *>(&rd);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// ---- plan hand-over between the Sinkhorn kernels and the kernels that consume their plans (see sinkhorn_common.cuh) ----
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
// ---- launch chaining: every kernel of the path is launched with programmatic stream serialization (launch_chained()
// below) and begins with pdl_prologue(): wait until the grids it depends on have completed and flushed -- exactly the
// ordering a plain launch gives -- THEN allow the next kernel in the stream to become resident.  Nothing runs early;
// what disappears is the drain -> launch -> ramp-up gap between the 16 small dependent kernels of a step (~2 us each).
// Wait-before-trigger keeps the dependency transitive: a grid is resident only after everything before it has finished.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_prologue() {
    pdl_wait();
    pdl_launch_dependents();
}
// consumer side: one thread spins, then the caller synchronises its group.  The spin is bounded: after ~2 s (a producer slowed
// down by a debugger / sanitizer / time-slicing, or a huge batch in the log-domain fallback) the thread stops polling and
// executes griddepcontrol.wait, which returns when the producer GRID has completed and flushed -- always correct, merely
// without the overlap.  Nothing traps: a slow producer must never cost the CUDA context.
__device__ __forceinline__ void await_problem(const unsigned *done, unsigned epoch, int p) {
    const long long t0 = clock64();
    unsigned v;
    do {
        asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(done + p) : "memory");
        if (v != epoch && clock64() - t0 > (4ll << 30)) {
            pdl_wait();
            return;
        }
    } while (v != epoch);
}

#endif  // __CUDACC__

// host side (sinkhorn.cu): log_optimal_transport2 with per-problem "plan complete" flags for a consumer launched right
// behind it (the composite entry points of regroup.cu).  *done == nullptr on return: the kernel for this shape does not
// publish -- launch the consumer in plain stream order.  edge_add: see SinkArgs::edge_add.
int sinkhorn_ot2_publish(const float *scores, const float *one, const float *ns, int b, int m, int n, int iters, float edge_add,
                         float *out, cudaStream_t st, const unsigned **done, unsigned *epoch);
// Process-wide switches (A/B timing and tests).  Atomics: a call samples each switch once, on entry; flipping one from another host
// thread never tears, it just applies to that thread's later calls.  They select between bit-identical code paths.
extern std::atomic<int> g_handover;  // pats_plan_handover(): 0 = never publish (plain stream order everywhere)
extern std::atomic<int> g_chain;     // pats_launch_chaining(): 0 = plain launches

// launch with programmatic stream serialization (see pdl_prologue); the kernel MUST start with pdl_prologue() (or, for
// the hand-over consumers, wait on the per-problem flags instead)
template <class... KArgs, class... Args>
inline cudaError_t launch_chained(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args &&...args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = g_chain.load(std::memory_order_relaxed) ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

}  // namespace pats
