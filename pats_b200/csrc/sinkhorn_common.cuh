// Shared device helpers of the Sinkhorn kernels (sinkhorn.cu, sinkhorn_grid.cu): problem descriptor, the virtual
// dustbin augmentation of models/modules.py:152-159 / :175-178, and the exact log-domain solver.
#pragma once

#include "common.cuh"

namespace pats {

enum { MODE_RAW = 0, MODE_OT = 1, MODE_OT2 = 2 };

struct SinkArgs {
    const float *Z;       // raw/OT2: [b,M,N]; OT: un-augmented scores [b,M-1,N-1]
    const float *log_mu;  // raw only [b,M]
    const float *log_nu;  // raw only [b,N]
    const float *alpha;   // OT: dustbin score (device scalar); OT2: `one` (device scalar)
    const float *ns;      // OT/OT2: [b,N-1] target areas
    float *out;           // [b,M,N]
    int b, M, N, iters, mode;
    int *fb_total;        // device counter: problems sent to the log-domain fallback
    // Producer side of the early-start hand-over to the kernel that consumes the plans (est_position / Compute_result):
    // done[p] = epoch once problem p's plan is complete in memory (nullptr: off).  See "plan hand-over" below.
    unsigned *done;
    unsigned epoch;
    // added to the dustbin column, then to the dustbin row, of the finished plan (second_layer.py:108-112 does this in
    // place right after the solve; the corner gets it twice).  Only the 145 x 145 kernel and the log-domain solver apply
    // it in their epilogue; run_sinkhorn() covers the other kernels with a separate pass.
    float edge_add;
    int fp_exit;  // 65 x 65 kernel: leave the iteration loop at a bitwise fixed point of beta (bit-identical results; fb_total[1] counts the skipped iterations)
};

struct Marg {
    float norm, lms, lnsum, fill;
};

__device__ __forceinline__ float z_at(const SinkArgs &a, const Marg &g, int p, int row, int col) {
    if (a.mode == MODE_OT) {
        const int zm = a.M - 1, zn = a.N - 1;  // couplings = [[scores, alpha],[alpha, alpha]]  (modules.py:152-156)
        return (row < zm && col < zn) ? __ldg(a.Z + ((size_t)p * zm + row) * zn + col) : g.fill;
    }
    return __ldg(a.Z + ((size_t)p * a.M + row) * a.N + col);
}

__device__ __forceinline__ float lmu_at(const SinkArgs &a, const Marg &g, int p, int row) {
    if (a.mode == MODE_RAW) return __ldg(a.log_mu + (size_t)p * a.M + row);
    return row < a.M - 1 ? g.norm : g.lnsum + g.norm;  // modules.py:159 / :178
}

__device__ __forceinline__ float lnu_at(const SinkArgs &a, const Marg &g, int p, int col) {
    if (a.mode == MODE_RAW) return __ldg(a.log_nu + (size_t)p * a.N + col);
    return col < a.N - 1 ? logf(__ldg(a.ns + (size_t)p * (a.N - 1) + col)) + g.norm : g.lms + g.norm;  // :158 / :177
}

// Per-problem scalars of the marginals (modules.py:157 / :175).  Every warp computes them
// redundantly in the same order, so all warps of a CTA hold bit-identical values.
__device__ __forceinline__ Marg problem_marginals(const SinkArgs &a, int p, int lane) {
    Marg g;
    g.norm = 0.f, g.lms = 0.f, g.lnsum = 0.f, g.fill = 0.f;
    if (a.mode != MODE_RAW) {
        const int nr = a.N - 1;
        float s = 0.f;
        for (int j = lane; j < nr; j += 32) s += __ldg(a.ns + (size_t)p * nr + j);
        s = warp_sum(s);
        const float sc = __ldg(a.alpha);
        const float ms = (a.mode == MODE_OT2) ? (float)(a.M - 1) * sc : (float)(a.M - 1);
        g.norm = -logf(ms + s);
        g.lms = logf(ms);
        g.lnsum = logf(s);
        g.fill = (a.mode == MODE_OT) ? sc : 0.f;
    }
    return g;
}

// Branch-free view of one problem's plan for the fixed-shape kernels: the core (rows < M-1, columns < N-1) is always
// in memory (row stride N-1 for log_optimal_transport's un-augmented scores, N otherwise); only the dustbin row /
// column entries depend on the mode (virtual `fill` for MODE_OT).  z_at() tests the mode per element, which put every
// load of the 64-element tiles into its own basic block (~8 instructions each).
struct PlanRef {
    const float *base;
    int stride;
    bool aug;
    float fill;
    __device__ __forceinline__ float core(int row_off, int col) const { return __ldg(base + (row_off + col)); }  // row_off = row * stride
    __device__ __forceinline__ float edge(int row, int col) const { return aug ? fill : __ldg(base + (row * stride + col)); }
};
__device__ __forceinline__ PlanRef plan_ref(const SinkArgs &a, const Marg &g, int p) {
    PlanRef r;
    r.aug = a.mode == MODE_OT;
    r.stride = r.aug ? a.N - 1 : a.N;
    r.base = a.Z + (size_t)p * (size_t)(r.aug ? (a.M - 1) * (a.N - 1) : a.M * a.N);
    r.fill = g.fill;
    return r;
}

// identity the compiler cannot see through: stops common-subexpression reuse across the iteration loop
__device__ __forceinline__ int opaque(int v) {
    asm volatile("" : "+r"(v));
    return v;
}

// ---- plan hand-over (programmatic dependent launch) ------------------------------------------------------------------
// The level-2 / level-3 solves end in a tail: 300 problems on 296 CTA slots, 4800 on 1184.  The consumer of the plans
// (area expansion, third-layer result) works per problem, so it is launched with programmatic stream serialization:
// its CTAs start as soon as every producer CTA has STARTED (griddepcontrol.launch_dependents at the top of the
// producer), wait for "their" problem's flag and overlap the producer's tail.  One thread of the producer publishes a
// problem after a barrier that follows all of its stores (fence + release store); the consumer acquires the flag and
// reads the plan with ld.global.cg.  Without a consumer the flags cost one store per problem.
__device__ __forceinline__ void publish_problem(const SinkArgs &a, int p) {  // one thread, after the problem's last store + barrier
    if (a.done) {
        __threadfence();
        asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(a.done + p), "r"(a.epoch) : "memory");
    }
}
struct BlockSync {
    __device__ __forceinline__ void operator()() const { __syncthreads(); }
};
struct WarpSync {
    __device__ __forceinline__ void operator()() const { __syncwarp(); }
};

// ---------------------------------------------------------------------------------------------
// Exact log-domain Sinkhorn by a group of GT threads (modules.py:137-143 as written: max-shifted
// log-sum-exp for rows, then columns).  u[M], v[N] live in shared memory; `red` holds 2*GT floats.
// Used (a) as the in-kernel fallback of the register kernels and (b) as the kernel for shapes
// that do not fit in registers.
// ---------------------------------------------------------------------------------------------
template <int GT, class Sync>
__device__ void log_domain_solve(const SinkArgs &a, const Marg &g, int p, float *u, float *v, float *red, int gtid,
                                 Sync gsync) {
    const int M = a.M, N = a.N;
    const int lane = gtid & 31, warp = gtid >> 5;
    constexpr int NW = GT / 32;
    for (int j = gtid; j < N; j += GT) v[j] = 0.f;
    for (int i = gtid; i < M; i += GT) u[i] = 0.f;
    // column pass geometry: NP columns (padded to a warp multiple) x G row groups
    const int NP = (N + 31) & ~31;
    const int G = (NP <= GT) ? GT / NP : 1;
    gsync();
    for (int it = 0; it < a.iters; ++it) {
        for (int i = warp; i < M; i += NW) {
            float mx = -INFINITY;
            for (int j = lane; j < N; j += 32) mx = fmaxf(mx, z_at(a, g, p, i, j) + v[j]);
            mx = warp_max(mx);
            const float mxs = (fabsf(mx) == INFINITY) ? 0.f : mx;
            float s = 0.f;
            for (int j = lane; j < N; j += 32) s += expf((z_at(a, g, p, i, j) + v[j]) - mxs);
            s = warp_sum(s);
            if (lane == 0) u[i] = lmu_at(a, g, p, i) - (logf(s) + mxs);
        }
        gsync();
        if (G > 1) {
            const int jj = gtid % NP, gi = gtid / NP;
            float mx = -INFINITY, s = 0.f;
            if (gi < G && jj < N) {
                for (int i = gi; i < M; i += G) mx = fmaxf(mx, z_at(a, g, p, i, jj) + u[i]);
                const float mxs = (fabsf(mx) == INFINITY) ? 0.f : mx;
                for (int i = gi; i < M; i += G) s += expf((z_at(a, g, p, i, jj) + u[i]) - mxs);
                red[2 * (gi * NP + jj)] = mx;
                red[2 * (gi * NP + jj) + 1] = s;
            }
            gsync();
            if (gi == 0 && jj < N) {
                float m2 = -INFINITY;
                for (int q = 0; q < G; ++q) m2 = fmaxf(m2, red[2 * (q * NP + jj)]);
                const float m2s = (fabsf(m2) == INFINITY) ? 0.f : m2;
                float s2 = 0.f;
                for (int q = 0; q < G; ++q) {
                    const float mq = red[2 * (q * NP + jj)];
                    const float mqs = (fabsf(mq) == INFINITY) ? 0.f : mq;
                    s2 += red[2 * (q * NP + jj) + 1] * expf(mqs - m2s);
                }
                v[jj] = lnu_at(a, g, p, jj) - (logf(s2) + m2s);
            }
        } else {
            for (int j = gtid; j < N; j += GT) {
                float mx = -INFINITY;
                for (int i = 0; i < M; ++i) mx = fmaxf(mx, z_at(a, g, p, i, j) + u[i]);
                const float mxs = (fabsf(mx) == INFINITY) ? 0.f : mx;
                float s = 0.f;
                for (int i = 0; i < M; ++i) s += expf((z_at(a, g, p, i, j) + u[i]) - mxs);
                v[j] = lnu_at(a, g, p, j) - (logf(s) + mxs);
            }
        }
        gsync();
    }
    const float shift = (a.mode == MODE_RAW) ? 0.f : g.norm;  // modules.py:161 / :181
    float *o = a.out + (size_t)p * M * N;
    for (int e = gtid; e < M * N; e += GT) {
        const int i = e / N, j = e - i * N;
        float r = ((z_at(a, g, p, i, j) + u[i]) + v[j]) - shift;
        if (a.edge_add != 0.f) {
            if (j == N - 1) r += a.edge_add;
            if (i == M - 1) r += a.edge_add;
        }
        o[e] = r;
    }
}

template <int GT>
__global__ void __launch_bounds__(GT) sinkhorn_generic_kernel(SinkArgs a) {
    extern __shared__ float sm[];
    float *u = sm, *v = sm + a.M, *red = sm + a.M + a.N;
    for (int p = blockIdx.x; p < a.b; p += gridDim.x) {
        const Marg g = problem_marginals(a, p, threadIdx.x & 31);
        log_domain_solve<GT>(a, g, p, u, v, red, threadIdx.x, BlockSync());
        __syncthreads();
    }
}

__device__ __forceinline__ float finite_or_zero(float m) { return (fabsf(m) == INFINITY) ? 0.f : m; }

// sinkhorn_grid.cu: problems larger than 512 x 512, rows split over co-resident CTAs (cooperative launch)
bool grid_plan_supported(int M, int N);
int launch_grid(const SinkArgs &a, cudaStream_t st);
// sinkhorn.cu: one CTA per problem, exact log-domain iteration from global memory (any shape)
int launch_generic(const SinkArgs &a, cudaStream_t st);

}  // namespace pats
