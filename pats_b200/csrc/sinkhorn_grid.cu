// Sinkhorn for plans that do not fit one thread-block cluster (M or N > 512: the 1025 x 1025 level-1 plan of a
// 1024 x 1024 image pair, the synthetic N = 1536 / 4096 cases of BASELINE.json) -- models/modules.py:137-182.
//
// The plan is STREAMED: this is the HBM-roofline kernel of the family (B_alg = b * 4 * M * N * (iters + 2) bytes).
//   * A problem is split by rows over G co-resident CTAs (cooperative launch; 8 warps x 2 CTAs per SM up to 1537
//     columns, G = floor(296 / b) but at least one row per warp); a warp owns whole rows, lane l the columns l + 32c.  One pass over a row does BOTH half-iterations: it loads Z_i (coalesced 128-B
//     warp loads, all of a row's loads in flight at once), forms K_ij = exp(Z_ij + u1_i + v1_j) in registers, reduces
//     r_i = sum_j K_ij beta_j with one warp butterfly, alpha_i = mu_i / r_i, and accumulates K_ij * alpha_i into the
//     lane's per-column registers.  The plan is read ONCE per iteration; nothing but two N-vectors is written.
//   * Column sums cross warps through shared memory and cross CTAs through an L2-resident [G][N] partial array:
//     barrier, each CTA finishes an N/G column slice (beta_j = nu_j / sum), barrier, everyone reloads beta.  The
//     barriers are per-problem arrival counters (red.release / ld.acquire at gpu scope).
//   * Iteration 1 is exact in the log domain (row log-sum-exp, then column max and column sum passes), as in the
//     register-resident kernels; scalings are monitored and a problem that leaves [1e-13, 1e13] is flagged and
//     re-solved by the log-domain kernel after this one.  No CPU path.
//   Measured on B200 (tools/ab_grid4096.py): b = 32, 1537 x 1537, 100 iterations: 4.94 ms = 6.2 TB/s algorithmic =
//   95 % of the measured HBM (copy) peak (5.5 ms before the exchange's loads were batched; the one-CTA-per-problem log-domain
//   kernel needs > 1 s); 1025 x 1025, b = 1: 0.64 ms; 4097 x 4097, b = 1, 200 iterations: 3.36 ms (sinkhorn_gridq_kernel below).
//   Variants that lost the A/B and were removed: register prefetch of the next row across the exchange (register
//   pressure), a one-barrier exchange where every CTA sums all partials, 16 warps x 1 CTA per SM (kept as a hook).
#include <atomic>
#include <map>
#include <mutex>

#include "sinkhorn_common.cuh"

namespace pats {

namespace {

struct GridArgs {
    SinkArgs s;
    int G;            // CTAs per problem
    int groups;       // problems in flight (grid = G * groups)
    int rpc;          // rows per CTA
    int slice;        // columns finished per CTA
    int npad;         // stride of the per-problem N-vectors in the workspace (floats)
    float *v1;        // [b][npad]
    float *beta;      // [b][npad]
    float *cref;      // [b][npad]   column maxima of iteration 1
    float *part;      // [b][G][npad] per-CTA column partials
    unsigned *bar;    // [b] arrival counters (zeroed before launch)
    unsigned *flag;   // [b] 1 = scalings left the safe range -> log-domain re-solve
    int l2_prefetch;  // 1: the batch of plans does not stay in L2 between iterations -> prefetch each warp's next row into L2
};

__device__ __forceinline__ void red_release_add(unsigned *p, unsigned v) {
    asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire(const unsigned *p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
// Barrier over the G CTAs of one problem.  Writes before it (st.global by any thread of any of the CTAs) are visible
// to __ldcg loads after it: bar.sync orders the CTA's writes before thread 0's release, the acquire orders the
// other CTAs' writes before the second bar.sync.
__device__ __forceinline__ void group_barrier(unsigned *ctr, unsigned &epoch, unsigned G) {
    epoch += G;
    __syncthreads();
    if (threadIdx.x == 0) {
        red_release_add(ctr, 1u);
        while (ld_acquire(ctr) < epoch) {
        }
    }
    __syncthreads();
}

// My column slice [c0, c1) of the G per-CTA partials of one problem: fin(j, total).  Few CTAs: a thread per column (coalesced
// over j); many CTAs: a warp per column, lanes over the partials (lane-strided running sum, then the butterfly: a fixed order, so
// the sum is deterministic).  The partial loads are L2 round trips: a warp issues those of TWO columns, up to eight per lane each,
// before it uses the first -- issued one by one (5 per lane at 148 CTAs, 10 at 296) they were a quarter of an iteration of a
// single large plan.
template <class Fin>
__device__ __forceinline__ void combine_slice(const float *partg, int npad, unsigned G, bool is_max, int c0, int c1, int w, int W, int lane,
                                              int tid, int T, Fin &&fin) {
#ifdef PATS_AB_SERIAL_COMBINE
    if (G >= 16u) {
        for (int j = c0 + w; j < c1; j += W) {
            float t = is_max ? -INFINITY : 0.f;
            for (unsigned q = lane; q < G; q += 32) {
                const float x = __ldcg(partg + (size_t)q * npad + j);
                t = is_max ? fmaxf(t, x) : t + x;
            }
            t = is_max ? warp_max(t) : warp_sum(t);
            if (lane == 0) fin(j, t);
        }
    } else
#endif
    if (G >= 16u) {
        const float idn = is_max ? -INFINITY : 0.f;
        for (int j = c0 + w; j < c1; j += 2 * W) {
            const int j2 = j + W;
            const bool two = j2 < c1;
            float t0 = idn, t1 = idn;
            for (unsigned q0 = 0; q0 < G; q0 += 256u) {
                float x0[8], x1[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const unsigned qq = q0 + lane + 32u * u;
                    x0[u] = qq < G ? __ldcg(partg + (size_t)qq * npad + j) : idn;
                    x1[u] = (two && qq < G) ? __ldcg(partg + (size_t)qq * npad + j2) : idn;
                }
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    t0 = is_max ? fmaxf(t0, x0[u]) : t0 + x0[u];
                    t1 = is_max ? fmaxf(t1, x1[u]) : t1 + x1[u];
                }
            }
            t0 = is_max ? warp_max(t0) : warp_sum(t0);
            t1 = is_max ? warp_max(t1) : warp_sum(t1);
            if (lane == 0) {
                fin(j, t0);
                if (two) fin(j2, t1);
            }
        }
    } else {
        for (int j = c0 + tid; j < c1; j += T) {
            float t = __ldcg(partg + j);
#pragma unroll 8
            for (unsigned q = 1; q < G; ++q) {  // unrolled: the partial loads of a column are issued together
                const float x = __ldcg(partg + (size_t)q * npad + j);
                t = is_max ? fmaxf(t, x) : t + x;
            }
            fin(j, t);
        }
    }
}

// dst (shared memory)[0, n) <- src (global, written by other CTAs before the barrier)[0, n): eight loads per thread in flight
template <int T>
__device__ __forceinline__ void reload_vector(float *dst, const float *src, int n, int tid) {
#ifdef PATS_AB_SERIAL_RELOAD
    for (int j = tid; j < n; j += T) dst[j] = __ldcg(src + j);
    return;
#endif
    for (int j0 = 0; j0 < n; j0 += 8 * T) {
        float x[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int j = j0 + tid + u * T;
            x[u] = j < n ? __ldcg(src + j) : 0.f;
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int j = j0 + tid + u * T;
            if (j < n) dst[j] = x[u];
        }
    }
}

template <int CPL, int W>
struct GridSmem {
    static constexpr int NC = 32 * CPL;  // padded core width
    // floats: v1[NC+1] beta[NC+1] colbuf[W][NC] last[W] + 3 * rpc (u1, mu, alpha)
    static size_t bytes(int rpc) { return sizeof(float) * ((size_t)2 * (NC + 4) + (size_t)W * NC + 32 + 3 * (size_t)rpc); }
};

// element (row, col) of the virtually augmented plan through a row base pointer (nullptr: the whole row is `fill`)
struct RowRef {
    const float *base;  // first core element of the row (always a readable row; ignored when is_fill)
    bool is_fill;       // log_optimal_transport's dustbin row: every core element is `fill`
    float fill;
    float last;         // value of the last column
};

// One sweep over a row: the loads of a chunk of CH column slots are all issued before the first use (unconditional,
// clamped addresses; padding and the virtual dustbin row are patched in afterwards), so a warp keeps CH x 128 B in
// flight.  f(c, z) is called for every column slot c < CPL with compile-time c after unrolling.
template <int CPL, int CH, class F>
__device__ __forceinline__ void for_row(const RowRef &rr, int lane, int NC, F &&f) {
    static_assert(CPL % CH == 0, "the chunks must tile the row");
#pragma unroll
    for (int ch = 0; ch < CPL / CH; ++ch) {
        float z[CH];
#pragma unroll
        for (int c = 0; c < CH; ++c) z[c] = __ldg(rr.base + min(lane + 32 * (ch * CH + c), NC - 1));
#pragma unroll
        for (int c = 0; c < CH; ++c) {
            const int cc = ch * CH + c;
            f(cc, (lane + 32 * cc < NC) ? (rr.is_fill ? rr.fill : z[c]) : -INFINITY);
        }
    }
}

// Pull a row this warp will sweep a little later from HBM into L2 with ONE instruction and no registers
// (cp.async.bulk.prefetch.L2): the sweep's own loads then see L2 latency instead of DRAM latency, which is what limits a
// warp that can only keep one row's loads in flight (a second row of loads in registers did not fit).  Needs a 16-byte
// aligned row; other shapes simply skip it.
__device__ __forceinline__ void prefetch_row_l2(const RowRef &rr, int NC, int lane) {
#ifndef PATS_AB_NO_L2PF
    if (lane == 0 && !rr.is_fill) {
        const unsigned bytes = (unsigned)(NC * sizeof(float)) & ~15u;
        if (bytes && (reinterpret_cast<uintptr_t>(rr.base) & 15) == 0)
            asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(rr.base), "r"(bytes) : "memory");
    }
#endif
}

// The sweep for rows whose CPL column slots are all real columns (NC == 32 * CPL): no clamps, immediate offsets; the
// virtual dustbin row of log_optimal_transport costs one select per element.
template <int CPL, int CH, class F>
__device__ __forceinline__ void for_row_full(const RowRef &rr, int lane, F &&f) {
    const float *zp = rr.base + lane;
#pragma unroll
    for (int ch = 0; ch < CPL / CH; ++ch) {
        float z[CH];
#pragma unroll
        for (int c = 0; c < CH; ++c) z[c] = __ldg(zp + 32 * (ch * CH + c));
#pragma unroll
        for (int c = 0; c < CH; ++c) f(ch * CH + c, rr.is_fill ? rr.fill : z[c]);
    }
}

__device__ __forceinline__ RowRef row_ref(const SinkArgs &a, const Marg &g, int p, int row) {
    RowRef r;
    r.fill = g.fill;
    if (a.mode == MODE_OT) {
        const int zm = a.M - 1, zn = a.N - 1;
        r.is_fill = row >= zm;
        r.base = a.Z + ((size_t)p * zm + (r.is_fill ? 0 : row)) * zn;
        r.last = g.fill;
    } else {
        r.is_fill = false;
        r.base = a.Z + ((size_t)p * a.M + row) * a.N;
        r.last = __ldg(r.base + a.N - 1);
    }
    return r;
}

__device__ __forceinline__ RowRef row_ref_base_only(const SinkArgs &a, int p, int row) {  // base / is_fill only (for the prefetch)
    RowRef r;
    r.fill = 0.f, r.last = 0.f;
    if (a.mode == MODE_OT) {
        const int zm = a.M - 1, zn = a.N - 1;
        r.is_fill = row >= zm;
        r.base = a.Z + ((size_t)p * zm + (r.is_fill ? 0 : row)) * zn;
    } else {
        r.is_fill = false;
        r.base = a.Z + ((size_t)p * a.M + row) * a.N;
    }
    return r;
}

// FULLONLY (used for the recompute variant, whose 128 column accumulators leave no registers for two sweep flavours in one
// instantiation): the host guarantees NC == 32 * CPL, and the main loop is compiled with the clamp-free sweep only.
template <int CPL, int W, int OCC, bool KEEP, bool FULLONLY = false>
__global__ void __launch_bounds__(W * 32, OCC) sinkhorn_grid_kernel(GridArgs ga) {
    using S = GridSmem<CPL, W>;
    constexpr int NCP = S::NC, T = W * 32;
    constexpr int CH = KEEP ? CPL : 32;  // column slots loaded per batch; !KEEP: rows are re-swept instead of kept
    static_assert(CPL % CH == 0, "chunking");
    extern __shared__ float sm[];
    float *v1s = sm;                    // [NCP + 4]  (index NCP = last column)
    float *bes = v1s + NCP + 4;         // [NCP + 4]
    float *colbuf = bes + NCP + 4;      // [W][NCP]
    float *lastbuf = colbuf + W * NCP;  // [32]
    float *u1s = lastbuf + 32;          // [rpc]
    float *mus = u1s + ga.rpc;          // [rpc]
    float *als = mus + ga.rpc;          // [rpc]
    const SinkArgs &a = ga.s;
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int M = a.M, N = a.N, NC = N - 1;  // NC real core columns + the last column
    const unsigned G = (unsigned)ga.G;
    const int grp = blockIdx.x / ga.G, gi = blockIdx.x - grp * ga.G;
    const int r0 = gi * ga.rpc, r1 = min(M, r0 + ga.rpc);
    const int c0 = gi * ga.slice, c1 = min(N, c0 + ga.slice);  // my column slice (global index; N-1 = last column)

    // cross-warp reduction of the per-lane column registers, then the CTA's partial goes to part[p][gi][.]
    auto publish = [&](float (&acc)[CPL], float acc_last, float *dst, bool is_max) {
#pragma unroll
        for (int c = 0; c < CPL; ++c) colbuf[w * NCP + lane + 32 * c] = acc[c];
        if (lane == 0) lastbuf[w] = acc_last;
        __syncthreads();
        for (int j = tid; j < NC; j += T) {
            float t = colbuf[j];
#pragma unroll
            for (int q = 1; q < W; ++q) t = is_max ? fmaxf(t, colbuf[q * NCP + j]) : t + colbuf[q * NCP + j];
            __stcg(dst + j, t);
        }
        if (tid == 0) {
            float t = lastbuf[0];
#pragma unroll
            for (int q = 1; q < W; ++q) t = is_max ? fmaxf(t, lastbuf[q]) : t + lastbuf[q];
            __stcg(dst + NC, t);
        }
    };

    auto combine = [&](const float *partg, bool is_max, auto &&fin) {
        combine_slice(partg, ga.npad, G, is_max, c0, c1, w, W, lane, tid, T, fin);
    };

    for (int p = grp; p < a.b; p += ga.groups) {
        const Marg g = problem_marginals(a, p, lane);
        float *v1g = ga.v1 + (size_t)p * ga.npad, *beg = ga.beta + (size_t)p * ga.npad, *crg = ga.cref + (size_t)p * ga.npad;
        float *partg = ga.part + (size_t)p * ga.G * ga.npad, *mypart = partg + (size_t)gi * ga.npad;
        unsigned *ctr = ga.bar + p;
        unsigned epoch = 0;
        const float shift = (a.mode == MODE_RAW) ? 0.f : g.norm;
        __syncthreads();
        for (int i = r0 + tid; i < r1; i += T) {
            mus[i - r0] = expf(lmu_at(a, g, p, i));
            u1s[i - r0] = 0.f;
            als[i - r0] = 1.f;
        }
        for (int j = tid; j < NCP + 4; j += T) v1s[j] = 0.f, bes[j] = (j < NC || j == NCP) ? 1.f : 0.f;
        __syncthreads();

        if (a.iters >= 1) {
            // ---- iteration 1: u1 = log_mu - LSE_j Z ; column maxima of Z + u1 ---------------------------------------------
            float cm[CPL], cml = -INFINITY;
#pragma unroll
            for (int c = 0; c < CPL; ++c) cm[c] = -INFINITY;
            for (int i = r0 + w; i < r1; i += W) {
                const RowRef rr = row_ref(a, g, p, i);
                float zk[KEEP ? CPL : 1];
                float mx = rr.last;
                for_row<CPL, CH>(rr, lane, NC, [&](int c, float zz) {
                    if (KEEP) zk[c] = zz;
                    mx = fmaxf(mx, zz);
                });
                mx = finite_or_zero(warp_max(mx));
                float sacc = 0.f;
                if constexpr (KEEP) {
#pragma unroll
                    for (int c = 0; c < CPL; ++c) sacc += fast_exp(zk[c] - mx);
                } else {
                    for_row<CPL, CH>(rr, lane, NC, [&](int, float zz) { sacc += fast_exp(zz - mx); });
                }
                sacc = warp_sum(sacc) + fast_exp(rr.last - mx);
                const float u = lmu_at(a, g, p, i) - (fast_log(sacc) + mx);
                if (lane == 0) u1s[i - r0] = u;
                if constexpr (KEEP) {
#pragma unroll
                    for (int c = 0; c < CPL; ++c) cm[c] = fmaxf(cm[c], zk[c] + u);
                } else {
                    for_row<CPL, CH>(rr, lane, NC, [&](int c, float zz) { cm[c] = fmaxf(cm[c], zz + u); });
                }
                cml = fmaxf(cml, rr.last + u);
            }
            publish(cm, cml, mypart, true);
            group_barrier(ctr, epoch, G);
            combine(partg, true, [&](int j, float t) { __stcg(crg + j, finite_or_zero(t)); });
            group_barrier(ctr, epoch, G);
            reload_vector<T>(bes, crg, NC, tid);  // bes holds the column reference for this pass
            if (tid == 0) bes[NCP] = __ldcg(crg + NC);
            __syncthreads();
            // ---- v1 = log_nu - LSE_i (Z + u1): column sums of exp(Z + u1 - cref) -----------------------------------------
            float cs[CPL], csl = 0.f;
#pragma unroll
            for (int c = 0; c < CPL; ++c) cs[c] = 0.f;
            const float crl = bes[NCP];
            for (int i = r0 + w; i < r1; i += W) {
                const RowRef rr = row_ref(a, g, p, i);
                const float u = u1s[i - r0];
                for_row<CPL, CH>(rr, lane, NC, [&](int c, float zz) { cs[c] += fast_exp((zz + u) - bes[lane + 32 * c]); });
                csl += fast_exp((rr.last + u) - crl);
            }
            __syncthreads();
            publish(cs, csl, mypart, false);
            group_barrier(ctr, epoch, G);
            combine(partg, false, [&](int j, float t) {
                __stcg(v1g + j, lnu_at(a, g, p, j) - (fast_log(t) + __ldcg(crg + j)));
                __stcg(beg + j, 1.f);
            });
            group_barrier(ctr, epoch, G);
            reload_vector<T>(v1s, v1g, NC, tid);
            for (int j = tid; j < NC; j += T) bes[j] = 1.f;
            if (tid == 0) v1s[NCP] = __ldcg(v1g + NC), bes[NCP] = 1.f;
            __syncthreads();
        }

        // ---- iterations 2..iters: one pass over the plan per iteration -------------------------------------------------
        float lo = INFINITY, hi = 0.f;
        float k[KEEP ? CPL : 1];  // KEEP: one row of exp(Z + u1 + v1), kept between the row sum and the column accumulation
        const int ifirst = r0 + w;
        const bool full_rows = NC == 32 * CPL;
        for (int it = 1; it < a.iters; ++it) {
            float cacc[CPL], cl = 0.f;
#pragma unroll
            for (int c = 0; c < CPL; ++c) cacc[c] = 0.f;
            const float v1l = v1s[NCP], bel = bes[NCP];
            const bool check = (it & 7) == 0 || it == a.iters - 1;
            for (int i = ifirst; i < r1; i += W) {
                const RowRef rr = row_ref(a, g, p, i);
                if (ga.l2_prefetch && i + W < r1) prefetch_row_l2(row_ref_base_only(a, p, i + W), NC, lane);
                const float u = u1s[i - r0];
                float rsum = 0.f;
                if constexpr (KEEP) {
#ifndef PATS_AB_NO_FULLROW
                    if (full_rows && !rr.is_fill) {
                        // every column slot of every lane is a real column (NC == 32 * CPL: 512, 1024, 1536, 2048): no index
                        // clamps, no padding selects, one base pointer with immediate offsets, exponent in base 2 with the
                        // row's potential folded into the FFMA -- 8 instructions per element instead of ~19 (ncu: the sweep
                        // was half issue-bound: 57 % issue slots, 43 % ALU pipe, at 54 % of DRAM throughput)
                        const float *zp = rr.base + lane;
                        const float u2 = u * kLog2e;
#pragma unroll
                        for (int c = 0; c < CPL; ++c) k[c] = __ldg(zp + 32 * c);
#pragma unroll
                        for (int c = 0; c < CPL; ++c) {
                            const int j = lane + 32 * c;
                            const float kk = fast_exp2(fmaf(k[c] + v1s[j], kLog2e, u2));
                            k[c] = kk;
                            rsum = fmaf(kk, bes[j], rsum);
                        }
                    } else
#endif
                    for_row<CPL, CH>(rr, lane, NC, [&](int c, float zz) {
                        const int j = lane + 32 * c;
                        const float kk = fast_exp((zz + u) + v1s[j]);
                        k[c] = kk;
                        rsum = fmaf(kk, bes[j], rsum);
                    });
                } else if constexpr (FULLONLY) {
                    const float u2 = u * kLog2e;
                    for_row_full<CPL, CH>(rr, lane, [&](int c, float zz) {
                        const int j = lane + 32 * c;
                        rsum = fmaf(fast_exp2(fmaf(zz + v1s[j], kLog2e, u2)), bes[j], rsum);
                    });
                } else {
                    for_row<CPL, CH>(rr, lane, NC, [&](int c, float zz) {
                        const int j = lane + 32 * c;
                        rsum = fmaf(fast_exp((zz + u) + v1s[j]), bes[j], rsum);
                    });
                }
                const float kl = fast_exp((rr.last + u) + v1l);
                rsum = fmaf(kl, bel, warp_sum(rsum));
                const float al = mus[i - r0] * fast_rcp(rsum);
                if (lane == 0) als[i - r0] = al;
                if (check) lo = fminf(lo, al), hi = fmaxf(hi, al);
                if constexpr (KEEP) {
#pragma unroll
                    for (int c = 0; c < CPL; ++c) cacc[c] = fmaf(k[c], al, cacc[c]);
                } else if constexpr (FULLONLY) {
                    const float u2 = u * kLog2e;  // the same expression as in the row sweep: bit-identical exponentials
                    for_row_full<CPL, CH>(rr, lane, [&](int c, float zz) {
                        cacc[c] = fmaf(fast_exp2(fmaf(zz + v1s[lane + 32 * c], kLog2e, u2)), al, cacc[c]);
                    });
                } else {
                    for_row<CPL, CH>(rr, lane, NC, [&](int c, float zz) {
                        cacc[c] = fmaf(fast_exp((zz + u) + v1s[lane + 32 * c]), al, cacc[c]);
                    });
                }
                cl = fmaf(kl, al, cl);
            }
            __syncthreads();  // colbuf of the previous iteration has been read by every thread (barrier inside publish)
            publish(cacc, cl, mypart, false);
            group_barrier(ctr, epoch, G);
            combine(partg, false, [&](int j, float t) {
                const float be = expf(lnu_at(a, g, p, j)) * fast_rcp(t);
                if (check) lo = fminf(lo, be), hi = fmaxf(hi, be);
                __stcg(beg + j, be);
            });
            group_barrier(ctr, epoch, G);
            reload_vector<T>(bes, beg, NC, tid);
            if (tid == 0) bes[NCP] = __ldcg(beg + NC);
            __syncthreads();
        }

        // ---- health verdict (any CTA of the problem may raise the flag), then the output pass -----------------------------
        const bool bad = !(lo >= 1e-13f && hi <= 1e13f) && a.iters >= 2;
        if (__syncthreads_or(bad ? 1 : 0)) {
            if (tid == 0) atomicExch(ga.flag + p, 1u);
        }
        group_barrier(ctr, epoch, G);
        const bool flagged = *reinterpret_cast<volatile unsigned *>(ga.flag + p) != 0u;
        if (!flagged) {
            bool nonfinite = false;
            float *o = a.out + (size_t)p * M * N;
            const float Vl = (v1s[NCP] + (a.iters >= 2 ? fast_log(bes[NCP]) : 0.f)) - shift;
            for (int j = tid; j < NC; j += T) {  // bes <- V = v1 + ln beta - norm
                const float V = (v1s[j] + (a.iters >= 2 ? fast_log(bes[j]) : 0.f)) - shift;
                if (!(fabsf(V) < INFINITY)) nonfinite = true;
                colbuf[j] = V;
            }
            if (!(fabsf(Vl) < INFINITY)) nonfinite = true;
            __syncthreads();
            for (int i = r0 + w; i < r1; i += W) {
                const RowRef rr = row_ref(a, g, p, i);
                const float U = u1s[i - r0] + (a.iters >= 2 ? fast_log(als[i - r0]) : 0.f);
                if (!(fabsf(U) < INFINITY)) nonfinite = true;
                float *orow = o + (size_t)i * N;
                for_row<CPL, 16>(rr, lane, NC, [&](int c, float zz) {
                    const int j = lane + 32 * c;
                    if (j < NC) orow[j] = (zz + U) + colbuf[j];
                });
                if (lane == 0) orow[NC] = (rr.last + U) + Vl;
            }
            // non-finite potentials (e.g. an all -inf row): the log-domain kernel reproduces the reference's result
            if (__syncthreads_or(nonfinite ? 1 : 0)) {
                if (tid == 0) atomicExch(ga.flag + p, 1u);
            }
        }
        __syncthreads();
    }
}

// ---- rows split over Q warps (core width 32 * CPL * Q exactly: BASELINE.json's N = 4096 stress size) ----------------------------
// At 4096 columns a warp that owns whole rows needs 128 column accumulators per lane: no room to keep the row's exponentials
// (they were recomputed in a second sweep), 32 loads in flight per warp, 8 warps per SM -- latency-bound at 28 us per iteration
// against ~10 us of data.  Here a row belongs to a QUAD of warps: warp q owns columns [1024 q, 1024 q + 1024), 32 per lane, so the
// exponentials stay in registers (one sweep, one ex2 per element), all 32 loads of a warp are in flight at once and 16 warps
// fit an SM.  The row sum (and, in the log-domain first iteration, the row maximum) crosses the four warps through a
// double-buffered shared-memory slot and one named barrier per reduction; the column side (publish / combine / the two
// problem-wide barriers per iteration) is the scheme of sinkhorn_grid_kernel.  Every warp of a quad derives alpha_i from the same
// four partials in the same order, so the quad agrees bit for bit.
template <int CPL, int Q, int RG>
struct GridQSmem {
    static constexpr int NC = 32 * CPL * Q;
    // floats: v1[NC+4] beta[NC+4] colbuf[RG][NC] last[32] xbuf[2][Q*RG <= 32] + 3 * rpc (u1, mu, alpha)
    static size_t bytes(int rpc) { return sizeof(float) * ((size_t)2 * (NC + 4) + (size_t)RG * NC + 32 + 64 + 3 * (size_t)rpc); }
};

template <int CPL, int Q, int RG>
__global__ void __launch_bounds__(Q * RG * 32, 1) sinkhorn_gridq_kernel(GridArgs ga) {
    constexpr int W = Q * RG, T = W * 32, NCP = 32 * CPL * Q, QC = 32 * CPL;
    static_assert(W <= 32, "xbuf holds one slot per warp");
    extern __shared__ float sm[];
    float *v1s = sm;                     // [NCP + 4]  (index NCP = last column)
    float *bes = v1s + NCP + 4;          // [NCP + 4]
    float *colbuf = bes + NCP + 4;       // [RG][NCP]
    float *lastbuf = colbuf + RG * NCP;  // [32]
    float *xbuf = lastbuf + 32;          // [2][32]
    float *u1s = xbuf + 64;              // [rpc]
    float *mus = u1s + ga.rpc;           // [rpc]
    float *als = mus + ga.rpc;           // [rpc]
    const SinkArgs &a = ga.s;
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int rg = w / Q, q = w - rg * Q;  // row group of the CTA, column quarter of the row
    const int jb = q * QC + lane;          // this lane's first column; its columns are jb + 32 c
    const int M = a.M, N = a.N, NC = N - 1;  // host: NC == NCP
    const unsigned G = (unsigned)ga.G;
    const int grp = blockIdx.x / ga.G, gi = blockIdx.x - grp * ga.G;
    const int r0 = gi * ga.rpc, r1 = min(M, r0 + ga.rpc);
    const int c0 = gi * ga.slice, c1 = min(N, c0 + ga.slice);
    unsigned xc = 0;  // reductions done by this quad so far (selects the slot)

    // value reduced over the Q warps of the row (x is warp-uniform); the same four operands in the same order in every warp
    auto quad_reduce = [&](float x, bool is_max) -> float {
        float *xb = xbuf + (xc & 1u) * 32 + rg * Q;
        ++xc;
        if (lane == 0) xb[q] = x;
        asm volatile("bar.sync %0, %1;" ::"r"(1 + rg), "n"(Q * 32) : "memory");
        float t = xb[0];
#pragma unroll
        for (int i = 1; i < Q; ++i) t = is_max ? fmaxf(t, xb[i]) : t + xb[i];
        return t;
    };

    auto publish = [&](float (&acc)[CPL], float acc_last, float *dst, bool is_max) {
#pragma unroll
        for (int c = 0; c < CPL; ++c) colbuf[rg * NCP + jb + 32 * c] = acc[c];
        if (lane == 0 && q == 0) lastbuf[rg] = acc_last;
        __syncthreads();
        for (int j = tid; j < NCP; j += T) {
            float t = colbuf[j];
#pragma unroll
            for (int r = 1; r < RG; ++r) t = is_max ? fmaxf(t, colbuf[r * NCP + j]) : t + colbuf[r * NCP + j];
            __stcg(dst + j, t);
        }
        if (tid == 0) {
            float t = lastbuf[0];
#pragma unroll
            for (int r = 1; r < RG; ++r) t = is_max ? fmaxf(t, lastbuf[r]) : t + lastbuf[r];
            __stcg(dst + NCP, t);
        }
    };

    auto combine = [&](const float *partg, bool is_max, auto &&fin) {
        combine_slice(partg, ga.npad, G, is_max, c0, c1, w, W, lane, tid, T, fin);
    };

    // this lane's CPL elements of a row (the virtual dustbin row of log_optimal_transport is `fill`)
    auto load_row = [&](const RowRef &rr, float (&z)[CPL]) {
        const float *zp = rr.base + jb;
#pragma unroll
        for (int c = 0; c < CPL; ++c) z[c] = __ldg(zp + 32 * c);
        if (rr.is_fill) {
#pragma unroll
            for (int c = 0; c < CPL; ++c) z[c] = rr.fill;
        }
    };

    for (int p = grp; p < a.b; p += ga.groups) {
        const Marg g = problem_marginals(a, p, lane);
        float *v1g = ga.v1 + (size_t)p * ga.npad, *beg = ga.beta + (size_t)p * ga.npad, *crg = ga.cref + (size_t)p * ga.npad;
        float *partg = ga.part + (size_t)p * ga.G * ga.npad, *mypart = partg + (size_t)gi * ga.npad;
        unsigned *ctr = ga.bar + p;
        unsigned epoch = 0;
        const float shift = (a.mode == MODE_RAW) ? 0.f : g.norm;
        __syncthreads();
        for (int i = r0 + tid; i < r1; i += T) {
            mus[i - r0] = expf(lmu_at(a, g, p, i));
            u1s[i - r0] = 0.f;
            als[i - r0] = 1.f;
        }
        for (int j = tid; j < NCP + 4; j += T) v1s[j] = 0.f, bes[j] = (j <= NCP) ? 1.f : 0.f;
        __syncthreads();

        if (a.iters >= 1) {
            // ---- iteration 1, exact in the log domain: u1 = log_mu - LSE_j Z ; column maxima of Z + u1 ----------------------
            float cm[CPL], cml = -INFINITY;
#pragma unroll
            for (int c = 0; c < CPL; ++c) cm[c] = -INFINITY;
            for (int i = r0 + rg; i < r1; i += RG) {
                const RowRef rr = row_ref(a, g, p, i);
                float zk[CPL];
                load_row(rr, zk);
                float mx = zk[0];
#pragma unroll
                for (int c = 1; c < CPL; ++c) mx = fmaxf(mx, zk[c]);
                mx = finite_or_zero(fmaxf(quad_reduce(warp_max(mx), true), rr.last));
                float sacc = 0.f;
#pragma unroll
                for (int c = 0; c < CPL; ++c) sacc += fast_exp(zk[c] - mx);
                sacc = quad_reduce(warp_sum(sacc), false) + fast_exp(rr.last - mx);
                const float u = lmu_at(a, g, p, i) - (fast_log(sacc) + mx);
                if (lane == 0 && q == 0) u1s[i - r0] = u;
#pragma unroll
                for (int c = 0; c < CPL; ++c) cm[c] = fmaxf(cm[c], zk[c] + u);
                cml = fmaxf(cml, rr.last + u);
            }
            publish(cm, cml, mypart, true);
            group_barrier(ctr, epoch, G);
            combine(partg, true, [&](int j, float t) { __stcg(crg + j, finite_or_zero(t)); });
            group_barrier(ctr, epoch, G);
            reload_vector<T>(bes, crg, NCP + 1, tid);  // bes holds the column reference for this pass
            __syncthreads();
            // ---- v1 = log_nu - LSE_i (Z + u1): column sums of exp(Z + u1 - cref) -------------------------------------------
            float cs[CPL], csl = 0.f;
#pragma unroll
            for (int c = 0; c < CPL; ++c) cs[c] = 0.f;
            const float crl = bes[NCP];
            for (int i = r0 + rg; i < r1; i += RG) {
                const RowRef rr = row_ref(a, g, p, i);
                const float u = u1s[i - r0];
                float zk[CPL];
                load_row(rr, zk);
#pragma unroll
                for (int c = 0; c < CPL; ++c) cs[c] += fast_exp((zk[c] + u) - bes[jb + 32 * c]);
                csl += fast_exp((rr.last + u) - crl);
            }
            __syncthreads();
            publish(cs, csl, mypart, false);
            group_barrier(ctr, epoch, G);
            combine(partg, false, [&](int j, float t) {
                __stcg(v1g + j, lnu_at(a, g, p, j) - (fast_log(t) + __ldcg(crg + j)));
                __stcg(beg + j, 1.f);
            });
            group_barrier(ctr, epoch, G);
            reload_vector<T>(v1s, v1g, NCP + 1, tid);
            for (int j = tid; j <= NCP; j += T) bes[j] = 1.f;
            __syncthreads();
        }

        // ---- iterations 2..iters: one sweep per iteration, the row's exponentials kept in registers ----------------------------
        float lo = INFINITY, hi = 0.f;
        for (int it = 1; it < a.iters; ++it) {
            float cacc[CPL], cl = 0.f;
#pragma unroll
            for (int c = 0; c < CPL; ++c) cacc[c] = 0.f;
            const float v1l = v1s[NCP], bel = bes[NCP];
            const bool check = (it & 7) == 0 || it == a.iters - 1;
            for (int i = r0 + rg; i < r1; i += RG) {
                const RowRef rr = row_ref(a, g, p, i);
                if (ga.l2_prefetch && q == 0 && i + RG < r1) prefetch_row_l2(row_ref_base_only(a, p, i + RG), NC, lane);
                const float u = u1s[i - r0];
                const float u2 = u * kLog2e;
                float k[CPL];
                load_row(rr, k);
                float rsum = 0.f;
#pragma unroll
                for (int c = 0; c < CPL; ++c) {
                    const int j = jb + 32 * c;
                    const float kk = fast_exp2(fmaf(k[c] + v1s[j], kLog2e, u2));
                    k[c] = kk;
                    rsum = fmaf(kk, bes[j], rsum);
                }
                const float kl = fast_exp((rr.last + u) + v1l);
                rsum = fmaf(kl, bel, quad_reduce(warp_sum(rsum), false));
                const float al = mus[i - r0] * fast_rcp(rsum);
                if (lane == 0 && q == 0) als[i - r0] = al;
                if (check) lo = fminf(lo, al), hi = fmaxf(hi, al);
#pragma unroll
                for (int c = 0; c < CPL; ++c) cacc[c] = fmaf(k[c], al, cacc[c]);
                cl = fmaf(kl, al, cl);
            }
            __syncthreads();  // colbuf of the previous iteration has been read by every thread (barrier inside publish)
            publish(cacc, cl, mypart, false);
            group_barrier(ctr, epoch, G);
            combine(partg, false, [&](int j, float t) {
                const float be = expf(lnu_at(a, g, p, j)) * fast_rcp(t);
                if (check) lo = fminf(lo, be), hi = fmaxf(hi, be);
                __stcg(beg + j, be);
            });
            group_barrier(ctr, epoch, G);
            reload_vector<T>(bes, beg, NCP + 1, tid);
            __syncthreads();
        }

        // ---- health verdict (any CTA of the problem may raise the flag), then the output pass -----------------------------
        const bool bad = !(lo >= 1e-13f && hi <= 1e13f) && a.iters >= 2;
        if (__syncthreads_or(bad ? 1 : 0)) {
            if (tid == 0) atomicExch(ga.flag + p, 1u);
        }
        group_barrier(ctr, epoch, G);
        const bool flagged = *reinterpret_cast<volatile unsigned *>(ga.flag + p) != 0u;
        if (!flagged) {
            bool nonfinite = false;
            float *o = a.out + (size_t)p * M * N;
            const float Vl = (v1s[NCP] + (a.iters >= 2 ? fast_log(bes[NCP]) : 0.f)) - shift;
            for (int j = tid; j < NCP; j += T) {  // colbuf <- V = v1 + ln beta - norm
                const float V = (v1s[j] + (a.iters >= 2 ? fast_log(bes[j]) : 0.f)) - shift;
                if (!(fabsf(V) < INFINITY)) nonfinite = true;
                colbuf[j] = V;
            }
            if (!(fabsf(Vl) < INFINITY)) nonfinite = true;
            __syncthreads();
            for (int i = r0 + rg; i < r1; i += RG) {
                const RowRef rr = row_ref(a, g, p, i);
                const float U = u1s[i - r0] + (a.iters >= 2 ? fast_log(als[i - r0]) : 0.f);
                if (!(fabsf(U) < INFINITY)) nonfinite = true;
                float *orow = o + (size_t)i * N;
                float zk[CPL];
                load_row(rr, zk);
#pragma unroll
                for (int c = 0; c < CPL; ++c) orow[jb + 32 * c] = (zk[c] + U) + colbuf[jb + 32 * c];
                if (lane == 0 && q == 0) orow[NCP] = (rr.last + U) + Vl;
            }
            // non-finite potentials (e.g. an all -inf row): the log-domain kernel reproduces the reference's result
            if (__syncthreads_or(nonfinite ? 1 : 0)) {
                if (tid == 0) atomicExch(ga.flag + p, 1u);
            }
        }
        __syncthreads();
    }
}

// log-domain re-solve of the flagged problems (one CTA each; rare)
__global__ void __launch_bounds__(1024) sinkhorn_grid_fallback_kernel(SinkArgs a, const unsigned *flag) {
    extern __shared__ float sm[];
    float *u = sm, *v = sm + a.M, *red = sm + a.M + a.N;
    for (int p = blockIdx.x; p < a.b; p += gridDim.x) {
        if (!flag[p]) continue;
        if (threadIdx.x == 0 && a.fb_total) atomicAdd(a.fb_total, 1);
        const Marg g = problem_marginals(a, p, threadIdx.x & 31);
        log_domain_solve<1024>(a, g, p, u, v, red, threadIdx.x, BlockSync());
        __syncthreads();
    }
}

void *grid_workspace(cudaStream_t st, size_t bytes) {
    static std::mutex mu;
    static std::map<std::pair<int, cudaStream_t>, std::pair<void *, size_t>> pool;  // (device, stream)
    const int dev = current_device();
    if (dev < 0) return nullptr;
    std::lock_guard<std::mutex> lk(mu);
    auto &e = pool[std::make_pair(dev, st)];
    if (e.second < bytes) {
        if (e.first) {
            cudaStreamSynchronize(st);
            cudaFree(e.first);
        }
        const size_t cap = bytes < (1u << 20) ? (1u << 20) : bytes * 2;
        if (cudaMalloc(&e.first, cap) != cudaSuccess) {
            e.first = nullptr, e.second = 0;
            return nullptr;
        }
        e.second = cap;
    }
    return e.first;
}

// flagged problems: exact log-domain iteration (device-side test of the flag, no host sync)
int launch_grid_fallback(const SinkArgs &a, const unsigned *flag, int sms, cudaStream_t st) {
    const size_t fsmem = sizeof(float) * ((size_t)a.M + a.N + 2 * 1024);
    if (fsmem > 200 * 1024) return PATS_OK;  // shapes beyond the log-domain kernel's budget keep the scaling result
    static PerDeviceOnce configured;
    const int dev = current_device();
    if (dev < 0) return PATS_E_CUDA;
    if (!configured.done(dev)) {
        PATS_CUDA_TRY(cudaFuncSetAttribute(sinkhorn_grid_fallback_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        configured.mark(dev);
    }
    sinkhorn_grid_fallback_kernel<<<a.b < sms ? a.b : sms, 1024, fsmem, st>>>(a, flag);
    PATS_LAUNCH_CHECK("sinkhorn_grid_fallback_kernel");
    return PATS_OK;
}

std::atomic<int> g_grid_ctas_per_problem{0};  // test hook: 0 = automatic
std::atomic<int> g_grid_variant{0};            // A/B hook, bits: 1 = 16 warps x 1 CTA per SM instead of 8 warps x 2 CTAs per SM (<= 1536 columns); 2 = 4096 columns: one warp per row

template <int CPL, int W, int OCC, bool KEEP, bool FULLONLY = false>
int launch_grid_cfg(const SinkArgs &a, cudaStream_t st) {
    auto kern = sinkhorn_grid_kernel<CPL, W, OCC, KEEP, FULLONLY>;
    const int sms = sm_count() > 0 ? sm_count() : 148;
    GridArgs ga;
    ga.s = a;
    // CTAs per problem: as many as are co-resident (OCC per SM), but at least one row per warp.  The row block and
    // with it the shared-memory size depend on G, so settle G against the occupancy the smaller block really gets.
    const int gmax = (a.M + W - 1) / W;
    int G = 1, slots = sms;
    size_t smem = 0;
    for (int occ_try = OCC; occ_try >= 1; --occ_try) {
        slots = sms * occ_try;
        G = slots / a.b;
        if (G < 1) G = 1;
        if (G > gmax) G = gmax;
        const int forced = g_grid_ctas_per_problem.load(std::memory_order_relaxed);
        if (forced > 0 && forced <= slots) G = forced < gmax ? forced : gmax;
        smem = GridSmem<CPL, W>::bytes((a.M + G - 1) / G);
        if (smem > 227 * 1024) continue;
        PATS_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        int occ = 0;
        PATS_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, W * 32, smem));
        if (occ >= occ_try) break;
        if (occ_try == 1) return invalid("sinkhorn (grid kernel): %d x %d does not fit an SM (%zu B of shared memory)", a.M, a.N, smem);
    }
    if (smem > 227 * 1024) return invalid("sinkhorn (grid kernel): %d x %d needs %zu B of shared memory", a.M, a.N, smem);
    ga.G = G;
    ga.groups = a.b < slots / G ? a.b : slots / G;
    ga.rpc = (a.M + G - 1) / G;
    ga.slice = (a.N + G - 1) / G;
    ga.npad = (a.N + 3) & ~3;
    ga.l2_prefetch = (size_t)a.b * a.M * a.N * sizeof(float) > ((size_t)64 << 20);  // measured: +3 % at 302 MB, -2.5 % at 4 MB
    const size_t vec = (size_t)a.b * ga.npad;
    const size_t floats = vec * (3 + (size_t)G);
    const size_t bytes = floats * sizeof(float) + 2 * (size_t)a.b * sizeof(unsigned);
    float *ws = static_cast<float *>(grid_workspace(st, bytes));
    if (!ws) return cuda_fail(cudaGetLastError(), "sinkhorn grid workspace");
    ga.v1 = ws, ga.beta = ws + vec, ga.cref = ws + 2 * vec, ga.part = ws + 3 * vec;
    ga.bar = reinterpret_cast<unsigned *>(ws + floats);
    ga.flag = ga.bar + a.b;
    PATS_CUDA_TRY(cudaMemsetAsync(ga.bar, 0, 2 * (size_t)a.b * sizeof(unsigned), st));
    void *params[] = {&ga};
    PATS_CUDA_TRY(cudaLaunchCooperativeKernel((void *)kern, dim3((unsigned)(G * ga.groups)), dim3(W * 32), params, smem, st));
    return launch_grid_fallback(a, ga.flag, sms, st);
}

// rows split over Q warps: core width exactly 32 * CPL * Q columns, one CTA of Q * RG warps per SM
#ifndef PATS_GRIDQ_CPL  // columns per lane, warps per row, row groups per CTA; measured: 32 x 4 x 4 3.49 ms, 16 x 8 x 4 (32 warps, 64 registers) 3.67 ms
#define PATS_GRIDQ_CPL 32
#define PATS_GRIDQ_Q 4
#define PATS_GRIDQ_RG 4
#endif
template <int CPL, int Q, int RG>
int launch_gridq_cfg(const SinkArgs &a, cudaStream_t st) {
    auto kern = sinkhorn_gridq_kernel<CPL, Q, RG>;
    constexpr int W = Q * RG;
    const int sms = sm_count() > 0 ? sm_count() : 148;
    GridArgs ga;
    ga.s = a;
    const int gmax = (a.M + RG - 1) / RG;  // at least one row per row group
    const int slots = sms;
    int G = slots / a.b;
    if (G < 1) G = 1;
    if (G > gmax) G = gmax;
    const int forced = g_grid_ctas_per_problem.load(std::memory_order_relaxed);
    if (forced > 0 && forced <= slots) G = forced < gmax ? forced : gmax;
    const size_t smem = GridQSmem<CPL, Q, RG>::bytes((a.M + G - 1) / G);
    if (smem > 227 * 1024) return invalid("sinkhorn (grid kernel): %d x %d needs %zu B of shared memory", a.M, a.N, smem);
    PATS_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int occ = 0;
    PATS_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, W * 32, smem));
    if (occ < 1) return invalid("sinkhorn (grid kernel): %d x %d does not fit an SM (%zu B of shared memory)", a.M, a.N, smem);
    ga.G = G;
    ga.groups = a.b < slots / G ? a.b : slots / G;
    ga.rpc = (a.M + G - 1) / G;
    ga.slice = (a.N + G - 1) / G;
    ga.npad = (a.N + 3) & ~3;
    // one 4097 x 4097 plan (67 MB) stays in L2 between iterations: prefetching the next row measured 3.49 vs 3.36 ms without
    // (profiles/r02_ab_gridq_configs.json); batches beyond L2 keep the prefetch of the one-warp-per-row kernel
    ga.l2_prefetch = (size_t)a.b * a.M * a.N * sizeof(float) > ((size_t)128 << 20);
    const size_t vec = (size_t)a.b * ga.npad;
    const size_t floats = vec * (3 + (size_t)G);
    const size_t bytes = floats * sizeof(float) + 2 * (size_t)a.b * sizeof(unsigned);
    float *ws = static_cast<float *>(grid_workspace(st, bytes));
    if (!ws) return cuda_fail(cudaGetLastError(), "sinkhorn grid workspace");
    ga.v1 = ws, ga.beta = ws + vec, ga.cref = ws + 2 * vec, ga.part = ws + 3 * vec;
    ga.bar = reinterpret_cast<unsigned *>(ws + floats);
    ga.flag = ga.bar + a.b;
    PATS_CUDA_TRY(cudaMemsetAsync(ga.bar, 0, 2 * (size_t)a.b * sizeof(unsigned), st));
    void *params[] = {&ga};
    PATS_CUDA_TRY(cudaLaunchCooperativeKernel((void *)kern, dim3((unsigned)(G * ga.groups)), dim3(W * 32), params, smem, st));
    return launch_grid_fallback(a, ga.flag, sms, st);
}

}  // namespace

bool grid_plan_supported(int M, int N) { return M >= 2 && N >= 2 && N - 1 <= 4096; }

int launch_grid(const SinkArgs &a, cudaStream_t st) {
    const int nc = a.N - 1;
    if (g_grid_variant & 1) {
        if (nc <= 512) return launch_grid_cfg<16, 16, 1, true>(a, st);
        if (nc <= 1024) return launch_grid_cfg<32, 16, 1, true>(a, st);
        if (nc <= 1536) return launch_grid_cfg<48, 16, 1, true>(a, st);
    }
    if (nc <= 512) return launch_grid_cfg<16, 8, 2, true>(a, st);
    if (nc <= 1024) return launch_grid_cfg<32, 8, 2, true>(a, st);
    if (nc <= 1536) return launch_grid_cfg<48, 8, 2, true>(a, st);
    if (nc <= 2048) return launch_grid_cfg<64, 8, 1, true>(a, st);
    if (nc == 4096)  // BASELINE.json's stress size: four warps per row (exponentials kept); variant 2: one warp per row, exponentials recomputed
        return (g_grid_variant & 2) ? launch_grid_cfg<128, 8, 1, false, true>(a, st) : launch_gridq_cfg<PATS_GRIDQ_CPL, PATS_GRIDQ_Q, PATS_GRIDQ_RG>(a, st);
    return launch_grid_cfg<128, 8, 1, false>(a, st);
}

}  // namespace pats

PATS_API void pats_sinkhorn_grid_ctas_per_problem(int g) { pats::g_grid_ctas_per_problem = g > 0 ? g : 0; }
PATS_API void pats_sinkhorn_grid_variant(int v) { pats::g_grid_variant = v & 3; }
