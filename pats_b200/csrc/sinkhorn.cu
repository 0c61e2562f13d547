// Sinkhorn optimal transport for sm_100a -- replaces models/modules.py:137-182 of zju3dv/pats.
//
//   log_sinkhorn_iterations (:137-143), log_optimal_transport (:145-162), log_optimal_transport2 (:165-182)
//
// Design (see DESIGN.md "Sinkhorn kernels"):
//   * The plan never leaves the chip between iterations.  Each problem is owned by one warp
//     (<= 72 x 68, the level-3 65 x 65 problems), one CTA (<= 160 x 160, the level-2 145 x 145
//     problems) or one thread-block CLUSTER of 8 CTAs (<= 320 x 320 / 512 x 512, the level-1
//     301 x 301 problem; column sums cross CTAs through distributed shared memory); every thread
//     keeps an RT x CT tile of the kernel matrix in REGISTERS.
//   * Iteration 1 is done exactly in the log domain (row / column log-sum-exp, as the reference
//     does).  Its potentials (u1, v1) are absorbed into the matrix, K = exp(Z + u1 + v1), so every
//     entry is a probability <= 1 and the remaining iterations are plain scaling updates
//        alpha_i = mu_i / sum_j K_ij beta_j ,   beta_j = nu_j / sum_i K_ij alpha_i
//     (u first, then v with the new u: modules.py:141-142) -- two packed-FP32 FMAs (FFMA2) per
//     element per iteration instead of two exp + two max passes.  Row sums reduce over the QC
//     lanes that share a row with warp shuffles; column sums reduce over lanes with shuffles and
//     over warps through a few hundred bytes of shared memory.
//   * Result = Z + (u1 + ln alpha) + (v1 + ln beta) - norm, with Z re-read (L2 hit).
//   * The scalings are monitored; a problem whose scalings leave [1e-13, 1e13] or whose potentials
//     are not finite is re-solved by the same threads with the exact log-domain iteration
//     (log_domain_solve), which is also the kernel for shapes that do not fit in registers.
//     No CPU path exists.
#include <cooperative_groups.h>

#include <atomic>
#include <map>
#include <mutex>

#include "sinkhorn_common.cuh"

namespace cg = cooperative_groups;

namespace pats {

// ---------------------------------------------------------------------------------------------
// Register-resident kernel.
//   A problem is owned by WARPS warps.  Threads form a PR x QC grid (QC = 2^QC_LOG2 lanes along a
//   row, PR = WARPS*32/QC row groups); thread (pr,qc) owns rows pr + PR*k (k < RT) and columns
//   qc + QC*c (c < CT).  Capacity MAXM x MAXN = PR*RT x QC*CT; the padding holds K = 1 with zero
//   marginals, which keeps every sum positive and every padded scaling exactly 0 (no NaN, no
//   select in the loop).
// ---------------------------------------------------------------------------------------------
//   With CL > 1 the problem is split by rows over the CL CTAs of a cluster: CTA `rank` owns rows
//   rank*PR*RT + pr + PR*k.  Row sums stay inside a CTA; column sums are pushed into every CTA's shared
//   memory (DSMEM) as a reduce-scatter + broadcast over the cluster: st.async stores that complete bytes on the
//   receiver's mbarrier (no cluster barrier, no MEMBAR in the iteration loop).
template <int WARPS_, int QC_LOG2_, int RT_, int CT_, int GROUPS_, int CL_ = 1, int HOP1_ = 0>
struct RegCfg {
    static constexpr int W = WARPS_, QCL = QC_LOG2_, RT = RT_, CT = CT_, GROUPS = GROUPS_, CL = CL_;
    static constexpr bool HOP1 = HOP1_ != 0;  // cluster exchange: all-to-all broadcast of partials (one DSMEM hop, CL x the traffic)
    static constexpr int GT = W * 32;       // threads per CTA working on the problem
    static constexpr int QC = 1 << QCL;     // lanes along a row
    static constexpr int PRW = 32 / QC;     // row groups per warp
    static constexpr int PR = W * PRW;      // row groups per CTA
    static constexpr int ROWS = PR * RT;    // rows per CTA
    static constexpr int MAXM = CL * ROWS, MAXN = QC * CT;
    static constexpr int CT2 = CT / 2;
    static constexpr bool ODD = (CT & 1) != 0;
    static constexpr int THREADS = GT * GROUPS;
    static constexpr bool BLOCK = (W > 1) || (CL > 1);  // reductions need block-level synchronisation
    static_assert(!BLOCK || GROUPS == 1, "multi-warp problems use __syncthreads: one problem per CTA");
};

// Shared-memory carve-up of one CTA (dynamic shared memory; sizes in floats).
template <class C>
struct Smem {
    static constexpr int MU = 0;                                   // [GROUPS][ROWS]  exp-domain row marginals
    static constexpr int NU = MU + C::GROUPS * C::ROWS;            // [GROUPS][MAXN]
    static constexpr int U1 = NU + C::GROUPS * C::MAXN;            // [GROUPS][ROWS]  potentials of iteration 1
    static constexpr int V1 = U1 + C::GROUPS * C::ROWS;            // [GROUPS][MAXN]
    static constexpr int PART = V1 + C::GROUPS * C::MAXN;          // [W][MAXN]       per-warp column partials
    static constexpr int TOT = PART + (C::BLOCK ? C::W * C::MAXN : 0);   // [MAXN]
    static constexpr int XBUF = TOT + (C::BLOCK ? C::MAXN : 0);          // [2][CL][MAXN/CL] cluster exchange: slice partials
    static constexpr int TOTB = XBUF + (C::CL > 1 ? 2 * C::MAXN * (C::HOP1 ? C::CL : 1) : 0);  // HOP1: XBUF is [2][CL][MAXN]
                                                                         // [2][MAXN]        cluster exchange: finished values
    static constexpr int MBAR = (TOTB + (C::CL > 1 ? 2 * C::MAXN : 0) + 3) & ~3;  // 4 mbarriers (8 B each), 16-B aligned
    static constexpr int XFLAG = MBAR + (C::CL > 1 ? 8 : 0);       // [CL]
    static constexpr int FB = XFLAG + (C::CL > 1 ? C::CL : 0);     // fallback scratch: u[MAXM], v[MAXN], red[2*GT] per group
    static constexpr int FB_PER = C::MAXM + C::MAXN + 2 * C::GT;
    static constexpr int FLOATS = FB + C::GROUPS * FB_PER;
    static constexpr size_t BYTES = sizeof(float) * (size_t)FLOATS;
};

template <class C>
__device__ __forceinline__ float row_allreduce_sum(float v) {
#pragma unroll
    for (int o = 1; o < C::QC; o <<= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
template <class C>
__device__ __forceinline__ float row_allreduce_max(float v) {
#pragma unroll
    for (int o = 1; o < C::QC; o <<= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

struct OpSum {
    __device__ __forceinline__ float operator()(float x, float y) const { return x + y; }
};
struct OpMax {
    __device__ __forceinline__ float operator()(float x, float y) const { return fmaxf(x, y); }
};

// ---- DSMEM exchange primitives: st.async (SASS STAS) completes bytes on the RECEIVER's mbarrier, so the steady state
//      needs no cluster barrier and no MEMBAR (cooperative_groups' cluster.sync() costs MEMBAR.ALL.GPU + CCTL.IVALL). ----
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ unsigned mapa_u32(unsigned addr, unsigned rank) {
    unsigned r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void st_async_f32(unsigned remote_addr, float v, unsigned remote_mbar) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b32 [%0], %1, [%2];" ::"r"(remote_addr),
                 "r"(__float_as_uint(v)), "r"(remote_mbar)
                 : "memory");
}
__device__ __forceinline__ void mbar_init(unsigned mb, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(mb), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned mb, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mb), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned mb, unsigned parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t"
        "}" ::"r"(mb),
        "r"(parity)
        : "memory");
}

// Column all-reduce over every row group of the problem (lanes, then warps through smem, then the CTAs
// of the cluster through DSMEM).  SCALE: the reducing thread turns a total t of column j into
// nu[j] / t (the beta update).  Every CTA sums the same CL partials in the same order, so all CTAs
// hold bit-identical results.
template <class C, bool SCALE, class Op>
__device__ __forceinline__ void col_reduce(float (&v)[C::CT], Op op, float *sm, const float *s_nu, int warp, int prw, int qc,
                                           int gtid, unsigned rank, int &parity) {
#pragma unroll
    for (int c = 0; c < C::CT; ++c) {
#pragma unroll
        for (int o = C::QC; o < 32; o <<= 1) v[c] = op(v[c], __shfl_xor_sync(0xffffffffu, v[c], o));
    }
    if (!C::BLOCK) {
        if (SCALE) {
#pragma unroll
            for (int c = 0; c < C::CT; ++c) v[c] = s_nu[qc + C::QC * c] * fast_rcp(v[c]);
        }
        return;
    }
    float *s_part = sm + Smem<C>::PART, *s_tot = sm + Smem<C>::TOT;
    if (prw == 0) {
#pragma unroll
        for (int c = 0; c < C::CT; ++c) s_part[warp * C::MAXN + qc + C::QC * c] = v[c];
    }
    __syncthreads();
    if (C::CL == 1) {
        for (int j = gtid; j < C::MAXN; j += C::GT) {
            float t = s_part[j];
#pragma unroll
            for (int w = 1; w < C::W; ++w) t = op(t, s_part[w * C::MAXN + j]);
            s_tot[j] = SCALE ? s_nu[j] * fast_rcp(t) : t;
        }
    } else {
        // reduce-scatter + broadcast across the cluster over DSMEM, synchronised by mbarriers (no cluster barrier):
        //   1. every CTA st.async's its partial sums of column slice r to CTA r      -> completes bytes on r's mb_x
        //   2. CTA r waits on mb_x, adds the CL partials of its slice in rank order, applies the scaling and
        //      st.async's the finished slice into every CTA's tot buffer             -> completes bytes on their mb_t
        //   3. every thread waits on mb_t and reads its columns.
        // Buffers and barriers are double-buffered by call parity; a CTA cannot run two calls ahead of any other
        // (step 3 of call n+1 needs every CTA's step 2 of call n+1), which orders all buffer reuse.
        constexpr int SL = C::MAXN / C::CL;
        static_assert(SL * C::CL == C::MAXN, "column slices must tile the padded width");
        const int q = parity & 1;
        const unsigned phase = (unsigned)(parity >> 1) & 1u;
        if (C::HOP1) {
            // One-hop variant: every CTA st.async's its CTA-level partials of ALL columns to every CTA (itself included),
            // each CTA then adds the CL partial rows in rank order itself.  CL x the DSMEM traffic, one hop instead of two.
            // Buffer reuse is ordered as above: a sender reaches call n+2 only after receiving every peer's call n+1
            // partials, which each peer sent after its own reads of call n.
            float *ab = sm + Smem<C>::XBUF + q * C::CL * C::MAXN;  // [CL][MAXN], one row per sender
            const unsigned mb_a = smem_u32(sm + Smem<C>::MBAR) + 16u * q;
            if (gtid == 0) mbar_expect_tx(mb_a, (unsigned)(C::CL * C::MAXN * sizeof(float)));
            for (int j = gtid; j < C::MAXN; j += C::GT) {
                float t = s_part[j];
#pragma unroll
                for (int w = 1; w < C::W; ++w) t = op(t, s_part[w * C::MAXN + j]);
                const unsigned a_loc = smem_u32(ab + rank * C::MAXN + j);
#pragma unroll
                for (unsigned r = 0; r < (unsigned)C::CL; ++r) st_async_f32(mapa_u32(a_loc, r), t, mapa_u32(mb_a, r));
            }
            mbar_wait(mb_a, phase);
            for (int j = gtid; j < C::MAXN; j += C::GT) {
                float t = ab[j];
#pragma unroll
                for (int r = 1; r < C::CL; ++r) t = op(t, ab[r * C::MAXN + j]);
                s_tot[j] = SCALE ? s_nu[j] * fast_rcp(t) : t;
            }
            __syncthreads();
#pragma unroll
            for (int c = 0; c < C::CT; ++c) v[c] = s_tot[qc + C::QC * c];
            parity += 1;
            return;
        }
        float *xb = sm + Smem<C>::XBUF + q * C::MAXN;     // [CL][SL] partials of my slice, one row per sender
        float *tb = sm + Smem<C>::TOTB + q * C::MAXN;     // [MAXN] finished values
        const unsigned mb_x = smem_u32(sm + Smem<C>::MBAR) + 16u * q, mb_t = mb_x + 8u;
        if (gtid == 0) {
            mbar_expect_tx(mb_x, (unsigned)(C::MAXN * sizeof(float)));
            mbar_expect_tx(mb_t, (unsigned)(C::MAXN * sizeof(float)));
        }
        for (int j = gtid; j < C::MAXN; j += C::GT) {
            float t = s_part[j];
#pragma unroll
            for (int w = 1; w < C::W; ++w) t = op(t, s_part[w * C::MAXN + j]);
            const unsigned dst = (unsigned)(j / SL);
            st_async_f32(mapa_u32(smem_u32(xb + rank * SL + (j - (int)dst * SL)), dst), t, mapa_u32(mb_x, dst));
        }
        if (gtid < SL) {
            mbar_wait(mb_x, phase);
            for (int jj = gtid; jj < SL; jj += C::GT) {
                float t = xb[jj];
#pragma unroll
                for (int r = 1; r < C::CL; ++r) t = op(t, xb[r * SL + jj]);
                const int col = (int)rank * SL + jj;
                const float val = SCALE ? s_nu[col] * fast_rcp(t) : t;
                const unsigned a_loc = smem_u32(tb + col);
#pragma unroll
                for (unsigned r = 0; r < (unsigned)C::CL; ++r) st_async_f32(mapa_u32(a_loc, r), val, mapa_u32(mb_t, r));
            }
        }
        mbar_wait(mb_t, phase);
#pragma unroll
        for (int c = 0; c < C::CT; ++c) v[c] = tb[qc + C::QC * c];
        parity += 1;
        return;
    }
    __syncthreads();
#pragma unroll
    for (int c = 0; c < C::CT; ++c) v[c] = s_tot[qc + C::QC * c];
}

template <class C>
__global__ void __launch_bounds__(C::THREADS) sinkhorn_reg_kernel(SinkArgs a) {
    pdl_prologue();
    constexpr int RT = C::RT, CT = C::CT, CT2 = C::CT2, QC = C::QC, PR = C::PR;
    constexpr bool ODD = C::ODD;
    extern __shared__ __align__(16) float sm[];

    const int group = threadIdx.x / C::GT, gtid = threadIdx.x % C::GT;
    const int lane = gtid & 31, warp = gtid >> 5;
    const int qc = lane & (QC - 1), prw = lane >> C::QCL;
    const int pr = warp * C::PRW + prw;
    unsigned rank = 0;
    if (C::CL > 1) rank = cg::this_cluster().block_rank();
    const int p = (C::CL > 1) ? (int)(blockIdx.x / C::CL) : (int)(blockIdx.x * C::GROUPS + group);
    if (p >= a.b) return;  // warp-uniform; block/cluster kernels have GROUPS==1 and exact grids, so nobody is left waiting

    const int M = a.M, N = a.N;
    const int row0 = (int)rank * C::ROWS;  // first row of this CTA's slab
    const Marg g = problem_marginals(a, p, lane);
    float *mu_s = sm + Smem<C>::MU + group * C::ROWS, *nu_s = sm + Smem<C>::NU + group * C::MAXN;
    float *u1_s = sm + Smem<C>::U1 + group * C::ROWS, *v1_s = sm + Smem<C>::V1 + group * C::MAXN;
    int parity = 0;  // number of cluster column reductions done so far (buffer / mbarrier parity and phase)
    if (C::CL > 1) {
        if (gtid == 0) {
            const unsigned mb0 = smem_u32(sm + Smem<C>::MBAR);
#pragma unroll
            for (int i = 0; i < 4; ++i) mbar_init(mb0 + 8u * i, 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        cg::this_cluster().sync();  // once: every CTA's mbarriers exist before any remote completion can arrive
    }

    // ---- load the tile (padding = -inf) --------------------------------------------------------
    float z[RT][CT];
#pragma unroll
    for (int k = 0; k < RT; ++k) {
        const int row = row0 + pr + PR * k;
#pragma unroll
        for (int c = 0; c < CT; ++c) {
            const int col = qc + QC * c;
            z[k][c] = (row < M && col < N) ? z_at(a, g, p, row, col) : -INFINITY;
        }
    }
    // marginals to shared memory (each value written by its qc==0 / pr==0 owner)
    if (qc == 0) {
#pragma unroll
        for (int k = 0; k < RT; ++k) {
            const int row = row0 + pr + PR * k;
            mu_s[pr + PR * k] = (row < M) ? expf(lmu_at(a, g, p, row)) : 0.f;
        }
    }
    if (pr == 0) {
#pragma unroll
        for (int c = 0; c < CT; ++c) {
            const int col = qc + QC * c;
            nu_s[col] = (col < N) ? expf(lnu_at(a, g, p, col)) : 0.f;
        }
    }

    // ---- iteration 1, exact in the log domain; K = exp(Z + u1 + v1) ------------------------------
    if (a.iters >= 1) {
        float u1[RT];
#pragma unroll
        for (int k = 0; k < RT; ++k) {
            const int row = row0 + pr + PR * k;
            float mx = -INFINITY;
#pragma unroll
            for (int c = 0; c < CT; ++c) mx = fmaxf(mx, z[k][c]);
            mx = row_allreduce_max<C>(mx);
            const float mxs = (fabsf(mx) == INFINITY) ? 0.f : mx;
            float s = 0.f;
#pragma unroll
            for (int c = 0; c < CT; ++c) s += fast_exp(z[k][c] - mxs);
            s = row_allreduce_sum<C>(s);
            u1[k] = (row < M) ? lmu_at(a, g, p, row) - (fast_log(s) + mxs) : 0.f;
            if (qc == 0) u1_s[pr + PR * k] = u1[k];
        }
        float cm[CT];
#pragma unroll
        for (int c = 0; c < CT; ++c) {
            float mx = -INFINITY;
#pragma unroll
            for (int k = 0; k < RT; ++k) mx = fmaxf(mx, z[k][c] + u1[k]);
            cm[c] = mx;
        }
        col_reduce<C, false>(cm, OpMax(), sm, nu_s, warp, prw, qc, gtid, rank, parity);
        float cs[CT];
#pragma unroll
        for (int c = 0; c < CT; ++c) {
            cm[c] = (fabsf(cm[c]) == INFINITY) ? 0.f : cm[c];
            float s = 0.f;
#pragma unroll
            for (int k = 0; k < RT; ++k) s += fast_exp((z[k][c] + u1[k]) - cm[c]);
            cs[c] = s;
        }
        col_reduce<C, false>(cs, OpSum(), sm, nu_s, warp, prw, qc, gtid, rank, parity);
#pragma unroll
        for (int c = 0; c < CT; ++c) {
            const int col = qc + QC * c;
            const float v1 = (col < N) ? lnu_at(a, g, p, col) - (fast_log(cs[c]) + cm[c]) : 0.f;
            if (pr == 0) v1_s[col] = v1;
#pragma unroll
            for (int k = 0; k < RT; ++k) {
                const int row = row0 + pr + PR * k;
                z[k][c] = (row < M && col < N) ? fast_exp((z[k][c] + u1[k]) + v1) : 1.0f;
            }
        }
    }
    if (C::BLOCK) __syncthreads(); else __syncwarp();

    // ---- iterations 2..iters: scaling updates on the register tile --------------------------------
    float2 Kp[RT][CT2 > 0 ? CT2 : 1];
    float Kl[RT];
#pragma unroll
    for (int k = 0; k < RT; ++k) {
#pragma unroll
        for (int h = 0; h < CT2; ++h) Kp[k][h] = make_float2(z[k][2 * h], z[k][2 * h + 1]);
        Kl[k] = ODD ? z[k][CT - 1] : 0.f;
    }
    float bcol[CT], al[RT];
#pragma unroll
    for (int c = 0; c < CT; ++c) bcol[c] = (qc + QC * c < N) ? 1.f : 0.f;
#pragma unroll
    for (int k = 0; k < RT; ++k) al[k] = 1.f;
    float lo = INFINITY, hi = 0.f;

    for (int it = 1; it < a.iters; ++it) {
        // alpha_i = mu_i / sum_j K_ij beta_j
        float2 acc[RT];
#pragma unroll
        for (int k = 0; k < RT; ++k) acc[k] = make_float2(0.f, 0.f);
#pragma unroll
        for (int h = 0; h < CT2; ++h) {
            const float2 bp = make_float2(bcol[2 * h], bcol[2 * h + 1]);
#pragma unroll
            for (int k = 0; k < RT; ++k) acc[k] = ffma2(Kp[k][h], bp, acc[k]);
        }
#pragma unroll
        for (int k = 0; k < RT; ++k) {
            float r = acc[k].x + acc[k].y;
            if (ODD) r = fmaf(Kl[k], bcol[CT - 1], r);
            r = row_allreduce_sum<C>(r);
            al[k] = mu_s[pr + PR * k] * fast_rcp(r);
        }
        // beta_j = nu_j / sum_i K_ij alpha_i
        float2 s2[CT2 > 0 ? CT2 : 1];
        float sl = 0.f;
#pragma unroll
        for (int h = 0; h < CT2; ++h) s2[h] = make_float2(0.f, 0.f);
#pragma unroll
        for (int k = 0; k < RT; ++k) {
            const float2 ak = make_float2(al[k], al[k]);
#pragma unroll
            for (int h = 0; h < CT2; ++h) s2[h] = ffma2(Kp[k][h], ak, s2[h]);
            if (ODD) sl = fmaf(Kl[k], al[k], sl);
        }
#pragma unroll
        for (int h = 0; h < CT2; ++h) {
            bcol[2 * h] = s2[h].x;
            bcol[2 * h + 1] = s2[h].y;
        }
        if (ODD) bcol[CT - 1] = sl;
        col_reduce<C, true>(bcol, OpSum(), sm, nu_s, warp, prw, qc, gtid, rank, parity);

        if ((it & 7) == 0 || it == a.iters - 1) {  // range monitor (valid entries only)
#pragma unroll
            for (int k = 0; k < RT; ++k)
                if (row0 + pr + PR * k < M) lo = fminf(lo, al[k]), hi = fmaxf(hi, al[k]);
#pragma unroll
            for (int c = 0; c < CT; ++c)
                if (qc + QC * c < N) lo = fminf(lo, bcol[c]), hi = fmaxf(hi, bcol[c]);
        }
    }

    // ---- potentials, health check, output ---------------------------------------------------------
    const float shift = (a.mode == MODE_RAW) ? 0.f : g.norm;
    float U[RT], V[CT];
    bool bad = !(lo >= 1e-13f && hi <= 1e13f);
#pragma unroll
    for (int k = 0; k < RT; ++k) {
        const int row = row0 + pr + PR * k;
        float t = 0.f;
        if (a.iters >= 1) t = u1_s[pr + PR * k];
        if (a.iters >= 2) t += fast_log(al[k]);
        U[k] = t;
        if (row < M && !(fabsf(t) < INFINITY)) bad = true;
    }
#pragma unroll
    for (int c = 0; c < CT; ++c) {
        const int col = qc + QC * c;
        float t = 0.f;
        if (a.iters >= 1) t = v1_s[col];
        if (a.iters >= 2) t += fast_log(bcol[c]);
        if (col < N && !(fabsf(t) < INFINITY)) bad = true;
        V[c] = t - shift;
    }
    bool any_bad;
    if (!C::BLOCK) {
        any_bad = __any_sync(0xffffffffu, bad) != 0;
    } else {
        any_bad = __syncthreads_or(bad ? 1 : 0) != 0;
        if (C::CL > 1) {  // agree across the cluster
            cg::cluster_group cluster = cg::this_cluster();
            float *xf = sm + Smem<C>::XFLAG;
            if (gtid < C::CL) cluster.map_shared_rank(xf, gtid)[rank] = any_bad ? 1.f : 0.f;
            cluster.sync();
            bool t = false;
#pragma unroll
            for (int r = 0; r < C::CL; ++r) t = t || (xf[r] != 0.f);
            any_bad = t;
        }
    }
    if (!any_bad) {
        float *o = a.out + (size_t)p * M * N;
#pragma unroll
        for (int k = 0; k < RT; ++k) {
            const int row = row0 + pr + PR * k;
#pragma unroll
            for (int c = 0; c < CT; ++c) {
                const int col = qc + QC * c;
                if (row < M && col < N) o[(size_t)row * N + col] = (z_at(a, g, p, row, col) + U[k]) + V[c];
            }
        }
    } else {
        // exact log-domain re-solve by this CTA (rank 0 of a cluster; the other CTAs are done)
        if (rank != 0) return;
        if (gtid == 0 && a.fb_total) atomicAdd(a.fb_total, 1);
        float *fb = sm + Smem<C>::FB + group * Smem<C>::FB_PER;
        if (!C::BLOCK)
            log_domain_solve<C::GT>(a, g, p, fb, fb + C::MAXM, fb + C::MAXM + C::MAXN, gtid, WarpSync());
        else
            log_domain_solve<C::GT>(a, g, p, fb, fb + C::MAXM, fb + C::MAXM + C::MAXN, gtid, BlockSync());
    }
}

// ---------------------------------------------------------------------------------------------
// Level-3 kernel: exactly 65 x 65 (64 x 64 real cells + dustbin row / column), one warp per problem.
//
//   Lane (pr = lane>>2, qc = lane&3) keeps an 8 x 16 tile of the 64 x 64 core (128 registers, no padding).
//   Row slot k of a lane holds logical row pr + 8*(k ^ rmask(qc)); column slot c holds logical column
//   qc + 4*(c ^ cmask(pr)).  With these lane-dependent slot permutations the recursive-halving
//   reduce-scatter ("keep slots [0,h), send slots [h,2h)") and the mirror all-gather need no selects:
//       reduce:  v[t] += shfl_xor(v[t+h], bit)        gather:  v[t+h] = shfl_xor(v[t], bit)
//   After a reduce every lane owns 2 complete row (column) sums, computes 2 scalings (2 MUFU.RCP instead of
//   8 / 16 redundant ones) and the gather hands every lane the 8 (16) scalings of its tile.
//   The dustbin row / column (129 values) are spread 2+2 per lane over the lanes that own the matching
//   row / column sums; their own sums are two 5-step warp all-reduces per iteration, issued early so their
//   latency hides under the FFMA2 stream.
// ---------------------------------------------------------------------------------------------
constexpr int W65_WARPS = 4;  // problems per CTA

template <class Op>
__device__ __forceinline__ void rs_rows_op(float (&v)[8], Op op) {  // reduce-scatter over qc (lane bits 0,1): 8 -> 2
#pragma unroll
    for (int t = 0; t < 4; ++t) v[t] = op(v[t], __shfl_xor_sync(0xffffffffu, v[t + 4], 1));
#pragma unroll
    for (int t = 0; t < 2; ++t) v[t] = op(v[t], __shfl_xor_sync(0xffffffffu, v[t + 2], 2));
}
__device__ __forceinline__ void rs_rows(float (&v)[8]) { rs_rows_op(v, OpSum()); }
__device__ __forceinline__ void ag_rows(float (&v)[8]) {  // all-gather over qc: 2 -> 8
#pragma unroll
    for (int t = 0; t < 2; ++t) v[t + 2] = __shfl_xor_sync(0xffffffffu, v[t], 2);
#pragma unroll
    for (int t = 0; t < 4; ++t) v[t + 4] = __shfl_xor_sync(0xffffffffu, v[t], 1);
}
template <class Op>
__device__ __forceinline__ void rs_cols_op(float (&v)[16], Op op) {  // reduce-scatter over pr (lane bits 2,3,4): 16 -> 2
#pragma unroll
    for (int t = 0; t < 8; ++t) v[t] = op(v[t], __shfl_xor_sync(0xffffffffu, v[t + 8], 4));
#pragma unroll
    for (int t = 0; t < 4; ++t) v[t] = op(v[t], __shfl_xor_sync(0xffffffffu, v[t + 4], 8));
#pragma unroll
    for (int t = 0; t < 2; ++t) v[t] = op(v[t], __shfl_xor_sync(0xffffffffu, v[t + 2], 16));
}
__device__ __forceinline__ void rs_cols(float (&v)[16]) { rs_cols_op(v, OpSum()); }
__device__ __forceinline__ void ag_cols(float (&v)[16]) {  // all-gather over pr: 2 -> 16
#pragma unroll
    for (int t = 0; t < 2; ++t) v[t + 2] = __shfl_xor_sync(0xffffffffu, v[t], 16);
#pragma unroll
    for (int t = 0; t < 4; ++t) v[t + 4] = __shfl_xor_sync(0xffffffffu, v[t], 8);
#pragma unroll
    for (int t = 0; t < 8; ++t) v[t + 8] = __shfl_xor_sync(0xffffffffu, v[t], 4);
}

template <int MIN_CTAS>  // 2: 255 registers, no spills in the loop; 3: 170 registers (3 warps per scheduler), ~19 spill ops per iteration
__global__ void __launch_bounds__(W65_WARPS * 32, MIN_CTAS) sinkhorn_w65_kernel(SinkArgs a) {
    constexpr int D = 64;  // dustbin index; M = N = 65
    __shared__ float s_fb[W65_WARPS][65 + 65 + 64];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int p = blockIdx.x * W65_WARPS + wib;
    if (p >= a.b) return;
    const int pr = lane >> 2, qc = lane & 3;
    const int rmask = ((qc & 1) << 2) | ((qc >> 1) << 1);
    const int cmask = ((pr & 1) << 3) | (((pr >> 1) & 1) << 2) | ((pr >> 2) << 1);
    const Marg g = problem_marginals(a, p, lane);
#define LROW(k) (pr + 8 * ((k) ^ rmask))
#define LCOL(c) (qc + 4 * ((c) ^ cmask))

    // ---- load: core tile; dustbin column / row entries of the 2 rows / 2 columns this lane owns; corner --------------
    // (slots 0,1 are the rows / columns whose complete sums land on this lane after a reduce-scatter; over the warp they
    //  cover all 64 rows / columns exactly once)
    float z[8][16], zc[2], zr[2];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
#pragma unroll
        for (int c = 0; c < 16; ++c) z[k][c] = z_at(a, g, p, LROW(k), LCOL(c));
    }
#pragma unroll
    for (int t = 0; t < 2; ++t) {
        zc[t] = z_at(a, g, p, LROW(t), D);
        zr[t] = z_at(a, g, p, D, LCOL(t));
    }
    const float zcorner = z_at(a, g, p, D, D);

    // owned rows / columns (slots 0,1 after a reduce-scatter) and their marginals
    float mu2[2], nu2[2], u1o[2] = {0.f, 0.f}, v1o[2] = {0.f, 0.f}, u1d = 0.f, v1d = 0.f;
#pragma unroll
    for (int t = 0; t < 2; ++t) {
        mu2[t] = expf(lmu_at(a, g, p, LROW(t)));
        nu2[t] = expf(lnu_at(a, g, p, LCOL(t)));
    }
    const float mud = expf(lmu_at(a, g, p, D)), nud = expf(lnu_at(a, g, p, D));
    float Dc[2] = {0.f, 0.f}, Dr[2] = {0.f, 0.f}, corner = 0.f;

    // ---- iteration 1, exact in the log domain; K = exp(Z + u1 + v1) -----------------------------------------------
    if (a.iters >= 1) {
        // NOTE: slot k of different lanes holds different logical rows, so every cross-lane combination goes through
        // the slot-aware reduce-scatter / all-gather (never a plain same-slot butterfly).
        float u1[8], v1[16];
        {  // rows: u1_i = log mu_i - LSE_j Z_ij
            float mx[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                float m = z[k][0];
#pragma unroll
                for (int c = 1; c < 16; ++c) m = fmaxf(m, z[k][c]);
                mx[k] = m;
            }
            rs_rows_op(mx, OpMax());
#pragma unroll
            for (int t = 0; t < 2; ++t) mx[t] = finite_or_zero(fmaxf(mx[t], zc[t]));
            ag_rows(mx);
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                float sacc = 0.f;
#pragma unroll
                for (int c = 0; c < 16; ++c) sacc += fast_exp(z[k][c] - mx[k]);
                u1[k] = sacc;
            }
            rs_rows(u1);
#pragma unroll
            for (int t = 0; t < 2; ++t) {
                const float sacc = u1[t] + fast_exp(zc[t] - mx[t]);
                u1[t] = lmu_at(a, g, p, LROW(t)) - (fast_log(sacc) + mx[t]);
            }
            ag_rows(u1);
            // dustbin row: every column is owned (slot 0/1) by exactly one lane
            const float m = finite_or_zero(fmaxf(warp_max(fmaxf(zr[0], zr[1])), zcorner));
            const float sacc = warp_sum(fast_exp(zr[0] - m) + fast_exp(zr[1] - m)) + fast_exp(zcorner - m);
            u1d = lmu_at(a, g, p, D) - (fast_log(sacc) + m);
        }
        {  // columns: v1_j = log nu_j - LSE_i (Z_ij + u1_i)
            float mx[16];
#pragma unroll
            for (int c = 0; c < 16; ++c) {
                float m = z[0][c] + u1[0];
#pragma unroll
                for (int k = 1; k < 8; ++k) m = fmaxf(m, z[k][c] + u1[k]);
                mx[c] = m;
            }
            rs_cols_op(mx, OpMax());
#pragma unroll
            for (int t = 0; t < 2; ++t) mx[t] = finite_or_zero(fmaxf(mx[t], zr[t] + u1d));
            ag_cols(mx);
#pragma unroll
            for (int c = 0; c < 16; ++c) {
                float sacc = 0.f;
#pragma unroll
                for (int k = 0; k < 8; ++k) sacc += fast_exp((z[k][c] + u1[k]) - mx[c]);
                v1[c] = sacc;
            }
            rs_cols(v1);
#pragma unroll
            for (int t = 0; t < 2; ++t) {
                const float sacc = v1[t] + fast_exp((zr[t] + u1d) - mx[t]);
                v1[t] = lnu_at(a, g, p, LCOL(t)) - (fast_log(sacc) + mx[t]);
            }
            ag_cols(v1);
            // dustbin column: every row is owned by exactly one lane
            const float m = finite_or_zero(fmaxf(warp_max(fmaxf(zc[0] + u1[0], zc[1] + u1[1])), zcorner + u1d));
            const float sacc = warp_sum(fast_exp((zc[0] + u1[0]) - m) + fast_exp((zc[1] + u1[1]) - m)) + fast_exp((zcorner + u1d) - m);
            v1d = lnu_at(a, g, p, D) - (fast_log(sacc) + m);
        }
#pragma unroll
        for (int k = 0; k < 8; ++k)
#pragma unroll
            for (int c = 0; c < 16; ++c) z[k][c] = fast_exp((z[k][c] + u1[k]) + v1[c]);
#pragma unroll
        for (int t = 0; t < 2; ++t) {
            Dc[t] = fast_exp((zc[t] + u1[t]) + v1d);
            Dr[t] = fast_exp((zr[t] + u1d) + v1[t]);
            u1o[t] = u1[t], v1o[t] = v1[t];
        }
        corner = fast_exp((zcorner + u1d) + v1d);
    }

    // ---- iterations 2..iters on the register tile -------------------------------------------------------------------
    float2 Kp[8][8];
#pragma unroll
    for (int k = 0; k < 8; ++k)
#pragma unroll
        for (int h = 0; h < 8; ++h) Kp[k][h] = make_float2(z[k][2 * h], z[k][2 * h + 1]);
    float be[16], al[8];  // beta of my 16 column slots, alpha of my 8 row slots
#pragma unroll
    for (int c = 0; c < 16; ++c) be[c] = 1.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) al[k] = 1.f;
    float ald = 1.f, bed = 1.f;
    float Sr = warp_sum(Dr[0] + Dr[1]);  // sum_j K[D][j] beta_j with beta = 1
    float lo = INFINITY, hi = 0.f;

    for (int it = 1; it < a.iters; ++it) {
        // alpha_i = mu_i / sum_j K_ij beta_j
        float2 acc[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[k] = make_float2(0.f, 0.f);
#pragma unroll
        for (int h = 0; h < 8; ++h) {
            const float2 bp = make_float2(be[2 * h], be[2 * h + 1]);
#pragma unroll
            for (int k = 0; k < 8; ++k) acc[k] = ffma2(Kp[k][h], bp, acc[k]);
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) al[k] = acc[k].x + acc[k].y;
        rs_rows(al);
#pragma unroll
        for (int t = 0; t < 2; ++t) al[t] = mu2[t] * fast_rcp(fmaf(Dc[t], bed, al[t]));
        ald = mud * fast_rcp(fmaf(corner, bed, Sr));
        float Sc = warp_sum(fmaf(Dc[0], al[0], Dc[1] * al[1]));  // sum_i K[i][D] alpha_i (hidden under the column pass)
        const float al0 = al[0], al1 = al[1];
        ag_rows(al);
        // beta_j = nu_j / sum_i K_ij alpha_i
        float2 s2[8];
#pragma unroll
        for (int h = 0; h < 8; ++h) s2[h] = make_float2(0.f, 0.f);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const float2 ak = make_float2(al[k], al[k]);
#pragma unroll
            for (int h = 0; h < 8; ++h) s2[h] = ffma2(Kp[k][h], ak, s2[h]);
        }
#pragma unroll
        for (int h = 0; h < 8; ++h) be[2 * h] = s2[h].x, be[2 * h + 1] = s2[h].y;
        rs_cols(be);
#pragma unroll
        for (int t = 0; t < 2; ++t) be[t] = nu2[t] * fast_rcp(fmaf(Dr[t], ald, be[t]));
        bed = nud * fast_rcp(fmaf(corner, ald, Sc));
        Sr = warp_sum(fmaf(Dr[0], be[0], Dr[1] * be[1]));
        if ((it & 7) == 0 || it == a.iters - 1) {
            lo = fminf(fminf(fminf(lo, al0), fminf(al1, ald)), fminf(fminf(be[0], be[1]), bed));
            hi = fmaxf(fmaxf(fmaxf(hi, al0), fmaxf(al1, ald)), fmaxf(fmaxf(be[0], be[1]), bed));
        }
        if (it == a.iters - 1) al[0] = al0, al[1] = al1;  // keep the owned alphas in slots 0,1 for the epilogue
        ag_cols(be);
    }

    // ---- potentials, health check, output ------------------------------------------------------------------------------
    const float shift = (a.mode == MODE_RAW) ? 0.f : g.norm;
    float U[8], V[16], Ud = 0.f, Vd = -shift;
    bool bad = !(lo >= 1e-13f && hi <= 1e13f);
#pragma unroll
    for (int t = 0; t < 2; ++t) {
        float tu = 0.f, tv = 0.f;
        if (a.iters >= 1) tu = u1o[t], tv = v1o[t];
        if (a.iters >= 2) tu += fast_log(al[t]), tv += fast_log(be[t]);
        if (!(fabsf(tu) < INFINITY) || !(fabsf(tv) < INFINITY)) bad = true;
        U[t] = tu, V[t] = tv - shift;
    }
    if (a.iters >= 1) Ud = u1d, Vd = v1d - shift;
    if (a.iters >= 2) Ud += fast_log(ald), Vd += fast_log(bed);
    if (!(fabsf(Ud) < INFINITY) || !(fabsf(Vd) < INFINITY)) bad = true;
    if (__any_sync(0xffffffffu, bad)) {
        if (lane == 0 && a.fb_total) atomicAdd(a.fb_total, 1);
        log_domain_solve<32>(a, g, p, s_fb[wib], s_fb[wib] + 65, s_fb[wib] + 130, lane, WarpSync());
        return;
    }
    ag_rows(U);
    ag_cols(V);
    float *o = a.out + (size_t)p * 65 * 65;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const int row = LROW(k);
#pragma unroll
        for (int c = 0; c < 16; ++c) o[row * 65 + LCOL(c)] = (z_at(a, g, p, row, LCOL(c)) + U[k]) + V[c];
        if (qc == 0) o[row * 65 + D] = (z_at(a, g, p, row, D) + U[k]) + Vd;
    }
    if (pr == 0) {
#pragma unroll
        for (int c = 0; c < 16; ++c) o[D * 65 + LCOL(c)] = (z_at(a, g, p, D, LCOL(c)) + Ud) + V[c];
    }
    if (lane == 0) o[D * 65 + D] = (z_at(a, g, p, D, D) + Ud) + Vd;
#undef LROW
#undef LCOL
}

// ---------------------------------------------------------------------------------------------
// Level-3 kernel, two warps per problem (the default for 65 x 65).
//   Warp w of the pair keeps columns [32w, 32w+32) of the 64 x 64 core: an 8 x 8 tile per lane (64 registers),
//   so four warps per scheduler stay resident and each iteration's dependency chain is half as long as in the
//   one-warp kernel.  Column sums are complete inside a warp; the partial ROW sums of the two warps meet through
//   128 floats of shared memory and one 64-thread named barrier per iteration (double-buffered by parity).
//   Both warps then compute bit-identical alphas (a + b == b + a).  Slot permutations as in sinkhorn_w65_kernel.
// ---------------------------------------------------------------------------------------------
#ifndef X2_PAIRS_N
#define X2_PAIRS_N 1  // measured: 1 problem per CTA (8 CTAs per SM) beats 2 (-1.4 % at 4800 problems, -2.3 % at 30 000), 4 loses 10 %
#endif
constexpr int X2_PAIRS = X2_PAIRS_N;  // problems per CTA (two warps each)
constexpr float kDirectZ = 12.f;  // |z| bound of the direct start (exp(+-12) ~ 1.6e5 / 6e-6: scalings stay far inside [1e-13, 1e13])
#ifndef X2_MIN_CTAS
#define X2_MIN_CTAS (8 / X2_PAIRS_N)
#endif

struct PairSync {
    int id;  // 1 or 2: literal barrier ids so that ptxas reserves 3 barriers, not all 16 (16 would cap the CTAs per SM at 4)
    __device__ __forceinline__ void operator()() const {
        if (id == 1) asm volatile("bar.sync 1, 64;" ::: "memory");
        else if (id == 2 || X2_PAIRS_N <= 2) asm volatile("bar.sync 2, 64;" ::: "memory");
        else if (id == 3) asm volatile("bar.sync 3, 64;" ::: "memory");
        else asm volatile("bar.sync 4, 64;" ::: "memory");
    }
};

template <class Op>
__device__ __forceinline__ void rs_cols8_op(float (&v)[8], Op op) {  // reduce-scatter over pr (lane bits 2,3,4): 8 -> 1
#pragma unroll
    for (int t = 0; t < 4; ++t) v[t] = op(v[t], __shfl_xor_sync(0xffffffffu, v[t + 4], 4));
#pragma unroll
    for (int t = 0; t < 2; ++t) v[t] = op(v[t], __shfl_xor_sync(0xffffffffu, v[t + 2], 8));
    v[0] = op(v[0], __shfl_xor_sync(0xffffffffu, v[1], 16));
}
__device__ __forceinline__ void ag_cols8(float (&v)[8]) {  // all-gather over pr: 1 -> 8
    v[1] = __shfl_xor_sync(0xffffffffu, v[0], 16);
#pragma unroll
    for (int t = 0; t < 2; ++t) v[t + 2] = __shfl_xor_sync(0xffffffffu, v[t], 8);
#pragma unroll
    for (int t = 0; t < 4; ++t) v[t + 4] = __shfl_xor_sync(0xffffffffu, v[t], 4);
}

// ---- bulk-copy (TMA engine) staging of a whole problem ----------------------------------------------------------------
// A [b,65,65] plan is 16 900 B per problem: problem starts are only 4-byte aligned, so neither a tiled tensor map nor an exact
// 1-D bulk copy applies (both need 16-byte aligned addresses and sizes).  The BULK variant copies the ALIGNED SUPERSET of the
// problem -- [start & ~15, (end + 15) & ~15), at most 16 928 B -- into a 16-byte aligned shared buffer with ONE
// cp.async.bulk (SASS UBLKCP) that completes on an mbarrier; element e of the problem then sits at float index
// `sh + e`, sh = (start >> 2) & 3.  The result is formed IN PLACE in that buffer and leaves with one cp.async.bulk store
// of the aligned interior (plus <= 3 + 3 scalar stores for the unaligned head / tail).  CTAs are persistent: the load of
// the CTA's next problem is issued as soon as the store has read the buffer.
constexpr int X2_STAGE_FLOATS = 4232 + 8;  // 16 928 B superset + slack, in floats
__device__ __forceinline__ void bulk_load(unsigned dst_smem, const void *src, unsigned bytes, unsigned mbar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_smem), "l"(src),
                 "r"(bytes), "r"(mbar)
                 : "memory");
}
__device__ __forceinline__ void bulk_store(void *dst, unsigned src_smem, unsigned bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src_smem), "r"(bytes) : "memory");
}

template <bool BULK>
__global__ void __launch_bounds__(X2_PAIRS * 64, X2_MIN_CTAS) sinkhorn_w65x2_kernel(SinkArgs a) {
    constexpr int D = 64;
    static_assert(!BULK || X2_PAIRS == 1, "the bulk-copy variant stages one problem per CTA");
    __shared__ __align__(16) float s_x[X2_PAIRS][2][2][68];  // [pair][parity][warp]: 64 owned-row values + scalars (64, 65: one float2)
    __shared__ float s_fb[X2_PAIRS][65 + 65 + 128];
    __shared__ __align__(128) float s_stage[BULK ? X2_STAGE_FLOATS : 4];
    __shared__ __align__(8) unsigned long long s_mbar;
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int pair = wib >> 1, w = wib & 1;
    // Launch chaining (common.cuh).  The persistent variant lets the next grid in only when a CTA begins its LAST problem: a
    // hand-over consumer that became resident at the start would spin on its flag for the whole solve, next to the producers.
    if constexpr (BULK) pdl_wait();
    else pdl_prologue();
    const PairSync psync{1 + pair};
    const int pr = lane >> 2, qc = lane & 3;
    const int rmask = ((qc & 1) << 2) | ((qc >> 1) << 1);
    const int cmask = ((pr & 1) << 2) | (((pr >> 1) & 1) << 1) | (pr >> 2);
#define LROW(k) (pr + 8 * ((k) ^ rmask))
#define LCOL(c) (32 * w + qc + 4 * ((c) ^ cmask))
    [[maybe_unused]] unsigned mb = 0, stage_u32 = 0, phase = 0;
    // first float of the aligned superset of problem q, its length in bytes, and the problem's offset inside it (floats)
    auto superset = [&](int q, const float *&src, unsigned &bytes) {
        const uintptr_t s0 = reinterpret_cast<uintptr_t>(a.Z + (size_t)q * 4225), e0 = s0 + 16900;
        src = reinterpret_cast<const float *>(s0 & ~(uintptr_t)15);
        bytes = (unsigned)(((e0 + 15) & ~(uintptr_t)15) - (s0 & ~(uintptr_t)15));
        return (int)((s0 & 15) >> 2);
    };
    if constexpr (BULK) {
        mb = smem_u32(&s_mbar), stage_u32 = smem_u32(s_stage);
        if (threadIdx.x == 0) {
            mbar_init(mb, 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            if ((int)blockIdx.x < a.b) {
                const float *src;
                unsigned bytes;
                superset((int)blockIdx.x, src, bytes);
                mbar_expect_tx(mb, bytes);
                bulk_load(stage_u32, src, bytes, mb);
            }
        }
        psync();
    }
    int p = BULK ? (int)blockIdx.x : (int)blockIdx.x * X2_PAIRS + pair;
    if (p >= a.b) return;  // uniform over the pair; the named barrier involves this pair only
    do {  // BULK: the problems of this persistent CTA; otherwise exactly one pass
    if constexpr (BULK) {
        if (p + (int)gridDim.x >= a.b) pdl_launch_dependents();
    }
    const Marg g = problem_marginals(a, p, lane);
    int par = 0;
    // exchange the two owned-row values and one scalar with the partner warp (double-buffered by parity).
    // (Precomputing the six smem addresses was measured SLOWER: +9% at 153.6 k problems -- more live registers.)
#define PAIR_XCHG(v0, v1, sc, o0, o1, osc)                               \
    do {                                                                 \
        float *mine_ = s_x[pair][par][w], *oth_ = s_x[pair][par][w ^ 1]; \
        mine_[LROW(0)] = (v0);                                           \
        mine_[LROW(1)] = (v1);                                           \
        if (lane == 0) mine_[64] = (sc);                                 \
        psync();                                                         \
        (o0) = oth_[LROW(0)];                                            \
        (o1) = oth_[LROW(1)];                                            \
        (osc) = oth_[64];                                                \
        par ^= 1;                                                        \
    } while (0)
    // the loop's exchange: one more scalar (the fixed-point vote of the previous iteration) in the same barrier phase
#define PAIR_XCHG_IT(v0, v1, sc, fx, o0, o1, osc, ofx)                   \
    do {                                                                 \
        float *mine_ = s_x[pair][par][w], *oth_ = s_x[pair][par][w ^ 1]; \
        mine_[LROW(0)] = (v0);                                           \
        mine_[LROW(1)] = (v1);                                           \
        if (lane == 0) *reinterpret_cast<float2 *>(mine_ + 64) = make_float2((sc), (fx)); \
        psync();                                                         \
        (o0) = oth_[LROW(0)];                                            \
        (o1) = oth_[LROW(1)];                                            \
        {                                                                \
            const float2 t_ = *reinterpret_cast<const float2 *>(oth_ + 64); \
            (osc) = t_.x, (ofx) = t_.y;                                  \
        }                                                                \
        par ^= 1;                                                        \
    } while (0)

    // ---- load ------------------------------------------------------------------------------------------------------
    float z[8][8], zc[2], zr, zcorner;
    [[maybe_unused]] int sh = 0;
    if constexpr (BULK) {
        const float *src_;
        unsigned bytes_;
        sh = superset(p, src_, bytes_);
        mbar_wait(mb, phase);  // the problem has landed in s_stage
        phase ^= 1u;
        const float *zs = s_stage + sh;
        int coff[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) coff[c] = LCOL(c);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const int roff = LROW(k) * 65;
#pragma unroll
            for (int c = 0; c < 8; ++c) z[k][c] = zs[roff + coff[c]];
        }
#pragma unroll
        for (int t = 0; t < 2; ++t) zc[t] = zs[LROW(t) * 65 + D];
        zr = zs[D * 65 + LCOL(0)];
        zcorner = zs[D * 65 + D];
    } else {
        const PlanRef pl = plan_ref(a, g, p);
        int coff[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) coff[c] = LCOL(c);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const int roff = LROW(k) * pl.stride;
#pragma unroll
            for (int c = 0; c < 8; ++c) z[k][c] = pl.core(roff, coff[c]);
        }
#pragma unroll
        for (int t = 0; t < 2; ++t) zc[t] = pl.edge(LROW(t), D);
        zr = pl.edge(D, LCOL(0));  // dustbin-row entry of the one column this lane owns
        zcorner = pl.edge(D, D);
    }
    float mu2[2];
#pragma unroll
    for (int t = 0; t < 2; ++t) mu2[t] = expf(lmu_at(a, g, p, LROW(t)));
    const float nu1 = expf(lnu_at(a, g, p, LCOL(0)));
    const float mud = expf(lmu_at(a, g, p, D)), nud = expf(lnu_at(a, g, p, D));
    float u1o[2] = {0.f, 0.f}, v1o = 0.f, u1d = 0.f, v1d = 0.f;
    float Dc[2] = {0.f, 0.f}, Dr = 0.f, corner = 0.f;

    // ---- iteration 1 ------------------------------------------------------------------------------------------------------
    // Moderate scores (every |z| <= kDirectZ, the case for descriptor correlations scaled by 0.1 / sqrt(d)) need no
    // log-domain first iteration: exp(z) cannot leave the f32 range, so the scaling iteration starts directly on
    // K = exp(Z) with alpha = beta = 1 and runs `iters` times -- the same iteration, u1 = v1 = 0 absorbed.  Anything
    // else (large or non-finite entries) takes the exact max-shifted log-sum-exp path below, as before.
    int it0 = 1;
#ifndef PATS_AB_NO_DIRECT
    {
        bool ok = fabsf(zr) <= kDirectZ && fabsf(zcorner) <= kDirectZ && fabsf(zc[0]) <= kDirectZ && fabsf(zc[1]) <= kDirectZ;
#pragma unroll
        for (int k = 0; k < 8; ++k)
#pragma unroll
            for (int c = 0; c < 8; ++c) ok = ok && (fabsf(z[k][c]) <= kDirectZ);  // NaN fails the comparison
        const float mine = __all_sync(0xffffffffu, ok) ? 1.f : 0.f;
        [[maybe_unused]] float d0, d1;
        float other;
        PAIR_XCHG(0.f, 0.f, mine, d0, d1, other);
        (void)d0, (void)d1;
        if (mine != 0.f && other != 0.f) it0 = 0;
    }
    if (it0 == 0) {
#pragma unroll
        for (int k = 0; k < 8; ++k)
#pragma unroll
            for (int c = 0; c < 8; ++c) z[k][c] = fast_exp(z[k][c]);
        Dc[0] = fast_exp(zc[0]), Dc[1] = fast_exp(zc[1]);
        Dr = fast_exp(zr);
        corner = fast_exp(zcorner);
    } else
#endif
    if (a.iters >= 1) {
        float u1[8], v1[8];
        {
            float mx[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                float m = z[k][0];
#pragma unroll
                for (int c = 1; c < 8; ++c) m = fmaxf(m, z[k][c]);
                mx[k] = m;
            }
            rs_rows_op(mx, OpMax());
            float o0, o1, osc;
            const float drm_part = warp_max(zr);
            PAIR_XCHG(mx[0], mx[1], drm_part, o0, o1, osc);
            mx[0] = finite_or_zero(fmaxf(fmaxf(mx[0], o0), zc[0]));
            mx[1] = finite_or_zero(fmaxf(fmaxf(mx[1], o1), zc[1]));
            const float drm = finite_or_zero(fmaxf(fmaxf(drm_part, osc), zcorner));
            ag_rows(mx);
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                float sacc = 0.f;
#pragma unroll
                for (int c = 0; c < 8; ++c) sacc += fast_exp(z[k][c] - mx[k]);
                u1[k] = sacc;
            }
            rs_rows(u1);
            const float drs_part = warp_sum(fast_exp(zr - drm));
            PAIR_XCHG(u1[0], u1[1], drs_part, o0, o1, osc);
#pragma unroll
            for (int t = 0; t < 2; ++t) {
                const float sacc = (u1[t] + (t == 0 ? o0 : o1)) + fast_exp(zc[t] - mx[t]);
                u1[t] = lmu_at(a, g, p, LROW(t)) - (fast_log(sacc) + mx[t]);
            }
            u1d = lmu_at(a, g, p, D) - (fast_log((drs_part + osc) + fast_exp(zcorner - drm)) + drm);
            ag_rows(u1);
        }
        {
            float mx[8];
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                float m = z[0][c] + u1[0];
#pragma unroll
                for (int k = 1; k < 8; ++k) m = fmaxf(m, z[k][c] + u1[k]);
                mx[c] = m;
            }
            rs_cols8_op(mx, OpMax());
            mx[0] = finite_or_zero(fmaxf(mx[0], zr + u1d));
            ag_cols8(mx);
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                float sacc = 0.f;
#pragma unroll
                for (int k = 0; k < 8; ++k) sacc += fast_exp((z[k][c] + u1[k]) - mx[c]);
                v1[c] = sacc;
            }
            rs_cols8_op(v1, OpSum());
            v1[0] = lnu_at(a, g, p, LCOL(0)) - (fast_log(v1[0] + fast_exp((zr + u1d) - mx[0])) + mx[0]);
            ag_cols8(v1);
            const float m = finite_or_zero(fmaxf(warp_max(fmaxf(zc[0] + u1[0], zc[1] + u1[1])), zcorner + u1d));
            const float sacc = warp_sum(fast_exp((zc[0] + u1[0]) - m) + fast_exp((zc[1] + u1[1]) - m)) + fast_exp((zcorner + u1d) - m);
            v1d = lnu_at(a, g, p, D) - (fast_log(sacc) + m);
        }
#pragma unroll
        for (int k = 0; k < 8; ++k)
#pragma unroll
            for (int c = 0; c < 8; ++c) z[k][c] = fast_exp((z[k][c] + u1[k]) + v1[c]);
#pragma unroll
        for (int t = 0; t < 2; ++t) {
            Dc[t] = fast_exp((zc[t] + u1[t]) + v1d);
            u1o[t] = u1[t];
        }
        Dr = fast_exp((zr + u1d) + v1[0]);
        v1o = v1[0];
        corner = fast_exp((zcorner + u1d) + v1d);
    }

    // ---- iterations 2..iters ------------------------------------------------------------------------------------------
    float2 Kp[8][4];
#pragma unroll
    for (int k = 0; k < 8; ++k)
#pragma unroll
        for (int h = 0; h < 4; ++h) Kp[k][h] = make_float2(z[k][2 * h], z[k][2 * h + 1]);
    float be[8], al[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) be[c] = 1.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) al[k] = 1.f;
    float ald = 1.f, bed = 1.f;
#ifndef PATS_AB_NO_PIPE_SRP
    // this warp's half of sum_j K[D][j] beta_j: the 5-level butterfly is split over the loop's back-edge -- two levels behind
    // the beta update, three between the FFMA2 groups of the next row pass (ptxas does not move code across the back-edge,
    // and an in-order warp otherwise sits on every SHFL -> FADD pair of the chain with nothing else to issue: -1.6 %)
    float Srp = Dr;
    Srp += __shfl_xor_sync(0xffffffffu, Srp, 16);
    Srp += __shfl_xor_sync(0xffffffffu, Srp, 8);
#else
    float Srp = warp_sum(Dr);  // this warp's half of sum_j K[D][j] beta_j
#endif
    float lo = INFINITY, hi = 0.f;
    // Fixed-point exit (SinkArgs::fp_exit).  beta_t == beta_{t-1} bit for bit -- every column of both warps and the dustbin -- makes
    // iteration t+1 a replay of iteration t (same operands through the same instructions), and so every later one: the remaining
    // iterations cannot change a bit of the result.  The vote of iteration t travels with iteration t+1's row exchange; the
    // pair then finishes that iteration's alpha update (identical to iteration t's by the same argument) and leaves.
    float pbe0 = 0.f, pbed = 0.f, fix = 0.f;  // a beta is never 0, so the first comparison fails

    for (int it = it0; it < a.iters; ++it) {
        float2 acc[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[k] = make_float2(0.f, 0.f);
#pragma unroll
        for (int h = 0; h < 4; ++h) {
            const float2 bp = make_float2(be[2 * h], be[2 * h + 1]);
#pragma unroll
            for (int k = 0; k < 8; ++k) acc[k] = ffma2(Kp[k][h], bp, acc[k]);
#ifndef PATS_AB_NO_PIPE_SRP
            if (h < 3) Srp += __shfl_xor_sync(0xffffffffu, Srp, 4 >> h);
#endif
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) al[k] = acc[k].x + acc[k].y;
        rs_rows(al);
        float o0, o1, oS, ofix;
        PAIR_XCHG_IT(al[0], al[1], Srp, fix, o0, o1, oS, ofix);
        al[0] = mu2[0] * fast_rcp(fmaf(Dc[0], bed, al[0] + o0));
        al[1] = mu2[1] * fast_rcp(fmaf(Dc[1], bed, al[1] + o1));
        ald = mud * fast_rcp(fmaf(corner, bed, Srp + oS));
        if (fix != 0.f && ofix != 0.f) {  // both warps: beta did not change in the previous iteration
            lo = fminf(fminf(fminf(lo, al[0]), fminf(al[1], ald)), fminf(be[0], bed));  // the sample the last iteration would have taken
            hi = fmaxf(fmaxf(fmaxf(hi, al[0]), fmaxf(al[1], ald)), fmaxf(be[0], bed));
            if (w == 0 && lane == 0 && a.fb_total) atomicAdd(a.fb_total + 1, a.iters - it);  // iterations not executed
            break;
        }
        // (deferring this butterfly and the beta_D update into the next row pass as well was measured 2.7 % SLOWER: beta_D's
        // reciprocal then sits in front of the alpha update)
        const float Sc = warp_sum(fmaf(Dc[0], al[0], Dc[1] * al[1]));
        ag_rows(al);
        float2 s2[4];
#pragma unroll
        for (int h = 0; h < 4; ++h) s2[h] = make_float2(0.f, 0.f);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const float2 ak = make_float2(al[k], al[k]);
#pragma unroll
            for (int h = 0; h < 4; ++h) s2[h] = ffma2(Kp[k][h], ak, s2[h]);
        }
#pragma unroll
        for (int h = 0; h < 4; ++h) be[2 * h] = s2[h].x, be[2 * h + 1] = s2[h].y;
        rs_cols8_op(be, OpSum());
        be[0] = nu1 * fast_rcp(fmaf(Dr, ald, be[0]));
        bed = nud * fast_rcp(fmaf(corner, ald, Sc));
        fix = (a.fp_exit && __all_sync(0xffffffffu, be[0] == pbe0 && bed == pbed)) ? 1.f : 0.f;
        pbe0 = be[0], pbed = bed;
#ifndef PATS_AB_NO_PIPE_SRP
        Srp = Dr * be[0];
        Srp += __shfl_xor_sync(0xffffffffu, Srp, 16);
        Srp += __shfl_xor_sync(0xffffffffu, Srp, 8);
#else
        Srp = warp_sum(Dr * be[0]);
#endif
        if ((it & 7) == 0 || it == a.iters - 1) {
            lo = fminf(fminf(fminf(lo, al[0]), fminf(al[1], ald)), fminf(be[0], bed));
            hi = fmaxf(fmaxf(fmaxf(hi, al[0]), fmaxf(al[1], ald)), fmaxf(be[0], bed));
        }
        ag_cols8(be);
    }

    // ---- potentials, health check (agreed over the pair), output -----------------------------------------------------------
    const float shift = (a.mode == MODE_RAW) ? 0.f : g.norm;
    const bool ran = a.iters > it0;  // the scaling loop ran at least once
    float U[8], V[8], Ud = 0.f, Vd = -shift;
    bool bad = !(lo >= 1e-13f && hi <= 1e13f);

#pragma unroll
    for (int t = 0; t < 2; ++t) {
        float tu = 0.f;
        if (a.iters >= 1) tu = u1o[t];
        if (ran) tu += fast_log(al[t]);
        if (!(fabsf(tu) < INFINITY)) bad = true;
        U[t] = tu;
    }
    {
        float tv = 0.f;
        if (a.iters >= 1) tv = v1o;
        if (ran) tv += fast_log(be[0]);
        if (!(fabsf(tv) < INFINITY)) bad = true;
        V[0] = tv - shift;
    }
    if (a.iters >= 1) Ud = u1d, Vd = v1d - shift;
    if (ran) Ud += fast_log(ald), Vd += fast_log(bed);
    if (!(fabsf(Ud) < INFINITY) || !(fabsf(Vd) < INFINITY)) bad = true;
    {
        const float mine = __any_sync(0xffffffffu, bad) ? 1.f : 0.f;
        [[maybe_unused]] float d0, d1;
        float other;
        PAIR_XCHG(0.f, 0.f, mine, d0, d1, other);
        (void)d0, (void)d1;
        bad = (mine != 0.f) || (other != 0.f);
    }
    if (bad) {
        if (w == 0 && lane == 0 && a.fb_total) atomicAdd(a.fb_total, 1);
        psync();  // everybody is done with s_x before the scratch is reused
        log_domain_solve<64>(a, g, p, s_fb[pair], s_fb[pair] + 65, s_fb[pair] + 130, w * 32 + lane, psync);
        psync();
        if (w == 0 && lane == 0) publish_problem(a, p);
    } else {
        ag_rows(U);
        ag_cols8(V);
        float *o = a.out + (size_t)p * 65 * 65;
        if constexpr (BULK) {
            // the result is formed in place in the staged copy (same expressions as the direct path: bit-identical) ...
            float *zs = s_stage + sh;
            int coff[8];
#pragma unroll
            for (int c = 0; c < 8; ++c) coff[c] = LCOL(c);
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const int row = LROW(k);
#pragma unroll
                for (int c = 0; c < 8; ++c) zs[row * 65 + coff[c]] = (zs[row * 65 + coff[c]] + U[k]) + V[c];
                if (w == 0 && qc == 0) zs[row * 65 + D] = (zs[row * 65 + D] + U[k]) + Vd;
            }
            if (pr == 0) {
#pragma unroll
                for (int c = 0; c < 8; ++c) zs[D * 65 + coff[c]] = (zs[D * 65 + coff[c]] + Ud) + V[c];
            }
            if (w == 0 && lane == 0) zs[D * 65 + D] = (zs[D * 65 + D] + Ud) + Vd;
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> visible to the bulk store
            psync();
            // ... and leaves with ONE bulk store of its 16-byte aligned interior; <= 3 floats on either side go by hand.  (The host
            // only picks this variant when input and output share their alignment phase, so the interior is aligned on both sides.)
            if (threadIdx.x == 0) {
                const int head = (int)(((16u - (unsigned)(reinterpret_cast<uintptr_t>(o) & 15)) & 15u) >> 2);
                const int body = (4225 - head) & ~3;
                for (int e = 0; e < head; ++e) o[e] = zs[e];
                for (int e = head + body; e < 4225; ++e) o[e] = zs[e];
                bulk_store(o + head, stage_u32 + 4u * (unsigned)(sh + head), (unsigned)body * 4u);
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                if (a.done) {
                    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");  // written, not just read: a consumer is waiting on the flag
                    publish_problem(a, p);
                } else {
                    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");  // the buffer may be overwritten
                }
            }
        } else {
            const PlanRef pl = plan_ref(a, g, opaque(p));  // recomputed: keeping the load-time copy alive costs registers in the loop
            int coff[8];
#pragma unroll
            for (int c = 0; c < 8; ++c) coff[c] = LCOL(c);
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const int row = LROW(k);
                const int roff = row * pl.stride;
                float zz[8];
#pragma unroll
                for (int c = 0; c < 8; ++c) zz[c] = pl.core(roff, coff[c]);  // all eight loads of the row in flight (L2 hits)
#pragma unroll
                for (int c = 0; c < 8; ++c) o[row * 65 + coff[c]] = (zz[c] + U[k]) + V[c];
                if (w == 0 && qc == 0) o[row * 65 + D] = (pl.edge(row, D) + U[k]) + Vd;
            }
            if (pr == 0) {
#pragma unroll
                for (int c = 0; c < 8; ++c) o[D * 65 + coff[c]] = (pl.edge(D, coff[c]) + Ud) + V[c];
            }
            if (w == 0 && lane == 0) o[D * 65 + D] = (pl.edge(D, D) + Ud) + Vd;
            if (a.done) {  // both warps of the pair have stored their halves
                psync();
                if (w == 0 && lane == 0) publish_problem(a, p);
            }
        }
    }
    if constexpr (BULK) {  // stage the CTA's next problem (the buffer is free: every thread passed the barrier above)
        if (threadIdx.x == 0) {
            const int q = p + (int)gridDim.x;
            if (q < a.b) {
                const float *src;
                unsigned bytes;
                superset(q, src, bytes);
                mbar_expect_tx(mb, bytes);
                bulk_load(stage_u32, src, bytes, mb);
            }
        }
    }
    if constexpr (BULK) p += (int)gridDim.x;
    } while (BULK && p < a.b);
#undef PAIR_XCHG
#undef PAIR_XCHG_IT
#undef LROW
#undef LCOL
}

// ---------------------------------------------------------------------------------------------
// Level-2 kernel: exactly 145 x 145 (144 x 144 real cells + dustbin row / column), one CTA of 9 warps per problem.
//   Warp w keeps the 16-column slab [16w, 16w+16) of the core; lane (pr = lane>>1, qc = lane&1) holds a 9 x 8 tile
//   (rows pr + 16k, columns 16w + qc + 2*(c ^ cmask(pr))): 72 registers, two CTAs per SM.
//   Column sums are complete inside a warp (select-free reduce-scatter over pr, as in the level-3 kernels); the row
//   partials of the 9 warps meet in shared memory, where thread t < 144 finishes row t (its dustbin-column entry,
//   marginal and first-iteration potential live in that thread's registers).  Two CTA barriers per iteration.
// ---------------------------------------------------------------------------------------------
constexpr int C145_W = 9, C145_T = 288;

template <class Op>
__device__ __forceinline__ void rs_c145(float (&v)[8], Op op) {  // over pr (lane bits 1..4): 8 -> 1 (twins pr, pr^8 both hold it)
#pragma unroll
    for (int t = 0; t < 4; ++t) v[t] = op(v[t], __shfl_xor_sync(0xffffffffu, v[t + 4], 2));
#pragma unroll
    for (int t = 0; t < 2; ++t) v[t] = op(v[t], __shfl_xor_sync(0xffffffffu, v[t + 2], 4));
    v[0] = op(v[0], __shfl_xor_sync(0xffffffffu, v[1], 8));
    v[0] = op(v[0], __shfl_xor_sync(0xffffffffu, v[0], 16));
}
__device__ __forceinline__ void ag_c145(float (&v)[8]) {  // 1 -> 8
    v[1] = __shfl_xor_sync(0xffffffffu, v[0], 8);
#pragma unroll
    for (int t = 0; t < 2; ++t) v[t + 2] = __shfl_xor_sync(0xffffffffu, v[t], 4);
#pragma unroll
    for (int t = 0; t < 4; ++t) v[t + 4] = __shfl_xor_sync(0xffffffffu, v[t], 2);
}

__global__ void __launch_bounds__(C145_T) sinkhorn_c145_kernel(SinkArgs a) {  // 9 warps; register allocation granularity allows one CTA per SM
    constexpr int D = 144;
    __shared__ float s_part[C145_W][144];  // per-warp row partials
    __shared__ float s_row[144];           // per-row values handed back to the tiles
    __shared__ float s_red[2][16];         // [0][w]: dustbin-row partial of warp w; [1][w]: dustbin-column partial of warp w
    __shared__ float s_keep[2][144];       // per-row first-iteration potential / final alpha (kept out of registers)
    __shared__ float s_fb[145 + 145 + 2 * C145_T];
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int p = blockIdx.x;
    if (p >= a.b) return;
    const int pr = lane >> 1, qc = lane & 1;
    const int cmask = ((pr & 1) << 2) | (((pr >> 1) & 1) << 1) | ((pr >> 2) & 1);
    const bool col_owner = (pr & 8) == 0;  // of the twins (pr, pr^8) that own a column, this one counts it
    // row finishing is spread over all warps: the qc==0 lane (w, pr) finishes row 16w + pr (144 rows = 9 warps x 16)
    const bool row_thread = qc == 0;
    const int myrow = 16 * w + pr;
    const Marg g = problem_marginals(a, p, lane);
#define LROW(k) (pr + 16 * (k))
#define LCOL(c) (16 * w + qc + 2 * ((c) ^ cmask))
    auto red9 = [&](int which) {  // fixed-order sum of the 9 per-warp partials
        float t = s_red[which][0];
#pragma unroll
        for (int i = 1; i < C145_W; ++i) t += s_red[which][i];
        return t;
    };
    auto max9 = [&](int which) {
        float t = s_red[which][0];
#pragma unroll
        for (int i = 1; i < C145_W; ++i) t = fmaxf(t, s_red[which][i]);
        return t;
    };

    // ---- load ------------------------------------------------------------------------------------------------------
    float z[9][8];
#pragma unroll
    for (int k = 0; k < 9; ++k)
#pragma unroll
        for (int c = 0; c < 8; ++c) z[k][c] = z_at(a, g, p, LROW(k), LCOL(c));
    const float zr = z_at(a, g, p, D, LCOL(0));                    // dustbin-row entry of my owned column
    const float zc = row_thread ? z_at(a, g, p, myrow, D) : 0.f;   // dustbin-column entry of my row
    const float zcorner = z_at(a, g, p, D, D);
    const float mu_t = row_thread ? expf(lmu_at(a, g, p, myrow)) : 0.f;
    const float nu_o = expf(lnu_at(a, g, p, LCOL(0)));
    const float mud = expf(lmu_at(a, g, p, D)), nud = expf(lnu_at(a, g, p, D));
    float v1_o = 0.f, u1d = 0.f, v1d = 0.f, Dc = 0.f, Dr = 0.f, corner = 0.f;

    // ---- iteration 1, exact in the log domain ---------------------------------------------------------------------
    if (a.iters >= 1) {
        float u1[9], v1[8], u1_t = 0.f;
        {  // row maxima
            float mx[9];
#pragma unroll
            for (int k = 0; k < 9; ++k) {
                float m = z[k][0];
#pragma unroll
                for (int c = 1; c < 8; ++c) m = fmaxf(m, z[k][c]);
                m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 1));
                if (qc == 0) s_part[w][LROW(k)] = m;
            }
            const float drm_w = warp_max(zr);
            if (lane == 0) s_red[0][w] = drm_w;
            __syncthreads();
            float rm = 0.f;
            if (row_thread) {
                rm = s_part[0][myrow];
#pragma unroll
                for (int i = 1; i < C145_W; ++i) rm = fmaxf(rm, s_part[i][myrow]);
                rm = finite_or_zero(fmaxf(rm, zc));
                s_row[myrow] = rm;
            }
            const float drm = finite_or_zero(fmaxf(max9(0), zcorner));
            __syncthreads();
#pragma unroll
            for (int k = 0; k < 9; ++k) mx[k] = s_row[LROW(k)];
#pragma unroll
            for (int k = 0; k < 9; ++k) {
                float sacc = 0.f;
#pragma unroll
                for (int c = 0; c < 8; ++c) sacc += fast_exp(z[k][c] - mx[k]);
                sacc += __shfl_xor_sync(0xffffffffu, sacc, 1);
                u1[k] = sacc;
            }
            __syncthreads();  // everybody has read s_row (maxima) and s_red[0]
#pragma unroll
            for (int k = 0; k < 9; ++k)
                if (qc == 0) s_part[w][LROW(k)] = u1[k];
            const float drs_w = warp_sum(col_owner ? fast_exp(zr - drm) : 0.f);
            if (lane == 0) s_red[0][w] = drs_w;
            __syncthreads();
            if (row_thread) {
                float sacc = s_part[0][myrow];
#pragma unroll
                for (int i = 1; i < C145_W; ++i) sacc += s_part[i][myrow];
                sacc += fast_exp(zc - rm);
                u1_t = lmu_at(a, g, p, myrow) - (fast_log(sacc) + rm);
                s_row[myrow] = u1_t;
                s_keep[0][myrow] = u1_t;
            }
            u1d = lmu_at(a, g, p, D) - (fast_log(red9(0) + fast_exp(zcorner - drm)) + drm);
            __syncthreads();
#pragma unroll
            for (int k = 0; k < 9; ++k) u1[k] = s_row[LROW(k)];
        }
        {  // columns (inside the warp) and the dustbin column (row threads)
            float mx[8];
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                float m = z[0][c] + u1[0];
#pragma unroll
                for (int k = 1; k < 9; ++k) m = fmaxf(m, z[k][c] + u1[k]);
                mx[c] = m;
            }
            rs_c145(mx, OpMax());
            mx[0] = finite_or_zero(fmaxf(mx[0], zr + u1d));
            ag_c145(mx);
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                float sacc = 0.f;
#pragma unroll
                for (int k = 0; k < 9; ++k) sacc += fast_exp((z[k][c] + u1[k]) - mx[c]);
                v1[c] = sacc;
            }
            rs_c145(v1, OpSum());
            v1[0] = lnu_at(a, g, p, LCOL(0)) - (fast_log(v1[0] + fast_exp((zr + u1d) - mx[0])) + mx[0]);
            ag_c145(v1);
            const float dcm_w = warp_max(row_thread ? zc + u1_t : -INFINITY);
            if (lane == 0) s_red[1][w] = dcm_w;
            __syncthreads();
            const float dcm = finite_or_zero(fmaxf(max9(1), zcorner + u1d));
            __syncthreads();
            const float dcs_w = warp_sum(row_thread ? fast_exp((zc + u1_t) - dcm) : 0.f);
            if (lane == 0) s_red[1][w] = dcs_w;
            __syncthreads();
            v1d = lnu_at(a, g, p, D) - (fast_log(red9(1) + fast_exp((zcorner + u1d) - dcm)) + dcm);
        }
#pragma unroll
        for (int k = 0; k < 9; ++k)
#pragma unroll
            for (int c = 0; c < 8; ++c) z[k][c] = fast_exp((z[k][c] + u1[k]) + v1[c]);
        Dc = row_thread ? fast_exp((zc + u1_t) + v1d) : 0.f;
        Dr = fast_exp((zr + u1d) + v1[0]);
        v1_o = v1[0];
        corner = fast_exp((zcorner + u1d) + v1d);
    }
    __syncthreads();

    // ---- iterations 2..iters ------------------------------------------------------------------------------------------
    // Per iteration: [row partials] B1 [finish rows: alpha] B2 [column pass: beta].  The two dustbin sums
    //   Sr = sum_j K[D][j] beta_j  and  Sc = sum_i K[i][D] alpha_i  are produced as per-warp partials AFTER B2 and consumed
    //   after the next B1, so their 5-step warp reductions overlap the FFMA2 stream instead of sitting between the barriers.
    float2 Kp[9][4];
#pragma unroll
    for (int k = 0; k < 9; ++k)
#pragma unroll
        for (int h = 0; h < 4; ++h) Kp[k][h] = make_float2(z[k][2 * h], z[k][2 * h + 1]);
    float be[8], al[9];
#pragma unroll
    for (int c = 0; c < 8; ++c) be[c] = 1.f;
    float al_t = 1.f, ald = 1.f, bed = 1.f;
    {
        const float srp = warp_sum(col_owner ? Dr : 0.f);
        if (lane == 0) s_red[0][w] = srp;
    }
    float lo = INFINITY, hi = 0.f;

    for (int it = 1; it < a.iters; ++it) {
        float2 acc[9];
#pragma unroll
        for (int k = 0; k < 9; ++k) acc[k] = make_float2(0.f, 0.f);
#pragma unroll
        for (int h = 0; h < 4; ++h) {
            const float2 bp = make_float2(be[2 * h], be[2 * h + 1]);
#pragma unroll
            for (int k = 0; k < 9; ++k) acc[k] = ffma2(Kp[k][h], bp, acc[k]);
        }
#pragma unroll
        for (int k = 0; k < 9; ++k) {
            float r = acc[k].x + acc[k].y;
            r += __shfl_xor_sync(0xffffffffu, r, 1);
            if (qc == 0) s_part[w][LROW(k)] = r;
        }
        __syncthreads();  // B1: row partials + both dustbin partials of the previous iteration are visible
        if (it > 1) bed = nud * fast_rcp(fmaf(corner, ald, red9(1)));  // beta_D of the previous iteration (Sc arrived late)
        ald = mud * fast_rcp(fmaf(corner, bed, red9(0)));
        if (row_thread) {
            float r = s_part[0][myrow];
#pragma unroll
            for (int i = 1; i < C145_W; ++i) r += s_part[i][myrow];
            al_t = mu_t * fast_rcp(fmaf(Dc, bed, r));
            s_row[myrow] = al_t;
        }
        __syncthreads();  // B2: alphas visible
#pragma unroll
        for (int k = 0; k < 9; ++k) al[k] = s_row[LROW(k)];
        {
            const float scp = warp_sum(row_thread ? Dc * al_t : 0.f);
            if (lane == 0) s_red[1][w] = scp;  // read after the next B1
        }
        float2 s2[4];
#pragma unroll
        for (int h = 0; h < 4; ++h) s2[h] = make_float2(0.f, 0.f);
#pragma unroll
        for (int k = 0; k < 9; ++k) {
            const float2 ak = make_float2(al[k], al[k]);
#pragma unroll
            for (int h = 0; h < 4; ++h) s2[h] = ffma2(Kp[k][h], ak, s2[h]);
        }
#pragma unroll
        for (int h = 0; h < 4; ++h) be[2 * h] = s2[h].x, be[2 * h + 1] = s2[h].y;
        rs_c145(be, OpSum());
        be[0] = nu_o * fast_rcp(fmaf(Dr, ald, be[0]));
        {
            const float srp = warp_sum(col_owner ? Dr * be[0] : 0.f);
            if (lane == 0) s_red[0][w] = srp;  // read after the next B1
        }
        if ((it & 7) == 0 || it == a.iters - 1) {
            const float at = row_thread ? al_t : ald;
            lo = fminf(fminf(lo, at), fminf(fminf(ald, be[0]), bed));
            hi = fmaxf(fmaxf(hi, at), fmaxf(fmaxf(ald, be[0]), bed));
        }
        ag_c145(be);
    }
    __syncthreads();
    if (a.iters >= 2) bed = nud * fast_rcp(fmaf(corner, ald, red9(1)));  // beta_D of the last iteration

    // ---- potentials, health check, output ------------------------------------------------------------------------------
    const float shift = (a.mode == MODE_RAW) ? 0.f : g.norm;
    bool bad = !(lo >= 1e-13f && hi <= 1e13f) || (a.iters >= 2 && !(bed >= 1e-13f && bed <= 1e13f));
    float U_t = 0.f, Ud = 0.f, Vd = -shift, V[8], U[9];
    if (a.iters >= 1) U_t = row_thread ? s_keep[0][myrow] : 0.f, Ud = u1d, Vd = v1d - shift;
    if (a.iters >= 2) U_t += fast_log(al_t), Ud += fast_log(ald), Vd += fast_log(bed);
    if (row_thread && !(fabsf(U_t) < INFINITY)) bad = true;
    if (!(fabsf(Ud) < INFINITY) || !(fabsf(Vd) < INFINITY)) bad = true;
    {
        float tv = 0.f;
        if (a.iters >= 1) tv = v1_o;
        if (a.iters >= 2) tv += fast_log(be[0]);
        if (!(fabsf(tv) < INFINITY)) bad = true;
        V[0] = tv - shift;
    }
    if (row_thread) s_row[myrow] = U_t;
    if (__syncthreads_or(bad ? 1 : 0)) {
        if (tid == 0 && a.fb_total) atomicAdd(a.fb_total, 1);
        log_domain_solve<C145_T>(a, g, p, s_fb, s_fb + 145, s_fb + 290, tid, BlockSync());
        return;
    }
#pragma unroll
    for (int k = 0; k < 9; ++k) U[k] = s_row[LROW(k)];
    ag_c145(V);
    float *o = a.out + (size_t)p * 145 * 145;
#pragma unroll
    for (int k = 0; k < 9; ++k) {
        const int row = LROW(k);
#pragma unroll
        for (int c = 0; c < 8; ++c) o[row * 145 + LCOL(c)] = (z_at(a, g, p, row, LCOL(c)) + U[k]) + V[c];
    }
    if (row_thread) o[myrow * 145 + D] = (z_at(a, g, p, myrow, D) + U_t) + Vd;
    if (col_owner) o[D * 145 + LCOL(0)] = (z_at(a, g, p, D, LCOL(0)) + Ud) + V[0];
    if (tid == 0) o[D * 145 + D] = (z_at(a, g, p, D, D) + Ud) + Vd;
#undef LROW
#undef LCOL
}

// ---------------------------------------------------------------------------------------------
// Level-2 kernel, 8-warp tiling (default for 145 x 145): warp w keeps the 18-column slab [18w, 18w+18) of the core,
// a 9 x 9 tile per lane (rows pr + 16k; columns 18w + qc + 2*(c ^ cmask(pr)) for slots c < 8 and 18w + qc + 16 for
// slot 8).  256 threads at <= 128 registers: TWO CTAs per SM (16 warps), so one CTA's barrier phases overlap the
// other's FFMA2 stream -- the 9-warp kernel above is limited to one CTA per SM by the per-scheduler register file.
// Slots 0..7 reduce with the select-free tree (twins pr, pr^8 own slot 0), slot 8 with a plain 4-step all-reduce.
// ---------------------------------------------------------------------------------------------
constexpr int C145B_W = 8, C145B_T = 256;

template <class Op>
__device__ __forceinline__ float allreduce_pr16(float v, Op op) {  // over pr (lane bits 1..4)
    v = op(v, __shfl_xor_sync(0xffffffffu, v, 2));
    v = op(v, __shfl_xor_sync(0xffffffffu, v, 4));
    v = op(v, __shfl_xor_sync(0xffffffffu, v, 8));
    return op(v, __shfl_xor_sync(0xffffffffu, v, 16));
}

__global__ void __launch_bounds__(C145B_T, 2) sinkhorn_c145b_kernel(SinkArgs a) {
    constexpr int D = 144;
    __shared__ float s_part[C145B_W][144];
    __shared__ float s_row[144];
    __shared__ float s_red[2][8];
    __shared__ float s_keep[144];
    __shared__ float s_fb[145 + 145 + 2 * C145B_T];
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int p = blockIdx.x;
    pdl_prologue();
    if (p >= a.b) return;
    const int pr = lane >> 1, qc = lane & 1;
    const int cmask = ((pr & 1) << 2) | (((pr >> 1) & 1) << 1) | ((pr >> 2) & 1);
    const bool col_owner = (pr & 8) == 0;   // counts the owned slot-0 column once per twin pair
    const bool col8_owner = pr == 0;        // counts the slot-8 column once per 16 lanes
    // rows are finished by the qc==0 lanes (row 16w + pr) and, for rows 128..143, the qc==1 lanes of warp 0
    const int myrow = qc == 0 ? 16 * w + pr : (w == 0 ? 128 + pr : -1);
    const bool row_thread = myrow >= 0;
    const Marg g = problem_marginals(a, p, lane);
#define LROW(k) (pr + 16 * (k))
#define LCOL(c) (18 * w + qc + 2 * ((c) ^ cmask))
#define LCOL8 (18 * w + qc + 16)
    auto red8 = [&](int which) {
        float t = s_red[which][0];
#pragma unroll
        for (int i = 1; i < C145B_W; ++i) t += s_red[which][i];
        return t;
    };
    auto max8 = [&](int which) {
        float t = s_red[which][0];
#pragma unroll
        for (int i = 1; i < C145B_W; ++i) t = fmaxf(t, s_red[which][i]);
        return t;
    };
    auto row_sum8 = [&](int row) {
        float t = s_part[0][row];
#pragma unroll
        for (int i = 1; i < C145B_W; ++i) t += s_part[i][row];
        return t;
    };

    // ---- load ------------------------------------------------------------------------------------------------------
    float z[9][9], zr, zr8, zc, zcorner;
    {
        const PlanRef pl = plan_ref(a, g, p);
        int coff[9];
#pragma unroll
        for (int c = 0; c < 8; ++c) coff[c] = LCOL(c);
        coff[8] = LCOL8;
#pragma unroll
        for (int k = 0; k < 9; ++k) {
            const int roff = LROW(k) * pl.stride;
#ifdef PATS_AB_C145_ZAT_LOAD
#pragma unroll
            for (int c = 0; c < 9; ++c) z[k][c] = z_at(a, g, p, LROW(k), coff[c]);
#else
#pragma unroll
            for (int c = 0; c < 9; ++c) z[k][c] = pl.core(roff, coff[c]);
#endif
        }
        zr = pl.edge(D, LCOL(0)), zr8 = pl.edge(D, LCOL8);
        zc = row_thread ? pl.edge(myrow, D) : 0.f;
        zcorner = pl.edge(D, D);
    }
    const float mu_t = row_thread ? expf(lmu_at(a, g, p, myrow)) : 0.f;
    const float nu_o = expf(lnu_at(a, g, p, LCOL(0))), nu8 = expf(lnu_at(a, g, p, LCOL8));
    const float mud = expf(lmu_at(a, g, p, D)), nud = expf(lnu_at(a, g, p, D));
    float v1_o = 0.f, v1_8 = 0.f, u1d = 0.f, v1d = 0.f, Dc = 0.f, Dr = 0.f, Dr8 = 0.f, corner = 0.f;

    // ---- iteration 1, exact in the log domain (the direct start of sinkhorn_w65x2_kernel was measured 3-4 % SLOWER here:
    //      the second entry path costs registers in the loop) ------------------------------------------------------------
    if (a.iters >= 1) {
        float u1[9], v1[8], u1_t = 0.f;
        {
            float mx[9];
#pragma unroll
            for (int k = 0; k < 9; ++k) {
                float m = z[k][0];
#pragma unroll
                for (int c = 1; c < 9; ++c) m = fmaxf(m, z[k][c]);
                m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 1));
                if (qc == 0) s_part[w][LROW(k)] = m;
            }
            const float drm_w = warp_max(fmaxf(zr, zr8));
            if (lane == 0) s_red[0][w] = drm_w;
            __syncthreads();
            float rm = 0.f;
            if (row_thread) {
                rm = s_part[0][myrow];
#pragma unroll
                for (int i = 1; i < C145B_W; ++i) rm = fmaxf(rm, s_part[i][myrow]);
                rm = finite_or_zero(fmaxf(rm, zc));
                s_row[myrow] = rm;
            }
            const float drm = finite_or_zero(fmaxf(max8(0), zcorner));
            __syncthreads();
#pragma unroll
            for (int k = 0; k < 9; ++k) mx[k] = s_row[LROW(k)];
#pragma unroll
            for (int k = 0; k < 9; ++k) {
                float sacc = 0.f;
#pragma unroll
                for (int c = 0; c < 9; ++c) sacc += fast_exp(z[k][c] - mx[k]);
                sacc += __shfl_xor_sync(0xffffffffu, sacc, 1);
                u1[k] = sacc;
            }
            __syncthreads();
#pragma unroll
            for (int k = 0; k < 9; ++k)
                if (qc == 0) s_part[w][LROW(k)] = u1[k];
            const float drs_w = warp_sum((col_owner ? fast_exp(zr - drm) : 0.f) + (col8_owner ? fast_exp(zr8 - drm) : 0.f));
            if (lane == 0) s_red[0][w] = drs_w;
            __syncthreads();
            if (row_thread) {
                const float sacc = row_sum8(myrow) + fast_exp(zc - rm);
                u1_t = lmu_at(a, g, p, myrow) - (fast_log(sacc) + rm);
                s_row[myrow] = u1_t;
                s_keep[myrow] = u1_t;
            }
            u1d = lmu_at(a, g, p, D) - (fast_log(red8(0) + fast_exp(zcorner - drm)) + drm);
            __syncthreads();
#pragma unroll
            for (int k = 0; k < 9; ++k) u1[k] = s_row[LROW(k)];
        }
        {
            float mx[8];
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                float m = z[0][c] + u1[0];
#pragma unroll
                for (int k = 1; k < 9; ++k) m = fmaxf(m, z[k][c] + u1[k]);
                mx[c] = m;
            }
            float m8 = z[0][8] + u1[0];
#pragma unroll
            for (int k = 1; k < 9; ++k) m8 = fmaxf(m8, z[k][8] + u1[k]);
            rs_c145(mx, OpMax());
            mx[0] = finite_or_zero(fmaxf(mx[0], zr + u1d));
            ag_c145(mx);
            m8 = finite_or_zero(fmaxf(allreduce_pr16(m8, OpMax()), zr8 + u1d));
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                float sacc = 0.f;
#pragma unroll
                for (int k = 0; k < 9; ++k) sacc += fast_exp((z[k][c] + u1[k]) - mx[c]);
                v1[c] = sacc;
            }
            float s8 = 0.f;
#pragma unroll
            for (int k = 0; k < 9; ++k) s8 += fast_exp((z[k][8] + u1[k]) - m8);
            rs_c145(v1, OpSum());
            v1[0] = lnu_at(a, g, p, LCOL(0)) - (fast_log(v1[0] + fast_exp((zr + u1d) - mx[0])) + mx[0]);
            ag_c145(v1);
            s8 = allreduce_pr16(s8, OpSum());
            v1_8 = lnu_at(a, g, p, LCOL8) - (fast_log(s8 + fast_exp((zr8 + u1d) - m8)) + m8);
            const float dcm_w = warp_max(row_thread ? zc + u1_t : -INFINITY);
            if (lane == 0) s_red[1][w] = dcm_w;
            __syncthreads();
            const float dcm = finite_or_zero(fmaxf(max8(1), zcorner + u1d));
            __syncthreads();
            const float dcs_w = warp_sum(row_thread ? fast_exp((zc + u1_t) - dcm) : 0.f);
            if (lane == 0) s_red[1][w] = dcs_w;
            __syncthreads();
            v1d = lnu_at(a, g, p, D) - (fast_log(red8(1) + fast_exp((zcorner + u1d) - dcm)) + dcm);
        }
#pragma unroll
        for (int k = 0; k < 9; ++k) {
#pragma unroll
            for (int c = 0; c < 8; ++c) z[k][c] = fast_exp((z[k][c] + u1[k]) + v1[c]);
            z[k][8] = fast_exp((z[k][8] + u1[k]) + v1_8);
        }
        Dc = row_thread ? fast_exp((zc + u1_t) + v1d) : 0.f;
        Dr = fast_exp((zr + u1d) + v1[0]);
        Dr8 = fast_exp((zr8 + u1d) + v1_8);
        v1_o = v1[0];
        corner = fast_exp((zcorner + u1d) + v1d);
    }
    __syncthreads();

    // ---- iterations 2..iters (same barrier structure as sinkhorn_c145_kernel).  Splitting the two dustbin butterflies over the
    //      back-edge as in sinkhorn_w65x2_kernel was measured 3-7 % SLOWER here (the carried partials cost registers in a loop
    //      that sits at the 128-register limit) --------------------------------------------------------------------------
    float2 Kp[9][4];
    float K8[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) {
#pragma unroll
        for (int h = 0; h < 4; ++h) Kp[k][h] = make_float2(z[k][2 * h], z[k][2 * h + 1]);
        K8[k] = z[k][8];
    }
    float be[8], be8 = 1.f;
#pragma unroll
    for (int c = 0; c < 8; ++c) be[c] = 1.f;
    float al_t = 1.f, ald = 1.f, bed = 1.f;
    {
        const float srp = warp_sum((col_owner ? Dr : 0.f) + (col8_owner ? Dr8 : 0.f));
        if (lane == 0) s_red[0][w] = srp;
    }
    float lo = INFINITY, hi = 0.f;

    for (int it = 1; it < a.iters; ++it) {
#pragma unroll
        for (int k = 0; k < 9; ++k) {
            float2 acc = make_float2(0.f, 0.f);
#pragma unroll
            for (int h = 0; h < 4; ++h) acc = ffma2(Kp[k][h], make_float2(be[2 * h], be[2 * h + 1]), acc);
            float r = fmaf(K8[k], be8, acc.x + acc.y);
            r += __shfl_xor_sync(0xffffffffu, r, 1);
            if (qc == 0) s_part[w][LROW(k)] = r;
        }
        __syncthreads();  // B1
        if (it > 1) bed = nud * fast_rcp(fmaf(corner, ald, red8(1)));
        ald = mud * fast_rcp(fmaf(corner, bed, red8(0)));
        if (row_thread) {
            al_t = mu_t * fast_rcp(fmaf(Dc, bed, row_sum8(myrow)));
            s_row[myrow] = al_t;
        }
        __syncthreads();  // B2
        {
            const float scp = warp_sum(row_thread ? Dc * al_t : 0.f);
            if (lane == 0) s_red[1][w] = scp;  // read after the next B1
        }
        float2 s2[4];
        float s8 = 0.f;
#pragma unroll
        for (int h = 0; h < 4; ++h) s2[h] = make_float2(0.f, 0.f);
#pragma unroll
        for (int k = 0; k < 9; ++k) {
            const float ak = s_row[LROW(k)];
            const float2 ak2 = make_float2(ak, ak);
#pragma unroll
            for (int h = 0; h < 4; ++h) s2[h] = ffma2(Kp[k][h], ak2, s2[h]);
            s8 = fmaf(K8[k], ak, s8);
        }
#pragma unroll
        for (int h = 0; h < 4; ++h) be[2 * h] = s2[h].x, be[2 * h + 1] = s2[h].y;
        rs_c145(be, OpSum());
        s8 = allreduce_pr16(s8, OpSum());
        be[0] = nu_o * fast_rcp(fmaf(Dr, ald, be[0]));
        be8 = nu8 * fast_rcp(fmaf(Dr8, ald, s8));
        {
            const float srp = warp_sum((col_owner ? Dr * be[0] : 0.f) + (col8_owner ? Dr8 * be8 : 0.f));
            if (lane == 0) s_red[0][w] = srp;  // read after the next B1
        }
        if ((it & 7) == 0 || it == a.iters - 1) {
            const float at = row_thread ? al_t : ald;
            lo = fminf(fminf(lo, at), fminf(fminf(ald, be[0]), fminf(be8, bed)));
            hi = fmaxf(fmaxf(hi, at), fmaxf(fmaxf(ald, be[0]), fmaxf(be8, bed)));
        }
        ag_c145(be);
    }
    __syncthreads();
    const bool ran = a.iters >= 2;  // the scaling loop ran at least once
    if (ran) bed = nud * fast_rcp(fmaf(corner, ald, red8(1)));

    // ---- potentials, health check, output ------------------------------------------------------------------------------
    const float shift = (a.mode == MODE_RAW) ? 0.f : g.norm;
    bool bad = !(lo >= 1e-13f && hi <= 1e13f) || (ran && !(bed >= 1e-13f && bed <= 1e13f));
    float U_t = 0.f, Ud = 0.f, Vd = -shift, V[8], V8;
    if (a.iters >= 1) U_t = row_thread ? s_keep[myrow] : 0.f, Ud = u1d, Vd = v1d - shift;
    if (ran) U_t += fast_log(al_t), Ud += fast_log(ald), Vd += fast_log(bed);
    if (row_thread && !(fabsf(U_t) < INFINITY)) bad = true;
    if (!(fabsf(Ud) < INFINITY) || !(fabsf(Vd) < INFINITY)) bad = true;
    {
        float tv = 0.f, t8 = 0.f;
        if (a.iters >= 1) tv = v1_o, t8 = v1_8;
        if (ran) tv += fast_log(be[0]), t8 += fast_log(be8);
        if (!(fabsf(tv) < INFINITY) || !(fabsf(t8) < INFINITY)) bad = true;
        V[0] = tv - shift;
        V8 = t8 - shift;
    }
    if (row_thread) s_row[myrow] = U_t;
    if (__syncthreads_or(bad ? 1 : 0)) {
        if (tid == 0 && a.fb_total) atomicAdd(a.fb_total, 1);
        log_domain_solve<C145B_T>(a, g, p, s_fb, s_fb + 145, s_fb + 290, tid, BlockSync());
        __syncthreads();
        if (tid == 0) publish_problem(a, p);
        return;
    }
    ag_c145(V);
    float *o = a.out + (size_t)p * 145 * 145;
    const PlanRef pl = plan_ref(a, g, opaque(p));  // recomputed: keeping the load-time copy alive costs registers in the loop
    {
        int coff[9];
#pragma unroll
        for (int c = 0; c < 8; ++c) coff[c] = LCOL(c);
        coff[8] = LCOL8;
#pragma unroll
        for (int k = 0; k < 9; ++k) {
            const int row = LROW(k);
            const int roff = row * pl.stride;
            const float Uk = s_row[row];
#ifdef PATS_AB_OUT_INTERLEAVE
#pragma unroll
            for (int c = 0; c < 8; ++c) o[row * 145 + coff[c]] = (pl.core(roff, coff[c]) + Uk) + V[c];
            o[row * 145 + coff[8]] = (pl.core(roff, coff[8]) + Uk) + V8;
#else
            float zz[9];
#pragma unroll
            for (int c = 0; c < 9; ++c) zz[c] = pl.core(roff, coff[c]);  // the nine loads of the row in flight together (L2 hits)
#pragma unroll
            for (int c = 0; c < 8; ++c) o[row * 145 + coff[c]] = (zz[c] + Uk) + V[c];
            o[row * 145 + coff[8]] = (zz[8] + Uk) + V8;
#endif
        }
    }
    const float ea = a.edge_add;  // second_layer.py:108-112: dustbin column += c, then dustbin row += c (x + 0.f == x bit for bit)
    if (row_thread) o[myrow * 145 + D] = ((pl.edge(myrow, D) + U_t) + Vd) + ea;
    if (col_owner) o[D * 145 + LCOL(0)] = ((pl.edge(D, LCOL(0)) + Ud) + V[0]) + ea;
    if (col8_owner) o[D * 145 + LCOL8] = ((pl.edge(D, LCOL8) + Ud) + V8) + ea;
    if (tid == 0) o[D * 145 + D] = (((pl.edge(D, D) + Ud) + Vd) + ea) + ea;
    if (a.done) {
        __syncthreads();
        if (tid == 0) publish_problem(a, p);
    }
#undef LROW
#undef LCOL
#undef LCOL8
}

// ---- host dispatch ------------------------------------------------------------------------------
using CfgTiny = RegCfg<1, 2, 4, 8, 4>;         // <= 32 x 32, one warp per problem, 4 problems per CTA
using CfgWarp = RegCfg<1, 2, 9, 17, 4>;        // <= 72 x 68  (level 3: 65 x 65)
using CfgCta = RegCfg<8, 4, 10, 10, 1>;        // <= 160 x 160 (level 2: 145 x 145), 16 x 16 threads
using CfgCl320 = RegCfg<8, 5, 5, 10, 1, 8>;    // <= 320 x 320 (level 1: 301 x 301), cluster of 8 CTAs x 256 threads
using CfgCl320h = RegCfg<8, 5, 5, 10, 1, 8, 1>;  // same tiling, one-hop all-to-all exchange (A/B variant 2)
using CfgCl320b = RegCfg<16, 5, 5, 10, 1, 4>;  // <= 320 x 320, cluster of 4 CTAs x 512 threads (A/B variant)
// <= 320 x 320, cluster of 10 CTAs x 256 threads, 4 x 10 tiles: 32 rows per CTA cover 301 rows without the 8-CTA tiling's idle
// row groups, and a column slice is exactly one warp wide.  Non-portable cluster size (one cluster per GPC at a time), so it is
// the default only for small batches.  Measured for one 301 x 301 problem through the wrapper: 121.8 us (8 x 256) -> 111.4 us;
// 16 CTAs x 256 (3 x 10 tiles) 118.8 us, 16 CTAs x 128 (6 x 10) 127.6 us.
using CfgCl320z = RegCfg<8, 5, 4, 10, 1, 10>;
using CfgCl512 = RegCfg<16, 5, 4, 16, 1, 8>;   // <= 512 x 512, cluster of 8 CTAs x 512 threads

static std::atomic<int> g_force_generic{0};
static std::atomic<int> g_cluster_variant{0};  // <= 320 x 320 plans: 0 = auto (10 x 256 for b <= 8, else 8 x 256), 1 = 4 CTAs x 512 threads, 2 = 8 x 256 one-hop exchange,
                                   // 3 = 8 x 256, 4 = 10 x 256
static std::atomic<int> g_disable_c145{0};  // 145 x 145 routing: 0 = 8-warp kernel, two CTAs per SM (default), 1 = padded 160 x 160 CTA kernel,
                                //                    2 = 9-warp kernel
static std::atomic<int> g_disable_w65{0};  // 65 x 65 routing: 0 = two warps per problem (default), 1 = padded 72 x 68 warp kernel,
                               //                  2 / 3 = one-warp 65 x 65 kernel at 2 / 3 CTAs per SM (tests / A-B timing)
static std::atomic<int> g_bulk_staging{1};  // pats_sinkhorn_bulk_staging(): 65 x 65 problems staged by cp.async.bulk (persistent CTAs); 0 = direct loads
static std::atomic<int> g_fp_exit{1};  // pats_sinkhorn_fixed_point_exit(): 65 x 65 kernel leaves its loop at a bitwise fixed point (results identical)
static int *g_fb_total[kMaxDevices];  // per device: counter of problems that took the log-domain fallback
static std::mutex g_mu;

static int ensure_counter(int **counter) {
    const int dev = current_device();
    if (dev < 0) return PATS_E_CUDA;
    std::lock_guard<std::mutex> lk(g_mu);
    if (!g_fb_total[dev]) {
        PATS_CUDA_TRY(cudaMalloc(&g_fb_total[dev], 2 * sizeof(int)));  // [0] fallbacks, [1] iterations skipped by the fixed-point exit
        PATS_CUDA_TRY(cudaMemset(g_fb_total[dev], 0, 2 * sizeof(int)));
    }
    *counter = g_fb_total[dev];
    return PATS_OK;
}

// ---- plan hand-over: per-stream flag pool (see sinkhorn_common.cuh) -----------------------------------------------------
namespace {
struct HandOver {
    unsigned *flags = nullptr;  // [cap] per stream, zero-initialised; epochs start at 1 and only grow
    size_t cap = 0;
    unsigned epoch = 0;
};
std::mutex g_ho_mu;
std::map<std::pair<int, cudaStream_t>, HandOver> g_ho;  // keyed by (device, stream): the default stream is 0 on every device

// flags for a launch of b problems on `st`; nullptr (hand-over off) if the pool cannot be grown
unsigned *handover_begin(cudaStream_t st, int b, unsigned *epoch) {
    const int dev = current_device();
    if (dev < 0) return nullptr;
    cudaStreamCaptureStatus cap_st = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(st, &cap_st) != cudaSuccess || cap_st != cudaStreamCaptureStatusNone) {
        cudaGetLastError();
        return nullptr;  // under stream capture: no pool growth (it synchronises), no spinning consumers -- plain stream order
    }
    std::lock_guard<std::mutex> lk(g_ho_mu);
    HandOver &h = g_ho[std::make_pair(dev, st)];
    if (h.cap < (size_t)b) {
        if (h.flags) {
            cudaStreamSynchronize(st);
            cudaFree(h.flags);
        }
        const size_t cap = (size_t)b * 2 < 4096 ? 4096 : (size_t)b * 2;
        h.flags = nullptr, h.cap = 0;
        if (cudaMalloc(&h.flags, cap * sizeof(unsigned)) != cudaSuccess || cudaMemset(h.flags, 0, cap * sizeof(unsigned)) != cudaSuccess) {
            h.flags = nullptr;
            cudaGetLastError();
            return nullptr;
        }
        h.cap = cap;
    }
    h.epoch += 1;
    if (h.epoch == 0) h.epoch = 1;  // 0 is the "never written" value
    *epoch = h.epoch;
    return h.flags;
}
}  // namespace

// dustbin column += add, then dustbin row += add (the corner gets both, in that order)
__global__ void edge_add_kernel(float *out, int b, int M, int N, float add) {
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int per = M + N;
    if (e >= (long long)b * per) return;
    const int p = (int)(e / per), t = (int)(e - (long long)p * per);
    float *o = out + (size_t)p * M * N;
    if (t < M - 1) o[(size_t)t * N + (N - 1)] += add;                        // dustbin column, rows 0..M-2
    else if (t == M - 1) o[(size_t)t * N + (N - 1)] = (o[(size_t)t * N + (N - 1)] + add) + add;  // corner
    else if (t - M < N - 1) o[(size_t)(M - 1) * N + (t - M)] += add;         // dustbin row, columns 0..N-2
}

static int kernel_kind(int M, int N) {
    if (g_force_generic) return 2;
    if (M <= CfgWarp::MAXM && N <= CfgWarp::MAXN) return 0;
    if (M <= CfgCta::MAXM && N <= CfgCta::MAXN) return 1;
    if (M <= CfgCl512::MAXM && N <= CfgCl512::MAXN) return 3;
    if (grid_plan_supported(M, N)) return 4;
    return 2;
}

template <class C>
static int launch_reg(const SinkArgs &a, cudaStream_t st) {
    constexpr size_t smem = Smem<C>::BYTES;
    static_assert(smem <= 200 * 1024, "shared-memory layout too large");
    const int dev = current_device();
    if (dev < 0) return PATS_E_CUDA;
    if (smem > 48 * 1024) {
        static PerDeviceOnce configured;  // per kernel instantiation and device
        if (!configured.done(dev)) {
            PATS_CUDA_TRY(cudaFuncSetAttribute(sinkhorn_reg_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            configured.mark(dev);
        }
    }
    if (C::CL > 8) {  // beyond the portable cluster size
        static PerDeviceOnce allowed;
        if (!allowed.done(dev)) {
            PATS_CUDA_TRY(cudaFuncSetAttribute(sinkhorn_reg_kernel<C>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
            allowed.mark(dev);
        }
    }
    cudaLaunchConfig_t cfg = {};
    cfg.blockDim = dim3(C::THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[2];
    int na = 0;
    if (g_chain) {  // launch chaining (common.cuh): the kernel starts with pdl_prologue()
        attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[na].val.programmaticStreamSerializationAllowed = 1;
        ++na;
    }
    if (C::CL > 1) {
        cfg.gridDim = dim3((unsigned)a.b * C::CL);
        attr[na].id = cudaLaunchAttributeClusterDimension;
        attr[na].val.clusterDim.x = C::CL;
        attr[na].val.clusterDim.y = 1;
        attr[na].val.clusterDim.z = 1;
        ++na;
    } else {
        cfg.gridDim = dim3((unsigned)((a.b + C::GROUPS - 1) / C::GROUPS));
    }
    cfg.attrs = attr;
    cfg.numAttrs = na;
    PATS_CUDA_TRY(cudaLaunchKernelEx(&cfg, sinkhorn_reg_kernel<C>, a));
    return PATS_OK;
}

int launch_generic(const SinkArgs &a, cudaStream_t st) {
    constexpr int GT = 1024;
    const size_t smem = sizeof(float) * ((size_t)a.M + a.N + 2 * GT);
    if (smem > 200 * 1024) return invalid("sinkhorn: M+N = %d exceeds the shared-memory budget of the generic kernel", a.M + a.N);
    if (smem > 48 * 1024) {
        PATS_CUDA_TRY(cudaFuncSetAttribute(sinkhorn_generic_kernel<GT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    }
    int grid = a.b;
    const int cap = 4 * (sm_count() > 0 ? sm_count() : 148);
    if (grid > cap) grid = cap;
    sinkhorn_generic_kernel<GT><<<grid, GT, smem, st>>>(a);
    PATS_LAUNCH_CHECK("sinkhorn_generic_kernel");
    return PATS_OK;
}

// publish: ask the kernel to raise per-problem "plan complete" flags (plan hand-over); *done stays nullptr when the kernel
// chosen for the shape does not publish.
// Can this device schedule the 10-CTA cluster shape (a non-portable cluster size)?  Asked once; any error or "zero clusters"
// keeps the portable 8-CTA shape.
static bool cluster10_ok() {
    static int ok_dev[kMaxDevices];  // per device: 0 = not asked, 1 = yes, 2 = no
    const int dev = current_device();
    if (dev < 0) return false;
    int ok = ok_dev[dev] == 0 ? -1 : (ok_dev[dev] == 1 ? 1 : 0);
    if (ok < 0) {
        using C = CfgCl320z;
        int n = 0;
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(C::CL);
        cfg.blockDim = dim3(C::THREADS);
        cfg.dynamicSmemBytes = Smem<C>::BYTES;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = C::CL;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        cudaError_t e = cudaSuccess;
        if (Smem<C>::BYTES > 48 * 1024)
            e = cudaFuncSetAttribute(sinkhorn_reg_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Smem<C>::BYTES);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(sinkhorn_reg_kernel<C>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
        if (e == cudaSuccess) e = cudaOccupancyMaxActiveClusters(&n, sinkhorn_reg_kernel<C>, &cfg);
        if (e != cudaSuccess) (void)cudaGetLastError();  // not sticky: clear it, the 8-CTA shape runs instead
        ok = (e == cudaSuccess && n >= 1) ? 1 : 0;
        ok_dev[dev] = ok ? 1 : 2;
    }
    return ok == 1;
}

static int run_sinkhorn(SinkArgs a, void *stream, bool publish = false, const unsigned **done = nullptr, unsigned *epoch = nullptr) {
    struct Report {  // hands the flags of the launch (if any) back on every return path
        SinkArgs &a;
        const unsigned **done;
        unsigned *epoch;
        ~Report() {
            if (done) *done = a.done, *epoch = a.epoch;
        }
    } report{a, done, epoch};
    if (a.b < 0 || a.M <= 0 || a.N <= 0 || a.iters < 0) return invalid("sinkhorn: bad sizes b=%d M=%d N=%d iters=%d", a.b, a.M, a.N, a.iters);
    if (a.mode != MODE_RAW && (a.M < 2 || a.N < 2)) return invalid("optimal transport needs at least one real row and column");
    if (a.b == 0) return PATS_OK;
    if (!a.Z || !a.out) return invalid("sinkhorn: null pointer");
    if ((long long)a.M * a.N > 0x7fffffffLL / 2) return invalid("sinkhorn: problem too large");
    int rc = ensure_counter(&a.fb_total);
    if (rc) return rc;
    a.fp_exit = g_fp_exit;
    cudaStream_t st = as_stream(stream);
    const bool c145b = kernel_kind(a.M, a.N) == 1 && a.M == 145 && a.N == 145 && g_disable_c145 == 0;
    if (a.edge_add != 0.f && !c145b && kernel_kind(a.M, a.N) != 2) {  // kernels without the edge epilogue: solve, then one more pass
        const float add = a.edge_add;
        a.edge_add = 0.f;
        rc = run_sinkhorn(a, stream);
        if (rc) return rc;
        const long long cells = (long long)a.b * (a.M + a.N);
        edge_add_kernel<<<(unsigned)((cells + 255) / 256), 256, 0, st>>>(a.out, a.b, a.M, a.N, add);
        PATS_LAUNCH_CHECK("edge_add_kernel");
        return PATS_OK;
    }
    switch (kernel_kind(a.M, a.N)) {
        case 0:
            if (a.M == 65 && a.N == 65 && g_disable_w65 == 0) {
                if (publish) a.done = handover_begin(st, a.b, &a.epoch);
                // bulk-copy staging (pats_sinkhorn_bulk_staging): contiguous [b,65,65] problems whose input and output share their
                // 16-byte phase; a trailing problem whose aligned superset would reach past the tensor goes through the direct kernel
                const bool bulk_ok = g_bulk_staging && X2_PAIRS == 1 && a.mode != MODE_OT && (reinterpret_cast<uintptr_t>(a.Z) & 15) == 0 &&
                                     (reinterpret_cast<uintptr_t>(a.out) & 15) == 0;
                if (bulk_ok) {
                    const int tail = (a.b % 4 != 0) ? 1 : 0, main_b = a.b - tail;
                    if (main_b > 0) {
                        SinkArgs m = a;
                        m.b = main_b;
                        const int sms = sm_count() > 0 ? sm_count() : 148;
                        const int grid = main_b < sms * X2_MIN_CTAS ? main_b : sms * X2_MIN_CTAS;
                        PATS_CUDA_TRY(launch_chained(sinkhorn_w65x2_kernel<true>, dim3(grid), dim3(64), 0, st, m));
                    }
                    if (tail) {
                        SinkArgs t = a;  // the last problem alone, flags and all
                        const size_t off = (size_t)main_b;
                        t.Z = a.Z + off * 4225, t.out = a.out + off * 4225, t.ns = a.ns ? a.ns + off * 64 : nullptr;
                        t.log_mu = a.log_mu ? a.log_mu + off * 65 : nullptr, t.log_nu = a.log_nu ? a.log_nu + off * 65 : nullptr;
                        t.done = a.done ? a.done + off : nullptr;
                        t.b = 1;
                        PATS_CUDA_TRY(launch_chained(sinkhorn_w65x2_kernel<false>, dim3(1), dim3(X2_PAIRS * 64), 0, st, t));
                    }
                    return PATS_OK;
                }
                PATS_CUDA_TRY(launch_chained(sinkhorn_w65x2_kernel<false>, dim3((a.b + X2_PAIRS - 1) / X2_PAIRS), dim3(X2_PAIRS * 64), 0, st, a));
                return PATS_OK;
            }
            if (a.M == 65 && a.N == 65 && (g_disable_w65 == 2 || g_disable_w65 == 3)) {
                if (g_disable_w65 == 2) sinkhorn_w65_kernel<2><<<(a.b + W65_WARPS - 1) / W65_WARPS, W65_WARPS * 32, 0, st>>>(a);
                else sinkhorn_w65_kernel<3><<<(a.b + W65_WARPS - 1) / W65_WARPS, W65_WARPS * 32, 0, st>>>(a);
                PATS_LAUNCH_CHECK("sinkhorn_w65_kernel");
                return PATS_OK;
            }
            if (a.M <= CfgTiny::MAXM && a.N <= CfgTiny::MAXN) return launch_reg<CfgTiny>(a, st);
            return launch_reg<CfgWarp>(a, st);
        case 1:
            if (a.M == 145 && a.N == 145 && g_disable_c145 == 0) {
                if (publish) a.done = handover_begin(st, a.b, &a.epoch);
                PATS_CUDA_TRY(launch_chained(sinkhorn_c145b_kernel, dim3(a.b), dim3(C145B_T), 0, st, a));
                return PATS_OK;
            }
            if (a.M == 145 && a.N == 145 && g_disable_c145 == 2) {
                sinkhorn_c145_kernel<<<a.b, C145_T, 0, st>>>(a);
                PATS_LAUNCH_CHECK("sinkhorn_c145_kernel");
                return PATS_OK;
            }
            return launch_reg<CfgCta>(a, st);
        case 3:
            if (a.M <= CfgCl320::MAXM && a.N <= CfgCl320::MAXN)
                return g_cluster_variant == 1   ? launch_reg<CfgCl320b>(a, st)
                       : g_cluster_variant == 2 ? launch_reg<CfgCl320h>(a, st)
                       : g_cluster_variant == 3 ? launch_reg<CfgCl320>(a, st)
                       : g_cluster_variant == 4 ? launch_reg<CfgCl320z>(a, st)
                       : (a.b <= 8 && cluster10_ok()) ? launch_reg<CfgCl320z>(a, st)   // auto: 10-CTA clusters while every GPC can host one
                                                      : launch_reg<CfgCl320>(a, st);
            return launch_reg<CfgCl512>(a, st);
        case 4:
            return launch_grid(a, st);
        default:
            return launch_generic(a, st);
    }
}

}  // namespace pats

using namespace pats;

PATS_API int pats_log_sinkhorn_iterations_f32(const float *Z, const float *log_mu, const float *log_nu, int b, int M, int N,
                                              int iters, float *out, void *stream) {
    if (b > 0 && (!log_mu || !log_nu)) return invalid("log_sinkhorn_iterations: null marginals");
    SinkArgs a{Z, log_mu, log_nu, nullptr, nullptr, out, b, M, N, iters, MODE_RAW, nullptr, nullptr, 0u, 0.f};
    return run_sinkhorn(a, stream);
}

PATS_API int pats_log_optimal_transport_f32(const float *scores, const float *alpha, const float *ns, int b, int m, int n,
                                            int iters, float *out, void *stream) {
    if (b > 0 && (!alpha || !ns)) return invalid("log_optimal_transport: null alpha / ns");
    SinkArgs a{scores, nullptr, nullptr, alpha, ns, out, b, m + 1, n + 1, iters, MODE_OT, nullptr, nullptr, 0u, 0.f};
    return run_sinkhorn(a, stream);
}

PATS_API int pats_log_optimal_transport2_f32(const float *scores, const float *one, const float *ns, int b, int m, int n,
                                             int iters, float *out, void *stream) {
    if (b > 0 && (!one || !ns)) return invalid("log_optimal_transport2: null one / ns");
    SinkArgs a{scores, nullptr, nullptr, one, ns, out, b, m, n, iters, MODE_OT2, nullptr, nullptr, 0u, 0.f};
    return run_sinkhorn(a, stream);
}

PATS_API int pats_sinkhorn_kernel_kind(int M, int N) { return kernel_kind(M, N); }
PATS_API void pats_sinkhorn_force_generic(int on) { g_force_generic = on ? 1 : 0; }
PATS_API void pats_plan_handover(int on) { g_handover = on ? 1 : 0; }
PATS_API void pats_launch_chaining(int on) { g_chain = on ? 1 : 0; }

namespace pats {
std::atomic<int> g_handover{1};
std::atomic<int> g_chain{1};
int sinkhorn_ot2_publish(const float *scores, const float *one, const float *ns, int b, int m, int n, int iters, float edge_add,
                         float *out, cudaStream_t st, const unsigned **done, unsigned *epoch) {
    if (b > 0 && (!one || !ns)) return invalid("log_optimal_transport2: null one / ns");
    SinkArgs a{scores, nullptr, nullptr, one, ns, out, b, m, n, iters, MODE_OT2, nullptr, nullptr, 0u, edge_add};
    *done = nullptr, *epoch = 0u;
    return run_sinkhorn(a, st, g_handover != 0, done, epoch);
}
}  // namespace pats
PATS_API void pats_sinkhorn_cluster_variant(int v) { g_cluster_variant = (v >= 0 && v <= 4) ? v : 0; }
PATS_API void pats_sinkhorn_disable_c145(int mode) { g_disable_c145 = (mode >= 0 && mode <= 2) ? mode : 0; }
PATS_API void pats_sinkhorn_disable_w65(int mode) { g_disable_w65 = (mode >= 0 && mode <= 3) ? mode : 0; }

PATS_API void pats_sinkhorn_fixed_point_exit(int on) { g_fp_exit = on ? 1 : 0; }
PATS_API void pats_sinkhorn_bulk_staging(int on) { g_bulk_staging = on ? 1 : 0; }

PATS_API long long pats_sinkhorn_iterations_skipped(int reset) {
    int *counter = nullptr;  // of the current device
    if (ensure_counter(&counter) != PATS_OK) return -1;
    int v = 0;
    if (cudaMemcpy(&v, counter + 1, sizeof(int), cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
    if (reset) cudaMemset(counter + 1, 0, sizeof(int));
    return (long long)v;
}

PATS_API int pats_sinkhorn_fallback_count(int reset) {
    int *counter = nullptr;  // of the current device
    if (ensure_counter(&counter) != PATS_OK) return -1;
    int v = 0;
    if (cudaMemcpy(&v, counter, sizeof(int), cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
    if (reset) cudaMemset(counter, 0, sizeof(int));
    return v;
}

// ---- host-buffer (end-to-end) variants -------------------------------------------------------------
namespace {
struct DevBuf {
    void *p = nullptr;
    ~DevBuf() {
        if (p) cudaFree(p);
    }
    int alloc(size_t n) { return cudaMalloc(&p, n ? n : 1) == cudaSuccess ? 0 : -1; }
};

int ot_host(int mode, const float *scores, float scalar, const float *ns, int b, int m, int n, int iters, float *out) {
    if (b < 0 || m <= 0 || n <= 0) return invalid("optimal transport (host): bad sizes");
    if (b == 0) return PATS_OK;
    if (!scores || !ns || !out) return invalid("optimal transport (host): null pointer");
    const int M = (mode == MODE_OT) ? m + 1 : m, N = (mode == MODE_OT) ? n + 1 : n;
    const size_t nin = (size_t)b * m * n, nns = (size_t)b * (N - 1), nout = (size_t)b * M * N;
    DevBuf d_in, d_ns, d_sc, d_out;
    if (d_in.alloc(nin * 4) || d_ns.alloc(nns * 4) || d_sc.alloc(4) || d_out.alloc(nout * 4))
        return cuda_fail(cudaGetLastError(), "cudaMalloc (host variant)");
    cudaStream_t st = nullptr;
    PATS_CUDA_TRY(cudaMemcpyAsync(d_in.p, scores, nin * 4, cudaMemcpyHostToDevice, st));
    PATS_CUDA_TRY(cudaMemcpyAsync(d_ns.p, ns, nns * 4, cudaMemcpyHostToDevice, st));
    PATS_CUDA_TRY(cudaMemcpyAsync(d_sc.p, &scalar, 4, cudaMemcpyHostToDevice, st));
    int rc = (mode == MODE_OT)
                 ? pats_log_optimal_transport_f32((const float *)d_in.p, (const float *)d_sc.p, (const float *)d_ns.p, b, m, n, iters, (float *)d_out.p, st)
                 : pats_log_optimal_transport2_f32((const float *)d_in.p, (const float *)d_sc.p, (const float *)d_ns.p, b, m, n, iters, (float *)d_out.p, st);
    if (rc) return rc;
    PATS_CUDA_TRY(cudaMemcpyAsync(out, d_out.p, nout * 4, cudaMemcpyDeviceToHost, st));
    PATS_CUDA_TRY(cudaStreamSynchronize(st));
    return PATS_OK;
}
}  // namespace

PATS_API int pats_log_optimal_transport_f32_host(const float *scores, float alpha, const float *ns, int b, int m, int n,
                                                 int iters, float *out) {
    return ot_host(MODE_OT, scores, alpha, ns, b, m, n, iters, out);
}

PATS_API int pats_log_optimal_transport2_f32_host(const float *scores, float one, const float *ns, int b, int m, int n,
                                                  int iters, float *out) {
    return ot_host(MODE_OT2, scores, one, ns, b, m, n, iters, out);
}
