// Area expansion, window regrouping, match assembly and the third-layer result for sm_100a.
// Replaces, from zju3dv/pats:
//   utils/utils.py:1179-1297  Iterative_expand_matrix (+ :1321-1340 Compute_scaling, est_position's argmaxes)
//   models/second_layer.py:137-238  merge_patches_old / merge_patches_new
//   utils/utils.py:189-213    get_result
//   models/third_layer.py:184-217, :166-167  ThirdLayer.Compute_result + the label test
//
// These kernels produce the integers (boxes, masks, match order) that define match-index parity, so they
// are written to be bit-identical to the CPU oracle: short f32 sums run sequentially in the oracle's
// order, long sums accumulate in f64, and this file is compiled with -fmad=false so every f32 expression
// rounds exactly like the C restatement.  They are latency-bound index kernels; the work is spread over
// (rows | window cells | matches) >> 148 SMs worth of warps.
#include <map>
#include <mutex>

#include "common.cuh"

namespace pats {

// Per-stream scratch that is all-zero whenever no kernel is using it: zeroed at allocation, and every user restores
// the zeros before it finishes (est_position's column maxima / arrival counters).
static void *zeroed_workspace(cudaStream_t st, size_t bytes) {
    static std::mutex mu;
    static std::map<std::pair<int, cudaStream_t>, std::pair<void *, size_t>> pool;  // (device, stream): stream 0 exists on every device
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
        e.first = nullptr, e.second = 0;
        if (cudaMalloc(&e.first, cap) != cudaSuccess) {
            e.first = nullptr;
            return nullptr;
        }
        if (cudaMemsetAsync(e.first, 0, cap, st) != cudaSuccess) {
            cudaFree(e.first);
            e.first = nullptr;
            return nullptr;
        }
        e.second = cap;
    }
    return e.first;
}

// =================================================================================================
// a8/a9  area expansion: one half-warp (16 lanes) per (problem, source row), 16 rows per CTA
// =================================================================================================
#ifndef EX_ROWS_N
#define EX_ROWS_N 16
#endif
constexpr int EX_ROWS = EX_ROWS_N;  // rows (half-warps) per CTA
constexpr int EX_SM = 5;         // shared per-CTA arrays in front of the rows: O, SX, SY, CM, SV
constexpr float kZero = 1e-14f;  // `zero` of utils/utils.py:1203

struct ExpandArgs {
    const float *scores;  // [b, m+1, n+1] = exp(Z)  (or Z when log_input)
    const float *sx, *sy; // [b, n]
    int b, m, n, grid_w, width, height, iters, log_input;  // log_input: scores hold Z, exp() applied on load
    float lb;
    float *whole, *core, *avg, *xs, *ys;
    int64_t *bound;
    uint8_t *nomatch;
    // fused column-argmax mask of est_position (log_input only): nm2[j] = (argmax_i Z[i][j] == dustbin row)
    uint8_t *nm2;
    unsigned *colmax;   // [b, n] order-preserving encoding of max_i<m Z[i][j]; zero on entry, zeroed again by the last CTA
    unsigned *counter;  // [b] CTAs finished; zero on entry, zeroed again by the last CTA
    // plan hand-over (sinkhorn_common.cuh): when set, problem bb's plan is ready once done[bb] == epoch
    const unsigned *done;
    unsigned epoch;
};

// monotone map float -> unsigned (atomicMax on floats of either sign); every key is > 0
__device__ __forceinline__ unsigned enc_f32(float f) {
    const unsigned b = __float_as_uint(f);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}

__device__ __forceinline__ double half_sum_f64(double v) {
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// WIDTH > 0: the strip length (= a.width) at compile time, so the strip loops unroll completely -- the shared-memory
// loads of a strip are then in flight together and only the additions stay sequential (the order is the oracle's).
// WIDTH == 0: any grid.
template <int WIDTH>
__global__ void __launch_bounds__(EX_ROWS * 16) area_expand_kernel(ExpandArgs a) {
    extern __shared__ float sm[];
    __shared__ bool s_last;
    const int n = a.n, m = a.m, bb = blockIdx.y;
    const int stride = n + 2;
    float *O = sm, *SX = sm + stride, *SY = sm + 2 * stride;  // dustbin row, target scales (shared by the CTA)
    unsigned *CM = reinterpret_cast<unsigned *>(sm + 3 * stride);  // per-CTA column maxima (encoded)
    float *SV = sm + 4 * stride;                                   // SX * SY, the product every strip sum reads
    const int hl = threadIdx.x & 15, half = threadIdx.x >> 4;      // lane within the half-warp, row slot
    const int hbase = threadIdx.x & 16;                            // first lane of this half inside its warp
    float *E = sm + (EX_SM + half) * stride;                          // this row (+ dustbin col, + zero slot)
    const float *opp = a.scores + ((size_t)bb * (m + 1) + m) * (n + 1);
    if (a.done) {  // launched early (plan hand-over): wait until the Sinkhorn kernel has published this problem
        pdl_launch_dependents();
        if (threadIdx.x == 0) await_problem(a.done, a.epoch, bb);
        __syncthreads();
    } else {
        pdl_prologue();
    }
    // the plan is read with ld.global.cg: it may have been written by a still-running producer grid
    for (int j = threadIdx.x; j < stride; j += blockDim.x) {
        O[j] = j < n ? (a.log_input ? expf(__ldcg(opp + j)) : __ldcg(opp + j)) : kZero;
        const float sxv = j < n ? a.sx[(size_t)bb * n + j] : kZero;  // slots n, n+1: SX*SY == 1e-14 exactly (the `zero` padding of
        const float syv = j < n ? a.sy[(size_t)bb * n + j] : 1.0f;   // expand_scale, utils.py:1208-1209)
        SX[j] = sxv, SY[j] = syv, SV[j] = sxv * syv;
        CM[j] = 0u;
    }
    __syncthreads();
    const int i_raw = blockIdx.x * EX_ROWS + half;
    const bool row_ok = i_raw < m;
    const int i = row_ok ? i_raw : m - 1;  // surplus half-warps recompute the last row and discard it (keeps shuffles full)
    const float *row = a.scores + ((size_t)bb * (m + 1) + i) * (n + 1);
    for (int j = hl; j < stride; j += 16) E[j] = j <= n ? __ldcg(row + j) : kZero;  // raw values (Z when log_input)
    __syncthreads();
    if (a.nm2) {
        // this CTA's column maxima of the raw Z: a thread per column over the staged rows (no shared-memory atomics --
        // sixteen half-warps hammering the same 144 words cost a quarter of the kernel), then one global atomic each
        const int rows_here = min(EX_ROWS, m - (int)blockIdx.x * EX_ROWS);
        for (int j = threadIdx.x; j < n; j += blockDim.x) {
            unsigned best = 0u;
            for (int r = 0; r < rows_here; ++r) best = max(best, enc_f32(sm[(EX_SM + r) * stride + j]));
            CM[j] = best;
        }
        __syncthreads();
    }
    if (a.log_input) {
        for (int j = hl; j <= n; j += 16) E[j] = expf(E[j]);  // each half-warp its own row
    }
    __syncthreads();
    if (a.nm2) {  // publish this CTA's column maxima; the last CTA of the problem writes the mask
        for (int j = threadIdx.x; j < n; j += blockDim.x) atomicMax(&a.colmax[(size_t)bb * n + j], CM[j]);
        __threadfence();
        __syncthreads();
        if (threadIdx.x == 0) s_last = (atomicAdd(&a.counter[bb], 1u) == gridDim.x - 1);
        __syncthreads();
        if (s_last) {
            __threadfence();
            for (int j = threadIdx.x; j < n; j += blockDim.x) {
                const unsigned best = *reinterpret_cast<volatile unsigned *>(&a.colmax[(size_t)bb * n + j]);
                a.nm2[(size_t)bb * n + j] = enc_f32(__ldcg(opp + j)) > best;  // strictly larger: ties go to the first (real) row
                a.colmax[(size_t)bb * n + j] = 0u;  // leave the scratch zeroed for the next call (no memset between the
            }                                       // Sinkhorn kernel and this one: it would serialise the early launch)
            if (threadIdx.x == 0) a.counter[bb] = 0u;
        }
    }
    auto Sval = [&](int idx) -> float { return SV[idx]; };
    const float lbv = a.lb;

    // ---- argmax over real targets (first maximum), dustbin test (utils.py:1182,1194) ------------------------
    float bv = -INFINITY;
    int bi = 0x7fffffff;
    for (int j = hl; j < n; j += 16) {
        const float v = E[j];
        if (bi == 0x7fffffff || v > bv) bv = v, bi = j;
    }
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (oi != 0x7fffffff && (bi == 0x7fffffff || ov > bv || (ov == bv && oi < bi))) bv = ov, bi = oi;
    }
    const int max0 = bi;
    const int maxall = (E[n] > E[max0]) ? n : max0;
    const bool nomatch = (maxall == m);

    int bd0, bd1, bd2, bd3, dy = 0, dx = 0, sdy = 0, sdx = 0;
    bd0 = bd1 = max0 / a.grid_w;
    bd2 = bd3 = max0 % a.grid_w;
    const int width = WIDTH > 0 ? WIDTH : a.width, height = a.height;
    float last_sum = E[max0], last_nom = O[max0];

    // ---- box growth (utils.py:1213-1243): lanes 0..11 of the half = (direction, quantity); the strip is summed
    //      sequentially in the reference's order ------------------------------------------------------------------
    const int d12 = (hl / 3) & 3, q12 = hl - 3 * (hl / 3);  // lanes 12..15: d12 wraps to 0 (discarded)
    for (int it = 0; it < a.iters; ++it) {
        sdy = dy, sdx = dx;
        float acc = 0.f;
        {
            // Every lane runs the same instruction stream (lanes 12..15 of the half compute a discarded copy of lane 0..3).
            // Inside the extent (t <= lim) ranges[] holds t, so the f32 index arithmetic of the reference is exact and
            // equals off + t*step in integers; beyond it ranges[] holds 1e7 -> slot n+1 -> every quantity adds `zero`.
            int off, lim;
            if (d12 == 0) off = bd2 + bd0 * width - width, lim = dx;
            else if (d12 == 1) off = bd2 + bd1 * width + width, lim = dx;
            else if (d12 == 2) off = bd2 + bd0 * width - 1, lim = dy;
            else off = bd3 + bd0 * width + 1, lim = dy;
            const int stepi = (d12 < 2) ? 1 : width;
            const float *const A12 = (q12 == 2) ? SV : E;
            if (WIDTH > 0) {
                // beyond the extent every quantity adds `zero` (slot n+1 holds it for E, O and SX*SY), so one loop of WIDTH
                // steps is the same sequence of additions as the extent loop followed by the padding loop below
#pragma unroll
                for (int t = 0; t < WIDTH; ++t) {
                    int sidx = off + t * stepi;
                    if (sidx < 0 || sidx > n - 1 || t > lim) sidx = n + 1;
                    const float x = A12[sidx], o = O[sidx];  // x: E for the plain / thresholded sums, SX*SY for the scale sum
                    acc += (q12 == 1) ? ((x > lbv) ? o : kZero) : x;
                }
            } else {
                const int tmax = min(max(dx, dy) + 1, width);  // uniform over the half-warp
                for (int t = 0; t < tmax; ++t) {
                    int sidx = off + t * stepi;
                    if (sidx < 0 || sidx > n - 1 || t > lim) sidx = n + 1;
                    const float x = A12[sidx], o = O[sidx];  // x: E for the plain / thresholded sums, SX*SY for the scale sum
                    acc += (q12 == 1) ? ((x > lbv) ? o : kZero) : x;
                }
                for (int t = tmax; t < width; ++t) acc += kZero;
            }
        }
        float es0 = __shfl_sync(0xffffffffu, acc, hbase + 0), es1 = __shfl_sync(0xffffffffu, acc, hbase + 3);
        float es2 = __shfl_sync(0xffffffffu, acc, hbase + 6), es3 = __shfl_sync(0xffffffffu, acc, hbase + 9);
        if (bd0 == 0) es0 = kZero;
        if (bd1 == height - 1) es1 = kZero;
        if (bd2 == 0) es2 = kZero;
        if (bd3 == width - 1) es3 = kZero;
        int arg = 0;
        float mx = es0;
        if (es1 > mx) mx = es1, arg = 1;
        if (es2 > mx) mx = es2, arg = 2;
        if (es3 > mx) mx = es3, arg = 3;
        const float en_arg = __shfl_sync(0xffffffffu, acc, hbase + arg * 3 + 1);
        float add_sum = kZero, add_nom = kZero;
        if (mx > a.lb) {
            if (arg == 0) bd0 -= 1;
            else if (arg == 1) bd1 += 1;
            else if (arg == 2) bd2 -= 1;
            else bd3 += 1;
            add_sum = mx, add_nom = en_arg;
        }
        dy = bd1 - bd0, dx = bd3 - bd2;
        last_sum += add_sum;
        last_nom += add_nom;
        // A box that did not grow is a fixed point: every later iteration sees the same strips, grows nothing and adds
        // `zero` to both sums.  Once neither row of this warp grows, only those additions are left (peaked plans -- what
        // the trained network produces -- stop after two or three of the 8 / 15 iterations).
#ifndef PATS_AB_NO_EARLY_EXIT
        if (__all_sync(0xffffffffu, !(mx > a.lb))) {
            for (int rest = it + 1; rest < a.iters; ++rest) last_sum += kZero, last_nom += kZero;
            break;
        }
#endif
    }
    const bool core_exist = (dy > 1) && (dx > 1);

    // ---- border strips of the final box with the previous iteration's extents (utils.py:1245-1253) ------------
    float acc = 0.f;
    {
        const int d = (hl >> 1) & 3, q = hl & 1;
        int off, lim;
        if (d == 0) off = bd2 + bd0 * width, lim = sdx;
        else if (d == 1) off = bd2 + bd1 * width, lim = sdx;
        else if (d == 2) off = bd2 + bd0 * width, lim = sdy;
        else off = bd3 + bd0 * width, lim = sdy;
        const int stepi = (d < 2) ? 1 : width;
        const float *const Aq = (q == 0) ? E : SV;
        if (WIDTH > 0) {
#pragma unroll
            for (int t = 0; t < WIDTH; ++t) {
                int sidx = off + t * stepi;
                if (sidx < 0 || sidx > n - 1 || t > lim) sidx = n + 1;
                acc += Aq[sidx];
            }
        } else {
            const int tmax = min(max(sdx, sdy) + 1, width);
            for (int t = 0; t < tmax; ++t) {
                int sidx = off + t * stepi;
                if (sidx < 0 || sidx > n - 1 || t > lim) sidx = n + 1;
                acc += Aq[sidx];
            }
            for (int t = tmax; t < width; ++t) acc += kZero;
        }
    }
    const float e0 = __shfl_sync(0xffffffffu, acc, hbase + 0), e1 = __shfl_sync(0xffffffffu, acc, hbase + 2);
    const float e2 = __shfl_sync(0xffffffffu, acc, hbase + 4), e3 = __shfl_sync(0xffffffffu, acc, hbase + 6);
    const float s0 = __shfl_sync(0xffffffffu, acc, hbase + 1), s1 = __shfl_sync(0xffffffffu, acc, hbase + 3);
    const float s2 = __shfl_sync(0xffffffffu, acc, hbase + 5), s3 = __shfl_sync(0xffffffffu, acc, hbase + 7);
    const float es4 = ((e0 + e1) + e2) + e3, ss4 = ((s0 + s1) + s2) + s3;

    // ---- weighted mean position / area scale inside the box (utils.py:1254-1268, 1321-1340), f64 sums ----------
    double wx = 0, wy = 0, sxs = 0, sys = 0, wsc = 0, psum = 0, ts = 0;
    int pr = hl / a.grid_w, pc = hl - pr * a.grid_w;
    for (int p = hl; p < n; p += 16, pc += 16) {
        while (pc >= a.grid_w) pc -= a.grid_w, ++pr;
        const bool in = pr >= bd0 && pr <= bd1 && pc >= bd2 && pc <= bd3;
        float ox = kZero, oy = kZero;
        if (in) {
            const float q = sqrtf(E[p] + 1e-7f);
            ox = q / SX[p];
            oy = q / SY[p];
        }
        wx += (double)(ox * (float)pc);
        wy += (double)(oy * (float)pr);
        sxs += (double)ox;
        sys += (double)oy;
        const float o = ox * oy;
        wsc += (double)(o * SV[p]);
        psum += (double)o;
    }
    for (int j = hl; j <= n; j += 16) ts += (double)E[j];
    wx = half_sum_f64(wx), wy = half_sum_f64(wy), sxs = half_sum_f64(sxs), sys = half_sum_f64(sys);
    wsc = half_sum_f64(wsc), psum = half_sum_f64(psum), ts = half_sum_f64(ts);

    if (hl == 0 && row_ok) {
        const size_t r = (size_t)bb * m + i;
        a.avg[2 * r + 1] = (float)wx / (float)sxs + 0.5f;
        a.avg[2 * r + 0] = (float)wy / (float)sys + 0.5f;
        const float avg_scale = sqrtf((float)wsc / (float)psum);
        a.xs[r] = 1.0f / (avg_scale / 1.0f);
        a.ys[r] = 1.0f / (avg_scale * 1.0f);
        int cn[4] = {bd0 * width + bd2, bd0 * width + bd3, bd1 * width + bd2, bd1 * width + bd3};
        float cps = 0.f, css = 0.f;
#pragma unroll
        for (int d = 0; d < 4; ++d) {
            if (cn[d] < 0 || cn[d] > n - 1) cn[d] = n + 1;
            cps += E[cn[d]];
            css += Sval(cn[d]);
        }
        const float the_scale = (float)ts;
        const float core_scale_sum = (the_scale - ss4) + css;
        const float core_sum = (last_sum - es4) + cps;
        a.core[r] = (core_exist && !nomatch) ? fabsf((core_sum - core_scale_sum) / the_scale) : kZero;
        a.whole[r] = nomatch ? kZero : (fabsf(the_scale - last_sum) + last_nom / 4.0f) / the_scale;
        a.bound[4 * r + 0] = bd0, a.bound[4 * r + 1] = bd1, a.bound[4 * r + 2] = bd2, a.bound[4 * r + 3] = bd3;
        if (a.nomatch) a.nomatch[r] = nomatch ? 1 : 0;
    }
}

// est_position's masks (first_layer.py:162-167): row / column argmax of Z equals the dustbin index.
// grid (b): warps take rows for nm1; for nm2 every warp scans a slab of rows for all columns (coalesced) and the
// per-warp (max, first index) pairs are combined in shared memory.
__global__ void __launch_bounds__(256) est_nomatching_kernel(const float *__restrict__ Z, int M, int N, int dust,
                                                             uint8_t *__restrict__ nm1, uint8_t *__restrict__ nm2) {
    extern __shared__ float sm[];  // [8][N] max, [8][N] index
    const int bb = blockIdx.x;
    const float *z = Z + (size_t)bb * M * N;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    for (int i = warp; nm1 != nullptr && i < M - 1; i += nw) {
        float bv = -INFINITY;
        int bi = 0x7fffffff;
        for (int j = lane; j < N; j += 32) {
            const float v = z[(size_t)i * N + j];
            if (bi == 0x7fffffff || v > bv) bv = v, bi = j;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (oi != 0x7fffffff && (bi == 0x7fffffff || ov > bv || (ov == bv && oi < bi))) bv = ov, bi = oi;
        }
        if (lane == 0) nm1[(size_t)bb * (M - 1) + i] = (bi == dust);
    }
    float *pmax = sm;
    int *pidx = reinterpret_cast<int *>(sm + (size_t)nw * N);
    const int rows_per = (M + nw - 1) / nw, r0 = warp * rows_per, r1 = min(M, r0 + rows_per);
    for (int j = lane; j < N - 1; j += 32) {
        float bv = -INFINITY;
        int bi = 0x7fffffff;
        for (int i = r0; i < r1; ++i) {
            const float v = z[(size_t)i * N + j];
            if (bi == 0x7fffffff || v > bv) bv = v, bi = i;
        }
        pmax[warp * N + j] = bv, pidx[warp * N + j] = bi;
    }
    __syncthreads();
    for (int j = threadIdx.x; j < N - 1; j += blockDim.x) {
        float bv = pmax[j];
        int bi = pidx[j];
        for (int w = 1; w < nw; ++w) {
            const float ov = pmax[w * N + j];
            const int oi = pidx[w * N + j];
            if (oi != 0x7fffffff && (bi == 0x7fffffff || ov > bv)) bv = ov, bi = oi;  // slabs are in row order: ties keep the earlier
        }
        nm2[(size_t)bb * (N - 1) + j] = (bi == dust);
    }
}

// =================================================================================================
// a11  merge_patches_new / merge_patches_old
// =================================================================================================
// Ordered index maps between matched level-1 patches q (row-major over [B,hw]) and window numbers p.
__global__ void __launch_bounds__(1024) window_maps_kernel(const uint8_t *__restrict__ nm_L1, int total, int *__restrict__ qmap,
                                                           int *__restrict__ pmap, int capacity, int *count) {
    pdl_prologue();
    __shared__ int warp_tot[32];
    __shared__ int base;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) base = 0;
    __syncthreads();
    for (int start = 0; start < total; start += blockDim.x) {
        const int q = start + threadIdx.x;
        const bool matched = q < total && nm_L1[q] == 0;
        const unsigned mm = __ballot_sync(0xffffffffu, matched);
        if (lane == 0) warp_tot[warp] = __popc(mm);
        __syncthreads();
        int before = base;
        for (int w = 0; w < warp; ++w) before += warp_tot[w];
        const int pos = before + __popc(mm & ((1u << lane) - 1u));
        if (q < total) pmap[q] = matched ? pos : -1;
        if (matched && pos < capacity) qmap[pos] = q;
        __syncthreads();
        if (threadIdx.x == 0) {
            int t = 0;
            for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += warp_tot[w];
            base += t;
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) *count = base;
}

// Pass 1 (per window cell): border-ring weighting and filtering (second_layer.py:190-201 / :138-149), scores into
// scores_back (:210 / :157).
__global__ void merge_rings_kernel(float *__restrict__ trust, uint8_t *__restrict__ nm_L2, double *__restrict__ scores_back,
                                   const int *__restrict__ qmap, int P, int merge_new) {
    pdl_prologue();
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= P * 144) return;
    const int p = e / 144, cell = e - p * 144;
    const int y = cell / 12, x = cell - 12 * y;
    float t = trust[e];
#pragma unroll
    for (int i = 0; i < 3; ++i)
        if (x < 3 - i || x > 7 + i || y < 3 - i || y > 7 + i) t *= 2.0f;
    uint8_t f = nm_L2[e];
    if (t > 2.0f) f = 1;
    if (x < 1 || x > 10 || y < 1 || y > 10) f = 1;
    if (merge_new && !f) t -= 10000.0f;
    trust[e] = t;
    nm_L2[e] = f;
    const int a = y >> 2, r = y & 3, c = x >> 2, s = x & 3;
    scores_back[((size_t)qmap[p] * 16 + r * 4 + s) * 9 + a * 3 + c] = (double)t;
}

// Which of several EQUAL minima `torch.argsort(x)[..., 0]` returns on CUDA tensors (second_layer.py:169 / :230).  argsort is
// unstable; ATen sorts slices of <= 32 elements with a bitonic network over 32 slots / 16 threads in which equal keys ARE
// exchanged (ATen/native/cuda/SortUtils.cuh: bitonicSort, LTOp, invalid slots sort to the end), so among tied minima the winner
// is a fixed function of WHICH slots hold the minimum -- never of the other values.  Ties are the rule here (windows that do not
// exist score exactly 0.0; matched cells are -10000 + trust in f32, i.e. quantised to ~1e-3), and the choice changes which
// window keeps a fine cell.  The table is indexed by the 9-bit mask of the minimum's positions; generated by emulating that
// network (`python tools/argsort_tie_probe.py table`) and checked against the live op on a B200 for 200 000 random tie
// patterns (profiles/r02_argsort_tie_probe.json: agreement 1.0; "first index", which is what the CPU kernel does, 0.55).
__constant__ uint8_t kArgsortTie9[512] = {
    0, 0, 1, 1, 2, 0, 1, 2, 3, 0, 1, 3, 2, 3, 3, 1, 4, 4, 4, 1, 4, 0, 1, 4, 4, 0, 1, 4, 2, 4, 4, 1,
    5, 5, 5, 1, 5, 0, 1, 5, 5, 0, 1, 5, 2, 5, 5, 1, 4, 0, 1, 4, 2, 4, 4, 0, 3, 4, 4, 0, 4, 0, 1, 1,
    6, 6, 6, 1, 6, 0, 1, 6, 6, 0, 1, 6, 2, 6, 6, 1, 6, 0, 1, 6, 2, 6, 6, 0, 3, 6, 6, 0, 6, 0, 1, 1,
    6, 0, 1, 6, 2, 6, 6, 0, 3, 6, 6, 0, 6, 0, 1, 1, 4, 6, 6, 1, 6, 0, 1, 1, 6, 0, 1, 1, 2, 2, 2, 1,
    7, 7, 7, 1, 7, 0, 1, 7, 7, 0, 1, 7, 2, 7, 7, 1, 7, 0, 1, 7, 2, 7, 7, 0, 3, 7, 7, 0, 7, 0, 1, 1,
    7, 0, 1, 7, 2, 7, 7, 0, 3, 7, 7, 0, 7, 0, 1, 1, 4, 7, 7, 1, 7, 0, 1, 1, 7, 0, 1, 1, 2, 2, 2, 1,
    7, 0, 1, 7, 2, 7, 7, 0, 3, 7, 7, 0, 7, 0, 1, 1, 4, 7, 7, 1, 7, 0, 1, 1, 7, 0, 1, 1, 2, 2, 2, 1,
    5, 7, 7, 1, 7, 0, 1, 1, 7, 0, 1, 1, 2, 2, 2, 1, 7, 0, 1, 1, 2, 0, 1, 1, 3, 0, 1, 1, 2, 2, 2, 7,
    8, 0, 1, 1, 2, 0, 1, 1, 3, 0, 1, 1, 2, 2, 2, 1, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 1,
    5, 5, 5, 5, 5, 5, 5, 5, 5, 5, 5, 5, 5, 5, 5, 1, 4, 4, 4, 4, 4, 4, 4, 0, 4, 4, 4, 0, 4, 0, 1, 1,
    6, 6, 6, 6, 6, 6, 6, 6, 6, 6, 6, 6, 6, 6, 6, 1, 6, 6, 6, 6, 6, 6, 6, 0, 6, 6, 6, 0, 6, 0, 1, 1,
    6, 6, 6, 6, 6, 6, 6, 0, 6, 6, 6, 0, 6, 0, 1, 1, 6, 6, 6, 1, 6, 0, 1, 1, 6, 0, 1, 1, 2, 2, 2, 6,
    7, 7, 7, 7, 7, 7, 7, 7, 7, 7, 7, 7, 7, 7, 7, 1, 7, 7, 7, 7, 7, 7, 7, 0, 7, 7, 7, 0, 7, 0, 1, 1,
    7, 7, 7, 7, 7, 7, 7, 0, 7, 7, 7, 0, 7, 0, 1, 1, 7, 7, 7, 1, 7, 0, 1, 1, 7, 0, 1, 1, 2, 2, 2, 7,
    7, 7, 7, 7, 7, 7, 7, 0, 7, 7, 7, 0, 7, 0, 1, 1, 7, 7, 7, 1, 7, 0, 1, 1, 7, 0, 1, 1, 2, 2, 2, 7,
    7, 7, 7, 1, 7, 0, 1, 1, 7, 0, 1, 1, 2, 2, 2, 7, 7, 0, 1, 1, 2, 0, 1, 7, 3, 0, 1, 7, 2, 7, 7, 7,
};
// index of the winning candidate: tie_first != 0 -> first minimum (ATen CPU; what the CPU-recorded fixtures hold)
__device__ __forceinline__ int argsort9_first(const double (&v)[9], int tie_first) {
    double best = v[0];
#pragma unroll
    for (int k = 1; k < 9; ++k) best = v[k] < best ? v[k] : best;
    unsigned mask = 0;
    int first = -1;
#pragma unroll
    for (int k = 8; k >= 0; --k)
        if (v[k] == best) mask |= 1u << k, first = k;
    if (first < 0) return 0;  // every comparison failed (NaN): not produced by the path
    return tie_first ? first : kArgsortTie9[mask];
}

// Pass 2 (per window cell): which of the <= 9 overlapping windows keeps the cell.  Gather formulation of the
// reference's shift / argsort / scatter (see DESIGN.md "merge_regroup" for the derivation).
__global__ void merge_select_kernel(const uint8_t *__restrict__ nm_L2, const double *__restrict__ scores_back,
                                    const int *__restrict__ qmap, const int *__restrict__ pmap, int P, int height, int width,
                                    int merge_new, int tie_first, uint8_t *__restrict__ out) {
    pdl_prologue();
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= P * 144) return;
    const int p = e / 144, cell = e - p * 144;
    const int cy = cell / 12, cx = cell - 12 * cy;
    const int a1 = cy >> 2, r = cy & 3, c1 = cx >> 2, s = cx & 3;
    const int hw = height * width, H4 = 4 * height, W4 = 4 * width;
    const int q = qmap[p], bb = q / hw, ij = q - bb * hw;
    const int i1 = ij / width, j1 = ij - i1 * width;
    const int Y = 4 * (i1 + a1 - 1) + r, X = 4 * (j1 + c1 - 1) + s;  // true fine-grid cell
    uint8_t res = 1;
    if (Y >= 0 && Y < H4 && X >= 0 && X < W4) {
        int ks = 0;
        double cand[9];
        if (merge_new) {
            // second_layer.py:225: argsort of the centre window's OWN nine scores (+1e5 where the neighbour is outside)
            const double *sb = scores_back + ((size_t)(bb * hw + (Y >> 2) * width + (X >> 2)) * 16 + r * 4 + s) * 9;
#pragma unroll
            for (int k = 0; k < 9; ++k) {
                const int yy = Y + 4 * (k / 3 - 1), xx = X + 4 * (k % 3 - 1);
                double v = sb[k];
                if (yy < 0 || yy >= H4 || xx < 0 || xx >= W4) v += 100000.0;
                cand[k] = v;
            }
            ks = argsort9_first(cand, tie_first);
            if (ks == (2 - a1) * 3 + (2 - c1)) res = nm_L2[e];
        } else {
            // second_layer.py:159-169: per-channel shifted scores (own value where the shift leaves the grid),
            // -1e4 where that window matched the cell
#pragma unroll
            for (int k = 0; k < 9; ++k) {
                const int a = k / 3, c = k % 3;
                int yy = Y - 4 * (a - 1), xx = X - 4 * (c - 1);
                if (yy < 0 || yy >= H4 || xx < 0 || xx >= W4) yy = Y, xx = X;
                const int qq = bb * hw + (yy >> 2) * width + (xx >> 2);
                double v = scores_back[((size_t)qq * 16 + (yy & 3) * 4 + (xx & 3)) * 9 + k];
                const int pp = pmap[qq];
                if (pp >= 0 && nm_L2[(size_t)pp * 144 + (4 * a + (yy & 3)) * 12 + 4 * c + (xx & 3)] == 0) v -= 10000.0;
                cand[k] = v;
            }
            ks = argsort9_first(cand, tie_first);
            if (ks == a1 * 3 + c1) res = nm_L2[e];
        }
    }
    out[e] = res;
}

// =================================================================================================
// a14  get_result: two-level ordered compaction + affine composition
// =================================================================================================
// Launched behind window_maps_kernel, which it does not depend on: it is resident only after every window_maps CTA has
// passed its own wait (so everything older is complete), runs concurrently with it, and waits for it at the END so that
// "this grid complete" still implies "everything before it complete" for the kernels chained behind.
__global__ void __launch_bounds__(256) count_rows_kernel(const uint8_t *__restrict__ nm1, int n1, int *__restrict__ cnt) {
    pdl_launch_dependents();
    __shared__ int part[8];
    const int p = blockIdx.x;
    int c = 0;
    const uint8_t *row = nm1 + (size_t)p * n1;
    if ((n1 & 15) == 0 && (reinterpret_cast<uintptr_t>(nm1) & 15) == 0) {
        // 16 flags per load (a 48 x 48 window row is 144 loads: one per thread, all in flight together)
        const uint4 *row4 = reinterpret_cast<const uint4 *>(row);
        for (int e = threadIdx.x; e < (n1 >> 4); e += blockDim.x) {
            const uint4 v = __ldg(row4 + e);
            c += (__popc(__vcmpeq4(v.x, 0u)) + __popc(__vcmpeq4(v.y, 0u)) + __popc(__vcmpeq4(v.z, 0u)) + __popc(__vcmpeq4(v.w, 0u))) >> 3;
        }
    } else {
        for (int e = threadIdx.x; e < n1; e += blockDim.x) c += row[e] == 0;
    }
    c = __reduce_add_sync(0xffffffffu, c);
    if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        int t = 0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += part[w];
        cnt[p] = t;
    }
    pdl_wait();
}

__global__ void __launch_bounds__(1024) exclusive_scan_kernel(const int *__restrict__ cnt, int P, long long *__restrict__ off,
                                                              long long *total) {
    pdl_prologue();
    __shared__ long long warp_tot[32];
    __shared__ long long base;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) base = 0;
    __syncthreads();
    for (int start = 0; start < P; start += blockDim.x) {
        const int p = start + threadIdx.x;
        const long long v = p < P ? cnt[p] : 0;
        long long incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const long long t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) warp_tot[warp] = incl;
        __syncthreads();
        long long before = base;
        for (int w = 0; w < warp; ++w) before += warp_tot[w];
        if (p < P) off[p] = before + incl - v;
        __syncthreads();
        if (threadIdx.x == 0) {
            long long t = 0;
            for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += warp_tot[w];
            base += t;
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) *total = base;
}

struct ResultArgs {
    const uint8_t *nm1;
    const float *pt0, *sc0, *pt1, *sc1;
    const int *qmap;
    const long long *off;   // exclusive scan of cnt (exclusive_scan_kernel), or nullptr: every CTA sums cnt[0..p) itself
    const int *cnt;         // matches per window
    long long *total;       // written by the last CTA when off == nullptr
    int P, n0, w0, ps0, n1, w1, ps1;
    long long capacity;
    float *ml, *mr;
};

__global__ void __launch_bounds__(256) assemble_matches_kernel(ResultArgs a) {
    pdl_prologue();
    __shared__ int warp_tot[8];
    __shared__ long long base;
    const int p = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int q = a.qmap[p], cell = q % a.n0;
    // level 0 (utils.py:205-207)
    float l0[2], r0[2];
    const float pos0[2] = {(float)((cell / a.w0) * a.ps0), (float)((cell % a.w0) * a.ps0)};
#pragma unroll
    for (int d = 0; d < 2; ++d) {
        const float dl = (pos0[d] + 0.5f * (float)a.ps0) - (1.5f * a.sc0[(size_t)q * 2 + 1]) * (float)a.ps0;
        const float dr = (a.pt0[(size_t)q * 2 + d] - 1.5f * a.sc0[(size_t)q * 2 + 0]) * (float)a.ps0;
        l0[d] = 0.f + dl;
        r0[d] = 0.f + dr;
    }
    if (a.off) {
        if (threadIdx.x == 0) base = a.off[p];
    } else {  // few windows: the prefix over the preceding windows' counts costs less than one more kernel in the chain
        long long t = 0;
        for (int i = threadIdx.x; i < p; i += blockDim.x) t += a.cnt[i];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
        __shared__ long long wsum[8];
        if (lane == 0) wsum[warp] = t;
        __syncthreads();
        if (threadIdx.x == 0) {
            long long b0 = 0;
            for (int w = 0; w < (int)(blockDim.x >> 5); ++w) b0 += wsum[w];
            base = b0;
            if (p == a.P - 1) *a.total = b0 + a.cnt[p];
        }
    }
    __syncthreads();
    constexpr int kMaxChunks = 32;  // fast path: windows of up to 32 * 256 cells
    const int nch = (a.n1 + (int)blockDim.x - 1) / (int)blockDim.x;
    const bool aligned8 = ((reinterpret_cast<uintptr_t>(a.sc1) | reinterpret_cast<uintptr_t>(a.pt1) | reinterpret_cast<uintptr_t>(a.ml) |
                            reinterpret_cast<uintptr_t>(a.mr)) & 7) == 0;  // (y, x) pairs move as one 8-byte access
    if (nch <= kMaxChunks && blockDim.x == 256 && aligned8) {
        // Every chunk's ballots first (one pass over the flags, all loads in flight), one scan over the (chunk, warp)
        // counts in output order, then the writes: two CTA barriers per window instead of three per chunk.  Lane-consecutive
        // cells still go to consecutive output rows, so the gathers of pt1 / sc1 and the stores stay coalesced.
        __shared__ unsigned ball[kMaxChunks * 8];
        __shared__ int pre[kMaxChunks * 8];
        for (int c = 0; c < nch; ++c) {
            const int c1 = c * 256 + threadIdx.x;
            const bool matched = c1 < a.n1 && a.nm1[(size_t)p * a.n1 + c1] == 0;
            const unsigned mm = __ballot_sync(0xffffffffu, matched);
            if (lane == 0) ball[c * 8 + warp] = mm;
        }
        __syncthreads();
        if (warp == 0) {
            int carry = 0;
            for (int s0 = 0; s0 < nch * 8; s0 += 32) {
                const int idx = s0 + lane;
                const int v = idx < nch * 8 ? __popc(ball[idx]) : 0;
                int incl = v;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const int t = __shfl_up_sync(0xffffffffu, incl, o);
                    if (lane >= o) incl += t;
                }
                if (idx < nch * 8) pre[idx] = carry + incl - v;
                carry += __shfl_sync(0xffffffffu, incl, 31);
            }
        }
        __syncthreads();
        const long long base0 = base;
        for (int c = 0; c < nch; ++c) {
            const int c1 = c * 256 + threadIdx.x;
            const unsigned mm = ball[c * 8 + warp];
            if (!((mm >> lane) & 1u)) continue;
            const long long pos = base0 + pre[c * 8 + warp] + __popc(mm & ((1u << lane) - 1u));
            if (pos >= a.capacity) continue;
            const size_t e = (size_t)p * a.n1 + c1;
            const float pos1[2] = {(float)((c1 / a.w1) * a.ps1), (float)((c1 % a.w1) * a.ps1)};
            const float2 sc = __ldg(reinterpret_cast<const float2 *>(a.sc1) + e), pt = __ldg(reinterpret_cast<const float2 *>(a.pt1) + e);
            const float ptv[2] = {pt.x, pt.y};
            float lv[2], rv[2];
#pragma unroll
            for (int d = 0; d < 2; ++d) {
                // last level (utils.py:209-210)
                const float dl = (pos1[d] + 0.5f * (float)a.ps1) * sc.y;
                const float dr = (ptv[d] * (float)a.ps1) * sc.x;
                lv[d] = l0[d] + dl;
                rv[d] = r0[d] + dr;
            }
            reinterpret_cast<float2 *>(a.ml)[pos] = make_float2(lv[0], lv[1]);
            reinterpret_cast<float2 *>(a.mr)[pos] = make_float2(rv[0], rv[1]);
        }
        return;
    }
    for (int start = 0; start < a.n1; start += blockDim.x) {
        const int c1 = start + threadIdx.x;
        const size_t e = (size_t)p * a.n1 + c1;
        const bool matched = c1 < a.n1 && a.nm1[e] == 0;
        const unsigned mm = __ballot_sync(0xffffffffu, matched);
        if (lane == 0) warp_tot[warp] = __popc(mm);
        __syncthreads();
        long long pos = base;
        for (int w = 0; w < warp; ++w) pos += warp_tot[w];
        pos += __popc(mm & ((1u << lane) - 1u));
        if (matched && pos < a.capacity) {
            const float pos1[2] = {(float)((c1 / a.w1) * a.ps1), (float)((c1 % a.w1) * a.ps1)};
#pragma unroll
            for (int d = 0; d < 2; ++d) {
                // last level (utils.py:209-210)
                const float dl = (pos1[d] + 0.5f * (float)a.ps1) * a.sc1[e * 2 + 1];
                const float dr = (a.pt1[e * 2 + d] * (float)a.ps1) * a.sc1[e * 2 + 0];
                a.ml[pos * 2 + d] = l0[d] + dl;
                a.mr[pos * 2 + d] = r0[d] + dr;
            }
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            int t = 0;
            for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += warp_tot[w];
            base += t;
        }
        __syncthreads();
    }
}

// =================================================================================================
// a13  ThirdLayer.Compute_result: one thread per (problem, inner source cell), rows staged in smem
// =================================================================================================
#ifndef TH_K_N
#define TH_K_N 8
#endif
constexpr int TH_K = TH_K_N;  // problems per CTA

__global__ void __launch_bounds__(TH_K * 16) third_result_kernel(const float *__restrict__ scores, const float *__restrict__ scale_x,
                                                                 const float *__restrict__ scale_y, const int64_t *__restrict__ p_s,
                                                                 const int64_t *__restrict__ p_t, int K, int log_input,
                                                                 float *__restrict__ mk0, float *__restrict__ mk1,
                                                                 uint8_t *__restrict__ if_matching1, const unsigned *done, unsigned epoch) {
    constexpr int W = 8, T = 5, NN = 65;
    __shared__ float rows[TH_K * 16][NN];
    __shared__ float gx[TH_K][64], gy[TH_K][64];
    const int k0 = blockIdx.x * TH_K;
    if (done) {  // launched early (plan hand-over, sinkhorn_common.cuh): wait for this CTA's problems
        pdl_launch_dependents();
        if (threadIdx.x < TH_K && k0 + threadIdx.x < K) await_problem(done, epoch, k0 + threadIdx.x);
        __syncthreads();
    } else {
        pdl_prologue();
    }
    // 8320 plan entries per CTA = 65 per thread: loaded 13 at a time so that the loads of a batch are in flight together
    // (one load per loop trip left the kernel waiting on L2 latency for half of its time)
    constexpr int PER = TH_K * 16 * NN / (TH_K * 16), BATCH = 13;
    static_assert(PER == NN && NN % BATCH == 0, "65 entries per thread in 5 batches of 13");
    float *rows_flat = &rows[0][0];
#pragma unroll 1
    for (int b0 = 0; b0 < PER; b0 += BATCH) {
        float v[BATCH];
        bool ok[BATCH];
#pragma unroll
        for (int u = 0; u < BATCH; ++u) {
            const int e = (b0 + u) * (TH_K * 16) + threadIdx.x;
            const int rr = e / NN, j = e - rr * NN;
            const int k = k0 + rr / 16, c16 = rr % 16;
            const int srow = (2 + c16 / 4) * W + 2 + c16 % 4;  // inner 4x4 of the 8x8 source window (:186)
            ok[u] = k < K;
            v[u] = ok[u] ? __ldcg(scores + ((size_t)k * NN + srow) * NN + j) : 0.f;  // .cg: the producer grid may still run
        }
#pragma unroll
        for (int u = 0; u < BATCH; ++u)  // third_layer.py:159 scores = exp(scores_origin)
            rows_flat[(b0 + u) * (TH_K * 16) + threadIdx.x] = (log_input && ok[u]) ? expf(v[u]) : v[u];
    }
    for (int e = threadIdx.x; e < TH_K * 64; e += blockDim.x) {
        const int k = k0 + e / 64;
        gx[e / 64][e % 64] = k < K ? __ldg(scale_x + (size_t)k * 64 + e % 64) : 1.f;
        gy[e / 64][e % 64] = k < K ? __ldg(scale_y + (size_t)k * 64 + e % 64) : 1.f;
    }
    __syncthreads();
    const int kk = threadIdx.x / 16, c16 = threadIdx.x % 16, k = k0 + kk;
    if (k >= K) return;
    const float *row = rows[threadIdx.x];
    int max0 = 0;
    for (int j = 1; j < 64; ++j)
        if (row[j] > row[max0]) max0 = j;
    int mall = 0;
    for (int j = 1; j < NN; ++j)
        if (row[j] + 1e-8f > row[mall] + 1e-8f) mall = j;  // label test, third_layer.py:166-167
    if_matching1[(size_t)k * 16 + c16] = (mall != W * W);
    const int mx = max0 % W, my = max0 / W;
    float wx = 0.f, wy = 0.f, sx = 0.f, sy = 0.f;
    for (int ty = 0; ty < T; ++ty)
        for (int tx = 0; tx < T; ++tx) {
            const int iy = my + ty - 2, ix = mx + tx - 2;
            const bool in = iy >= 0 && iy < W && ix >= 0 && ix < W;
            const float sc = in ? row[iy * W + ix] : 0.f;       // ZeroPad2d(2)             (:185)
            const float sgx = in ? gx[kk][iy * W + ix] : 1e-2f;  // ConstantPad2d(2, 1e-2)   (:195-196)
            const float sgy = in ? gy[kk][iy * W + ix] : 1e-2f;
            const float qv = sqrtf(sc + 1e-7f);
            const float ux = qv / sgx, uy = qv / sgy;
            wx += ux * (float)(tx * 2 - (T - 1));
            wy += uy * (float)(ty * 2 - (T - 1));
            sx += ux;
            sy += uy;
        }
    float *o1 = mk1 + ((size_t)k * 16 + c16) * 2, *o0 = mk0 + ((size_t)k * 16 + c16) * 2;
    o1[0] = (wx / sx + ((float)mx + 0.5f - (float)W / 2) * 2.0f) + (float)p_t[(size_t)k * 2 + 0];
    o1[1] = (wy / sy + ((float)my + 0.5f - (float)W / 2) * 2.0f) + (float)p_t[(size_t)k * 2 + 1];
    o0[0] = ((float)p_s[(size_t)k * 2 + 0] + (float)(c16 % 4) * 2.0f) - 3.0f;
    o0[1] = ((float)p_s[(size_t)k * 2 + 1] + (float)(c16 / 4) * 2.0f) - 3.0f;
}

template <int WIDTH>
static int launch_area_expand_w(const ExpandArgs &a, dim3 grid, size_t smem, cudaStream_t st, bool early) {
    if (smem > 48 * 1024)
        PATS_CUDA_TRY(cudaFuncSetAttribute(area_expand_kernel<WIDTH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = dim3(EX_ROWS * 16);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;  // plan hand-over: start behind a still-running producer
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = (early || g_chain) ? 1 : 0;  // hand-over consumer, or plain launch chaining (the kernel then starts with pdl_prologue)
    PATS_CUDA_TRY(cudaLaunchKernelEx(&cfg, area_expand_kernel<WIDTH>, a));
    return PATS_OK;
}
// strip length 20 (level 1: 15 x 20 coarse grid, one problem) is compiled in; anything else takes the loops
static int launch_area_expand(const ExpandArgs &a, dim3 grid, size_t smem, cudaStream_t st, bool early) {
    if (a.width == 20 && a.b * a.m <= 4096) return launch_area_expand_w<20>(a, grid, smem, st, early);  // few rows: latency-bound
    return launch_area_expand_w<0>(a, grid, smem, st, early);  // many rows: issue-bound, the short extent loops win
}

}  // namespace pats

using namespace pats;

PATS_API int pats_iterative_expand_matrix_f32(const float *scores_in, const float *scalex, const float *scaley, int b, int m,
                                              int grid_h, int grid_w, float lower_bound, int iter_num, float *whole_cost,
                                              float *core_cost, float *average_point, float *x_scale, float *y_scale,
                                              int64_t *bound, uint8_t *if_nomatching, void *stream) {
    const long long n = (long long)grid_h * grid_w;
    if (b < 0 || m <= 0 || grid_h <= 0 || grid_w <= 0 || iter_num < 1) return invalid("iterative_expand_matrix: bad sizes");
    if (b == 0) return PATS_OK;
    if (!scores_in || !scalex || !scaley || !whole_cost || !core_cost || !average_point || !x_scale || !y_scale || !bound)
        return invalid("iterative_expand_matrix: null pointer");
    ExpandArgs a;
    a.scores = scores_in, a.sx = scalex, a.sy = scaley;
    a.b = b, a.m = m, a.n = (int)n, a.grid_w = grid_w;
    a.width = grid_h > grid_w ? grid_h : grid_w;  // ranges.shape[0]   (utils.py:1181)
    a.height = (int)(n / a.width);
    a.iters = iter_num, a.lb = lower_bound, a.log_input = 0;
    a.whole = whole_cost, a.core = core_cost, a.avg = average_point, a.xs = x_scale, a.ys = y_scale;
    a.bound = bound, a.nomatch = if_nomatching;
    a.nm2 = nullptr, a.colmax = nullptr, a.counter = nullptr;
    a.done = nullptr, a.epoch = 0u;
    const size_t smem = sizeof(float) * (size_t)(EX_SM + EX_ROWS) * (n + 2);
    if (smem > 200 * 1024) return invalid("iterative_expand_matrix: grid of %lld cells exceeds the shared-memory budget", n);
    if (b > 65535) return invalid("iterative_expand_matrix: batch %d exceeds gridDim.y", b);
    return launch_area_expand(a, dim3((m + EX_ROWS - 1) / EX_ROWS, b), smem, as_stream(stream), false);
}

PATS_API int pats_est_nomatching_f32(const float *Z, int b, int M, int N, int dust, uint8_t *nm1, uint8_t *nm2, void *stream) {
    if (b < 0 || M < 2 || N < 2) return invalid("est_nomatching: bad sizes");
    if (b == 0) return PATS_OK;
    if (!Z || !nm1 || !nm2) return invalid("est_nomatching: null pointer");
    const size_t smem = sizeof(float) * 2 * 8 * (size_t)N;
    if (smem > 200 * 1024) return invalid("est_nomatching: N = %d exceeds the shared-memory budget", N);
    if (smem > 48 * 1024)
        PATS_CUDA_TRY(cudaFuncSetAttribute(est_nomatching_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    est_nomatching_kernel<<<b, 256, smem, as_stream(stream)>>>(Z, M, N, dust, nm1, nm2);
    PATS_LAUNCH_CHECK("est_nomatching_kernel");
    return PATS_OK;
}

PATS_API int pats_merge_patches(int merge_new, float *trust_score, const uint8_t *nm_L1, uint8_t *nm_L2, double *scores_back, int B,
                                int height, int width, int P, uint8_t *out, int *workspace, void *stream) {
    if (B <= 0 || height <= 0 || width <= 0 || P < 0) return invalid("merge_patches: bad sizes");
    if (!nm_L1 || !scores_back || !workspace) return invalid("merge_patches: null pointer");
    if (P > 0 && (!trust_score || !nm_L2 || !out)) return invalid("merge_patches: null pointer");
    cudaStream_t st = as_stream(stream);
    const int total = B * height * width;
    int *qmap = workspace, *pmap = workspace + total, *count = workspace + 2 * total;  // workspace: 2*B*hw+1 ints
    PATS_CUDA_TRY(launch_chained(window_maps_kernel, dim3(1), dim3(1024), 0, st, nm_L1, total, qmap, pmap, P, count));
    PATS_LAUNCH_CHECK("window_maps_kernel");
    if (P == 0) return PATS_OK;
    const int cells = P * 144;
    const int tie_first = (merge_new & PATS_MERGE_TIE_FIRST) ? 1 : 0;
    merge_new &= 1;
    PATS_CUDA_TRY(launch_chained(merge_rings_kernel, dim3((cells + 255) / 256), dim3(256), 0, st, trust_score, nm_L2, scores_back, qmap, P, merge_new ? 1 : 0));
    PATS_LAUNCH_CHECK("merge_rings_kernel");
    PATS_CUDA_TRY(launch_chained(merge_select_kernel, dim3((cells + 255) / 256), dim3(256), 0, st, nm_L2, scores_back, qmap, pmap, P, height, width, merge_new ? 1 : 0, tie_first, out));
    PATS_LAUNCH_CHECK("merge_select_kernel");
    if (!merge_new) PATS_CUDA_TRY(cudaMemsetAsync(scores_back, 0, sizeof(double) * (size_t)total * 144, st));  // second_layer.py:186
    return PATS_OK;
}

__global__ void argsort9_first_kernel(const double *__restrict__ x, int n, int tie_first, int *__restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double v[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) v[k] = x[(size_t)i * 9 + k];
    out[i] = argsort9_first(v, tie_first);
}

PATS_API int pats_argsort9_first_f64(const double *x, int n, int tie_first, int *out, void *stream) {
    if (n < 0) return invalid("argsort9_first: bad size");
    if (n == 0) return PATS_OK;
    if (!x || !out) return invalid("argsort9_first: null pointer");
    argsort9_first_kernel<<<(n + 255) / 256, 256, 0, as_stream(stream)>>>(x, n, tie_first, out);
    PATS_LAUNCH_CHECK("argsort9_first_kernel");
    return PATS_OK;
}

PATS_API int pats_get_result_f32(const uint8_t *nm0, const float *pt0, const float *sc0, int B, int ps0, int h0, int w0,
                                 const uint8_t *nm1, const float *pt1, const float *sc1, int P, int ps1, int h1, int w1,
                                 float *matches_l, float *matches_r, long long capacity, long long *total, void *workspace,
                                 void *stream) {
    if (B <= 0 || h0 <= 0 || w0 <= 0 || h1 <= 0 || w1 <= 0 || P < 0 || capacity < 0) return invalid("get_result: bad sizes");
    if (!nm0 || !pt0 || !sc0 || !total || !workspace) return invalid("get_result: null pointer");
    cudaStream_t st = as_stream(stream);
    const int n0 = h0 * w0, n1 = h1 * w1, tot0 = B * n0;
    // workspace layout: off[P+1] (i64) | qmap[tot0] | pmap[tot0] | count | cnt[P]
    long long *off = (long long *)workspace;
    int *qmap = (int *)(off + P + 1), *pmap = qmap + tot0, *count = pmap + tot0, *cnt = count + 1;
    PATS_CUDA_TRY(launch_chained(window_maps_kernel, dim3(1), dim3(1024), 0, st, nm0, tot0, qmap, pmap, P, count));
    PATS_LAUNCH_CHECK("window_maps_kernel");
    if (P == 0) {
        PATS_CUDA_TRY(cudaMemsetAsync(total, 0, sizeof(long long), st));
        return PATS_OK;
    }
    if (!nm1 || !pt1 || !sc1 || !matches_l || !matches_r) return invalid("get_result: null pointer");
    PATS_CUDA_TRY(launch_chained(count_rows_kernel, dim3(P), dim3(256), 0, st, nm1, n1, cnt));
    PATS_LAUNCH_CHECK("count_rows_kernel");
    const bool own_prefix = P <= 2048;
    if (!own_prefix) {
        PATS_CUDA_TRY(launch_chained(exclusive_scan_kernel, dim3(1), dim3(1024), 0, st, cnt, P, off, total));
        PATS_LAUNCH_CHECK("exclusive_scan_kernel");
    }
    ResultArgs a{nm1, pt0, sc0, pt1, sc1, qmap, own_prefix ? nullptr : off, cnt, total, P, n0, w0, ps0, n1, w1, ps1, capacity, matches_l, matches_r};
    PATS_CUDA_TRY(launch_chained(assemble_matches_kernel, dim3(P), dim3(256), 0, st, a));
    PATS_LAUNCH_CHECK("assemble_matches_kernel");
    return PATS_OK;
}

PATS_API int pats_third_compute_result_f32(const float *scores, const float *scale_x, const float *scale_y, const int64_t *p_s,
                                           const int64_t *p_t, int K, float *mkpts0_f, float *mkpts1_f, uint8_t *if_matching1,
                                           void *stream) {
    if (K < 0) return invalid("third_compute_result: bad sizes");
    if (K == 0) return PATS_OK;
    if (!scores || !scale_x || !scale_y || !p_s || !p_t || !mkpts0_f || !mkpts1_f || !if_matching1)
        return invalid("third_compute_result: null pointer");
    third_result_kernel<<<(K + TH_K - 1) / TH_K, TH_K * 16, 0, as_stream(stream)>>>(scores, scale_x, scale_y, p_s, p_t, K, 0, mkpts0_f,
                                                                                  mkpts1_f, if_matching1, nullptr, 0u);
    PATS_LAUNCH_CHECK("third_result_kernel");
    return PATS_OK;
}

// third-layer result from the log-domain plans; `done` != nullptr: launched early behind the Sinkhorn kernel that is still
// producing them (plan hand-over, sinkhorn_common.cuh)
static int third_result_from_log_launch(const float *Z, const float *scale_x, const float *scale_y, const int64_t *p_s, const int64_t *p_t,
                                        int K, float *mkpts0_f, float *mkpts1_f, uint8_t *if_matching1, cudaStream_t st,
                                        const unsigned *done, unsigned epoch) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((K + TH_K - 1) / TH_K);
    cfg.blockDim = dim3(TH_K * 16);
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = (done || g_chain) ? 1 : 0;
    PATS_CUDA_TRY(cudaLaunchKernelEx(&cfg, third_result_kernel, Z, scale_x, scale_y, p_s, p_t, K, 1, mkpts0_f, mkpts1_f, if_matching1, done,
                                     epoch));
    return PATS_OK;
}

PATS_API int pats_third_result_from_log_f32(const float *Z, const float *scale_x, const float *scale_y, const int64_t *p_s,
                                            const int64_t *p_t, int K, float *mkpts0_f, float *mkpts1_f, uint8_t *if_matching1,
                                            void *stream) {
    if (K < 0) return invalid("third_result_from_log: bad sizes");
    if (K == 0) return PATS_OK;
    if (!Z || !scale_x || !scale_y || !p_s || !p_t || !mkpts0_f || !mkpts1_f || !if_matching1)
        return invalid("third_result_from_log: null pointer");
    return third_result_from_log_launch(Z, scale_x, scale_y, p_s, p_t, K, mkpts0_f, mkpts1_f, if_matching1, as_stream(stream), nullptr, 0u);
}

// third_layer.py:158-167 in one call: log_optimal_transport2 (65 x 65, 100 it) -> exp -> Compute_result + label test.
// The result kernel starts on finished problems while the solve's last wave is still running.
PATS_API int pats_third_layer_match_f32(const float *scores, const float *one, const float *ns, const float *scale_x, const float *scale_y,
                                        const int64_t *p_s, const int64_t *p_t, int K, int iters, float *Z_out, float *mkpts0_f,
                                        float *mkpts1_f, uint8_t *if_matching1, void *stream) {
    if (K < 0 || iters < 0) return invalid("third_layer_match: bad sizes");
    if (K == 0) return PATS_OK;
    if (!scores || !one || !ns || !scale_x || !scale_y || !p_s || !p_t || !Z_out || !mkpts0_f || !mkpts1_f || !if_matching1)
        return invalid("third_layer_match: null pointer");
    cudaStream_t st = as_stream(stream);
    const unsigned *done = nullptr;
    unsigned epoch = 0u;
    const int rc = sinkhorn_ot2_publish(scores, one, ns, K, 65, 65, iters, 0.f, Z_out, st, &done, &epoch);
    if (rc) return rc;
    return third_result_from_log_launch(Z_out, scale_x, scale_y, p_s, p_t, K, mkpts0_f, mkpts1_f, if_matching1, st, done, epoch);
}

static int est_position_launch(const float *Z, const float *scalex, const float *scaley, int b, int grid_h, int grid_w, float lower_bound,
                               int iter_num, float *trust_score, float *average_point, float *x_scale, float *y_scale,
                               uint8_t *if_nomatching1, uint8_t *if_nomatching2, float *core_cost, int64_t *bound, cudaStream_t st,
                               const unsigned *done, unsigned epoch) {
    const long long n = (long long)grid_h * grid_w;
    if (b < 0 || grid_h <= 0 || grid_w <= 0 || iter_num < 1) return invalid("est_position: bad sizes");
    if (b == 0) return PATS_OK;
    if (!Z || !scalex || !scaley || !trust_score || !average_point || !x_scale || !y_scale || !if_nomatching1 || !if_nomatching2 ||
        !core_cost || !bound)
        return invalid("est_position: null pointer");
    // square plan [b, n+1, n+1]: row argmax == n is exactly the mask the expansion computes (utils.py:1194 with m == n)
    ExpandArgs a;
    a.scores = Z, a.sx = scalex, a.sy = scaley;
    a.b = b, a.m = (int)n, a.n = (int)n, a.grid_w = grid_w;
    a.width = grid_h > grid_w ? grid_h : grid_w;
    a.height = (int)(n / a.width);
    a.iters = iter_num, a.lb = lower_bound, a.log_input = 1;
    a.whole = trust_score, a.core = core_cost, a.avg = average_point, a.xs = x_scale, a.ys = y_scale;
    a.bound = bound, a.nomatch = if_nomatching1;
    const size_t smem = sizeof(float) * (size_t)(EX_SM + EX_ROWS) * (n + 2);
    if (smem > 200 * 1024) return invalid("est_position: grid of %lld cells exceeds the shared-memory budget", n);
    if (b > 65535) return invalid("est_position: batch %d exceeds gridDim.y", b);
    // column-argmax mask fused into the expansion kernel: encoded column maxima + per-problem arrival counters, in a
    // scratch buffer that is zero between calls (zeroed when allocated, re-zeroed by the kernel's last CTA per problem;
    // a memset here would sit between the Sinkhorn kernel and this one and serialise the early launch)
    const size_t ws_bytes = sizeof(unsigned) * ((size_t)b * n + b);
    unsigned *ws = static_cast<unsigned *>(zeroed_workspace(st, ws_bytes));
    if (!ws) return cuda_fail(cudaGetLastError(), "est_position workspace");
    a.nm2 = if_nomatching2, a.colmax = ws, a.counter = ws + (size_t)b * n;
    a.done = done, a.epoch = epoch;
    return launch_area_expand(a, dim3(((int)n + EX_ROWS - 1) / EX_ROWS, b), smem, st, done != nullptr);
}

PATS_API int pats_est_position_f32(const float *Z, const float *scalex, const float *scaley, int b, int grid_h, int grid_w,
                                   float lower_bound, int iter_num, float *trust_score, float *average_point, float *x_scale,
                                   float *y_scale, uint8_t *if_nomatching1, uint8_t *if_nomatching2, float *core_cost,
                                   int64_t *bound, void *stream) {
    return est_position_launch(Z, scalex, scaley, b, grid_h, grid_w, lower_bound, iter_num, trust_score, average_point, x_scale, y_scale,
                               if_nomatching1, if_nomatching2, core_cost, bound, as_stream(stream), nullptr, 0u);
}

// second_layer.py:103-116 in one call: log_optimal_transport2 on the [b, n+1, n+1] scores -> dustbin column / row +=
// edge_add (log 2 outdoor, log 3 indoor; :108-112) -> est_position.  Z_out is the plan the reference returns as 'scores'.
// The expansion starts on finished problems while the solve's last wave is still running (plan hand-over).
PATS_API int pats_second_layer_match_f32(const float *scores, const float *one, const float *ns, const float *scalex, const float *scaley,
                                         int b, int grid_h, int grid_w, int iters, float edge_add, float lower_bound, int iter_num,
                                         float *Z_out, float *trust_score, float *average_point, float *x_scale, float *y_scale,
                                         uint8_t *if_nomatching1, uint8_t *if_nomatching2, float *core_cost, int64_t *bound, void *stream) {
    const long long n = (long long)grid_h * grid_w;
    if (b < 0 || grid_h <= 0 || grid_w <= 0 || iters < 0 || n > 4096) return invalid("second_layer_match: bad sizes");
    if (b == 0) return PATS_OK;
    if (!scores || !one || !ns || !Z_out) return invalid("second_layer_match: null pointer");
    cudaStream_t st = as_stream(stream);
    const unsigned *done = nullptr;
    unsigned epoch = 0u;
    const int rc = sinkhorn_ot2_publish(scores, one, ns, b, (int)n + 1, (int)n + 1, iters, edge_add, Z_out, st, &done, &epoch);
    if (rc) return rc;
    return est_position_launch(Z_out, scalex, scaley, b, grid_h, grid_w, lower_bound, iter_num, trust_score, average_point, x_scale, y_scale,
                               if_nomatching1, if_nomatching2, core_cost, bound, st, done, epoch);
}
