// N3 -- the attention network in front of every matching level, on the tcgen05 tensor cores.
// Replaces, from zju3dv/pats (models/modules.py):
//   AttentionalGNN.forward            :119-134   L layers, 'self' / 'cross', residual update of both descriptor sets
//   AttentionalPropagation.forward    :108-117   message = attn(x, source, source);  mlp(cat([x, message]))
//   MultiHeadedAttention.forward      :100-106   three 1x1 convolutions, view(b, dim, heads, n), attention, merge convolution
//   attention                         :84-88     softmax(q^T k / sqrt(dim)) v per head
//   MLP([2D, 2D, D])                  :58-69     Conv1d(2D,2D) -> BatchNorm1d (inference or batch statistics) -> ReLU -> Conv1d(2D,D)
// called at first_layer.py:106 (D = 448, n = 300, 18 layers), second_layer.py:93 (D = 264, n = 145, 18 layers, b = P windows) and
// third_layer.py:148 (D = 128, n = 65, 10 layers, b = K points).  The reference runs ~25 ATen kernels per layer and side (17 000
// launches per image pair); with the Sinkhorn path at ~1 ms these networks are 2/3 of what is left of a forward pass
// (tools/profile_forward.py).
//
// Formulation.  Activations are kept TOKEN-major inside the network: X[t][c], t = (side * b + problem) * n + token, so that every
// 1x1 convolution is one GEMM over all tokens of all problems with both operands K-major (rows of D / 2D floats, 16-byte aligned):
//     QKV = X  Wqkv^T + bqkv                       [T, 3D]   (rows of Wq / Wk / Wv permuted head-major: c' = h * dim + d <- c = d * heads + h)
//     O   = softmax(Q_h K_h^T / sqrt(dim)) V_h     [T, D]    per (problem, side, head); K / V of the same side ('self') or the other ('cross')
//     Y   = relu([X | O] W1f^T + b1f)              [T, 2D]   W1f = bn_scale * [W1[:, :D] | W1[:, D:] Wm]: the merge convolution and the
//                                                            inference BatchNorm are affine and are folded into the first MLP layer when
//                                                            the weights are packed (products accumulated in FP64)
//     X  += Y W2^T + b2                            [T, D]
// Four kernels per layer, launch-chained (griddepcontrol), no host synchronisation; the [b, D, n] <-> [T, D] transpositions are
// one pass each at entry and exit.  Problems are processed in chunks of as many as fit the caller's workspace (28 n D floats each);
// larger chunks measured faster (fewer kernel tails) than chunks that would stay inside the 126 MB L2.
//
// GEMM kernels: gnn_gemm_tma_kernel (default; operands pre-split into TF32 halves by their producers, tiles by TMA, warp-specialised --
// see its header below) and gnn_gemm_kernel, the first generation (one CTA of 256 threads per (128-token block, <= 256-output block),
// two CTAs per SM: per 32-wide K chunk every thread loads its 16-byte pieces of the FP32 activation and weight tiles -- issued BEFORE it
// waits for the previous chunk's MMAs -- rounds them to TF32, forms the remainder tile and stores both in the canonical K-major
// no-swizzle UMMA layout, lane = (k4 % 4) * 8 + row % 8: conflict-free 128-bit stores; one thread issues tcgen05.mma.kind::tf32, M = 128,
// N <= 256, K = 8, accumulator in TMEM).  Both issue hi*hi + hi*lo + lo*hi per K step ("3xTF32": FP32-class accuracy; the reference's
// own convolutions run as single TF32 through cuDNN, which `pats_gnn_precision(1)` mirrors) in the same order and agree bit for bit.
// Attention kernels: FP32 on the CUDA cores, as the reference computes it (torch.einsum -> SGEMM): one CTA per (problem, side, head),
// K^T / V in shared memory, a warp owns 8 query rows x all keys in registers, softmax by warp shuffles, P through a per-warp shared
// buffer into the P V product.  gnn_attention2_kernel (default: row pairs on the packed FP32 pipe), gnn_attention_kernel (first
// generation, bit-identical), gnn_attention_flash_kernel (any token count: keys in chunks, online softmax).
// train() mode (batch-statistics BatchNorm, pats_attentional_gnn_train_f32): the first MLP GEMM writes FP32, gnn_bn_stats_kernel and
// gnn_bn_apply_kernel normalise per side and update the running statistics.
#include <stdlib.h>

#include <cuda.h>  // CUtensorMap and its enums only: cuTensorMapEncodeTiled is looked up at run time (no link against libcuda)

#include "common.cuh"
#include "tcgen05.cuh"

namespace pats {
namespace {

using namespace tc;

constexpr int GKC = 32;             // K extent staged per chunk
constexpr int GKC4 = GKC / 4;       // core matrices along K per chunk
constexpr int GEMM_THREADS = 256;   // 8 warps stage and drain; warps w and w + 4 share a TMEM lane quadrant
constexpr int GEMM_M = 128;

std::atomic<int> g_precision{3};    // 3 = 3xTF32 (default), 1 = single TF32 (what cuDNN gives the reference's Conv1d)
std::atomic<int> g_gemm_variant{0}; // 0 = TMA-fed warp-specialised GEMM in CTA pairs, 1 = register-staged GEMM, 2 = TMA-fed, single CTAs (pats_gnn_gemm_variant)
std::atomic<int> g_att_variant{0};  // A/B of the level-2 attention tiling (pats_gnn_attention_variant)

struct GemmArgs {
    const float *A1, *A2;  // [T, K1] (row stride lda1), [T, K2] (row stride lda2; K2 = 0: absent): the K extents are concatenated
    const float *W;        // [Nout, K1 + K2] row-major (row stride ldw)
    const float *bias;     // [Nout]
    float *out;            // [T, Nout] at row stride ldo
    int lda1, lda2, K1, K2, ldw, ldo;
    int T, Nout;
    int nb, nblocks, mblocks;  // outputs per block (a multiple of 16, <= 256), blocks along Nout, 128-token blocks
    int relu, accumulate;      // out = relu(acc + bias)   /   out += acc + bias
    int tmem_cols;
};

__device__ __forceinline__ float4 ldg4(const float *p) { return __ldg(reinterpret_cast<const float4 *>(p)); }

// round to TF32 (nearest, ties away from zero in magnitude) with two integer instructions; cvt.rna.tf32.f32 adds a NaN / Inf test
// (FSETP + SEL per value) that the activations of this network never need: an Inf / NaN input stays Inf / NaN through the mask
__device__ __forceinline__ float tf32_round(float x) { return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xffffe000u); }

template <bool SPLIT>
__device__ __forceinline__ void store_split(float *hi, float *lo, int idx, float4 x) {
    const float4 h = make_float4(tf32_round(x.x), tf32_round(x.y), tf32_round(x.z), tf32_round(x.w));
    *reinterpret_cast<float4 *>(hi + idx) = h;
    if (SPLIT) *reinterpret_cast<float4 *>(lo + idx) = make_float4(tf32_round(x.x - h.x), tf32_round(x.y - h.y), tf32_round(x.z - h.z), tf32_round(x.w - h.w));
}

template <bool SPLIT>
__global__ void __launch_bounds__(GEMM_THREADS, 2) gnn_gemm_kernel(GemmArgs a) {
    extern __shared__ __align__(128) float smem[];
    __shared__ __align__(8) unsigned long long s_bar;
    __shared__ unsigned s_tmem;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int kq = lane >> 3, rr = lane & 7;  // lane = (k4 % 4) * 8 + row % 8
    const int nbpad_max = (a.nb + 15) & ~15;
    float *a_hi = smem;
    float *a_lo = a_hi + (SPLIT ? GEMM_M * GKC : 0);
    float *b_hi = a_lo + GEMM_M * GKC;
    float *b_lo = b_hi + (SPLIT ? nbpad_max * GKC : 0);
    const unsigned mb = smem_addr(&s_bar);
    pdl_prologue();
    if (warp == 0) tmem_alloc(&s_tmem, (unsigned)a.tmem_cols);
    if (tid == 0) {
        mbar_init1(mb);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const unsigned tmem = s_tmem;
    unsigned phase = 0;
    const int nch1 = (a.K1 + GKC - 1) / GKC, nch = nch1 + (a.K2 + GKC - 1) / GKC;

    for (int unit = blockIdx.x; unit < a.mblocks * a.nblocks; unit += gridDim.x) {
        const int mblk = unit / a.nblocks, nblk = unit - mblk * a.nblocks;
        const int r0 = mblk * GEMM_M, n0 = nblk * a.nb;
        const int ncols = min(a.nb, a.Nout - n0), npad = (ncols + 15) & ~15;
        const int bgroups = npad >> 3;
        bool inflight = false;
        for (int ch = 0; ch < nch; ++ch) {
            const bool second = ch >= nch1;
            const float *src = second ? a.A2 : a.A1;
            const int lda = second ? a.lda2 : a.lda1;
            const int k0 = (second ? ch - nch1 : ch) * GKC;
            const int kn = min(GKC, (second ? a.K2 : a.K1) - k0);
            const int wc0 = (second ? a.K1 : 0) + k0;
            const int kb_sh = kn > 16 ? 1 : 0;  // warp items along K: one (k4 0..3) or two (k4 0..7)
            // ---- global -> registers (nothing in shared memory is touched yet: the previous chunk's MMAs may still be reading it) ----
            float4 av[4], bv[8];
            const int a_items = 16 << kb_sh, b_items = bgroups << kb_sh;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int w = warp + i * 8;
                const int g8 = w >> kb_sh, k4 = ((w & kb_sh) << 2) + kq;
                const int row = r0 + g8 * 8 + rr;
                av[i] = (w < a_items && row < a.T && k4 * 4 < kn) ? ldg4(src + (size_t)row * lda + k0 + k4 * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int w = warp + i * 8;
                const int g8 = w >> kb_sh, k4 = ((w & kb_sh) << 2) + kq;
                const int nrow = n0 + g8 * 8 + rr;
                bv[i] = (w < b_items && nrow < a.Nout && k4 * 4 < kn) ? ldg4(a.W + (size_t)nrow * a.ldw + wc0 + k4 * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
            if (inflight) {  // the previous chunk's MMAs have finished reading the operand tiles
                mbar_wait_parity(mb, phase);
                phase ^= 1u;
                fence_after_sync();
            }
            // ---- registers -> split -> shared memory (canonical K-major layout) ----
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int w = warp + i * 8;
                const int g8 = w >> kb_sh, k4 = ((w & kb_sh) << 2) + kq;
                if (w < a_items) store_split<SPLIT>(a_hi, a_lo, (g8 * GKC4 + k4) * 32 + rr * 4, av[i]);
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int w = warp + i * 8;
                const int g8 = w >> kb_sh, k4 = ((w & kb_sh) << 2) + kq;
                if (w < b_items) store_split<SPLIT>(b_hi, b_lo, (g8 * GKC4 + k4) * 32 + rr * 4, bv[i]);
            }
            fence_async_smem();
            __syncthreads();
            if (tid == 0) {
                fence_after_sync();
                const unsigned idesc = umma_idesc(npad);
                const int steps = (kn + 7) >> 3;
                for (int s = 0; s < steps; ++s) {
                    const unsigned koff = (unsigned)s * 256u;  // two core matrices (K = 8) per step
                    const unsigned long long ah = umma_desc(smem_addr(a_hi) + koff, GKC4 * 128u), bh = umma_desc(smem_addr(b_hi) + koff, GKC4 * 128u);
                    umma_tf32(tmem, ah, bh, idesc, (ch == 0 && s == 0) ? 0u : 1u);
                    if (SPLIT) {
                        const unsigned long long al = umma_desc(smem_addr(a_lo) + koff, GKC4 * 128u), bl = umma_desc(smem_addr(b_lo) + koff, GKC4 * 128u);
                        umma_tf32(tmem, ah, bl, idesc, 1u);
                        umma_tf32(tmem, al, bh, idesc, 1u);
                    }
                }
                umma_commit(mb);
            }
            inflight = true;
        }
        // ---- epilogue: warp w owns TMEM lanes (= tokens) 32 (w % 4) .. + 31; warps w and w + 4 alternate over the 16-column groups.
        // The residual / bias values of a group are loaded BEFORE the wait for the accumulator (first group) or while the previous
        // group is being finished, so their latency does not add to every group.
        const int quad = warp & 3;
        const int row = r0 + quad * 32 + lane;
        const bool live = row < a.T;
        float *orow = a.out + (size_t)(live ? row : 0) * a.ldo + n0;
        float4 res[4], bia[4];
        auto prefetch = [&](int c0) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const bool ok = c0 + 4 * j < ncols;
                bia[j] = ok ? ldg4(a.bias + n0 + c0 + 4 * j) : make_float4(0.f, 0.f, 0.f, 0.f);
                res[j] = (ok && live && a.accumulate) ? *reinterpret_cast<const float4 *>(orow + c0 + 4 * j) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
        };
        int c0 = (warp >> 2) * 16;
        if (c0 < npad) prefetch(c0);
        mbar_wait_parity(mb, phase);
        phase ^= 1u;
        fence_after_sync();
        for (; c0 < npad; c0 += 32) {
            unsigned v[16];
            tmem_ld16(tmem + ((unsigned)(quad * 32) << 16) + (unsigned)c0, v);
            float4 r[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                r[j] = make_float4(__uint_as_float(v[4 * j]) + bia[j].x + res[j].x, __uint_as_float(v[4 * j + 1]) + bia[j].y + res[j].y,
                                   __uint_as_float(v[4 * j + 2]) + bia[j].z + res[j].z, __uint_as_float(v[4 * j + 3]) + bia[j].w + res[j].w);
                if (a.relu) r[j].x = fmaxf(r[j].x, 0.f), r[j].y = fmaxf(r[j].y, 0.f), r[j].z = fmaxf(r[j].z, 0.f), r[j].w = fmaxf(r[j].w, 0.f);
            }
            const int cur = c0;
            if (c0 + 32 < npad) prefetch(c0 + 32);
            if (live) {
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (cur + 4 * j < ncols) *reinterpret_cast<float4 *>(orow + cur + 4 * j) = r[j];
            }
        }
        fence_before_sync();
        __syncthreads();  // every warp has drained its accumulator rows before the next unit's first MMA overwrites them
        fence_after_sync();
    }
    if (warp == 0) tmem_dealloc(tmem, (unsigned)a.tmem_cols);
}

// ---- second GEMM generation: TMA-fed, warp-specialised --------------------------------------------------------------------------
// The register pass of gnn_gemm_kernel (LDG -> round -> STS) is what bounds it: 12 LDG.128 + 24 STS.128 + ~100 ALU instructions per
// thread and K chunk, all CTAs pulling their tiles through L2 at once.  Here the operands arrive ALREADY split: every producer of an
// activation (transposition, attention, the epilogues below) writes its TF32 halves, the weights are packed as halves, and the
// tiles go global -> shared memory by TMA (cp.async.bulk.tensor, 128-byte swizzle = the K-major UMMA layout) with no thread touching
// them.  One persistent CTA per SM, six warps:
//     warp 0   producer: waits for a free stage, arms its mbarrier with the byte count, issues the 2 (single-pass) or 4 tile loads
//     warp 1   MMA: waits for a full stage, issues the tcgen05.mma of its K steps, tcgen05.commit -> "stage free"; after a unit's last
//              chunk a second commit -> "accumulator full".  Two accumulators of 256 TMEM columns alternate between units.
//     warps 2-5 epilogue: wait for "accumulator full", tcgen05.ld their lane quadrant, bias / ReLU / residual, store FP32 and / or
//              the TF32 halves the next GEMM reads, arrive on "accumulator free" -- while the MMA warp is already in the next unit.
constexpr int TMA_THREADS = 192;
constexpr int ACC_COLS = 256;

struct TmaGemmArgs {
    const float *bias;
    float *out;            // FP32 result (QKV; the residual stream X, read and written in place) or nullptr
    float *out_h, *out_l;  // TF32 halves of the result or nullptr
    int ldo, T, Nout, K1, K2;
    int nb, nblocks, mblocks, layer, stages, split, relu, accumulate;
    int cluster;  // 1, or 2: CTA pairs (a thread-block cluster) work on two token blocks of the same output block and each loads HALF of the
                  // weight tile, multicast into both CTAs' shared memory -- the weight bytes pulled through L2 per CTA halve
};

__device__ __forceinline__ void mbar_init(unsigned mb, unsigned count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(mb), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(unsigned mb, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mb), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned mb) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(mb) : "memory"); }
__device__ __forceinline__ void tma_load_2d(unsigned dst, const CUtensorMap *map, int c0, int c1, unsigned mb) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst), "l"(map), "r"(mb),
                 "r"(c0), "r"(c1)
                 : "memory");
}
__device__ __forceinline__ void tma_load_3d(unsigned dst, const CUtensorMap *map, int c0, int c1, int c2, unsigned mb) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(dst), "l"(map),
                 "r"(mb), "r"(c0), "r"(c1), "r"(c2)
                 : "memory");
}
__device__ __forceinline__ void tma_load_3d_multicast(unsigned dst, const CUtensorMap *map, int c0, int c1, int c2, unsigned mb, unsigned short mask) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4, %5}], [%2], %6;" ::"r"(dst),
        "l"(map), "r"(mb), "r"(c0), "r"(c1), "r"(c2), "h"(mask)
        : "memory");
}
// arrives (once the MMAs issued so far have completed) on the mbarrier at this shared-memory offset in EVERY CTA of the mask
__device__ __forceinline__ void umma_commit_multicast(unsigned mb, unsigned short mask) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(mb), "h"(mask) : "memory");
}
__device__ __forceinline__ unsigned cluster_ctarank() {
    unsigned r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// K-major tile with 128-byte rows, SWIZZLE_128B (8-row atoms of 1024 bytes): start address, LBO unused (1), SBO = 1024 B, layout type 2
__device__ __forceinline__ unsigned long long umma_desc_sw128(unsigned saddr) {
    unsigned long long d = 0;
    d |= (unsigned long long)((saddr & 0x3FFFFu) >> 4);
    d |= 1ull << 16;
    d |= (unsigned long long)(1024u >> 4) << 32;
    d |= 1ull << 46;
    d |= 2ull << 61;
    return d;
}

__global__ void __launch_bounds__(TMA_THREADS, 1)
gnn_gemm_tma_kernel(const __grid_constant__ CUtensorMap a1h, const __grid_constant__ CUtensorMap a1l, const __grid_constant__ CUtensorMap a2h,
                    const __grid_constant__ CUtensorMap a2l, const __grid_constant__ CUtensorMap wh, const __grid_constant__ CUtensorMap wl, TmaGemmArgs a) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    __shared__ __align__(8) unsigned long long s_full[6], s_empty[6], s_accf[2], s_acce[2];
    __shared__ unsigned s_tmem;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    // stage layout: A hi | A lo | B hi | B lo (the lo tiles only when split), every tile 1024-byte aligned
    const unsigned sbase = (smem_addr(smem_raw) + 1023u) & ~1023u;
    const int nbpad = (a.nb + 15) & ~15;
    const unsigned a_bytes = GEMM_M * 128u, b_bytes = (unsigned)((nbpad + 7) & ~7) * 128u;
    const unsigned stage_bytes = (a_bytes + b_bytes) * (a.split ? 2u : 1u);
    pdl_prologue();
    if (warp == 1) tmem_alloc(&s_tmem, 2 * ACC_COLS);
    if (tid == 0) {
        for (int i = 0; i < a.stages; ++i) mbar_init(smem_addr(&s_full[i]), 1), mbar_init(smem_addr(&s_empty[i]), (unsigned)a.cluster);
        for (int i = 0; i < 2; ++i) mbar_init(smem_addr(&s_accf[i]), 1), mbar_init(smem_addr(&s_acce[i]), 4);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const unsigned tmem = s_tmem;
    const int nch1 = (a.K1 + GKC - 1) / GKC, nch = nch1 + (a.K2 + GKC - 1) / GKC;
    // cluster == 2: the two CTAs of a pair walk the same sequence of (token-block pair, output block) units in lock step (the stage
    // barriers tie them together); CTA `crank` takes token block 2 * pair + crank (beyond the last block: all rows out of range -> zero
    // tiles, nothing stored).  unit_of() maps a position of that sequence to this CTA's (token block, output block).
    const int crank = a.cluster == 2 ? (int)cluster_ctarank() : 0;
    const unsigned short cmask = (unsigned short)((1u << a.cluster) - 1u);
    if (a.cluster == 2) cluster_sync_all();  // the peer's barriers exist before anything of ours can reach them
    const int mgroups = (a.mblocks + a.cluster - 1) / a.cluster;
    const int units = mgroups * a.nblocks;
    const int first_unit = (int)blockIdx.x / a.cluster, unit_step = (int)gridDim.x / a.cluster;
    const unsigned b_half_bytes = b_bytes / (unsigned)a.cluster;

    if (warp == 0) {
        if (lane == 0) {
            int q = 0;
            for (int unit = first_unit; unit < units; unit += unit_step) {
                const int mgrp = unit / a.nblocks, nblk = unit - mgrp * a.nblocks;
                const int r0 = (mgrp * a.cluster + crank) * GEMM_M, n0 = nblk * a.nb;
                for (int ch = 0; ch < nch; ++ch, ++q) {
                    const int s = q % a.stages;
                    // free in BOTH CTAs of a pair: our half of the weight tile lands in the peer's stage as well
                    if (q >= a.stages) mbar_wait_parity(smem_addr(&s_empty[s]), (unsigned)(((q / a.stages) - 1) & 1));
                    const bool second = ch >= nch1;
                    const int k0 = (second ? ch - nch1 : ch) * GKC, wk = (second ? a.K1 : 0) + k0;
                    const unsigned mb = smem_addr(&s_full[s]);
                    const unsigned st = sbase + (unsigned)s * stage_bytes;
                    const unsigned bh = st + (a.split ? 2u : 1u) * a_bytes, bl = bh + b_bytes;
                    mbar_expect_tx(mb, stage_bytes);  // own activation tiles + the whole weight tile (our half and the peer's)
                    tma_load_2d(st, second ? &a2h : &a1h, k0, r0, mb);
                    if (a.split) tma_load_2d(st + a_bytes, second ? &a2l : &a1l, k0, r0, mb);
                    if (a.cluster == 2) {
                        const int nh = (int)(b_half_bytes / 128u);  // rows of half a weight tile
                        tma_load_3d_multicast(bh + (unsigned)crank * b_half_bytes, &wh, wk, n0 + crank * nh, a.layer, mb, cmask);
                        if (a.split) tma_load_3d_multicast(bl + (unsigned)crank * b_half_bytes, &wl, wk, n0 + crank * nh, a.layer, mb, cmask);
                    } else {
                        tma_load_3d(bh, &wh, wk, n0, a.layer, mb);
                        if (a.split) tma_load_3d(bl, &wl, wk, n0, a.layer, mb);
                    }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            int q = 0, u = 0;
            for (int unit = first_unit; unit < units; unit += unit_step, ++u) {
                const int nblk = unit % a.nblocks;
                const int ncols = min(a.nb, a.Nout - nblk * a.nb), npad = (ncols + 15) & ~15;
                const unsigned idesc = umma_idesc(npad);
                const int t = u & 1;
                if (u >= 2) mbar_wait_parity(smem_addr(&s_acce[t]), (unsigned)(((u >> 1) - 1) & 1));  // the epilogue has drained this accumulator
                fence_after_sync();
                const unsigned acc = tmem + (unsigned)(t * ACC_COLS);
                for (int ch = 0; ch < nch; ++ch, ++q) {
                    const int s = q % a.stages;
                    mbar_wait_parity(smem_addr(&s_full[s]), (unsigned)((q / a.stages) & 1));
                    fence_after_sync();
                    const bool second = ch >= nch1;
                    const int kn = min(GKC, (second ? a.K2 : a.K1) - (second ? ch - nch1 : ch) * GKC);
                    const unsigned st = sbase + (unsigned)s * stage_bytes;
                    // the four tile descriptors once per stage; a K step (8 TF32 values = 32 bytes inside the 128-byte swizzle row) advances
                    // the start-address field (16-byte units) by 2
                    const unsigned long long ah0 = umma_desc_sw128(st), al0 = umma_desc_sw128(st + a_bytes);
                    const unsigned long long bh0 = umma_desc_sw128(st + (a.split ? 2u : 1u) * a_bytes), bl0 = umma_desc_sw128(st + (a.split ? 2u : 1u) * a_bytes + b_bytes);
                    const int steps = (kn + 7) >> 3;
                    for (int k = 0; k < steps; ++k) {
                        const unsigned long long kadv = (unsigned long long)(2 * k);
                        umma_tf32(acc, ah0 + kadv, bh0 + kadv, idesc, (ch == 0 && k == 0) ? 0u : 1u);
                        if (a.split) {
                            umma_tf32(acc, ah0 + kadv, bl0 + kadv, idesc, 1u);
                            umma_tf32(acc, al0 + kadv, bh0 + kadv, idesc, 1u);
                        }
                    }
                    if (a.cluster == 2)
                        umma_commit_multicast(smem_addr(&s_empty[s]), cmask);  // ... in both CTAs' books: either may refill the other's stage
                    else
                        umma_commit(smem_addr(&s_empty[s]));  // the stage is free once these MMAs have read it
                }
                umma_commit(smem_addr(&s_accf[t]));  // ... and the accumulator complete once they have written it
            }
        }
    } else {
        const int quad = warp & 3;  // TMEM lane quadrant this warp may read
        int u = 0;
        for (int unit = first_unit; unit < units; unit += unit_step, ++u) {
            const int mgrp = unit / a.nblocks, nblk = unit - mgrp * a.nblocks;
            const int r0 = (mgrp * a.cluster + crank) * GEMM_M, n0 = nblk * a.nb;
            const int ncols = min(a.nb, a.Nout - n0), npad = (ncols + 15) & ~15;
            const int t = u & 1;
            const int row = r0 + quad * 32 + lane;
            const bool live = row < a.T;
            const size_t obase = (size_t)(live ? row : 0) * a.ldo + n0;
            float4 res[4], bia[4];
            auto prefetch = [&](int c0) {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const bool ok = c0 + 4 * j < ncols;
                    bia[j] = ok ? ldg4(a.bias + n0 + c0 + 4 * j) : make_float4(0.f, 0.f, 0.f, 0.f);
                    res[j] = (ok && live && a.accumulate) ? *reinterpret_cast<const float4 *>(a.out + obase + c0 + 4 * j) : make_float4(0.f, 0.f, 0.f, 0.f);
                }
            };
            prefetch(0);
            mbar_wait_parity(smem_addr(&s_accf[t]), (unsigned)((u >> 1) & 1));
            fence_after_sync();
            for (int c0 = 0; c0 < npad; c0 += 16) {
                unsigned v[16];
                tmem_ld16(tmem + ((unsigned)(quad * 32) << 16) + (unsigned)(t * ACC_COLS + c0), v);
                float4 r[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    r[j] = make_float4(__uint_as_float(v[4 * j]) + bia[j].x + res[j].x, __uint_as_float(v[4 * j + 1]) + bia[j].y + res[j].y,
                                       __uint_as_float(v[4 * j + 2]) + bia[j].z + res[j].z, __uint_as_float(v[4 * j + 3]) + bia[j].w + res[j].w);
                    if (a.relu) r[j].x = fmaxf(r[j].x, 0.f), r[j].y = fmaxf(r[j].y, 0.f), r[j].z = fmaxf(r[j].z, 0.f), r[j].w = fmaxf(r[j].w, 0.f);
                }
                const int cur = c0;
                if (c0 + 16 < npad) prefetch(c0 + 16);
                if (live) {
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        if (cur + 4 * j < ncols) {
                            if (a.out) *reinterpret_cast<float4 *>(a.out + obase + cur + 4 * j) = r[j];
                            if (a.out_h) {
                                const float4 h = make_float4(tf32_round(r[j].x), tf32_round(r[j].y), tf32_round(r[j].z), tf32_round(r[j].w));
                                *reinterpret_cast<float4 *>(a.out_h + obase + cur + 4 * j) = h;
                                *reinterpret_cast<float4 *>(a.out_l + obase + cur + 4 * j) =
                                    make_float4(tf32_round(r[j].x - h.x), tf32_round(r[j].y - h.y), tf32_round(r[j].z - h.z), tf32_round(r[j].w - h.w));
                            }
                        }
                }
            }
            fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_addr(&s_acce[t]));
        }
    }
    __syncthreads();
    if (a.cluster == 2) cluster_sync_all();  // the peer may still be arriving on our barriers
    if (warp == 1) tmem_dealloc(tmem, 2 * ACC_COLS);
}

// ---- attention ----------------------------------------------------------------------------------------------------------------
struct AttArgs {
    const float *qkv;  // [T, 3D]: q | k | v, each head-major (column h * dim + d)
    float *o;          // [T, D] head-major
    float *oh, *ol;    // when set: O as its TF32 halves (the operand form the TMA-fed GEMM reads) instead of `o`
    int Bc, N, D, heads, dim, cross;
    float c;           // log2(e) / sqrt(dim)
};

// one value of O: FP32, or split into its TF32 halves
__device__ __forceinline__ void put_o(const AttArgs &a, size_t idx, float v) {
    if (a.oh) {
        const float h = __uint_as_float((__float_as_uint(v) + 0x1000u) & 0xffffe000u);
        a.oh[idx] = h;
        a.ol[idx] = __uint_as_float((__float_as_uint(v - h) + 0x1000u) & 0xffffe000u);
    } else {
        a.o[idx] = v;
    }
}

__device__ __forceinline__ float comp(const float4 &q, int dd) {
    return dd == 0 ? q.x : dd == 1 ? q.y : dd == 2 ? q.z : q.w;
}

// NJ = ceil(n / 32) key slots per lane; R query rows per warp pass; NW warps.  Value layout: PV2 = false: DI = ceil(dim / 32) scalar
// slots per lane (dim <= 32 DI: level 3, dim = 32); PV2 = true: dims [0, 64) as one float2 per lane, the dim - 64 <= 4 tail dims by a
// key-parallel reduction (level 2, dim = 66: a third scalar slot would idle 30 of 32 lanes).
template <int NJ, int DI, bool PV2, int R, int NW>
__global__ void __launch_bounds__(NW * 32) gnn_attention_kernel(AttArgs a) {
    extern __shared__ __align__(16) float sm[];
    constexpr int NP = NJ * 32 + 1, NP32 = NJ * 32, DV = PV2 ? 64 : DI * 32, NT = NW * 32, TAILMAX = 4;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int N = a.N, dim = a.dim, dim4 = (dim + 3) & ~3, D3 = 3 * a.D;
    const int tail = PV2 ? dim - 64 : 0;
    float *Q = sm;                         // [N][dim4]
    float *P = Q + N * dim4;               // [NW][NP32][R]
    float *V = P + NW * NP32 * R;          // [N][DV]
    float *Vt = V + N * DV;                // [N][TAILMAX]   (PV2 only)
    float *Kt = Vt + (PV2 ? N * TAILMAX : 0);  // [dim4][NP]
    pdl_prologue();
    const int ps = blockIdx.x / a.heads, h = blockIdx.x - ps * a.heads;
    const int side = ps / a.Bc, b = ps - side * a.Bc;
    const int sps = a.cross ? (1 - side) * a.Bc + b : ps;
    const float *qbase = a.qkv + (size_t)ps * N * D3 + h * dim;
    const float *kbase = a.qkv + (size_t)sps * N * D3 + a.D + h * dim;
    const float *vbase = kbase + a.D;
    const int half = dim4 >> 1;
    // staging: four pieces (12 loads) per thread in flight
    for (int it0 = tid; it0 < N * half; it0 += NT * 4) {
        float2 q2[4], k2[4], v2[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int it = it0 + u * NT;
            const int n = it / half, d = (it - n * half) * 2;
            q2[u] = k2[u] = v2[u] = make_float2(0.f, 0.f);
            if (it < N * half && d < dim) {
                q2[u] = __ldg(reinterpret_cast<const float2 *>(qbase + (size_t)n * D3 + d));
                k2[u] = __ldg(reinterpret_cast<const float2 *>(kbase + (size_t)n * D3 + d));
                v2[u] = __ldg(reinterpret_cast<const float2 *>(vbase + (size_t)n * D3 + d));
            }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int it = it0 + u * NT;
            if (it < N * half) {
                const int n = it / half, d = (it - n * half) * 2;
                *reinterpret_cast<float2 *>(Q + n * dim4 + d) = q2[u];
                Kt[d * NP + n] = k2[u].x, Kt[(d + 1) * NP + n] = k2[u].y;
                if (d < DV)
                    *reinterpret_cast<float2 *>(V + n * DV + d) = v2[u];
                else if (PV2)
                    Vt[n * TAILMAX + d - 64] = v2[u].x, Vt[n * TAILMAX + d - 63] = v2[u].y;
            }
        }
    }
    for (int it = tid; it < dim4 * (NP32 - N); it += NT) {  // keys beyond n: finite (masked after the products)
        const int d = it / (NP32 - N), m = N + it - d * (NP32 - N);
        Kt[d * NP + m] = 0.f;
    }
    if (!PV2 && DV > dim4)
        for (int it = tid; it < N * (DV - dim4); it += NT) {  // value columns beyond dim: zero
            const int n = it / (DV - dim4), d = dim4 + it - n * (DV - dim4);
            V[n * DV + d] = 0.f;
        }
    __syncthreads();
    float *Pw = P + warp * NP32 * R;
    for (int n0 = warp * R; n0 < N; n0 += NW * R) {
        float s[R][NJ];
#pragma unroll
        for (int r = 0; r < R; ++r)
#pragma unroll
            for (int j = 0; j < NJ; ++j) s[r][j] = 0.f;
        const float *qrow[R];
#pragma unroll
        for (int r = 0; r < R; ++r) qrow[r] = Q + min(n0 + r, N - 1) * dim4;
        for (int d = 0; d < dim4; d += 4) {
            float4 q[R];
#pragma unroll
            for (int r = 0; r < R; ++r) q[r] = *reinterpret_cast<const float4 *>(qrow[r] + d);
#pragma unroll
            for (int dd = 0; dd < 4; ++dd) {
                float kv[NJ];
#pragma unroll
                for (int j = 0; j < NJ; ++j) kv[j] = Kt[(d + dd) * NP + lane + 32 * j];
#pragma unroll
                for (int r = 0; r < R; ++r)
#pragma unroll
                    for (int j = 0; j < NJ; ++j) s[r][j] = fmaf(comp(q[r], dd), kv[j], s[r][j]);
            }
        }
        // softmax(s / sqrt(dim)) over the keys: exp2((s - max) * log2(e) / sqrt(dim)) / sum
#pragma unroll
        for (int r = 0; r < R; ++r) {
            float mx = -INFINITY;
#pragma unroll
            for (int j = 0; j < NJ; ++j) {
                if (lane + 32 * j >= N) s[r][j] = -INFINITY;
                mx = fmaxf(mx, s[r][j]);
            }
            mx = warp_max(mx);
            float sum = 0.f;
#pragma unroll
            for (int j = 0; j < NJ; ++j) {
                s[r][j] = fast_exp2((s[r][j] - mx) * a.c);
                sum += s[r][j];
            }
            sum = warp_sum(sum);
            const float inv = 1.0f / sum;
#pragma unroll
            for (int j = 0; j < NJ; ++j) s[r][j] *= inv;
        }
#pragma unroll
        for (int j = 0; j < NJ; ++j)
#pragma unroll
            for (int r4 = 0; r4 < R; r4 += 4)
                *reinterpret_cast<float4 *>(Pw + (lane + 32 * j) * R + r4) = make_float4(s[r4][j], s[r4 + 1][j], s[r4 + 2][j], s[r4 + 3][j]);
        __syncwarp();
        constexpr int OS = PV2 ? 2 : DI;  // value slots per lane
        float o[R][OS];
#pragma unroll
        for (int r = 0; r < R; ++r)
#pragma unroll
            for (int i = 0; i < OS; ++i) o[r][i] = 0.f;
#pragma unroll 2
        for (int m = 0; m < N; ++m) {
            float p[R];
#pragma unroll
            for (int r4 = 0; r4 < R; r4 += 4) {
                const float4 t = *reinterpret_cast<const float4 *>(Pw + m * R + r4);
                p[r4] = t.x, p[r4 + 1] = t.y, p[r4 + 2] = t.z, p[r4 + 3] = t.w;
            }
            float vv[OS];
            if (PV2) {
                const float2 t = *reinterpret_cast<const float2 *>(V + m * DV + 2 * lane);
                vv[0] = t.x, vv[OS - 1] = t.y;
            } else {
#pragma unroll
                for (int i = 0; i < OS; ++i) vv[i] = V[m * DV + lane + 32 * i];
            }
#pragma unroll
            for (int r = 0; r < R; ++r)
#pragma unroll
                for (int i = 0; i < OS; ++i) o[r][i] = fmaf(p[r], vv[i], o[r][i]);
        }
#pragma unroll
        for (int r = 0; r < R; ++r)
            if (n0 + r < N) {
                const size_t obase = ((size_t)ps * N + n0 + r) * a.D + h * dim;
                if (PV2) {
                    put_o(a, obase + 2 * lane, o[r][0]);
                    put_o(a, obase + 2 * lane + 1, o[r][OS - 1]);
                } else {
#pragma unroll
                    for (int i = 0; i < OS; ++i)
                        if (lane + 32 * i < dim) put_o(a, obase + lane + 32 * i, o[r][i]);
                }
            }
        if (PV2) {  // tail dims 64 .. dim - 1: every lane sums its own keys (the probabilities are still in registers), then one warp reduction
            for (int t = 0; t < tail; ++t) {
                float vt[NJ];
#pragma unroll
                for (int j = 0; j < NJ; ++j) vt[j] = lane + 32 * j < N ? Vt[(lane + 32 * j) * TAILMAX + t] : 0.f;
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    float acc = 0.f;
#pragma unroll
                    for (int j = 0; j < NJ; ++j) acc = fmaf(s[r][j], vt[j], acc);
                    acc = warp_sum(acc);
                    if (lane == 0 && n0 + r < N) put_o(a, ((size_t)ps * N + n0 + r) * a.D + h * dim + 64 + t, acc);
                }
            }
        }
        __syncwarp();  // the P buffer is rewritten by the next pass
    }
}

// Second generation of the resident-key kernel: the same sums in the same order (bit-identical to gnn_attention_kernel), on the packed
// FP32 pipe.  A warp owns 8 query rows as 4 row PAIRS: Q of its block sits transposed in the warp's own buffer ([dim][8 rows]: two
// broadcast LDS.128 per dimension deliver the four (q_r, q_r+1) pairs), every accumulator is a float2 over a row pair, and one
// fma.rn.f32x2 (SASS FFMA2) does the work of two FFMAs -- the key / value element is duplicated into a register pair with one MOV that
// serves four FFMA2.  Per dimension and 8 rows x 160 keys: 2 + 5 loads, 5 MOVs, 20 FFMA2 (first generation: 52 instructions, this: 32);
// per key in P V: 13 instead of 20.  The warp's buffer holds Q^T during Q K^T and P afterwards; Q is no longer staged for the CTA.
// VMODE 0: dim <= 32, one value slot per lane; VMODE 1: dims [0, 64) as a float2 per lane + the <= 4 tail dims by a key-parallel reduction.
template <int NJ, int VMODE, int NW, int R>
__global__ void __launch_bounds__(NW * 32) gnn_attention2_kernel(AttArgs a) {
    extern __shared__ __align__(16) float sm[];
    constexpr int NP = NJ * 32 + 1, NP32 = NJ * 32, DV = VMODE ? 64 : 32, NT = NW * 32, TAILMAX = 4;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int N = a.N, dim = a.dim, dim4 = (dim + 3) & ~3, D3 = 3 * a.D;
    const int tail = VMODE ? dim - 64 : 0;
    float *P = sm;                               // [NW][NP32][R]   (Q^T of the warp's block first: [dim4][R])
    float *V = P + NW * NP32 * R;                // [N][DV]
    float *Vt = V + N * DV;                      // [N][TAILMAX]    (VMODE 1 only)
    float *Kt = Vt + (VMODE ? N * TAILMAX : 0);  // [dim4][NP]
    pdl_prologue();
    const int ps = blockIdx.x / a.heads, h = blockIdx.x - ps * a.heads;
    const int side = ps / a.Bc, b = ps - side * a.Bc;
    const int sps = a.cross ? (1 - side) * a.Bc + b : ps;
    const float *qbase = a.qkv + (size_t)ps * N * D3 + h * dim;
    const float *kbase = a.qkv + (size_t)sps * N * D3 + a.D + h * dim;
    const float *vbase = kbase + a.D;
    const int half = dim4 >> 1;
    for (int it0 = tid; it0 < N * half; it0 += NT * 4) {
        float2 k2[4], v2[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int it = it0 + u * NT;
            const int n = it / half, d = (it - n * half) * 2;
            k2[u] = v2[u] = make_float2(0.f, 0.f);
            if (it < N * half && d < dim) {
                k2[u] = __ldg(reinterpret_cast<const float2 *>(kbase + (size_t)n * D3 + d));
                v2[u] = __ldg(reinterpret_cast<const float2 *>(vbase + (size_t)n * D3 + d));
            }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int it = it0 + u * NT;
            if (it < N * half) {
                const int n = it / half, d = (it - n * half) * 2;
                Kt[d * NP + n] = k2[u].x, Kt[(d + 1) * NP + n] = k2[u].y;
                if (d < DV)
                    *reinterpret_cast<float2 *>(V + n * DV + d) = v2[u];
                else if (VMODE)
                    Vt[n * TAILMAX + d - 64] = v2[u].x, Vt[n * TAILMAX + d - 63] = v2[u].y;
            }
        }
    }
    for (int it = tid; it < dim4 * (NP32 - N); it += NT) {  // keys beyond n: finite (masked after the products)
        const int d = it / (NP32 - N), m = N + it - d * (NP32 - N);
        Kt[d * NP + m] = 0.f;
    }
    if (!VMODE && DV > dim4)
        for (int it = tid; it < N * (DV - dim4); it += NT) {  // value columns beyond dim: zero
            const int n = it / (DV - dim4), d = dim4 + it - n * (DV - dim4);
            V[n * DV + d] = 0.f;
        }
    __syncthreads();
    float *Pw = P + warp * NP32 * R;
    for (int n0 = warp * R; n0 < N; n0 += NW * R) {
        // ---- Q^T of this block into the warp's buffer: Pw[d * 8 + r] ----
        for (int it = lane; it < R * half; it += 32) {
            const int r = it / half, d = (it - r * half) * 2;
            float2 q2 = make_float2(0.f, 0.f);
            if (d < dim) q2 = __ldg(reinterpret_cast<const float2 *>(qbase + (size_t)min(n0 + r, N - 1) * D3 + d));
            Pw[d * R + r] = q2.x, Pw[(d + 1) * R + r] = q2.y;
        }
        __syncwarp();
        float2 s2[R / 2][NJ];
#pragma unroll
        for (int rp = 0; rp < R / 2; ++rp)
#pragma unroll
            for (int j = 0; j < NJ; ++j) s2[rp][j] = make_float2(0.f, 0.f);
#pragma unroll 4
        for (int d = 0; d < dim4; ++d) {
            float2 qp[R / 2];
#pragma unroll
            for (int r4 = 0; r4 < R; r4 += 4) {
                const float4 t = *reinterpret_cast<const float4 *>(Pw + d * R + r4);
                qp[r4 / 2] = make_float2(t.x, t.y), qp[r4 / 2 + 1] = make_float2(t.z, t.w);
            }
#pragma unroll
            for (int j = 0; j < NJ; ++j) {
                const float kv = Kt[d * NP + lane + 32 * j];
                const float2 kk = make_float2(kv, kv);
#pragma unroll
                for (int rp = 0; rp < R / 2; ++rp) s2[rp][j] = ffma2(qp[rp], kk, s2[rp][j]);
            }
        }
        __syncwarp();  // Q^T has been read; the buffer becomes P
        float s[R][NJ];
#pragma unroll
        for (int rp = 0; rp < R / 2; ++rp)
#pragma unroll
            for (int j = 0; j < NJ; ++j) s[2 * rp][j] = s2[rp][j].x, s[2 * rp + 1][j] = s2[rp][j].y;
#pragma unroll
        for (int r = 0; r < R; ++r) {
            float mx = -INFINITY;
#pragma unroll
            for (int j = 0; j < NJ; ++j) {
                if (lane + 32 * j >= N) s[r][j] = -INFINITY;
                mx = fmaxf(mx, s[r][j]);
            }
            mx = warp_max(mx);
            float sum = 0.f;
#pragma unroll
            for (int j = 0; j < NJ; ++j) {
                s[r][j] = fast_exp2((s[r][j] - mx) * a.c);
                sum += s[r][j];
            }
            sum = warp_sum(sum);
            const float inv = 1.0f / sum;
#pragma unroll
            for (int j = 0; j < NJ; ++j) s[r][j] *= inv;
        }
#pragma unroll
        for (int j = 0; j < NJ; ++j)
#pragma unroll
            for (int r4 = 0; r4 < R; r4 += 4)
                *reinterpret_cast<float4 *>(Pw + (lane + 32 * j) * R + r4) = make_float4(s[r4][j], s[r4 + 1][j], s[r4 + 2][j], s[r4 + 3][j]);
        __syncwarp();
        constexpr int OS = VMODE ? 2 : 1;  // value slots per lane
        float2 o2[R / 2][OS];
#pragma unroll
        for (int rp = 0; rp < R / 2; ++rp)
#pragma unroll
            for (int i = 0; i < OS; ++i) o2[rp][i] = make_float2(0.f, 0.f);
#pragma unroll 4
        for (int m = 0; m < N; ++m) {
            float2 pp[R / 2];
#pragma unroll
            for (int r4 = 0; r4 < R; r4 += 4) {
                const float4 t = *reinterpret_cast<const float4 *>(Pw + m * R + r4);
                pp[r4 / 2] = make_float2(t.x, t.y), pp[r4 / 2 + 1] = make_float2(t.z, t.w);
            }
            float vv[OS];
            if (VMODE) {
                const float2 t = *reinterpret_cast<const float2 *>(V + m * DV + 2 * lane);
                vv[0] = t.x, vv[OS - 1] = t.y;
            } else {
                vv[0] = V[m * DV + lane];
            }
#pragma unroll
            for (int i = 0; i < OS; ++i) {
                const float2 vd = make_float2(vv[i], vv[i]);
#pragma unroll
                for (int rp = 0; rp < R / 2; ++rp) o2[rp][i] = ffma2(pp[rp], vd, o2[rp][i]);
            }
        }
#pragma unroll
        for (int r = 0; r < R; ++r)
            if (n0 + r < N) {
                const size_t obase = ((size_t)ps * N + n0 + r) * a.D + h * dim;
                const float2 *op = o2[r >> 1];
                if (VMODE) {
                    put_o(a, obase + 2 * lane, (r & 1) ? op[0].y : op[0].x);
                    put_o(a, obase + 2 * lane + 1, (r & 1) ? op[OS - 1].y : op[OS - 1].x);
                } else if (lane < dim) {
                    put_o(a, obase + lane, (r & 1) ? op[0].y : op[0].x);
                }
            }
        if (VMODE) {  // tail dims 64 .. dim - 1: every lane sums its own keys (the probabilities are still in registers), then one warp reduction
            for (int t = 0; t < tail; ++t) {
                float vt[NJ];
#pragma unroll
                for (int j = 0; j < NJ; ++j) vt[j] = lane + 32 * j < N ? Vt[(lane + 32 * j) * TAILMAX + t] : 0.f;
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    float acc = 0.f;
#pragma unroll
                    for (int j = 0; j < NJ; ++j) acc = fmaf(s[r][j], vt[j], acc);
                    acc = warp_sum(acc);
                    if (lane == 0 && n0 + r < N) put_o(a, ((size_t)ps * N + n0 + r) * a.D + h * dim + 64 + t, acc);
                }
            }
        }
        __syncwarp();  // the buffer is rewritten by the next pass
    }
}

// Any token count (level 1: n = 300 at 640 x 480, 1024 at 1024 x 1024; head dimension 112): one CTA per (problem, side, head, tile of
// NW * R * RB query rows), the keys in chunks of NJ * 32 with the running maximum / sum of an online softmax, the output rescaled when
// the maximum moves ("flash" formulation; identical to the plain softmax up to FP32 rounding).
template <int NJ, int DI, int R, int NW, int RB>
__global__ void __launch_bounds__(NW * 32) gnn_attention_flash_kernel(AttArgs a, int qtiles) {
    extern __shared__ __align__(16) float sm[];
    constexpr int NP = NJ * 32 + 1, NP32 = NJ * 32, DV = DI * 32, NT = NW * 32, QT = NW * R * RB;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int N = a.N, dim = a.dim, dim4 = (dim + 3) & ~3, D3 = 3 * a.D;
    float *Q = sm;                     // [QT][dim4]
    float *P = Q + QT * dim4;          // [NW][NP32][R]
    float *V = P + NW * NP32 * R;      // [NP32][DV]
    float *Kt = V + NP32 * DV;         // [dim4][NP]
    pdl_prologue();
    const int qt = blockIdx.x % qtiles, rest = blockIdx.x / qtiles;
    const int ps = rest / a.heads, h = rest - ps * a.heads;
    const int side = ps / a.Bc, b = ps - side * a.Bc;
    const int sps = a.cross ? (1 - side) * a.Bc + b : ps;
    const float *qbase = a.qkv + (size_t)ps * N * D3 + h * dim;
    const float *kbase = a.qkv + (size_t)sps * N * D3 + a.D + h * dim;
    const float *vbase = kbase + a.D;
    const int q0 = qt * QT, half = dim4 >> 1;
    for (int it = tid; it < QT * half; it += NT) {
        const int n = it / half, d = (it - n * half) * 2;
        float2 q2 = make_float2(0.f, 0.f);
        if (q0 + n < N && d < dim) q2 = __ldg(reinterpret_cast<const float2 *>(qbase + (size_t)(q0 + n) * D3 + d));
        *reinterpret_cast<float2 *>(Q + n * dim4 + d) = q2;
    }
    float o[RB][R][DI], mrun[RB][R], lrun[RB][R];
#pragma unroll
    for (int rb = 0; rb < RB; ++rb)
#pragma unroll
        for (int r = 0; r < R; ++r) {
            mrun[rb][r] = -INFINITY, lrun[rb][r] = 0.f;
#pragma unroll
            for (int i = 0; i < DI; ++i) o[rb][r][i] = 0.f;
        }
    float *Pw = P + warp * NP32 * R;
    for (int kc0 = 0; kc0 < N; kc0 += NP32) {
        const int kn = min(NP32, N - kc0);
        __syncthreads();  // the previous chunk's K / V are no longer read
        for (int it0 = tid; it0 < NP32 * half; it0 += NT * 4) {
            float2 k2[4], v2[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int it = it0 + u * NT;
                const int n = it / half, d = (it - n * half) * 2;
                k2[u] = v2[u] = make_float2(0.f, 0.f);
                if (it < NP32 * half && n < kn && d < dim) {
                    k2[u] = __ldg(reinterpret_cast<const float2 *>(kbase + (size_t)(kc0 + n) * D3 + d));
                    v2[u] = __ldg(reinterpret_cast<const float2 *>(vbase + (size_t)(kc0 + n) * D3 + d));
                }
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int it = it0 + u * NT;
                if (it < NP32 * half) {
                    const int n = it / half, d = (it - n * half) * 2;
                    Kt[d * NP + n] = k2[u].x, Kt[(d + 1) * NP + n] = k2[u].y;
                    *reinterpret_cast<float2 *>(V + n * DV + d) = v2[u];
                }
            }
        }
        if (DV > dim4)
            for (int it = tid; it < NP32 * (DV - dim4); it += NT) {
                const int n = it / (DV - dim4), d = dim4 + it - n * (DV - dim4);
                V[n * DV + d] = 0.f;
            }
        __syncthreads();
#pragma unroll
        for (int rb = 0; rb < RB; ++rb) {
            const int l0 = (rb * NW + warp) * R;  // first row of this block inside the tile
            if (q0 + l0 >= N) continue;
            float s[R][NJ];
#pragma unroll
            for (int r = 0; r < R; ++r)
#pragma unroll
                for (int j = 0; j < NJ; ++j) s[r][j] = 0.f;
            for (int d = 0; d < dim4; d += 4) {
                float4 q[R];
#pragma unroll
                for (int r = 0; r < R; ++r) q[r] = *reinterpret_cast<const float4 *>(Q + (l0 + r) * dim4 + d);
#pragma unroll
                for (int dd = 0; dd < 4; ++dd) {
                    float kv[NJ];
#pragma unroll
                    for (int j = 0; j < NJ; ++j) kv[j] = Kt[(d + dd) * NP + lane + 32 * j];
#pragma unroll
                    for (int r = 0; r < R; ++r)
#pragma unroll
                        for (int j = 0; j < NJ; ++j) s[r][j] = fmaf(comp(q[r], dd), kv[j], s[r][j]);
                }
            }
#pragma unroll
            for (int r = 0; r < R; ++r) {
                float mx = -INFINITY;
#pragma unroll
                for (int j = 0; j < NJ; ++j) {
                    if (lane + 32 * j >= kn) s[r][j] = -INFINITY;
                    mx = fmaxf(mx, s[r][j]);
                }
                const float mnew = fmaxf(mrun[rb][r], warp_max(mx));
                const float sc = fast_exp2((mrun[rb][r] - mnew) * a.c);  // 0 on the first chunk (running maximum -inf)
                float sum = 0.f;
#pragma unroll
                for (int j = 0; j < NJ; ++j) {
                    s[r][j] = fast_exp2((s[r][j] - mnew) * a.c);
                    sum += s[r][j];
                }
                lrun[rb][r] = lrun[rb][r] * sc + warp_sum(sum);
                mrun[rb][r] = mnew;
#pragma unroll
                for (int i = 0; i < DI; ++i) o[rb][r][i] *= sc;
            }
#pragma unroll
            for (int j = 0; j < NJ; ++j)
#pragma unroll
                for (int r4 = 0; r4 < R; r4 += 4)
                    *reinterpret_cast<float4 *>(Pw + (lane + 32 * j) * R + r4) = make_float4(s[r4][j], s[r4 + 1][j], s[r4 + 2][j], s[r4 + 3][j]);
            __syncwarp();
#pragma unroll 2
            for (int m = 0; m < kn; ++m) {
                float p[R];
#pragma unroll
                for (int r4 = 0; r4 < R; r4 += 4) {
                    const float4 t = *reinterpret_cast<const float4 *>(Pw + m * R + r4);
                    p[r4] = t.x, p[r4 + 1] = t.y, p[r4 + 2] = t.z, p[r4 + 3] = t.w;
                }
                float vv[DI];
#pragma unroll
                for (int i = 0; i < DI; ++i) vv[i] = V[m * DV + lane + 32 * i];
#pragma unroll
                for (int r = 0; r < R; ++r)
#pragma unroll
                    for (int i = 0; i < DI; ++i) o[rb][r][i] = fmaf(p[r], vv[i], o[rb][r][i]);
            }
            __syncwarp();  // the P buffer is rewritten by the next block
        }
    }
#pragma unroll
    for (int rb = 0; rb < RB; ++rb)
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int n = q0 + (rb * NW + warp) * R + r;
            if (n < N) {
                const float inv = 1.0f / lrun[rb][r];
                const size_t obase = ((size_t)ps * N + n) * a.D + h * dim;
#pragma unroll
                for (int i = 0; i < DI; ++i)
                    if (lane + 32 * i < dim) put_o(a, obase + lane + 32 * i, o[rb][r][i] * inv);
            }
        }
}

// ---- layout changes at entry / exit: [b, D, n] (the reference's Conv1d layout) <-> token-major [T, D] --------------------------
struct TransArgs {
    const float *d0, *d1;  // entry: inputs [B, D, N] per side
    float *o0, *o1;        // exit: outputs [B, D, N] per side
    float *X;              // [2 * Bc * N, D]
    float *Xh, *Xl;        // entry, when set: the TF32 halves of X as well
    int b0, Bc, D, N;      // problems [b0, b0 + Bc) of the batch
    int to_tokens;
};

__global__ void __launch_bounds__(256) gnn_transpose_kernel(TransArgs a) {
    __shared__ float tile[32][33];
    pdl_prologue();
    const int ps = blockIdx.z, side = ps / a.Bc, b = ps - side * a.Bc;
    const int nb = blockIdx.x * 32, db = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const size_t off = (size_t)(a.b0 + b) * a.D * a.N;
    float *X = a.X + (size_t)ps * a.N * a.D;
    if (a.to_tokens) {
        const float *src = (side ? a.d1 : a.d0) + off;
        for (int i = ty; i < 32; i += 8)
            if (db + i < a.D && nb + tx < a.N) tile[i][tx] = src[(size_t)(db + i) * a.N + nb + tx];
        __syncthreads();
        for (int i = ty; i < 32; i += 8)
            if (nb + i < a.N && db + tx < a.D) {
                const size_t idx = (size_t)(nb + i) * a.D + db + tx;
                const float v = tile[tx][i];
                X[idx] = v;
                if (a.Xh) {
                    const float h = tf32_round(v);
                    a.Xh[(size_t)ps * a.N * a.D + idx] = h;
                    a.Xl[(size_t)ps * a.N * a.D + idx] = tf32_round(v - h);
                }
            }
    } else {
        float *dst = (side ? a.o1 : a.o0) + off;
        for (int i = ty; i < 32; i += 8)
            if (nb + i < a.N && db + tx < a.D) tile[i][tx] = X[(size_t)(nb + i) * a.D + db + tx];
        __syncthreads();
        for (int i = ty; i < 32; i += 8)
            if (db + i < a.D && nb + tx < a.N) dst[(size_t)(db + i) * a.N + nb + tx] = tile[tx][i];
    }
}

// ---- weight packing (device, FP64 accumulation) ----------------------------------------------------------------------------------
// raw layer layout (floats), the reference's parameters in this order (models/modules.py:92-112):
//   Wq[D*D] bq[D] Wk[D*D] bk[D] Wv[D*D] bv[D]       attn.proj.0 / 1 / 2  (Conv1d weight [out, in, 1])
//   Wm[D*D] bm[D]                                   attn.merge
//   W1[2D*2D] b1[2D]                                mlp.0
//   gamma[2D] beta[2D] mean[2D] var[2D]             mlp.1 (BatchNorm1d weight, bias, running_mean, running_var)
//   W2[D*2D] b2[D]                                  mlp.3
// packed layer layout: Wqkv[3D*D] bqkv[3D] W1f[2D*2D] b1f[2D] W2[D*2D] b2[D]
__host__ __device__ inline size_t raw_layer_floats(int D) { return (size_t)4 * D * D + 4 * D + (size_t)4 * D * D + 2 * D + 8 * D + (size_t)2 * D * D + D; }
__host__ __device__ inline size_t packed_layer_floats(int D) { return (size_t)9 * D * D + 6 * D; }
// behind the FP32 layers: per layer Wqkv_hi[3DD] Wqkv_lo[3DD] W1f_hi[4DD] W1f_lo[4DD] W2_hi[2DD] W2_lo[2DD] -- the TF32 halves the TMA-fed GEMM reads
__host__ __device__ inline size_t halves_layer_floats(int D) { return (size_t)18 * D * D; }

__global__ void gnn_pack_kernel(const float *raw, float *packed, int layers, int D, int heads, float eps, int fold_bn) {
    const int dim = D / heads, D2 = 2 * D;
    const size_t per = packed_layer_floats(D);
    const size_t total = per * layers;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int l = (int)(i / per);
        size_t e = i - (size_t)l * per;
        const float *r = raw + (size_t)l * raw_layer_floats(D);
        const size_t DD = (size_t)D * D;
        const float *Wm = r + 3 * (DD + D), *bm = Wm + DD;
        const float *W1 = bm + D, *b1 = W1 + (size_t)D2 * D2;
        const float *gamma = b1 + D2, *beta = gamma + D2, *mean = beta + D2, *var = mean + D2;
        const float *W2 = var + D2;
        float val;
        if (e < 3 * DD) {  // Wqkv: row c' = h * dim + d of block p <- row c = d * heads + h of proj.p
            const int p = (int)(e / DD);
            const int cp = (int)((e - p * DD) / D), k = (int)(e - p * DD - (size_t)cp * D);
            const int hh = cp / dim, d = cp - hh * dim;
            val = r[(size_t)p * (DD + D) + (size_t)(d * heads + hh) * D + k];
        } else if ((e -= 3 * DD) < (size_t)3 * D) {
            const int p = (int)(e / D), cp = (int)(e - (size_t)p * D);
            const int hh = cp / dim, d = cp - hh * dim;
            val = r[(size_t)p * (DD + D) + DD + d * heads + hh];
        } else if ((e -= 3 * D) < (size_t)D2 * D2) {  // W1f
            const int o = (int)(e / D2), c = (int)(e - (size_t)o * D2);
            const double sc = fold_bn ? (double)gamma[o] / sqrt((double)var[o] + (double)eps) : 1.0;  // train(): the BatchNorm stays a separate step
            double acc;
            if (c < D) {
                acc = W1[(size_t)o * D2 + c];
            } else {
                const int cp = c - D, hh = cp / dim, d = cp - hh * dim, old = d * heads + hh;
                acc = 0.0;
                for (int k = 0; k < D; ++k) acc += (double)W1[(size_t)o * D2 + D + k] * (double)Wm[(size_t)k * D + old];
            }
            val = (float)(acc * sc);
        } else if ((e -= (size_t)D2 * D2) < (size_t)D2) {  // b1f = (b1 + W1[:, D:] bm - mean) * scale + beta
            const int o = (int)e;
            const double sc = (double)gamma[o] / sqrt((double)var[o] + (double)eps);
            double acc = b1[o];
            for (int k = 0; k < D; ++k) acc += (double)W1[(size_t)o * D2 + D + k] * (double)bm[k];
            val = fold_bn ? (float)((acc - (double)mean[o]) * sc + (double)beta[o]) : (float)acc;
        } else {  // W2, b2 verbatim
            e -= D2;
            val = W2[e];
        }
        packed[i] = val;
    }
}

// ---- BatchNorm1d in train() mode (models/pats.py:112-119 keeps the third layer's network in train() when `if_local` is False): the
// first MLP convolution writes its FP32 result Z, the statistics of each side's batch -- `layer(desc0, src0)` and `layer(desc1, src1)`
// are two BatchNorm calls (models/modules.py:131) -- are taken over all its tokens, and the normalisation + ReLU produces the TF32
// halves the second convolution reads.  Running statistics are updated as torch does (momentum, unbiased variance), side 0 first.
struct BnArgs {
    const float *Z;          // [T, C] pre-normalisation, T = 2 * Th (side-major)
    double *stats;           // [2 sides][C][2]: sum, sum of squares (zero on entry of the statistics kernel)
    const float *gamma, *beta;
    float *running;          // [2][C]: running_mean, running_var of this layer (updated in place)
    float *Yh, *Yl;          // [T, C] TF32 halves of relu(bn(Z))
    int Th, C;
    float eps, momentum;
};

constexpr int BN_ROWS = 128;

__global__ void __launch_bounds__(256) gnn_bn_stats_kernel(BnArgs a) {
    pdl_prologue();
    const int side = blockIdx.y, r0 = blockIdx.x * BN_ROWS, r1 = min(r0 + BN_ROWS, a.Th);
    for (int c = threadIdx.x; c < a.C; c += blockDim.x) {
        const float *z = a.Z + ((size_t)side * a.Th + r0) * a.C + c;
        double s = 0.0, q = 0.0;
        for (int r = r0; r < r1; ++r, z += a.C) {
            const double v = (double)__ldg(z);
            s += v, q += v * v;
        }
        atomicAdd(a.stats + ((size_t)side * a.C + c) * 2, s);
        atomicAdd(a.stats + ((size_t)side * a.C + c) * 2 + 1, q);
    }
}

__global__ void __launch_bounds__(256) gnn_bn_apply_kernel(BnArgs a) {
    pdl_prologue();
    const int side = blockIdx.y, r0 = blockIdx.x * BN_ROWS, r1 = min(r0 + BN_ROWS, a.Th);
    const double n = (double)a.Th;
    for (int c4 = threadIdx.x * 4; c4 < a.C; c4 += blockDim.x * 4) {
        float sc[4], sh[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const double *st = a.stats + ((size_t)side * a.C + c4 + j) * 2;
            const double mean = st[0] / n, var = fmax(st[1] / n - mean * mean, 0.0);
            const double scale = (double)a.gamma[c4 + j] / sqrt(var + (double)a.eps);
            sc[j] = (float)scale, sh[j] = (float)((double)a.beta[c4 + j] - mean * scale);
        }
        for (int r = r0; r < r1; ++r) {
            const size_t idx = ((size_t)side * a.Th + r) * a.C + c4;
            const float4 z = *reinterpret_cast<const float4 *>(a.Z + idx);
            const float4 y = make_float4(fmaxf(fmaf(z.x, sc[0], sh[0]), 0.f), fmaxf(fmaf(z.y, sc[1], sh[1]), 0.f), fmaxf(fmaf(z.z, sc[2], sh[2]), 0.f),
                                         fmaxf(fmaf(z.w, sc[3], sh[3]), 0.f));
            const float4 h = make_float4(tf32_round(y.x), tf32_round(y.y), tf32_round(y.z), tf32_round(y.w));
            *reinterpret_cast<float4 *>(a.Yh + idx) = h;
            *reinterpret_cast<float4 *>(a.Yl + idx) = make_float4(tf32_round(y.x - h.x), tf32_round(y.y - h.y), tf32_round(y.z - h.z), tf32_round(y.w - h.w));
        }
    }
    if (blockIdx.x == 0 && blockIdx.y == 0) {  // running statistics: the two BatchNorm calls of the layer in the reference's order
        for (int c = threadIdx.x; c < a.C; c += blockDim.x) {
            float rm = a.running[c], rv = a.running[a.C + c];
            for (int sd = 0; sd < 2; ++sd) {
                const double *st = a.stats + ((size_t)sd * a.C + c) * 2;
                const double mean = st[0] / n, var = fmax(st[1] / n - mean * mean, 0.0);
                const double unbiased = n > 1.0 ? var * n / (n - 1.0) : var;
                rm = (1.0f - a.momentum) * rm + a.momentum * (float)mean;
                rv = (1.0f - a.momentum) * rv + a.momentum * (float)unbiased;
            }
            a.running[c] = rm, a.running[a.C + c] = rv;
        }
    }
}

__global__ void gnn_split_weights_kernel(float *packed, int layers, int D) {
    const size_t DD = (size_t)D * D, per = packed_layer_floats(D), per2 = halves_layer_floats(D);
    const size_t total = (size_t)layers * 9 * DD;
    float *halves = packed + (size_t)layers * per;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int l = (int)(i / (9 * DD));
        const size_t e = i - (size_t)l * 9 * DD;
        const float *w = packed + (size_t)l * per;
        float *h = halves + (size_t)l * per2;
        float x;
        size_t hi_at, lo_at;
        if (e < 3 * DD) x = w[e], hi_at = e, lo_at = 3 * DD + e;                                              // Wqkv
        else if (e < 7 * DD) x = w[3 * DD + 3 * D + (e - 3 * DD)], hi_at = 6 * DD + (e - 3 * DD), lo_at = 10 * DD + (e - 3 * DD);  // W1f
        else x = w[7 * DD + 5 * D + (e - 7 * DD)], hi_at = 14 * DD + (e - 7 * DD), lo_at = 16 * DD + (e - 7 * DD);                  // W2
        const float hv = tf32_round(x);
        h[hi_at] = hv;
        h[lo_at] = tf32_round(x - hv);
    }
}

// ---- tensor maps (host) ------------------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                                  const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_tiled() {
    static std::atomic<void *> cached{nullptr};
    void *fn = cached.load(std::memory_order_acquire);
    if (!fn) {
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) return nullptr;
        cached.store(fn, std::memory_order_release);
    }
    return reinterpret_cast<EncodeTiledFn>(fn);
}
// [rows, cols] FP32, row stride ld floats; box = 32 columns (128 bytes, the swizzle span) x box_rows; out-of-range elements read as zero
int make_map_2d(CUtensorMap *m, const float *base, int rows, int cols, int ld, int box_rows) {
    EncodeTiledFn enc = encode_tiled();
    if (!enc) return invalid("attentional_gnn: cuTensorMapEncodeTiled is not available from this driver");
    const cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows}, strides[1] = {(cuuint64_t)ld * 4};
    const cuuint32_t box[2] = {32u, (cuuint32_t)box_rows}, es[2] = {1u, 1u};
    const CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float *>(base), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                           CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? PATS_OK : invalid("attentional_gnn: cuTensorMapEncodeTiled failed (%d) for a [%d, %d] tensor", (int)r, rows, cols);
}
// [layers, rows, cols] weights, one layer every `layer_stride` floats
int make_map_3d(CUtensorMap *m, const float *base, int layers, int rows, int cols, size_t layer_stride, int box_rows) {
    EncodeTiledFn enc = encode_tiled();
    if (!enc) return invalid("attentional_gnn: cuTensorMapEncodeTiled is not available from this driver");
    const cuuint64_t dims[3] = {(cuuint64_t)cols, (cuuint64_t)rows, (cuuint64_t)layers}, strides[2] = {(cuuint64_t)cols * 4, (cuuint64_t)layer_stride * 4};
    const cuuint32_t box[3] = {32u, (cuuint32_t)box_rows, 1u}, es[3] = {1u, 1u, 1u};
    const CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float *>(base), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                           CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? PATS_OK : invalid("attentional_gnn: cuTensorMapEncodeTiled failed (%d) for [%d, %d, %d] weights", (int)r, layers, rows, cols);
}

// block size along Nout for the TMA-fed kernel and the number of stages that fit
struct TmaShape {
    int nb, nblocks, stages;
    size_t smem;
};
// The widest block that fits three stages moves the fewest operand bytes per flop -- what counts when there are many waves of units.
// With few token blocks (the chunked calls of an `if_local` forward: 91 blocks at 40 windows) the last wave decides: among the block
// widths that fit, take the one with the smallest  waves x (128 + nb)  -- rounds of the persistent CTAs times operand bytes per K chunk.
TmaShape tma_shape(int Nout, bool split, int mblocks, int sms) {
    const size_t budget = 225 * 1024;
    int max_nb = split ? 160 : 256;  // three stages of (128 + nb) x 128 B x 2 halves must fit
    if (const char *e = getenv("PATS_GNN_MAX_NB")) {  // A/B of the block width (tools/gnn_small.py)
        const int v = atoi(e);
        if (v >= 16 && v <= 256) max_nb = v & ~15;
    }
    long long best_cost = -1;
    TmaShape best = {};
    const int first = (Nout + max_nb - 1) / max_nb;
    for (int nblocks = first; nblocks <= first + 6; ++nblocks) {
        TmaShape t;
        t.nb = (((Nout + nblocks - 1) / nblocks) + 15) & ~15;
        t.nblocks = (Nout + t.nb - 1) / t.nb;
        const long long units = (long long)mblocks * t.nblocks, waves = (units + sms - 1) / sms;
        const long long cost = waves * (GEMM_M + t.nb);
        if (best_cost < 0 || cost < best_cost) best_cost = cost, best = t;
        if (t.nb <= 48) break;
    }
    const size_t stage = (size_t)(GEMM_M + best.nb) * 128 * (split ? 2 : 1);
    best.stages = (int)((budget - 1024) / stage);
    if (best.stages > 6) best.stages = 6;
    best.smem = (size_t)best.stages * stage + 1024;
    return best;
}

template <int NJ, int DI, bool PV2, int R, int NW>
int launch_attention(const AttArgs &a, cudaStream_t st, int dev) {
    const int dim4 = (a.dim + 3) & ~3;
    const size_t smem = sizeof(float) * ((size_t)a.N * dim4 + (size_t)NW * NJ * 32 * R + (size_t)a.N * (PV2 ? 64 + 4 : DI * 32) + (size_t)dim4 * (NJ * 32 + 1));
    static PerDeviceOnce configured;
    if (!configured.done(dev)) {
        PATS_CUDA_TRY(cudaFuncSetAttribute(gnn_attention_kernel<NJ, DI, PV2, R, NW>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
        configured.mark(dev);
    }
    if (smem > 220 * 1024) return invalid("attentional_gnn: attention tile of %zu bytes exceeds shared memory (n = %d, head dim = %d)", smem, a.N, a.dim);
    PATS_CUDA_TRY(launch_chained(gnn_attention_kernel<NJ, DI, PV2, R, NW>, dim3((unsigned)(2 * a.Bc * a.heads)), dim3(NW * 32), smem, st, a));
    return PATS_OK;
}

template <int NJ, int VMODE, int NW, int R = 8>
int launch_attention2(const AttArgs &a, cudaStream_t st, int dev) {
    const int dim4 = (a.dim + 3) & ~3;
    const size_t smem = sizeof(float) * ((size_t)NW * NJ * 32 * R + (size_t)a.N * (VMODE ? 64 + 4 : 32) + (size_t)dim4 * (NJ * 32 + 1));
    static PerDeviceOnce configured;
    if (!configured.done(dev)) {
        PATS_CUDA_TRY(cudaFuncSetAttribute(gnn_attention2_kernel<NJ, VMODE, NW, R>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
        configured.mark(dev);
    }
    if (smem > 220 * 1024) return invalid("attentional_gnn: attention tile of %zu bytes exceeds shared memory (n = %d, head dim = %d)", smem, a.N, a.dim);
    PATS_CUDA_TRY(launch_chained(gnn_attention2_kernel<NJ, VMODE, NW, R>, dim3((unsigned)(2 * a.Bc * a.heads)), dim3(NW * 32), smem, st, a));
    return PATS_OK;
}

template <int NJ, int DI, int R, int NW, int RB>
int launch_attention_flash(const AttArgs &a, cudaStream_t st, int dev) {
    constexpr int QT = NW * R * RB;
    const int dim4 = (a.dim + 3) & ~3;
    const size_t smem = sizeof(float) * ((size_t)QT * dim4 + (size_t)NW * NJ * 32 * R + (size_t)NJ * 32 * DI * 32 + (size_t)dim4 * (NJ * 32 + 1));
    static PerDeviceOnce configured;
    if (!configured.done(dev)) {
        PATS_CUDA_TRY(cudaFuncSetAttribute(gnn_attention_flash_kernel<NJ, DI, R, NW, RB>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
        configured.mark(dev);
    }
    if (smem > 220 * 1024) return invalid("attentional_gnn: attention tile of %zu bytes exceeds shared memory (head dim = %d)", smem, a.dim);
    const int qtiles = (a.N + QT - 1) / QT;
    PATS_CUDA_TRY(launch_chained(gnn_attention_flash_kernel<NJ, DI, R, NW, RB>, dim3((unsigned)(2 * a.Bc * a.heads * qtiles)), dim3(NW * 32), smem, st, a, qtiles));
    return PATS_OK;
}

int launch_gemm(GemmArgs a, cudaStream_t st, int dev, int sms) {
    // split Nout evenly into blocks of <= 256 outputs, each a multiple of 16
    a.nblocks = (a.Nout + 255) / 256;
    a.nb = (((a.Nout + a.nblocks - 1) / a.nblocks) + 15) & ~15;
    a.nblocks = (a.Nout + a.nb - 1) / a.nb;
    a.mblocks = (a.T + GEMM_M - 1) / GEMM_M;
    a.tmem_cols = 32;
    while (a.tmem_cols < a.nb) a.tmem_cols <<= 1;
    const bool split = g_precision.load(std::memory_order_relaxed) != 1;
    const size_t smem = sizeof(float) * (size_t)GKC * (GEMM_M + a.nb) * (split ? 2 : 1);
    static PerDeviceOnce configured;
    if (!configured.done(dev)) {
        PATS_CUDA_TRY(cudaFuncSetAttribute(gnn_gemm_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        PATS_CUDA_TRY(cudaFuncSetAttribute(gnn_gemm_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        configured.mark(dev);
    }
    // co-resident CTAs share the SM's 512 TMEM columns and its shared memory; a CTA that cannot allocate would wait for one that can
    int per_sm = 512 / a.tmem_cols;
    const int by_smem = (int)((224 * 1024) / (smem + 2048));
    if (per_sm > by_smem) per_sm = by_smem;
    if (per_sm > 2) per_sm = 2;  // __launch_bounds__(256, 2): the register file
    if (per_sm < 1) per_sm = 1;
    long long grid = (long long)a.mblocks * a.nblocks;
    if (grid > (long long)sms * per_sm) grid = (long long)sms * per_sm;
    if (split)
        PATS_CUDA_TRY(launch_chained(gnn_gemm_kernel<true>, dim3((unsigned)grid), dim3(GEMM_THREADS), smem, st, a));
    else
        PATS_CUDA_TRY(launch_chained(gnn_gemm_kernel<false>, dim3((unsigned)grid), dim3(GEMM_THREADS), smem, st, a));
    return PATS_OK;
}

}  // namespace
}  // namespace pats

using namespace pats;

PATS_API void pats_gnn_attention_variant(int v) { g_att_variant.store(v, std::memory_order_relaxed); }
PATS_API void pats_gnn_precision(int passes) { g_precision.store(passes == 1 ? 1 : 3, std::memory_order_relaxed); }

PATS_API long long pats_gnn_raw_floats(int layers, int D) { return (long long)(raw_layer_floats(D) * (size_t)layers); }
PATS_API long long pats_gnn_packed_floats(int layers, int D) { return (long long)((packed_layer_floats(D) + halves_layer_floats(D)) * (size_t)layers); }
PATS_API long long pats_gnn_workspace_floats(int chunk, int D, int N) { return (long long)28 * chunk * N * D + 32 * D; }
PATS_API void pats_gnn_gemm_variant(int v) { g_gemm_variant.store(v, std::memory_order_relaxed); }

namespace {
int pack_impl(const float *raw, int layers, int D, int heads, float bn_eps, int fold_bn, float *packed, void *stream) {
    if (layers <= 0 || D <= 0 || heads <= 0 || D % heads != 0) return invalid("gnn_pack: bad sizes layers=%d D=%d heads=%d", layers, D, heads);
    if (!raw || !packed) return invalid("gnn_pack: null pointer");
    gnn_pack_kernel<<<1184, 256, 0, as_stream(stream)>>>(raw, packed, layers, D, heads, bn_eps, fold_bn);
    PATS_LAUNCH_CHECK("gnn_pack_kernel");
    gnn_split_weights_kernel<<<1184, 256, 0, as_stream(stream)>>>(packed, layers, D);
    PATS_LAUNCH_CHECK("gnn_split_weights_kernel");
    return PATS_OK;
}
}  // namespace

PATS_API int pats_gnn_pack_f32(const float *raw, int layers, int D, int heads, float bn_eps, float *packed, void *stream) {
    return pack_impl(raw, layers, D, heads, bn_eps, 1, packed, stream);
}
PATS_API int pats_gnn_pack_train_f32(const float *raw, int layers, int D, int heads, float *packed, void *stream) {
    return pack_impl(raw, layers, D, heads, 0.f, 0, packed, stream);
}

namespace {
int launch_gemm_tma(const CUtensorMap &a1h, const CUtensorMap &a1l, const CUtensorMap &a2h, const CUtensorMap &a2l, const CUtensorMap &wh, const CUtensorMap &wl,
                    TmaGemmArgs a, const TmaShape &shape, int cluster, cudaStream_t st, int dev, int sms) {
    a.nb = shape.nb, a.nblocks = shape.nblocks, a.stages = shape.stages, a.cluster = cluster;
    a.mblocks = (a.T + GEMM_M - 1) / GEMM_M;
    static PerDeviceOnce configured;
    if (!configured.done(dev)) {
        PATS_CUDA_TRY(cudaFuncSetAttribute(gnn_gemm_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024));
        configured.mark(dev);
    }
    const long long groups = (long long)((a.mblocks + cluster - 1) / cluster) * a.nblocks;
    long long clusters = sms / cluster;
    if (clusters > groups) clusters = groups;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(clusters * cluster));
    cfg.blockDim = dim3(TMA_THREADS);
    cfg.dynamicSmemBytes = shape.smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[2];
    int n = 0;
    if (g_chain.load(std::memory_order_relaxed)) {
        attr[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[n].val.programmaticStreamSerializationAllowed = 1;
        ++n;
    }
    if (cluster > 1) {
        attr[n].id = cudaLaunchAttributeClusterDimension;
        attr[n].val.clusterDim.x = (unsigned)cluster, attr[n].val.clusterDim.y = 1, attr[n].val.clusterDim.z = 1;
        ++n;
    }
    cfg.attrs = attr;
    cfg.numAttrs = n;
    PATS_CUDA_TRY(cudaLaunchKernelEx(&cfg, gnn_gemm_tma_kernel, a1h, a1l, a2h, a2l, wh, wl, a));
    return PATS_OK;
}
}  // namespace

namespace {
struct TrainBn {
    const float *raw;  // the reference's parameters (gamma / beta of every layer's BatchNorm are read from here)
    float *running;    // [layers][2][2D] running_mean, running_var -- updated in place
    float momentum, eps;
};

int gnn_impl(const float *desc0, const float *desc1, int B, int D, int N, const float *packed, const unsigned char *cross, int layers, int heads, float *out0,
             float *out1, float *workspace, long long workspace_floats, void *stream, const TrainBn *train) {
    if (B < 0 || D <= 0 || N <= 0 || layers <= 0 || heads <= 0) return invalid("attentional_gnn: bad sizes B=%d D=%d N=%d layers=%d heads=%d", B, D, N, layers, heads);
    if (B == 0) return PATS_OK;
    if (!desc0 || !desc1 || !packed || !cross || !out0 || !out1 || !workspace) return invalid("attentional_gnn: null pointer");
    if (D % heads != 0 || D % 8 != 0 || (D / heads) % 2 != 0)
        return invalid("attentional_gnn: D = %d, heads = %d: D must be a multiple of 8 and of heads, the head dimension even", D, heads);
    const int dim = D / heads;
    const int NJ = (N + 31) / 32, DI = (dim + 31) / 32;
    if (DI > 4) return invalid("attentional_gnn: head dimension %d exceeds the 128 this build has attention kernels for", dim);
    const long long per_problem = (long long)28 * N * D;
    const long long ws_problems = (workspace_floats - 32 * D) / per_problem;
    int chunk = (int)(ws_problems < B ? ws_problems : B);
    if (train && chunk < B)
        return invalid("attentional_gnn (train): batch statistics need the whole batch in one chunk: workspace of %lld floats, %lld needed", workspace_floats,
                       per_problem * B + 32 * D);
    if (chunk > 16384) chunk = 16384;  // 2 * chunk is a grid z extent
    if (chunk < 1) return invalid("attentional_gnn: workspace of %lld floats holds no problem (%lld floats each)", workspace_floats, per_problem);
    const int dev = current_device();
    if (dev < 0) return PATS_E_CUDA;
    const int sms = sm_count() > 0 ? sm_count() : 148;
    cudaStream_t st = as_stream(stream);
    const size_t per = packed_layer_floats(D), per2 = halves_layer_floats(D);
    const size_t DD = (size_t)D * D;
    const int gv = g_gemm_variant.load(std::memory_order_relaxed);
    const bool tma = gv != 1 || train != nullptr;
    const int cluster = gv == 0 ? 2 : 1;  // default: CTA pairs with the weight tile multicast (level 2, 300 windows: GEMMs 22.2 -> 20.5 ms)
    const bool split = g_precision.load(std::memory_order_relaxed) != 1;
    // weights as TF32 halves, one 3-D tensor map per matrix over all layers (box height = the output block of the call's shape)
    CUtensorMap m_qkv_h, m_qkv_l, m_w1_h, m_w1_l, m_w2_h, m_w2_l;
    TmaShape s_qkv = {}, s_w1 = {}, s_w2 = {};
    int maps_for_mblocks = -1;
    for (int b0 = 0; b0 < B; b0 += chunk) {
        const int Bc = B - b0 < chunk ? B - b0 : chunk;
        const int T = 2 * Bc * N;
        const size_t TD = (size_t)T * D;
        // stats (32 D) | X | Xh | Xl | QKV (3) | O or Oh, Ol (2) | Y or Yh, Yl (4) | Z (4, train only)
        double *stats = reinterpret_cast<double *>(workspace);
        float *X = workspace + 32 * D, *Xh = X + TD, *Xl = Xh + TD, *QKV = Xl + TD, *O = QKV + 3 * TD, *Oh = O, *Ol = O + TD, *Y = Ol + TD, *Yh = Y, *Yl = Y + 2 * TD, *Z = Yl + 2 * TD;
        CUtensorMap m_xh, m_xl, m_oh, m_ol, m_yh, m_yl;
        if (tma) {
            const int mblocks = (T + GEMM_M - 1) / GEMM_M;
            if (mblocks != maps_for_mblocks) {  // first chunk, and a shorter last one
                const float *halves = packed + (size_t)layers * per;
                s_qkv = tma_shape(3 * D, split, mblocks, sms), s_w1 = tma_shape(2 * D, split, mblocks, sms), s_w2 = tma_shape(D, split, mblocks, sms);
                int rcw = make_map_3d(&m_qkv_h, halves, layers, 3 * D, D, per2, s_qkv.nb / cluster);
                if (!rcw) rcw = make_map_3d(&m_qkv_l, halves + 3 * DD, layers, 3 * D, D, per2, s_qkv.nb / cluster);
                if (!rcw) rcw = make_map_3d(&m_w1_h, halves + 6 * DD, layers, 2 * D, 2 * D, per2, s_w1.nb / cluster);
                if (!rcw) rcw = make_map_3d(&m_w1_l, halves + 10 * DD, layers, 2 * D, 2 * D, per2, s_w1.nb / cluster);
                if (!rcw) rcw = make_map_3d(&m_w2_h, halves + 14 * DD, layers, D, 2 * D, per2, s_w2.nb / cluster);
                if (!rcw) rcw = make_map_3d(&m_w2_l, halves + 16 * DD, layers, D, 2 * D, per2, s_w2.nb / cluster);
                if (rcw) return rcw;
                maps_for_mblocks = mblocks;
            }
            int rc = make_map_2d(&m_xh, Xh, T, D, D, GEMM_M);
            if (!rc) rc = make_map_2d(&m_xl, Xl, T, D, D, GEMM_M);
            if (!rc) rc = make_map_2d(&m_oh, Oh, T, D, D, GEMM_M);
            if (!rc) rc = make_map_2d(&m_ol, Ol, T, D, D, GEMM_M);
            if (!rc) rc = make_map_2d(&m_yh, Yh, T, 2 * D, 2 * D, GEMM_M);
            if (!rc) rc = make_map_2d(&m_yl, Yl, T, 2 * D, 2 * D, GEMM_M);
            if (rc) return rc;
        }
        TransArgs t;
        t.d0 = desc0, t.d1 = desc1, t.o0 = out0, t.o1 = out1, t.X = X, t.Xh = tma ? Xh : nullptr, t.Xl = tma ? Xl : nullptr;
        t.b0 = b0, t.Bc = Bc, t.D = D, t.N = N, t.to_tokens = 1;
        const dim3 tgrid((unsigned)((N + 31) / 32), (unsigned)((D + 31) / 32), (unsigned)(2 * Bc));
        PATS_CUDA_TRY(launch_chained(gnn_transpose_kernel, tgrid, dim3(256), 0, st, t));
        for (int l = 0; l < layers; ++l) {
            const float *w = packed + (size_t)l * per;
            const float *Wqkv = w, *bqkv = Wqkv + 3 * DD, *W1f = bqkv + 3 * D, *b1f = W1f + 4 * DD, *W2 = b1f + 2 * D, *b2 = W2 + 2 * DD;
            int rc;
            GemmArgs g = {};
            TmaGemmArgs ta = {};
            ta.T = T, ta.layer = l, ta.split = split ? 1 : 0;
            if (tma) {
                ta.bias = bqkv, ta.out = QKV, ta.out_h = ta.out_l = nullptr, ta.ldo = 3 * D, ta.Nout = 3 * D, ta.K1 = D, ta.K2 = 0, ta.relu = 0, ta.accumulate = 0;
                rc = launch_gemm_tma(m_xh, m_xl, m_xh, m_xl, m_qkv_h, m_qkv_l, ta, s_qkv, cluster, st, dev, sms);
            } else {
                g.A1 = X, g.lda1 = D, g.K1 = D, g.A2 = nullptr, g.lda2 = 0, g.K2 = 0, g.W = Wqkv, g.ldw = D, g.bias = bqkv, g.out = QKV, g.ldo = 3 * D;
                g.T = T, g.Nout = 3 * D, g.relu = 0, g.accumulate = 0;
                rc = launch_gemm(g, st, dev, sms);
            }
            if (rc) return rc;
            AttArgs at;
            at.qkv = QKV, at.o = O, at.oh = tma ? Oh : nullptr, at.ol = tma ? Ol : nullptr;
            at.Bc = Bc, at.N = N, at.D = D, at.heads = heads, at.dim = dim, at.cross = cross[l] ? 1 : 0;
            at.c = 1.4426950408889634f / sqrtf((float)dim);
            const int av = g_att_variant.load(std::memory_order_relaxed);  // 0: packed-FP32 generation; 1: first generation (same sums, same order)
            if (NJ <= 3 && DI == 1)
                // 65 rows = 9 blocks of 8: one pass per warp.  Measured (2800 points, 560 per launch): 9 warps 158 us, 5 warps 165 us, 4 warps 188 us, 3 warps 242 us
                rc = av == 0 ? launch_attention2<3, 0, 9>(at, st, dev) : launch_attention<3, 1, false, 8, 4>(at, st, dev);
            else if (NJ <= 5 && dim >= 64 && dim <= 68)
                // first generation, measured (tools/gnn_kernels.py, 89 windows per launch): 4 rows x 20 warps 206 us, 8 rows x 10 warps 232 us;
                // packed generation, 300 windows x 18 launches (tools/ab_attention_rows.py): 8 x 20 10.7 ms, 12 x 13 (av 3) 12.2 ms, 16 x 10 (av 2) 14.6 ms
                rc = av == 0 ? launch_attention2<5, 1, 20>(at, st, dev) : av == 2 ? launch_attention2<5, 1, 10, 16>(at, st, dev)
                             : av == 3 ? launch_attention2<5, 1, 13, 12>(at, st, dev)
                             : launch_attention<5, 3, true, 4, 20>(at, st, dev);
            else if (NJ <= 5 && DI <= 3)
                rc = launch_attention<5, 3, false, 4, 8>(at, st, dev);
            else
                rc = launch_attention_flash<5, 4, 4, 8, 2>(at, st, dev);
            if (rc) return rc;
            if (train) {
                // Z = [X | O] W1f^T + b1f (BatchNorm not folded), batch statistics per side, normalise + ReLU -> the halves of Y
                ta.bias = b1f, ta.out = Z, ta.out_h = ta.out_l = nullptr, ta.ldo = 2 * D, ta.Nout = 2 * D, ta.K1 = D, ta.K2 = D, ta.relu = 0, ta.accumulate = 0;
                rc = launch_gemm_tma(m_xh, m_xl, m_oh, m_ol, m_w1_h, m_w1_l, ta, s_w1, cluster, st, dev, sms);
                if (rc) return rc;
                PATS_CUDA_TRY(cudaMemsetAsync(stats, 0, sizeof(double) * 8 * D, st));
                BnArgs bn;
                const float *rl = train->raw + (size_t)l * raw_layer_floats(D);
                bn.Z = Z, bn.stats = stats, bn.gamma = rl + 4 * (DD + D) + 4 * DD + 2 * D, bn.beta = bn.gamma + 2 * D;
                bn.running = train->running + (size_t)l * 4 * D, bn.Yh = Yh, bn.Yl = Yl, bn.Th = T / 2, bn.C = 2 * D, bn.eps = train->eps, bn.momentum = train->momentum;
                const dim3 bgrid((unsigned)((T / 2 + BN_ROWS - 1) / BN_ROWS), 2u);
                gnn_bn_stats_kernel<<<bgrid, 256, 0, st>>>(bn);
                PATS_LAUNCH_CHECK("gnn_bn_stats_kernel");
                PATS_CUDA_TRY(launch_chained(gnn_bn_apply_kernel, bgrid, dim3(256), 0, st, bn));
            } else if (tma) {
                ta.bias = b1f, ta.out = nullptr, ta.out_h = Yh, ta.out_l = Yl, ta.ldo = 2 * D, ta.Nout = 2 * D, ta.K1 = D, ta.K2 = D, ta.relu = 1, ta.accumulate = 0;
                rc = launch_gemm_tma(m_xh, m_xl, m_oh, m_ol, m_w1_h, m_w1_l, ta, s_w1, cluster, st, dev, sms);
            } else {
                g.A1 = X, g.lda1 = D, g.K1 = D, g.A2 = O, g.lda2 = D, g.K2 = D, g.W = W1f, g.ldw = 2 * D, g.bias = b1f, g.out = Y, g.ldo = 2 * D;
                g.Nout = 2 * D, g.relu = 1, g.accumulate = 0;
                rc = launch_gemm(g, st, dev, sms);
            }
            if (rc) return rc;
            if (tma) {
                ta.bias = b2, ta.out = X, ta.out_h = Xh, ta.out_l = Xl, ta.ldo = D, ta.Nout = D, ta.K1 = 2 * D, ta.K2 = 0, ta.relu = 0, ta.accumulate = 1;
                rc = launch_gemm_tma(m_yh, m_yl, m_yh, m_yl, m_w2_h, m_w2_l, ta, s_w2, cluster, st, dev, sms);
            } else {
                g.A1 = Y, g.lda1 = 2 * D, g.K1 = 2 * D, g.A2 = nullptr, g.lda2 = 0, g.K2 = 0, g.W = W2, g.ldw = 2 * D, g.bias = b2, g.out = X, g.ldo = D;
                g.Nout = D, g.relu = 0, g.accumulate = 1;
                rc = launch_gemm(g, st, dev, sms);
            }
            if (rc) return rc;
        }
        t.to_tokens = 0;
        PATS_CUDA_TRY(launch_chained(gnn_transpose_kernel, tgrid, dim3(256), 0, st, t));
    }
    return PATS_OK;
}
}  // namespace

PATS_API int pats_attentional_gnn_f32(const float *desc0, const float *desc1, int B, int D, int N, const float *packed, const unsigned char *cross,
                                      int layers, int heads, float *out0, float *out1, float *workspace, long long workspace_floats, void *stream) {
    return gnn_impl(desc0, desc1, B, D, N, packed, cross, layers, heads, out0, out1, workspace, workspace_floats, stream, nullptr);
}

PATS_API int pats_attentional_gnn_train_f32(const float *desc0, const float *desc1, int B, int D, int N, const float *packed, const float *raw, float *running,
                                            float momentum, float bn_eps, const unsigned char *cross, int layers, int heads, float *out0, float *out1,
                                            float *workspace, long long workspace_floats, void *stream) {
    if (!raw || !running) return invalid("attentional_gnn (train): null pointer");
    if (D % 4 != 0) return invalid("attentional_gnn (train): D = %d", D);
    TrainBn t = {raw, running, momentum, bn_eps};
    return gnn_impl(desc0, desc1, B, D, N, packed, cross, layers, heads, out0, out1, workspace, workspace_floats, stream, &t);
}
