// N2 -- the descriptor correlation that feeds every Sinkhorn solve, on the 5th-generation tensor cores (tcgen05 + TMEM).
// Replaces, from zju3dv/pats:
//   models/first_layer.py:110-114    scores = einsum('bdn,bdm->bnm', mdesc0, mdesc1) / 448**.5 ;  0.1 * scores
//   models/second_layer.py:100-104   ... / 264**.5 ;  0.1 * scores          (b = P windows, n = m = 145)
//   models/third_layer.py:156-158    ... / 128**.5 ;  0.1 * scores          (b = K points,  n = m = 65)
// i.e. one batched GEMM and two full elementwise passes over the [b,n,m] score tensor.  Here: one kernel that writes
// Z = scale * d0^T d1 (scale = 0.1 / sqrt(d)) straight into the contiguous [b,n,m] layout the Sinkhorn kernels stage with one
// bulk copy per problem.
//
// Arithmetic.  The reference multiplies in FP32 (cuBLAS SGEMM; torch's allow_tf32 is off by default) and the plans must agree
// to 1e-4 with identical argmax, so plain TF32 (10-bit mantissa: ~5e-4 relative per product) is not enough.  Every operand is
// split x = hi + lo with hi = x rounded to TF32 and lo = x - hi (exact in FP32) rounded to TF32, and three MMAs per K step accumulate hi*hi + hi*lo + lo*hi in the FP32 accumulator ("3xTF32"; the
// dropped lo*lo term is ~2^-22 relative).  Measured against the FP32 einsum: tests/test_gpu_correlation.py.
//
// Structure (one CTA of 256 threads per (problem, 128-row block)):
//   * operands are staged by all threads: global [d, n] rows (contiguous along n) -> registers -> split -> shared memory in the
//     canonical K-major no-swizzle UMMA layout (core matrix = 8 rows x 16 B; the transposition happens in the store index, and
//     lane = (k % 4) * 8 + (row % 8) makes the 32 stores of a warp hit 32 different banks).  TMA cannot do this copy: rows are
//     65 / 145 / 300 floats, i.e. not 16-byte aligned.
//   * ONE thread issues tcgen05.mma.cta_group::1.kind::tf32 (M = 128, N = a multiple of 16 up to 160, K = 8) per K step,
//     accumulator in TMEM, and tcgen05.commit arrives on an mbarrier that the CTA waits on before restaging the buffers;
//   * the eight warps read the 32 TMEM lanes of their quadrant with tcgen05.ld (32x32b.x16; warps w and w + 4 alternate over the
//     16-column groups), scale, and store the valid rows / columns.
//   SASS: UTCMMA (the MMA), LDTM (TMEM load), UTCBAR (commit), STS / LDG for the staging.
#include "common.cuh"
#include "tcgen05.cuh"

namespace pats {
namespace {

constexpr int KC = 32;            // K extent staged per chunk
constexpr int KC4 = KC / 4;       // core matrices along K per chunk
constexpr int CORR_THREADS = 256;  // 8 warps stage; warps w and w + 4 share a TMEM lane quadrant and split the accumulator columns

using namespace tc;  // csrc/tcgen05.cuh: descriptors, MMA issue, commit / mbarrier, TMEM allocation and loads, TF32 rounding

// K-major, no swizzle (tcgen05.cuh): LBO (K direction) = 128 B, SBO (row direction) = KC4 * 128 B
__device__ __forceinline__ unsigned long long umma_desc(unsigned saddr) { return tc::umma_desc(saddr, KC4 * 128u); }

struct CorrArgs {
    const float *d0, *d1;  // [b,d,n], [b,d,m]
    float *out;            // [b,n,m]
    int b, d, n, m;
    float scale;
    int mblocks;           // ceil(n / 128)
    int npad;              // m rounded up to 16
    int tile_n, ntiles;    // N extent per MMA (multiple of 16, <= 160), tiles per accumulator row block
    int tmem_cols;         // power of two >= max(32, npad)
};

// stage rows [r0, r0 + rows) x k in [k0, k0 + kn) of src[d, ld] into the hi / lo tiles (rows beyond `valid` and k beyond kn are zero).
// A warp moves one 8-row x 4-k core matrix per step; UN steps are in flight together (the loop is latency-bound otherwise: one
// dependent LDG -> split -> 2 STS chain per thread at a time kept a CTA at ~60 us per 145-row problem).
template <int UN>
__device__ __forceinline__ void stage_tile(const float *__restrict__ src, int ld, int valid, int r0, int rows, int k0, int kn, float *hi,
                                           float *lo, int warp, int lane, int nwarps) {
    const int kq = lane >> 3, rr = lane & 7;  // lane = (k % 4) * 8 + row % 8
    const int groups = (rows + 7) >> 3, nblk = groups * KC4;
    for (int blk0 = warp; blk0 < nblk; blk0 += nwarps * UN) {
        float x[UN];
#pragma unroll
        for (int u = 0; u < UN; ++u) {
            const int blk = blk0 + u * nwarps;
            const int g8 = blk / KC4, k4 = blk - g8 * KC4;
            const int row = r0 + g8 * 8 + rr, k = k4 * 4 + kq;
            x[u] = (blk < nblk && row < valid && k < kn) ? __ldg(src + (size_t)(k0 + k) * ld + row) : 0.f;
        }
#pragma unroll
        for (int u = 0; u < UN; ++u) {
            const int blk = blk0 + u * nwarps;
            if (blk < nblk) {
                const int g8 = blk / KC4, k4 = blk - g8 * KC4;
                // hi = x rounded to TF32 (nearest), lo = the exact remainder, rounded to TF32 as well: the tensor core would TRUNCATE
                // the low 13 bits of a 32-bit container, and a truncation error has one sign -- it adds up linearly over K
                const float h = tf32_rn(x[u]);
                const int idx = (g8 * KC4 + k4) * 32 + rr * 4 + kq;
                hi[idx] = h;
                lo[idx] = tf32_rn(x[u] - h);
            }
        }
    }
}

__global__ void __launch_bounds__(CORR_THREADS) correlation_tcgen05_kernel(CorrArgs a) {
    extern __shared__ __align__(128) float smem[];
    __shared__ __align__(8) unsigned long long s_bar;
    __shared__ unsigned s_tmem;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    float *a_hi = smem, *a_lo = a_hi + 128 * KC;        // [16][KC4][8][4]
    float *b_hi = a_lo + 128 * KC, *b_lo = b_hi + a.npad * KC;
    const unsigned mb = smem_addr(&s_bar);
    if (warp == 0) {
        tmem_alloc(&s_tmem, (unsigned)a.tmem_cols);
    }
    if (tid == 0) {
        mbar_init1(mb);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const unsigned tmem = s_tmem;
    unsigned phase = 0;

    for (int work = blockIdx.x; work < a.b * a.mblocks; work += gridDim.x) {
        const int p = work / a.mblocks, mblk = work - p * a.mblocks;
        const float *d0 = a.d0 + (size_t)p * a.d * a.n, *d1 = a.d1 + (size_t)p * a.d * a.m;
        const int r0 = mblk * 128;
        unsigned first = 1;
        for (int k0 = 0; k0 < a.d; k0 += KC) {
            const int kn = min(KC, a.d - k0);
            // only the row groups that hold real rows are written; what the other accumulator rows / columns see is never stored
            stage_tile<16>(d0, a.n, a.n, r0, min(128, (a.n - r0 + 7) & ~7), k0, kn, a_hi, a_lo, warp, lane, CORR_THREADS / 32);
            stage_tile<16>(d1, a.m, a.m, 0, (a.m + 7) & ~7, k0, kn, b_hi, b_lo, warp, lane, CORR_THREADS / 32);
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy stores -> visible to the tensor core's reads
            __syncthreads();
            if (tid == 0) {
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const int steps = (kn + 7) >> 3;
                for (int t = 0; t < a.ntiles; ++t) {
                    const int c0 = t * a.tile_n, cn = min(a.tile_n, a.npad - c0);
                    const unsigned idesc = umma_idesc(cn);
                    const unsigned dcol = tmem + (unsigned)c0;
                    // B rows c0.. of this tile: (c0 / 8) row groups further on
                    const unsigned boff = (unsigned)((c0 >> 3) * KC4 * 128);
                    for (int s = 0; s < steps; ++s) {
                        const unsigned koff = (unsigned)s * 256u;  // two core matrices (K = 8) per step
                        const unsigned long long ah = umma_desc(smem_addr(a_hi) + koff), al = umma_desc(smem_addr(a_lo) + koff);
                        const unsigned long long bh = umma_desc(smem_addr(b_hi) + boff + koff), bl = umma_desc(smem_addr(b_lo) + boff + koff);
                        umma_tf32(dcol, ah, bh, idesc, (first && s == 0) ? 0u : 1u);
                        umma_tf32(dcol, ah, bl, idesc, 1u);
                        umma_tf32(dcol, al, bh, idesc, 1u);
                    }
                }
                // arrives on the mbarrier when every MMA issued so far has finished reading shared memory and writing TMEM
                umma_commit(mb);
            }
            first = 0;
            mbar_wait_parity(mb, phase);
            phase ^= 1u;
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        }
        // ---- epilogue: warp w owns TMEM lanes (= accumulator rows) 32w .. 32w+31 --------------------------------------
        const int quad = warp & 3;  // a warp may only touch the TMEM lanes of its quadrant (warp id % 4)
        const int row = r0 + quad * 32 + lane;
        float *orow = a.out + ((size_t)p * a.n + row) * a.m;
        for (int c0 = (warp >> 2) * 16; c0 < a.npad; c0 += 32) {
            unsigned v[16];
            const unsigned taddr = tmem + ((unsigned)(quad * 32) << 16) + (unsigned)c0;
            tmem_ld16(taddr, v);
            if (row < a.n) {
#pragma unroll
                for (int j = 0; j < 16; ++j)
                    if (c0 + j < a.m) orow[c0 + j] = __uint_as_float(v[j]) * a.scale;
            }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();  // every warp has drained its accumulator rows before the next block's first MMA overwrites them
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    }
    if (warp == 0) tmem_dealloc(tmem, (unsigned)a.tmem_cols);
}

}  // namespace
}  // namespace pats

using namespace pats;

PATS_API int pats_correlation_f32(const float *d0, const float *d1, int b, int d, int n, int m, float scale, float *out, void *stream) {
    if (b < 0 || d <= 0 || n <= 0 || m <= 0) return invalid("correlation: bad sizes b=%d d=%d n=%d m=%d", b, d, n, m);
    if (b == 0) return PATS_OK;
    if (!d0 || !d1 || !out) return invalid("correlation: null pointer");
    if (d % 8 != 0) return invalid("correlation: d = %d is not a multiple of 8 (the K extent of one TF32 MMA)", d);
    CorrArgs a;
    a.d0 = d0, a.d1 = d1, a.out = out, a.b = b, a.d = d, a.n = n, a.m = m, a.scale = scale;
    a.mblocks = (n + 127) / 128;
    a.npad = (m + 15) & ~15;
    if (a.npad > 512) return invalid("correlation: m = %d exceeds the 512 accumulator columns of tensor memory", m);
    a.ntiles = (a.npad + 159) / 160;
    a.tile_n = (((a.npad + a.ntiles - 1) / a.ntiles) + 15) & ~15;
    a.tmem_cols = 32;
    while (a.tmem_cols < a.npad) a.tmem_cols <<= 1;
    const size_t smem = sizeof(float) * (size_t)KC * 2 * (128 + a.npad);
    const int dev = current_device();
    if (dev < 0) return PATS_E_CUDA;
    static PerDeviceOnce configured;
    if (!configured.done(dev)) {
        PATS_CUDA_TRY(cudaFuncSetAttribute(correlation_tcgen05_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        configured.mark(dev);
    }
    // co-resident CTAs share the SM's 512 TMEM columns and its shared memory; a CTA that cannot allocate would wait for one that can
    int per_sm = 512 / a.tmem_cols;
    const int by_smem = (int)((200 * 1024) / (smem + 1024));
    if (per_sm > by_smem) per_sm = by_smem;
    if (per_sm < 1) per_sm = 1;
    const int sms = sm_count() > 0 ? sm_count() : 148;
    long long grid = (long long)b * a.mblocks;
    if (grid > (long long)sms * per_sm) grid = (long long)sms * per_sm;
    correlation_tcgen05_kernel<<<(unsigned)grid, CORR_THREADS, smem, as_stream(stream)>>>(a);
    PATS_LAUNCH_CHECK("correlation_tcgen05_kernel");
    return PATS_OK;
}
