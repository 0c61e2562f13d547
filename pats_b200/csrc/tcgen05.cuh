// tcgen05 / TMEM / mbarrier building blocks shared by the tensor-core kernels (correlation.cu, gnn.cu).  sm_100a only.
// Operand tiles live in shared memory in the canonical K-major, no-swizzle UMMA layout: a tile is a grid of 8-row x 16-byte "core
// matrices" (8 x 4 floats); element (row, k) of a tile with KC4 core matrices along K sits at float index
//     ((row / 8) * KC4 + k / 4) * 32 + (row % 8) * 4 + k % 4
// i.e. leading byte offset (between core matrices along K) = 128 B, stride byte offset (between 8-row groups) = KC4 * 128 B.
#pragma once

#include "common.cuh"

#ifdef __CUDACC__
namespace pats {
namespace tc {

__device__ __forceinline__ unsigned smem_addr(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }

// shared-memory matrix descriptor, K-major, SWIZZLE_NONE; sbo_bytes = distance between 8-row groups
__device__ __forceinline__ unsigned long long umma_desc(unsigned saddr, unsigned sbo_bytes) {
    unsigned long long d = 0;
    d |= (unsigned long long)((saddr & 0x3FFFFu) >> 4);   // start address, bits [0,14)
    d |= (unsigned long long)(128u >> 4) << 16;           // leading byte offset, bits [16,30)
    d |= (unsigned long long)(sbo_bytes >> 4) << 32;      // stride byte offset, bits [32,46)
    d |= 1ull << 46;                                      // descriptor version (Blackwell)
    return d;                                             // base offset 0, layout type SWIZZLE_NONE
}
// instruction descriptor: kind::tf32, FP32 accumulate, both operands K-major, M = 128, N = n (a multiple of 16, <= 256)
__device__ __forceinline__ unsigned umma_idesc(int n) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((unsigned)(n >> 3) << 17) | ((128u >> 4) << 24);
}
// D[tmem] (+)= A[smem] * B[smem]; issued by ONE thread
__device__ __forceinline__ void umma_tf32(unsigned d_tmem, unsigned long long a, unsigned long long b, unsigned idesc, unsigned accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(d_tmem),
        "l"(a), "l"(b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// arrives on the mbarrier when every MMA issued so far by this thread has finished reading shared memory and writing TMEM
__device__ __forceinline__ void umma_commit(unsigned mb) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(mb) : "memory");
}
__device__ __forceinline__ void mbar_init1(unsigned mb) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mb) : "memory"); }
__device__ __forceinline__ void mbar_wait_parity(unsigned mb, unsigned parity) {
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
__device__ __forceinline__ void fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// generic-proxy shared-memory stores -> visible to the tensor core's (async proxy) reads
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// TMEM allocation by one whole warp (cols = a power of two >= 32); the base address lands in *slot (shared memory)
__device__ __forceinline__ void tmem_alloc(unsigned *slot, unsigned cols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_addr(slot)), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(unsigned tmem, unsigned cols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(cols) : "memory");
}
// 16 consecutive accumulator columns of this lane's TMEM lane (= accumulator row); warp-collective; returns after the data arrived
__device__ __forceinline__ void tmem_ld16(unsigned taddr, unsigned (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
          "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// round to TF32 (nearest).  The tensor core would TRUNCATE the low 13 bits of a 32-bit container, and a truncation error has one
// sign -- it adds up linearly over K.
__device__ __forceinline__ float tf32_rn(float x) {
    unsigned r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}

}  // namespace tc
}  // namespace pats
#endif  // __CUDACC__
