// tcgen05.cuh — the Blackwell (sm_100a) tensor-core primitives used by the Alpha0.5 policy net:
// TMEM allocation, shared-memory matrix descriptors, tcgen05.mma (kind::f16, one issuing thread,
// fp32 accumulator in TMEM), tcgen05.commit onto an mbarrier, tcgen05.ld for the epilogue.
// Inline PTX only.  Descriptor bit layouts follow the PTX ISA "tcgen05 matrix descriptor" /
// "instruction descriptor" tables.
#pragma once
#include <cstdint>
#include "tma.cuh"

namespace nimmt {

// --- TMEM allocation (one warp, warp-collective) ------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_slot, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_slot)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
// Several allocations by one CTA (each a power of two): allocate them all, then give up the permit once.
__device__ __forceinline__ void tmem_alloc_keep_permit(uint32_t* smem_slot, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_slot)), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_relinquish_permit() { asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// --- shared-memory operand descriptor, K-major, no swizzle --------------------------------------
// Canonical layout (in 16-byte units): ((8, n), 2) : ((1, SBO), LBO) — "core matrices" of 8 rows x
// 16 bytes stored contiguously (128 B); the two core matrices that make up K = 16 bf16 of one MMA are
// LBO bytes apart, consecutive 8-row groups SBO bytes apart.
__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((smem_addr & 0x3FFFFu) >> 4)         // bits  0..13 start address >> 4
           | ((uint64_t)(lbo_bytes >> 4) << 16)            // bits 16..29 leading-dimension byte offset >> 4
           | ((uint64_t)(sbo_bytes >> 4) << 32)            // bits 32..45 stride-dimension byte offset >> 4
           | (1ull << 46);                                 // bits 46..47 descriptor version 1 (sm_100); layout_type 0 = no swizzle
}

// Byte offset of element (row, k) of a K-major bf16 operand with `kchunks` 8-element chunks per row.
__host__ __device__ constexpr uint32_t canon_off(uint32_t row, uint32_t k, uint32_t kchunks) {
    return (row >> 3) * (kchunks * 128u) + (k >> 3) * 128u + (row & 7u) * 16u + (k & 7u) * 2u;
}

// --- instruction descriptor: D fp32 = A bf16 (K-major) x B bf16 (K-major), M x N ----------------
__host__ __device__ constexpr uint32_t umma_idesc_bf16(uint32_t M, uint32_t N) {
    return (1u << 4)            // bits 4..5   D format: F32
           | (1u << 7)          // bits 7..9   A format: BF16
           | (1u << 10)         // bits 10..12 B format: BF16
           | ((N >> 3) << 17)   // bits 17..22 N >> 3
           | ((M >> 4) << 24);  // bits 24..28 M >> 4      (a_major = b_major = 0: K-major)
}

// D[tmem] (+)= A[smem] * B[smem]; issued by ONE thread on behalf of the CTA.
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}

// D[tmem] (+)= A[tmem] * B[smem]: the A operand (128 rows x 16 bf16) is read from tensor memory — row m in lane m, elements
// 2 j, 2 j + 1 of the row packed into the 32-bit column tmem_a + j (8 columns per K = 16 step) — which is where an epilogue
// thread can leave the previous layer's activations with tcgen05.st, without a trip through shared memory.
__device__ __forceinline__ void umma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n"
        "}\n" ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}

// Arrive on `bar` once every MMA issued so far by this thread has completed (implies
// tcgen05.fence::before_thread_sync).
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// 16 consecutive fp32 columns of this thread's TMEM lane (warp w of the CTA owns lanes 32 (w % 4) ..).
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
          "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&v)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "r"(taddr)
                 : "memory");
}
__device__ __forceinline__ void tmem_ld4(uint32_t taddr, uint32_t (&v)[4]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]) : "r"(taddr) : "memory");
}
// Registers -> 8 / 4 consecutive 32-bit columns of this thread's TMEM lane.
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&v)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]),
                 "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
                 : "memory");
}
__device__ __forceinline__ void tmem_st4(uint32_t taddr, const uint32_t (&v)[4]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]) : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// Wait for an MMA-completion phase.  Product build: mbarrier.try_wait in a loop (the instruction itself suspends the thread
// for a hardware-bounded time, so this neither burns issue slots nor can it kill the context when a profiler, a sanitizer or
// time-slicing stretches a legitimate wait).  -DNIMMT_DEBUG_TRAP (bring-up): give up after 2^24 polls and trap instead of
// hanging the GPU.
__device__ __forceinline__ void mbar_wait_mma(uint64_t* bar, uint32_t parity) {
#ifdef NIMMT_DEBUG_TRAP
    for (uint32_t spin = 0; spin < (1u << 24); ++spin) {
#else
    for (;;) {
#endif
        uint32_t done;
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
        if (done) return;
    }
#ifdef NIMMT_DEBUG_TRAP
    __trap();
#endif
}

// mbar_wait_mma on a 32-bit shared-memory address taken once, outside the loop.
__device__ __forceinline__ void mbar_wait_mma_a(uint32_t bar, uint32_t parity) {
#ifdef NIMMT_DEBUG_TRAP
    for (uint32_t spin = 0; spin < (1u << 24); ++spin) {
#else
    for (;;) {
#endif
        uint32_t done;
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
        if (done) return;
    }
#ifdef NIMMT_DEBUG_TRAP
    __trap();
#endif
}

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) { mbar_arrive_a(smem_u32(bar)); }
// tcgen05.commit on a 32-bit shared-memory address taken once, outside the loop.
__device__ __forceinline__ void umma_commit_a(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

}  // namespace nimmt
