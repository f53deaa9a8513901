// policy_tile.cuh — one 128-row tile of the Alpha0.5 policy net on tcgen05: the pieces shared by the batched leaf-evaluation
// kernel (policy_kernels.cu, which arranges them into a warp-specialised pipeline), the masked-policy kernel and the
// policy-rollout kernel (which calls mlp_tile, the whole tile with one group of 128 threads).  See policy_kernels.cu for the overview.
//
// All three kernels are bound by what happens around the MMAs — the latency of one tile's chain, with one tile in flight in the
// search kernel and three in the batched one — so the tile is built to leave the threads little:
//   * biases ride in the GEMMs: features 48, 49 of every row are the constant 1 and carry b1 split into two
//     bf16 terms (hi + lo, ~16 mantissa bits); layer 1's units 100, 101 are the constant 1 and carry b2 the
//     same way.  Epilogue 1 is cvt.relu.bf16x2 + store; epilogue 2 is ONE fma per column, because the linear half
//     of the ReLU head is computed by the GEMM as well (epilogue2_chunks).
//   * layer 2's A operand never touches shared memory: epilogue 1 writes the bf16 activations back into TENSOR memory
//     (tcgen05.st, a thread's row = its TMEM lane, two units per 32-bit column) on top of the accumulator columns it has
//     just read, and layer 2 runs with A from TMEM (umma_bf16_ts).  Per tile that removes 28 KB of shared-memory stores, the
//     tensor core's 28 KB read of them and the async-proxy fence between the two.
//   * feature rows are assembled from bf16 data that is already laid out in 16-byte chunks (six vector
//     loads and stores per row) instead of 48 scalar conversions.
#pragma once
#include <cstring>

#include <cuda_bf16.h>

#include "abi_common.cuh"
#include "tcgen05.cuh"

namespace nimmt {

constexpr int kIn = 48, kInPad = 64, kHid = 100, kHidPad = 112, kObs = 47;
constexpr int kBiasCol = 48;     // features 48, 49 = 1.0 (b1 hi, lo); 50..63 = 0
constexpr int kOneUnit = 100;    // hidden-1 units 100, 101 = 1.0 (b2 hi, lo); 102..111 = 0
constexpr int kTileRows = 128, kSlots = 10, kDecPerTile = 12, kDecPerWarp = 3;
constexpr int kInChunks = kInPad / 8, kHidChunks = kHidPad / 8, kFeatChunks = kIn / 8;
constexpr uint32_t kW1Bytes = (kHidPad / 8) * kInChunks * 128;    // 14336
constexpr uint32_t kW2Bytes = (kHidPad / 8) * kHidChunks * 128;   // 25088
constexpr uint32_t kOffW1 = 0, kOffW2 = kOffW1 + kW1Bytes, kOffW3 = kOffW2 + kW2Bytes, kOffB3 = kOffW3 + kHidPad * 4;
constexpr uint32_t kBlobBytes = kOffB3 + 16;                      // 39888
constexpr uint32_t kA1Bytes = (kTileRows / 8) * kInChunks * 128;  // 16384
constexpr uint32_t kA2Bytes = (kTileRows / 8) * kHidChunks * 128; // 28672: a 128 x 112 bf16 operand in SHARED memory (masked_policy.cu)
// dynamic shared memory: the weight blob, then the buffers of the kernel's tile group(s) (a group = 128 threads that push
// tiles through the net independently of the other groups of the CTA, sharing only the weights); each kernel lays out its
// own groups — all this file needs is a 16 KB layer-1 A operand (`a1buf`) per tile in flight
constexpr uint32_t kSmemBlob = 0, kSmemGroups = (kBlobBytes + 127) / 128 * 128;
// Tensor-memory columns of one tile group: layer 1 accumulates into [0, 112); epilogue 1 compacts them IN PLACE to layer 2's
// A operand, bf16 pairs in [0, 56) (column j is written after columns 2 j, 2 j + 1 have been read); layer 2 accumulates
// into [56, 168), whose first 56 columns are layer 1's, all read by then.
constexpr uint32_t kTmemA2Cols = kHidPad / 2, kTmemAcc2 = kTmemA2Cols, kTmemColsPerGroup = kTmemAcc2 + kHidPad;   // 56, 56, 168

// The search kernel's tile (mlp_tile) squeezes into 160 columns — two allocations, 128 + 32, so that THREE of its CTAs share an
// SM's 512 (one allocation would be 256: two CTAs) — by giving layer 1 and layer 2 the same accumulator columns and parking the
// operand elsewhere: block A [0, 112) both accumulators (layer 2 starts after epilogue 1 has read layer 1's), block A [112, 128)
// operand K-chunks 0, 1, block B [0, 32) K-chunks 2..5, and the last K-chunk (units 96..111) in SHARED memory — a tcgen05.mma
// picks its A operand from either memory per instruction.
constexpr uint32_t kSearchTmemA = 128, kSearchTmemB = 32, kSearchA2InA = 112;
constexpr uint32_t kA2TailBytes = (kTileRows / 8) * 2 * 128;   // 4096: a [128 x 16] bf16 operand, canonical layout with two K-chunks

// Barrier over the 128 threads of one tile group (named barrier `id`; id 0 with a single group is __syncthreads).
__device__ __forceinline__ void group_sync(int id) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(kTileRows) : "memory"); }

// segments of the 48-vector [action, obs47] and their normalisation ranges (preprocessing.py:21-47)
struct Segment { int begin, end; float lo, hi; };
static const Segment kSegments[7] = {{0, 1, 0, 103}, {1, 11, 0, 103}, {11, 12, 0, 6}, {12, 16, 1, 5}, {16, 20, 0, 103}, {20, 24, 1, 10}, {24, 48, 0, 103}};

static inline uint16_t float_to_bf16_rne(float f) {
    uint32_t u;
    memcpy(&u, &f, 4);
    if ((u & 0x7F800000u) == 0x7F800000u) return (uint16_t)(u >> 16);   // inf / nan: truncate
    u += 0x7FFFu + ((u >> 16) & 1u);
    return (uint16_t)(u >> 16);
}
static inline float bf16_to_float(uint16_t h) {
    const uint32_t u = (uint32_t)h << 16;
    float f;
    memcpy(&f, &u, 4);
    return f;
}

// max(x, 0) rounded to bf16, two at a time, in one instruction (x1 lands in the upper half).
__device__ __forceinline__ uint32_t relu_pack_bf16x2(float x0, float x1) {
    uint32_t d;
    asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(x1), "f"(x0));
    return d;
}
// bf16 bit pattern of a small integer (|v| <= 256 is exact).
__device__ __forceinline__ uint32_t bf16_bits(int v) { return (uint32_t)__bfloat16_as_ushort(__int2bfloat16_rn(v)); }
__device__ __forceinline__ uint32_t bf16x2_bits(int lo, int hi) { return bf16_bits(lo) | (bf16_bits(hi) << 16); }

// Bring-up instrument (-DNIMMT_PHASE_CLOCKS): thread 0 of block 0 accumulates clock64() deltas per phase
// and prints them when the kernel ends.  Compiled out of the product build.
struct PhaseClock {
#ifdef NIMMT_PHASE_CLOCKS
    long long acc[16] = {0}, last = 0;
    __device__ __forceinline__ void start() { last = clock64(); }
    __device__ __forceinline__ void mark(int i) {
        const long long now = clock64();
        acc[i] += now - last;
        last = now;
    }
    __device__ void print(const char* const* names, int n) {
        if (blockIdx.x == 0 && threadIdx.x == 0)
            for (int i = 0; i < n; ++i) printf("phase %-10s %12lld cycles\n", names[i], acc[i]);
    }
#else
    __device__ __forceinline__ void start() {}
    __device__ __forceinline__ void mark(int) {}
#endif
};

// Once per kernel: the constant chunks (features 48..63) of this thread's row of the layer-1 A operand.
__device__ __forceinline__ void init_feature_constants(uint8_t* a1buf, int row) {
    *reinterpret_cast<uint4*>(a1buf + canon_off(row, kBiasCol, kInChunks)) = make_uint4(0x3F803F80u, 0u, 0u, 0u);   // 1.0, 1.0, 0 ..
    *reinterpret_cast<uint4*>(a1buf + canon_off(row, kBiasCol + 8, kInChunks)) = make_uint4(0u, 0u, 0u, 0u);
}
// Once per kernel (mlp_tile's callers): hidden units 104..111 of this thread's row of the shared-memory tail of layer 2's A operand
// do not exist.
__device__ __forceinline__ void init_a2_tail(uint8_t* a2tail, int row) {
    *reinterpret_cast<uint4*>(a2tail + canon_off(row, 8, 2)) = make_uint4(0u, 0u, 0u, 0u);
}
// Chunk c (features 8 c .. 8 c + 7, bf16) of one row of the layer-1 A operand.
__device__ __forceinline__ void store_feature_chunk(uint8_t* a1buf, int row, int c, uint4 v) {
    *reinterpret_cast<uint4*>(a1buf + canon_off(row, 0, kInChunks) + c * 128) = v;
}

// See mlp_tile's while_mma1: marks a value as needed at this point of the instruction stream.
__device__ __forceinline__ void pin_result(float& x) { asm volatile("" : "+f"(x)); }

// ReLU epilogue for hidden chunks [C0, C1) of 16 columns: TMEM -> max(x, 0) -> bf16 pairs -> TMEM, 16 accumulator columns at
// `src` into 8 operand columns at `dst`.  The chunk group's loads are all issued before the single wait, so their latencies
// overlap — and so that, when dst == src, every column a store overwrites has been read (in-place compaction: kTmemA2Cols).
template <int C0, int C1>
__device__ __forceinline__ void relu_chunks_to_operand(uint32_t src, uint32_t dst) {
    uint32_t v[C1 - C0][16];
#pragma unroll
    for (int c = C0; c < C1; ++c) tmem_ld16(src + c * 16, v[c - C0]);
    tmem_ld_wait();
#pragma unroll
    for (int c = C0; c < C1; ++c) {
        uint32_t packed[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) packed[i] = relu_pack_bf16x2(__uint_as_float(v[c - C0][2 * i]), __uint_as_float(v[c - C0][2 * i + 1]));
        tmem_st8(dst + c * 8, packed);
    }
}

// A whole accumulator row (this thread's TMEM lane, 112 columns at `src`) -> the next layer's A operand (56 columns at `dst`;
// dst == src or disjoint from it).  4 + 2 chunks + a half bound the live registers.  TMEM reads are the scarce resource of
// the tile (64 B/clk per SM sub-partition): only the 104 columns that exist are read; units 100, 101 are the constant-1 units
// that carry the next layer's bias, k = 104..111 do not exist and are written as zeros.  Ends with the wait that makes the
// stores visible to a tcgen05.mma issued after the next barrier.
__device__ __forceinline__ void relu_to_operand(uint32_t src, uint32_t dst) {
    relu_chunks_to_operand<0, 4>(src, dst);
    relu_chunks_to_operand<4, 6>(src, dst);
    uint32_t v[8];
    tmem_ld8(src + 96, v);
    tmem_ld_wait();
    uint32_t packed[8];
#pragma unroll
    for (int i = 0; i < 4; ++i) packed[i] = relu_pack_bf16x2(__uint_as_float(v[2 * i]), __uint_as_float(v[2 * i + 1]));
#pragma unroll
    for (int i = 4; i < 8; ++i) packed[i] = 0u;
    tmem_st8(dst + 48, packed);
    tmem_st_wait();
}

// Epilogue 2 + layer 3 for hidden chunks [C0, C1): part += (w3 / 2) . |acc|, four independent chains.  relu(h) is
// (h + |h|) / 2: the linear half of the head, sum_n (w3_n / 2) h_n, is one more output unit of layer 2's GEMM (units
// 100..102: three bf16 terms of the combined row, see nimmt_policy_pack_weights), so the threads only add the |h| half —
// one FFMA with an |x| operand per column instead of a max and an FFMA.
template <int C0, int C1>
__device__ __forceinline__ void epilogue2_chunks(uint32_t lane_taddr, const float* w3, float (&part)[4]) {
    uint32_t v[C1 - C0][16];
#pragma unroll
    for (int c = C0; c < C1; ++c) tmem_ld16(lane_taddr + c * 16, v[c - C0]);
    tmem_ld_wait();
#pragma unroll
    for (int c = C0; c < C1; ++c)
#pragma unroll
        for (int i = 0; i < 16; ++i) part[i & 3] = fmaf(fabsf(__uint_as_float(v[c - C0][i])), w3[c * 16 + i], part[i & 3]);
}

// Epilogue 2 + layer 3 for one row: logit = w3 . relu(acc) + b3 = (linear half from the GEMM) + (w3 / 2) . |acc| + b3, fp32.
// acc2_taddr: this thread's lane of layer 2's accumulator.  When it returns, every TMEM load has completed.
__device__ __forceinline__ float head_from_acc2(uint32_t acc2_taddr, const float* w3, float b3) {
    float part[4] = {0.0f, 0.0f, 0.0f, 0.0f};
    epilogue2_chunks<0, 4>(acc2_taddr, w3, part);
    epilogue2_chunks<4, 6>(acc2_taddr, w3, part);
    float linear;
    {
        uint32_t v[8];   // units 96..99: the last ones that exist; units 100..102: the linear half of the head (three terms)
        tmem_ld8(acc2_taddr + 96, v);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 4; ++i) part[i] = fmaf(fabsf(__uint_as_float(v[i])), w3[96 + i], part[i]);
        linear = (__uint_as_float(v[6]) + __uint_as_float(v[5])) + __uint_as_float(v[4]);
    }
    return (b3 + linear) + ((part[0] + part[1]) + (part[2] + part[3]));
}

// One 128-row tile through the three layers.  The 128 threads of a tile group call this together.
//   blob: the weights; a1buf: the tile's layer-1 A operand, filled by the caller (chunks 0..5 per tile, the constant chunks
//   once); a2tail: kA2TailBytes of shared memory (init_a2_tail once); tmem_base, tmem_b: the group's two tensor-memory blocks
//   (kSearchTmemA, kSearchTmemB columns); tid: 0..127 within the group; bar_id: the group's
//   named barrier; while_mma1(token = 0), while_mma2(): work the caller wants done while layer 1's / layer 2's MMAs run (the
//   threads would only sleep on the mbarrier) — the search kernel draws its Gumbel variates there.  A group barrier lies
//   between the two.  Returns this thread's row's logit.
template <class F1, class F2>
__device__ __forceinline__ float mlp_tile(const uint8_t* blob, const uint8_t* a1buf, uint8_t* a2tail, uint32_t tmem_base, uint32_t tmem_b, uint64_t* bar,
                                          uint32_t& phase, int tid, int bar_id, PhaseClock& pc, F1 while_mma1, F2 while_mma2) {
    const int warp = tid >> 5;
    const uint32_t a1 = smem_u32(a1buf), a2t = smem_u32(a2tail);
    const uint32_t w1 = smem_u32(blob + kOffW1), w2 = smem_u32(blob + kOffW2);
    const float* w3 = reinterpret_cast<const float*>(blob + kOffW3);
    const float b3 = *reinterpret_cast<const float*>(blob + kOffB3);
    constexpr uint32_t idesc = umma_idesc_bf16(kTileRows, kHidPad);
    const uint32_t lane_taddr = tmem_base + ((uint32_t)(warp * 32) << 16), lane_b = tmem_b + ((uint32_t)(warp * 32) << 16);

    // ---- layer 1 ----
    fence_async_smem();          // the caller's feature stores -> visible to the tensor-core (async) proxy
    tc_fence_before_sync();
    group_sync(bar_id);
    pc.mark(1);
    // (thread 0 issues from a divergent branch on purpose: the other lanes of its warp run while_mma1 while it sits in the
    // blocking MMA issue; the warp-uniform form — elected lane + __syncwarp, as the batched kernel's tile warpgroups use — was
    // measured here and costs the search ~100 cycles per turn)
    if (tid == 0) {
        tc_fence_after_sync();
#pragma unroll
        for (int ks = 0; ks < kInPad / 16; ++ks)
            umma_bf16(tmem_base, umma_desc(a1 + ks * 256, 128, kInChunks * 128), umma_desc(w1 + ks * 256, 128, kInChunks * 128), idesc, ks > 0);
        umma_commit(bar);
    }
    {   // caller's work, pinned between the MMA issue and the wait: its input passes through a volatile asm placed
        // here and its results through another one (see pin_result), which keeps the compiler from sinking it to the
        // first use after the epilogues
        uint32_t token;
        asm volatile("mov.u32 %0, 0;" : "=r"(token));
        while_mma1(token);
    }
    mbar_wait_mma(bar, phase);
    phase ^= 1u;
    tc_fence_after_sync();
    pc.mark(2);
    // epilogue 1: ReLU, round to bf16; K-chunks 0, 1 of layer 2's A operand into block A behind the accumulator, 2..5 into block B,
    // the last one (units 96..103; 104..111 are zeros written once, init_a2_tail) into shared memory
    relu_chunks_to_operand<0, 2>(lane_taddr, lane_taddr + kSearchA2InA);
    relu_chunks_to_operand<2, 6>(lane_taddr, lane_b - 16);               // chunk c -> lane_b + 8 (c - 2)
    {
        uint32_t v[8];
        tmem_ld8(lane_taddr + 96, v);
        tmem_ld_wait();
        uint32_t packed[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) packed[i] = relu_pack_bf16x2(__uint_as_float(v[2 * i]), __uint_as_float(v[2 * i + 1]));
        *reinterpret_cast<uint4*>(a2tail + canon_off(tid, 0, 2)) = make_uint4(packed[0], packed[1], packed[2], packed[3]);
    }
    tmem_st_wait();
    fence_async_smem();          // the tail chunk -> visible to the tensor core's (async) proxy
    // ---- layer 2: A from tensor memory (six K-steps) and from shared memory (the last), accumulator over layer 1's ----
    tc_fence_before_sync();
    pc.mark(3);
    group_sync(bar_id);          // every lane's operand row is written and its accumulator read: columns [0, 112) may be overwritten
    pc.mark(4);
    if (tid == 0) {
        tc_fence_after_sync();
#pragma unroll
        for (int ks = 0; ks < kHidPad / 16 - 1; ++ks)
            umma_bf16_ts(tmem_base, ks < 2 ? tmem_base + kSearchA2InA + ks * 8 : tmem_b + (ks - 2) * 8, umma_desc(w2 + ks * 256, 128, kHidChunks * 128), idesc,
                         ks > 0);
        umma_bf16(tmem_base, umma_desc(a2t, 128, 2 * 128), umma_desc(w2 + (kHidPad / 16 - 1) * 256, 128, kHidChunks * 128), idesc, 1);
        umma_commit(bar);
    }
    while_mma2();
    mbar_wait_mma(bar, phase);
    phase ^= 1u;
    tc_fence_after_sync();
    pc.mark(5);
    const float logit = head_from_acc2(lane_taddr, w3, b3);
    tc_fence_before_sync();      // ordered before the caller's next barrier / the next tile's MMA
    pc.mark(6);
    return logit;
}

// Softmax over one decision's hand slots by warp shuffles.  The decision's ten rows sit in lanes
// first .. first + 9 of the warp; live_mask has bit s set when slot s holds a card.  Every lane of the
// decision ends up with all ten un-normalised weights e[s] = exp(logit_s - max) (0 for empty slots) and
// their sum z and the maximum m; probabilities are e[s] / z (agents/mcts.py:207,227: Softmax(dim=0) over the
// rows).  Must be called by all 32 lanes.
__device__ __forceinline__ void decision_softmax(float logit, int first, uint32_t live_mask, float (&e)[kSlots], float& z, float& m) {
    m = -INFINITY;
#pragma unroll
    for (int s = 0; s < kSlots; ++s) {
        e[s] = __shfl_sync(0xffffffffu, logit, first + s);
        if ((live_mask >> s) & 1u) m = fmaxf(m, e[s]);
    }
    z = 0.0f;
#pragma unroll
    for (int s = 0; s < kSlots; ++s) {
        e[s] = (live_mask >> s) & 1u ? __expf(e[s] - m) : 0.0f;
        z += e[s];
    }
}

}  // namespace nimmt
