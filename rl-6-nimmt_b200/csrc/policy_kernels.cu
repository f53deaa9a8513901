// policy_kernels.cu — Alpha0.5 leaf evaluation on the 5th-generation tensor cores.
//
// Replaces PolicyMCSAgent._compute_policy (agents/mcts.py:219-228) = SechsNimmtStateNormalization
// (utils/preprocessing.py:12-57) + MultiHeadedMLP(48, (100, 100), (1,)) with ReLU
// (utils/nets.py:100-132) + softmax over the legal cards of one decision.
//
// One tile = 128 rows = 12 decisions x 10 card slots (row = [card | obs47 | 1 1 0..], raw integer features;
// the affine normalisation is folded into layer 1 when the weights are packed, and so are the biases).
//   layer 1: [128 x 64] x [64 x 112]   4 x tcgen05.mma (K = 16 each), A and B bf16 in shared memory,
//            fp32 accumulator in TMEM; epilogue tcgen05.ld -> ReLU -> bf16 -> tcgen05.st back into TMEM
//   layer 2: [128 x 112] x [112 x 112] 7 x tcgen05.mma with A from TMEM, B in shared memory; epilogue ReLU, and
//   layer 3 (100 -> 1) as an fp32 dot product in the same epilogue (N = 1 is not a tensor-core shape)
//   softmax over each decision's legal slots by two warp reductions, probabilities out.
// Hidden width 100 is padded to 112 (tcgen05 N granularity 16 at M = 128); padded weights are zero.
// Features are built on chip from int8 observations, so the GEMMs never read activations from HBM.
#include "policy_tile.cuh"

namespace nimmt {

// obs: int8 [D][47]; probs: float [D][10] (0 for empty slots); logits (optional): float [D][10].
constexpr uint32_t kProbTmemCols = 512;   // allocations are powers of two; 3 x 168 = 504 are used
constexpr int kObsWords = kDecPerTile * kObs / 4;   // 141 32-bit words of observation bytes per tile
constexpr uint16_t kBf16NoCard = 0xBF80;            // -1: an empty hand slot (env.py:209-210)

// k_policy_probs: grid-stride over tiles of 12 decisions, warp-specialised.
// One persistent CTA per SM, 16 warps with two jobs, decoupled by mbarriers:
//   warps 12..15  PRODUCERS: observation bytes (fetched two iterations ahead) -> bf16 rows -> the layer-1 A operand of ring
//                 stage k % 6, then fence.proxy.async + arrive on a1_full[stage]; two tiles per iteration, so that the two
//                 instruction streams interleave (one warp per scheduler: the job is bound by instruction latency)
//   warps 0..11   three TILE warpgroups: warpgroup s owns tensor-memory slot s (168 columns, policy_tile.cuh) and every third
//                 tile, and drives it through the net by itself: its warp 0 issues the tile's MMAs (warp-uniform control flow,
//                 the elected lane issues: under `if (tid == 0)` ptxas wraps every tcgen05 instruction in an ELECT / BRA.U.ANY
//                 loop and rebuilds the barriers' addresses from SR_CgaCtaId each time), each layer followed by a tcgen05.commit
//                 onto the mbarrier the warpgroup then sleeps on:
//                   layer 1 -> ReLU -> bf16 back into tensor memory -> named barrier -> layer 2 -> first half of the head ->
//                   named barrier -> LAYER 1 OF THE WARPGROUP'S NEXT TILE (the columns it overwrites are in registers by then) ->
//                   rest of the head, softmax over the decision (two redux.sync), probabilities out — under that layer 1.
//                 No hand-off goes through a third party: a central MMA-issuing warp was measured, and every hop through it
//                 (arrive -> poll -> issue) cost ~250 cycles of a chain that has only three tiles in flight to hide it.
constexpr int kWsSlots = 3;                                   // tensor-memory slots = tile warpgroups
constexpr int kWsStages = 6;                                  // layer-1 operand ring (even: a producer iteration fills two)
constexpr int kWsProdGroups = 1;                              // producer warpgroups, alternating pairs of tiles (two were measured: slower,
                                                              // the producers' and the tile warpgroups' instruction streams add up)
constexpr int kWsProdWarp0 = 4 * kWsSlots, kWsThreads = 32 * (kWsProdWarp0 + 4 * kWsProdGroups);   // 12, 512
constexpr uint32_t kWsSmemA1 = kSmemGroups, kWsSmemRows = kWsSmemA1 + kWsStages * kA1Bytes;
constexpr uint32_t kWsRowsBytes = (kDecPerTile * kIn * 2 + 127) / 128 * 128;   // bf16 [12][48] staged decisions; four of them
constexpr uint32_t kWsSmemBytes = kWsSmemRows + kWsProdGroups * 4 * kWsRowsBytes;
// Bring-up instruments (compiled out of the product build; profiles/r02_n_*): -DWS_L1_STEPS=1 -DWS_L2_STEPS=1 issue one MMA per
// layer, -DWS_SKIP_PRODUCE / _EPI / _SOFTMAX / _STORE drop one job's work and keep its hand-offs — wrong results, right protocol —
// which is how the time of a tile was split between the tensor core, the epilogues, the producers and the hand-offs.
#ifndef WS_L1_STEPS
#define WS_L1_STEPS (kInPad / 16)
#define WS_L2_STEPS (kHidPad / 16)
#endif
constexpr int kWsTileBarrier0 = 1, kWsProdBarrier0 = kWsTileBarrier0 + kWsSlots;   // named barriers: tile warpgroup s = 1 + s; producer warpgroup p = 4 + p
#ifdef NIMMT_PHASE_CLOCKS
#define WS_CLK_DECL(n) long long wsc[n] = {0}, wsl = clock64()
#define WS_CLK(i) do { const long long now_ = clock64(); wsc[i] += now_ - wsl; wsl = now_; } while (0)
#else
#define WS_CLK_DECL(n)
#define WS_CLK(i)
#endif

// Softmax over one decision's hand slots with two warp reductions: the decision's ten rows sit in lanes first .. first + 9 of the
// warp (dmask), lanes 30, 31 are dead rows and reduce among themselves.  The maximum travels as an order-preserving integer key;
// the sum as 8.24 fixed point (every term is in (0, 1], the largest is 1: relative error of the sum < 3e-7).
__device__ __forceinline__ float decision_prob_redux(float logit, bool has_card, uint32_t dmask) {
    uint32_t key = __float_as_uint(logit);
    key = (key & 0x80000000u) ? ~key : (key | 0x80000000u);
    const uint32_t kmax = __reduce_max_sync(dmask, has_card ? key : 0u);
    const float m = __uint_as_float((kmax & 0x80000000u) ? (kmax & 0x7FFFFFFFu) : ~kmax);
    const float e = has_card ? __expf(logit - m) : 0.0f;
    const uint32_t z = __reduce_add_sync(dmask, (uint32_t)(e * 16777216.0f + 0.5f));
    return has_card ? __fdividef(e * 16777216.0f, (float)z) : 0.0f;
}

__global__ void __launch_bounds__(kWsThreads, 1)
k_policy_probs(const int8_t* __restrict__ obs, int64_t D, const uint8_t* __restrict__ blob, float* __restrict__ probs, float* __restrict__ logits) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ uint64_t a1_full[kWsStages], a1_empty[kWsStages], acc1_full[kWsSlots], acc2_full[kWsSlots];
    __shared__ uint32_t tmem_slot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    for (uint32_t i = threadIdx.x * 16; i < kBlobBytes; i += blockDim.x * 16)
        *reinterpret_cast<uint4*>(smem + kSmemBlob + i) = *reinterpret_cast<const uint4*>(blob + i);
    if (threadIdx.x == 0) {
        for (int s = 0; s < kWsStages; ++s) {
            mbar_init(&a1_full[s], 4);             // one lane of every producer warp arrives
            mbar_init(&a1_empty[s], 1);            // tcgen05.commit: the MMAs that read the stage have completed
        }
        for (int s = 0; s < kWsSlots; ++s) {
            mbar_init(&acc1_full[s], 1);           // tcgen05.commit
            mbar_init(&acc2_full[s], 1);
        }
        fence_barrier_init();
    }
    if (warp == 0) tmem_alloc(&tmem_slot, kProbTmemCols);
    if (threadIdx.x < kTileRows)
        for (int s = 0; s < kWsStages; ++s) init_feature_constants(smem + kWsSmemA1 + s * kA1Bytes, threadIdx.x);
    fence_async_smem();          // the weights and the constant chunks -> visible to the tensor core's (async) proxy
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = tmem_slot;
    const int64_t num_tiles = (D + kDecPerTile - 1) / kDecPerTile;
    // this CTA's tiles: blockIdx.x + k gridDim.x, k = 0 .. n_local - 1; tile k uses tensor-memory slot k % 3 and operand stage k % 6
    const int64_t n_local = (int64_t)blockIdx.x < num_tiles ? (num_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;

    if (warp >= kWsProdWarp0) {
        const int pw = (warp - kWsProdWarp0) >> 2;                       // producer warpgroup: iterations pw, pw + 2, ..
        const int tid = (threadIdx.x - 32 * kWsProdWarp0) & (kTileRows - 1);   // = the row of the tile this thread builds
        const int dloc = lane / kSlots, slot = lane % kSlots;
        const int dec_local = (tid >> 5) * kDecPerWarp + (dloc < kDecPerWarp ? dloc : kDecPerWarp - 1);
        // a tile's 12 x 47 observation bytes as 141 coalesced 4-byte loads (564 B, 4-byte aligned)
        const int64_t total_bytes = D * kObs;
        auto load_word = [&](int64_t tile, int wd) -> uint32_t {
            const int64_t byte = tile * (kDecPerTile * kObs) + 4 * wd;
            uint32_t v = 0;
            if (byte + 4 <= total_bytes) v = *reinterpret_cast<const uint32_t*>(obs + byte);
            else for (int i = 0; i < 4; ++i) if (byte + i < total_bytes) v |= (uint32_t)(uint8_t)obs[byte + i] << (8 * i);
            return v;
        };
        auto fetch = [&](int64_t k, uint32_t& w0, uint32_t& w1) {
            if (k < n_local) {
                const int64_t tile = blockIdx.x + k * gridDim.x;
                w0 = load_word(tile, tid);
                if (tid < kObsWords - kTileRows) w1 = load_word(tile, kTileRows + tid);
            }
        };
        // int8 -> bf16 straight from the registers: byte j of the tile is feature 1 + j % 47 of decision j / 47
        auto stage_word = [&](uint16_t* rows, uint32_t w, int wd) {
            const int j = 4 * wd, d = j / kObs, k = j - d * kObs;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int wrap = k + i >= kObs;
                const float f = (float)(int8_t)(w >> (8 * i));                     // |f| <= 128: the upper half of the fp32 is the exact bf16
                rows[(d + wrap) * kIn + 1 + k + i - wrap * kObs] = (uint16_t)(__float_as_uint(f) >> 16);
            }
        };
        // this thread's row of tile k, [card | obs47], from the staged decisions into operand stage `stage`
        auto build_row = [&](int64_t k, const uint16_t* rows, int stage) {
            const int64_t dec = (blockIdx.x + k * gridDim.x) * kDecPerTile + dec_local;
            const bool in_range = dloc < kDecPerWarp && dec < D;
            const uint32_t card = in_range ? rows[dec_local * kIn + 1 + slot] : kBf16NoCard;   // hand slot: the candidate card, -1 if empty
            uint8_t* a1buf = smem + kWsSmemA1 + stage * kA1Bytes;
#pragma unroll
            for (int c = 0; c < kFeatChunks; ++c) {
                uint4 v = *reinterpret_cast<const uint4*>(rows + dec_local * kIn + 8 * c);
                if (c == 0) v.x = (v.x & 0xFFFF0000u) | card;
                store_feature_chunk(a1buf, tid, c, v);
            }
        };
        // iteration `it` of the CTA fills tiles 2 it and 2 it + 1 (operand stages 2 (it % 3) and + 1, use it / 3 of them)
        uint32_t ca0 = 0, ca1 = 0, cb0 = 0, cb1 = 0, na0 = 0, na1 = 0, nb0 = 0, nb1 = 0;   // this iteration's tiles; the warpgroup's next iteration's
        fetch(2 * pw, ca0, ca1);
        fetch(2 * pw + 1, cb0, cb1);
        fetch(2 * (pw + kWsProdGroups), na0, na1);
        fetch(2 * (pw + kWsProdGroups) + 1, nb0, nb1);
        const int prod_bar = kWsProdBarrier0 + pw;
        uint8_t* my_rows = smem + kWsSmemRows + pw * 4 * kWsRowsBytes;
        uint32_t flip = 0;
        WS_CLK_DECL(4);
        for (uint32_t it = pw; 2 * (int64_t)it < n_local; it += kWsProdGroups, flip ^= 2u) {
            const int64_t k = 2 * (int64_t)it;
            const bool two = k + 1 < n_local;
            const uint32_t use = it / 3u;
            const int stage = 2 * (int)(it - 3u * use);
            uint16_t* rows_a = reinterpret_cast<uint16_t*>(my_rows + (flip + 0) * kWsRowsBytes);
            uint16_t* rows_b = reinterpret_cast<uint16_t*>(my_rows + (flip + 1) * kWsRowsBytes);
#ifndef WS_SKIP_PRODUCE
            stage_word(rows_a, ca0, tid);
            stage_word(rows_b, cb0, tid);
            if (tid < kObsWords - kTileRows) {
                stage_word(rows_a, ca1, kTileRows + tid);
                stage_word(rows_b, cb1, kTileRows + tid);
            }
#endif
            ca0 = na0, ca1 = na1, cb0 = nb0, cb1 = nb1;
            fetch(k + 4 * kWsProdGroups, na0, na1);
            fetch(k + 4 * kWsProdGroups + 1, nb0, nb1);
            WS_CLK(0);
            // the staging buffers of this parity were last read two of this warpgroup's iterations ago, and every thread has
            // passed the barrier of the iteration in between since
            asm volatile("bar.sync %0, %1;" ::"r"(prod_bar), "n"(kTileRows) : "memory");
            WS_CLK(1);
            if (use > 0) {           // the MMAs that read these operand stages six tiles ago have completed
                mbar_wait_mma(&a1_empty[stage], (use - 1u) & 1u);
                if (two) mbar_wait_mma(&a1_empty[stage + 1], (use - 1u) & 1u);
            }
            WS_CLK(2);
#ifndef WS_SKIP_PRODUCE
            build_row(k, rows_a, stage);
            if (two) build_row(k + 1, rows_b, stage + 1);
#endif
            fence_async_smem();
            __syncwarp();
            if (lane == 0) {
                mbar_arrive(&a1_full[stage]);
                if (two) mbar_arrive(&a1_full[stage + 1]);
            }
            WS_CLK(3);
        }
#ifdef NIMMT_PHASE_CLOCKS
        if (blockIdx.x == 0 && tid == 0 && pw == 0) printf("producer: stage %lld, barrier %lld, wait stage free %lld, build %lld cycles\n", wsc[0], wsc[1], wsc[2], wsc[3]);
#endif
    } else {
        const int s = warp >> 2, q = warp & 3;                     // slot; TMEM lane quarter (a warp reads lanes 32 (warp % 4) ..)
        const uint32_t slot_taddr = tmem_base + s * kTmemColsPerGroup;
        const uint32_t lane_taddr = slot_taddr + ((uint32_t)(q * 32) << 16), acc2_taddr = lane_taddr + kTmemAcc2;
        const int dloc = lane / kSlots, slot = lane % kSlots;
        const uint32_t dmask = dloc < kDecPerWarp ? 0x3FFu << (dloc * kSlots) : 0xC0000000u;
        const int dec_local = q * kDecPerWarp + (dloc < kDecPerWarp ? dloc : kDecPerWarp - 1);
        const float* w3 = reinterpret_cast<const float*>(smem + kSmemBlob + kOffW3);
        const float b3 = *reinterpret_cast<const float*>(smem + kSmemBlob + kOffB3);
        // the MMA side (warp 0 of the warpgroup, all lanes in step; the elected lane issues)
        const uint32_t w1 = smem_u32(smem + kSmemBlob + kOffW1), w2 = smem_u32(smem + kSmemBlob + kOffW2), a1 = smem_u32(smem + kWsSmemA1);
        const uint32_t b_a1_full = smem_u32(a1_full), b_a1_empty = smem_u32(a1_empty), b_acc1 = smem_u32(&acc1_full[s]), b_acc2 = smem_u32(&acc2_full[s]);
        constexpr uint32_t idesc = umma_idesc_bf16(kTileRows, kHidPad);
        const bool issuer = elect_one() != 0u;
        const int bar_id = kWsTileBarrier0 + s;
        // layer 1 of this warpgroup's i-th tile (tile k = s + 3 i: operand stage s or s + 3 alternately, use i / 2 of it)
        auto issue_layer1 = [&](uint32_t i) {
            const uint32_t stage = (uint32_t)s + 3u * (i & 1u);
            mbar_wait_mma_a(b_a1_full + 8u * stage, (i >> 1) & 1u);
            tc_fence_after_sync();
            if (issuer) {
                const uint32_t a = a1 + stage * kA1Bytes;
#pragma unroll
                for (int ks = 0; ks < WS_L1_STEPS; ++ks)
                    umma_bf16(slot_taddr, umma_desc(a + ks * 256, 128, kInChunks * 128), umma_desc(w1 + ks * 256, 128, kInChunks * 128), idesc, ks > 0);
                umma_commit_a(b_acc1);
                umma_commit_a(b_a1_empty + 8u * stage);
            }
            __syncwarp();
        };
        // this row's hand slot, -1 if empty (env.py:209-210): one byte of the observations, fetched ONE TILE AHEAD of its use by a
        // volatile load (left to the compiler, the load sinks to its first use, after the head, and its latency lands on the
        // critical path of the slot)
        const int64_t dec_step = (int64_t)kWsSlots * gridDim.x * kDecPerTile;
        int64_t dec = ((int64_t)blockIdx.x + (int64_t)s * gridDim.x) * kDecPerTile + dec_local;
        auto load_card = [&](int64_t d) -> int {
            int c = -1;
            if (dloc < kDecPerWarp && d < D) asm volatile("ld.global.nc.s8 %0, [%1];" : "=r"(c) : "l"(obs + d * kObs + slot));
            return c;
        };
        int next_card = load_card(dec);
        const uint32_t n_mine = n_local > s ? (uint32_t)((n_local - s + kWsSlots - 1) / kWsSlots) : 0u;
        if (q == 0 && n_mine > 0) issue_layer1(0);
        WS_CLK_DECL(6);
        for (uint32_t i = 0; i < n_mine; ++i, dec += dec_step) {
            const uint32_t par = i & 1u;
            const bool in_range = dloc < kDecPerWarp && dec < D;
            const int card = next_card;
            next_card = load_card(dec + dec_step);
            WS_CLK(5);
            mbar_wait_mma_a(b_acc1, par);
            WS_CLK(0);
            tc_fence_after_sync();
#ifndef WS_SKIP_EPI
            relu_to_operand(lane_taddr, lane_taddr);
#endif
            tc_fence_before_sync();
            asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "n"(kTileRows) : "memory");   // every row's operand is in tensor memory, its accumulator read
            if (q == 0) {
                tc_fence_after_sync();
                if (issuer) {
#pragma unroll
                    for (int ks = 0; ks < WS_L2_STEPS; ++ks)
                        umma_bf16_ts(slot_taddr + kTmemAcc2, slot_taddr + ks * 8, umma_desc(w2 + ks * 256, 128, kHidChunks * 128), idesc, ks > 0);
                    umma_commit_a(b_acc2);
                }
                __syncwarp();
            }
            WS_CLK(1);
            mbar_wait_mma_a(b_acc2, par);
            WS_CLK(2);
            tc_fence_after_sync();
            // the head (policy_tile.cuh::head_from_acc2) in two halves: accumulator columns 0..63 are tensor-memory columns 56..119,
            // which cover everything of this slot that the next tile's layer 1 overwrites ([0, 112)) — that layer 1 is issued
            // between the halves
            float part[4] = {0.0f, 0.0f, 0.0f, 0.0f};
#ifndef WS_SKIP_EPI
            epilogue2_chunks<0, 4>(acc2_taddr, w3, part);
#endif
            tc_fence_before_sync();
            asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "n"(kTileRows) : "memory");
            if (q == 0 && i + 1 < n_mine) issue_layer1(i + 1);
#ifndef WS_SKIP_EPI
            epilogue2_chunks<4, 6>(acc2_taddr, w3, part);
#endif
            float linear;
            {
                uint32_t v[8];   // units 96..99: the last ones that exist; units 100..102: the linear half of the head (three terms)
                tmem_ld8(acc2_taddr + 96, v);
                tmem_ld_wait();
#pragma unroll
                for (int i2 = 0; i2 < 4; ++i2) part[i2] = fmaf(fabsf(__uint_as_float(v[i2])), w3[96 + i2], part[i2]);
                linear = (__uint_as_float(v[6]) + __uint_as_float(v[5])) + __uint_as_float(v[4]);
            }
            const float logit = (b3 + linear) + ((part[0] + part[1]) + (part[2] + part[3]));
            tc_fence_before_sync();          // the rest of the accumulator has been read before this warpgroup's next layer 2
            WS_CLK(3);
            const bool has_card = card >= 0;
#ifdef WS_SKIP_SOFTMAX
            const float prob = logit;
#else
            const float prob = decision_prob_redux(logit, has_card, dmask);
#endif
#ifdef WS_SKIP_STORE
            if (in_range && prob == 12345.678f) {
#else
            if (in_range) {
#endif
                probs[dec * kSlots + slot] = prob;
                if (logits) logits[dec * kSlots + slot] = has_card ? logit : 0.0f;
            }
            WS_CLK(4);
        }
#ifdef NIMMT_PHASE_CLOCKS
        if (blockIdx.x == 0 && lane == 0 && q == 0)
            printf("tile wg %d: wait acc1 %lld, relu + issue L2 %lld, wait acc2 %lld, head + issue L1 %lld, softmax+out %lld, loop top %lld cycles\n", s, wsc[0], wsc[1], wsc[2], wsc[3], wsc[4], wsc[5]);
#endif
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base, kProbTmemCols);
}

}  // namespace nimmt

using namespace nimmt;

extern "C" {

size_t nimmt_policy_weights_bytes(void) { return kBlobBytes; }

int nimmt_policy_pack_weights(const float* w1, const float* b1, const float* w2, const float* b2, const float* w3, float b3,
                              void* blob_host) {
    if (!w1 || !b1 || !w2 || !b2 || !w3 || !blob_host) return NIMMT_E_BADARG;
    uint8_t* blob = static_cast<uint8_t*>(blob_host);
    memset(blob, 0, kBlobBytes);
    double scale[kIn], shift[kIn];
    for (const Segment& s : kSegments)
        for (int k = s.begin; k < s.end; ++k) {
            scale[k] = 2.0 / ((double)s.hi - s.lo);                 // preprocessing.py:56-57 with out range [-1, 1]
            shift[k] = -1.0 - 2.0 * s.lo / ((double)s.hi - s.lo);
        }
    auto put1 = [&](int n, int k, uint16_t v) { *reinterpret_cast<uint16_t*>(blob + kOffW1 + canon_off(n, k, kInChunks)) = v; };
    auto put2 = [&](int n, int k, uint16_t v) { *reinterpret_cast<uint16_t*>(blob + kOffW2 + canon_off(n, k, kHidChunks)) = v; };
    // a bias as two bf16 terms: hi = bf16(b), lo = bf16(b - hi); both meet a constant-1 input in the GEMM
    auto split = [](float b, uint16_t& hi, uint16_t& lo) {
        hi = float_to_bf16_rne(b);
        lo = float_to_bf16_rne(b - bf16_to_float(hi));
    };
    float* w3p = reinterpret_cast<float*>(blob + kOffW3);
    for (int n = 0; n < kHid; ++n) {
        double acc = b1[n];
        for (int k = 0; k < kIn; ++k) {
            const double w = w1[n * kIn + k];
            acc += w * shift[k];
            put1(n, k, float_to_bf16_rne((float)(w * scale[k])));
        }
        uint16_t hi, lo;
        split((float)acc, hi, lo);
        put1(n, kBiasCol, hi);
        put1(n, kBiasCol + 1, lo);
        for (int k = 0; k < kHid; ++k) put2(n, k, float_to_bf16_rne(w2[n * kHid + k]));
        split(b2[n], hi, lo);
        put2(n, kOneUnit, hi);
        put2(n, kOneUnit + 1, lo);
        w3p[n] = 0.5f * w3[n];                 // the |h| half of relu(h) = (h + |h|) / 2
    }
    // the linear half, sum_n (w3_n / 2) h_n, as output units 100, 101, 102 of layer 2 (three bf16 terms: the two halves of
    // the head cancel wherever units are inactive, so this row is carried to ~24 bits): the combined row is formed from
    // the bf16 weights the tensor core really multiplies by, bias columns included
    for (int k = 0; k < kOneUnit + 2; ++k) {
        double acc = 0.0;
        for (int n = 0; n < kHid; ++n)
            acc += 0.5 * (double)w3[n] * (double)bf16_to_float(*reinterpret_cast<const uint16_t*>(blob + kOffW2 + canon_off(n, k, kHidChunks)));
        float rest = (float)acc;
        for (int t = 0; t < 3; ++t) {
            const uint16_t term = float_to_bf16_rne(rest);
            put2(kOneUnit + t, k, term);
            rest -= bf16_to_float(term);
        }
    }
    put1(kOneUnit, kBiasCol, 0x3F80);       // units 100, 101 of layer 1: relu(1 * 1) = 1, the inputs that carry b2
    put1(kOneUnit + 1, kBiasCol, 0x3F80);
    *reinterpret_cast<float*>(blob + kOffB3) = b3;
    return NIMMT_OK;
}

int nimmt_policy_probs(const int8_t* obs, int64_t num_decisions, const void* weights, float* probs, float* logits, void* stream) {
    if (!obs || !weights || !probs || num_decisions < 0) return NIMMT_E_BADARG;
    if (!aligned16(weights) || !aligned16(probs)) return NIMMT_E_ALIGN;
    if (num_decisions == 0) return NIMMT_OK;
    if (reinterpret_cast<uintptr_t>(obs) & 3u) return NIMMT_E_ALIGN;   // the kernel reads the observations with 32-bit loads
    const int num_sms = device_sms(current_device());
    const int64_t tiles = (num_decisions + kDecPerTile - 1) / kDecPerTile;
    static int occ_cache[kMaxDevices];
    blocks_per_sm_cached(k_policy_probs, kWsThreads, (int)kWsSmemBytes, occ_cache);   // per-device opt-in
    const unsigned blocks = (unsigned)(tiles < num_sms ? tiles : num_sms);   // persistent: one CTA per SM
    k_policy_probs<<<blocks, kWsThreads, kWsSmemBytes, (cudaStream_t)stream>>>(obs, num_decisions, static_cast<const uint8_t*>(weights), probs, logits);
    return check_launch();
}

}  // extern "C"
