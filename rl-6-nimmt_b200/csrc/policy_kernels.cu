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
//   softmax over each decision's legal slots by warp shuffles, probabilities out.
// Hidden width 100 is padded to 112 (tcgen05 N granularity 16 at M = 128); padded weights are zero.
// Features are built on chip from int8 observations, so the GEMMs never read activations from HBM.
#include "policy_tile.cuh"

namespace nimmt {

// grid-stride over tiles of 12 decisions; every 128-thread group of the CTA takes its own tiles.
// obs: int8 [D][47]; probs: float [D][10] (0 for empty slots); logits (optional): float [D][10].
// kProbGroups groups share one copy of the weights in shared memory and each own 168 TMEM columns (three groups fill the
// SM's 512), an mbarrier and a named barrier, so one group's epilogue (TMEM -> registers -> bf16 -> TMEM)
// runs under the other groups' MMAs.  Within a group, warp w carries decisions 3 w .. 3 w + 2 in lanes 0..29 (lanes 30, 31
// are dead rows), so the softmax is ten shuffles.
// A group is software-pipelined over its tiles: while layer 1 of tile i runs on the tensor core the threads convert tile
// i + 1's observation bytes (prefetched into registers one tile earlier) to bf16, and while layer 2 runs they assemble tile
// i + 1's rows into the OTHER of two layer-1 operand buffers — work that would otherwise sit between two tiles with the tensor
// core idle and the threads, later, asleep on the mbarrier.
constexpr int kProbGroups = 3;
constexpr uint32_t kProbTmemCols = 512;   // allocations are powers of two; 3 x 168 = 504 are used
static_assert(kProbGroups * kTmemColsPerGroup <= kProbTmemCols, "tensor memory");
constexpr int kObsWords = kDecPerTile * kObs / 4;   // 141 32-bit words of observation bytes per tile
// a group's shared memory: two layer-1 operands, then the tile's decisions as bf16 [12][48] ([d][0] unused, [d][1 + k] = obs k)
constexpr uint32_t kPRows = 2 * kA1Bytes, kProbGroupBytes = (kPRows + kDecPerTile * kIn * 2 + 127) / 128 * 128;   // 33920
constexpr uint32_t kProbSmemBytes = kSmemGroups + kProbGroups * kProbGroupBytes;
constexpr uint16_t kBf16NoCard = 0xBF80;            // -1: an empty hand slot (env.py:209-210)

__global__ void __launch_bounds__(kTileRows * kProbGroups, 1)
k_policy_probs(const int8_t* __restrict__ obs, int64_t D, const uint8_t* __restrict__ blob, float* __restrict__ probs, float* __restrict__ logits) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ uint64_t bars[kProbGroups];
    __shared__ uint32_t tmem_slot;

    for (uint32_t i = threadIdx.x * 16; i < kBlobBytes; i += blockDim.x * 16)
        *reinterpret_cast<uint4*>(smem + kSmemBlob + i) = *reinterpret_cast<const uint4*>(blob + i);
    if (threadIdx.x < kProbGroups) mbar_init(&bars[threadIdx.x], 1);
    if (threadIdx.x == 0) fence_barrier_init();
    if (threadIdx.x < 32) tmem_alloc(&tmem_slot, kProbTmemCols);
    const int group = threadIdx.x / kTileRows, tid = threadIdx.x % kTileRows, bar_id = 1 + group;
    uint8_t* gbuf = smem + kSmemGroups + group * kProbGroupBytes;
    init_feature_constants(gbuf, tid);
    init_feature_constants(gbuf + kA1Bytes, tid);
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = tmem_slot + group * kTmemColsPerGroup;
    uint32_t phase = 0;
    PhaseClock pc;
    uint16_t* tile_rows = reinterpret_cast<uint16_t*>(gbuf + kPRows);

    const int lane = tid & 31, warp = tid >> 5;
    const int dloc = lane / kSlots, slot = lane % kSlots;
    const int first = (dloc < kDecPerWarp ? dloc : kDecPerWarp - 1) * kSlots;
    const int dec_local = warp * kDecPerWarp + (dloc < kDecPerWarp ? dloc : kDecPerWarp - 1);

    // the tile's 12 x 47 observation bytes as 141 coalesced 4-byte loads (564 B, 4-byte aligned), fetched into registers
    // two tiles ahead of the GEMMs that consume them
    const int64_t total_bytes = D * kObs;
    auto load_word = [&](int64_t tile, int wd) -> uint32_t {
        const int64_t byte = tile * (kDecPerTile * kObs) + 4 * wd;
        uint32_t v = 0;
        if (byte + 4 <= total_bytes) v = *reinterpret_cast<const uint32_t*>(obs + byte);
        else for (int i = 0; i < 4; ++i) if (byte + i < total_bytes) v |= (uint32_t)(uint8_t)obs[byte + i] << (8 * i);
        return v;
    };
    const int64_t num_tiles = (D + kDecPerTile - 1) / kDecPerTile, stride = (int64_t)gridDim.x * kProbGroups;
    uint32_t w0 = 0, w1 = 0;
    auto fetch = [&](int64_t tile) {
        if (tile < num_tiles) {
            w0 = load_word(tile, tid);
            if (tid < kObsWords - kTileRows) w1 = load_word(tile, kTileRows + tid);
        }
    };
    // int8 -> bf16 straight from the prefetched registers: byte j of the tile is feature 1 + j % 47 of decision j / 47
    auto stage_word = [&](uint32_t w, int wd) {
        const int j = 4 * wd, d = j / kObs, k = j - d * kObs;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int wrap = k + i >= kObs;
            const float f = (float)(int8_t)(w >> (8 * i));                     // |f| <= 128: the upper half of the fp32 is the exact bf16
            tile_rows[(d + wrap) * kIn + 1 + k + i - wrap * kObs] = (uint16_t)(__float_as_uint(f) >> 16);
        }
    };
    auto stage = [&]() {
        stage_word(w0, tid);
        if (tid < kObsWords - kTileRows) stage_word(w1, kTileRows + tid);
    };
    // this thread's row of `tile`, [card | obs47], into the layer-1 operand `a1buf`; returns the card's bf16 bits
    auto build_rows = [&](int64_t tile, uint8_t* a1buf) -> uint32_t {
        const int64_t dec = tile * kDecPerTile + dec_local;
        const bool in_range = dloc < kDecPerWarp && dec < D;
        const uint32_t card = in_range ? tile_rows[dec_local * kIn + 1 + slot] : kBf16NoCard;   // hand slot: the candidate card, -1 if empty
#pragma unroll
        for (int c = 0; c < kFeatChunks; ++c) {
            uint4 v = *reinterpret_cast<const uint4*>(tile_rows + dec_local * kIn + 8 * c);
            if (c == 0) v.x = (v.x & 0xFFFF0000u) | card;
            store_feature_chunk(a1buf, tid, c, v);
        }
        return card;
    };

    int64_t tile = (int64_t)blockIdx.x * kProbGroups + group;
    uint32_t card_bits = kBf16NoCard, buf = 0;
    fetch(tile);
    if (tile < num_tiles) {
        stage();
        group_sync(bar_id);
        card_bits = build_rows(tile, gbuf);
    }
    fetch(tile + stride);
    pc.start();
    for (; tile < num_tiles; tile += stride) {
        const bool more = tile + stride < num_tiles;   // the same for every thread of the group
        uint32_t next_card = kBf16NoCard;
        const float logit = mlp_tile(
            smem + kSmemBlob, gbuf + buf * kA1Bytes, tmem_base, &bars[group], phase, tid, bar_id, pc,
            [&](uint32_t) {
                // every thread built this tile's rows before the barrier that preceded the MMA issue: tile_rows is free
                if (more) stage();
                fetch(tile + 2 * stride);
            },
            [&] {
                // the staged rows are visible (mlp_tile's barrier before layer 2); the other operand buffer was last read by the
                // MMAs of the tile before this one, which completed before that tile's epilogue
                if (more) next_card = build_rows(tile + stride, gbuf + (buf ^ 1u) * kA1Bytes);
            });
        const bool has_card = (card_bits & 0x8000u) == 0u;
        const uint32_t live = (__ballot_sync(0xffffffffu, has_card) >> first) & 0x3FFu;
        float e[kSlots], z, m;
        decision_softmax(logit, first, live, e, z, m);
        const int64_t dec = tile * kDecPerTile + dec_local;
        if (dloc < kDecPerWarp && dec < D) {
            probs[dec * kSlots + slot] = has_card ? __expf(logit - m) / z : 0.0f;
            if (logits) logits[dec * kSlots + slot] = has_card ? logit : 0.0f;
        }
        card_bits = next_card;
        buf ^= 1u;
        pc.mark(8);
    }
#ifdef NIMMT_PHASE_CLOCKS
    {
        static const char* const names[] = {"", "sync+fence", "mma1 wait", "epilogue1", "sync", "mma2 wait", "epilogue2", "", "softmax+out"};
        pc.print(names, 9);
    }
#endif
    tc_fence_before_sync();
    __syncthreads();
    if (threadIdx.x < 32) tmem_dealloc(tmem_slot, kProbTmemCols);
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
    static int occ_cache[kMaxDevices];
    blocks_per_sm_cached(k_policy_probs, kTileRows * kProbGroups, (int)kProbSmemBytes, occ_cache);   // per-device opt-in
    const int num_sms = device_sms(current_device());
    const int64_t tiles = (num_decisions + kDecPerTile - 1) / kDecPerTile, ctas = (tiles + kProbGroups - 1) / kProbGroups;
    const unsigned blocks = (unsigned)(ctas < num_sms ? ctas : num_sms);   // persistent: one 4-group CTA per SM
    k_policy_probs<<<blocks, kTileRows * kProbGroups, kProbSmemBytes, (cudaStream_t)stream>>>(obs, num_decisions, static_cast<const uint8_t*>(weights), probs, logits);
    return check_launch();
}

}  // extern "C"
