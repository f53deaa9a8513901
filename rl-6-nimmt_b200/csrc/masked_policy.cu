// masked_policy.cu — the state-only 47 -> 100 -> 100 -> 104 nets of the model-free agents on tcgen05:
// MaskedReinforceAgent.forward up to the sampling (agents/policy.py:45-60; a DQN's Q-values have the same shape,
// agents/dqn.py:219-230) = SechsNimmtStateNormalization(action=False) (utils/preprocessing.py:12-57) +
// MultiHeadedMLP(47, (100, 100), (104,)) (utils/nets.py:100-132) + softmax over the cards in hand.
//
// One tile = 128 DECISIONS (one row per seat, not per card as in policy_kernels.cu):
//   layer 1: [128 x 64] x [64 x 112]     features = [0 | obs47 | 1 1 0 ..] raw integers, normalisation and b1 folded into the weights
//   layer 2: [128 x 112] x [112 x 112]   b2 rides on the constant-1 units 100, 101 of layer 1
//   layer 3: [128 x 112] x [112 x 112]   one output column per card (104 of 112 used); b3 rides on units 100, 101 of layer 2,
//                                        which the packed W2 holds at the constant 1
// all three as tcgen05.mma (kind::f16, fp32 accumulator in TMEM) issued by one thread; layer 1 reads both operands from shared
// memory, layers 2 and 3 read A from TENSOR memory: epilogues 1 and 2 are policy_tile.cuh's relu_to_operand (TMEM -> ReLU ->
// bf16 pairs -> TMEM), so the activations never pass through shared memory; epilogue 3 stages each row's 104
// logits in shared memory, and the row's own thread gathers the <= 10 cards of its hand and normalises them.
#include "policy_tile.cuh"

namespace nimmt {

constexpr uint32_t kMOffW1 = 0, kMOffW2 = kMOffW1 + kW1Bytes, kMOffW3 = kMOffW2 + kW2Bytes;
constexpr uint32_t kMaskedBlobBytes = kMOffW3 + kW2Bytes;                        // 64512
constexpr int kCardsOut = 104, kLogitStride = 33;                                // a 32-column window per row; odd stride: a warp's rows hit 32 banks
constexpr uint32_t kMSmemBlob = 0, kMSmemA1 = (kMaskedBlobBytes + 127) / 128 * 128;
constexpr uint32_t kMSmemObs = kMSmemA1 + kA1Bytes;                              // int8 [128][47] (+ pad to words)
constexpr uint32_t kMaskedTmemCols = 256;                                        // 168 used (policy_tile.cuh::kTmemColsPerGroup)
constexpr uint32_t kMSmemLogits = (kMSmemObs + kTileRows * kObs + 127) / 128 * 128;
constexpr uint32_t kMaskedSmemBytes = kMSmemLogits + kTileRows * kLogitStride * 4;

__global__ void __launch_bounds__(kTileRows, 1)
k_masked_probs(const int8_t* __restrict__ obs, int64_t D, const uint8_t* __restrict__ blob, float* __restrict__ probs, float* __restrict__ logits_out) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5;

    for (uint32_t i = tid * 16; i < kMaskedBlobBytes; i += kTileRows * 16)
        *reinterpret_cast<uint4*>(smem + kMSmemBlob + i) = *reinterpret_cast<const uint4*>(blob + i);
    if (tid == 0) {
        mbar_init(&bar, 1);
        fence_barrier_init();
    }
    if (tid < 32) tmem_alloc(&tmem_slot, kMaskedTmemCols);
    uint8_t* a1_row = smem + kMSmemA1 + canon_off(tid, 0, kInChunks);
    // once: the constant chunks of this thread's row (features 48..63 = 1 1 0 ..)
    *reinterpret_cast<uint4*>(a1_row + (kBiasCol / 8) * 128) = make_uint4(0x3F803F80u, 0u, 0u, 0u);
    *reinterpret_cast<uint4*>(a1_row + (kBiasCol / 8 + 1) * 128) = make_uint4(0u, 0u, 0u, 0u);
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = tmem_slot;
    const uint32_t lane_taddr = tmem_base + ((uint32_t)(warp * 32) << 16);
    const uint32_t a1 = smem_u32(smem + kMSmemA1);
    const uint32_t w1 = smem_u32(smem + kMOffW1), w2 = smem_u32(smem + kMOffW2), w3 = smem_u32(smem + kMOffW3);
    constexpr uint32_t idesc = umma_idesc_bf16(kTileRows, kHidPad);
    uint32_t phase = 0;
    int8_t* tile_obs = reinterpret_cast<int8_t*>(smem + kMSmemObs);
    float* my_logits = reinterpret_cast<float*>(smem + kMSmemLogits) + tid * kLogitStride;

    const int64_t total_bytes = D * kObs, num_tiles = (D + kTileRows - 1) / kTileRows;
    for (int64_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        // the tile's 128 x 47 observation bytes as coalesced 32-bit loads (the base is 4-byte aligned, a tile is 6016 bytes)
        const int64_t tile_byte0 = tile * (kTileRows * kObs);
        for (int wd = tid; wd < kTileRows * kObs / 4; wd += kTileRows) {
            const int64_t byte = tile_byte0 + 4 * wd;
            uint32_t v = 0;
            if (byte + 4 <= total_bytes) v = *reinterpret_cast<const uint32_t*>(obs + byte);
            else for (int i = 0; i < 4; ++i) if (byte + i < total_bytes) v |= (uint32_t)(uint8_t)obs[byte + i] << (8 * i);
            reinterpret_cast<uint32_t*>(tile_obs)[wd] = v;
        }
        __syncthreads();
        // this thread's row: [0 | obs 0..46] as bf16, six 16-byte chunks of the layer-1 A operand
        const int8_t* o = tile_obs + tid * kObs - 1;     // o[k] = feature k (k >= 1)
#pragma unroll
        for (int c = 0; c < kFeatChunks; ++c) {
            uint32_t p[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) p[i] = bf16x2_bits(c == 0 && i == 0 ? 0 : o[8 * c + 2 * i], o[8 * c + 2 * i + 1]);
            *reinterpret_cast<uint4*>(a1_row + c * 128) = make_uint4(p[0], p[1], p[2], p[3]);
        }
        // ---- layer 1 ----
        fence_async_smem();
        tc_fence_before_sync();
        __syncthreads();
        if (tid == 0) {
            tc_fence_after_sync();
#pragma unroll
            for (int ks = 0; ks < kInPad / 16; ++ks)
                umma_bf16(tmem_base, umma_desc(a1 + ks * 256, 128, kInChunks * 128), umma_desc(w1 + ks * 256, 128, kInChunks * 128), idesc, ks > 0);
            umma_commit(&bar);
        }
        mbar_wait_mma(&bar, phase);
        phase ^= 1u;
        tc_fence_after_sync();
        relu_to_operand(lane_taddr, lane_taddr);     // in place: operand columns [0, 56) over accumulator columns already read
        // ---- layers 2 and 3: the same [128 x 112] x [112 x 112] product with different weights, A from tensor memory
        // (columns [0, 56)), accumulator in columns [56, 168) ----
#pragma unroll
        for (int layer = 2; layer <= 3; ++layer) {
            tc_fence_before_sync();
            __syncthreads();          // every lane's accumulator has been read and its operand row written
            if (tid == 0) {
                tc_fence_after_sync();
                const uint32_t w = layer == 2 ? w2 : w3;
#pragma unroll
                for (int ks = 0; ks < kHidPad / 16; ++ks)
                    umma_bf16_ts(tmem_base + kTmemAcc2, tmem_base + ks * 8, umma_desc(w + ks * 256, 128, kHidChunks * 128), idesc, ks > 0);
                umma_commit(&bar);
            }
            mbar_wait_mma(&bar, phase);
            phase ^= 1u;
            tc_fence_after_sync();
            if (layer == 2) relu_to_operand(lane_taddr + kTmemAcc2, lane_taddr);   // the MMAs that read the old operand have completed
        }
        // ---- epilogue 3: this row's 104 logits pass through a 32-column window of shared memory (the only way to index them by
        // card: a register file has no dynamic index), four passes; after each the cards of the hand that fall into the window
        // are picked up.  The window is the thread's own row: no barrier between its stores and its loads. ----
        const int64_t d = tile * kTileRows + tid;
        float l[kSlots];
        int cards[kSlots];
#pragma unroll
        for (int s = 0; s < kSlots; ++s) {
            cards[s] = tile_obs[tid * kObs + s];                           // hand slot s: a card, or -1 (env.py:209-210)
            l[s] = -INFINITY;
        }
#pragma unroll
        for (int pass = 0; pass < 4; ++pass) {
            if (pass < 3) {
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    uint32_t v[16];
                    tmem_ld16(lane_taddr + kTmemAcc2 + pass * 32 + c * 16, v);
                    tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < 16; ++i) my_logits[c * 16 + i] = __uint_as_float(v[i]);
                }
            } else {
                uint32_t v[8];
                tmem_ld8(lane_taddr + kTmemAcc2 + 96, v);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 8; ++i) my_logits[i] = __uint_as_float(v[i]);
            }
            const int width = pass < 3 ? 32 : kCardsOut - 96;
#pragma unroll
            for (int s = 0; s < kSlots; ++s) {
                const int off = cards[s] - pass * 32;
                if (off >= 0 && off < width) l[s] = my_logits[off];
            }
        }
        tc_fence_before_sync();
        if (d < D) {
            float m = -INFINITY;
#pragma unroll
            for (int s = 0; s < kSlots; ++s) m = fmaxf(m, l[s]);
            float e[kSlots], z = 0.0f;
#pragma unroll
            for (int s = 0; s < kSlots; ++s) {
                e[s] = l[s] > -INFINITY ? __expf(l[s] - m) : 0.0f;
                z += e[s];
            }
#pragma unroll
            for (int s = 0; s < kSlots; ++s) {
                probs[d * kSlots + s] = z > 0.0f ? e[s] / z : 0.0f;
                if (logits_out) logits_out[d * kSlots + s] = l[s] > -INFINITY ? l[s] : 0.0f;
            }
        }
        __syncthreads();   // tile_obs is rewritten by the next tile
    }
    tc_fence_before_sync();
    __syncthreads();
    if (tid < 32) tmem_dealloc(tmem_base, kMaskedTmemCols);
}

}  // namespace nimmt

using namespace nimmt;

extern "C" {

size_t nimmt_masked_weights_bytes(void) { return kMaskedBlobBytes; }

int nimmt_masked_pack_weights(const float* w1, const float* b1, const float* w2, const float* b2, const float* w3, const float* b3, void* blob_host) {
    if (!w1 || !b1 || !w2 || !b2 || !w3 || !b3 || !blob_host) return NIMMT_E_BADARG;
    uint8_t* blob = static_cast<uint8_t*>(blob_host);
    memset(blob, 0, kMaskedBlobBytes);
    double scale[kIn], shift[kIn];
    for (const Segment& s : kSegments)
        for (int k = s.begin; k < s.end; ++k) {
            scale[k] = 2.0 / ((double)s.hi - s.lo);                 // preprocessing.py:56-57 with out range [-1, 1]
            shift[k] = -1.0 - 2.0 * s.lo / ((double)s.hi - s.lo);
        }
    auto put = [&](uint32_t base, int n, int k, int kchunks, uint16_t v) { *reinterpret_cast<uint16_t*>(blob + base + canon_off(n, k, kchunks)) = v; };
    auto split = [](float b, uint16_t& hi, uint16_t& lo) {
        hi = float_to_bf16_rne(b);
        lo = float_to_bf16_rne(b - bf16_to_float(hi));
    };
    uint16_t hi, lo;
    for (int n = 0; n < kHid; ++n) {
        double acc = b1[n];
        for (int k = 1; k < kIn; ++k) {                              // feature k = observation entry k - 1; feature 0 (the candidate card) does not exist here
            const double w = w1[n * kObs + (k - 1)];
            acc += w * shift[k];
            put(kMOffW1, n, k, kInChunks, float_to_bf16_rne((float)(w * scale[k])));
        }
        split((float)acc, hi, lo);
        put(kMOffW1, n, kBiasCol, kInChunks, hi);
        put(kMOffW1, n, kBiasCol + 1, kInChunks, lo);
        for (int k = 0; k < kHid; ++k) put(kMOffW2, n, k, kHidChunks, float_to_bf16_rne(w2[n * kHid + k]));
        split(b2[n], hi, lo);
        put(kMOffW2, n, kOneUnit, kHidChunks, hi);
        put(kMOffW2, n, kOneUnit + 1, kHidChunks, lo);
    }
    for (int t = 0; t < 2; ++t) {
        put(kMOffW1, kOneUnit + t, kBiasCol, kInChunks, 0x3F80);      // layer 1's units 100, 101: relu(1 * 1) = 1, the inputs that carry b2
        put(kMOffW2, kOneUnit + t, kOneUnit, kHidChunks, 0x3F80);     // layer 2's units 100, 101: the same constant, carrying b3
    }
    for (int c = 0; c < kCardsOut; ++c) {
        for (int k = 0; k < kHid; ++k) put(kMOffW3, c, k, kHidChunks, float_to_bf16_rne(w3[c * kHid + k]));
        split(b3[c], hi, lo);
        put(kMOffW3, c, kOneUnit, kHidChunks, hi);
        put(kMOffW3, c, kOneUnit + 1, kHidChunks, lo);
    }
    return NIMMT_OK;
}

int nimmt_masked_probs(const int8_t* obs, int64_t num_decisions, const void* weights, float* probs, float* logits, void* stream) {
    if (!obs || !weights || !probs || num_decisions < 0) return NIMMT_E_BADARG;
    if (!aligned16(weights) || (reinterpret_cast<uintptr_t>(obs) & 3u)) return NIMMT_E_ALIGN;
    if (num_decisions == 0) return NIMMT_OK;
    static int occ_cache[kMaxDevices];
    blocks_per_sm_cached(k_masked_probs, kTileRows, (int)kMaskedSmemBytes, occ_cache);   // per-device shared-memory opt-in
    const int64_t tiles = (num_decisions + kTileRows - 1) / kTileRows;
    // two CTAs per SM: 256 tensor-memory columns and 101 KB of shared memory each (the logits pass through a 32-column window
    // instead of a 104-column row for exactly this).  The occupancy API answers 1 for this kernel; the hardware runs two — measured:
    // 0.300 ms per 2^20 decisions with 148 CTAs, 0.200 ms with 296, 0.255 ms with 444 (a second wave).
    const int64_t slots = 2 * (int64_t)device_sms(current_device());
    const unsigned blocks = (unsigned)(tiles < slots ? tiles : slots);
    k_masked_probs<<<blocks, kTileRows, kMaskedSmemBytes, (cudaStream_t)stream>>>(obs, num_decisions, static_cast<const uint8_t*>(weights), probs, logits);
    return check_launch();
}

}  // extern "C"
