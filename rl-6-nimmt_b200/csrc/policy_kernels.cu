// policy_kernels.cu — Alpha0.5 leaf evaluation on the 5th-generation tensor cores.
//
// Replaces PolicyMCSAgent._compute_policy (agents/mcts.py:219-228) = SechsNimmtStateNormalization
// (utils/preprocessing.py:12-57) + MultiHeadedMLP(48, (100, 100), (1,)) with ReLU
// (utils/nets.py:100-132) + softmax over the legal cards of one decision.
//
// One CTA = one 128-row tile = 12 decisions x 10 card slots (row = [card | obs47], raw integer
// features; the affine normalisation is folded into layer 1 when the weights are packed).
//   layer 1: [128 x 48] x [48 x 112]   3 x tcgen05.mma (K = 16 each), A and B bf16 in shared memory,
//            fp32 accumulator in TMEM; epilogue tcgen05.ld -> + b1 -> ReLU -> bf16 -> shared memory
//   layer 2: [128 x 112] x [112 x 112] 7 x tcgen05.mma; epilogue + b2 -> ReLU, and
//   layer 3 (100 -> 1) as an fp32 dot product in the same epilogue (N = 1 is not a tensor-core shape)
//   softmax over each decision's legal slots, probabilities out.
// Hidden width 100 is padded to 112 (tcgen05 N granularity 16 at M = 128); padded weights are zero.
// Features are built on chip from int8 observations, so the GEMMs never read activations from HBM.
#include "policy_tile.cuh"

namespace nimmt {

// grid-stride over tiles of 12 decisions; every 128-thread group of the CTA takes its own tiles.
// obs: int8 [D][47]; probs: float [D][10] (0 for empty slots); logits (optional): float [D][10].
// kProbGroups groups share one copy of the weights in shared memory and each own 128 TMEM columns, an
// mbarrier and a named barrier, so one group's epilogue (TMEM -> registers -> bf16 -> shared memory)
// runs under the other groups' MMAs.
constexpr int kProbGroups = 4;
__global__ void __launch_bounds__(kTileRows * kProbGroups, 1)
k_policy_probs(const int8_t* __restrict__ obs, int64_t D, const uint8_t* __restrict__ blob, float* __restrict__ probs, float* __restrict__ logits) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ uint64_t bars[kProbGroups];
    __shared__ uint32_t tmem_slot;

    for (uint32_t i = threadIdx.x * 16; i < kBlobBytes; i += blockDim.x * 16)
        *reinterpret_cast<uint4*>(smem + kSmemBlob + i) = *reinterpret_cast<const uint4*>(blob + i);
    if (threadIdx.x < kProbGroups) mbar_init(&bars[threadIdx.x], 1);
    if (threadIdx.x == 0) fence_barrier_init();
    if (threadIdx.x < 32) tmem_alloc(&tmem_slot, kTmemColsPerGroup * kProbGroups);
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const int group = threadIdx.x / kTileRows, tid = threadIdx.x % kTileRows, bar_id = 1 + group;
    const uint32_t tmem_base = tmem_slot + group * kTmemColsPerGroup;
    uint8_t* gbuf = smem + kSmemGroups + group * kGroupBytes;
    uint32_t phase = 0;
    PhaseClock pc;
    float* tile_logits = reinterpret_cast<float*>(gbuf + kGLogits);
    int8_t* tile_obs = reinterpret_cast<int8_t*>(gbuf + kGObs);

    const int64_t num_tiles = (D + kDecPerTile - 1) / kDecPerTile;
    for (int64_t tile = (int64_t)blockIdx.x * kProbGroups + group; tile < num_tiles; tile += (int64_t)gridDim.x * kProbGroups) {
        const int dec_local = tid / kSlots, slot = tid % kSlots;
        const int64_t dec = tile * kDecPerTile + dec_local;
        const bool in_range = tid < kDecPerTile * kSlots && dec < D;
        // stage the tile's 12 x 47 observation bytes with coalesced 4-byte loads (564 B, 4-byte aligned)
        {
            const int64_t first_byte = tile * (kDecPerTile * kObs), total_bytes = D * kObs;
            for (int wd = tid; wd < kDecPerTile * kObs / 4; wd += kTileRows) {
                const int64_t byte = first_byte + 4 * wd;
                uint32_t v = 0;
                if (byte + 4 <= total_bytes) v = *reinterpret_cast<const uint32_t*>(obs + byte);
                else for (int i = 0; i < 4; ++i) if (byte + i < total_bytes) v |= (uint32_t)(uint8_t)obs[byte + i] << (8 * i);
                reinterpret_cast<uint32_t*>(tile_obs)[wd] = v;
            }
        }
        group_sync(bar_id);
        const int8_t* o = tile_obs + dec_local * kObs;
        const int card = in_range ? o[slot] : -1;      // hand slot: the candidate card, -1 if empty (env.py:209-210)
        const bool live = in_range && card >= 0;
        write_feature_row(gbuf, tid, [&](int k) -> float { return live ? (float)(k == 0 ? card : o[k - 1]) : 0.0f; });

        const float logit = mlp_tile(smem + kSmemBlob, gbuf, tmem_base, &bars[group], phase, tid, bar_id, pc);

        tile_logits[tid] = logit;
        group_sync(bar_id);
        if (in_range) {
            // softmax over the decision's legal slots (agents/mcts.py:207,227: Softmax(dim=0) over the rows)
            float m = -INFINITY;
            for (int s = 0; s < kSlots; ++s)
                if (o[s] >= 0) m = fmaxf(m, tile_logits[dec_local * kSlots + s]);
            float z = 0.0f;
            for (int s = 0; s < kSlots; ++s)
                if (o[s] >= 0) z += __expf(tile_logits[dec_local * kSlots + s] - m);
            probs[dec * kSlots + slot] = card >= 0 ? __expf(logit - m) / z : 0.0f;
            if (logits) logits[dec * kSlots + slot] = card >= 0 ? logit : 0.0f;
        }
        group_sync(bar_id);   // tile_logits, tile_obs and the A operands are free again
    }
    tc_fence_before_sync();
    __syncthreads();
    if (threadIdx.x < 32) tmem_dealloc(tmem_slot, kTmemColsPerGroup * kProbGroups);
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
    float* b1p = reinterpret_cast<float*>(blob + kOffB1);
    float* b2p = reinterpret_cast<float*>(blob + kOffB2);
    float* w3p = reinterpret_cast<float*>(blob + kOffW3);
    for (int n = 0; n < kHid; ++n) {
        double acc = b1[n];
        for (int k = 0; k < kIn; ++k) {
            const double w = w1[n * kIn + k];
            acc += w * shift[k];
            *reinterpret_cast<uint16_t*>(blob + kOffW1 + canon_off(n, k, kInChunks)) = float_to_bf16_rne((float)(w * scale[k]));
        }
        b1p[n] = (float)acc;
        for (int k = 0; k < kHid; ++k)
            *reinterpret_cast<uint16_t*>(blob + kOffW2 + canon_off(n, k, kHidChunks)) = float_to_bf16_rne(w2[n * kHid + k]);
        b2p[n] = b2[n];
        w3p[n] = w3[n];
    }
    *reinterpret_cast<float*>(blob + kOffB3) = b3;
    return NIMMT_OK;
}

int nimmt_policy_probs(const int8_t* obs, int64_t num_decisions, const void* weights, float* probs, float* logits, void* stream) {
    if (!obs || !weights || !probs || num_decisions < 0) return NIMMT_E_BADARG;
    if (!aligned16(weights) || !aligned16(probs)) return NIMMT_E_ALIGN;
    if (num_decisions == 0) return NIMMT_OK;
    static int num_sms = 0;
    if (num_sms == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
        cudaFuncSetAttribute(k_policy_probs, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)policy_smem_bytes(kProbGroups));
    }
    const int64_t tiles = (num_decisions + kDecPerTile - 1) / kDecPerTile, ctas = (tiles + kProbGroups - 1) / kProbGroups;
    const unsigned blocks = (unsigned)(ctas < num_sms ? ctas : num_sms);   // persistent: one 4-group CTA per SM
    k_policy_probs<<<blocks, kTileRows * kProbGroups, policy_smem_bytes(kProbGroups), (cudaStream_t)stream>>>(obs, num_decisions, static_cast<const uint8_t*>(weights), probs, logits);
    return check_launch();
}

}  // extern "C"
