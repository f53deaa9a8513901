// env_step.cu — batched env.step (SechsNimmtEnv.step, env.py:64-77), uniformly random actions (DrunkHamster.forward,
// agents/random.py:8-10) and the two fused; the multi-turn, free-row-choice and packed-transfer variants of the step.
// One game per lane; the throughput kernel (k_step_tiles) is warp-specialised and HBM-bandwidth bound.
#include <cstdlib>

#include "abi_common.cuh"
#include "step_tile.cuh"
#include "tma.cuh"

namespace nimmt {

// ------------------------------------------------------------------------------------------
// k_step — SechsNimmtEnv.step (env.py:64-77) without the observation rebuild.
// Reads (12 P + 24) + P bytes and writes (4 P + 24) + P + 1 (+1) bytes per game: the dealt cards are immutable.
// kRandom: the actions are drawn in-kernel (DrunkHamster, agents/random.py:8-10) instead of read.
// ------------------------------------------------------------------------------------------
template <int P, bool kRandom>
__global__ void __launch_bounds__(kStepThreads)
k_step(StateView s, const uint8_t* __restrict__ actions_in, uint8_t* __restrict__ actions_out, int8_t* __restrict__ rewards,
       uint8_t* __restrict__ done, uint8_t* __restrict__ illegal, uint64_t seed, uint32_t turn, uint64_t game0,
       int64_t first_game, const uint8_t* __restrict__ rows = nullptr) {
    __shared__ uint8_t values[128];
    const int64_t g_raw = first_game + (int64_t)blockIdx.x * kStepThreads + threadIdx.x;
    const bool valid = g_raw < s.B;
    const int64_t g = valid ? g_raw : s.B - 1;  // tail threads shadow the last game and write nothing

    // Issue every global load before the value table is staged, so that the block barrier below
    // overlaps the memory latency instead of preceding it.
    RawGame<P> raw;
    int act[P];
    if constexpr (!kRandom) load_bytes<P>(actions_in, g, act);
    load_raw<P>(s, g, raw);
    stage_card_values(values);
    __syncthreads();
    if (!valid) return;
    GameRec<P> gm;
    unpack_raw<P>(raw, gm);
    if constexpr (kRandom) {
        random_actions_game<P>(gm, seed, game0 + (uint64_t)g, turn, act);
        if (actions_out) store_bytes<P>(actions_out, g, act);
    }

    int penalty[P], choice[P];
    if (rows) load_bytes<P>(rows, g, choice);   // free-row-choice mode (env.py:156 TODO)
    const bool legal = step_game<P>(gm, act, values, penalty, rows ? choice : nullptr);
    if (legal) store_step<P>(s, g, gm);

    int rew[P];
#pragma unroll
    for (int p = 0; p < P; ++p) rew[p] = -penalty[p];
    store_bytes<P>(reinterpret_cast<uint8_t*>(rewards), g, rew);
    done[g] = game_done<P>(gm);
    if (illegal) illegal[g] = !legal;
}

// ------------------------------------------------------------------------------------------
// k_step_tiles — the throughput path of env.step.  Warp-specialised: W consumer warps and ONE producer warp per block.
//
// Work unit = a GROUP of W consecutive 32-game tiles.  Tile records are contiguous in HBM (game.cuh::StateView), so a
// group arrives with two 1-D bulk copies — W tile records in one run, W x 32 x P action bytes in another — and leaves
// with W (the tiles' mutable blocks).  All copies are issued by the producer warp with warp-uniform control flow and
// operands, so address arithmetic and mbarrier bookkeeping live on the uniform datapath and cost the consumer warps
// nothing (the first version of this kernel had every warp's lane 0 issue its own tile's copies: a fifth of all
// warp-instructions).  S stages per block:
//     producer:  wait done[s] -> bulk-store the W mutable blocks of stage s -> wait until the engine has read them ->
//                bulk-load group i + S into stage s (full[s])
//     consumer w: wait full[s] -> step tile w IN PLACE in shared memory (one game per lane) -> fence -> arrive on done[s]
//                -> store its games' reward bytes and done / illegal flags straight from registers (coalesced)
// The dealt cards never travel back: a step writes 4 bytes per player + the 24-byte row record.
// ------------------------------------------------------------------------------------------
// Tuning knobs (profiles/tools/build_variants.sh builds and times the alternatives): consumer warps = tiles per group, and
// pipeline stages, for small (P <= 5) and large tables.
#ifndef NIMMT_STEP_WARPS_SMALL
#define NIMMT_STEP_WARPS_SMALL 8
#endif
#ifndef NIMMT_STEP_WARPS_LARGE
#define NIMMT_STEP_WARPS_LARGE 4
#endif
#ifndef NIMMT_STEP_STAGES_SMALL
#define NIMMT_STEP_STAGES_SMALL 3
#endif
#ifndef NIMMT_STEP_STAGES_LARGE
#define NIMMT_STEP_STAGES_LARGE 2
#endif
#ifndef NIMMT_STEP_STAGES_SMALL_DEEP
#define NIMMT_STEP_STAGES_SMALL_DEEP 4
#endif
// Measured on the B200 (profiles/r02_step_shapes.txt): P = 4 at 2^20 games W8S3 24.1 us, W8S4 24.4, W16S2 26.4, W6S3 24.0, W4S3
// 24.4; at 2^24 games W8S3 343 us, W8S4 326, W16S2 326; P = 10 at 2^20 / 2^24: W4S3 48.3 / 686, W4S2 47.7 / 662, W6S2 50.7 / 726.
// A deeper pipeline pays once the batch is far beyond one wave of blocks, so the stage count is chosen per launch.
template <int P>
struct StepShape {
    static constexpr int kWarps = P <= 5 ? NIMMT_STEP_WARPS_SMALL : NIMMT_STEP_WARPS_LARGE;
    static constexpr int kStages = P <= 5 ? NIMMT_STEP_STAGES_SMALL : NIMMT_STEP_STAGES_LARGE;
    static constexpr int kStagesDeep = P <= 5 ? NIMMT_STEP_STAGES_SMALL_DEEP : NIMMT_STEP_STAGES_LARGE;   // batches >= kDeepTiles tiles
};
constexpr int64_t kDeepTiles = 1 << 17;   // 2^22 games

template <int P, int W, bool kChoice = false, bool kPacked = false>
struct StageLayout {
    using L = TileLayout<P>;
    static constexpr int kTiles = 0;                                          // W tile records
    static constexpr int kActions = W * L::kTileBytes;                        // W x 32 x P action bytes (kPacked: 4-bit slots)
    static constexpr int kActTile = kPacked ? kTileGames * packed_action_bytes<P>() : L::kActBytes;   // action bytes of one tile
    static constexpr int kChoices = kActions + W * kActTile;                  // kChoice: W x 32 x P row-choice bytes
    static constexpr int kResults = kChoices;                                 // kPacked: W x 32 bit-packed result records (bulk-stored)
    static constexpr int kResTile = kTileGames * packed_result_bytes<P>();
    static constexpr int kBytes = kChoices + (kChoice ? W * L::kActBytes : 0) + (kPacked ? W * kResTile : 0);
    static constexpr int kStride = (kBytes + 127) / 128 * 128;
    static_assert(kActTile % 16 == 0 && kResTile % 16 == 0, "bulk copies need 16-byte alignment");
};

// kMany: `turns` consecutive env steps per launch (nimmt_step_many / nimmt_step_random_many).  A group's tiles stay in shared
// memory for all of them: the state travels once per launch instead of once per step, the per-turn inputs (actions [T][B][P])
// arrive with one more bulk copy per turn, and the per-turn outputs (rewards [T][B][P], done / illegal [T][B]) leave from
// registers as before.  The stage size depends on `turns`, so the stage stride is a run-time value in this variant.
template <int P, int W, bool kChoice>
__host__ __device__ constexpr uint32_t stage_stride_many(int turns) {
    return (uint32_t)((StageLayout<P, W, kChoice>::kActions + turns * W * TileLayout<P>::kActBytes + 127) / 128 * 128);
}

// kPacked: the compact transfer format (step_tile.cuh): `actions` holds 4-bit hand slots, and `rewards` receives ONE bit-packed
// record per game (bull heads, done, illegal) — staged in shared memory by the lanes and bulk-stored by the producer.
template <int P, bool kRandom, bool kChoice = false, bool kMany = false, bool kPacked = false, int S = StepShape<P>::kStages>
__global__ void __launch_bounds__((StepShape<P>::kWarps + 1) * 32)
k_step_tiles(StateView s, const uint8_t* __restrict__ actions, uint8_t* __restrict__ actions_out, int8_t* __restrict__ rewards,
             uint8_t* __restrict__ done, uint8_t* __restrict__ illegal, int num_tiles, uint64_t seed, uint32_t turn, uint64_t game0,
             const uint8_t* __restrict__ rows = nullptr, int turns = 1) {
    constexpr int W = StepShape<P>::kWarps;
    using L = TileLayout<P>;
    using G = StageLayout<P, W, kChoice, kPacked>;
    static_assert(!(kMany && kChoice), "the multi-turn launch carries no row choices");
    static_assert(!kPacked || (!kRandom && !kChoice && !kMany), "the packed format exists for the plain step only");
    const uint32_t kStride = kMany ? stage_stride_many<P, W, kChoice>(kRandom ? 0 : turns) : (uint32_t)G::kStride;   // compile-time unless kMany
    const int64_t turn_bytes = (int64_t)num_tiles * L::kActBytes;   // kMany: B * P, the stride between the turns of the [T][B][P] arrays
    extern __shared__ __align__(128) uint8_t stage_smem[];   // S x kStride
    __shared__ uint64_t full[S], computed[S];
    __shared__ uint8_t values5[128];
    __shared__ uint4 keys_w[W * 32], keys_u[W * 32];          // each lane's row keys, indexable
    __shared__ uint32_t sel8[kRandom ? 256 : 1];              // kRandom: the slot-selection table (handrec.cuh::rec_select_slot)

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int num_groups = (num_tiles + W - 1) / W;
    // Programmatic dependent launch: this grid may have been launched while the previous kernel of the stream was still draining
    // (launch_step_tiles sets the attribute), and lets the next one do the same.  Everything up to griddepcontrol.wait — barrier
    // initialisation, the tables — touches no memory another kernel writes and overlaps the previous kernel's tail.
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    if (threadIdx.x == 0) {
#pragma unroll
        for (int i = 0; i < S; ++i) {
            mbar_init(&full[i], 1);
            mbar_init(&computed[i], W);
        }
        fence_barrier_init();
    }
    stage_card_values5(values5);
    if constexpr (kRandom) stage_select8(sel8);
    __syncthreads();   // the only block-wide barrier: tables + barrier init
    asm volatile("griddepcontrol.wait;" ::: "memory");   // the previous kernel's writes (this state, the outputs' last readers) are complete and visible
    const uint32_t full_a = smem_u32(full), computed_a = smem_u32(computed), stage_a = smem_u32(stage_smem);

    if (warp == W) {
        // ---------------- producer warp: every copy of the block; warp-uniform control flow, one elected lane issues ----------------
        auto load_group = [&](int group, int stage) {
            const int first = group * W;
            const uint32_t n = (uint32_t)min(W, num_tiles - first);       // the last group may be short
            const uint32_t buf = stage_a + (uint32_t)stage * kStride, bar = full_a + 8u * (uint32_t)stage;
            const uint32_t action_loads = kRandom ? 0u : (kMany ? (uint32_t)turns : 1u);
            mbar_arrive_expect_tx_a(bar, n * ((uint32_t)L::kTileBytes + action_loads * (uint32_t)G::kActTile + (kChoice ? (uint32_t)L::kActBytes : 0u)));
            bulk_load_a(buf, s.tile_ptr(first), n * L::kTileBytes, bar);
            if constexpr (!kRandom) {
                for (uint32_t t = 0; t < action_loads; ++t)
                    bulk_load_a(buf + G::kActions + t * (uint32_t)(W * G::kActTile), actions + (int64_t)t * turn_bytes + (int64_t)first * G::kActTile,
                                n * G::kActTile, bar);
            }
            if constexpr (kChoice) bulk_load_a(buf + G::kChoices, rows + (int64_t)first * L::kActBytes, n * L::kActBytes, bar);
        };
        const bool issuer = elect_one() != 0u;
        {
            int stage = 0;
            for (int group = blockIdx.x; stage < S && group < num_groups; group += gridDim.x, ++stage)
                if (issuer) load_group(group, stage);
        }
        int stage = 0;
        uint32_t phase = 0;
        for (int group = blockIdx.x; group < num_groups; group += gridDim.x) {
            mbar_wait_a(computed_a + 8u * (uint32_t)stage, phase);
            if (issuer) {
                const int first = group * W;
                const int n = min(W, num_tiles - first);
                const uint32_t buf = stage_a + (uint32_t)stage * kStride + L::kMeta;
                for (int t = 0; t < n; ++t) bulk_store_a(s.mut_ptr(first + t), buf + (uint32_t)t * L::kTileBytes, L::kMutBytes);
                if constexpr (kPacked)
                    bulk_store_a(reinterpret_cast<uint8_t*>(rewards) + (int64_t)first * G::kResTile, stage_a + (uint32_t)stage * kStride + G::kResults,
                                 (uint32_t)n * G::kResTile);
                bulk_commit();
                const int next = group + S * (int)gridDim.x;
                if (next < num_groups) {
                    bulk_wait_read<0>();   // the engine has read the stage: it may be refilled
                    load_group(next, stage);
                }
            }
            if (++stage == S) { stage = 0; phase ^= 1u; }
        }
        if (issuer) bulk_wait_all();   // stores must land before the block's shared memory is released
        return;
    }

    // ---------------- consumer warps: warp w steps tile w of every group ----------------
    uint32_t* kw = reinterpret_cast<uint32_t*>(&keys_w[threadIdx.x]);
    uint32_t* ku = reinterpret_cast<uint32_t*>(&keys_u[threadIdx.x]);
    uint8_t* my_tile = stage_smem + warp * L::kTileBytes;
    const uint32_t lane_game = (uint32_t)(warp * kTileGames + lane);       // this lane's game within a group
    int stage = 0;
    uint32_t phase = 0;
    for (int group = blockIdx.x; group < num_groups; group += gridDim.x) {
        uint8_t* tile = my_tile + stage * kStride;
        mbar_wait_a(full_a + 8u * (uint32_t)stage, phase);
        if (group * W + warp < num_tiles) {
            const int64_t g0 = (int64_t)group * (W * kTileGames);           // warp-uniform: lives on the uniform datapath
            uint8_t* stage_base = stage_smem + stage * kStride;
            if constexpr (kPacked) {
                step_lane<P, false, false, true>(tile, stage_base + G::kActions + warp * G::kActTile, lane, values5, kw, ku,
                                                 stage_base + G::kResults + (warp * kTileGames + lane) * packed_result_bytes<P>(), nullptr, nullptr, nullptr, 0, 0, 0);
            } else if constexpr (!kMany) {
                step_lane<P, kRandom, kChoice>(tile, stage_base + G::kActions + warp * L::kActBytes, lane, values5, kw, ku,
                                               reinterpret_cast<uint8_t*>(rewards) + g0 * P + lane_game * P, done + g0 + lane_game,
                                               illegal ? illegal + g0 + lane_game : nullptr, actions_out ? actions_out + g0 * P + lane_game * P : nullptr,
                                               seed, game0 + (uint64_t)g0 + lane_game, turn, stage_base + G::kChoices + warp * L::kActBytes, sel8);
            } else {
                const int64_t games = (int64_t)num_tiles * kTileGames;     // B: the stride between the turns of the per-game outputs
                for (int t = 0; t < turns; ++t)                             // the tile never leaves shared memory between the turns
                    step_lane<P, kRandom, false>(tile, stage_base + G::kActions + (t * W + warp) * L::kActBytes, lane, values5, kw, ku,
                                                 reinterpret_cast<uint8_t*>(rewards) + (t * games + g0 + lane_game) * P, done + t * games + g0 + lane_game,
                                                 illegal ? illegal + t * games + g0 + lane_game : nullptr,
                                                 actions_out ? actions_out + (t * games + g0 + lane_game) * P : nullptr, seed,
                                                 game0 + (uint64_t)g0 + lane_game, turn + (uint32_t)t, nullptr, sel8);
            }
            fence_async_smem();   // make this lane's shared-memory writes visible to the TMA engine
        }
        __syncwarp();
        if (lane == 0) mbar_arrive_a(computed_a + 8u * (uint32_t)stage);
        if (++stage == S) { stage = 0; phase ^= 1u; }
    }
}

// kRandom = false: actions is the tape to play; true: actions (may be NULL) receives the cards drawn in the kernel.
template <int P, bool kRandom, bool kChoice, bool kPacked, int S>
static void launch_step_tiles(const StateView& s, uint8_t* actions, int8_t* rewards, uint8_t* done, uint8_t* illegal, uint64_t seed, uint32_t turn,
                              uint64_t game0, cudaStream_t st, const uint8_t* rows, int64_t num_tiles) {
    constexpr int W = StepShape<P>::kWarps;
    constexpr int kSmem = S * StageLayout<P, W, kChoice, kPacked>::kStride;
    constexpr int kThreads = (W + 1) * 32;
    static int occ_cache[kMaxDevices];   // per device: the shared-memory opt-in and the occupancy are device properties
    const int blocks_per_sm = blocks_per_sm_cached(k_step_tiles<P, kRandom, kChoice, false, kPacked, S>, kThreads, kSmem, occ_cache);
    const int num_sms = device_sms(current_device());
    // persistent grid: one resident wave; block b walks groups b, b + #blocks, ...
    const int64_t groups = (num_tiles + W - 1) / W;
    const unsigned blocks = (unsigned)min(groups, (int64_t)num_sms * blocks_per_sm);
    // launched with programmatic stream serialisation: the grid may start its prologue while the previous kernel of the stream
    // drains (it waits at griddepcontrol.wait before touching memory).  NIMMT_STEP_PDL=0 launches it the plain way.
    static const bool pdl = [] { const char* v = getenv("NIMMT_STEP_PDL"); return !(v && v[0] == '0'); }();
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(blocks);
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = kSmem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 1 : 0;
    cudaLaunchKernelEx(&cfg, k_step_tiles<P, kRandom, kChoice, false, kPacked, S>, s, (const uint8_t*)actions, kRandom ? actions : (uint8_t*)nullptr, rewards, done,
                       illegal, (int)num_tiles, seed, turn, game0, rows, 1);
}

// kRandom = false: actions is the tape to play; true: actions (may be NULL) receives the cards drawn in the kernel.
template <int P, bool kRandom, bool kChoice = false, bool kPacked = false>
static int launch_step(const StateView& s, uint8_t* actions, int8_t* rewards, uint8_t* done, uint8_t* illegal, uint64_t seed, uint32_t turn,
                       uint64_t game0, cudaStream_t st, const uint8_t* rows = nullptr) {
    const int64_t num_tiles = s.B / kTileGames;
    if (num_tiles > 0) {
        constexpr int S0 = StepShape<P>::kStages, S1 = StepShape<P>::kStagesDeep;
        if (S1 != S0 && !kChoice && !kPacked && num_tiles >= kDeepTiles)
            launch_step_tiles<P, kRandom, false, false, S1>(s, actions, rewards, done, illegal, seed, turn, game0, st, rows, num_tiles);
        else
            launch_step_tiles<P, kRandom, kChoice, kPacked, S0>(s, actions, rewards, done, illegal, seed, turn, game0, st, rows, num_tiles);
    }
    const int64_t tail0 = num_tiles * kTileGames;
    if (tail0 < s.B) {   // ragged tail (< 32 games): plain loads
        if constexpr (kPacked) return NIMMT_E_UNSUPPORTED;   // checked by the entry point: the packed format needs whole tiles
        else k_step<P, kRandom><<<1, kStepThreads, 0, st>>>(s, actions, kRandom ? actions : nullptr, rewards, done, illegal, seed, turn, game0, tail0, rows);
    }
    return 0;
}

// `turns` consecutive steps in one launch.  Needs whole tiles (B % 32 == 0: the turns of the [T][B][P] arrays then start on
// 16-byte boundaries); otherwise, and for a single turn, the caller falls back to one launch per turn.
template <int P, bool kRandom>
static int launch_step_many(const StateView& s, uint8_t* actions, int8_t* rewards, uint8_t* done, uint8_t* illegal, uint64_t seed, uint32_t turn,
                            uint64_t game0, int turns, cudaStream_t st) {
    constexpr int W = StepShape<P>::kWarps, S = StepShape<P>::kStages;
    constexpr int kThreads = (W + 1) * 32;
    const int64_t num_tiles = s.B / kTileGames;
    const int smem = S * (int)stage_stride_many<P, W, false>(kRandom ? 0 : turns);
    cudaFuncSetAttribute(k_step_tiles<P, kRandom, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);   // depends on `turns`: set per launch
    int occ = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_step_tiles<P, kRandom, false, true>, kThreads, smem);
    if (occ < 1) return NIMMT_E_UNSUPPORTED;
    const int64_t groups = (num_tiles + W - 1) / W;
    const unsigned blocks = (unsigned)min(groups, (int64_t)device_sms(current_device()) * occ);
    k_step_tiles<P, kRandom, false, true><<<blocks, kThreads, smem, st>>>(s, actions, kRandom ? actions : nullptr, rewards, done, illegal, (int)num_tiles, seed,
                                                                         turn, game0, nullptr, turns);
    return 0;
}

// ------------------------------------------------------------------------------------------
// k_random_actions — DrunkHamster.forward for every (game, player) (agents/random.py:8-10).
// ------------------------------------------------------------------------------------------
constexpr int kActionThreads = 256;   // = the entries of the slot-selection table: every thread stages exactly one

template <int P>
__global__ void __launch_bounds__(kActionThreads)
k_random_actions(StateView s, uint8_t* __restrict__ actions, uint64_t seed, uint32_t turn, uint64_t game0) {
    __shared__ uint32_t sel8[256];
    sel8[threadIdx.x] = d_select8[threadIdx.x];
    const int64_t g = (int64_t)blockIdx.x * kActionThreads + threadIdx.x;
    HandRec hand[P];
    if (g < s.B) load_hands<P>(s, g, hand);   // the loads fly while the table is staged
    __syncthreads();
    if (g >= s.B) return;
    int act[P];
    random_actions_rec<P>(hand, sel8, seed, game0 + (uint64_t)g, turn, act);
    store_bytes<P>(actions, g, act);
}

// ------------------------------------------------------------------------------------------
// k_pack_flags — flag bytes -> bits (one ballot per 32 games), for results that travel over PCIe.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_pack_flags(const uint8_t* __restrict__ flags, uint32_t* __restrict__ bits, int64_t B) {
    const int64_t g = (int64_t)blockIdx.x * 256 + threadIdx.x;   // B is padded to whole warps by the launch
    const uint32_t word = __ballot_sync(0xffffffffu, g < B && flags[g] != 0);
    if ((threadIdx.x & 31) == 0 && g < B) bits[g >> 5] = word;
}

}  // namespace nimmt

using namespace nimmt;

extern "C" {

int nimmt_step(void* state, const uint8_t* actions, int8_t* rewards, uint8_t* done, uint8_t* illegal, int64_t B,
               int num_players, void* stream) {
    if (int rc = check_common(state, B, num_players)) return rc;
    if (!actions || !rewards || !done) return NIMMT_E_BADARG;
    if (!aligned16(actions) || !aligned16(rewards) || !aligned16(done) || (illegal && !aligned16(illegal))) return NIMMT_E_ALIGN;
    if (B == 0) return NIMMT_OK;
    StateView s(state, B, num_players);
    NIMMT_DISPATCH_P(num_players, (launch_step<P, false>(s, const_cast<uint8_t*>(actions), rewards, done, illegal, 0, 0, 0, (cudaStream_t)stream)));
    return check_launch();
}

int nimmt_step_many(void* state, const uint8_t* actions, int8_t* rewards, uint8_t* done, uint8_t* illegal, int64_t B, int num_players, int turns,
                    void* stream) {
    if (int rc = check_common(state, B, num_players)) return rc;
    if (!actions || !rewards || !done || turns < 0 || turns > 10) return NIMMT_E_BADARG;
    if (!aligned16(actions) || !aligned16(rewards) || !aligned16(done) || (illegal && !aligned16(illegal))) return NIMMT_E_ALIGN;
    if (B == 0 || turns == 0) return NIMMT_OK;
    if (turns == 1 || B % kTileGames != 0) {   // one launch per turn (ragged batches: the turns of [T][B][P] are not 16-byte aligned)
        if (B % 16 != 0 && turns > 1) return NIMMT_E_ALIGN;
        for (int t = 0; t < turns; ++t)
            if (int rc = nimmt_step(state, actions + (int64_t)t * B * num_players, rewards + (int64_t)t * B * num_players, done + (int64_t)t * B,
                                    illegal ? illegal + (int64_t)t * B : nullptr, B, num_players, stream))
                return rc;
        return NIMMT_OK;
    }
    StateView s(state, B, num_players);
    int rc = 0;
    NIMMT_DISPATCH_P(num_players, (rc = launch_step_many<P, false>(s, const_cast<uint8_t*>(actions), rewards, done, illegal, 0, 0, 0, turns, (cudaStream_t)stream)));
    return rc ? rc : check_launch();
}

int nimmt_step_random_many(void* state, uint8_t* actions, int8_t* rewards, uint8_t* done, int64_t B, int num_players, uint64_t seed, uint32_t turn,
                           uint64_t game0, int turns, void* stream) {
    if (int rc = check_common(state, B, num_players)) return rc;
    if (!rewards || !done || turns < 0 || turns > 10) return NIMMT_E_BADARG;
    if ((actions && !aligned16(actions)) || !aligned16(rewards) || !aligned16(done)) return NIMMT_E_ALIGN;
    if (B == 0 || turns == 0) return NIMMT_OK;
    if (turns == 1 || B % kTileGames != 0) {
        if (B % 16 != 0 && turns > 1) return NIMMT_E_ALIGN;
        for (int t = 0; t < turns; ++t)
            if (int rc = nimmt_step_random(state, actions ? actions + (int64_t)t * B * num_players : nullptr, rewards + (int64_t)t * B * num_players,
                                           done + (int64_t)t * B, B, num_players, seed, turn + (uint32_t)t, game0, stream))
                return rc;
        return NIMMT_OK;
    }
    StateView s(state, B, num_players);
    int rc = 0;
    NIMMT_DISPATCH_P(num_players, (rc = launch_step_many<P, true>(s, actions, rewards, done, nullptr, seed, turn, game0, turns, (cudaStream_t)stream)));
    return rc ? rc : check_launch();
}

int nimmt_packed_bytes(int num_players, int* action_bytes, int* result_bytes) {
    if (num_players < 1 || num_players > kMaxPlayers) return NIMMT_E_BADARG;
    if (action_bytes) *action_bytes = (num_players + 1) / 2;
    if (result_bytes) *result_bytes = (5 * num_players + 2 + 7) / 8;
    return NIMMT_OK;
}

int nimmt_step_packed(void* state, const uint8_t* slots, uint8_t* results, int64_t B, int num_players, void* stream) {
    if (int rc = check_common(state, B, num_players)) return rc;
    if (!slots || !results) return NIMMT_E_BADARG;
    if (!aligned16(slots) || !aligned16(results)) return NIMMT_E_ALIGN;
    if (B % kTileGames != 0) return NIMMT_E_UNSUPPORTED;   // whole 32-game tiles only (pad the batch)
    if (B == 0) return NIMMT_OK;
    StateView s(state, B, num_players);
    int rc = 0;
    NIMMT_DISPATCH_P(num_players, (rc = launch_step<P, false, false, true>(s, const_cast<uint8_t*>(slots), reinterpret_cast<int8_t*>(results), nullptr, nullptr, 0,
                                                                           0, 0, (cudaStream_t)stream)));
    return rc ? rc : check_launch();
}

int nimmt_step_choice(void* state, const uint8_t* actions, const uint8_t* rows, int8_t* rewards, uint8_t* done, uint8_t* illegal, int64_t B,
                      int num_players, void* stream) {
    if (int rc = check_common(state, B, num_players)) return rc;
    if (!actions || !rows || !rewards || !done) return NIMMT_E_BADARG;
    if (!aligned16(actions) || !aligned16(rows) || !aligned16(rewards) || !aligned16(done) || (illegal && !aligned16(illegal))) return NIMMT_E_ALIGN;
    if (B == 0) return NIMMT_OK;
    StateView s(state, B, num_players);
    NIMMT_DISPATCH_P(num_players, (launch_step<P, false, true>(s, const_cast<uint8_t*>(actions), rewards, done, illegal, 0, 0, 0, (cudaStream_t)stream, rows)));
    return check_launch();
}

int nimmt_pack_flags(const uint8_t* flags, uint32_t* bits, int64_t B, void* stream) {
    if (!flags || !bits || B < 0) return NIMMT_E_BADARG;
    if (reinterpret_cast<uintptr_t>(bits) & 3u) return NIMMT_E_ALIGN;
    if (B == 0) return NIMMT_OK;
    k_pack_flags<<<blocks_for(B, 256), 256, 0, (cudaStream_t)stream>>>(flags, bits, B);
    return check_launch();
}

int nimmt_step_random(void* state, uint8_t* actions, int8_t* rewards, uint8_t* done, int64_t B, int num_players,
                      uint64_t seed, uint32_t turn, uint64_t game0, void* stream) {
    if (int rc = check_common(state, B, num_players)) return rc;
    if (!rewards || !done) return NIMMT_E_BADARG;
    if ((actions && !aligned16(actions)) || !aligned16(rewards) || !aligned16(done)) return NIMMT_E_ALIGN;
    if (B == 0) return NIMMT_OK;
    StateView s(state, B, num_players);
    NIMMT_DISPATCH_P(num_players, (launch_step<P, true>(s, actions, rewards, done, nullptr, seed, turn, game0, (cudaStream_t)stream)));
    return check_launch();
}

int nimmt_random_actions(const void* state, uint8_t* actions, int64_t B, int num_players, uint64_t seed, uint32_t turn,
                         uint64_t game0, void* stream) {
    if (int rc = check_common(state, B, num_players)) return rc;
    if (!actions) return NIMMT_E_BADARG;
    if (!aligned16(actions)) return NIMMT_E_ALIGN;
    if (B == 0) return NIMMT_OK;
    StateView s(const_cast<void*>(state), B, num_players);
    NIMMT_DISPATCH_P(num_players, k_random_actions<P><<<blocks_for(B, kActionThreads), kActionThreads, 0, (cudaStream_t)stream>>>(
                                      s, actions, seed, turn, game0));
    return check_launch();
}

}  // extern "C"
