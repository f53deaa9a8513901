// env_step.cu — batched env.step, random actions, and the two fused (SechsNimmtEnv.step, env.py:64-77;
// DrunkHamster.forward, agents/random.py:8-10).  One thread per game; HBM-bandwidth bound:
// every thread issues its P + 2 plane loads up front, works in registers, writes the planes back.
#include "abi_common.cuh"
#include "tma.cuh"

namespace nimmt {

// ------------------------------------------------------------------------------------------
// k_step — SechsNimmtEnv.step (env.py:64-77) without the observation rebuild.
// Reads (16 P + 24) + P bytes and writes (16 P + 24) + P + 1 (+1) bytes per game.
// kRandom: the actions are drawn in-kernel (DrunkHamster, agents/random.py:8-10) instead of read.
// ------------------------------------------------------------------------------------------
template <int P, bool kRandom>
__global__ void __launch_bounds__(kStepThreads)
k_step(StateView s, const uint8_t* __restrict__ actions_in, uint8_t* __restrict__ actions_out, int8_t* __restrict__ rewards,
       uint8_t* __restrict__ done, uint8_t* __restrict__ illegal, uint64_t seed, uint32_t turn, uint64_t game0,
       int64_t first_game) {
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
    Game<P> gm;
    unpack_raw<P>(raw, gm);
    if constexpr (kRandom) {
        random_actions_game<P>(gm, seed, game0 + (uint64_t)g, turn, act);
        if (actions_out) store_bytes<P>(actions_out, g, act);
    }

    int penalty[P];
    const bool legal = step_game<P>(gm, act, values, penalty);
    if (legal) store_game<P>(s, g, gm);

    int rew[P];
#pragma unroll
    for (int p = 0; p < P; ++p) rew[p] = -penalty[p];
    store_bytes<P>(reinterpret_cast<uint8_t*>(rewards), g, rew);
    done[g] = game_done<P>(gm);
    if (illegal) illegal[g] = !legal;
}

// ------------------------------------------------------------------------------------------
// k_step_tiles — the same step, for whole tiles of 128 games, with the loads taken off the
// threads: one elected thread asks the TMA engine (cp.async.bulk, 1-D) to stream the next tile's
// planes — P hand planes of 2 KB, the 2 KB + 1 KB row planes, 128 P action bytes, each one
// contiguous in HBM — into the other half of a double buffer while the block steps the current
// tile out of shared memory.  A thread therefore never sits on outstanding loads with its 70
// registers pinned, and every block keeps one to two tiles (12-24 KB for P = 4) in flight the
// whole time, which is what it takes to cover HBM latency at ~40 % occupancy.
// Stores go straight from registers (16-byte, warp-contiguous).  Each block owns kTilesPerBlock
// consecutive tiles; the ragged tail of the batch is left to k_step.
// ------------------------------------------------------------------------------------------
constexpr int kTileGames = 128;
constexpr int kTilesPerBlock = 4;

template <int P>
struct TileLayout {
    static constexpr int kHandPlane = kTileGames * 16;
    static constexpr int kRowsA = P * kHandPlane;
    static constexpr int kRowsB = kRowsA + kTileGames * 16;
    static constexpr int kActions = kRowsB + kTileGames * 8;
    static constexpr int kBytes = kActions + kTileGames * P;   // multiple of 16 for every P
    static constexpr int kStride = (kBytes + 127) / 128 * 128;
};

template <int P>
__device__ __forceinline__ void issue_tile(const StateView& s, const uint8_t* actions, int64_t tile, uint8_t* dst, uint64_t* bar) {
    using L = TileLayout<P>;
    const int64_t g0 = tile * kTileGames;
    mbar_arrive_expect_tx(bar, L::kBytes);
#pragma unroll
    for (int p = 0; p < P; ++p) bulk_load(dst + p * L::kHandPlane, s.hand + (int64_t)p * s.B + g0, L::kHandPlane, bar);
    bulk_load(dst + L::kRowsA, s.rows_a + g0, kTileGames * 16, bar);
    bulk_load(dst + L::kRowsB, s.rows_b + g0, kTileGames * 8, bar);
    bulk_load(dst + L::kActions, actions + g0 * P, kTileGames * P, bar);
}

template <int P>
__global__ void __launch_bounds__(kTileGames)
k_step_tiles(StateView s, const uint8_t* __restrict__ actions, int8_t* __restrict__ rewards, uint8_t* __restrict__ done,
             uint8_t* __restrict__ illegal, int64_t num_tiles) {
    using L = TileLayout<P>;
    extern __shared__ __align__(128) uint8_t tile_smem[];   // 2 x L::kStride
    __shared__ uint64_t full[2];
    __shared__ uint8_t values[128];

    const int64_t first = (int64_t)blockIdx.x * kTilesPerBlock;
    const int n_tiles = (int)min((int64_t)kTilesPerBlock, num_tiles - first);
    if (threadIdx.x == 0) {
        mbar_init(&full[0], 1);
        mbar_init(&full[1], 1);
        fence_barrier_init();
        issue_tile<P>(s, actions, first, tile_smem, &full[0]);
        if (n_tiles > 1) issue_tile<P>(s, actions, first + 1, tile_smem + L::kStride, &full[1]);
    }
    stage_card_values(values);
    __syncthreads();

    for (int it = 0; it < n_tiles; ++it) {
        const int stage = it & 1;
        const uint8_t* buf = tile_smem + stage * L::kStride;
        mbar_wait(&full[stage], (uint32_t)(it >> 1) & 1u);

        RawGame<P> raw;
#pragma unroll
        for (int p = 0; p < P; ++p) raw.hand[p] = reinterpret_cast<const uint4*>(buf + p * L::kHandPlane)[threadIdx.x];
        raw.rows_a = reinterpret_cast<const uint4*>(buf + L::kRowsA)[threadIdx.x];
        raw.rows_b = reinterpret_cast<const uint2*>(buf + L::kRowsB)[threadIdx.x];
        int act[P];
        load_bytes<P>(buf + L::kActions, threadIdx.x, act);
        __syncthreads();   // every thread holds its game: this half of the buffer is free again
        if (threadIdx.x == 0 && it + 2 < n_tiles) issue_tile<P>(s, actions, first + it + 2, tile_smem + stage * L::kStride, &full[stage]);

        const int64_t g = (first + it) * kTileGames + threadIdx.x;
        Game<P> gm;
        unpack_raw<P>(raw, gm);
        int penalty[P];
        const bool legal = step_game<P>(gm, act, values, penalty);
        if (legal) store_game<P>(s, g, gm);
        int rew[P];
#pragma unroll
        for (int p = 0; p < P; ++p) rew[p] = -penalty[p];
        store_bytes<P>(reinterpret_cast<uint8_t*>(rewards), g, rew);
        done[g] = game_done<P>(gm);
        if (illegal) illegal[g] = !legal;
    }
}

template <int P>
static int launch_step(const StateView& s, const uint8_t* actions, int8_t* rewards, uint8_t* done, uint8_t* illegal, cudaStream_t st) {
    using L = TileLayout<P>;
    constexpr int kSmem = 2 * L::kStride;
    const int64_t num_tiles = s.B / kTileGames;
    if (num_tiles > 0) {
        static bool configured = false;   // benign race: the attribute is idempotent
        if (!configured) {
            cudaFuncSetAttribute(k_step_tiles<P>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem);
            configured = true;
        }
        const unsigned blocks = (unsigned)((num_tiles + kTilesPerBlock - 1) / kTilesPerBlock);
        k_step_tiles<P><<<blocks, kTileGames, kSmem, st>>>(s, actions, rewards, done, illegal, num_tiles);
    }
    const int64_t tail0 = num_tiles * kTileGames;
    if (tail0 < s.B)   // ragged tail (< 128 games): plain loads
        k_step<P, false><<<1, kStepThreads, 0, st>>>(s, actions, nullptr, rewards, done, illegal, 0, 0, 0, tail0);
    return 0;
}

// ------------------------------------------------------------------------------------------
// k_random_actions — DrunkHamster.forward for every (game, player) (agents/random.py:8-10).
// ------------------------------------------------------------------------------------------
template <int P>
__global__ void __launch_bounds__(kStepThreads)
k_random_actions(StateView s, uint8_t* __restrict__ actions, uint64_t seed, uint32_t turn, uint64_t game0) {
    const int64_t g = (int64_t)blockIdx.x * kStepThreads + threadIdx.x;
    if (g >= s.B) return;
    Game<P> gm;
#pragma unroll
    for (int p = 0; p < P; ++p) gm.hand[p] = s.hand[(int64_t)p * s.B + g];
    int act[P];
    random_actions_game<P>(gm, seed, game0 + (uint64_t)g, turn, act);
    store_bytes<P>(actions, g, act);
}

}  // namespace nimmt

using namespace nimmt;

extern "C" {

int nimmt_step(void* state, const uint8_t* actions, int8_t* rewards, uint8_t* done, uint8_t* illegal, int64_t B,
               int num_players, void* stream) {
    if (int rc = check_common(state, B, num_players)) return rc;
    if (!actions || !rewards || !done) return NIMMT_E_BADARG;
    if (!aligned16(actions) || !aligned16(rewards)) return NIMMT_E_ALIGN;
    if (B == 0) return NIMMT_OK;
    StateView s(state, B, num_players);
    NIMMT_DISPATCH_P(num_players, launch_step<P>(s, actions, rewards, done, illegal, (cudaStream_t)stream));
    return check_launch();
}

int nimmt_step_random(void* state, uint8_t* actions, int8_t* rewards, uint8_t* done, int64_t B, int num_players,
                      uint64_t seed, uint32_t turn, uint64_t game0, void* stream) {
    if (int rc = check_common(state, B, num_players)) return rc;
    if (!rewards || !done) return NIMMT_E_BADARG;
    if ((actions && !aligned16(actions)) || !aligned16(rewards)) return NIMMT_E_ALIGN;
    if (B == 0) return NIMMT_OK;
    StateView s(state, B, num_players);
    NIMMT_DISPATCH_P(num_players, k_step<P, true><<<blocks_for(B, kStepThreads), kStepThreads, 0, (cudaStream_t)stream>>>(
                                      s, nullptr, actions, rewards, done, nullptr, seed, turn, game0, 0));
    return check_launch();
}

int nimmt_random_actions(const void* state, uint8_t* actions, int64_t B, int num_players, uint64_t seed, uint32_t turn,
                         uint64_t game0, void* stream) {
    if (int rc = check_common(state, B, num_players)) return rc;
    if (!actions) return NIMMT_E_BADARG;
    if (!aligned16(actions)) return NIMMT_E_ALIGN;
    if (B == 0) return NIMMT_OK;
    StateView s(const_cast<void*>(state), B, num_players);
    NIMMT_DISPATCH_P(num_players, k_random_actions<P><<<blocks_for(B, kStepThreads), kStepThreads, 0, (cudaStream_t)stream>>>(
                                      s, actions, seed, turn, game0));
    return check_launch();
}

}  // extern "C"
