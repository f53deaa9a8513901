// env_step.cu — batched env.step, random actions, and the two fused (SechsNimmtEnv.step, env.py:64-77;
// DrunkHamster.forward, agents/random.py:8-10).  One thread per game; HBM-bandwidth bound:
// every thread issues its P + 2 plane loads up front, works in registers, writes the planes back.
#include "abi_common.cuh"

namespace nimmt {

// ------------------------------------------------------------------------------------------
// k_step — SechsNimmtEnv.step (env.py:64-77) without the observation rebuild.
// Reads (16 P + 24) + P bytes and writes (16 P + 24) + P + 1 (+1) bytes per game.
// kRandom: the actions are drawn in-kernel (DrunkHamster, agents/random.py:8-10) instead of read.
// ------------------------------------------------------------------------------------------
template <int P, bool kRandom>
__global__ void __launch_bounds__(kStepThreads)
k_step(StateView s, const uint8_t* __restrict__ actions_in, uint8_t* __restrict__ actions_out, int8_t* __restrict__ rewards,
       uint8_t* __restrict__ done, uint8_t* __restrict__ illegal, uint64_t seed, uint32_t turn, uint64_t game0) {
    __shared__ uint8_t values[128];
    stage_card_values(values);
    __syncthreads();
    const int64_t g = (int64_t)blockIdx.x * kStepThreads + threadIdx.x;
    if (g >= s.B) return;

    Game<P> gm;
    int act[P];
    if constexpr (!kRandom) load_bytes<P>(actions_in, g, act);  // issue with the state loads
    load_game<P>(s, g, gm);
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
    NIMMT_DISPATCH_P(num_players, k_step<P, false><<<blocks_for(B, kStepThreads), kStepThreads, 0, (cudaStream_t)stream>>>(
                                      s, actions, nullptr, rewards, done, illegal, 0, 0, 0));
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
                                      s, nullptr, actions, rewards, done, nullptr, seed, turn, game0));
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
