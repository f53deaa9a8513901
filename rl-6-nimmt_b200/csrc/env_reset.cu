// env_reset.cu — deal, deal_from_perm, reset_to, scores (SechsNimmtEnv.reset / _deal / reset_to,
// env.py:43-62, 99-112) and the library-wide helpers of the C ABI.  One thread per game.
#include "abi_common.cuh"

namespace nimmt {

thread_local char g_last_error[256] = "";

// ------------------------------------------------------------------------------------------
// k_deal — SechsNimmtEnv.reset/_deal (env.py:43-51, 99-112).  Write-only: (12 P + 24) B/game.
// ------------------------------------------------------------------------------------------
constexpr int kDealThreads = 64;   // 104 words x 64 threads = 26 KB of interleaved decks per block

template <int P>
__global__ void __launch_bounds__(kDealThreads) k_deal(StateView s, uint64_t seed, uint64_t game0) {
    __shared__ uint8_t values[128];
    __shared__ uint32_t decks[kCards * kDealThreads];
    stage_card_values(values);
    __syncthreads();
    const int64_t g = (int64_t)blockIdx.x * kDealThreads + threadIdx.x;
    if (g >= s.B) return;
    GameRec<P> gm;
    deal_game<P>(gm, seed, game0 + (uint64_t)g, values, decks + threadIdx.x, kDealThreads);
    store_game<P>(s, g, gm);
}

// k_deal_from_perm — the same, from caller-supplied shuffled decks (env.py:103-112).
template <int P>
__global__ void __launch_bounds__(kStepThreads) k_deal_from_perm(StateView s, const uint8_t* __restrict__ perm) {
    __shared__ uint8_t values[128];
    stage_card_values(values);
    __syncthreads();
    const int64_t g = (int64_t)blockIdx.x * kStepThreads + threadIdx.x;
    if (g >= s.B) return;
    const uint8_t* deck = perm + g * kCards;
    GameRec<P> gm;
#pragma unroll
    for (int p = 0; p < P; ++p) {
        int cards[kHand];
#pragma unroll
        for (int i = 0; i < kHand; ++i) cards[i] = deck[p * kHand + i];
        set_dealt_hand<P>(gm, p, cards);
    }
#pragma unroll
    for (int r = 0; r < kRows; ++r) {
        const uint32_t card = deck[kCards - 1 - r];
        gm.board.set_row(r, card, card, 1u, values[card & 127]);
    }
    store_game<P>(s, g, gm);
}

// ------------------------------------------------------------------------------------------
// k_reset_to — SechsNimmtEnv.reset_to (env.py:53-62) from -1 padded board / hands arrays.
// ------------------------------------------------------------------------------------------
template <int P>
__global__ void __launch_bounds__(kStepThreads)
k_reset_to(StateView s, const int8_t* __restrict__ board, const int8_t* __restrict__ hands, uint8_t* __restrict__ invalid) {
    __shared__ uint8_t values[128];
    stage_card_values(values);
    __syncthreads();
    const int64_t g = (int64_t)blockIdx.x * kStepThreads + threadIdx.x;
    if (g >= s.B) return;
    GameRec<P> gm;
    uint4 seen = make_uint4(0, 0, 0, 0);
    bool bad = false;
    const int8_t* bsrc = board + g * (kRows * 6);
#pragma unroll
    for (int r = 0; r < kRows; ++r) {
        uint64_t cards = 0;
        uint32_t len = 0, sum = 0, top = 0;
        bool open = true;
#pragma unroll
        for (int i = 0; i < 6; ++i) {
            const int c = bsrc[r * 6 + i];
            if (c < 0) open = false;  // first -1 ends the row, as list building from the observation does (mcts.py:75-85)
            if (open) {
                const bool ok = c < kCards && i < 5 && !mask_has(seen, (uint32_t)c);
                bad = bad || !ok;
                if (ok) {
                    mask_set(seen, (uint32_t)c);
                    cards |= (uint64_t)(uint32_t)c << (8u * len);
                    ++len;
                    sum += values[c];
                    top = (uint32_t)c;
                }
            }
        }
        bad = bad || len == 0;
        if (len == 0) { len = 1; sum = values[0]; }  // keep the packed invariant 1 <= len <= 5
        gm.board.set_row(r, cards, top, len, sum);
    }
    const int8_t* hsrc = hands + g * (P * kHand);
#pragma unroll
    for (int p = 0; p < P; ++p) {
        uint4 h = make_uint4(0, 0, 0, 0);
#pragma unroll
        for (int i = 0; i < kHand; ++i) {
            const int c = hsrc[p * kHand + i];
            if (c >= 0) {
                const bool ok = c < kCards && !mask_has(seen, (uint32_t)c);
                bad = bad || !ok;
                if (ok) { mask_set(seen, (uint32_t)c); mask_set(h, (uint32_t)c); }
            }
        }
        gm.hand[p] = rec_from_mask(h);   // slots in ascending card order whatever order the caller listed them in
    }
    store_game<P>(s, g, gm);
    if (invalid) invalid[g] = bad;
}

// ------------------------------------------------------------------------------------------
// k_scores — SechsNimmtEnv._scores (env.py:32,167).
// ------------------------------------------------------------------------------------------
template <int P>
__global__ void __launch_bounds__(kStepThreads) k_scores(StateView s, uint8_t* __restrict__ scores) {
    const int64_t g = (int64_t)blockIdx.x * kStepThreads + threadIdx.x;
    if (g >= s.B) return;
    int sc[P];
#pragma unroll
    for (int p = 0; p < P; ++p) sc[p] = (int)(*s.meta_ptr(g, p) >> kRecScoreShift);
    store_bytes<P>(scores, g, sc);
}

}  // namespace nimmt

using namespace nimmt;

extern "C" {

int nimmt_abi_version(void) { return 3; }

const char* nimmt_last_cuda_error(void) { return g_last_error; }

size_t nimmt_state_bytes(int64_t num_games, int num_players) {
    if (num_games < 0 || num_players < 1 || num_players > kMaxPlayers) return 0;
    return (size_t)StateView::bytes(num_games, num_players);   // whole tiles of 32 games, (12 P + 24) bytes per game
}

int nimmt_obs_len(int include_summaries) { return include_summaries ? 47 : 35; }

int nimmt_card_value(int card) { return (card >= 0 && card < kCards) ? h_card_value[card] : -1; }

int nimmt_deal(void* state, int64_t B, int num_players, uint64_t seed, uint64_t game0, void* stream) {
    if (int rc = check_common(state, B, num_players)) return rc;
    if (B == 0) return NIMMT_OK;
    StateView s(state, B, num_players);
    NIMMT_DISPATCH_P(num_players, k_deal<P><<<blocks_for(B, kDealThreads), kDealThreads, 0, (cudaStream_t)stream>>>(s, seed, game0));
    return check_launch();
}

int nimmt_deal_from_perm(void* state, const uint8_t* perm, int64_t B, int num_players, void* stream) {
    if (int rc = check_common(state, B, num_players)) return rc;
    if (!perm) return NIMMT_E_BADARG;
    if (B == 0) return NIMMT_OK;
    StateView s(state, B, num_players);
    NIMMT_DISPATCH_P(num_players, k_deal_from_perm<P><<<blocks_for(B, kStepThreads), kStepThreads, 0, (cudaStream_t)stream>>>(s, perm));
    return check_launch();
}

int nimmt_reset_to(void* state, const int8_t* board, const int8_t* hands, uint8_t* invalid, int64_t B, int num_players,
                   void* stream) {
    if (int rc = check_common(state, B, num_players)) return rc;
    if (!board || !hands) return NIMMT_E_BADARG;
    if (B == 0) return NIMMT_OK;
    StateView s(state, B, num_players);
    NIMMT_DISPATCH_P(num_players,
                     k_reset_to<P><<<blocks_for(B, kStepThreads), kStepThreads, 0, (cudaStream_t)stream>>>(s, board, hands, invalid));
    return check_launch();
}

int nimmt_scores(const void* state, uint8_t* scores, int64_t B, int num_players, void* stream) {
    if (int rc = check_common(state, B, num_players)) return rc;
    if (!scores) return NIMMT_E_BADARG;
    if (!aligned16(scores)) return NIMMT_E_ALIGN;
    if (B == 0) return NIMMT_OK;
    StateView s(const_cast<void*>(state), B, num_players);
    NIMMT_DISPATCH_P(num_players, k_scores<P><<<blocks_for(B, kStepThreads), kStepThreads, 0, (cudaStream_t)stream>>>(s, scores));
    return check_launch();
}

}  // extern "C"
