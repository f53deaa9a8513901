// mc_batch.cu — the Monte-Carlo agents' host-side bookkeeping, batched on the device, so that B games
// with MC agents at some seats can be played in lock-step without leaving the GPU (the batched
// GameSession of SURVEY.md §8f):
//   k_mc_roots   BaseMCAgent._initialize_game / _memorize_cards / _board_from_state (agents/mcts.py:62-89):
//                per game the agent's "available cards" mask loses its own hand and every card lying on the
//                board at decision time (cards played and swept within one step are never seen — the
//                reference's stale memory is kept), then the 64-byte root is written.
//   k_mc_choose  BaseMCAgent._choose_action_from_outcomes (agents/mcts.py:156-165) + the n == 1 shortcut
//                (:52-53): argmax of the per-card mean with strict '>', scanning cards in ascending order.
#include "abi_common.cuh"

namespace nimmt {

template <int P>
__global__ void __launch_bounds__(kStepThreads)
k_mc_roots(StateView s, uint4* __restrict__ available, nimmt_root* __restrict__ roots, int seat, int initialize) {
    const int64_t g = (int64_t)blockIdx.x * kStepThreads + threadIdx.x;
    if (g >= s.B) return;
    HandRec rec;
    rec.lo = *s.cards_ptr(g, seat);
    rec.meta = *s.meta_ptr(g, seat);
    const uint4 hand = rec_to_mask(rec);
    uint4 av = initialize ? make_uint4(0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, kHighCardMask) : available[g];
    av.x &= ~hand.x; av.y &= ~hand.y; av.z &= ~hand.z; av.w &= ~hand.w & kHighCardMask;
    Board b;
    load_rows(s, g, b);
    nimmt_root r;
#pragma unroll
    for (int row = 0; row < kRows; ++row) {
        const int len = b.k.len(row);
#pragma unroll
        for (int i = 0; i < 6; ++i) {
            const uint32_t c = (uint32_t)(b.cards[row] >> (8 * i)) & 0xFFu;
            const bool on = i < len;
            if (on) mask_clear(av, c);
            r.rows[row][i] = on ? (uint8_t)c : (uint8_t)255;
        }
    }
    available[g] = av;
    r.own[0] = hand.x; r.own[1] = hand.y; r.own[2] = hand.z; r.own[3] = hand.w & kHighCardMask;
    r.available[0] = av.x; r.available[1] = av.y; r.available[2] = av.z; r.available[3] = av.w;
    r.num_players = (uint8_t)P;
#pragma unroll
    for (int i = 0; i < 7; ++i) r.pad[i] = 0;
    const uint4* src = reinterpret_cast<const uint4*>(&r);
    uint4* dst = reinterpret_cast<uint4*>(roots + g);
#pragma unroll
    for (int i = 0; i < 4; ++i) dst[i] = src[i];
}

template <int P>
__global__ void __launch_bounds__(kStepThreads)
k_mc_choose(StateView s, const long long* __restrict__ stats, uint8_t* __restrict__ actions, int seat) {
    const int64_t g = (int64_t)blockIdx.x * kStepThreads + threadIdx.x;
    if (g >= s.B) return;
    HandRec rec;
    rec.lo = *s.cards_ptr(g, seat);
    rec.meta = *s.meta_ptr(g, seat);
    const int n = rec_count(rec);
    int best = 0;
    if (n > 1) {
        double best_mean = -INFINITY;
        for (int a = 0; a < n; ++a) {
            const long long sum = stats[(g * 10 + a) * 3], cnt = stats[(g * 10 + a) * 3 + 2];
            const double mean = cnt > 0 ? (double)sum / (double)cnt : NAN;   // np.mean([]) is NaN and never wins
            if (mean > best_mean) { best_mean = mean; best = a; }
        }
    }
    actions[g * P + seat] = n > 0 ? (uint8_t)rec_select(rec, (uint32_t)best) : (uint8_t)255;
}

// ------------------------------------------------------------------------------------------
// k_elo_scan — Tournament._compute_elos (tournament.py:157-164) for B finished games IN ORDER: the multiplayer Elo of the
// `multi_elo` package the reference calls (`calc_elo(players, k)`), i.e. every pair of players of a game is a two-player match,
//     K = k / (n - 1),   S_ij = 1 if place_i < place_j, 1/2 if equal, 0 otherwise,   E_ij = 1 / (1 + 10^((R_j - R_i) / 400)),
//     R_i <- R_i + K * sum_{j != i} (S_ij - E_ij)          (all players of the game updated from the ratings before the game)
// with place = the tie-averaged rank by score (tournament.py:240-247; only its order matters here).  Elo is sequential by
// nature — game b+1 sees the ratings left by game b — so one warp walks the games; lane p is seat p of the current game.
// `multi_elo` is a third-party dependency absent from the reference tree and from this image: the formula above is its
// published algorithm (no rounding of the terms), parity is UNPINNED and says so in DESIGN.md.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(32)
k_elo_scan(const int* __restrict__ scores, const int* __restrict__ agents, double* __restrict__ ratings, int64_t B, int P, double k,
           double* __restrict__ history) {
    const int lane = threadIdx.x;
    const double K = k / (double)(P > 1 ? P - 1 : 1);
    for (int64_t b = 0; b < B; ++b) {
        const bool live = lane < P;
        const int me = live ? (agents ? agents[b * P + lane] : lane) : 0;
        const int score = live ? scores[b * P + lane] : 0;
        const double r = live ? ratings[me] : 0.0;
        double delta = 0.0;
        for (int j = 0; j < P; ++j) {
            const int sj = __shfl_sync(0xffffffffu, score, j);
            const double rj = __shfl_sync(0xffffffffu, r, j);
            if (live && j != lane) {
                const double s = score > sj ? 1.0 : (score == sj ? 0.5 : 0.0);   // higher score = better place (scores are <= 0)
                delta += s - 1.0 / (1.0 + pow(10.0, (rj - r) / 400.0));
            }
        }
        __syncwarp();
        if (live) {
            ratings[me] = r + K * delta;
            if (history) history[b * P + lane] = r + K * delta;
        }
        __syncwarp();
    }
}

}  // namespace nimmt

using namespace nimmt;

extern "C" {

int nimmt_mc_roots(const void* state, void* available, nimmt_root* roots, int64_t B, int num_players, int seat, int initialize,
                   void* stream) {
    if (int rc = check_common(state, B, num_players)) return rc;
    if (!available || !roots || seat < 0 || seat >= num_players) return NIMMT_E_BADARG;
    if (!aligned16(available) || !aligned16(roots)) return NIMMT_E_ALIGN;
    if (B == 0) return NIMMT_OK;
    StateView s(const_cast<void*>(state), B, num_players);
    NIMMT_DISPATCH_P(num_players, k_mc_roots<P><<<blocks_for(B, kStepThreads), kStepThreads, 0, (cudaStream_t)stream>>>(
                                      s, static_cast<uint4*>(available), roots, seat, initialize));
    return check_launch();
}

int nimmt_mc_choose(const void* state, const int64_t* stats, uint8_t* actions, int64_t B, int num_players, int seat, void* stream) {
    if (int rc = check_common(state, B, num_players)) return rc;
    if (!stats || !actions || seat < 0 || seat >= num_players) return NIMMT_E_BADARG;
    if (B == 0) return NIMMT_OK;
    StateView s(const_cast<void*>(state), B, num_players);
    NIMMT_DISPATCH_P(num_players, k_mc_choose<P><<<blocks_for(B, kStepThreads), kStepThreads, 0, (cudaStream_t)stream>>>(
                                      s, reinterpret_cast<const long long*>(stats), actions, seat));
    return check_launch();
}

int nimmt_elo_scan(const int32_t* scores, const int32_t* agents, double* ratings, int64_t num_games, int num_players, double k,
                   double* history, void* stream) {
    if (!scores || !ratings || num_games < 0 || num_players < 1 || num_players > 32 || !(k >= 0.0)) return NIMMT_E_BADARG;
    if (num_games == 0) return NIMMT_OK;
    k_elo_scan<<<1, 32, 0, (cudaStream_t)stream>>>(scores, agents, ratings, num_games, num_players, k, history);
    return check_launch();
}

}  // extern "C"
