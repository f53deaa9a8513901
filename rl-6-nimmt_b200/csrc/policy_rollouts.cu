// policy_rollouts.cu — PolicyMCSAgent / PUCTAgent ("Alpha0.5") search, one decision per tree, whole
// searches resident on chip (agents/mcts.py:91-154 with _choose_action_mc of :209-217 and :276-293).
//
// The reference plays n_mc rollouts one after the other; in every rollout every player's move — the
// opponents' and, after the first move, player 0's own — is sampled from softmax(policy net) over that
// player's legal cards, and player 0's first move is chosen by PUCT over the outcomes of the previous
// rollouts.  That makes a tree strictly sequential (rollout j needs the outcomes of 0..j-1), so the
// parallelism is across trees and across the (player, card) rows of one turn:
//
//   one CTA = floor(12 / P) trees in lock-step; 128-row tile = (tree, player, hand slot), three decisions per warp;
//   per turn:  build the [card | observation] rows from the trees' state in shared memory
//              -> policy net on tcgen05 (policy_tile.cuh) -> per (tree, player) softmax + sample (or PUCT)
//              -> per tree one env step (game.cuh::RowKeys) -> next turn;
//   per rollout: restore the root, deal the opponents from the agent's unseen cards (partial
//              Fisher-Yates), play to the end, file the outcome under the first card.
// Nothing leaves the SM between the first and the last rollout; the only HBM traffic is the 64-byte
// root, the 37 KB weight blob and the 240-byte result per tree.
#include "policy_tile.cuh"
#include "puct.cuh"
#include "rollout.cuh"

namespace nimmt {

constexpr int kModePuct = 0, kModeStratified = 2;   // mode 1: the root move is sampled from the policy like every other move

// Board block of a tree in observation order (env.py:186-207), 36 bytes so that it is restored with word copies.
struct BoardBlock {
    int8_t board[kRows][6];            // -1 padded
    uint8_t len[kRows], top[kRows], sum[kRows];
};

struct TreeState {
    int8_t hand[kMaxPlayers][kHand];   // each ascending, -1 padded (the observation's hand block)
    alignas(4) BoardBlock cur, root;
    int8_t action[kMaxPlayers];
    int8_t root_hand[kHand];
    uint8_t deck[kCards];              // the agent's unseen cards (any order), then garbage
    int n_avail, root_n, valid, outcome, first_index;
    float root_prob[kHand], cur_prob[kHand];
    RootStats stats;
};
static_assert(sizeof(BoardBlock) == 36, "BoardBlock is copied as nine 32-bit words");

__device__ __forceinline__ float uniform01(uint32_t w) { return (float)(w >> 8) * (1.0f / 16777216.0f); }

template <int P>
__global__ void __launch_bounds__(kTileRows, 1)
k_policy_rollouts(const nimmt_root* __restrict__ roots, int D, const uint8_t* __restrict__ blob, int n_mc, float c_puct, int mode,
                  uint64_t seed, unsigned long long* __restrict__ stats_out, float* __restrict__ root_probs_out) {
    constexpr int T = 12 / P;                      // trees per CTA
    constexpr unsigned kFull = 0xffffffffu;
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_slot;
    __shared__ TreeState trees[T];
    __shared__ uint8_t values[128];

    for (uint32_t i = threadIdx.x * 16; i < kBlobBytes; i += kTileRows * 16)
        *reinterpret_cast<uint4*>(smem + kSmemBlob + i) = *reinterpret_cast<const uint4*>(blob + i);
    stage_card_values(values);
    if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
        fence_barrier_init();
    }
    if (threadIdx.x < 32) tmem_alloc(&tmem_slot, kTmemColsPerGroup);
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = tmem_slot;
    uint32_t phase = 0;
    PhaseClock pc;
    uint8_t* gbuf = smem + kSmemGroups;

    // thread roles: row = (decision, hand slot).  Warp w carries decisions 3 w .. 3 w + 2 in lanes 0..29, so the
    // softmax, the sampling and the hand update of a decision are warp shuffles; lanes 30 and 31 are dead rows.
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int dloc = lane / kSlots, slot = lane % kSlots;
    const int dec = warp * 3 + dloc;
    const int gbase = (dloc < 3 ? dloc : 2) * kSlots;                        // first lane of this row's decision
    const bool row_live = dloc < 3 && dec < T * P;
    const int tree_l = row_live ? dec / P : 0, player = row_live ? dec % P : 0;
    const bool is_dec = row_live && slot == 0;                               // one thread per (tree, player)
    const bool is_tree = is_dec && player == 0;                              // one thread per tree
    const int tree_g = blockIdx.x * T + tree_l;
    TreeState& ts = trees[tree_l];

    // ---- decode the root (BaseMCAgent's view, agents/mcts.py:62-89) ----
    if (is_tree) {
        ts.valid = 0;
        if (tree_g < D) {
            const nimmt_root root = roots[tree_g];
            uint4 own, pool;
            BoardLite lite;
            if (decode_root<P>(root, values, own, pool, lite)) {
                ts.valid = 1;
                int n = 0;
                for (int c = 0; c < kCards; ++c)
                    if (mask_has(own, c)) ts.root_hand[n++] = (int8_t)c;
                ts.root_n = n;
                for (int i = n; i < kHand; ++i) ts.root_hand[i] = -1;
                int na = 0;
                for (int c = 0; c < kCards; ++c)
                    if (mask_has(pool, c)) ts.deck[na++] = (uint8_t)c;
                ts.n_avail = na;
                for (int r = 0; r < kRows; ++r) {
                    int len = 0;
                    for (int i = 0; i < 6; ++i) {
                        const int c = root.rows[r][i];
                        const bool ok = c < kCards && len == i && i < 5;
                        ts.root.board[r][i] = ok ? (int8_t)c : (int8_t)-1;
                        len += ok;
                    }
                    ts.root.len[r] = (uint8_t)lite.len(r);
                    ts.root.top[r] = (uint8_t)lite.top(r);
                    ts.root.sum[r] = (uint8_t)lite.sum(r);
                }
                root_stats_clear(ts.stats);
                for (int i = 0; i < kHand; ++i) ts.root_prob[i] = 0.0f;
            }
        }
    }
    __syncthreads();
    int n_root_max = 0;
    for (int t = 0; t < T; ++t) n_root_max = max(n_root_max, trees[t].valid ? trees[t].root_n : 0);
    const bool tree_ok = row_live && ts.valid;
    const int root_n = tree_ok ? ts.root_n : 0;

    pc.start();
    for (int j = 0; j < n_mc; ++j) {
        // ---- new rollout: restore the root, deal the opponents (agents/mcts.py:108-127) ----
        if (tree_ok && player == 0) {
            ts.hand[0][slot] = ts.root_hand[slot];
            if (slot < 9) reinterpret_cast<uint32_t*>(&ts.cur)[slot] = reinterpret_cast<const uint32_t*>(&ts.root)[slot];
        }
        if (is_tree && tree_ok) {
            ts.outcome = 0;
            ts.first_index = -1;
            // partial Fisher-Yates over the unseen cards: the first (P-1) n entries become the opponents' hands
            Philox rng(seed, ((uint64_t)tree_g << 24) | (uint64_t)j, 0x6465616cu, 0);
            uint4 rw = make_uint4(0, 0, 0, 0);
            const int need = (P - 1) * root_n, n_avail = ts.n_avail;
            for (int i = 0; i < need; ++i) {
                if ((i & 3) == 0) rw = rng.next();
                const uint32_t w = (i & 3) == 0 ? rw.x : (i & 3) == 1 ? rw.y : (i & 3) == 2 ? rw.z : rw.w;
                const int k = i + (int)below(w, (uint32_t)(n_avail - i));
                const uint8_t a = ts.deck[i], b = ts.deck[k];
                ts.deck[i] = b; ts.deck[k] = a;
            }
        }
        __syncthreads();
        {   // each opponent sorts its chunk (mcts.py:124): rank sort across the decision's ten lanes
            const bool opp = tree_ok && player > 0;
            const int v = opp && slot < root_n ? (int)ts.deck[(player - 1) * root_n + slot] : 127;
            int rank = 0;
#pragma unroll
            for (int s = 0; s < kSlots; ++s) rank += __shfl_sync(kFull, v, gbase + s) < v;
            if (opp) {
                if (slot < root_n) ts.hand[player][rank] = (int8_t)v;   // cards are distinct, so ranks are a permutation
                else ts.hand[player][slot] = -1;
            }
        }
        __syncthreads();
        pc.mark(10);

        for (int turn = 0; turn < n_root_max; ++turn) {
            const bool playing = tree_ok && turn < root_n;   // shorter roots idle until the longest finishes
            const int h = root_n - turn;                      // cards in every hand of this tree
            // ---- features of every (tree, player, slot) row (env.py:174-212 layout behind the candidate card) ----
            const int card = playing ? ts.hand[player][slot] : -1;
            const bool live = playing && card >= 0;
            write_feature_row(gbuf, threadIdx.x, [&](int k) -> float {
                if (!live) return 0.0f;
                if (k == 0) return (float)card;
                if (k <= 10) return (float)ts.hand[player][k - 1];
                if (k == 11) return (float)P;
                if (k <= 15) return (float)ts.cur.len[k - 12];
                if (k <= 19) return (float)ts.cur.top[k - 16];
                if (k <= 23) return (float)ts.cur.sum[k - 20];
                return (float)ts.cur.board[(k - 24) / 6][(k - 24) % 6];
            });
            pc.mark(0);
            const float logit = mlp_tile(smem + kSmemBlob, gbuf, tmem_base, &bar, phase, threadIdx.x, 0, pc);

            // ---- softmax over the decision's hand, every lane of the decision redundantly (mcts.py:219-228) ----
            float pr[kSlots];
            float m = -INFINITY;
#pragma unroll
            for (int s = 0; s < kSlots; ++s) {
                pr[s] = __shfl_sync(kFull, logit, gbase + s);
                if (s < h) m = fmaxf(m, pr[s]);
            }
            float z = 0.0f;
#pragma unroll
            for (int s = 0; s < kSlots; ++s) {
                pr[s] = s < h ? __expf(pr[s] - m) : 0.0f;
                z += pr[s];
            }
#pragma unroll
            for (int s = 0; s < kSlots; ++s) pr[s] /= z;

            // ---- player 0's first move: PUCT over the earlier outcomes, or the stratified schedule ----
            int pick = -1;
            if (turn == 0) {
                if (playing && player == 0) {
                    const float mine = slot < h ? __expf(logit - m) / z : 0.0f;
                    ts.cur_prob[slot] = mine;
                    if (j == 0) ts.root_prob[slot] = mine;
                }
                __syncwarp();
                if (is_tree && playing) {
                    if (mode == kModePuct) pick = puct_choose(ts.stats, ts.cur_prob, h, c_puct, nullptr);   // mcts.py:281-293
                    else if (mode == kModeStratified) pick = j % h;
                }
                __syncwarp();
                pick = __shfl_sync(kFull, pick, gbase);
            }
            // ---- every other move: Categorical(probs).sample() (mcts.py:212-213) by inverse CDF ----
            {
                Philox rng(seed, ((uint64_t)tree_g << 24) | (uint64_t)j, 0x73616d70u + (uint32_t)turn, (uint32_t)player);
                const float u = uniform01(rng.next().x);
                float acc = 0.0f;
                int below_u = 0;   // the CDF is non-decreasing: the first s with u < cdf[s] is the number of s with u >= cdf[s]
#pragma unroll
                for (int s = 0; s < kSlots; ++s) {
                    acc += pr[s];
                    below_u += (s < h && !(u < acc)) ? 1 : 0;
                }
                if (pick < 0) pick = min(below_u, h - 1);
            }
            // ---- hand.remove(card): the lanes behind the pick shift down by one ----
            const int chosen = __shfl_sync(kFull, card, gbase + max(pick, 0));
            const int next = __shfl_sync(kFull, card, (lane + 1) & 31);
            if (playing) {
                ts.hand[player][slot] = (int8_t)(slot < pick ? card : (slot < kSlots - 1 ? next : -1));
                if (slot == 0) {
                    ts.action[player] = (int8_t)chosen;
                    if (player == 0 && turn == 0) ts.first_index = pick;
                }
            }
            pc.mark(7);
            __syncthreads();
            pc.mark(8);

            // ---- one thread per tree: env.step (env.py:120-136) on the shared-memory board ----
            if (is_tree && playing) {
                RowKeys rk;
                for (int r = 0; r < kRows; ++r) rk.set_row(r, ts.cur.top[r], ts.cur.len[r], ts.cur.sum[r]);
                int keys[P];
#pragma unroll
                for (int p = 0; p < P; ++p) keys[p] = ((int)ts.action[p] << 4) | p;
                sort_keys<P>(keys);
                int outcome = ts.outcome;
#pragma unroll
                for (int i = 0; i < P; ++i) {
                    const int c = keys[i] >> 4;
                    int row;
                    uint32_t keep_len;
                    const int pen = rk.place(c, values[c], row, keep_len);
                    if (keep_len == 0)
                        for (int s = 1; s < 6; ++s) ts.cur.board[row][s] = -1;
                    ts.cur.board[row][keep_len] = (int8_t)c;
                    if ((keys[i] & 15) == 0) outcome -= pen;                                         // mcts.py:150
                }
                ts.outcome = outcome;
                for (int r = 0; r < kRows; ++r) { ts.cur.len[r] = (uint8_t)rk.len(r); ts.cur.top[r] = (uint8_t)rk.top(r); ts.cur.sum[r] = (uint8_t)rk.sum(r); }
            }
            __syncthreads();
            pc.mark(9);
        }
        if (is_tree && tree_ok) root_stats_add(ts.stats, ts.first_index, ts.outcome);                // mcts.py:100
    }

#ifdef NIMMT_PHASE_CLOCKS
    {
        static const char* const names[] = {"features", "sync+fence", "mma1 wait", "epilogue1", "sync", "mma2 wait", "epilogue2", "softmax", "sync", "env step", "deal+sort"};
        pc.print(names, 11);
    }
#endif
    // ---- results ----
    if (tree_ok && player == 0) {
        const int a = slot;
        unsigned long long* o = stats_out + ((int64_t)tree_g * 10 + a) * 3;
        o[0] = (unsigned long long)(long long)ts.stats.sum[a];
        o[1] = (unsigned long long)ts.stats.sumsq[a];
        o[2] = (unsigned long long)ts.stats.count[a];
        root_probs_out[(int64_t)tree_g * 10 + a] = ts.root_prob[a];
    }
    tc_fence_before_sync();
    __syncthreads();
    if (threadIdx.x < 32) tmem_dealloc(tmem_base, kTmemColsPerGroup);
}

}  // namespace nimmt

using namespace nimmt;

extern "C" {

int nimmt_policy_rollouts(const nimmt_root* roots, int num_roots, int num_players, const void* weights, int n_mc, float c_puct,
                          int mode, uint64_t seed, int64_t* stats, float* root_probs, void* stream) {
    if (!roots || !weights || !stats || !root_probs || num_roots < 0 || n_mc < 0 || n_mc >= (1 << 24) || mode < 0 || mode > 2 ||
        num_players < 1 || num_players > kMaxPlayers)
        return NIMMT_E_BADARG;
    if (!aligned16(roots) || !aligned16(weights) || (reinterpret_cast<uintptr_t>(stats) & 7u)) return NIMMT_E_ALIGN;
    if (num_roots == 0) return NIMMT_OK;
    const int trees_per_block = 12 / num_players;
    const unsigned blocks = (unsigned)((num_roots + trees_per_block - 1) / trees_per_block);
    unsigned long long* st = reinterpret_cast<unsigned long long*>(stats);
    const uint8_t* blob = static_cast<const uint8_t*>(weights);
    cudaStream_t cs = (cudaStream_t)stream;
    switch (num_players) {
#define CASE(P_)                                                                                                             \
    case P_:                                                                                                                 \
        cudaFuncSetAttribute(k_policy_rollouts<P_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)policy_smem_bytes(1));            \
        k_policy_rollouts<P_><<<blocks, kTileRows, policy_smem_bytes(1), cs>>>(roots, num_roots, blob, n_mc, c_puct, mode, seed, st, root_probs); \
        break;
        CASE(1) CASE(2) CASE(3) CASE(4) CASE(5) CASE(6) CASE(7) CASE(8) CASE(9) CASE(10)
#undef CASE
        default: return NIMMT_E_BADARG;
    }
    return check_launch();
}

}  // extern "C"
