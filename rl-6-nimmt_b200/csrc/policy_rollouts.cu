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
#include <climits>

namespace nimmt {

constexpr int kModePuct = 0, kModeStratified = 2;   // mode 1: the root move is sampled from the policy like every other move

// Board features of a tree as bf16 bit patterns, in the order and chunking the policy rows consume them
// (env.py:186-207): 80 bytes, restored from the root with five 16-byte copies.
struct alignas(16) BoardFeat {
    uint16_t len[kRows], pad[4];       // features 12..15: second half of chunk 1
    uint16_t top[kRows], sum[kRows];   // chunk 2
    uint16_t board[kRows][6];          // chunks 3..5, -1 padded
};
static_assert(sizeof(BoardFeat) == 80, "BoardFeat is copied as five uint4");

struct TreeState {
    alignas(16) uint16_t hand[kMaxPlayers][16];   // bf16: [p][0] unused (the row's own card goes there), [p][1..10] the hand,
                                                  // ascending and -1 padded (the observation's hand block), [p][11] = P
    BoardFeat cur, root;
    int8_t sorted[kMaxPlayers][kHand];            // the opponents' freshly dealt hands (integers), per rollout
    int8_t action[kMaxPlayers];
    int8_t root_hand[kHand];
    uint8_t deck[kCards];              // the agent's unseen cards (any order), then garbage
    uint8_t draw[kCards];              // this rollout's Fisher-Yates swap targets
    int n_avail, root_n, valid;
    float root_prob[kHand];
    RootStats stats;
};

__device__ __forceinline__ float uniform01(uint32_t w) { return (float)(w >> 8) * (1.0f / 16777216.0f); }
constexpr uint16_t kBf16MinusOne = 0xBF80;
constexpr uint32_t kSearchSmemBytes = kSmemGroups + kA1Bytes + kA2TailBytes;   // the weights + one layer-1 operand + the tail of layer 2's
// tensor memory: 128 + 32 columns (policy_tile.cuh) — three CTAs per SM, which is what lets the 4 x 86 CTAs of a self-play turn
// (four seats' searches of 256 trees on four streams) run as ONE wave of 148 x 3 = 444 slots instead of two of 296

template <int P>
__global__ void __launch_bounds__(kTileRows, 3)   // three CTAs per SM: <= 168 registers
k_policy_rollouts(const nimmt_root* __restrict__ roots, int D, const uint8_t* __restrict__ blob, int n_mc, float c_puct, int mode,
                  uint64_t seed, unsigned long long* __restrict__ stats_out, float* __restrict__ root_probs_out) {
    constexpr int T = 12 / P;                      // trees per CTA
    constexpr unsigned kFull = 0xffffffffu;
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_slot, tmem_slot_b;
    __shared__ TreeState trees[T];
    __shared__ uint8_t values[128];

    for (uint32_t i = threadIdx.x * 16; i < kBlobBytes; i += kTileRows * 16)
        *reinterpret_cast<uint4*>(smem + kSmemBlob + i) = *reinterpret_cast<const uint4*>(blob + i);
    stage_card_values(values);
    if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
        fence_barrier_init();
    }
    if (threadIdx.x < 32) {
        tmem_alloc_keep_permit(&tmem_slot, kSearchTmemA);
        tmem_alloc_keep_permit(&tmem_slot_b, kSearchTmemB);
        tmem_relinquish_permit();
    }
    uint8_t* gbuf = smem + kSmemGroups;
    uint8_t* a2tail = gbuf + kA1Bytes;
    init_feature_constants(gbuf, threadIdx.x);
    init_a2_tail(a2tail, threadIdx.x);
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = tmem_slot, tmem_b = tmem_slot_b;
    uint32_t phase = 0;
    PhaseClock pc;

    // thread roles: row = (decision, hand slot).  Warp w carries decisions 3 w .. 3 w + 2 in lanes 0..29, so the
    // softmax, the sampling and the hand update of a decision are warp shuffles; lanes 30 and 31 are dead rows.
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int dloc = lane / kSlots, slot = lane % kSlots;
    const int dec = warp * kDecPerWarp + dloc;
    const int gbase = (dloc < kDecPerWarp ? dloc : kDecPerWarp - 1) * kSlots;   // first lane of this row's decision
    const bool row_live = dloc < kDecPerWarp && dec < T * P;
    const int tree_l = row_live ? dec / P : 0, player = row_live ? dec % P : 0;
    const bool is_dec = row_live && slot == 0;                               // one thread per (tree, player)
    const bool is_tree = is_dec && player == 0;                              // one thread per tree
    const int tree_g = blockIdx.x * T + tree_l;
    TreeState& ts = trees[tree_l];
    RowKeys rk_root, rk;                                                     // live in the tree thread's registers
#pragma unroll
    for (int r = 0; r < kRows; ++r) { rk_root.set_row(r, 0, 1, 0); rk.set_row(r, 0, 1, 0); }

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
#pragma unroll
                for (int r = 0; r < kRows; ++r) {
                    int len = 0;
                    for (int i = 0; i < 6; ++i) {
                        const int c = root.rows[r][i];
                        const bool ok = c < kCards && len == i && i < 5;
                        ts.root.board[r][i] = ok ? (uint16_t)bf16_bits(c) : kBf16MinusOne;
                        len += ok;
                    }
                    ts.root.len[r] = (uint16_t)bf16_bits(lite.len(r));
                    ts.root.top[r] = (uint16_t)bf16_bits(lite.top(r));
                    ts.root.sum[r] = (uint16_t)bf16_bits(lite.sum(r));
                    ts.root.pad[r] = 0;
                    rk_root.set_row(r, lite.top(r), lite.len(r), lite.sum(r));
                }
                for (int p = 0; p < P; ++p)
                    for (int i = 0; i < 16; ++i) ts.hand[p][i] = i == 11 ? (uint16_t)bf16_bits(P) : (uint16_t)0;
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
    const int n_avail = tree_ok ? ts.n_avail : 0;
    const int need = (P - 1) * root_n;                // cards dealt to the opponents per rollout
    uint16_t* my_hand = ts.hand[player];

    pc.start();
    for (int j = 0; j < n_mc; ++j) {
        const uint64_t rollout_id = ((uint64_t)tree_g << 24) | (uint64_t)j;
        // ---- new rollout: restore the root, deal the opponents (agents/mcts.py:108-127) ----
        int card = -1;                                // the card in this row's hand slot
        if (tree_ok) {
            if (player == 0) {
                card = ts.root_hand[slot];
                my_hand[1 + slot] = (uint16_t)bf16_bits(card);
                if (slot < 5) reinterpret_cast<uint4*>(&ts.cur)[slot] = reinterpret_cast<const uint4*>(&ts.root)[slot];
            }
            // partial Fisher-Yates over the unseen cards, the first (P-1) n entries become the opponents' hands:
            // the swap targets are drawn in parallel (one per row) and resolved in parallel below
            const int i = player * kSlots + slot;
            if (i < need) {
                Philox rng(seed, rollout_id, 0x6465616cu, (uint32_t)i);
                ts.draw[i] = (uint8_t)(i + (int)below(rng.next<7>().x, (uint32_t)(n_avail - i)));   // swap target k_i of Fisher-Yates step i
            }
        }
        int outcome = 0, first_index = -1;
        rk = rk_root;
        pc.mark(11);
        __syncthreads();
        pc.mark(12);
        {   // Resolve the shuffle without running it: draw i takes what sits at position k_i after swaps 0..i-1.  Position p
            // holds its original card unless an earlier swap j (the latest with k_j == p) moved position j's content there,
            // and so on back — a walk over j = i-1 .. 0 that every row does for its own draw in parallel, instead of a chain
            // of (P-1) n dependent shared-memory swaps on one thread.  The deck itself is never modified.  Then each opponent
            // sorts its chunk (mcts.py:124): rank sort across the decision's ten lanes.
            const bool opp = tree_ok && player > 0;
            const int i = (player - 1) * root_n + slot;          // this row's draw: opponent `player`, card `slot` of its chunk
            int v = 127;
            if (opp && slot < root_n) {
                v = (int)ts.deck[fisher_yates_source<(P - 1) * kHand>(ts.draw, i)];
            }
            pc.mark(13);
            int rank = 0;
#pragma unroll
            for (int s = 0; s < kSlots; ++s) rank += __shfl_sync(kFull, v, gbase + s) < v;
            if (opp) ts.sorted[player][slot < root_n ? rank : slot] = (int8_t)(slot < root_n ? v : -1);   // distinct cards: ranks are a permutation
            __syncwarp();
            if (opp) {
                card = ts.sorted[player][slot];
                my_hand[1 + slot] = (uint16_t)bf16_bits(card);
            }
            __syncwarp();
        }
        pc.mark(10);

        for (int turn = 0; turn < n_root_max; ++turn) {
            const bool playing = tree_ok && turn < root_n;   // shorter roots idle until the longest finishes
            const int h = root_n - turn;                      // cards in every hand of this tree
            // ---- features of every (tree, player, slot) row: [card | env.py:174-212 observation | 1 1 0 ..], six 16-byte chunks ----
            {
                uint4 c0 = *reinterpret_cast<const uint4*>(my_hand);
                c0.x = (c0.x & 0xFFFF0000u) | bf16_bits(card);
                const uint2 h1 = *reinterpret_cast<const uint2*>(my_hand + 8), ln = *reinterpret_cast<const uint2*>(ts.cur.len);
                store_feature_chunk(gbuf, threadIdx.x, 0, c0);
                store_feature_chunk(gbuf, threadIdx.x, 1, make_uint4(h1.x, h1.y, ln.x, ln.y));
                store_feature_chunk(gbuf, threadIdx.x, 2, *reinterpret_cast<const uint4*>(ts.cur.top));
#pragma unroll
                for (int c = 0; c < 3; ++c) store_feature_chunk(gbuf, threadIdx.x, 3 + c, reinterpret_cast<const uint4*>(ts.cur.board)[c]);
            }
            pc.mark(0);
            // Categorical(probs).sample() (mcts.py:212-213) by the Gumbel-max trick: argmax_s (logit_s - log(-log u_s)) is
            // a draw from softmax(logits), so a sampled move needs no exponentials, no sum and no CDF.  The row's
            // Gumbel variate is computed while the tensor core runs layer 1.
            float gumbel = 0.0f;
            const float logit = mlp_tile(smem + kSmemBlob, gbuf, a2tail, tmem_base, tmem_b, &bar, phase, threadIdx.x, 0, pc, [&](uint32_t token) {
                Philox rng(seed, rollout_id, (0x73616d70u + (uint32_t)turn) ^ token, (uint32_t)(player * 16 + slot));
                const float u = ((float)(rng.next<7>().x >> 8) + 0.5f) * (1.0f / 16777216.0f);   // in (0, 1)
                gumbel = -__logf(-__logf(u));
                pin_result(gumbel);
            }, [] {});

            // ---- player 0's first move: PUCT over the earlier outcomes, or the stratified schedule ----
            int pick = -1;
            if (turn == 0) {
                // softmax over the decision's hand (mcts.py:219-228), every lane of the decision redundantly
                float e[kSlots], z, m;
                decision_softmax(logit, gbase, playing ? (1u << h) - 1u : 0u, e, z, m);
                const float mine = slot < h ? __expf(logit - m) / z : 0.0f;
                if (playing && player == 0 && j == 0) ts.root_prob[slot] = mine;
                if (mode == kModePuct) {
                    // mcts.py:281-302: lane a computes the PUCT value of card a; the strict '>' scan over ascending cards
                    // (NaN never wins, so 0/0 everywhere picks the first card) runs on the gathered values
                    double mine_puct;
                    const int choice = puct_choose_lanes(ts.stats, playing && player == 0, slot, h, gbase, mine, c_puct, mine_puct);
                    if (player == 0) pick = choice;
                } else if (mode == kModeStratified) {
                    if (playing && player == 0) pick = j % h;
                }
            }
            pc.mark(14);
            // ---- every other move: the largest perturbed logit among the decision's cards ----
            {
                // order-preserving integer image of the float, hand slot in the low four bits (ties: lowest slot)
                const uint32_t bits = __float_as_uint(logit + gumbel);
                const int ordered = (int)(bits ^ ((uint32_t)((int)bits >> 31) >> 1));
                const int key = playing && slot < h ? (ordered & ~15) | (15 - slot) : INT_MIN;
                int winner = INT_MIN;   // ten independent shuffles + a max tree (a redux.sync per decision serialises over the masks)
#pragma unroll
                for (int s = 0; s < kSlots; ++s) winner = max(winner, __shfl_sync(kFull, key, gbase + s));
                if (pick < 0) pick = 15 - (winner & 15);
            }
            pc.mark(15);
            // ---- hand.remove(card): the lanes behind the pick shift down by one ----
            const int chosen = __shfl_sync(kFull, card, gbase + max(pick, 0));
            const int next = __shfl_sync(kFull, card, (lane + 1) & 31);
            if (playing) {
                const int moved = slot < pick ? card : (slot < kSlots - 1 ? next : -1);
                if (moved != card) my_hand[1 + slot] = (uint16_t)bf16_bits(moved);
                card = moved;
                if (slot == 0) {
                    ts.action[player] = (int8_t)chosen;
                    if (player == 0 && turn == 0) first_index = pick;
                }
            }
            pc.mark(7);
            __syncthreads();
            pc.mark(8);

            // ---- one thread per tree: env.step (env.py:120-136); row keys in registers, features in shared memory ----
            if (is_tree && playing) {
                int keys[P];
#pragma unroll
                for (int p = 0; p < P; ++p) keys[p] = ((int)ts.action[p] << 4) | p;
                sort_keys<P>(keys);
#pragma unroll
                for (int i = 0; i < P; ++i) {
                    const int c = keys[i] >> 4;
                    int row;
                    uint32_t keep_len;
                    const int pen = rk.place(c, values[c], row, keep_len);
                    if (keep_len == 0) {
#pragma unroll
                        for (int s = 1; s < 6; ++s) ts.cur.board[row][s] = kBf16MinusOne;
                    }
                    ts.cur.board[row][keep_len] = (uint16_t)bf16_bits(c);
                    if ((keys[i] & 15) == 0) outcome -= pen;                                         // mcts.py:150
                }
                *reinterpret_cast<uint2*>(ts.cur.len) = make_uint2(bf16x2_bits(rk.len(0), rk.len(1)), bf16x2_bits(rk.len(2), rk.len(3)));
                *reinterpret_cast<uint4*>(ts.cur.top) = make_uint4(bf16x2_bits(rk.top(0), rk.top(1)), bf16x2_bits(rk.top(2), rk.top(3)),
                                                                   bf16x2_bits(rk.sum(0), rk.sum(1)), bf16x2_bits(rk.sum(2), rk.sum(3)));
            }
            __syncthreads();
            pc.mark(9);
        }
        if (is_tree && tree_ok) root_stats_add(ts.stats, first_index, outcome);                      // mcts.py:100
    }

#ifdef NIMMT_PHASE_CLOCKS
    {
        static const char* const names[] = {"features", "sync+fence", "mma1 wait", "epilogue1", "sync", "mma2 wait", "epilogue2", "hand update", "sync", "env step",
                                            "rank sort", "restore+draw", "deal sync", "fy walk", "root rule", "argmax"};
        pc.print(names, 16);
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
    if (threadIdx.x < 32) {
        tmem_dealloc(tmem_base, kSearchTmemA);
        tmem_dealloc(tmem_b, kSearchTmemB);
    }
}

// PUCTAgent._compute_pucts / _normalize_q / the choice (agents/mcts.py:276-315) for a batch of decisions whose outcome lists are
// given — the root rule exactly as k_policy_rollouts evaluates it (same RootStats accumulation, same puct_choose_lanes with the
// decision's cards in lanes 10 d .. 10 d + 9 of a warp), exposed so that it can be called and checked on its own.
__global__ void __launch_bounds__(32)
k_puct_cases(const int* __restrict__ offsets, const int* __restrict__ action_index, const int* __restrict__ outcomes,
             const float* __restrict__ probs, const int* __restrict__ n_legal, int num_decisions, float c_puct, double* __restrict__ pucts,
             int* __restrict__ choice) {
    __shared__ RootStats st[kDecPerWarp];
    const int lane = threadIdx.x, dloc = lane / kSlots, slot = lane % kSlots;
    const int d = blockIdx.x * kDecPerWarp + dloc;
    const bool live = dloc < kDecPerWarp && d < num_decisions;
    const int gbase = (dloc < kDecPerWarp ? dloc : kDecPerWarp - 1) * kSlots;
    if (live && slot == 0) {
        RootStats& s = st[dloc];
        root_stats_clear(s);
        for (int i = offsets[d]; i < offsets[d + 1]; ++i) root_stats_add(s, action_index[i], outcomes[i]);   // agents/mcts.py:100
    }
    __syncwarp();
    const int h = live ? n_legal[d] : 0;
    const float prob = live && slot < h ? probs[(int64_t)d * kSlots + slot] : 0.0f;
    double mine;
    const int c = puct_choose_lanes(st[dloc < kDecPerWarp ? dloc : 0], live, slot, h, gbase, prob, c_puct, mine);
    if (live) {
        pucts[(int64_t)d * kSlots + slot] = slot < h ? mine : 0.0;
        if (slot == 0) choice[d] = c;
    }
}

}  // namespace nimmt

using namespace nimmt;

extern "C" {

int nimmt_policy_rollouts(const nimmt_root* roots, int num_roots, int num_players, const void* weights, int n_mc, float c_puct,
                          int mode, uint64_t seed, int64_t* stats, float* root_probs, void* stream) {
    // PUCT reads the median of all outcomes from a histogram that counts in 16 bits (puct.cuh::RootStats): n_mc <= 65535 there;
    // the other root rules never read it
    if (!roots || !weights || !stats || !root_probs || num_roots < 0 || n_mc < 0 || n_mc >= (1 << 24) || mode < 0 || mode > 2 ||
        (mode == NIMMT_ROOT_PUCT && n_mc > 65535) ||
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
        cudaFuncSetAttribute(k_policy_rollouts<P_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSearchSmemBytes);            \
        k_policy_rollouts<P_><<<blocks, kTileRows, kSearchSmemBytes, cs>>>(roots, num_roots, blob, n_mc, c_puct, mode, seed, st, root_probs); \
        break;
        CASE(1) CASE(2) CASE(3) CASE(4) CASE(5) CASE(6) CASE(7) CASE(8) CASE(9) CASE(10)
#undef CASE
        default: return NIMMT_E_BADARG;
    }
    return check_launch();
}

int nimmt_puct_choose(const int32_t* offsets, const int32_t* action_index, const int32_t* outcomes, const float* probs,
                      const int32_t* n_legal, int num_decisions, float c_puct, double* pucts, int32_t* choice, void* stream) {
    if (!offsets || !action_index || !outcomes || !probs || !n_legal || !pucts || !choice || num_decisions < 0) return NIMMT_E_BADARG;
    if (num_decisions == 0) return NIMMT_OK;
    k_puct_cases<<<(unsigned)((num_decisions + kDecPerWarp - 1) / kDecPerWarp), 32, 0, (cudaStream_t)stream>>>(
        offsets, action_index, outcomes, probs, n_legal, num_decisions, c_puct, pucts, choice);
    return check_launch();
}

}  // extern "C"
