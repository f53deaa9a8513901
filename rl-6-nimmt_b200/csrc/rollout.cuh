// rollout.cuh — one uniform-random playout of BaseMCAgent._mcts (agents/mcts.py:91-154) with
// MCSAgent._choose_action_mc (:187-188), on registers.
//
// What the reference does per rollout: shuffle the cards it believes are still unseen, hand the
// first (P-1) n of them to the opponents in sorted chunks of n (:116-127), then let every player
// — player 0 included — play a uniformly random card of their hand each turn (:139-147), and add
// up player 0's rewards (:150).
//
// What this does instead, with the same law: an opponent who plays a uniformly random card of a
// uniformly random hand plays, over the n turns, a uniformly random ordering of a uniformly random
// n-subset; jointly over the P-1 opponents that is a uniformly random injection of the (P-1) n
// (opponent, turn) slots into the unseen cards.  So the opponents' hands never need to exist: each
// turn P-1 cards are drawn without replacement from the shrinking pool of unseen cards.  Row
// contents are not needed either — only each row's top card, length and bull-head sum influence
// the rest of the game (env.py:138-172).  A rollout is therefore ~20 registers of state.
// DESIGN.md §5 has the argument in full; tests check it against exact enumeration.
#pragma once
#include "../../include/nimmt_b200.h"
#include "game.cuh"

namespace nimmt {

// Rows without their card lists (RowKeys, game.cuh) are enough for playouts.
using BoardLite = RowKeys;

// Decodes a root position (BaseMCAgent's view of the game, agents/mcts.py:62-89) into rollout
// state.  Cards in the own hand or lying on the board are never "unseen" (mcts.py:66-73), whatever
// the caller's mask says.  Returns false if the root cannot be played: wrong player count or
// fewer than (P-1) |own| unseen cards (the reference would deal short hands and then raise).
template <int P>
NIMMT_HD bool decode_root(const nimmt_root& root, const uint8_t* values, uint4& own, uint4& pool, BoardLite& board) {
    own = make_uint4(root.own[0], root.own[1], root.own[2], root.own[3] & kHighCardMask);
    pool = make_uint4(root.available[0] & ~own.x, root.available[1] & ~own.y, root.available[2] & ~own.z,
                      root.available[3] & kHighCardMask & ~own.w);
#pragma unroll
    for (int r = 0; r < kRows; ++r) {
        uint32_t len = 0, sum = 0, top = 0;
#pragma unroll
        for (int i = 0; i < 6; ++i) {
            const uint32_t c = root.rows[r][i];
            if (c < (uint32_t)kCards) {
                mask_clear(pool, c);
                if (len == (uint32_t)i && i < 5) { ++len; sum += values[c]; top = c; }
            }
        }
        if (len == 0) return false;
        board.set_row(r, top, len, sum);
    }
    return root.num_players == P && mask_count(pool) >= (P - 1) * mask_count(own);
}

// Plays one rollout.  `own` = the deciding player's hand, `pool` = cards the agent believes unseen
// (BaseMCAgent.available_cards), `first` = the card forced as player 0's first move (stratified
// root: the caller runs the same number of rollouts for every legal card).  Returns the outcome
// (sum of player 0's rewards, <= 0).  Requires |pool| >= (P-1) |own|.
template <int P>
NIMMT_HD int rollout(uint4 own, uint4 pool, BoardLite board, int first, const uint8_t* values, uint64_t seed, uint64_t rollout_id) {
    Philox rng(seed, rollout_id, /*stream=*/0x6d637300u, 0);
    int n_own = mask_count(own);
    uint32_t n_pool = (uint32_t)mask_count(pool);
    int outcome = 0;
    uint4 r = make_uint4(0, 0, 0, 0);
    int used = 4;
    auto next_word = [&]() -> uint32_t {
        if (used == 4) { r = rng.next(); used = 0; }
        const uint32_t w = used == 0 ? r.x : used == 1 ? r.y : used == 2 ? r.z : r.w;
        ++used;
        return w;
    };
    bool first_turn = true;
    while (n_own > 0) {
        int keys[P];
        // player 0: forced card on the first turn, uniform afterwards (mcts.py:140-145, 187-188)
        {
            const uint32_t w = next_word();
            const int card = first_turn ? first : (int)mask_select(own, below(w, (uint32_t)n_own));
            mask_clear(own, (uint32_t)card);
            --n_own;
            keys[0] = card << 4;
        }
#pragma unroll
        for (int p = 1; p < P; ++p) {
            const uint32_t card = mask_select(pool, below(next_word(), n_pool));
            mask_clear(pool, card);
            --n_pool;
            keys[p] = (int)(card << 4) | p;
        }
        sort_keys<P>(keys);
#pragma unroll
        for (int i = 0; i < P; ++i) {
            const int card = keys[i] >> 4;
            int row;
            uint32_t keep_len;
            const int pen = board.place(card, values[card], row, keep_len);
            outcome -= (keys[i] & 15) == 0 ? pen : 0;   // mcts.py:150: player 0's rewards only
        }
        first_turn = false;
    }
    return outcome;
}

}  // namespace nimmt
