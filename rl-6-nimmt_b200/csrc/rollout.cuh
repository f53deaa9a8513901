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

// Root position as the rollouts consume it: the unseen cards and the own hand as byte lists (one
// 116-byte record: 104 pool slots, 10 own slots, 2 pad — copied into each thread's private scratch at
// the start of every rollout), plus the row keys.
constexpr int kRolloutDeckStride = 116;   // 29 words: odd word stride => private decks never share a bank per index
constexpr int kOwnOffset = 104;

struct RolloutRoot {
    alignas(4) uint8_t deck[kRolloutDeckStride];
    int n_pool, n_own;
    BoardLite board;
};

template <int P>
NIMMT_HD bool make_rollout_root(const nimmt_root& root, const uint8_t* values, RolloutRoot& rr) {
    uint4 own, pool;
    if (!decode_root<P>(root, values, own, pool, rr.board)) return false;
    int n = 0;
    for (int c = 0; c < kCards; ++c)
        if (mask_has(pool, (uint32_t)c)) rr.deck[n++] = (uint8_t)c;
    rr.n_pool = n;
    for (int i = n; i < kOwnOffset; ++i) rr.deck[i] = 0;
    n = 0;
    for (int c = 0; c < kCards; ++c)
        if (mask_has(own, (uint32_t)c)) rr.deck[kOwnOffset + n++] = (uint8_t)c;
    rr.n_own = n;
    for (int i = kOwnOffset + n; i < kRolloutDeckStride; ++i) rr.deck[i] = 0;
    return true;
}

// Fisher-Yates without running it.  Step i of the shuffle swaps positions i and k_i = target[i] >= i and outputs what
// then sits at position i.  Position p holds its original entry unless an earlier step j (the latest with k_j == p) moved
// position j's content there, and so on back: walking j = i-1 .. 0 gives the ORIGINAL position whose entry step i outputs.
// Every step can be resolved independently (one thread each) and the array is never modified.
template <int MAX_STEPS>
NIMMT_HD int fisher_yates_source(const uint8_t* target, int i) {
    int pos = target[i];
    for (int j = MAX_STEPS - 1; j >= 0; --j) pos = (j < i && (int)target[j] == pos) ? j : pos;
    return pos;
}

// The opponents' cards of one turn: player p draws the next card of the shrinking pool (compile-time recursion so that
// every draw knows which half of which random word it consumes).
template <int P, int I>
NIMMT_HD void draw_opponents(TurnWords<P>& words, uint8_t* deck, uint32_t& drawn, uint32_t n_pool, int (&keys)[P]) {
    if constexpr (I < P) {
        const uint32_t j = drawn + words.template draw<I>(n_pool - drawn);
        NIMMT_CHECK(j < n_pool && n_pool <= (uint32_t)kOwnOffset);
        const uint32_t card = deck[j];
        deck[j] = deck[drawn];   // position `drawn` is never read again: half a swap suffices
        ++drawn;
        keys[I] = (int)(card << 10) | (I << 6);   // place_v3's key: card << 10, the player (non-zero here) below
        draw_opponents<P, I + 1>(words, deck, drawn, n_pool, keys);
    }
}

// Plays one rollout.  `first_index` = rank (ascending) of the card forced as player 0's first move
// (stratified root: the caller runs the same number of rollouts for every legal card).  `deck` is 116
// bytes of scratch private to the caller.  Cards are drawn by partial Fisher-Yates: the opponents'
// P-1 cards per turn from the pool prefix, player 0's card by swap-remove from its own list — a few
// shared-memory byte accesses per draw instead of a popcount search through a 104-bit mask.
// Returns the outcome (sum of player 0's rewards, <= 0).  Depends on (seed, rollout_id) only.
// keys_w / keys_u: the caller's private row keys (game.cuh::place_v3), row r at word r * KEY_STRIDE.  values5: bull heads << 5.
template <int P, int KEY_STRIDE = 1>
NIMMT_HD int rollout(const RolloutRoot& rr, int first_index, const uint8_t* values5, uint8_t* deck, uint32_t* keys_w, uint32_t* keys_u, uint64_t seed,
                     uint64_t rollout_id) {
#pragma unroll
    for (int w = 0; w < kRolloutDeckStride / 4; ++w) reinterpret_cast<uint32_t*>(deck)[w] = reinterpret_cast<const uint32_t*>(rr.deck)[w];
    Philox rng(seed, rollout_id, /*stream=*/0x6d637300u, 0);
#pragma unroll
    for (int r = 0; r < kRows; ++r) { keys_w[r * KEY_STRIDE] = (uint32_t)rr.board.w[r]; keys_u[r * KEY_STRIDE] = key_u_from_w((uint32_t)rr.board.w[r]); }
    int n_own = rr.n_own;
    uint32_t drawn = 0;
    const uint32_t n_pool = (uint32_t)rr.n_pool;
    uint32_t taken5 = 0;   // player 0's bull heads so far, << 5
    TurnWords<P> words;
    int turn = 0;
    while (n_own > 0) {
        int keys[P];
        words.begin_turn(rng, turn);
        // player 0: forced card on the first turn, uniform afterwards (mcts.py:140-145, 187-188)
        {
            const uint32_t pick = words.template draw<0>((uint32_t)n_own);
            const uint32_t idx = turn == 0 ? (uint32_t)first_index : pick;
            NIMMT_CHECK(idx < (uint32_t)n_own && n_own <= kHand);
            const int card = deck[kOwnOffset + idx];
            --n_own;
            deck[kOwnOffset + idx] = deck[kOwnOffset + n_own];   // swap-remove
            keys[0] = card << 10;
        }
        draw_opponents<P, 1>(words, deck, drawn, n_pool, keys);
        sort_keys<P>(keys);
#pragma unroll
        for (int i = 0; i < P; ++i) {
            uint32_t row, keep4;
            const uint32_t pen5 = place_v3<KEY_STRIDE>(keys_w, keys_u, (uint32_t)keys[i], values5, row, keep4);
            taken5 += (keys[i] & 0x3C0) == 0 ? pen5 : 0u;   // mcts.py:150: player 0's rewards only
        }
        ++turn;
    }
    return -(int)(taken5 >> 5);
}

}  // namespace nimmt
