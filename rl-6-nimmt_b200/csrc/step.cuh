// step.cuh — one game's worth of each environment operation, on registers.
//
// The kernels in env_kernels.cu are thin: load packed state (coalesced SoA planes), call one of
// these, store.  Reference semantics per function are cited inline (paths relative to the
// reference repository).
#pragma once
#include "game.cuh"

namespace nimmt {

template <int P>
struct Game {
    uint4 hand[P];  // bits 0..103 cards, bits 120..127 cumulative score
    Board board;
};

// Player-indexed penalties of one step packed 6 bits per player (a player plays one card per
// step, so at most one row sum <= 27 lands in each field).
template <int P>
struct PenaltyPack {
    using type = uint64_t;
};
template <> struct PenaltyPack<1> { using type = uint32_t; };
template <> struct PenaltyPack<2> { using type = uint32_t; };
template <> struct PenaltyPack<3> { using type = uint32_t; };
template <> struct PenaltyPack<4> { using type = uint32_t; };
template <> struct PenaltyPack<5> { using type = uint32_t; };

// SechsNimmtEnv.step minus the observation rebuild (env.py:64-77).
//   act[p]      card played by player p
//   values      104-entry bull-head table (shared memory on the device)
//   penalty[p]  bull heads taken by player p this step (reward = -penalty, env.py:169)
// Returns false — and leaves the game untouched — if any card is not in its owner's hand
// (env.py:68-69: every move is checked before anything is mutated).
template <int P>
NIMMT_HD bool step_game(Game<P>& g, const int (&act)[P], const uint8_t* values, int (&penalty)[P]) {
    bool legal = true;
#pragma unroll
    for (int p = 0; p < P; ++p) legal = legal && mask_has(g.hand[p], (uint32_t)act[p]);
#pragma unroll
    for (int p = 0; p < P; ++p) penalty[p] = 0;
    if (!legal) return false;

    // sorted((card, player)) ascending by card (env.py:124-125)
    int keys[P];
#pragma unroll
    for (int p = 0; p < P; ++p) keys[p] = (act[p] << 4) | p;
    sort_keys<P>(keys);

    typename PenaltyPack<P>::type packed = 0;
#pragma unroll
    for (int i = 0; i < P; ++i) {
        const int card = keys[i] >> 4, player = keys[i] & 15;
        const int pen = g.board.place(card, values[card]);  // env.py:126-134
        packed += (typename PenaltyPack<P>::type)pen << (6 * player);
    }
#pragma unroll
    for (int p = 0; p < P; ++p) {
        penalty[p] = (int)((packed >> (6 * p)) & 63u);
        mask_clear(g.hand[p], (uint32_t)act[p]);              // env.py:131
        g.hand[p].w += (uint32_t)penalty[p] << kScoreShift;   // env.py:167
    }
    return true;
}

// SechsNimmtEnv._is_done (env.py:246-249): player 0 has no cards left.
template <int P>
NIMMT_HD bool game_done(const Game<P>& g) {
    return (g.hand[0].x | g.hand[0].y | g.hand[0].z | (g.hand[0].w & kHighCardMask)) == 0u;
}

// SechsNimmtEnv._deal (env.py:99-112) with a counter RNG: 10 P + 4 draws without replacement
// from the 104-card deck.  Draw i < 10 P goes to hand i / 10 (the reference's perm[10p .. 10p+9]),
// draw 10 P + r opens row r (the reference's perm[103 - r]): a prefix plus four more entries of a
// uniform permutation, which is all the reference's shuffle provides.
template <int P>
NIMMT_HD void deal_game(Game<P>& g, uint64_t seed, uint64_t game_id, const uint8_t* values) {
    Philox rng(seed, game_id, /*stream=*/0x6e696d74u, 0);
    uint4 deck = make_uint4(0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, kHighCardMask);
    uint32_t left = kCards;
    uint4 r = make_uint4(0, 0, 0, 0);
    auto draw = [&](int i) -> uint32_t {
        if ((i & 3) == 0) r = rng.next();
        const uint32_t word = (i & 3) == 0 ? r.x : (i & 3) == 1 ? r.y : (i & 3) == 2 ? r.z : r.w;
        const uint32_t card = mask_select(deck, below(word, left));
        mask_clear(deck, card);
        --left;
        return card;
    };
#pragma unroll
    for (int p = 0; p < P; ++p) {
        uint4 h = make_uint4(0, 0, 0, 0);
#pragma unroll
        for (int i = 0; i < kHand; ++i) mask_set(h, draw(p * kHand + i));
        g.hand[p] = h;
    }
#pragma unroll
    for (int row = 0; row < kRows; ++row) {
        const uint32_t card = draw(P * kHand + row);
        g.board.tk[row] = (int)(card * 4u) + row;
        g.board.meta[row] = 1u | ((uint32_t)values[card] << 3);
        g.board.cards[row] = card;
    }
}

// DrunkHamster.forward (agents/random.py:8-10): a uniform card of each non-empty hand.
// Random words come from the stream (seed, game_id, turn); player p uses word p.
template <int P>
NIMMT_HD void random_actions_game(const Game<P>& g, uint64_t seed, uint64_t game_id, uint32_t turn, int (&act)[P]) {
    Philox rng(seed, game_id, /*stream=*/0x61637400u + turn, 0);
    uint4 r = make_uint4(0, 0, 0, 0);
#pragma unroll
    for (int p = 0; p < P; ++p) {
        if ((p & 3) == 0) r = rng.next();
        const uint32_t word = (p & 3) == 0 ? r.x : (p & 3) == 1 ? r.y : (p & 3) == 2 ? r.z : r.w;
        const uint32_t n = (uint32_t)mask_count(g.hand[p]);
        act[p] = n ? (int)mask_select(g.hand[p], below(word, n)) : 255;
    }
}

}  // namespace nimmt
