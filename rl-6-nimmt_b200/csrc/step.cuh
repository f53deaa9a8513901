// step.cuh — one game's worth of each environment operation, on registers.
//
// The kernels in env_kernels.cu are thin: load packed state (coalesced SoA planes), call one of
// these, store.  Reference semantics per function are cited inline (paths relative to the
// reference repository).
#pragma once
#include "handrec.cuh"

namespace nimmt {

// P consecutive bytes at base + g * P, using the widest access the alignment of g * P guarantees
// (base is 16-byte aligned; checked in the C entry points).
template <int P>
NIMMT_HD void load_bytes(const uint8_t* base, int64_t g, int (&v)[P]) {
    const uint8_t* p = base + g * P;
    if constexpr (P % 4 == 0) {
#pragma unroll
        for (int i = 0; i < P / 4; ++i) {
            const uint32_t w = reinterpret_cast<const uint32_t*>(p)[i];
            v[4 * i] = w & 0xFF; v[4 * i + 1] = (w >> 8) & 0xFF; v[4 * i + 2] = (w >> 16) & 0xFF; v[4 * i + 3] = w >> 24;
        }
    } else if constexpr (P % 2 == 0) {
#pragma unroll
        for (int i = 0; i < P / 2; ++i) {
            const uint32_t w = reinterpret_cast<const uint16_t*>(p)[i];
            v[2 * i] = w & 0xFF; v[2 * i + 1] = w >> 8;
        }
    } else {
#pragma unroll
        for (int i = 0; i < P; ++i) v[i] = p[i];
    }
}

template <int P>
NIMMT_HD void store_bytes(uint8_t* base, int64_t g, const int (&v)[P]) {
    uint8_t* p = base + g * P;
    if constexpr (P % 4 == 0) {
#pragma unroll
        for (int i = 0; i < P / 4; ++i)
            reinterpret_cast<uint32_t*>(p)[i] = (uint32_t)(v[4 * i] & 0xFF) | ((uint32_t)(v[4 * i + 1] & 0xFF) << 8) |
                                                ((uint32_t)(v[4 * i + 2] & 0xFF) << 16) | ((uint32_t)(v[4 * i + 3] & 0xFF) << 24);
    } else if constexpr (P % 2 == 0) {
#pragma unroll
        for (int i = 0; i < P / 2; ++i)
            reinterpret_cast<uint16_t*>(p)[i] = (uint16_t)((v[2 * i] & 0xFF) | ((v[2 * i + 1] & 0xFF) << 8));
    } else {
#pragma unroll
        for (int i = 0; i < P; ++i) p[i] = (uint8_t)v[i];
    }
}


template <int P>
struct Game {
    uint4 hand[P];  // bits 0..103 cards, bits 120..127 cumulative score
    Board board;
};

// The same game in its stored form (handrec.cuh): what the kernels load, step and store.
template <int P>
struct GameRec {
    HandRec hand[P];
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

// SechsNimmtEnv._play_cards (env.py:120-136) on a board: the cards are resolved in ascending order, penalty[p]
// receives the bull heads player p takes this step (reward = -penalty, env.py:169).
// `choice` (may be NULL): the free-row-choice mode — choice[p] is the row player p takes if their card undercuts every row.
template <int P>
NIMMT_HD void play_cards(Board& board, const int (&act)[P], const uint8_t* values, int (&penalty)[P], const int* choice = nullptr) {
    // sorted((card, player)) ascending by card (env.py:124-125)
    int keys[P];
#pragma unroll
    for (int p = 0; p < P; ++p) keys[p] = (act[p] << 4) | p;
    sort_keys<P>(keys);

    typename PenaltyPack<P>::type packed = 0;
#pragma unroll
    for (int i = 0; i < P; ++i) {
        const int card = keys[i] >> 4, player = keys[i] & 15;
        int chosen = -1;
        if (choice) {
#pragma unroll
            for (int p = 0; p < P; ++p) chosen = p == player ? choice[p] : chosen;   // static indices: no local memory
        }
        const int pen = board.place(card, values[card], chosen);  // env.py:126-134
        packed += (typename PenaltyPack<P>::type)pen << (6 * player);
    }
#pragma unroll
    for (int p = 0; p < P; ++p) penalty[p] = (int)((packed >> (6 * p)) & 63u);
}

// SechsNimmtEnv.step minus the observation rebuild (env.py:64-77).
//   act[p]      card played by player p
//   values      104-entry bull-head table (shared memory on the device)
//   penalty[p]  bull heads taken by player p this step (reward = -penalty, env.py:169)
// Returns false — and leaves the game untouched — if any card is not in its owner's hand
// (env.py:68-69: every move is checked before anything is mutated).
template <int P>
NIMMT_HD bool step_game(Game<P>& g, const int (&act)[P], const uint8_t* values, int (&penalty)[P], const int* choice = nullptr) {
    // env.py:68-69 + :131 — every card is checked (and tentatively removed) before anything is committed
    uint4 hand[P];
    bool legal = true;
#pragma unroll
    for (int p = 0; p < P; ++p) {
        hand[p] = g.hand[p];
        legal = mask_take(hand[p], (uint32_t)act[p], true) && legal;
        if (choice) legal = legal && (uint32_t)choice[p] < (uint32_t)kRows;
        penalty[p] = 0;
    }
    if (!legal) return false;
    play_cards<P>(g.board, act, values, penalty, choice);
#pragma unroll
    for (int p = 0; p < P; ++p) {
        hand[p].w += (uint32_t)penalty[p] << kScoreShift;     // env.py:167
        g.hand[p] = hand[p];
    }
    return true;
}

// The same step on the stored form: a played card sets its slot's bit, a take adds to the score field.
template <int P>
NIMMT_HD bool step_game(GameRec<P>& g, const int (&act)[P], const uint8_t* values, int (&penalty)[P], const int* choice = nullptr) {
    uint32_t meta[P];
    bool legal = true;
#pragma unroll
    for (int p = 0; p < P; ++p) {
        legal = rec_take(g.hand[p], (uint32_t)act[p], meta[p]) && legal;
        if (choice) legal = legal && (uint32_t)choice[p] < (uint32_t)kRows;
        penalty[p] = 0;
    }
    if (!legal) return false;
    play_cards<P>(g.board, act, values, penalty, choice);
#pragma unroll
    for (int p = 0; p < P; ++p) g.hand[p].meta = meta[p] + ((uint32_t)penalty[p] << kRecScoreShift);   // env.py:167
    return true;
}

// SechsNimmtEnv._is_done (env.py:246-249): player 0 has no cards left.
template <int P>
NIMMT_HD bool game_done(const Game<P>& g) {
    return (g.hand[0].x | g.hand[0].y | g.hand[0].z | (g.hand[0].w & kHighCardMask)) == 0u;
}

template <int P>
NIMMT_HD bool game_done(const GameRec<P>& g) { return rec_empty(g.hand[0]); }

// SechsNimmtEnv._deal (env.py:99-112) with a counter RNG: a partial Fisher-Yates shuffle of the
// 104-card deck, 10 P + 4 draws.  Draw i < 10 P goes to hand i / 10 (the reference's
// perm[10p .. 10p+9]), draw 10 P + r opens row r (the reference's perm[103 - r]): a prefix plus
// four more entries of a uniform permutation, which is all the reference's shuffle provides.
//   deck         the game's private 104-entry scratch, one BYTE per card at a stride of `deck_stride` bytes, already
//                holding the identity (entry j = card j).  On the device the decks of a block are interleaved so that a
//                thread's entries all live in its own shared-memory bank (env_reset.cu): a warp's 32 games hit 32
//                different banks whatever positions they draw.
//   sink         receives the result as it is produced: ten cards per hand in ASCENDING order (the reference sorts each hand,
//                env.py:105) and sink.row(r, card).  Hands are sorted two at a time — both sequences ride in one register
//                per position, 16 bits each, through one comparator network (game.cuh::Pair16) — and handed over in that
//                form, sink.hand_pair(p, k); the last hand of an odd table arrives as sink.hand(p, cards).
// Uniform integers by 32-bit multiply-high, two per random word (below_keep): one Philox4x32-7 call serves eight cards.
template <int P, class Sink>
NIMMT_HD void deal_game(uint64_t seed, uint64_t game_id, uint8_t* deck, int deck_stride, Sink&& sink) {
    Philox rng(seed, game_id, /*stream=*/0x6e696d74u, 0);
    uint4 r = make_uint4(0, 0, 0, 0);
    uint32_t spare = 0;
    auto draw = [&](int i) -> uint32_t {
        if ((i & 7) == 0) r = rng.next<7>();
        uint32_t off;
        if ((i & 1) == 0) {
            spare = (i & 7) == 0 ? r.x : (i & 7) == 2 ? r.y : (i & 7) == 4 ? r.z : r.w;   // i is a compile-time constant after unrolling
            off = below_keep(spare, (uint32_t)(kCards - i));
        } else {
            off = below(spare, (uint32_t)(kCards - i));
        }
        NIMMT_CHECK((uint32_t)i + off < (uint32_t)kCards);
        uint8_t* at = deck + i * deck_stride + off * (uint32_t)deck_stride;   // entry j = i + off
        const uint32_t card = *at;
        *at = deck[i * deck_stride];   // position i is never read again, so only half of the swap is needed
        return card;
    };
#pragma unroll
    for (int p = 0; p + 1 < P; p += 2) {   // hands p and p + 1 together
        Pair16 k[kHand];
#pragma unroll
        for (int i = 0; i < kHand; ++i) k[i].v = draw(p * kHand + i);
#pragma unroll
        for (int i = 0; i < kHand; ++i) k[i].v += draw((p + 1) * kHand + i) << 16;
        sort_keys<kHand>(k);
        sink.hand_pair(p, k);   // hand p in the low halves, hand p + 1 in the high halves
    }
    if constexpr (P % 2 == 1) {
        int k[kHand];
#pragma unroll
        for (int i = 0; i < kHand; ++i) k[i] = (int)draw((P - 1) * kHand + i);
        sort_keys<kHand>(k);
        uint32_t a[kHand];
#pragma unroll
        for (int i = 0; i < kHand; ++i) a[i] = (uint32_t)k[i];
        sink.hand(P - 1, a);
    }
#pragma unroll
    for (int row = 0; row < kRows; ++row) sink.row(row, draw(P * kHand + row));
}

// Sinks that fill a game held in registers (tests/host_sim; the kernel stores straight to HBM instead).
template <int P>
struct DealIntoGame {
    Game<P>& g;
    const uint8_t* values;
    NIMMT_HD void hand(int p, const uint32_t (&cards)[kHand]) {
        uint4 h = make_uint4(0, 0, 0, 0);
#pragma unroll
        for (int i = 0; i < kHand; ++i) mask_set(h, cards[i]);
        g.hand[p] = h;
    }
    NIMMT_HD void hand_pair(int p, const Pair16 (&k)[kHand]) {
        uint32_t a[kHand], b[kHand];
#pragma unroll
        for (int i = 0; i < kHand; ++i) { a[i] = k[i].v & 0xFFFFu; b[i] = k[i].v >> 16; }
        hand(p, a);
        hand(p + 1, b);
    }
    NIMMT_HD void row(int r, uint32_t card) { g.board.set_row(r, card, card, 1u, values[card]); }
};
template <int P>
struct DealIntoGameRec {
    GameRec<P>& g;
    const uint8_t* values;
    NIMMT_HD void hand(int p, const uint32_t (&cards)[kHand]) { g.hand[p] = rec_from_sorted(cards, kHand, 0u); }
    NIMMT_HD void hand_pair(int p, const Pair16 (&k)[kHand]) {
        uint32_t a[kHand], b[kHand];
#pragma unroll
        for (int i = 0; i < kHand; ++i) { a[i] = k[i].v & 0xFFFFu; b[i] = k[i].v >> 16; }
        hand(p, a);
        hand(p + 1, b);
    }
    NIMMT_HD void row(int r, uint32_t card) { g.board.set_row(r, card, card, 1u, values[card]); }
};

// The ten cards dealt to player p (any order) become its hand (deal from caller-supplied decks, env_reset.cu).
template <int P>
NIMMT_HD void set_dealt_hand(GameRec<P>& g, int p, int (&cards)[kHand]) {
    sort_keys<kHand>(cards);   // the reference sorts each hand (env.py:105); slots are in ascending card order
    uint32_t c[kHand];
#pragma unroll
    for (int i = 0; i < kHand; ++i) c[i] = (uint32_t)cards[i];
    g.hand[p] = rec_from_sorted(c, kHand, 0u);
}

// SechsNimmtEnv._create_agent_state + _create_game_state for one seat (env.py:186-212), one int8 per entry:
// [own hand ascending, -1 padded (10) | P | cards per row (4) top cards (4) bull heads per row (4) — if summaries | board 4 x 6, -1 padded].
template <int P>
NIMMT_HD void fill_observation(const GameRec<P>& g, int p, int8_t* o, bool summaries) {
    int n = 0;
    for (int i = 0; i < kHand; ++i)
        if (!rec_slot_empty(g.hand[p], i)) o[n++] = (int8_t)rec_card(g.hand[p], i);
    for (; n < kHand; ++n) o[n] = -1;
    o[n++] = (int8_t)P;
    if (summaries) {
        for (int r = 0; r < kRows; ++r) o[n + r] = (int8_t)g.board.k.len(r);
        for (int r = 0; r < kRows; ++r) o[n + 4 + r] = (int8_t)g.board.k.top(r);
        for (int r = 0; r < kRows; ++r) o[n + 8 + r] = (int8_t)g.board.k.sum(r);
        n += 12;
    }
    for (int r = 0; r < kRows; ++r)
        for (int i = 0; i < 6; ++i) o[n + r * 6 + i] = i < g.board.k.len(r) ? (int8_t)((g.board.cards[r] >> (8 * i)) & 0xFF) : (int8_t)-1;
}

// DrunkHamster.forward (agents/random.py:8-10): a uniform card of each non-empty hand.
// Random words come from the stream (seed, game_id, turn); player p uses word p.
NIMMT_HD int hand_count(const uint4& h) { return mask_count(h); }
NIMMT_HD int hand_count(const HandRec& h) { return rec_count(h); }
NIMMT_HD uint32_t hand_select(const uint4& h, uint32_t k) { return mask_select(h, k); }
NIMMT_HD uint32_t hand_select(const HandRec& h, uint32_t k) { return rec_select(h, k); }

// The stored form with the slot-selection table (handrec.cuh::rec_select_slot): what the kernels run.
template <int P>
NIMMT_HD void random_actions_rec(const HandRec (&hand)[P], const uint32_t* sel8, uint64_t seed, uint64_t game_id, uint32_t turn, int (&act)[P]) {
    Philox rng(seed, game_id, /*stream=*/0x61637400u + turn, 0);
    uint4 r = make_uint4(0, 0, 0, 0);
#pragma unroll
    for (int p = 0; p < P; ++p) {
        if ((p & 3) == 0) r = rng.next<7>();
        const uint32_t word = (p & 3) == 0 ? r.x : (p & 3) == 1 ? r.y : (p & 3) == 2 ? r.z : r.w;
        const uint32_t n = (uint32_t)rec_count(hand[p]);
        act[p] = n ? (int)rec_card_dyn(hand[p], rec_select_slot(sel8, hand[p].meta, below(word, n))) : 255;
    }
}

template <int P, class G>
NIMMT_HD void random_actions_game(const G& g, uint64_t seed, uint64_t game_id, uint32_t turn, int (&act)[P]) {
    Philox rng(seed, game_id, /*stream=*/0x61637400u + turn, 0);
    uint4 r = make_uint4(0, 0, 0, 0);
#pragma unroll
    for (int p = 0; p < P; ++p) {
        if ((p & 3) == 0) r = rng.next<7>();
        const uint32_t word = (p & 3) == 0 ? r.x : (p & 3) == 1 ? r.y : (p & 3) == 2 ? r.z : r.w;
        const uint32_t n = (uint32_t)hand_count(g.hand[p]);
        act[p] = n ? (int)hand_select(g.hand[p], below(word, n)) : 255;
    }
}

}  // namespace nimmt
