// host_sim.cu — TEST-ONLY harness: runs the product's __host__ __device__ per-game logic
// (rl-6-nimmt_b200/csrc/game.cuh, step.cuh) on the CPU so that tests can diff it against the
// oracle without a GPU.  It is never linked into libnimmt_b200.so and never used as a fallback.
#include <cstdint>
#include <cstring>
#include <type_traits>

#include "../../rl-6-nimmt_b200/csrc/puct.cuh"
#include "../../rl-6-nimmt_b200/csrc/rollout.cuh"
#include "../../rl-6-nimmt_b200/csrc/step.cuh"
#include "../../rl-6-nimmt_b200/csrc/step_tile.cuh"

using namespace nimmt;

template <int P>
static void unpack_to_arrays(const Game<P>& g, int8_t* hands /*[P][10]*/, int8_t* board /*[4][6]*/, int16_t* scores) {
    for (int p = 0; p < P; ++p) {
        int n = 0;
        for (int c = 0; c < kCards; ++c)
            if (mask_has(g.hand[p], c)) hands[p * 10 + n++] = (int8_t)c;
        for (; n < 10; ++n) hands[p * 10 + n] = -1;
        scores[p] = (int16_t)(g.hand[p].w >> kScoreShift);
    }
    // go through pack/unpack so the stored representation is what gets checked
    uint64_t q0, q1, q2;
    g.board.pack(q0, q1, q2);
    Board b;
    b.unpack(q0, q1, q2);
    for (int r = 0; r < kRows; ++r) {
        const int len = b.k.len(r);
        for (int i = 0; i < 6; ++i) board[r * 6 + i] = i < len ? (int8_t)((b.cards[r] >> (8 * i)) & 0xFF) : -1;
    }
}

// The stored form (GameRec, handrec.cuh) goes through the same checks: converted to card sets for reading,
// built from card sets for initialisation.
template <int P>
static void unpack_to_arrays(const GameRec<P>& r, int8_t* hands, int8_t* board, int16_t* scores) {
    Game<P> g;
    for (int p = 0; p < P; ++p) g.hand[p] = rec_to_mask(r.hand[p]);
    g.board = r.board;
    unpack_to_arrays<P>(g, hands, board, scores);
}

template <int P>
static void init_game(Game<P>& g, const int8_t* rows0 /*[4][6]*/, const int8_t* hands0 /*[P][10]*/);

template <int P>
static void init_game(GameRec<P>& r, const int8_t* rows0, const int8_t* hands0) {
    Game<P> g;
    init_game<P>(g, rows0, hands0);
    for (int p = 0; p < P; ++p) r.hand[p] = rec_from_mask(g.hand[p]);
    r.board = g.board;
}

static int g_form = 0;   // 0: card sets (Game<P>), 1: stored form (GameRec<P>), 2: in place in a tile record (step_tile.cuh, what k_step_tiles runs)

template <int P>
static void init_game(Game<P>& g, const int8_t* rows0 /*[4][6]*/, const int8_t* hands0 /*[P][10]*/) {
    for (int r = 0; r < kRows; ++r) {
        uint64_t cards = 0;
        uint32_t len = 0, sum = 0, top = 0;
        for (int i = 0; i < 6 && rows0[r * 6 + i] >= 0; ++i) {
            const uint32_t c = rows0[r * 6 + i];
            cards |= (uint64_t)c << (8 * len);
            ++len; sum += h_card_value[c]; top = c;
        }
        g.board.set_row(r, cards, top, len, sum);
    }
    for (int p = 0; p < P; ++p) {
        g.hand[p] = make_uint4(0, 0, 0, 0);
        for (int i = 0; i < 10; ++i)
            if (hands0[p * 10 + i] >= 0) mask_set(g.hand[p], hands0[p * 10 + i]);
    }
}

static const int8_t* g_row_choice = nullptr;   // [n][turns][P] or NULL: the free-row-choice mode

template <int P, class G>
static void replay(int n, int turns, const int8_t* rows0, const int8_t* hands0, const int8_t* actions, int8_t* rewards,
                   uint8_t* done, uint8_t* illegal, int8_t* hands, int8_t* boards, int16_t* scores) {
    for (int gi = 0; gi < n; ++gi) {
        G g;
        init_game<P>(g, rows0 + gi * 24, hands0 + gi * P * 10);
        for (int t = 0; t < turns; ++t) {
            const size_t gt = (size_t)gi * turns + t;
            int act[P], pen[P], choice[P];
            for (int p = 0; p < P; ++p) act[p] = (uint8_t)actions[gt * P + p];
            if (g_row_choice)
                for (int p = 0; p < P; ++p) choice[p] = (uint8_t)g_row_choice[gt * P + p];
            const bool legal = step_game<P>(g, act, h_card_value, pen, g_row_choice ? choice : nullptr);
            if (!legal)
                for (int p = 0; p < P; ++p) pen[p] = 0;
            illegal[gt] = !legal;
            done[gt] = game_done<P>(g);
            for (int p = 0; p < P; ++p) rewards[gt * P + p] = (int8_t)(-pen[p]);
            unpack_to_arrays<P>(g, hands + gt * P * 10, boards + gt * 24, scores + gt * P);
        }
    }
}


// ---- form 2: the games live in tile records (32 games per tile, exactly the HBM / shared-memory layout) and are stepped in
// place by step_tile.cuh::step_lane, the per-lane code of k_step_tiles ----
template <int P>
static void tile_put(uint8_t* tile, int lane, const GameRec<P>& r) {
    using L = TileLayout<P>;
    for (int p = 0; p < P; ++p) {
        reinterpret_cast<uint2*>(tile)[p * kTileGames + lane] = r.hand[p].lo;
        reinterpret_cast<uint32_t*>(tile + L::kMeta)[p * kTileGames + lane] = r.hand[p].meta;
    }
    uint64_t q[3];
    r.board.pack(q[0], q[1], q[2]);
    memcpy(tile + L::kRows + lane * 24, q, 24);
}
template <int P>
static void tile_get(const uint8_t* tile, int lane, GameRec<P>& r) {
    using L = TileLayout<P>;
    for (int p = 0; p < P; ++p) {
        r.hand[p].lo = reinterpret_cast<const uint2*>(tile)[p * kTileGames + lane];
        r.hand[p].meta = reinterpret_cast<const uint32_t*>(tile + L::kMeta)[p * kTileGames + lane];
    }
    uint64_t q[3];
    memcpy(q, tile + L::kRows + lane * 24, 24);
    r.board.unpack(q[0], q[1], q[2]);
}

// kRandom: `actions` is OUTPUT (the cards the fused random step drew), keyed (seed, game0 + game, turn0 + t).
template <int P, bool kRandom, bool kChoice = false, bool kPacked = false>
static void replay_tiles(int n, int turns, const int8_t* rows0, const int8_t* hands0, int8_t* actions, int8_t* rewards, uint8_t* done,
                         uint8_t* illegal, int8_t* hands, int8_t* boards, int16_t* scores, uint64_t seed, uint64_t game0) {
    using L = TileLayout<P>;
    alignas(16) static uint8_t tile[L::kTileBytes];
    alignas(16) uint8_t acts[L::kActBytes], chosen[L::kActBytes];
    uint8_t values5[128];
    for (int c = 0; c < 128; ++c) values5[c] = (uint8_t)(h_card_value[c] << 5);
    for (int first = 0; first < n; first += kTileGames) {
        const int lanes = n - first < kTileGames ? n - first : kTileGames;
        memset(tile, 0, sizeof(tile));
        for (int lane = 0; lane < lanes; ++lane) {
            GameRec<P> r;
            init_game<P>(r, rows0 + (size_t)(first + lane) * 24, hands0 + (size_t)(first + lane) * P * 10);
            tile_put<P>(tile, lane, r);
        }
        for (int t = 0; t < turns; ++t) {
            for (int lane = 0; lane < lanes; ++lane) {
                const size_t gt = (size_t)(first + lane) * turns + t;
                if (!kRandom && !kPacked)
                    for (int p = 0; p < P; ++p) acts[lane * P + p] = (uint8_t)actions[gt * P + p];
                if (kPacked) {   // the transfer format: the card's slot in the hand as dealt (hands0, ascending), 15 if it was never there
                    for (int b = 0; b < packed_action_bytes<P>(); ++b) acts[lane * packed_action_bytes<P>() + b] = 0;
                    for (int p = 0; p < P; ++p) {
                        int slot = 15;
                        for (int i = 0; i < 10; ++i)
                            if (hands0[((size_t)(first + lane) * P + p) * 10 + i] == actions[gt * P + p] && actions[gt * P + p] >= 0) slot = i;
                        acts[lane * packed_action_bytes<P>() + (p >> 1)] |= (uint8_t)(slot << (4 * (p & 1)));
                    }
                }
                if (kChoice)
                    for (int p = 0; p < P; ++p) chosen[lane * P + p] = (uint8_t)g_row_choice[gt * P + p];
                alignas(16) uint32_t kw[4], ku[4];
                uint8_t rew[P + 8], dn = 0, ill = 0, drawn[P];
                step_lane<P, kRandom, kChoice, kPacked>(tile, acts, lane, values5, kw, ku, rew, &dn, &ill, kRandom ? drawn : nullptr, seed,
                                                        game0 + (uint64_t)(first + lane), (uint32_t)t, chosen, h_select8);
                if (kPacked) {   // unpack the bit record: 5 bits of bull heads per player, done, illegal
                    uint64_t rec = 0;
                    for (int b = 0; b < packed_result_bytes<P>(); ++b) rec |= (uint64_t)rew[b] << (8 * b);
                    for (int p = 0; p < P; ++p) rew[p] = (uint8_t)(0 - (int)((rec >> (5 * p)) & 31u));
                    dn = (uint8_t)((rec >> (5 * P)) & 1u);
                    ill = (uint8_t)((rec >> (5 * P + 1)) & 1u);
                }
                for (int p = 0; p < P; ++p) {
                    rewards[gt * P + p] = (int8_t)rew[p];
                    if (kRandom) actions[gt * P + p] = (int8_t)drawn[p];
                }
                done[gt] = dn;
                illegal[gt] = ill;
                GameRec<P> r;
                tile_get<P>(tile, lane, r);
                unpack_to_arrays<P>(r, hands + gt * P * 10, boards + gt * 24, scores + gt * P);
            }
        }
    }
}

template <int P, class G>
static void deal(int n, uint64_t seed, uint64_t game0, int8_t* hands, int8_t* boards) {
    int16_t sc[P];
    for (int gi = 0; gi < n; ++gi) {
        G g;
        uint8_t deck[kCards];
        for (int c = 0; c < kCards; ++c) deck[c] = (uint8_t)c;
        if constexpr (std::is_same<G, Game<P>>::value) deal_game<P>(seed, game0 + gi, deck, 1, DealIntoGame<P>{g, h_card_value});
        else deal_game<P>(seed, game0 + gi, deck, 1, DealIntoGameRec<P>{g, h_card_value});
        unpack_to_arrays<P>(g, hands + (size_t)gi * P * 10, boards + (size_t)gi * 24, sc);
    }
}

template <int P, class G>
static void rand_act(int n, const int8_t* rows0, const int8_t* hands0, uint64_t seed, uint64_t game0, uint32_t turn, uint8_t* actions) {
    for (int gi = 0; gi < n; ++gi) {
        G g;
        init_game<P>(g, rows0 + gi * 24, hands0 + gi * P * 10);
        int act[P];
        random_actions_game<P>(g, seed, game0 + gi, turn, act);
        for (int p = 0; p < P; ++p) actions[(size_t)gi * P + p] = (uint8_t)act[p];
    }
}

#define DISPATCH(P_, CALL)                                  \
    switch (P_) {                                           \
        case 1: { constexpr int P = 1; CALL; } break;       \
        case 2: { constexpr int P = 2; CALL; } break;       \
        case 3: { constexpr int P = 3; CALL; } break;       \
        case 4: { constexpr int P = 4; CALL; } break;       \
        case 5: { constexpr int P = 5; CALL; } break;       \
        case 6: { constexpr int P = 6; CALL; } break;       \
        case 7: { constexpr int P = 7; CALL; } break;       \
        case 8: { constexpr int P = 8; CALL; } break;       \
        case 9: { constexpr int P = 9; CALL; } break;       \
        case 10: { constexpr int P = 10; CALL; } break;     \
        default: return -1;                                 \
    }

template <int N>
static int check_network() {
    // 0-1 principle: a comparator network sorts everything iff it sorts all 2^N bit vectors
    for (unsigned m = 0; m < (1u << N); ++m) {
        int k[N];
        for (int i = 0; i < N; ++i) k[i] = (m >> i) & 1;
        sort_keys<N>(k);
        for (int i = 1; i < N; ++i)
            if (k[i - 1] > k[i]) return N;
    }
    return 0;
}

template <int P>
static int mcs(const nimmt_root& root, int64_t R, uint64_t seed, int rank, int world, int64_t* stats) {
    RolloutRoot rr;
    if (!make_rollout_root<P>(root, h_card_value, rr)) return -2;
    alignas(4) uint8_t deck[kRolloutDeckStride];
    uint8_t values5[128];
    for (int c = 0; c < 128; ++c) values5[c] = (uint8_t)(h_card_value[c] << 5);
    const int n = rr.n_own;
    for (int a = 0; a < n; ++a) {
        for (int64_t j = rank; j < R; j += world) {
            const uint64_t id = ((uint64_t)a << 40) | (uint64_t)j;  // root index d = 0
            alignas(16) uint32_t kw[4], ku[4];
            const int out = rollout<P>(rr, a, values5, deck, kw, ku, seed, id);
            stats[a * 3 + 0] += out; stats[a * 3 + 1] += (int64_t)out * out; stats[a * 3 + 2] += 1;
        }
    }
    return 0;
}

extern "C" {
// PUCT root rule on explicit outcome lists: (action index, outcome) pairs in visiting order.
int sim_puct(int n, int n_outcomes, const int* action_index, const int* outcome, const float* probs, float c_puct, double* pucts) {
    RootStats s;
    root_stats_clear(s);
    for (int i = 0; i < n_outcomes; ++i) root_stats_add(s, action_index[i], outcome[i]);
    return puct_choose(s, probs, n, c_puct, pucts);
}
int sim_mcs(int P_, const nimmt_root* root, int64_t R, uint64_t seed, int rank, int world, int64_t* stats) {
    DISPATCH(P_, return mcs<P>(*root, R, seed, rank, world, stats));
    return 0;
}
int sim_replay(int P_, int n, int turns, const int8_t* rows0, const int8_t* hands0, const int8_t* actions, int8_t* rewards,
               uint8_t* done, uint8_t* illegal, int8_t* hands, int8_t* boards, int16_t* scores) {
    if (g_form == 3) {
        DISPATCH(P_, (replay_tiles<P, false, false, true>(n, turns, rows0, hands0, const_cast<int8_t*>(actions), rewards, done, illegal, hands, boards, scores, 0, 0)));
    } else if (g_form == 2) {
        if (g_row_choice) {
            DISPATCH(P_, (replay_tiles<P, false, true>(n, turns, rows0, hands0, const_cast<int8_t*>(actions), rewards, done, illegal, hands, boards, scores, 0, 0)));
        } else {
            DISPATCH(P_, (replay_tiles<P, false>(n, turns, rows0, hands0, const_cast<int8_t*>(actions), rewards, done, illegal, hands, boards, scores, 0, 0)));
        }
    } else if (g_form) { DISPATCH(P_, (replay<P, GameRec<P>>(n, turns, rows0, hands0, actions, rewards, done, illegal, hands, boards, scores))); }
    else { DISPATCH(P_, (replay<P, Game<P>>(n, turns, rows0, hands0, actions, rewards, done, illegal, hands, boards, scores))); }
    return 0;
}
// Random-vs-random play through the fused per-lane step (step_lane<P, true>): `actions` receives the cards drawn.
int sim_play_random_tiles(int P_, int n, int turns, const int8_t* rows0, const int8_t* hands0, uint64_t seed, uint64_t game0, int8_t* actions,
                          int8_t* rewards, uint8_t* done, uint8_t* illegal, int8_t* hands, int8_t* boards, int16_t* scores) {
    DISPATCH(P_, (replay_tiles<P, true>(n, turns, rows0, hands0, actions, rewards, done, illegal, hands, boards, scores, seed, game0)));
    return 0;
}
int sim_deal(int P_, int n, uint64_t seed, uint64_t game0, int8_t* hands, int8_t* boards) {
    if (g_form) { DISPATCH(P_, (deal<P, GameRec<P>>(n, seed, game0, hands, boards))); }
    else { DISPATCH(P_, (deal<P, Game<P>>(n, seed, game0, hands, boards))); }
    return 0;
}
int sim_random_actions(int P_, int n, const int8_t* rows0, const int8_t* hands0, uint64_t seed, uint64_t game0, uint32_t turn,
                       uint8_t* actions) {
    if (g_form) { DISPATCH(P_, (rand_act<P, GameRec<P>>(n, rows0, hands0, seed, game0, turn, actions))); }
    else { DISPATCH(P_, (rand_act<P, Game<P>>(n, rows0, hands0, seed, game0, turn, actions))); }
    return 0;
}
void sim_set_form(int form) { g_form = form; }
void sim_set_row_choice(const int8_t* row_choice) { g_row_choice = row_choice; }
// handrec.cuh directly: a hand of n ascending cards with `played` slots already empty; for every card id 0..255 the slot
// rec_find reports (-1: not dealt), whether rec_take accepts it and the meta word it would commit; then the record's
// views: rec_card per slot, rec_count, rec_to_mask words, rec_select for every k.
void sim_handrec(const uint8_t* cards, int n, uint32_t played, uint32_t score, int* find, uint8_t* take_ok, uint32_t* take_meta,
                 uint8_t* slot_card, int* count, uint32_t* mask, uint8_t* select) {
    uint32_t c[kHand];
    for (int i = 0; i < kHand; ++i) c[i] = i < n ? cards[i] : 0;
    HandRec h = rec_from_sorted(c, n, score);
    h.meta |= rec_spread(played & kSlotBits);
    for (int card = 0; card < 256; ++card) {
        find[card] = rec_find(h, (uint32_t)card);
        uint32_t meta = 0;
        take_ok[card] = rec_take(h, (uint32_t)card, meta);
        take_meta[card] = rec_empties(meta) | ((meta >> kRecScoreShift) << 10);   // canonical view: slot-order empty bits | score << 10
    }
    for (int i = 0; i < kHand; ++i) slot_card[i] = (uint8_t)rec_card(h, i);
    *count = rec_count(h);
    const uint4 m = rec_to_mask(h);
    mask[0] = m.x; mask[1] = m.y; mask[2] = m.z; mask[3] = m.w;
    for (int k = 0; k < *count; ++k) select[k] = (uint8_t)rec_select(h, (uint32_t)k);
    // and back: a record rebuilt from the card set holds the same cards in the same order
    const HandRec back = rec_from_mask(m);
    for (int k = 0; k < *count; ++k)
        if (rec_card(back, k) != select[k] || rec_score(back) != score) *count = -1;
}
// source[i] = original position of the entry that Fisher-Yates step i outputs, for steps 0..n-1 (n <= 90)
void sim_fisher_yates_sources(const uint8_t* target, int n, int* source) {
    for (int i = 0; i < n; ++i) source[i] = fisher_yates_source<90>(target, i);
}
int sim_check_sort_networks() {
    int bad = 0;
    bad |= check_network<2>(); bad |= check_network<3>(); bad |= check_network<4>(); bad |= check_network<5>();
    bad |= check_network<6>(); bad |= check_network<7>(); bad |= check_network<8>(); bad |= check_network<9>();
    bad |= check_network<10>();
    return bad;
}
unsigned sim_select(uint32_t x, uint32_t y, uint32_t z, uint32_t w, uint32_t k) { return mask_select(make_uint4(x, y, z, w), k); }
}
