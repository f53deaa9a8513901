// game.cuh — device-side primitives of the 6 nimmt! hot path (sm_100a).
//
// Everything here is register-resident integer code: card sets as 128-bit words, rows as two comparison keys
// per row plus a 24-byte record; handrec.cuh adds the stored form of a hand (include/nimmt_b200.h).  No dynamic register indexing: every
// data-dependent row / player choice is a predicated select, so nothing spills to local memory
// and a warp of 32 independent games never diverges inside a placement.
//
// Reference semantics: rl_6_nimmt/env.py:64-77,114-172,214-249 (restated in SURVEY.md §3.5).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

// The per-game logic is __host__ __device__ so that tests/host_sim can run the very same code on
// the CPU against the oracle before GPU time is spent (SURVEY.md §7 hard part 9).  That harness is
// test-only: libnimmt_b200.so exports no host implementation of any entry point.
#define NIMMT_HD __host__ __device__ __forceinline__

// -DNIMMT_BOUNDS_CHECK: every data-dependent shared-memory index of the hot kernels is asserted in range (a device assert
// traps the launch).  compute-sanitizer is closed on this pool's B200s, so this build — run over the small cases of
// profiles/tools/sanitize_small.py and the GPU test suite — is the memory-safety evidence (profiles/README.md).
#ifdef NIMMT_BOUNDS_CHECK
#include <cassert>
#define NIMMT_CHECK(cond) assert(cond)
#else
#define NIMMT_CHECK(cond) ((void)0)
#endif

namespace nimmt {

NIMMT_HD int popc32(uint32_t x) {
#ifdef __CUDA_ARCH__
    return __popc(x);
#else
    return __builtin_popcount(x);
#endif
}
NIMMT_HD uint32_t umulhi32(uint32_t a, uint32_t b) {
#ifdef __CUDA_ARCH__
    return __umulhi(a, b);
#else
    return (uint32_t)(((uint64_t)a * b) >> 32);
#endif
}
// PRMT: byte i of the result = byte (sel >> 4 i) & 7 of the 8-byte pair {b, a} (the sign-replication mode is not used here).
NIMMT_HD uint32_t byte_perm(uint32_t a, uint32_t b, uint32_t sel) {
#ifdef __CUDA_ARCH__
    return __byte_perm(a, b, sel);
#else
    const uint64_t v = ((uint64_t)b << 32) | a;
    uint32_t r = 0;
    for (int i = 0; i < 4; ++i) r |= (uint32_t)((v >> (8 * ((sel >> (4 * i)) & 7u))) & 0xFFu) << (8 * i);
    return r;
#endif
}
NIMMT_HD int imin(int a, int b) { return a < b ? a : b; }
NIMMT_HD int imax(int a, int b) { return a > b ? a : b; }
NIMMT_HD uint32_t umin32(uint32_t a, uint32_t b) { return a < b ? a : b; }

constexpr int kRows = 4;
constexpr int kCards = 104;
constexpr int kHand = 10;
constexpr int kMaxPlayers = 10;
constexpr int kScoreShift = 24;  // score byte = bits 24..31 of hand.w (bits 120..127 of the word)

// SechsNimmtEnv._card_value (env.py:224-239) for cards 0..103.
#define NIMMT_CARD_VALUES                                                                          \
    1, 1, 1, 1, 2, 1, 1, 1, 1, 3, 5, 1, 1, 1, 2, 1, 1, 1, 1, 3, 1, 5, 1, 1, 2, 1, 1, 1, 1, 3, 1, 1, \
    5, 1, 2, 1, 1, 1, 1, 3, 1, 1, 1, 5, 2, 1, 1, 1, 1, 3, 1, 1, 1, 1, 7, 1, 1, 1, 1, 3, 1, 1, 1, 1, \
    2, 5, 1, 1, 1, 3, 1, 1, 1, 1, 2, 1, 5, 1, 1, 3, 1, 1, 1, 1, 2, 1, 1, 5, 1, 3, 1, 1, 1, 1, 2, 1, \
    1, 1, 5, 3, 1, 1, 1, 1

__device__ __constant__ uint8_t c_card_value[128] = {NIMMT_CARD_VALUES};
static const uint8_t h_card_value[128] = {NIMMT_CARD_VALUES};  // host copy (nimmt_card_value, host_sim)

// The same table with every value << 5 (place_v3's unit), built in shared memory from the constant table.
__device__ __forceinline__ void stage_card_values5(uint8_t* smem) {
    for (uint32_t c = threadIdx.x; c < 128u; c += blockDim.x) smem[c] = (uint8_t)(c_card_value[c] << 5);   // any block size
}

// Copies the 104-entry value table into shared memory (26 words -> 26 distinct banks, so a warp
// of random lookups is conflict-free).  `smem` must hold 128 bytes.  Call before __syncthreads.
__device__ __forceinline__ void stage_card_values(uint8_t* smem) {
    const uint32_t* src = reinterpret_cast<const uint32_t*>(c_card_value);
    uint32_t* dst = reinterpret_cast<uint32_t*>(smem);
    if (threadIdx.x < 32) dst[threadIdx.x] = src[threadIdx.x];
}

// ----------------------------------------------------------------------------------------------
// Philox4x32-10 counter RNG (Salmon et al. 2011).  Keyed by the caller's seed; the counter is
// (global game / rollout id, stream, block) so that results never depend on launch geometry.
// ----------------------------------------------------------------------------------------------
struct Philox {
    uint32_t c[4];
    uint32_t k[2];
    NIMMT_HD Philox(uint64_t seed, uint64_t id, uint32_t stream, uint32_t block) {
        c[0] = (uint32_t)id;
        c[1] = (uint32_t)(id >> 32);
        c[2] = stream;
        c[3] = block;
        k[0] = (uint32_t)seed;
        k[1] = (uint32_t)(seed >> 32);
    }
    // Returns 4 fresh 32-bit words and advances the block counter.  ROUNDS = 10 is the standard
    // generator; 7 is the smallest round count that still passes BigCrush (Salmon et al. 2011, table 2)
    // and is used where the RNG is a large share of the work (rollouts).
    template <int ROUNDS = 10>
    NIMMT_HD uint4 next() {
        uint32_t x0 = c[0], x1 = c[1], x2 = c[2], x3 = c[3], k0 = k[0], k1 = k[1];
#pragma unroll
        for (int r = 0; r < ROUNDS; ++r) {
            const uint32_t lo0 = 0xD2511F53u * x0, hi0 = umulhi32(0xD2511F53u, x0);
            const uint32_t lo1 = 0xCD9E8D57u * x2, hi1 = umulhi32(0xCD9E8D57u, x2);
            x0 = hi1 ^ x1 ^ k0;
            x1 = lo1;
            x2 = hi0 ^ x3 ^ k1;
            x3 = lo0;
            k0 += 0x9E3779B9u;
            k1 += 0xBB67AE85u;
        }
        c[3] += 1;
        return make_uint4(x0, x1, x2, x3);
    }
};

// Uniform integer in [0, n) from 32 random bits by multiply-high.  Bias <= n / 2^32 (< 2.5e-8 for
// n <= 104): far below any test in this repo can resolve; stated in DESIGN.md §6.
NIMMT_HD uint32_t below(uint32_t r, uint32_t n) { return umulhi32(r, n); }

// Two uniform integers from one 32-bit word: the high half of r * n is the draw, the low half — the position of r
// within its bucket, spread over the full 32-bit range on a lattice of spacing n — serves as the random word of a second
// draw.  Joint bias <= n1 * n2 / 2^32 (< 2.6e-6 for n <= 104).  `r` is replaced by the low half.
NIMMT_HD uint32_t below_keep(uint32_t& r, uint32_t n) {
    const uint64_t prod = (uint64_t)r * n;
    r = (uint32_t)prod;
    return (uint32_t)(prod >> 32);
}

// The random words of one turn of a P-player game: two draws per word, so ceil(P / 2) words per turn, taken from one
// Philox4x32-7 call per floor(4 / words) turns (P = 4: one call per two turns; P >= 9 needs two calls per turn).
template <int P>
struct TurnWords {
    static constexpr int kWords = (P + 1) / 2;
    static constexpr int kTurnsPerCall = kWords <= 4 ? 4 / kWords : 1;
    uint4 r;
    uint32_t w[kWords];
    NIMMT_HD TurnWords() : r(make_uint4(0, 0, 0, 0)) {}
    template <class Rng>
    NIMMT_HD void begin_turn(Rng& rng, int turn) {
        const int phase = turn % kTurnsPerCall;
        if (phase == 0) r = rng.template next<7>();
        const uint32_t c[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
        for (int i = 0; i < kWords && i < 4; ++i) {
            uint32_t v = c[i];                                 // phase 0
#pragma unroll
            for (int ph = 1; ph < kTurnsPerCall; ++ph) v = phase == ph ? c[(ph * kWords + i) & 3] : v;
            w[i] = v;
        }
        if constexpr (kWords > 4) w[4] = rng.template next<7>().x;
    }
    // draw number i of the turn (compile-time i), uniform in [0, n)
    template <int I>
    NIMMT_HD uint32_t draw(uint32_t n) {
        if constexpr ((I & 1) == 0) return below_keep(w[I >> 1], n);
        else return below(w[I >> 1], n);
    }
};

// ----------------------------------------------------------------------------------------------
// 104-bit card sets in a uint4 (x = cards 0..31, y = 32..63, z = 64..95, w bits 0..7 = 96..103).
// ----------------------------------------------------------------------------------------------
constexpr uint32_t kHighCardMask = 0xFFu;  // card bits living in .w

NIMMT_HD uint32_t mask_word(const uint4& m, uint32_t word) {
    uint32_t v = m.x;
    v = word == 1 ? m.y : v;
    v = word == 2 ? m.z : v;
    v = word == 3 ? (m.w & kHighCardMask) : v;
    return v;
}

NIMMT_HD bool mask_has(const uint4& m, uint32_t card) {
    return card < (uint32_t)kCards && ((mask_word(m, card >> 5) >> (card & 31)) & 1u);
}

NIMMT_HD void mask_clear(uint4& m, uint32_t card) {
    const uint32_t word = card >> 5, bit = 1u << (card & 31);
    m.x &= ~(word == 0 ? bit : 0u);
    m.y &= ~(word == 1 ? bit : 0u);
    m.z &= ~(word == 2 ? bit : 0u);
    m.w &= ~(word == 3 ? bit : 0u);
}

// Removes `card` from the set; returns whether it was there (false also for card ids >= 104).
// One decode of the card into four word masks serves both the test and the removal.
NIMMT_HD bool mask_take(uint4& m, uint32_t card, bool commit) {
    const uint32_t word = card >> 5, bit = 1u << (card & 31);
    const uint32_t b0 = word == 0 ? bit : 0u, b1 = word == 1 ? bit : 0u, b2 = word == 2 ? bit : 0u;
    const uint32_t b3 = word == 3 ? (bit & kHighCardMask) : 0u;
    const bool had = ((m.x & b0) | (m.y & b1) | (m.z & b2) | (m.w & b3)) != 0u;
    if (commit) { m.x &= ~b0; m.y &= ~b1; m.z &= ~b2; m.w &= ~b3; }
    return had;
}

NIMMT_HD void mask_set(uint4& m, uint32_t card) {
    const uint32_t word = card >> 5, bit = 1u << (card & 31);
    m.x |= (word == 0 ? bit : 0u);
    m.y |= (word == 1 ? bit : 0u);
    m.z |= (word == 2 ? bit : 0u);
    m.w |= (word == 3 ? bit : 0u);
}

NIMMT_HD int mask_count(const uint4& m) {
    return popc32(m.x) + popc32(m.y) + popc32(m.z) + popc32(m.w & kHighCardMask);
}

// Position of the k-th (0-based) set bit of a non-zero 32-bit word, k < popc(w): 5-level
// popcount bisection (no __fns: it is a slow emulated loop).
NIMMT_HD uint32_t select_bit32(uint32_t w, uint32_t k) {
    uint32_t pos = 0, c;
    c = popc32(w & 0xFFFFu);
    if (k >= c) { k -= c; pos = 16; w >>= 16; }
    c = popc32(w & 0xFFu);
    if (k >= c) { k -= c; pos += 8; w >>= 8; }
    c = popc32(w & 0xFu);
    if (k >= c) { k -= c; pos += 4; w >>= 4; }
    c = popc32(w & 0x3u);
    if (k >= c) { k -= c; pos += 2; w >>= 2; }
    c = w & 1u;
    if (k >= c) pos += 1;
    return pos;
}

// Card id of the k-th smallest card in the set, k < mask_count(m).
NIMMT_HD uint32_t mask_select(const uint4& m, uint32_t k) {
    const uint32_t c0 = popc32(m.x), c1 = c0 + popc32(m.y), c2 = c1 + popc32(m.z);
    uint32_t w = m.x, base = 0, skip = 0;
    if (k >= c0) { w = m.y; base = 32; skip = c0; }
    if (k >= c1) { w = m.z; base = 64; skip = c1; }
    if (k >= c2) { w = m.w & kHighCardMask; base = 96; skip = c2; }
    return base + select_bit32(w, k - skip);
}

// ----------------------------------------------------------------------------------------------
// Rows in registers.
//
// RowKeys is everything the dynamics depend on (env.py:138-172): per row the top card, the number
// of cards and the bull-head sum, folded into two comparison keys so that a placement needs no
// gather over rows:
//   w[r] = top << 10 | sum << 5 | len << 2 | r     "largest top below the card" is max over the
//                                                   rows with w < card << 10, and the winner's
//                                                   sum, len and index come along in its low bits
//   u[r] = sum << 2 | r                             the undercut rule (lowest sum, lowest index on
//                                                   ties) is min over u
// Board adds the card lists (needed for observations / bit-exact state, not for the dynamics):
//   cards[r] = byte i = i-th card of the row, oldest first; bytes at index >= len are unspecified
//              (zero after a deal / reset_to, stale cards after a take).
// ----------------------------------------------------------------------------------------------
struct RowKeys {
    int w[kRows];
    int u[kRows];

    NIMMT_HD void set_row(int r, uint32_t top, uint32_t len, uint32_t sum) {
        w[r] = (int)((top << 10) | (sum << 5) | (len << 2) | (uint32_t)r);
        u[r] = (int)((sum << 2) | (uint32_t)r);
    }
    NIMMT_HD int top(int r) const { return w[r] >> 10; }
    NIMMT_HD int len(int r) const { return (w[r] >> 2) & 7; }
    NIMMT_HD int sum(int r) const { return (w[r] >> 5) & 31; }

    // One placement (env.py:126-134 + _find_row :138-152 + _pick_row_to_replace :154-159 +
    // _score_row :161-172).  Returns the bull heads taken (0 if none); `row` and `keep_len` tell the
    // caller where the card went (keep_len = cards of the row that stay under it: 0 after a take).
    //   undercut (card below every top) -> row with the smallest bull-head sum, lowest index on
    //                                      ties; the player takes the whole old row
    //   otherwise                       -> row with the largest top below the card; if it already
    //                                      holds five cards the player takes those five
    // In both take cases the penalty is the row's sum BEFORE the append and the row restarts with
    // the played card alone.
    // `choice` >= 0: the free-row-choice mode (env.py:156 TODO) — on an undercut the player takes row `choice` instead of the
    // cheapest one.
    NIMMT_HD int place(int card, int value, int& row, uint32_t& keep_len, int choice = -1) {
        const int ck = card << 10;
        // largest w below ck == smallest positive ck - w; rows above the card wrap to huge unsigned values
        const uint32_t d0 = (uint32_t)(ck - w[0]), d1 = (uint32_t)(ck - w[1]);
        const uint32_t d2 = (uint32_t)(ck - w[2]), d3 = (uint32_t)(ck - w[3]);
        const uint32_t dmin = umin32(umin32(d0, d1), umin32(d2, d3));
        const int best = ck - (int)dmin;
        int cheapest = imin(imin(u[0], u[1]), imin(u[2], u[3]));
        if (choice >= 0) cheapest = choice == 0 ? u[0] : choice == 1 ? u[1] : choice == 2 ? u[2] : u[3];
        const bool under = dmin > (uint32_t)ck;   // card below every top
        const int r = (under ? cheapest : best) & 3;
        const uint32_t len = ((uint32_t)best >> 2) & 7u;  // garbage when under; take is true then
        const uint32_t sum = under ? (uint32_t)cheapest >> 2 : ((uint32_t)best >> 5) & 31u;
        const bool take = under || len == 5u;
        keep_len = take ? 0u : len;
        const uint32_t new_sum = (take ? 0u : sum) + (uint32_t)value;
        const int new_w = ck | (int)((new_sum << 5) | ((keep_len + 1u) << 2)) | r;
        const int new_u = (int)(new_sum << 2) | r;
#pragma unroll
        for (int i = 0; i < kRows; ++i) {  // selects, not branches: a warp's 32 games pick different rows
            const bool hit = r == i;
            w[i] = hit ? new_w : w[i];
            u[i] = hit ? new_u : u[i];
        }
        row = r;
        return take ? (int)sum : 0;
    }
};

// ----------------------------------------------------------------------------------------------
// Placement on row keys that live in INDEXABLE memory (this thread's four w and four u words, each group 16-byte aligned: shared
// memory on the device; STRIDE = n interleaves the threads of a block, row r of this thread at word r * n, so that the store to
// the data-dependent row is free of bank conflicts) — the throughput kernels k_step_tiles and k_mcs_rollouts.  The four rows are
// read with two 128-bit loads and the one row that changes is written with two 32-bit stores, instead of the eight predicated
// selects (and four compares) that keeping the rows in registers costs.  Same rule as RowKeys::place, with the bit
// fields arranged so that almost nothing has to be shifted:
//   W[r] = top << 10 | sum << 5 | len << 2 | r        as before
//   U[r] =             sum << 5            | r        the undercut key, its fields ALIGNED with W's
//   key  = card << 10 | anything below bit 10         the caller keeps the player there (it never reaches W)
//   values5[c] = bull heads of card c, << 5           (<= 7 << 5 = 224: still a byte)
// so "which row" is sel & 3 and "the old row's bull heads" is sel & 0x3E0 whether sel came from the undercut minimum or
// from the best row, 4 * len + r — the byte of the 24-byte row record the card lands in — is best & 0x1F, and every
// new field is added, not shifted and or-ed.  Returns the bull heads taken, << 5 (0 if none); `row` = the row; `keep4` =
// 4 * (cards of the row that stay under the new card) = byte offset of the card within the record, minus `row`.
// ----------------------------------------------------------------------------------------------
constexpr uint32_t kSumField = 0x3E0u, kLenField = 0x1Cu;

NIMMT_HD uint32_t key_u_from_w(uint32_t w) { return w & (kSumField | 3u); }

// kChoice: the free-row-choice mode (env.py:156 TODO, README.md:11) — on an undercut the row is `choice` (0..3), not the cheapest.
template <int STRIDE = 1, bool kChoice = false>
NIMMT_HD uint32_t place_v3(uint32_t* w, uint32_t* u, uint32_t key, const uint8_t* values5, uint32_t& row, uint32_t& keep4, uint32_t choice = 0) {
    uint4 W, U;
    if constexpr (STRIDE == 1) {
        W = *reinterpret_cast<const uint4*>(w);
        U = *reinterpret_cast<const uint4*>(u);
    } else {
        W = make_uint4(w[0], w[STRIDE], w[2 * STRIDE], w[3 * STRIDE]);
        U = make_uint4(u[0], u[STRIDE], u[2 * STRIDE], u[3 * STRIDE]);
    }
    const uint32_t dmin = umin32(umin32(key - W.x, key - W.y), umin32(key - W.z, key - W.w));   // rows above the card wrap to huge values
    const uint32_t cheapest = kChoice ? u[choice * STRIDE] : umin32(umin32(U.x, U.y), umin32(U.z, U.w));
    const bool under = dmin > key;                          // card below every top (env.py:143)
    const uint32_t best = key - dmin;                       // == W of the row with the largest top below the card
    const uint32_t sel = under ? cheapest : best;
    const uint32_t r = sel & 3u, sumf = sel & kSumField, len4 = best & kLenField;
    const bool take = under || len4 == 20u;                 // env.py:133: replaced, or the sixth card
    keep4 = take ? 0u : len4;
    const uint32_t base = take ? 0u : sumf;
    NIMMT_CHECK((key >> 10) < (uint32_t)kCards && r < (uint32_t)kRows && keep4 <= 16u && (!kChoice || choice < (uint32_t)kRows));
    const uint32_t new_u = base + values5[key >> 10] + r;
    w[r * STRIDE] = (key & ~1023u) + new_u + keep4 + 4u;
    u[r * STRIDE] = new_u;
    row = r;
    return sumf - base;                                     // the row's sum BEFORE the append (env.py:164), << 5
}

struct Board {
    RowKeys k;
    uint64_t cards[kRows];

    NIMMT_HD void set_row(int r, uint64_t row_cards, uint32_t top, uint32_t len, uint32_t sum) {
        k.set_row(r, top, len, sum);
        cards[r] = row_cards;
    }

    // The stored 24-byte row record is slot-major: byte 4 j + r = j-th card (oldest first) of row r
    // for j = 0..4, byte 20 + r = meta of row r (len | sum << 3).  A placement touches exactly one
    // byte of it (plus the meta bytes), which is what lets k_step_tiles update it in place.
    // q0 / q1 / q2 are its three little-endian 64-bit words.
    NIMMT_HD void unpack(uint64_t q0, uint64_t q1, uint64_t q2) {
        const uint32_t slot[5] = {(uint32_t)q0, (uint32_t)(q0 >> 32), (uint32_t)q1, (uint32_t)(q1 >> 32), (uint32_t)q2};
        const uint32_t metas = (uint32_t)(q2 >> 32);
#pragma unroll
        for (int r = 0; r < kRows; ++r) {
            uint64_t c = 0;
#pragma unroll
            for (int j = 0; j < 5; ++j) c |= (uint64_t)((slot[j] >> (8 * r)) & 0xFFu) << (8 * j);
            cards[r] = c;
            const uint32_t meta = (metas >> (8 * r)) & 0xFFu;
            const uint32_t len = meta & 7u;
            const uint32_t top = (uint32_t)(c >> (8u * (len - 1u))) & 0xFFu;
            k.set_row(r, top, len, meta >> 3);
        }
    }

    NIMMT_HD void pack(uint64_t& q0, uint64_t& q1, uint64_t& q2) const {
        uint32_t slot[5] = {0, 0, 0, 0, 0}, metas = 0;
#pragma unroll
        for (int r = 0; r < kRows; ++r) {
#pragma unroll
            for (int j = 0; j < 5; ++j) slot[j] |= (uint32_t)((cards[r] >> (8 * j)) & 0xFFu) << (8 * r);
            metas |= ((uint32_t)k.len(r) | ((uint32_t)k.sum(r) << 3)) << (8 * r);
        }
        q0 = (uint64_t)slot[0] | ((uint64_t)slot[1] << 32);
        q1 = (uint64_t)slot[2] | ((uint64_t)slot[3] << 32);
        q2 = (uint64_t)slot[4] | ((uint64_t)metas << 32);
    }

    NIMMT_HD int place(int card, int value, int choice = -1) {
        int r;
        uint32_t keep_len;
        const int penalty = k.place(card, value, r, keep_len, choice);
        // write the one slot the card lands in; slots at index >= len keep whatever they held
        // (they are unspecified in the stored record: k_step_tiles writes a single byte too)
        const uint32_t shift = 8u * keep_len;
        const uint64_t clear = ~(0xFFull << shift), put = (uint64_t)(uint32_t)card << shift;
#pragma unroll
        for (int i = 0; i < kRows; ++i) {
            const uint64_t written = (cards[i] & clear) | put;
            cards[i] = r == i ? written : cards[i];
        }
        return penalty;
    }
};

// ----------------------------------------------------------------------------------------------
// Sorting network for the (card << 4 | player) keys of one step (env.py:124-125).  Optimal-size
// networks for n <= 10 (Knuth TAOCP 3, 5.3.4); fully unrolled compare-exchange on registers.
// ----------------------------------------------------------------------------------------------
NIMMT_HD void cswap(int& a, int& b) {
    const int lo = imin(a, b), hi = imax(a, b);
    a = lo;
    b = hi;
}

// Two 16-bit keys per word, compared lane-wise (VIMNMX.U16x2): one network sorts two independent sequences at once.
struct Pair16 {
    uint32_t v;
};
NIMMT_HD void cswap(Pair16& a, Pair16& b) {
#ifdef __CUDA_ARCH__
    const uint32_t lo = __vminu2(a.v, b.v), hi = __vmaxu2(a.v, b.v);
#else
    const uint32_t al = a.v & 0xFFFFu, ah = a.v >> 16, bl = b.v & 0xFFFFu, bh = b.v >> 16;
    const uint32_t lo = (al < bl ? al : bl) | ((ah < bh ? ah : bh) << 16), hi = (al < bl ? bl : al) | ((ah < bh ? bh : ah) << 16);
#endif
    a.v = lo;
    b.v = hi;
}

template <int N, class T>
NIMMT_HD void sort_keys(T (&k)[N]) {
#define CS(i, j) cswap(k[i], k[j])
    if constexpr (N == 2) {
        CS(0, 1);
    } else if constexpr (N == 3) {
        CS(0, 2); CS(0, 1); CS(1, 2);
    } else if constexpr (N == 4) {
        CS(0, 2); CS(1, 3); CS(0, 1); CS(2, 3); CS(1, 2);
    } else if constexpr (N == 5) {
        CS(0, 3); CS(1, 4); CS(0, 2); CS(1, 3); CS(0, 1); CS(2, 4); CS(1, 2); CS(3, 4); CS(2, 3);
    } else if constexpr (N == 6) {
        CS(0, 5); CS(1, 3); CS(2, 4); CS(1, 2); CS(3, 4); CS(0, 3); CS(2, 5); CS(0, 1); CS(2, 3);
        CS(4, 5); CS(1, 2); CS(3, 4);
    } else if constexpr (N == 7) {
        CS(0, 6); CS(2, 3); CS(4, 5); CS(0, 2); CS(1, 4); CS(3, 6); CS(0, 1); CS(2, 5); CS(3, 4);
        CS(1, 2); CS(4, 6); CS(2, 3); CS(4, 5); CS(1, 2); CS(3, 4); CS(5, 6);
    } else if constexpr (N == 8) {
        CS(0, 2); CS(1, 3); CS(4, 6); CS(5, 7); CS(0, 4); CS(1, 5); CS(2, 6); CS(3, 7); CS(0, 1);
        CS(2, 3); CS(4, 5); CS(6, 7); CS(2, 4); CS(3, 5); CS(1, 4); CS(3, 6); CS(1, 2); CS(3, 4);
        CS(5, 6);
    } else if constexpr (N == 9) {
        CS(0, 3); CS(1, 7); CS(2, 5); CS(4, 8); CS(0, 7); CS(2, 4); CS(3, 8); CS(5, 6); CS(0, 2);
        CS(1, 3); CS(4, 5); CS(7, 8); CS(1, 4); CS(3, 6); CS(5, 7); CS(0, 1); CS(2, 4); CS(3, 5);
        CS(6, 8); CS(2, 3); CS(4, 5); CS(6, 7); CS(1, 2); CS(3, 4); CS(5, 6);
    } else if constexpr (N == 10) {
        CS(0, 8); CS(1, 9); CS(2, 7); CS(3, 5); CS(4, 6); CS(0, 2); CS(1, 4); CS(5, 8); CS(7, 9);
        CS(0, 3); CS(2, 4); CS(5, 7); CS(6, 9); CS(0, 1); CS(3, 6); CS(8, 9); CS(1, 5); CS(2, 3);
        CS(4, 8); CS(6, 7); CS(1, 2); CS(3, 5); CS(4, 6); CS(7, 8); CS(2, 3); CS(4, 5); CS(6, 7);
        CS(3, 4); CS(5, 6);
    }
#undef CS
}

// ----------------------------------------------------------------------------------------------
// Packed-state addressing (include/nimmt_b200.h).  Games are stored in tiles of 32, one contiguous record
// per tile: an immutable block (the dealt cards) followed by a mutable block (slot bits + scores, row records), so that
// the step kernel loads a tile with ONE bulk copy and stores its mutable tail with one:
//     [tile]{ uint2 cards[P][32];                                  256 P bytes, written by deal / reset_to only
//             uint32 meta[P][32]; uint8 rows[32][24] }             128 P + 768 bytes
// A warp of 32 consecutive games reads contiguous 256- and 128-byte runs per player.
// ----------------------------------------------------------------------------------------------
constexpr int kTileGames = 32;

struct StateView {
    uint8_t* base;
    int64_t B;
    int P;
    __host__ __device__ StateView(void* base_, int64_t num_games, int num_players)
        : base(reinterpret_cast<uint8_t*>(base_)), B(num_games), P(num_players) {}
    static __host__ __device__ constexpr int64_t cards_tile_bytes(int P) { return (int64_t)P * kTileGames * 8; }
    static __host__ __device__ constexpr int64_t mut_tile_bytes(int P) { return (int64_t)P * kTileGames * 4 + kTileGames * 24; }
    static __host__ __device__ constexpr int64_t tile_bytes(int P) { return cards_tile_bytes(P) + mut_tile_bytes(P); }
    static __host__ __device__ constexpr int64_t bytes(int64_t B, int P) { return (B + kTileGames - 1) / kTileGames * tile_bytes(P); }
    __host__ __device__ uint8_t* tile_ptr(int64_t tile) const { return base + tile * tile_bytes(P); }
    __host__ __device__ uint8_t* mut_ptr(int64_t tile) const { return tile_ptr(tile) + cards_tile_bytes(P); }
    __host__ __device__ uint2* cards_ptr(int64_t g, int p) const {
        return reinterpret_cast<uint2*>(tile_ptr(g >> 5)) + p * kTileGames + (g & 31);
    }
    __host__ __device__ uint32_t* meta_ptr(int64_t g, int p) const {
        return reinterpret_cast<uint32_t*>(mut_ptr(g >> 5)) + p * kTileGames + (g & 31);
    }
    __host__ __device__ uint2* rows_ptr(int64_t g) const {   // three uint2 = the 24-byte row record
        return reinterpret_cast<uint2*>(mut_ptr(g >> 5) + (int64_t)P * kTileGames * 4) + 3 * (g & 31);
    }
};

NIMMT_HD void load_rows(const StateView& s, int64_t g, Board& b) {
    const uint2* r = s.rows_ptr(g);
    const uint2 a = r[0], c = r[1], d = r[2];
    b.unpack((uint64_t)a.x | ((uint64_t)a.y << 32), (uint64_t)c.x | ((uint64_t)c.y << 32), (uint64_t)d.x | ((uint64_t)d.y << 32));
}

NIMMT_HD void store_rows(const StateView& s, int64_t g, const Board& b) {
    uint64_t q0, q1, q2;
    b.pack(q0, q1, q2);
    uint2* r = s.rows_ptr(g);
    r[0] = make_uint2((uint32_t)q0, (uint32_t)(q0 >> 32));
    r[1] = make_uint2((uint32_t)q1, (uint32_t)(q1 >> 32));
    r[2] = make_uint2((uint32_t)q2, (uint32_t)(q2 >> 32));
}

}  // namespace nimmt
