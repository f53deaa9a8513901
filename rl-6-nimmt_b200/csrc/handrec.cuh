// handrec.cuh — a player's hand as it is STORED, and the conversions to and from the 104-bit
// card sets the per-game logic of step.cuh works on.
//
// A hand only ever shrinks, and by exactly one card per step.  So the ten cards a player was dealt are written
// once (by the deal) and never again; a step reads them to find the slot of the played card and writes back one
// 32-bit word per player — ten "slot is empty" bits plus the running Hornochsen score — instead of a 128-bit set:
//
//     uint2    lo      bytes 0..7: the cards of slots 0..7, ascending; 0xFF = no card was dealt to the slot
//     uint32_t meta    bits  0..9   slot i is empty (played, or never dealt)
//                      bits 10..17  Hornochsen taken so far (<= 171)
//                      bits 18..24  card of slot 8, bits 25..31 card of slot 9 (127 = none) — immutable, they
//                                   ride in the mutable word because 10 cards are 2 bytes more than a uint2
// Slots are in ascending card order, so "the k-th unplayed slot" is "the k-th smallest card in hand", which is
// what the observation's hand block (env.py:209-210) and DrunkHamster's uniform choice need.
#pragma once
#include "game.cuh"

namespace nimmt {

struct HandRec {
    uint2 lo;
    uint32_t meta;
};

constexpr uint32_t kSlotBits = 0x3FFu;
constexpr int kRecScoreShift = 10;
constexpr uint32_t kNoCard7 = 127u;

NIMMT_HD int ffs32(uint32_t x) {   // 1-based index of the lowest set bit, 0 if none
#ifdef __CUDA_ARCH__
    return __ffs((int)x);
#else
    return __builtin_ffs((int)x);
#endif
}

NIMMT_HD uint32_t rec_score(const HandRec& h) { return (h.meta >> kRecScoreShift) & 0xFFu; }
NIMMT_HD int rec_count(const HandRec& h) { return kHand - popc32(h.meta & kSlotBits); }
NIMMT_HD bool rec_empty(const HandRec& h) { return (h.meta & kSlotBits) == kSlotBits; }

// Card in `slot` (0..9), whether or not it has been played; >= 104 if the slot never held a card.
NIMMT_HD uint32_t rec_card(const HandRec& h, int slot) {
    const uint32_t word = slot < 4 ? h.lo.x : h.lo.y;
    const uint32_t low = (word >> (8 * (slot & 3))) & 0xFFu;
    const uint32_t high = (h.meta >> (slot == 8 ? 18 : 25)) & 127u;
    return slot < 8 ? low : high;
}

// Flags (bit 7 of each byte) the bytes of `w` that equal the byte replicated in `pattern`.  Bits above the
// lowest flag may be spurious (borrow), so callers use the lowest flag only; cards in a hand are distinct.
NIMMT_HD uint32_t eq_byte_flags(uint32_t w, uint32_t pattern) {
    const uint32_t x = w ^ pattern;
    return (x - 0x01010101u) & ~x & 0x80808080u;
}

// Slot holding `card`, or -1 (card ids >= 104 never match: stored "none" bytes are 0xFF / 127).
NIMMT_HD int rec_find(const HandRec& h, uint32_t card) {
    const uint32_t pattern = (card & 0xFFu) * 0x01010101u;
    const uint32_t f0 = eq_byte_flags(h.lo.x, pattern), f1 = eq_byte_flags(h.lo.y, pattern);
    int slot = -1;
    slot = ((h.meta >> 25) & 127u) == card ? 9 : slot;
    slot = ((h.meta >> 18) & 127u) == card ? 8 : slot;
    slot = f1 ? 4 + ((ffs32(f1) - 1) >> 3) : slot;
    slot = f0 ? (ffs32(f0) - 1) >> 3 : slot;
    return card < (uint32_t)kCards ? slot : -1;
}

// env.py:114-118 + :131 — if `card` is in the hand, marks its slot empty in `meta` and returns true.
NIMMT_HD bool rec_take(const HandRec& h, uint32_t card, uint32_t& meta) {
    const int slot = rec_find(h, card);
    const uint32_t bit = slot >= 0 ? 1u << slot : 0u;
    const bool held = bit != 0u && (h.meta & bit) == 0u;
    meta = h.meta | (held ? bit : 0u);
    return held;
}

// Card of the k-th (0-based) unplayed slot = the k-th smallest card in hand; k < rec_count(h).
NIMMT_HD uint32_t rec_select(const HandRec& h, uint32_t k) { return rec_card(h, (int)select_bit32(~h.meta & kSlotBits, k)); }

// The hand as a 104-bit set with the score in the top byte (the form step.cuh's Game<P> uses).
NIMMT_HD uint4 rec_to_mask(const HandRec& h) {
    uint4 m = make_uint4(0, 0, 0, 0);
#pragma unroll
    for (int i = 0; i < kHand; ++i)
        if (!((h.meta >> i) & 1u)) mask_set(m, rec_card(h, i));
    m.w = (m.w & kHighCardMask) | (rec_score(h) << kScoreShift);
    return m;
}

// A fresh record from up to ten cards in ascending order (`cards[i]` for i < n).
NIMMT_HD HandRec rec_from_sorted(const uint32_t (&cards)[kHand], int n, uint32_t score) {
    uint32_t c[kHand];
#pragma unroll
    for (int i = 0; i < kHand; ++i) c[i] = i < n ? cards[i] : 0xFFu;
    HandRec h;
    h.lo.x = c[0] | (c[1] << 8) | (c[2] << 16) | (c[3] << 24);
    h.lo.y = c[4] | (c[5] << 8) | (c[6] << 16) | (c[7] << 24);
    const uint32_t empty = (kSlotBits << n) & kSlotBits;
    h.meta = empty | (score << kRecScoreShift) | ((c[8] & 127u) << 18) | ((c[9] & 127u) << 25);
    return h;
}

// A fresh record from a card set in Game<P> form (at most ten cards; extra cards are dropped).
NIMMT_HD HandRec rec_from_mask(const uint4& m) {
    uint32_t cards[kHand];
    int n = 0;
    const uint32_t words[4] = {m.x, m.y, m.z, m.w & kHighCardMask};
#pragma unroll
    for (int wi = 0; wi < 4; ++wi) {
        uint32_t w = words[wi];
        while (w) {
            const uint32_t card = (uint32_t)(wi * 32 + ffs32(w) - 1);
            w &= w - 1;
#pragma unroll
            for (int i = 0; i < kHand; ++i)   // static indices: cards[] stays in registers
                if (i == n) cards[i] = card;
            n += n < kHand;
        }
    }
    return rec_from_sorted(cards, n, m.w >> kScoreShift);
}

}  // namespace nimmt
