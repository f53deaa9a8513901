// handrec.cuh — a player's hand as it is STORED, and the conversions to and from the 104-bit
// card sets the per-game logic of step.cuh works on.
//
// A hand only ever shrinks, and by exactly one card per step.  So the ten cards a player was dealt are written
// once (by the deal) and never again; a step reads them to find the slot of the played card and writes back one
// 32-bit word per player — ten "slot is empty" bits plus the running Hornochsen score — instead of a 128-bit set:
//
//     uint2    lo      bytes 0..7: the cards of slots 0..7, ascending; 0x7F = no card was dealt to the slot
//     uint32_t meta    bits  0..6   card of slot 8 (127 = none)      bit  7  slot 8 is empty (played, or never dealt)
//                      bits  8..14  card of slot 9 (127 = none)      bit 15  slot 9 is empty
//                      bits 16..23  slots 0..7 are empty
//                      bits 24..31  Hornochsen taken so far (<= 171)
//                      (the two card fields are immutable; they ride in the mutable word because 10 cards are 2 bytes
//                      more than a uint2)
// The layout is chosen for the slot search of the step kernel: comparing bytes 0 and 1 of the word with the played card
// leaves the "found" flags of slots 8 and 9 on bits 7 and 15 — exactly where those slots' empty bits live — and the flags
// of slots 0..7 are one multiply-add away from bits 16..23.
// Slots are in ascending card order, so "the k-th unplayed slot" is "the k-th smallest card in hand", which is
// what the observation's hand block (env.py:209-210) and DrunkHamster's uniform choice need.
#pragma once
#include "game.cuh"

namespace nimmt {

struct HandRec {
    uint2 lo;
    uint32_t meta;
};

constexpr uint32_t kEmptyBits = 0x00FF8080u;   // the ten slot-empty bits, in the word's own positions
constexpr uint32_t kSlotBits = 0x3FFu;         // the same ten bits in slot order (rec_empties)
constexpr int kRecScoreShift = 24;
constexpr uint32_t kNoCard7 = 127u;

NIMMT_HD int ffs32(uint32_t x) {   // 1-based index of the lowest set bit, 0 if none
#ifdef __CUDA_ARCH__
    return __ffs((int)x);
#else
    return __builtin_ffs((int)x);
#endif
}

// Slot-order view of the empty bits (bit i = slot i is empty) and its inverse for one slot.
NIMMT_HD uint32_t rec_empties(uint32_t meta) { return ((meta >> 16) & 0xFFu) | ((meta & 0x80u) << 1) | ((meta & 0x8000u) >> 6); }
NIMMT_HD uint32_t rec_slot_mask(uint32_t slot) { return slot < 8u ? 0x10000u << slot : (slot == 8u ? 0x80u : 0x8000u); }
// Ten slot-order bits -> the word's positions.
NIMMT_HD uint32_t rec_spread(uint32_t bits10) { return ((bits10 & 0xFFu) << 16) | ((bits10 & 0x100u) >> 1) | ((bits10 & 0x200u) << 6); }

NIMMT_HD uint32_t rec_score(const HandRec& h) { return h.meta >> kRecScoreShift; }
NIMMT_HD int rec_count(const HandRec& h) { return kHand - popc32(h.meta & kEmptyBits); }
NIMMT_HD bool rec_empty(const HandRec& h) { return (h.meta & kEmptyBits) == kEmptyBits; }
NIMMT_HD bool rec_slot_empty(const HandRec& h, int slot) { return (h.meta & rec_slot_mask((uint32_t)slot)) != 0u; }

// Card in `slot` (0..9), whether or not it has been played; >= 104 if the slot never held a card.
NIMMT_HD uint32_t rec_card(const HandRec& h, int slot) {
    const uint32_t word = slot < 4 ? h.lo.x : h.lo.y;
    const uint32_t low = (word >> (8 * (slot & 3))) & 0xFFu;
    const uint32_t high = (h.meta >> (slot == 8 ? 0 : 8)) & 127u;
    return slot < 8 ? low : high;
}

// k-th (0-based) UNPLAYED slot of a hand, in slot order (= the k-th smallest card in hand), k < rec_count: one table look-up for
// slots 0..7 (sel8[m] lists the positions of m's set bits, 3 bits each) and two bit tests for slots 8 and 9 — a fifth of the
// instructions of a popcount bisection.  `sel8` = the 256-word table (shared memory on the device: every lane indexes its own entry).
// (global memory, not __constant__: the staging loop reads 32 different words per warp, which the constant cache would serialise)
__device__ const uint32_t d_select8[256] = {
#include "select8_table.inc"
};
static const uint32_t h_select8[256] = {
#include "select8_table.inc"
};
__device__ __forceinline__ void stage_select8(uint32_t* smem) {
    for (uint32_t i = threadIdx.x; i < 256u; i += blockDim.x) smem[i] = d_select8[i];
}
NIMMT_HD uint32_t rec_select_slot(const uint32_t* sel8, uint32_t meta, uint32_t k) {
    const uint32_t avail8 = (~meta >> 16) & 0xFFu, n8 = (uint32_t)popc32(avail8);
    const uint32_t low = (sel8[avail8] >> (3u * k)) & 7u;            // garbage when k >= n8: not selected below
    const uint32_t high = (k == n8 && !(meta & 0x80u)) ? 8u : 9u;    // slot 8 if it is the next unplayed one, else slot 9
    return k < n8 ? low : high;
}

// The same for a slot known only at run time, by byte permutes instead of variable shifts (slot < 10).
NIMMT_HD uint32_t rec_card_dyn(const HandRec& h, uint32_t slot) {
    const uint32_t low = byte_perm(h.lo.x, h.lo.y, slot & 7u), high = byte_perm(h.meta, 0u, slot & 1u);
    return (slot < 8u ? low : high) & 0x7Fu;   // stored bytes are < 0x80 (0x7F = none); bit 7 of the meta bytes is an empty flag
}

// The slot that was dealt `card`, as that slot's EMPTY BIT of the meta word (0 if the hand never held the card).  Exact for
// card < 128 (ids 104..126 match nothing; 127 matches only never-dealt slots, whose empty bits are set from the start);
// larger ids may return garbage and must be rejected by the caller.  Branch-free and mostly multiplies, which run on the FMA
// pipe while the rest of a step saturates the ALU pipe:
//   * all stored bytes and the pattern bytes are < 0x80, so (x + 0x7F) sets bit 7 of a byte iff the byte is non-zero and
//     never carries into its neighbour: ~(x + 0x7F7F7F7F) & 0x80808080 flags exactly the equal bytes;
//   * the high word of flags * (2^25 + 2^18 + 2^11 + 2^4) has the flag of byte i at bit i (stray partial products land
//     at bits >= 8; at most one flag is set because the cards of a hand are distinct);
//   * the same byte test on bytes 0 and 1 of the meta word (the cards of slots 8 and 9; their bit 7 / 15 masked off) leaves
//     its flags on bits 7 and 15.
NIMMT_HD uint32_t rec_slot_bit(const HandRec& h, uint32_t card) {
    const uint32_t pattern = card * 0x01010101u;
    const uint32_t f0 = ~((h.lo.x ^ pattern) + 0x7F7F7F7Fu) & 0x80808080u;
    const uint32_t f1 = ~((h.lo.y ^ pattern) + 0x7F7F7F7Fu) & 0x80808080u;
    const uint32_t low = (umulhi32(f0, 0x02040810u) | umulhi32(f1, 0x20408100u)) & 0xFFu;
    const uint32_t f89 = ~(((h.meta ^ pattern) & 0x7F7Fu) + 0x7F7Fu) & 0x8080u;
    return low * 0x10000u + f89;
}

// Slot holding `card`, or -1.
NIMMT_HD int rec_find(const HandRec& h, uint32_t card) {
    return card < (uint32_t)kCards ? ffs32(rec_empties(rec_slot_bit(h, card))) - 1 : -1;
}

// env.py:114-118 + :131 — if `card` is in the hand, marks its slot empty in `meta` and returns true.
NIMMT_HD bool rec_take(const HandRec& h, uint32_t card, uint32_t& meta) {
    const uint32_t bit = rec_slot_bit(h, card) & ~h.meta;   // the slot must still hold its card
    meta = h.meta | bit;                                    // only committed by the caller if every move is legal
    return bit != 0u && card < (uint32_t)kCards;
}

// Card of the k-th (0-based) unplayed slot = the k-th smallest card in hand; k < rec_count(h).
NIMMT_HD uint32_t rec_select(const HandRec& h, uint32_t k) { return rec_card(h, (int)select_bit32(~rec_empties(h.meta) & kSlotBits, k)); }

// The hand as a 104-bit set with the score in the top byte (the form step.cuh's Game<P> uses).
NIMMT_HD uint4 rec_to_mask(const HandRec& h) {
    uint4 m = make_uint4(0, 0, 0, 0);
#pragma unroll
    for (int i = 0; i < kHand; ++i)
        if (!rec_slot_empty(h, i)) mask_set(m, rec_card(h, i));
    m.w = (m.w & kHighCardMask) | (rec_score(h) << kScoreShift);
    return m;
}

// A fresh record from up to ten cards in ascending order (`cards[i]` for i < n).
NIMMT_HD HandRec rec_from_sorted(const uint32_t (&cards)[kHand], int n, uint32_t score) {
    uint32_t c[kHand];
#pragma unroll
    for (int i = 0; i < kHand; ++i) c[i] = i < n ? cards[i] : kNoCard7;
    HandRec h;
    h.lo.x = c[0] | (c[1] << 8) | (c[2] << 16) | (c[3] << 24);
    h.lo.y = c[4] | (c[5] << 8) | (c[6] << 16) | (c[7] << 24);
    const uint32_t empty = (kSlotBits << n) & kSlotBits;
    h.meta = rec_spread(empty) | (score << kRecScoreShift) | (c[8] & 127u) | ((c[9] & 127u) << 8);
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
