// step_tile.cuh — env.step on one game IN PLACE in its tile record (the form k_step_tiles runs, one game per lane, the
// record in shared memory).  __host__ __device__: tests/host_sim runs the very same code over a tile in host memory and
// diffs it against the oracle (test harness only; the library exports no host implementation).
#pragma once
#include "step.cuh"

namespace nimmt {

template <int P>
struct TileLayout {
    static constexpr int kCardsBytes = P * kTileGames * 8;                    // uint2 [P][32]
    static constexpr int kMeta = kCardsBytes;                                 // uint32 [P][32]
    static constexpr int kRows = kMeta + P * kTileGames * 4;                  // 24-byte records
    static constexpr int kMutBytes = P * kTileGames * 4 + kTileGames * 24;
    static constexpr int kTileBytes = kCardsBytes + kMutBytes;
    static constexpr int kActBytes = kTileGames * P;
    static_assert(kTileBytes % 16 == 0 && kActBytes % 16 == 0 && kCardsBytes % 16 == 0, "bulk copies need 16-byte alignment");
};

// One game, in place in its tile's shared-memory record; `lane` selects the game.
//   tile     the tile record (cards | meta | rows) in shared memory
//   acts     the tile's action bytes in shared memory (kRandom: unused)
//   keys_w/u this lane's row keys (game.cuh::place_v3), 16-byte aligned
//   rew_out / done_out / illegal_out / act_out   where this game's P reward bytes, its flags and (kRandom) its drawn cards go
//            (global memory on the device: a warp's lanes write neighbouring addresses)
//   kChoice  free-row-choice mode: `rows` holds, like `acts`, one byte per (game, player) — the row that player takes if their
//            card undercuts every row (0..3; anything else rejects the step like an illegal card)
//   kPacked  the compact transfer format for hosts on the far side of PCIe (nimmt_step_packed): `acts` holds one 4-bit HAND SLOT
//            per player (player p = nibble p & 1 of byte p >> 1 of the game's ceil(P / 2) bytes; slot s = the s-th card of the
//            hand AS DEALT, ascending) instead of card bytes, and the results leave as one bit-packed record per game through
//            `rew_out`: bits [5 p, 5 p + 5) the bull heads player p took, bit 5 P done, bit 5 P + 1 illegal
template <int P>
constexpr int packed_action_bytes() { return (P + 1) / 2; }
template <int P>
constexpr int packed_result_bytes() { return (5 * P + 2 + 7) / 8; }

template <int P, bool kRandom, bool kChoice = false, bool kPacked = false>
NIMMT_HD void step_lane(uint8_t* tile, const uint8_t* acts, int lane, const uint8_t* values5, uint32_t* keys_w, uint32_t* keys_u,
                                          uint8_t* rew_out, uint8_t* done_out, uint8_t* illegal_out, uint8_t* act_out, uint64_t seed,
                                          uint64_t game_id, uint32_t turn, const uint8_t* rows = nullptr, const uint32_t* sel8 = nullptr) {
    using L = TileLayout<P>;
    const uint2* cards0 = reinterpret_cast<const uint2*>(tile) + lane;              // + p * kTileGames
    uint32_t* meta0 = reinterpret_cast<uint32_t*>(tile + L::kMeta) + lane;           // + p * kTileGames
    uint8_t* rec = tile + L::kRows + lane * 24;

    uint32_t act[P], meta[P];
    bool legal = true;
    if constexpr (kRandom) {
        // DrunkHamster for every seat (agents/random.py:8-10), drawn exactly as k_random_actions draws (step.cuh::random_actions_game):
        // the chosen slot is known, no search is needed
        Philox rng(seed, game_id, /*stream=*/0x61637400u + turn, 0);
        uint4 r = make_uint4(0, 0, 0, 0);
#pragma unroll
        for (int p = 0; p < P; ++p) {
            if ((p & 3) == 0) r = rng.next<7>();
            const uint32_t word = (p & 3) == 0 ? r.x : (p & 3) == 1 ? r.y : (p & 3) == 2 ? r.z : r.w;
            HandRec h;
            h.lo = cards0[p * kTileGames];
            h.meta = meta0[p * kTileGames];
            const uint32_t n = (uint32_t)rec_count(h);
            const uint32_t slot = rec_select_slot(sel8, h.meta, below(word, n ? n : 1u));   // the k-th unplayed slot: one table look-up
            act[p] = n ? rec_card_dyn(h, slot) : 255u;
            meta[p] = h.meta | (n ? rec_slot_mask(slot) : 0u);
            legal = legal && n != 0u;                      // an empty hand "plays" 255: rejected like any illegal card
        }
        if (act_out) {
            int a[P];
#pragma unroll
            for (int p = 0; p < P; ++p) a[p] = (int)act[p];
            store_bytes<P>(act_out, 0, a);
        }
    } else if constexpr (kPacked) {
        // the slot is given: the card is read, not searched; legal iff the slot exists and still holds its card
#pragma unroll
        for (int p = 0; p < P; ++p) {
            const uint32_t slot = ((uint32_t)acts[lane * packed_action_bytes<P>() + (p >> 1)] >> (4 * (p & 1))) & 15u;
            HandRec h;
            h.lo = cards0[p * kTileGames];
            h.meta = meta0[p * kTileGames];
            const uint32_t bit = slot < (uint32_t)kHand ? rec_slot_mask(slot) & ~h.meta : 0u;
            act[p] = rec_card_dyn(h, slot < (uint32_t)kHand ? slot : 0u);
            meta[p] = h.meta | bit;
            legal = legal && bit != 0u;
        }
    } else {
        // env.py:68-69 — every card is checked before anything is touched
        uint32_t any = 0;
#pragma unroll
        for (int p = 0; p < P; ++p) {
            act[p] = acts[lane * P + p];
            any |= act[p];
            HandRec h;
            h.lo = cards0[p * kTileGames];
            h.meta = meta0[p * kTileGames];
            const uint32_t bit = rec_slot_bit(h, act[p]) & ~h.meta;   // the slot must still hold its card
            meta[p] = h.meta | bit;
            legal = legal && bit != 0u;
        }
        // rec_slot_bit is exact for ids < 128 (104..127 match no stored card; 127 only never-dealt slots, which are marked empty)
        legal = legal && any < 128u;
        if constexpr (kChoice) {
            uint32_t all = 0;
#pragma unroll
            for (int p = 0; p < P; ++p) all |= rows[lane * P + p];
            legal = legal && all < (uint32_t)kRows;
        }
    }

    uint32_t gain[P];                                    // score words before - after: the top byte is -(bull heads taken) in two's complement
#pragma unroll
    for (int p = 0; p < P; ++p) gain[p] = 0u;
    bool done = (meta0[0] & kEmptyBits) == kEmptyBits;   // an illegal step leaves the game as it was
    if (legal) {
        // the row keys from the record: top card = byte 4 (len - 1) + r, meta byte r = len | sum << 3  =>  W = top << 10 | meta << 2 | r
        {
            const uint32_t metas = *reinterpret_cast<const uint32_t*>(rec + 20);
            uint32_t w[kRows], u[kRows];
#pragma unroll
            for (int r = 0; r < kRows; ++r) {
                const uint32_t m = byte_perm(metas, 0u, 0x4440u + (uint32_t)r);
                NIMMT_CHECK((m & 7u) >= 1u && (m & 7u) <= 5u);   // 1 <= len <= 5 between placements (env.py:133,170)
                const uint32_t top = rec[4u * (m & 7u) + (uint32_t)r - 4u];
                w[r] = top * 1024u + (m * 4u + (uint32_t)r);
                u[r] = key_u_from_w(w[r]);
            }
            *reinterpret_cast<uint4*>(keys_w) = make_uint4(w[0], w[1], w[2], w[3]);
            *reinterpret_cast<uint4*>(keys_u) = make_uint4(u[0], u[1], u[2], u[3]);
        }
        // env.py:131: the played slots are empty from here on; takes are added to the score fields in place below
#pragma unroll
        for (int p = 0; p < P; ++p) meta0[p * kTileGames] = meta[p];

        int keys[P];
#pragma unroll
        for (int p = 0; p < P; ++p) keys[p] = (int)(act[p] * 1024u + (uint32_t)(p << 6));
        sort_keys<P>(keys);   // env.py:124-125

#pragma unroll
        for (int i = 0; i < P; ++i) {
            const uint32_t key = (uint32_t)keys[i];
            uint32_t row, keep4;
            uint32_t choice = 0;
            if constexpr (kChoice) choice = rows[lane * P + ((key >> 6) & 15u)];          // the row this card's player named
            const uint32_t pen5 = place_v3<1, kChoice>(keys_w, keys_u, key, values5, row, keep4, choice);   // env.py:126-134
            NIMMT_CHECK(keep4 + row < 20u && ((key >> 6) & 15u) < (uint32_t)P);
            rec[keep4 + row] = (uint8_t)(key >> 10);                                      // the one byte of the record a placement changes
            // env.py:167-169: the player of this card takes pen bull heads — added to its score field where the word lives,
            // in shared memory (the player is data: a register array would need a select per player)
            uint32_t* mp = reinterpret_cast<uint32_t*>(reinterpret_cast<uint8_t*>(meta0) + 2u * (key & 0x3C0u));
            *mp += pen5 << (kRecScoreShift - 5);
        }
        // rewards = what the score fields gained; the record's meta bytes from the final keys
#pragma unroll
        for (int p = 0; p < P; ++p) gain[p] = meta[p] - meta0[p * kTileGames];
        const uint4 fw = *reinterpret_cast<const uint4*>(keys_w);
        *reinterpret_cast<uint32_t*>(rec + 20) = byte_perm(byte_perm(fw.x >> 2, fw.y >> 2, 0x0040u), byte_perm(fw.z >> 2, fw.w >> 2, 0x0040u), 0x5410u);
        done = (meta[0] & kEmptyBits) == kEmptyBits;   // env.py:246-249
    }
    if constexpr (kPacked) {
        // one bit-packed record per game: 5 bits of bull heads per player (<= 27), then done and illegal
        uint64_t rec64 = 0;
#pragma unroll
        for (int p = 0; p < P; ++p) rec64 |= (uint64_t)((0u - (gain[p] >> kRecScoreShift)) & 31u) << (5 * p);   // gain's top byte is -(bull heads)
        rec64 |= (uint64_t)(done ? 1u : 0u) << (5 * P) | (uint64_t)(legal ? 0u : 1u) << (5 * P + 1);
#pragma unroll
        for (int b = 0; b < packed_result_bytes<P>(); ++b) rew_out[b] = (uint8_t)(rec64 >> (8 * b));
        return;
    }
    // rewards (env.py:169): byte p = top byte of gain[p]; gathered with byte permutes, stored with the widest access P allows
    if constexpr (P % 4 == 0) {
#pragma unroll
        for (int i = 0; i < P / 4; ++i)
            reinterpret_cast<uint32_t*>(rew_out)[i] =
                byte_perm(byte_perm(gain[4 * i], gain[4 * i + 1], 0x0073u), byte_perm(gain[4 * i + 2], gain[4 * i + 3], 0x0073u), 0x5410u);
    } else if constexpr (P % 2 == 0) {
#pragma unroll
        for (int i = 0; i < P / 2; ++i) reinterpret_cast<uint16_t*>(rew_out)[i] = (uint16_t)byte_perm(gain[2 * i], gain[2 * i + 1], 0x0073u);
    } else {
#pragma unroll
        for (int p = 0; p < P; ++p) rew_out[p] = (uint8_t)(gain[p] >> kRecScoreShift);
    }
    done_out[0] = done;
    if (illegal_out) illegal_out[0] = !legal;
}

}  // namespace nimmt
