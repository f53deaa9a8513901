// env_step.cu — batched env.step, random actions, and the two fused (SechsNimmtEnv.step, env.py:64-77;
// DrunkHamster.forward, agents/random.py:8-10).  One thread per game; HBM-bandwidth bound.
#include "abi_common.cuh"
#include "tma.cuh"

namespace nimmt {

// ------------------------------------------------------------------------------------------
// k_step — SechsNimmtEnv.step (env.py:64-77) without the observation rebuild.
// Reads (12 P + 24) + P bytes and writes (4 P + 24) + P + 1 (+1) bytes per game: the dealt cards are immutable.
// kRandom: the actions are drawn in-kernel (DrunkHamster, agents/random.py:8-10) instead of read.
// ------------------------------------------------------------------------------------------
template <int P, bool kRandom>
__global__ void __launch_bounds__(kStepThreads)
k_step(StateView s, const uint8_t* __restrict__ actions_in, uint8_t* __restrict__ actions_out, int8_t* __restrict__ rewards,
       uint8_t* __restrict__ done, uint8_t* __restrict__ illegal, uint64_t seed, uint32_t turn, uint64_t game0,
       int64_t first_game) {
    __shared__ uint8_t values[128];
    const int64_t g_raw = first_game + (int64_t)blockIdx.x * kStepThreads + threadIdx.x;
    const bool valid = g_raw < s.B;
    const int64_t g = valid ? g_raw : s.B - 1;  // tail threads shadow the last game and write nothing

    // Issue every global load before the value table is staged, so that the block barrier below
    // overlaps the memory latency instead of preceding it.
    RawGame<P> raw;
    int act[P];
    if constexpr (!kRandom) load_bytes<P>(actions_in, g, act);
    load_raw<P>(s, g, raw);
    stage_card_values(values);
    __syncthreads();
    if (!valid) return;
    GameRec<P> gm;
    unpack_raw<P>(raw, gm);
    if constexpr (kRandom) {
        random_actions_game<P>(gm, seed, game0 + (uint64_t)g, turn, act);
        if (actions_out) store_bytes<P>(actions_out, g, act);
    }

    int penalty[P];
    const bool legal = step_game<P>(gm, act, values, penalty);
    if (legal) store_step<P>(s, g, gm);

    int rew[P];
#pragma unroll
    for (int p = 0; p < P; ++p) rew[p] = -penalty[p];
    store_bytes<P>(reinterpret_cast<uint8_t*>(rewards), g, rew);
    done[g] = game_done<P>(gm);
    if (illegal) illegal[g] = !legal;
}

// ------------------------------------------------------------------------------------------
// k_step_smem — the throughput path of env.step.  HBM <-> shared memory is done entirely by the TMA
// engine (1-D cp.async.bulk, both directions); the threads only store their game's few output bytes.
//
// Each WARP runs its own double-buffered pipeline over tiles of 32 games, with no block-level
// synchronisation at all:
//     lane 0:  bulk-load tile i+1 — the tile's immutable block (dealt cards, 256 P bytes), its mutable
//              block (slot bits + scores + row records, 128 P + 768 bytes) and 32 P action bytes, three
//              contiguous runs of HBM — into the other buffer                      [mbarrier]
//     lanes:   step tile i IN PLACE in shared memory — find the played card's slot among the ten dealt
//              cards, set one bit of the player's meta word, write one card byte of the row record,
//              add a take to the score field; the only per-row state kept in registers is the two
//              comparison keys of game.cuh::RowKeys
//     lane 0:  bulk-store the mutable block                                        [bulk group]
//     lanes:   store their game's reward bytes and done / illegal flags (coalesced)
// The dealt cards never travel back: a step writes 4 bytes per player instead of the 16 of a card set.
// ------------------------------------------------------------------------------------------
constexpr int kSmemWarps = 4;   // warps per block; each is independent

template <int P>
struct TileLayout {
    static constexpr int kCardsBytes = P * kTileGames * 8;                    // uint2 [P][32]
    static constexpr int kMeta = kCardsBytes;                                 // uint32 [P][32]
    static constexpr int kRows = kMeta + P * kTileGames * 4;                  // 24-byte records
    static constexpr int kMutBytes = P * kTileGames * 4 + kTileGames * 24;
    static constexpr int kActions = kRows + kTileGames * 24;
    static constexpr int kLoadBytes = kActions + kTileGames * P;              // everything above is loaded
    static constexpr int kRewards = kLoadBytes;
    static constexpr int kDone = kRewards + kTileGames * P;
    static constexpr int kIllegal = kDone + kTileGames;
    static constexpr int kBytes = kIllegal + kTileGames;
    static constexpr int kStride = (kBytes + 127) / 128 * 128;
    static_assert(kActions % 16 == 0 && kRewards % 16 == 0 && kDone % 16 == 0 && kIllegal % 16 == 0, "bulk copies need 16-byte alignment");
};

// kRandom: the actions are drawn in the kernel, there is no action tape to load.
template <int P, bool kRandom>
__device__ __forceinline__ void issue_tile_loads(const StateView& s, const uint8_t* actions, int64_t tile, uint8_t* buf, uint64_t* bar) {
    using L = TileLayout<P>;
    mbar_arrive_expect_tx(bar, kRandom ? L::kActions : L::kLoadBytes);
    bulk_load(buf, s.tile_ptr(tile), L::kCardsBytes + L::kMutBytes, bar);   // the whole tile record: cards, meta words, row records
    if constexpr (!kRandom) bulk_load(buf + L::kActions, actions + tile * (kTileGames * P), kTileGames * P, bar);
}

template <int P>
__device__ __forceinline__ void issue_tile_stores(const StateView& s, int8_t* rewards, uint8_t* done, uint8_t* illegal, uint8_t* actions_out,
                                                  int64_t tile, const uint8_t* buf) {
    using L = TileLayout<P>;
    const int64_t g0 = tile * kTileGames;
    bulk_store(s.mut_ptr(tile), buf + L::kMeta, L::kMutBytes);
    if (actions_out) bulk_store(actions_out + g0 * P, buf + L::kActions, kTileGames * P);
    bulk_commit();
}

// One game, in place in the tile buffer.  `lane` selects the game.
// kRandom: every player plays a uniformly random card of its hand (DrunkHamster, agents/random.py:8-10), drawn exactly as
// k_random_actions draws it (step.cuh::random_actions_game: same Philox stream, same word per player), so the fused step
// equals k_random_actions followed by a step; the chosen slot is known, no search is needed.
template <int P, bool kRandom>
__device__ __forceinline__ void step_in_smem(uint8_t* buf, int lane, const uint8_t* values, int* keys_w, int* keys_u, uint8_t* rewards_out,
                                             uint8_t* done_out, uint8_t* illegal_out, uint64_t seed, uint64_t game_id, uint32_t turn) {
    using L = TileLayout<P>;
    const uint2* cards0 = reinterpret_cast<const uint2*>(buf) + lane;             // + p * kTileGames
    uint32_t* meta0 = reinterpret_cast<uint32_t*>(buf + L::kMeta) + lane;         // + p * kTileGames
    uint8_t* rec = buf + L::kRows + lane * 24;

    int act[P];
    uint32_t meta[P];
    bool legal = true;
    if constexpr (kRandom) {
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
            const uint32_t slot = select_bit32(~h.meta & kSlotBits, below(word, n ? n : 1u));
            act[p] = n ? (int)rec_card(h, (int)slot) : 255;
            meta[p] = h.meta | (n ? 1u << slot : 0u);
            legal = legal && n != 0u;                      // an empty hand "plays" 255: rejected like any illegal card
        }
        store_bytes<P>(buf + L::kActions, lane, act);
    } else {
        load_bytes<P>(buf + L::kActions, lane, act);
        // env.py:68-69 — check every card before touching anything
#pragma unroll
        for (int p = 0; p < P; ++p) {
            HandRec h;
            h.lo = cards0[p * kTileGames];
            h.meta = meta0[p * kTileGames];
            legal = rec_take(h, (uint32_t)act[p], meta[p]) && legal;
        }
    }

    // rewards default to 0 (env.py:122); a take overwrites its player's byte below
    uint8_t* rw = buf + L::kRewards + lane * P;
    if constexpr (P % 4 == 0) {
#pragma unroll
        for (int i = 0; i < P / 4; ++i) reinterpret_cast<uint32_t*>(rw)[i] = 0u;
    } else if constexpr (P % 2 == 0) {
#pragma unroll
        for (int i = 0; i < P / 2; ++i) reinterpret_cast<uint16_t*>(rw)[i] = 0;
    } else {
#pragma unroll
        for (int p = 0; p < P; ++p) rw[p] = 0;
    }

    bool done = (meta0[0] & kSlotBits) == kSlotBits;   // an illegal step leaves the game as it was
    if (legal) {
        // comparison keys of the four rows (game.cuh::RowKeys) from the record, kept in this lane's indexable scratch
        // (game.cuh::place_indexed)
        {
            int w[kRows], u[kRows];
            const uint32_t metas = *reinterpret_cast<const uint32_t*>(rec + 20);
#pragma unroll
            for (int r = 0; r < kRows; ++r) {
                const uint32_t m = (metas >> (8 * r)) & 0xFFu;
                const uint32_t top = rec[4 * ((m & 7u) - 1u) + r];
                w[r] = (int)((top << 10) | (m << 2) | (uint32_t)r);
                u[r] = (int)(((m >> 3) << 2) | (uint32_t)r);
            }
            *reinterpret_cast<int4*>(keys_w) = make_int4(w[0], w[1], w[2], w[3]);
            *reinterpret_cast<int4*>(keys_u) = make_int4(u[0], u[1], u[2], u[3]);
        }

        int keys[P];
#pragma unroll
        for (int p = 0; p < P; ++p) keys[p] = (act[p] << 4) | p;
        sort_keys<P>(keys);   // env.py:124-125

#pragma unroll
        for (int i = 0; i < P; ++i) {
            const int card = keys[i] >> 4, player = keys[i] & 15;
            int row;
            uint32_t keep_len;
            const int pen = place_indexed(keys_w, keys_u, card, values[card], row, keep_len);   // env.py:126-134
            rec[4 * keep_len + row] = (uint8_t)card;   // the one byte of the record a placement changes
            if (pen != 0) {                            // a take (rare): env.py:167-169
#pragma unroll
                for (int p = 0; p < P; ++p) meta[p] += p == player ? (uint32_t)pen << kRecScoreShift : 0u;
                rw[player] = (uint8_t)(0 - pen);
            }
        }
#pragma unroll
        for (int p = 0; p < P; ++p) meta0[p * kTileGames] = meta[p];   // env.py:131: the played slots are empty now

        const int4 fw = *reinterpret_cast<const int4*>(keys_w);
        const int w_final[kRows] = {fw.x, fw.y, fw.z, fw.w};
        uint32_t new_metas = 0;
#pragma unroll
        for (int r = 0; r < kRows; ++r) new_metas |= (((uint32_t)w_final[r] >> 2) & 0xFFu) << (8 * r);
        *reinterpret_cast<uint32_t*>(rec + 20) = new_metas;
        done = (meta[0] & kSlotBits) == kSlotBits;     // env.py:246-249
    }
    // the per-game outputs go out as coalesced stores of the lanes (P reward bytes, two flag bytes): cheaper than three more
    // bulk copies by lane 0
    {
        int rew[P];
        load_bytes<P>(buf + L::kRewards, lane, rew);
        store_bytes<P>(rewards_out, 0, rew);
    }
    done_out[0] = done;
    if (illegal_out) illegal_out[0] = !legal;
}

template <int P, bool kRandom>
__global__ void __launch_bounds__(kSmemWarps * 32, 8)
k_step_smem(StateView s, const uint8_t* __restrict__ actions, uint8_t* __restrict__ actions_out, int8_t* __restrict__ rewards,
            uint8_t* __restrict__ done, uint8_t* __restrict__ illegal, int64_t num_tiles, uint64_t seed, uint32_t turn, uint64_t game0) {
    using L = TileLayout<P>;
    extern __shared__ __align__(128) uint8_t tile_smem[];   // kSmemWarps x 2 x L::kStride
    __shared__ uint64_t full[kSmemWarps][2];
    __shared__ uint8_t values[128];
    __shared__ int4 keys_w[kSmemWarps * 32], keys_u[kSmemWarps * 32];   // each lane's row keys, indexable

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint8_t* bufs = tile_smem + warp * 2 * L::kStride;
    const int64_t stride = (int64_t)gridDim.x * kSmemWarps;
    int64_t tile = (int64_t)blockIdx.x * kSmemWarps + warp;

    if (lane == 0) {
        mbar_init(&full[warp][0], 1);
        mbar_init(&full[warp][1], 1);
        fence_barrier_init();
        if (tile < num_tiles) issue_tile_loads<P, kRandom>(s, actions, tile, bufs, &full[warp][0]);
        if (tile + stride < num_tiles) issue_tile_loads<P, kRandom>(s, actions, tile + stride, bufs + L::kStride, &full[warp][1]);
    }
    stage_card_values(values);
    __syncthreads();   // the only block-wide barrier: value table + barrier init

    for (int it = 0; tile < num_tiles; tile += stride, ++it) {
        const int b = it & 1;
        uint8_t* buf = bufs + b * L::kStride;
        mbar_wait(&full[warp][b], (uint32_t)(it >> 1) & 1u);
        const int64_t g = tile * kTileGames + lane;
        step_in_smem<P, kRandom>(buf, lane, values, reinterpret_cast<int*>(&keys_w[threadIdx.x]), reinterpret_cast<int*>(&keys_u[threadIdx.x]),
                                 reinterpret_cast<uint8_t*>(rewards) + g * P, done + g, illegal ? illegal + g : nullptr, seed, game0 + (uint64_t)g, turn);
        fence_async_smem();   // make this lane's shared-memory writes visible to the TMA engine
        __syncwarp();
        if (lane == 0) {
            issue_tile_stores<P>(s, rewards, done, illegal, actions_out, tile, buf);
            if (tile + 2 * stride < num_tiles) {
                bulk_wait_read0();   // the engine has read the buffer: it may be refilled
                issue_tile_loads<P, kRandom>(s, actions, tile + 2 * stride, buf, &full[warp][b]);
            }
        }
        __syncwarp();
    }
    if (lane == 0) bulk_wait_all();   // stores must land before the block's shared memory is released
}

// kRandom = false: actions is the tape to play; true: actions (may be NULL) receives the cards drawn in the kernel.
template <int P, bool kRandom>
static int launch_step(const StateView& s, uint8_t* actions, int8_t* rewards, uint8_t* done, uint8_t* illegal, uint64_t seed, uint32_t turn,
                       uint64_t game0, cudaStream_t st) {
    using L = TileLayout<P>;
    constexpr int kSmem = kSmemWarps * 2 * L::kStride;
    const int64_t num_tiles = s.B / kTileGames;
    if (num_tiles > 0) {
        static int occ_cache[kMaxDevices];   // per device: the shared-memory opt-in and the occupancy are device properties
        const int blocks_per_sm = blocks_per_sm_cached(k_step_smem<P, kRandom>, kSmemWarps * 32, kSmem, occ_cache);
        const int num_sms = device_sms(current_device());
        // persistent grid: one resident wave; every warp walks tiles warp_id, warp_id + #warps, ...
        const int64_t want = (num_tiles + kSmemWarps - 1) / kSmemWarps;
        const unsigned blocks = (unsigned)min(want, (int64_t)num_sms * blocks_per_sm);
        k_step_smem<P, kRandom><<<blocks, kSmemWarps * 32, kSmem, st>>>(s, actions, kRandom ? actions : nullptr, rewards, done, illegal, num_tiles, seed,
                                                                        turn, game0);
    }
    const int64_t tail0 = num_tiles * kTileGames;
    if (tail0 < s.B)   // ragged tail (< 32 games): plain loads
        k_step<P, kRandom><<<1, kStepThreads, 0, st>>>(s, actions, kRandom ? actions : nullptr, rewards, done, illegal, seed, turn, game0, tail0);
    return 0;
}

// ------------------------------------------------------------------------------------------
// k_random_actions — DrunkHamster.forward for every (game, player) (agents/random.py:8-10).
// ------------------------------------------------------------------------------------------
template <int P>
__global__ void __launch_bounds__(kStepThreads)
k_random_actions(StateView s, uint8_t* __restrict__ actions, uint64_t seed, uint32_t turn, uint64_t game0) {
    const int64_t g = (int64_t)blockIdx.x * kStepThreads + threadIdx.x;
    if (g >= s.B) return;
    GameRec<P> gm;
    load_hands<P>(s, g, gm.hand);
    int act[P];
    random_actions_game<P>(gm, seed, game0 + (uint64_t)g, turn, act);
    store_bytes<P>(actions, g, act);
}

// ------------------------------------------------------------------------------------------
// k_pack_flags — flag bytes -> bits (one ballot per 32 games), for results that travel over PCIe.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_pack_flags(const uint8_t* __restrict__ flags, uint32_t* __restrict__ bits, int64_t B) {
    const int64_t g = (int64_t)blockIdx.x * 256 + threadIdx.x;   // B is padded to whole warps by the launch
    const uint32_t word = __ballot_sync(0xffffffffu, g < B && flags[g] != 0);
    if ((threadIdx.x & 31) == 0 && g < B) bits[g >> 5] = word;
}

}  // namespace nimmt

using namespace nimmt;

extern "C" {

int nimmt_step(void* state, const uint8_t* actions, int8_t* rewards, uint8_t* done, uint8_t* illegal, int64_t B,
               int num_players, void* stream) {
    if (int rc = check_common(state, B, num_players)) return rc;
    if (!actions || !rewards || !done) return NIMMT_E_BADARG;
    if (!aligned16(actions) || !aligned16(rewards) || !aligned16(done) || (illegal && !aligned16(illegal))) return NIMMT_E_ALIGN;
    if (B == 0) return NIMMT_OK;
    StateView s(state, B, num_players);
    NIMMT_DISPATCH_P(num_players, (launch_step<P, false>(s, const_cast<uint8_t*>(actions), rewards, done, illegal, 0, 0, 0, (cudaStream_t)stream)));
    return check_launch();
}

int nimmt_pack_flags(const uint8_t* flags, uint32_t* bits, int64_t B, void* stream) {
    if (!flags || !bits || B < 0) return NIMMT_E_BADARG;
    if (reinterpret_cast<uintptr_t>(bits) & 3u) return NIMMT_E_ALIGN;
    if (B == 0) return NIMMT_OK;
    k_pack_flags<<<blocks_for(B, 256), 256, 0, (cudaStream_t)stream>>>(flags, bits, B);
    return check_launch();
}

int nimmt_step_random(void* state, uint8_t* actions, int8_t* rewards, uint8_t* done, int64_t B, int num_players,
                      uint64_t seed, uint32_t turn, uint64_t game0, void* stream) {
    if (int rc = check_common(state, B, num_players)) return rc;
    if (!rewards || !done) return NIMMT_E_BADARG;
    if ((actions && !aligned16(actions)) || !aligned16(rewards) || !aligned16(done)) return NIMMT_E_ALIGN;
    if (B == 0) return NIMMT_OK;
    StateView s(state, B, num_players);
    NIMMT_DISPATCH_P(num_players, (launch_step<P, true>(s, actions, rewards, done, nullptr, seed, turn, game0, (cudaStream_t)stream)));
    return check_launch();
}

int nimmt_random_actions(const void* state, uint8_t* actions, int64_t B, int num_players, uint64_t seed, uint32_t turn,
                         uint64_t game0, void* stream) {
    if (int rc = check_common(state, B, num_players)) return rc;
    if (!actions) return NIMMT_E_BADARG;
    if (!aligned16(actions)) return NIMMT_E_ALIGN;
    if (B == 0) return NIMMT_OK;
    StateView s(const_cast<void*>(state), B, num_players);
    NIMMT_DISPATCH_P(num_players, k_random_actions<P><<<blocks_for(B, kStepThreads), kStepThreads, 0, (cudaStream_t)stream>>>(
                                      s, actions, seed, turn, game0));
    return check_launch();
}

}  // extern "C"
