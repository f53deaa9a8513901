// env_reset.cu — deal, deal_from_perm, reset_to, scores (SechsNimmtEnv.reset / _deal / reset_to,
// env.py:43-62, 99-112) and the library-wide helpers of the C ABI.  One thread per game.
#include "abi_common.cuh"

namespace nimmt {

thread_local char g_last_error[256] = "";

// ------------------------------------------------------------------------------------------
// k_deal — SechsNimmtEnv.reset/_deal (env.py:43-51, 99-112).  Write-only: (12 P + 24) B/game.
// ------------------------------------------------------------------------------------------
// One thread per game.  The per-thread decks are BYTES, interleaved so that every entry of a thread lives in that thread's
// own bank: entry j of thread (warp w, lane l) sits at byte 128 j + 4 l + w — a block of four warps shares 104 x 32 words,
// each word holding the same entry of the four warps' lane-l games.  A random access is conflict-free whatever it draws, its
// address is one multiply-add, and a deck costs 104 bytes of shared memory (the first version used a 32-bit word per
// card: 416 bytes per thread, 16 warps per SM).  The identity is written cooperatively, 26 words per thread.
// Finished hands and rows go straight to HBM (coalesced: a warp's games are contiguous in every plane).
constexpr int kDealThreads = 128;

template <int P>
struct DealToState {
    const StateView& s;
    int64_t g;
    const uint8_t* values;
    uint32_t row_cards = 0, row_metas = 0;
    __device__ __forceinline__ void hand(int p, const uint32_t (&c)[kHand]) {
        // rec_from_sorted for a full hand of ten: slots 0..7 in the uint2, slots 8 and 9 in the low bytes of the meta word
        *s.cards_ptr(g, p) = make_uint2(c[0] | (c[1] << 8) | (c[2] << 16) | (c[3] << 24), c[4] | (c[5] << 8) | (c[6] << 16) | (c[7] << 24));
        *s.meta_ptr(g, p) = c[8] | (c[9] << 8);
    }
    // two sorted hands, 16 bits per card, hand p in the low halves: every record word is gathered with byte permutes
    // (byte 0 / byte 2 of each register; bytes 1 and 3 are zero and serve as the zero source)
    __device__ __forceinline__ void hand_pair(int p, const Pair16 (&k)[kHand]) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const uint32_t pick = h ? 0x0062u : 0x0040u;   // {a.b0, b.b0} or {a.b2, b.b2} into the low half
            const uint32_t x = __byte_perm(__byte_perm(k[0].v, k[1].v, pick), __byte_perm(k[2].v, k[3].v, pick), 0x5410u);
            const uint32_t y = __byte_perm(__byte_perm(k[4].v, k[5].v, pick), __byte_perm(k[6].v, k[7].v, pick), 0x5410u);
            *s.cards_ptr(g, p + h) = make_uint2(x, y);
            *s.meta_ptr(g, p + h) = __byte_perm(k[8].v, k[9].v, h ? 0x1162u : 0x1140u);   // slots 8, 9; no empty bits, score 0
        }
    }
    __device__ __forceinline__ void row(int r, uint32_t card) {
        row_cards |= card << (8 * r);
        row_metas |= (1u | ((uint32_t)values[card] << 3)) << (8 * r);   // one card, its bull heads
        if (r == kRows - 1) {   // slot-major record: byte r = first card of row r, bytes 4..19 unused, bytes 20..23 the row metas
            uint2* rec = s.rows_ptr(g);
            rec[0] = make_uint2(row_cards, 0u);
            rec[1] = make_uint2(0u, 0u);
            rec[2] = make_uint2(0u, row_metas);
        }
    }
};

template <int P>
__global__ void __launch_bounds__(kDealThreads) k_deal(StateView s, uint64_t seed, uint64_t game0) {
    __shared__ uint8_t values[128];
    __shared__ __align__(16) uint8_t decks[kCards * kDealThreads];
    stage_card_values(values);
    {   // 16-byte chunk c = words 4 c .. 4 c + 3 = entry c / 8 of sixteen games = that card id in every byte; thread t
        // writes chunks t, t + 128, ... (832 chunks)
        uint32_t v = (threadIdx.x >> 3) * 0x01010101u;
#pragma unroll
        for (int c = threadIdx.x; c < kCards * 8; c += kDealThreads, v += (kDealThreads / 8) * 0x01010101u)
            reinterpret_cast<uint4*>(decks)[c] = make_uint4(v, v, v, v);
    }
    __syncthreads();
    const int64_t g = (int64_t)blockIdx.x * kDealThreads + threadIdx.x;
    if (g >= s.B) return;
    deal_game<P>(seed, game0 + (uint64_t)g, decks + (threadIdx.x & 31) * 4 + (threadIdx.x >> 5), 128, DealToState<P>{s, g, values});
}

// k_deal_from_perm — the same, from caller-supplied shuffled decks (env.py:103-112).
template <int P>
__global__ void __launch_bounds__(kStepThreads) k_deal_from_perm(StateView s, const uint8_t* __restrict__ perm) {
    __shared__ uint8_t values[128];
    stage_card_values(values);
    __syncthreads();
    const int64_t g = (int64_t)blockIdx.x * kStepThreads + threadIdx.x;
    if (g >= s.B) return;
    const uint8_t* deck = perm + g * kCards;
    GameRec<P> gm;
#pragma unroll
    for (int p = 0; p < P; ++p) {
        int cards[kHand];
#pragma unroll
        for (int i = 0; i < kHand; ++i) cards[i] = deck[p * kHand + i];
        set_dealt_hand<P>(gm, p, cards);
    }
#pragma unroll
    for (int r = 0; r < kRows; ++r) {
        const uint32_t card = deck[kCards - 1 - r];
        gm.board.set_row(r, card, card, 1u, values[card & 127]);
    }
    store_game<P>(s, g, gm);
}

// ------------------------------------------------------------------------------------------
// k_reset_to — SechsNimmtEnv.reset_to (env.py:53-62) from -1 padded board / hands arrays.
// ------------------------------------------------------------------------------------------
template <int P>
__global__ void __launch_bounds__(kStepThreads)
k_reset_to(StateView s, const int8_t* __restrict__ board, const int8_t* __restrict__ hands, uint8_t* __restrict__ invalid) {
    __shared__ uint8_t values[128];
    stage_card_values(values);
    __syncthreads();
    const int64_t g = (int64_t)blockIdx.x * kStepThreads + threadIdx.x;
    if (g >= s.B) return;
    GameRec<P> gm;
    uint4 seen = make_uint4(0, 0, 0, 0);
    bool bad = false;
    const int8_t* bsrc = board + g * (kRows * 6);
#pragma unroll
    for (int r = 0; r < kRows; ++r) {
        uint64_t cards = 0;
        uint32_t len = 0, sum = 0, top = 0;
        bool open = true;
#pragma unroll
        for (int i = 0; i < 6; ++i) {
            const int c = bsrc[r * 6 + i];
            if (c < 0) open = false;  // first -1 ends the row, as list building from the observation does (mcts.py:75-85)
            if (open) {
                const bool ok = c < kCards && i < 5 && !mask_has(seen, (uint32_t)c);
                bad = bad || !ok;
                if (ok) {
                    mask_set(seen, (uint32_t)c);
                    cards |= (uint64_t)(uint32_t)c << (8u * len);
                    ++len;
                    sum += values[c];
                    top = (uint32_t)c;
                }
            }
        }
        bad = bad || len == 0;
        if (len == 0) { len = 1; sum = values[0]; }  // keep the packed invariant 1 <= len <= 5
        gm.board.set_row(r, cards, top, len, sum);
    }
    const int8_t* hsrc = hands + g * (P * kHand);
#pragma unroll
    for (int p = 0; p < P; ++p) {
        uint4 h = make_uint4(0, 0, 0, 0);
#pragma unroll
        for (int i = 0; i < kHand; ++i) {
            const int c = hsrc[p * kHand + i];
            if (c >= 0) {
                const bool ok = c < kCards && !mask_has(seen, (uint32_t)c);
                bad = bad || !ok;
                if (ok) { mask_set(seen, (uint32_t)c); mask_set(h, (uint32_t)c); }
            }
        }
        gm.hand[p] = rec_from_mask(h);   // slots in ascending card order whatever order the caller listed them in
    }
    store_game<P>(s, g, gm);
    if (invalid) invalid[g] = bad;
}

// ------------------------------------------------------------------------------------------
// k_scores — SechsNimmtEnv._scores (env.py:32,167).
// ------------------------------------------------------------------------------------------
template <int P>
__global__ void __launch_bounds__(kStepThreads) k_scores(StateView s, uint8_t* __restrict__ scores) {
    const int64_t g = (int64_t)blockIdx.x * kStepThreads + threadIdx.x;
    if (g >= s.B) return;
    int sc[P];
#pragma unroll
    for (int p = 0; p < P; ++p) sc[p] = (int)(*s.meta_ptr(g, p) >> kRecScoreShift);
    store_bytes<P>(scores, g, sc);
}

}  // namespace nimmt

using namespace nimmt;

extern "C" {

int nimmt_abi_version(void) { return 3; }

const char* nimmt_last_cuda_error(void) { return g_last_error; }

size_t nimmt_state_bytes(int64_t num_games, int num_players) {
    if (num_games < 0 || num_players < 1 || num_players > kMaxPlayers) return 0;
    return (size_t)StateView::bytes(num_games, num_players);   // whole tiles of 32 games, (12 P + 24) bytes per game
}

int nimmt_obs_len(int include_summaries) { return include_summaries ? 47 : 35; }

int nimmt_card_value(int card) { return (card >= 0 && card < kCards) ? h_card_value[card] : -1; }

int nimmt_deal(void* state, int64_t B, int num_players, uint64_t seed, uint64_t game0, void* stream) {
    if (int rc = check_common(state, B, num_players)) return rc;
    if (B == 0) return NIMMT_OK;
    StateView s(state, B, num_players);
    NIMMT_DISPATCH_P(num_players, k_deal<P><<<blocks_for(B, kDealThreads), kDealThreads, 0, (cudaStream_t)stream>>>(s, seed, game0));
    return check_launch();
}

int nimmt_deal_from_perm(void* state, const uint8_t* perm, int64_t B, int num_players, void* stream) {
    if (int rc = check_common(state, B, num_players)) return rc;
    if (!perm) return NIMMT_E_BADARG;
    if (B == 0) return NIMMT_OK;
    StateView s(state, B, num_players);
    NIMMT_DISPATCH_P(num_players, k_deal_from_perm<P><<<blocks_for(B, kStepThreads), kStepThreads, 0, (cudaStream_t)stream>>>(s, perm));
    return check_launch();
}

int nimmt_reset_to(void* state, const int8_t* board, const int8_t* hands, uint8_t* invalid, int64_t B, int num_players,
                   void* stream) {
    if (int rc = check_common(state, B, num_players)) return rc;
    if (!board || !hands) return NIMMT_E_BADARG;
    if (B == 0) return NIMMT_OK;
    StateView s(state, B, num_players);
    NIMMT_DISPATCH_P(num_players,
                     k_reset_to<P><<<blocks_for(B, kStepThreads), kStepThreads, 0, (cudaStream_t)stream>>>(s, board, hands, invalid));
    return check_launch();
}

int nimmt_scores(const void* state, uint8_t* scores, int64_t B, int num_players, void* stream) {
    if (int rc = check_common(state, B, num_players)) return rc;
    if (!scores) return NIMMT_E_BADARG;
    if (!aligned16(scores)) return NIMMT_E_ALIGN;
    if (B == 0) return NIMMT_OK;
    StateView s(const_cast<void*>(state), B, num_players);
    NIMMT_DISPATCH_P(num_players, k_scores<P><<<blocks_for(B, kStepThreads), kStepThreads, 0, (cudaStream_t)stream>>>(s, scores));
    return check_launch();
}

}  // extern "C"
