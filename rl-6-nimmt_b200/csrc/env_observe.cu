// env_observe.cu — observation vectors (SechsNimmtEnv._create_states, env.py:174-212).
#include "abi_common.cuh"

namespace nimmt {

// ------------------------------------------------------------------------------------------
// k_observe — SechsNimmtEnv._create_states (env.py:174-212).
// Each thread expands its game into shared memory in exactly the output order [p][k]; the block
// then streams its contiguous output region with 16-byte stores (the block's games are adjacent,
// so element e of the region is byte e of the staging buffer: no index arithmetic on the way out).
// ------------------------------------------------------------------------------------------
template <int P>
struct ObserveCfg {
    static constexpr int kThreads = P <= 5 ? 128 : 64;  // staging buffer <= 30 KB
};

template <typename T>
__device__ __forceinline__ T obs_cast(int v) { return (T)v; }

template <int P, typename T, bool kSummaries>
__global__ void __launch_bounds__(ObserveCfg<P>::kThreads)
k_observe(StateView s, T* __restrict__ obs, uint8_t* __restrict__ n_legal) {
    constexpr int TPB = ObserveCfg<P>::kThreads;
    constexpr int L = kSummaries ? 47 : 35;
    constexpr int REC = P * L;  // staged bytes per game
    __shared__ __align__(16) int8_t stage[TPB * REC];

    const int64_t g0 = (int64_t)blockIdx.x * TPB;
    const int64_t g = g0 + threadIdx.x;
    const int nb = (int)min((int64_t)TPB, s.B - g0);  // games in this block
    if (g < s.B) {
        GameRec<P> gm;
        load_game<P>(s, g, gm);
        int8_t* rec = stage + threadIdx.x * REC;
        // shared part: P, (len, top, sum per row)?, board 4x6 (env.py:188-204)
        int8_t common[L - 10];
        int n = 0;
        common[n++] = (int8_t)P;
        if constexpr (kSummaries) {
#pragma unroll
            for (int r = 0; r < kRows; ++r) common[n + r] = (int8_t)gm.board.k.len(r);
#pragma unroll
            for (int r = 0; r < kRows; ++r) common[n + 4 + r] = (int8_t)gm.board.k.top(r);
#pragma unroll
            for (int r = 0; r < kRows; ++r) common[n + 8 + r] = (int8_t)gm.board.k.sum(r);
            n += 12;
        }
#pragma unroll
        for (int r = 0; r < kRows; ++r) {
            const int len = gm.board.k.len(r);
#pragma unroll
            for (int i = 0; i < 6; ++i)
                common[n + r * 6 + i] = i < len ? (int8_t)((gm.board.cards[r] >> (8 * i)) & 0xFF) : (int8_t)-1;
        }
        // the common block as aligned words (zero padded), for the funnel shifts below
        constexpr int C = L - 10, CW = (C + 3) / 4;
        uint32_t cw[CW + 1];
#pragma unroll
        for (int j = 0; j < CW; ++j) {
            uint32_t v = 0;
#pragma unroll
            for (int b = 0; b < 4; ++b)
                if (4 * j + b < C) v |= (uint32_t)(uint8_t)common[4 * j + b] << (8 * b);
            cw[j] = v;
        }
        cw[CW] = 0;
        int nl[P];
#pragma unroll
        for (int p = 0; p < P; ++p) {
            int8_t* h = rec + p * L;
            // The block shared by all players of the game goes behind the hand, as 32-bit words: byte stores cost four
            // times the shared-memory wavefronts, which is what bounds the narrow-dtype observation kernel.  The block
            // starts s bytes into a word (s is the same for every thread when a record is a whole number of words, else it
            // depends on the thread): word k of the destination holds common[4 k - s ..], a funnel shift of two
            // neighbouring aligned words.  The first word's low s bytes land on this player's hand and are overwritten by
            // the hand stores that follow; the bytes past the last whole word are stored one by one so that nothing of
            // the next record is touched.
            {
                const uint32_t s = (uint32_t)(reinterpret_cast<uintptr_t>(h + 10) & 3u);
                uint32_t* w = reinterpret_cast<uint32_t*>(h + 10 - s);
                const uint32_t shift = 8u * (4u - s);          // 32 when s = 0: the clamped funnel shift returns the high word
                constexpr int kWhole = C / 4;                  // words that are whole for every s
#pragma unroll
                for (int k = 0; k < kWhole; ++k) w[k] = __funnelshift_rc(k ? cw[k - 1] : 0u, cw[k], shift);
                const uint32_t last = __funnelshift_rc(kWhole ? cw[kWhole - 1] : 0u, cw[kWhole], shift);   // common[4 kWhole - s ..]
                const int left = C - 4 * kWhole + (int)s;     // bytes of the block that are still to be written (1 .. 4 + ..)
                int8_t* t = reinterpret_cast<int8_t*>(w + kWhole);
#pragma unroll
                for (int b = 0; b < 4; ++b)
                    if (b < left) t[b] = (int8_t)(last >> (8 * b));
                if (left > 4) {                                // s pushed a fifth..seventh byte into one more word
                    const uint32_t more = __funnelshift_rc(cw[kWhole], cw[kWhole + 1 <= CW ? kWhole + 1 : CW], shift);
#pragma unroll
                    for (int b = 0; b < 3; ++b)
                        if (4 + b < left) t[4 + b] = (int8_t)(more >> (8 * b));
                }
            }
            // own hand ascending, -1 padded at the end (env.py:209-210): the unplayed slots, in slot order
            int cnt = 0;
#pragma unroll
            for (int i = 0; i < kHand; ++i) {
                if (!rec_slot_empty(gm.hand[p], i)) h[cnt++] = (int8_t)rec_card(gm.hand[p], i);
            }
            nl[p] = cnt;
            for (int i = cnt; i < kHand; ++i) h[i] = -1;
        }
        if (n_legal) store_bytes<P>(n_legal, g, nl);
    }
    __syncthreads();

    constexpr int VEC = 16 / (int)sizeof(T) > 16 ? 16 : 16 / (int)sizeof(T);  // elements per 16-byte store
    T* out = obs + g0 * REC;
    const int total = nb * REC;
    if (nb == TPB) {  // full block: total is a multiple of 16, the region is 16-byte aligned
        for (int e = threadIdx.x * VEC; e < total; e += TPB * VEC) {
            if constexpr (sizeof(T) == 1) {   // int8: the staged bytes are the output
                *reinterpret_cast<uint4*>(out + e) = *reinterpret_cast<const uint4*>(stage + e);
            } else {
                alignas(16) T v[VEC];
#pragma unroll
                for (int i = 0; i < VEC; ++i) v[i] = obs_cast<T>(stage[e + i]);
                *reinterpret_cast<uint4*>(out + e) = *reinterpret_cast<const uint4*>(v);
            }
        }
    } else {
        for (int e = threadIdx.x; e < total; e += TPB) out[e] = obs_cast<T>(stage[e]);
    }
}

template <int P, typename T>
static void launch_observe(const StateView& s, void* obs, uint8_t* n_legal, int64_t B, int summaries, cudaStream_t st) {
    constexpr int TPB = ObserveCfg<P>::kThreads;
    if (summaries)
        k_observe<P, T, true><<<blocks_for(B, TPB), TPB, 0, st>>>(s, reinterpret_cast<T*>(obs), n_legal);
    else
        k_observe<P, T, false><<<blocks_for(B, TPB), TPB, 0, st>>>(s, reinterpret_cast<T*>(obs), n_legal);
}

// ------------------------------------------------------------------------------------------
// k_step1 — the whole env.step of ONE game in one launch, for the B = 1 drop-in (SechsNimmtEnv, env.py:64-77): check and
// play the cards (passed by value: no host-to-device copy), then rewards, done, illegal, the cumulative scores and the
// observations of all P seats go out as ONE 512-byte record written with a single coalesced 16-byte store per lane — into
// mapped pinned host memory, so the host needs one stream synchronisation and no copy at all.  With do_step = 0 the
// record describes the current state (reset / reset_to).
//   record: [0, P) rewards int8 | [P] done | [P + 1] illegal | [16, 16 + P) scores uint8 | [32, 32 + P L) observations int8 [P][L]
// ------------------------------------------------------------------------------------------
struct Step1Cards {
    uint8_t card[16];
    uint8_t row[16];   // free-row-choice mode: the row each player takes on an undercut
};

template <int P>
__global__ void __launch_bounds__(32) k_step1(StateView s, int64_t g, Step1Cards cards, int do_step, int choose_rows, int summaries, uint4* __restrict__ out) {
    __shared__ uint8_t values[128];
    __shared__ __align__(16) uint8_t rec[512];
    __shared__ GameRec<P> gm;
    stage_card_values(values);
    reinterpret_cast<uint4*>(rec)[threadIdx.x] = make_uint4(0, 0, 0, 0);
    __syncwarp();
    if (threadIdx.x == 0) {
        load_game<P>(s, g, gm);
        bool legal = true;
        if (do_step) {
            int act[P], pen[P], choice[P];
#pragma unroll
            for (int p = 0; p < P; ++p) { act[p] = cards.card[p]; choice[p] = cards.row[p]; }
            legal = step_game<P>(gm, act, values, pen, choose_rows ? choice : nullptr);
            if (legal) store_step<P>(s, g, gm);
#pragma unroll
            for (int p = 0; p < P; ++p) rec[p] = (uint8_t)(legal ? -pen[p] : 0);
        }
        rec[P] = game_done<P>(gm);
        rec[P + 1] = !legal;
#pragma unroll
        for (int p = 0; p < P; ++p) rec[16 + p] = (uint8_t)rec_score(gm.hand[p]);
    }
    __syncwarp();
    const int L = summaries ? 47 : 35;
    if (threadIdx.x < P) fill_observation<P>(gm, threadIdx.x, reinterpret_cast<int8_t*>(rec) + 32 + threadIdx.x * L, summaries != 0);
    __syncwarp();
    out[threadIdx.x] = reinterpret_cast<const uint4*>(rec)[threadIdx.x];
}

}  // namespace nimmt

using namespace nimmt;

extern "C" {

int nimmt_observe(const void* state, void* obs, uint8_t* n_legal, int64_t B, int num_players, int include_summaries,
                  int dtype, void* stream) {
    if (int rc = check_common(state, B, num_players)) return rc;
    if (!obs || dtype < NIMMT_DT_I8 || dtype > NIMMT_DT_I64) return NIMMT_E_BADARG;
    if (!aligned16(obs) || (n_legal && !aligned16(n_legal))) return NIMMT_E_ALIGN;
    if (B == 0) return NIMMT_OK;
    StateView s(const_cast<void*>(state), B, num_players);
    cudaStream_t st = (cudaStream_t)stream;
    NIMMT_DISPATCH_P(num_players, switch (dtype) {
        case NIMMT_DT_I8: launch_observe<P, int8_t>(s, obs, n_legal, B, include_summaries, st); break;
        case NIMMT_DT_I16: launch_observe<P, int16_t>(s, obs, n_legal, B, include_summaries, st); break;
        case NIMMT_DT_F32: launch_observe<P, float>(s, obs, n_legal, B, include_summaries, st); break;
        default: launch_observe<P, int64_t>(s, obs, n_legal, B, include_summaries, st); break;
    });
    return check_launch();
}

int nimmt_step1(void* state, int64_t game, const uint8_t* cards_host, const uint8_t* rows_host, int num_players, int include_summaries, void* record,
                void* stream) {
    if (int rc = check_common(state, game + 1, num_players)) return rc;
    if (!record || game < 0) return NIMMT_E_BADARG;
    if (!aligned16(record)) return NIMMT_E_ALIGN;
    Step1Cards c;
    for (int p = 0; p < 16; ++p) {
        c.card[p] = cards_host && p < num_players ? cards_host[p] : (uint8_t)255;
        c.row[p] = rows_host && p < num_players ? rows_host[p] : (uint8_t)0;
    }
    StateView s(state, game + 1, num_players);
    NIMMT_DISPATCH_P(num_players, k_step1<P><<<1, 32, 0, (cudaStream_t)stream>>>(s, game, c, cards_host != nullptr, rows_host != nullptr, include_summaries,
                                                                                 static_cast<uint4*>(record)));
    return check_launch();
}

}  // extern "C"
