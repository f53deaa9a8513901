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
        int nl[P];
#pragma unroll
        for (int p = 0; p < P; ++p) {
            // own hand ascending, -1 padded at the end (env.py:209-210): the unplayed slots, in slot order
            int8_t* h = rec + p * L;
            int cnt = 0;
#pragma unroll
            for (int i = 0; i < kHand; ++i) {
                if (!((gm.hand[p].meta >> i) & 1u)) h[cnt++] = (int8_t)rec_card(gm.hand[p], i);
            }
            nl[p] = cnt;
            for (int i = cnt; i < kHand; ++i) h[i] = -1;
            // the block shared by all players of the game, behind the hand.  When a game's record is a whole number of
            // words (P % 4 == 0) its position in the staging buffer is word aligned for every thread, and the bytes are
            // stored as 32-bit words between the unaligned edges: a quarter of the shared-memory wavefronts of byte
            // stores, which is what bounds the narrow-dtype observation kernel.
            if constexpr ((P * L) % 4 == 0) {
                const int d = p * L + 10;                    // byte offset of the common block in the record (static)
                const int lead = (4 - (d & 3)) & 3, words = (L - 10 - lead) / 4, tail = (L - 10 - lead) % 4;
#pragma unroll
                for (int k = 0; k < lead; ++k) h[10 + k] = common[k];
                uint32_t* w = reinterpret_cast<uint32_t*>(h + 10 + lead);
#pragma unroll
                for (int j = 0; j < words; ++j) {
                    const int k = lead + 4 * j;
                    w[j] = (uint32_t)(uint8_t)common[k] | ((uint32_t)(uint8_t)common[k + 1] << 8) | ((uint32_t)(uint8_t)common[k + 2] << 16) |
                           ((uint32_t)(uint8_t)common[k + 3] << 24);
                }
#pragma unroll
                for (int k = 0; k < tail; ++k) h[10 + lead + 4 * words + k] = common[lead + 4 * words + k];
            } else {
#pragma unroll
                for (int k = 0; k < L - 10; ++k) h[10 + k] = common[k];
            }
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

}  // extern "C"
