// mcs_kernels.cu — Monte-Carlo search rollouts (BaseMCAgent._mcts, agents/mcts.py:91-154) for
// batches of decisions.  Integer-issue bound: the root (64 B) is read once per block; a rollout's state is ~20 registers
// plus two private shared-memory scratch areas (its 116-byte deck and its row keys, rollout.cuh); each block leaves
// three 64-bit atomics behind.
#include "abi_common.cuh"
#include "rollout.cuh"

namespace nimmt {

constexpr int kMcsThreads = 256;

static_assert(sizeof(nimmt_root) == 64, "nimmt_root must be 64 bytes");

// grid = (chunks, 10 candidate ranks, D roots).  Block (c, a, d) runs local rollouts
// [c * kMcsThreads * iters, (c + 1) * kMcsThreads * iters) of candidate a of root d; local rollout
// i is global rollout j = rank + i * world, and the RNG is keyed by (d, a, j) only.
template <int P>
__global__ void __launch_bounds__(kMcsThreads)
k_mcs_rollouts(const nimmt_root* __restrict__ roots, int64_t local_rollouts, int iters, uint64_t seed, int rank, int world,
               unsigned long long* __restrict__ stats) {
    __shared__ uint8_t values[128], values5[128];
    __shared__ nimmt_root root;
    __shared__ long long red[3][kMcsThreads / 32];
    stage_card_values(values);
    stage_card_values5(values5);
    const int d = blockIdx.z, a = blockIdx.y;
    if (threadIdx.x < 16) reinterpret_cast<uint32_t*>(&root)[threadIdx.x] = reinterpret_cast<const uint32_t*>(roots + d)[threadIdx.x];
    __syncthreads();

    __shared__ RolloutRoot rr;
    __shared__ int root_ok;
    __shared__ __align__(4) uint8_t decks[kMcsThreads * kRolloutDeckStride];
    __shared__ uint32_t keys_w[kRows * kMcsThreads], keys_u[kRows * kMcsThreads];   // row r of thread t at [r * threads + t]: conflict-free
    // cooperative version of make_rollout_root (rollout.cuh): thread c places card c at its rank in the
    // ascending pool / own lists, so building the 116-byte record costs ~20 instructions per thread
    {
        __shared__ uint4 s_own, s_pool;
        if (threadIdx.x == 0) {
            uint4 own, pool;
            root_ok = decode_root<P>(root, values, own, pool, rr.board) ? 1 : 0;
            s_own = own; s_pool = pool;
            rr.n_pool = mask_count(pool);
            rr.n_own = mask_count(own);
        }
        if (threadIdx.x < kRolloutDeckStride / 4) reinterpret_cast<uint32_t*>(rr.deck)[threadIdx.x] = 0u;
        __syncthreads();
        const uint32_t c = threadIdx.x;
        if (c < (uint32_t)kCards) {
            const uint32_t word = c >> 5, below_bits = (1u << (c & 31)) - 1u;
            auto rank_in = [&](const uint4& m) -> int {   // number of set cards below c
                const uint32_t w0 = m.x, w1 = m.y, w2 = m.z, w3 = m.w & kHighCardMask;
                int n = 0;
                n += word > 0 ? __popc(w0) : __popc(w0 & below_bits);
                if (word >= 1) n += word > 1 ? __popc(w1) : __popc(w1 & below_bits);
                if (word >= 2) n += word > 2 ? __popc(w2) : __popc(w2 & below_bits);
                if (word >= 3) n += __popc(w3 & below_bits);
                return n;
            };
            if (mask_has(s_pool, c)) rr.deck[rank_in(s_pool)] = (uint8_t)c;
            if (mask_has(s_own, c)) rr.deck[kOwnOffset + rank_in(s_own)] = (uint8_t)c;
        }
    }
    __syncthreads();
    if (!root_ok || a >= rr.n_own) return;   // whole block; stats stay zero
    uint8_t* deck = decks + threadIdx.x * kRolloutDeckStride;

    long long s = 0, ss = 0, cnt = 0;
    const int64_t base = ((int64_t)blockIdx.x * iters) * kMcsThreads + threadIdx.x;
    for (int it = 0; it < iters; ++it) {
        const int64_t i = base + (int64_t)it * kMcsThreads;
        if (i < local_rollouts) {
            const uint64_t j = (uint64_t)rank + (uint64_t)i * (uint64_t)world;
            const uint64_t id = ((uint64_t)d << 44) | ((uint64_t)a << 40) | j;
            const int out = rollout<P, kMcsThreads>(rr, a, values5, deck, keys_w + threadIdx.x, keys_u + threadIdx.x, seed, id);
            s += out; ss += (long long)out * out; cnt += 1;
        }
    }
    // warp reduce, then one atomic triple per block
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        s += __shfl_down_sync(0xFFFFFFFFu, s, off);
        ss += __shfl_down_sync(0xFFFFFFFFu, ss, off);
        cnt += __shfl_down_sync(0xFFFFFFFFu, cnt, off);
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) { red[0][warp] = s; red[1][warp] = ss; red[2][warp] = cnt; }
    __syncthreads();
    if (threadIdx.x < 3) {
        long long t = 0;
#pragma unroll
        for (int w = 0; w < kMcsThreads / 32; ++w) t += red[threadIdx.x][w];
        if (t != 0) atomicAdd(stats + ((int64_t)d * 10 + a) * 3 + threadIdx.x, (unsigned long long)t);
    }
}

}  // namespace nimmt

using namespace nimmt;

extern "C" {

int nimmt_mcs_rollouts(const nimmt_root* roots, int num_roots, int num_players, int64_t rollouts_per_action, uint64_t seed,
                       int rank, int world, int64_t* stats, void* stream) {
    if (!roots || !stats || num_roots < 0 || rollouts_per_action < 0 || world < 1 || rank < 0 || rank >= world || num_players < 1 ||
        num_players > kMaxPlayers)
        return NIMMT_E_BADARG;
    if (num_roots > 32768) {   // grid.z limit: split the batch (roots are independent; ids keep the global root index out of the RNG key only per chunk)
        for (int first = 0; first < num_roots; first += 32768) {
            const int n = num_roots - first < 32768 ? num_roots - first : 32768;
            const int rc = nimmt_mcs_rollouts(roots + first, n, num_players, rollouts_per_action, seed + 0x9E3779B97F4A7C15ull * (uint64_t)(first / 32768 + 1),
                                              rank, world, stats + (int64_t)first * 30, stream);
            if (rc) return rc;
        }
        return NIMMT_OK;
    }
    if (rollouts_per_action >= ((int64_t)1 << 40)) return NIMMT_E_BADARG;
    if ((reinterpret_cast<uintptr_t>(roots) & 15u) || (reinterpret_cast<uintptr_t>(stats) & 7u)) return NIMMT_E_ALIGN;
    const int64_t local = rollouts_per_action > rank ? (rollouts_per_action - rank + world - 1) / world : 0;
    if (num_roots == 0 || local == 0) return NIMMT_OK;
    // rollouts per thread: aim for ~6 resident waves of blocks (each block pays a root decode and three
    // barriers), then shrink `iters` to the smallest value that still covers `local` with that many chunks
    const int64_t total_threads = local * 10 * (int64_t)num_roots;
    int64_t target = total_threads / ((int64_t)kMcsThreads * 148 * 3 * 6);
    target = target < 1 ? 1 : target > 64 ? 64 : target;
    const int64_t chunks = (local + kMcsThreads * target - 1) / (kMcsThreads * target);
    const int iters = (int)((local + kMcsThreads * chunks - 1) / (kMcsThreads * chunks));
    if (chunks > 0x7FFFFFFF) return NIMMT_E_BADARG;
    dim3 grid((unsigned)chunks, 10, (unsigned)num_roots);
    unsigned long long* st = reinterpret_cast<unsigned long long*>(stats);
    cudaStream_t cs = (cudaStream_t)stream;
    switch (num_players) {
#define CASE(P_) case P_: k_mcs_rollouts<P_><<<grid, kMcsThreads, 0, cs>>>(roots, local, iters, seed, rank, world, st); break;
        CASE(1) CASE(2) CASE(3) CASE(4) CASE(5) CASE(6) CASE(7) CASE(8) CASE(9) CASE(10)
#undef CASE
        default: return NIMMT_E_BADARG;
    }
    return check_launch();
}

}  // extern "C"
