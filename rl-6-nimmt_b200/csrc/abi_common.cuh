// abi_common.cuh — shared by the translation units that implement include/nimmt_b200.h:
// packed-state plane I/O, argument checks, the per-player-count dispatch.
#pragma once
#include <cstdio>

#include "../../include/nimmt_b200.h"
#include "step.cuh"

namespace nimmt {

constexpr int kStepThreads = 128;

// Stored form in, stored form out (coalesced: a warp's 32 games are contiguous in every plane).
template <int P>
__device__ __forceinline__ void load_hands(const StateView& s, int64_t g, HandRec (&hand)[P]) {
#pragma unroll
    for (int p = 0; p < P; ++p) {
        hand[p].lo = *s.cards_ptr(g, p);
        hand[p].meta = *s.meta_ptr(g, p);
    }
}

template <int P>
__device__ __forceinline__ void load_game(const StateView& s, int64_t g, GameRec<P>& gm) {
    load_hands<P>(s, g, gm.hand);
    load_rows(s, g, gm.board);
}

// After a step: the dealt cards are immutable, only the meta words and the row record go back.
template <int P>
__device__ __forceinline__ void store_step(const StateView& s, int64_t g, const GameRec<P>& gm) {
#pragma unroll
    for (int p = 0; p < P; ++p) *s.meta_ptr(g, p) = gm.hand[p].meta;
    store_rows(s, g, gm.board);
}

// After a deal / reset_to: everything.
template <int P>
__device__ __forceinline__ void store_game(const StateView& s, int64_t g, const GameRec<P>& gm) {
#pragma unroll
    for (int p = 0; p < P; ++p) *s.cards_ptr(g, p) = gm.hand[p].lo;
    store_step<P>(s, g, gm);
}

// The planes of one game exactly as stored: lets a kernel issue all its loads first and unpack later.
template <int P>
struct RawGame {
    HandRec hand[P];
    uint2 rows[3];
};

template <int P>
__device__ __forceinline__ void load_raw(const StateView& s, int64_t g, RawGame<P>& raw) {
    load_hands<P>(s, g, raw.hand);
    const uint2* r = s.rows_ptr(g);
#pragma unroll
    for (int k = 0; k < 3; ++k) raw.rows[k] = r[k];
}

template <int P>
__device__ __forceinline__ void unpack_raw(const RawGame<P>& raw, GameRec<P>& gm) {
#pragma unroll
    for (int p = 0; p < P; ++p) gm.hand[p] = raw.hand[p];
    gm.board.unpack((uint64_t)raw.rows[0].x | ((uint64_t)raw.rows[0].y << 32), (uint64_t)raw.rows[1].x | ((uint64_t)raw.rows[1].y << 32),
                    (uint64_t)raw.rows[2].x | ((uint64_t)raw.rows[2].y << 32));
}

// ------------------------------------------------------------------------------------------
// host-side helpers of the C entry points
// ------------------------------------------------------------------------------------------
extern thread_local char g_last_error[256];  // defined in env_reset.cu

inline int check_launch() {
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        snprintf(g_last_error, sizeof(g_last_error), "%s: %s", cudaGetErrorName(e), cudaGetErrorString(e));
        return NIMMT_E_CUDA;
    }
    return NIMMT_OK;
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

inline int check_common(const void* state, int64_t B, int P) {
    if (!state || B < 0 || P < 1 || P > kMaxPlayers) return NIMMT_E_BADARG;
    if (!aligned16(state)) return NIMMT_E_ALIGN;
    if (B > (int64_t)INT32_MAX * 64) return NIMMT_E_BADARG;
    return NIMMT_OK;
}

#define NIMMT_DISPATCH_P(P_, ...)                              \
    switch (P_) {                                              \
        case 1: { constexpr int P = 1; __VA_ARGS__; } break;   \
        case 2: { constexpr int P = 2; __VA_ARGS__; } break;   \
        case 3: { constexpr int P = 3; __VA_ARGS__; } break;   \
        case 4: { constexpr int P = 4; __VA_ARGS__; } break;   \
        case 5: { constexpr int P = 5; __VA_ARGS__; } break;   \
        case 6: { constexpr int P = 6; __VA_ARGS__; } break;   \
        case 7: { constexpr int P = 7; __VA_ARGS__; } break;   \
        case 8: { constexpr int P = 8; __VA_ARGS__; } break;   \
        case 9: { constexpr int P = 9; __VA_ARGS__; } break;   \
        case 10: { constexpr int P = 10; __VA_ARGS__; } break; \
        default: return NIMMT_E_BADARG;                        \
    }

// Per-device launch facts.  A process may drive several GPUs (one env per device): SM counts, opt-in shared-memory sizes
// (cudaFuncSetAttribute is per device) and occupancy are cached per device ordinal, never in a single static.
constexpr int kMaxDevices = 64;
inline int current_device() {
    int dev = 0;
    cudaGetDevice(&dev);
    return dev >= 0 && dev < kMaxDevices ? dev : 0;
}
inline int device_sms(int dev) {
    static int sms[kMaxDevices];   // benign race: the query is idempotent
    if (sms[dev] == 0) cudaDeviceGetAttribute(&sms[dev], cudaDevAttrMultiProcessorCount, dev);
    return sms[dev];
}
// Opts `kernel` in to `smem` bytes of dynamic shared memory on the current device (once per device) and returns how many
// of its blocks fit on one SM.
template <class Kernel>
inline int blocks_per_sm_cached(Kernel kernel, int threads, int smem, int (&cache)[kMaxDevices]) {
    const int dev = current_device();
    if (cache[dev] == 0) {
        cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);   // these kernels live in shared memory, not in L1
        int occ = 0;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, threads, smem);
        cache[dev] = occ > 0 ? occ : 1;
    }
    return cache[dev];
}

inline unsigned blocks_for(int64_t B, int threads) { return (unsigned)((B + threads - 1) / threads); }

}  // namespace nimmt
