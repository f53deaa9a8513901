// tma_copy_probe.cu — measurement tool (not product code): how fast can per-warp double-buffered
// TMA tile pipelines move the k_step traffic pattern, with no compute at all?  Compares the
// plane-per-player layout (several 512-byte streams per tile) with one contiguous chunk per tile,
// and 32- vs 128-game tiles.  Built and run by hand:
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I rl-6-nimmt_b200/csrc -o /tmp/probe profiles/tools/tma_copy_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include "tma.cuh"
using namespace nimmt;

// NCH chunks per tile, each CH bytes, chunk c of tile t at base[c] + t*CH.  Load all, store all.
template <int NCH, int CH, int WARPS>
__global__ void __launch_bounds__(WARPS * 32, 8) k_copy(uint8_t* const* src, uint8_t* const* dst, int64_t num_tiles) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ uint64_t full[WARPS][2];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr int STRIDE = NCH * CH;
    uint8_t* bufs = smem + warp * 2 * STRIDE;
    const int64_t stride = (int64_t)gridDim.x * WARPS;
    int64_t tile = (int64_t)blockIdx.x * WARPS + warp;
    auto load = [&](int64_t t, uint8_t* buf, uint64_t* bar) {
        mbar_arrive_expect_tx(bar, STRIDE);
        for (int c = 0; c < NCH; ++c) bulk_load(buf + c * CH, src[c] + t * CH, CH, bar);
    };
    if (lane == 0) {
        mbar_init(&full[warp][0], 1); mbar_init(&full[warp][1], 1); fence_barrier_init();
        if (tile < num_tiles) load(tile, bufs, &full[warp][0]);
        if (tile + stride < num_tiles) load(tile + stride, bufs + STRIDE, &full[warp][1]);
    }
    __syncthreads();
    for (int it = 0; tile < num_tiles; tile += stride, ++it) {
        const int b = it & 1;
        uint8_t* buf = bufs + b * STRIDE;
        mbar_wait(&full[warp][b], (uint32_t)(it >> 1) & 1u);
        // touch: each lane flips one byte so the store is not a pure pass-through
        buf[lane * 16] ^= 1;
        fence_async_smem();
        __syncwarp();
        if (lane == 0) {
            for (int c = 0; c < NCH; ++c) bulk_store(dst[c] + tile * CH, buf + c * CH, CH);
            bulk_commit();
            if (tile + 2 * stride < num_tiles) { bulk_wait_read0(); load(tile + 2 * stride, buf, &full[warp][b]); }
        }
        __syncwarp();
    }
    if (lane == 0) bulk_wait_all();
}

// plain vectorised copy for reference (grid-stride uint4)
__global__ void k_plain(const uint4* __restrict__ s, uint4* __restrict__ d, int64_t n) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) d[i] = s[i];
}

template <int NCH, int CH, int WARPS>
float run(const char* name, int64_t total_bytes, int nsets) {
    const int64_t tiles = total_bytes / (NCH * CH);
    std::vector<uint8_t*> bufs(nsets);
    std::vector<uint8_t**> tabs(nsets);
    for (int s = 0; s < nsets; ++s) {
        cudaMalloc(&bufs[s], total_bytes);
        cudaMemset(bufs[s], 1, total_bytes);
        uint8_t* h[NCH];
        for (int c = 0; c < NCH; ++c) h[c] = bufs[s] + (int64_t)c * tiles * CH;
        cudaMalloc(&tabs[s], sizeof(h));
        cudaMemcpy(tabs[s], h, sizeof(h), cudaMemcpyHostToDevice);
    }
    constexpr int SMEM = WARPS * 2 * NCH * CH;
    cudaFuncSetAttribute(k_copy<NCH, CH, WARPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
    int occ = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_copy<NCH, CH, WARPS>, WARPS * 32, SMEM);
    const int blocks = (int)std::min<int64_t>((tiles + WARPS - 1) / WARPS, 148LL * occ);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int i = 0; i < 8; ++i) k_copy<NCH, CH, WARPS><<<blocks, WARPS * 32, SMEM>>>(tabs[i % nsets], tabs[i % nsets], tiles);
    cudaEventRecord(e0);
    const int reps = 40;
    for (int i = 0; i < reps; ++i) k_copy<NCH, CH, WARPS><<<blocks, WARPS * 32, SMEM>>>(tabs[i % nsets], tabs[i % nsets], tiles);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    ms /= reps;
    printf("%-34s chunks=%d x %5d B warps/blk=%d occ=%d blocks=%d smem=%d : %.2f us  %.0f GB/s (read+write) err=%s\n", name, NCH, CH, WARPS, occ,
           blocks, SMEM, ms * 1e3, 2.0 * total_bytes / (ms * 1e-3) / 1e9, cudaGetErrorString(cudaGetLastError()));
    for (int s = 0; s < nsets; ++s) { cudaFree(bufs[s]); cudaFree(tabs[s]); }
    return ms;
}

int main(int argc, char** argv) {
    const int64_t total = (argc > 1 ? atoll(argv[1]) : 92LL) << 20;   // bytes moved each way per launch, like k_step<4> on 2^20 games (88 + 4)
    const int nsets = argc > 2 ? atoi(argv[2]) : 4;
    {   // reference: plain copy
        uint4 *a[8], *b[8];
        for (int s = 0; s < nsets; ++s) { cudaMalloc(&a[s], total); cudaMalloc(&b[s], total); cudaMemset(a[s], 1, total); }
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        for (int i = 0; i < 4; ++i) k_plain<<<148 * 8, 256>>>(a[i % nsets], b[i % nsets], total / 16);
        cudaEventRecord(e0);
        for (int i = 0; i < 40; ++i) k_plain<<<148 * 8, 256>>>(a[i % nsets], b[i % nsets], total / 16);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 40;
        printf("%-34s : %.2f us  %.0f GB/s (read+write)\n", "plain uint4 grid-stride copy", ms * 1e3, 2.0 * total / (ms * 1e-3) / 1e9);
        for (int s = 0; s < nsets; ++s) { cudaFree(a[s]); cudaFree(b[s]); }
    }
    run<6, 512, 4>("planes, 32-game tiles (6x512)", total / 3072 * 3072, nsets);
    run<1, 3072, 4>("one chunk, 32-game tiles (1x3072)", total / 3072 * 3072, nsets);
    run<6, 2048, 1>("planes, 128-game tiles (6x2048)", total / 12288 * 12288, nsets);
    run<1, 12288, 1>("one chunk, 128-game tiles", total / 12288 * 12288, nsets);
    run<6, 1024, 2>("planes, 64-game tiles (6x1024)", total / 6144 * 6144, nsets);
    run<1, 6144, 2>("one chunk, 64-game tiles", total / 6144 * 6144, nsets);
    run<1, 3072, 8>("one chunk, 32-game, 8 warps/blk", total / 3072 * 3072, nsets);
    return 0;
}
