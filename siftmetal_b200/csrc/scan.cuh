// scan.cuh — deterministic stream compaction primitives (count → scan → scatter).
//
// The reference appends to its lists with unordered atomics (SIFTExtrema.metal:92-108), so its
// output order changes run to run. Here every list is compacted by an exclusive scan, which
// fixes the canonical order (frame, octave, scale, y, x) by construction and needs no sort.
// A CTA of 256 threads owns kScanChunk = 2048 consecutive items, 8 consecutive per thread.
#pragma once
#include "common.cuh"

namespace sift {

// Exclusive scan of one int per thread across a 256-thread CTA. Returns the exclusive prefix;
// *total receives the CTA sum. `sh` is 9 ints of shared memory.
__device__ __forceinline__ int blockExclusiveScan256(int v, int* sh, int* total) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    int inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int n = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= d) inc += n;
    }
    if (lane == 31) sh[wid] = inc;
    __syncthreads();
    if (wid == 0) {
        int ws = lane < 8 ? sh[lane] : 0;
        int winc = ws;
#pragma unroll
        for (int d = 1; d < 8; d <<= 1) {
            const int n = __shfl_up_sync(0xffffffffu, winc, d);
            if (lane >= d) winc += n;
        }
        if (lane < 8) sh[lane] = winc - ws;
        if (lane == 7) sh[8] = winc;
    }
    __syncthreads();
    const int r = inc - v + sh[wid];
    *total = sh[8];
    __syncthreads();
    return r;
}

// Phase A: blockSums[b] = sum of value(i) over the chunk of block b. `Value` is a functor
// int operator()(int i) that already returns 0 beyond the live range.
template <class Value>
__global__ void __launch_bounds__(kScanThreads) scanBlockSumsKernel(Value value, int* blockSums) {
    pdlPrologue();
    __shared__ int sh[9];
    const int base = blockIdx.x * kScanChunk + threadIdx.x * kScanItemsPerThread;
    int s = 0;
#pragma unroll
    for (int k = 0; k < kScanItemsPerThread; k++) s += value(base + k);
    int total;
    blockExclusiveScan256(s, sh, &total);
    if (threadIdx.x == 0) blockSums[blockIdx.x] = total;
}

// Phase B: in-place exclusive scan of n block sums by a single CTA of 1024 threads;
// *totalOut = grand total clamped to `capacity` (overflow bit set in *overflow when clamped).
// Optional: segment starts of the compacted list straight from the scanned sums, when the items
// are laid out segment by segment in whole blocks (the extrema mask: [frame][octave] block ranges).
struct MaskSegments {
    int* segStart = nullptr;      // [nSegs + 1] out; null = none
    int nSegs = 0;
    int blocksPerFrame = 0;
    int blockBegin = 0;           // the scan covers global blocks [blockBegin, blockBegin + n)
    int octaveBlockStart[kOctaves] = {};
};
cudaError_t launchScanOffsets(int* blockSums, int n, int* totalOut, int capacity, int* overflow,
                              int overflowBit, cudaStream_t st, const MaskSegments* segs = nullptr);

}  // namespace sift
