// match.cu — brute-force descriptor matching with the reference's ratio test on tcgen05.
//
// Replaces SIFTDescriptor.match(source:target:absoluteThreshold:relativeThreshold:)
// (SIFTDescriptor.swift:298-361), whose inner loop is vDSP.distanceSquared over 128 floats per
// (source, target) pair (Utilities/Vector.swift:226-235).
//
// ‖a − b‖² = ‖a‖² + ‖b‖² − 2 a·b. The n_source x n_target dot products of the uint8 feature rows
// are one dense GEMM: A = source rows, B = target rows, both K-major with K = 128 bytes = exactly
// one 128-byte swizzle row, accumulated exactly in int32 by tcgen05.mma kind::i8 (products
// <= 255², sums <= 128·255² < 2^31) into TMEM. The reference's scan is fused into the epilogue:
// every epilogue thread owns one source row (= one TMEM lane) and walks the targets in order
// keeping best (strict <), its first index and `second` = the best before the last improvement
// (SIFTDescriptor.swift:339-343 — not the true second smallest).
//
// Two persistent CTAs per SM, warp-specialised:
//   warp 0      TMEM allocation; TMA producer: A tile (128 rows x 128 B) per work item, B tiles (128 rows x 128 B)
//               through a 3-stage ring, SWIZZLE_128B tensor maps, mbarrier complete_tx
//   warp 1      MMA issuer: one elected lane issues 4 x tcgen05.mma (M 128, N 128, K 32) per B
//               tile into one of two TMEM accumulator stages (2 x 128 columns; two CTAs per SM
//               share the 512 TMEM columns), tcgen05.commit frees the smem stage and publishes
//               the accumulator
//   warps 2..5  epilogue: tcgen05.ld 32 lanes x 32 columns at a time, v = ‖b‖² − 2 a·b (‖a‖² is
//               constant per row and added at the end), min over groups of 8 columns and the
//               sequential update only when some lane's group minimum improves its best
// Work item = (block of 128 source rows, contiguous segment of target tiles); segments of one
// row block are merged in target order by matchFinishKernel with the same sequential rule, so
// the split does not change the result.
#include <algorithm>
#include <atomic>

#include "common.cuh"
#include "tma.cuh"

namespace sift {

namespace {

constexpr int kM = 128;          // source rows per work item (TMEM lanes)
constexpr int kN = 128;          // target rows per MMA tile (TMEM columns per accumulator stage)
constexpr int kK = 128;          // feature bytes per row = one swizzle row
#ifndef SIFT_MATCH_CTAS
#define SIFT_MATCH_CTAS 4
#endif
// The epilogue (a compare-and-select scan per element on the CUDA cores) outweighs the MMA, and a
// CTA has one epilogue warp per scheduler: several CTAs per SM put that many epilogue warps on
// every scheduler to cover each other's latencies. Four CTAs: 6 warps (<= 85 registers), one
// accumulator stage of 128 TMEM columns and two B stages (48 KB of shared memory) each.
constexpr int kCtasPerSm = SIFT_MATCH_CTAS;
constexpr int kStages = kCtasPerSm >= 4 ? 2 : 3;      // B tiles in flight
constexpr int kAccStages = kCtasPerSm >= 3 ? 1 : 2;   // kCtasPerSm * kAccStages * kN <= 512 TMEM columns
constexpr int kThreads = 192;    // warp 0 TMA + TMEM allocation, warp 1 MMA, warps 2..5 epilogue
constexpr int kSentinel = 0x3fffffff;   // "no distance yet"; also the padded target norm

constexpr uint32_t kABytes = kM * kK;
constexpr uint32_t kBBytes = kN * kK;
constexpr size_t kSmemBytes = 1024 /* alignment slack */ + kABytes + kStages * kBBytes + 256 /* barriers */;

__device__ __forceinline__ void umma(uint32_t tmemD, uint64_t descA, uint64_t descB, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t"
        "}" ::"r"(tmemD), "l"(descA), "l"(descB), "r"(idesc), "r"(accumulate), "r"(0u) : "memory");
}
__device__ __forceinline__ void ummaCommit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smemAddr(bar)) : "memory");
}
__device__ __forceinline__ void tmemLoad32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
          "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
          "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmemLoadWait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tcFenceBefore() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcFenceAfter() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// K-major operand tile, rows of 128 bytes, SWIZZLE_128B: 8-row groups 1024 bytes apart
// (cute::UMMA::SmemDescriptor: start >> 4 at [0,14), LBO >> 4 at [16,30) (1: unused for swizzled
// K-major), SBO >> 4 at [32,46), version 1 at [46,48), layout SWIZZLE_128B = 2 at [61,64)).
__device__ __forceinline__ uint64_t smemDesc(uint32_t addr) {
    return (uint64_t)((addr >> 4) & 0x3fffu) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
// cute::UMMA::InstrDescriptor for kind::i8: D = S32 (2 at [4,6)), A = B = unsigned 8 bit (0 at
// [7,10), [10,13)), both K-major (0 at [15], [16]), N >> 3 at [17,23), M >> 4 at [24,29).
constexpr uint32_t kIdesc = (2u << 4) | ((uint32_t)(kN >> 3) << 17) | ((uint32_t)(kM >> 4) << 24);

struct MatchParams {
    int nSource, nTarget;
    int nRowBlocks, nSegs, tilesPerSeg, nColTiles;
    const int* normB;      // [nColTiles * kN], padded with kSentinel
    int* segBest;          // [nSegs][nRowBlocks * kM]
    int* segIndex;
    int* segSecond;
};

// Sequential update of SIFTDescriptor.swift:339-343 for one target.
__device__ __forceinline__ void scanOne(int v, int j, int& best, int& index, int& second) {
    if (v < best) {
        second = best;
        best = v;
        index = j;
    }
}

__global__ void __launch_bounds__(kThreads, kCtasPerSm)
matchKernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
            const MatchParams p) {
    extern __shared__ uint8_t smemRaw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smemRaw + 1023) & ~(uintptr_t)1023);
    uint8_t* sA = smem;
    uint8_t* sB = smem + kABytes;
    uint64_t* bars = (uint64_t*)(smem + kABytes + kStages * kBBytes);
    uint64_t* fullB = bars;                    // [kStages]
    uint64_t* emptyB = bars + kStages;         // [kStages]
    uint64_t* fullA = bars + 2 * kStages;      // [1]
    uint64_t* emptyA = fullA + 1;              // [1]
    uint64_t* tmemFull = emptyA + 1;           // [kAccStages]
    uint64_t* tmemEmpty = tmemFull + kAccStages;
    uint32_t* tmemBaseSlot = (uint32_t*)(tmemEmpty + kAccStages);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (warp == 1 && lane == 0) {
        for (int s = 0; s < kStages; s++) { mbarInit(&fullB[s], 1); mbarInit(&emptyB[s], 1); }
        mbarInit(fullA, 1);
        mbarInit(emptyA, 1);
        for (int s = 0; s < kAccStages; s++) { mbarInit(&tmemFull[s], 1); mbarInit(&tmemEmpty[s], 4); }
        mbarInitFence();
    } else if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smemAddr(tmemBaseSlot)),
                     "r"((uint32_t)(kAccStages * kN)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tcFenceBefore();
    __syncthreads();
    tcFenceAfter();
    const uint32_t tmemBase = *tmemBaseSlot;

    const int nItems = p.nRowBlocks * p.nSegs;

    if (warp == 0) {
        if (lane == 0) {
            uint32_t it = 0, itemCount = 0;
            for (int item = blockIdx.x; item < nItems; item += gridDim.x, itemCount++) {
                const int rb = item % p.nRowBlocks, seg = item / p.nRowBlocks;
                mbarWait(emptyA, (itemCount & 1) ^ 1);
                mbarExpectTx(fullA, kABytes);
                tmaLoad2d(&mapA, fullA, sA, 0, rb * kM);
                const int t0 = seg * p.tilesPerSeg, t1 = min(t0 + p.tilesPerSeg, p.nColTiles);
                for (int t = t0; t < t1; t++, it++) {
                    const uint32_t s = it % kStages, ph = (it / kStages) & 1;
                    mbarWait(&emptyB[s], ph ^ 1);
                    mbarExpectTx(&fullB[s], kBBytes);
                    tmaLoad2d(&mapB, &fullB[s], sB + s * kBBytes, 0, t * kN);
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            uint32_t it = 0, itemCount = 0;
            const uint32_t aBase = smemAddr(sA);
            for (int item = blockIdx.x; item < nItems; item += gridDim.x, itemCount++) {
                const int seg = item / p.nRowBlocks;
                mbarWait(fullA, itemCount & 1);
                const int t0 = seg * p.tilesPerSeg, t1 = min(t0 + p.tilesPerSeg, p.nColTiles);
                for (int t = t0; t < t1; t++, it++) {
                    const uint32_t s = it % kStages, ph = (it / kStages) & 1;
                    const uint32_t acc = it % kAccStages, accPh = (it / kAccStages) & 1;
                    mbarWait(&tmemEmpty[acc], accPh ^ 1);
                    mbarWait(&fullB[s], ph);
                    tcFenceAfter();
                    const uint32_t bBase = smemAddr(sB + s * kBBytes);
#pragma unroll
                    for (int k = 0; k < kK / 32; k++)
                        umma(tmemBase + acc * kN, smemDesc(aBase + k * 32), smemDesc(bBase + k * 32), kIdesc, k > 0);
                    ummaCommit(&emptyB[s]);        // the smem stage is free once these MMAs have read it
                    ummaCommit(&tmemFull[acc]);    // ... and the accumulator is complete
                }
                ummaCommit(emptyA);
            }
        }
    } else if (warp >= 2) {
        const int q = warp & 3;                        // TMEM lane quarter this warp may read (warps 2..5: 2, 3, 0, 1)
        const uint32_t laneBase = (uint32_t)(q * 32) << 16;
        uint32_t it = 0;
        for (int item = blockIdx.x; item < nItems; item += gridDim.x) {
            const int rb = item % p.nRowBlocks, seg = item / p.nRowBlocks;
            int best = kSentinel, index = -1, second = kSentinel;
            const int t0 = seg * p.tilesPerSeg, t1 = min(t0 + p.tilesPerSeg, p.nColTiles);
            for (int t = t0; t < t1; t++, it++) {
                const uint32_t acc = it % kAccStages, accPh = (it / kAccStages) & 1;
                mbarWait(&tmemFull[acc], accPh);
                tcFenceAfter();
                const uint32_t taddr = tmemBase + acc * kN + laneBase;
                const int4* __restrict__ nb4 = reinterpret_cast<const int4*>(p.normB + (size_t)t * kN);
#pragma unroll 1
                for (int c0 = 0; c0 < kN; c0 += 32) {
                    uint32_t d[32];
                    tmemLoad32(taddr + c0, d);
                    int nb[32];
#pragma unroll
                    for (int g = 0; g < 8; g++) {
                        const int4 n4 = __ldg(nb4 + (c0 >> 2) + g);   // same address in every lane: one broadcast
                        nb[4 * g] = n4.x; nb[4 * g + 1] = n4.y; nb[4 * g + 2] = n4.z; nb[4 * g + 3] = n4.w;
                    }
                    tmemLoadWait();
                    int v[32];
#pragma unroll
                    for (int k = 0; k < 32; k++) v[k] = nb[k] - 2 * (int)d[k];
#pragma unroll
                    for (int g = 0; g < 4; g++) {
                        const int m01 = min(v[8 * g], v[8 * g + 1]), m23 = min(v[8 * g + 2], v[8 * g + 3]);
                        const int m45 = min(v[8 * g + 4], v[8 * g + 5]), m67 = min(v[8 * g + 6], v[8 * g + 7]);
                        const int gmin = min(min(m01, m23), min(m45, m67));
                        // improvements are rare (O(log n) per row): only then walk the 8 targets in order
                        if (__any_sync(0xffffffffu, gmin < best)) {
                            const int j0 = t * kN + c0 + 8 * g;
#pragma unroll
                            for (int k = 0; k < 8; k++) scanOne(v[8 * g + k], j0 + k, best, index, second);
                        }
                    }
                }
                tcFenceBefore();
                __syncwarp();
                if (lane == 0) mbarArrive(&tmemEmpty[acc]);
            }
            const size_t o = (size_t)seg * p.nRowBlocks * kM + (size_t)rb * kM + q * 32 + lane;
            p.segBest[o] = best;
            p.segIndex[o] = index;
            p.segSecond[o] = second;
        }
    }

    tcFenceBefore();
    __syncthreads();
    if (warp == 0) {
        tcFenceAfter();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmemBase),
                     "r"((uint32_t)(kAccStages * kN)) : "memory");
    }
}

// ‖row‖² of a [n][128] uint8 matrix; rows [n, nPadded) receive `pad`.
__global__ void __launch_bounds__(256)
featureNormsKernel(const uint8_t* __restrict__ f, int n, int nPadded, int pad, int* __restrict__ norms) {
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row >= nPadded) return;
    if (row >= n) {
        if (lane == 0) norms[row] = pad;
        return;
    }
    const uint32_t w = __ldg(reinterpret_cast<const uint32_t*>(f + (size_t)row * kK) + lane);
    int s = (int)__dp4a(w, w, 0u);
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) s += __shfl_xor_sync(0xffffffffu, s, d);
    if (lane == 0) norms[row] = s;
}

// Merges the target segments of every source row in target order with the sequential rule (a
// segment whose minimum improves on the running best makes `second` = min(running best, the
// segment's own prefix minimum)), then applies the reference's two tests:
//   distance = sqrt(distanceSquared) over features / 255   (SIFTDescriptor.swift:37-41, Vector.swift:237-239)
//   keep iff best < absoluteThreshold and best < second * relativeThreshold   (:353-359)
__global__ void __launch_bounds__(256)
matchFinishKernel(const MatchParams p, const int* __restrict__ normA, float absThr, float relThr,
                  SiftMatch* __restrict__ rows) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= p.nSource) return;
    int best = kSentinel, index = -1, second = kSentinel;
    const size_t stride = (size_t)p.nRowBlocks * kM;
    for (int s = 0; s < p.nSegs; s++) {
        const int m = p.segBest[s * stride + i];
        if (m < best) {
            second = min(best, p.segSecond[s * stride + i]);
            best = m;
            index = p.segIndex[s * stride + i];
        }
    }
    SiftMatch r;
    r.source = i;
    r.target = -1;
    r.distance = 0.0f;
    if (index >= 0) {
        const int na = normA[i];
        const float dBest = __fdiv_rn(__fsqrt_rn((float)(best + na)), 255.0f);
        const float dSecond = second == kSentinel ? 3.402823466e+38f   // .greatestFiniteMagnitude (:332)
                                                  : __fdiv_rn(__fsqrt_rn((float)(second + na)), 255.0f);
        r.distance = dBest;
        if (dBest < absThr && dBest < __fmul_rn(dSecond, relThr)) r.target = index;
    }
    rows[i] = r;
}

// [rows][128] uint8, box = `boxRows` whole rows, 128-byte swizzle (one row = one swizzle span).
cudaError_t featureMap(CUtensorMap* map, const uint8_t* base, int rows, int boxRows) {
    auto enc = tensorMapEncoder();
    if (!enc) return cudaErrorNotSupported;
    const cuuint64_t dims[2] = {(cuuint64_t)kK, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)kK};
    const cuuint32_t box[2] = {(cuuint32_t)kK, (cuuint32_t)boxRows};
    const cuuint32_t elem[2] = {1, 1};
    const CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<uint8_t*>(base), dims, strides, box, elem,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? cudaSuccess : cudaErrorInvalidValue;
}

}  // namespace

size_t matchScratchInts(int nSource, int nTarget, int smCount) {
    const int nRowBlocks = (nSource + kM - 1) / kM, nColTiles = (nTarget + kN - 1) / kN;
    const int nSegs = std::max(1, std::min(nColTiles, (2 * kCtasPerSm * smCount + nRowBlocks - 1) / std::max(nRowBlocks, 1)));
    return (size_t)nRowBlocks * kM /* normA */ + (size_t)nColTiles * kN /* normB */ +
           3 * (size_t)nSegs * nRowBlocks * kM;
}

// rows[i] for every source row (target = -1: rejected). `scratch` holds matchScratchInts ints.
cudaError_t launchMatch(const uint8_t* source, int nSource, const uint8_t* target, int nTarget,
                        float absThr, float relThr, int* scratch, SiftMatch* rows, int smCount,
                        cudaStream_t st) {
    if (nSource < 1 || nTarget < 1) return cudaSuccess;
    MatchParams p;
    p.nSource = nSource;
    p.nTarget = nTarget;
    p.nRowBlocks = (nSource + kM - 1) / kM;
    p.nColTiles = (nTarget + kN - 1) / kN;
    // enough (row block, segment) items to fill the SMs about twice; one segment when the source
    // side alone does
    p.nSegs = std::max(1, std::min(p.nColTiles, (2 * kCtasPerSm * smCount + p.nRowBlocks - 1) / p.nRowBlocks));
    p.tilesPerSeg = (p.nColTiles + p.nSegs - 1) / p.nSegs;
    p.nSegs = (p.nColTiles + p.tilesPerSeg - 1) / p.tilesPerSeg;
    int* normA = scratch;
    int* normB = normA + (size_t)p.nRowBlocks * kM;
    p.normB = normB;
    p.segBest = normB + (size_t)p.nColTiles * kN;
    p.segIndex = p.segBest + (size_t)p.nSegs * p.nRowBlocks * kM;
    p.segSecond = p.segIndex + (size_t)p.nSegs * p.nRowBlocks * kM;

    featureNormsKernel<<<(p.nRowBlocks * kM + 7) / 8, 256, 0, st>>>(source, nSource, p.nRowBlocks * kM, 0, normA);
    featureNormsKernel<<<(p.nColTiles * kN + 7) / 8, 256, 0, st>>>(target, nTarget, p.nColTiles * kN, kSentinel, normB);
    SIFT_CUDA_TRY(cudaGetLastError());

    CUtensorMap mapA, mapB;
    SIFT_CUDA_TRY(featureMap(&mapA, source, nSource, kM));
    SIFT_CUDA_TRY(featureMap(&mapB, target, nTarget, kN));
    static std::atomic<unsigned long long> configured{0};
    int dev = 0;
    cudaGetDevice(&dev);
    if (!((configured.load(std::memory_order_acquire) >> (dev & 63)) & 1ull)) {
        SIFT_CUDA_TRY(cudaFuncSetAttribute(matchKernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes));
        configured.fetch_or(1ull << (dev & 63), std::memory_order_release);
    }
    const int nItems = p.nRowBlocks * p.nSegs;
    matchKernel<<<std::min(nItems, kCtasPerSm * smCount), kThreads, kSmemBytes, st>>>(mapA, mapB, p);
    SIFT_CUDA_TRY(cudaGetLastError());
    matchFinishKernel<<<(nSource + 255) / 256, 256, 0, st>>>(p, normA, absThr, relThr, rows);
    return cudaGetLastError();
}

}  // namespace sift
