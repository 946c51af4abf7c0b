// common.cuh — shared declarations of the sm_100a SIFT engine (libsiftcuda.so).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include <vector>

#include "../../include/siftcuda.h"

namespace sift {

constexpr int kOctaves = SIFT_NUM_OCTAVES;
constexpr int kScales = SIFT_SCALES_PER_OCTAVE;
constexpr int kGaussians = SIFT_NUM_GAUSSIANS;
constexpr int kDogs = SIFT_NUM_DOGS;
constexpr int kOriBins = SIFT_ORIENTATION_HISTOGRAM_BINS;
constexpr int kMaxTaps = SIFT_CONVOLUTION_WEIGHTS_LENGTH;

// Scan granularity: one CTA of 256 threads owns 2048 consecutive items (8 per thread).
constexpr int kScanThreads = 256;
constexpr int kScanItemsPerThread = 8;
constexpr int kScanChunk = kScanThreads * kScanItemsPerThread;

struct Taps {
    float w[kMaxTaps];
};

// Geometry + device planes of one octave. Planes are pitched linear float arrays laid out
// [frame][slice][y][x]; `pitch` is in floats and a multiple of 32 (128-byte rows).
struct OctaveDev {
    int w, h, pitch;
    float delta;
    float sigmas[kGaussians];
    float log2SigmaRatio;
    float* G;          // [frame][6][h][pitch]
    float* D;          // [frame][5][h][pitch]
    float2* grad;      // [frame][3][h][pitch]  slices 1..3 (orientation, magnitude)
    size_t plane;      // pitch * h
    // extrema bitmask: [frame][3][h][maskRowWords] inside the global mask array
    int maskRowWords;  // ceil(w / 32)
    int maskWords;     // 3 * h * maskRowWords
    int maskBlocks;    // ceil(maskWords / kScanChunk)
    int maskBlockStart;  // first scan block of this octave inside a frame
};

struct EngineParams {
    OctaveDev oct[kOctaves];
    int blocksPerFrame;        // scan blocks of one frame's masks (all octaves)
    float dogThreshold, edgeThreshold, maxOffset;
    int maxIterations, border;
    float lambdaOri, oriThreshold;
    int oriSmoothIterations;
};

// Candidate = SIFTExtremaResult (include/SIFTExtrema.h:14-18) packed: x | y << 15 | s << 30,
// plus the (frame, octave) segment it belongs to.
struct Candidate {
    uint32_t xys;
    int32_t seg;  // frame * 7 + octave
};
__host__ __device__ inline uint32_t packXYS(int x, int y, int s) {
    return (uint32_t)x | ((uint32_t)y << 15) | ((uint32_t)s << 30);
}

// Result columns (the wire format of include/siftcuda.h: SiftKeypointColumns /
// SiftDescriptorColumns). The pointers are device memory (staged path) or pinned host memory
// mapped into the device address space (host-buffer calls: the kernels' own stores are the D2H).
struct KeypointColumnsDev {
    float *absX, *absY, *sigma, *value, *subScale;
    short2* scaledXY;
    uchar2* octaveScale;
};
struct DescriptorColumnsDev {
    uint8_t* features;   // [n][128]
    float* theta;
    int32_t* keypoint;
};

// Device-side counters of one execute (zeroed at its start).
struct Counters {
    int nCandidates;
    int nKeypoints;
    int nDescriptors;
    int overflow;  // bit 0 candidates, bit 1 keypoints, bit 2 descriptors
    int oriNext;   // work queues of the orientation / descriptor kernels: warps take their first
    int descNext;  // item by position and every further one from these counters (zero per call)
};

// Programmatic dependent launch (PDL). A kernel launched through pdlLaunch() may become resident
// while the previous kernel of its stream is still running; pdlPrologue() at its very top lets
// its own successor start launching and then blocks until that previous kernel has completed
// and flushed. Both instructions are no-ops in a kernel launched the ordinary way. Used for the
// chains of small kernels (list compaction, small-plane blurs) whose cost is launch latency.
#ifdef __CUDACC__
__device__ __forceinline__ void pdlPrologue() {
    asm volatile("griddepcontrol.launch_dependents;\n" ::: "memory");
    asm volatile("griddepcontrol.wait;\n" ::: "memory");
}
// Every launch that needs an attribute goes through here. `priority` (kNoPriority = leave the
// stream's) is set as a per-launch attribute as well as through the stream it is launched on: a
// stream's priority is not carried into the kernel nodes of a captured CUDA graph, a launch
// attribute is.
constexpr int kNoPriority = 1 << 30;
template <class... KArgs, class... Args>
inline cudaError_t launchKernel(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                                bool pdl, int priority, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[2];
    static const bool pdlEnabled = !(getenv("SIFTCUDA_PDL") && atoi(getenv("SIFTCUDA_PDL")) == 0);   // tuning switch
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = (pdl && pdlEnabled) ? 1 : 0;
    cfg.numAttrs = 1;
    if (priority != kNoPriority) {
        attr[1].id = cudaLaunchAttributePriority;
        attr[1].val.priority = priority;
        cfg.numAttrs = 2;
    }
    cfg.attrs = attr;
    return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}
template <class... KArgs, class... Args>
inline cudaError_t pdlLaunch(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                             bool pdl, Args... args) {
    return launchKernel(kernel, grid, block, smem, st, pdl, kNoPriority, args...);
}
#endif

#define SIFT_CUDA_TRY(expr)                                  \
    do {                                                     \
        cudaError_t _e = (expr);                             \
        if (_e != cudaSuccess) return _e;                    \
    } while (0)

// ---- launchers (each returns cudaGetLastError of its launches) ---------------------------------

// pyramid.cu
// bytesPerPixel 4: BGRA8; 1: GRAY8 / the luma plane of NV12
cudaError_t launchGrayUpsample(const uint8_t* pixels, int bytesPerPixel, int pitchBytes,
                               int64_t frameStrideBytes, float* gray, int W, int H, float* scaled,
                               int w2, int h2, int pitch2, size_t scaledFrameStride, int frames,
                               cudaStream_t st);
// TMA tensor maps of one stack of planes for every blur configuration: [tap count 11/15/17/21/27]
// [tile 64 / 32][full row group / last row group]. Host memory; the two maps a launch needs
// travel as __grid_constant__ kernel parameters.
struct BlurTmaSet {
    CUtensorMap map[5][2][2];
    int valid;
};
cudaError_t makeBlurTmaSet(BlurTmaSet* set, const float* base, int pitch, int h, int nz, size_t planeFloats);

// out = blur(in); optional dog = out - in; optional decimated copy of out (every other pixel)
struct BlurArgs {
    const float* in;
    float* out;
    float* dog;      // may be null
    float* half;     // may be null: next octave's slice 0, receives out[2y][2x]
    int w, h, pitch;
    size_t inFrameStride, outFrameStride, dogFrameStride;
    int halfW, halfH, halfPitch;
    size_t halfFrameStride;
    int frames;
    int yBegin, yEnd;  // output rows [yBegin, yEnd) of the plane (0, 0 = all rows): row bands
    int debugMode;   // 0 normal; 1 skip the X/Y FMA loops; 2 skip the stores; 3 both (tuning only)
    int pdl;         // 1: programmatic dependent launch behind the previous kernel of the stream
                     //    (small planes: the launch latency of scale s + 1 overlaps scale s)
    // TMA tile loads: `in` = slice tmaZ0 (+ frame * tmaZStride) of the stack tmaSet describes
    const BlurTmaSet* tmaSet;   // host pointer, null = cp.async loads only
    int tmaZ0, tmaZStride;
    int tma;                    // set by the launcher
    int priority;               // launch priority (kNoPriority / 0 = the stream's)
};
cudaError_t launchBlur(const BlurArgs& a, const Taps& taps, int ntaps, cudaStream_t st);
// octaves >= tailStartOctave(P) run whole in one launch (one CTA per frame, planes in shared memory)
int tailStartOctave(const EngineParams& P);
cudaError_t launchTailOctaves(const EngineParams& P, int oStart, const Taps* taps, const int* ntaps, uint32_t* mask,
                              int frames, cudaStream_t st, int priority);
cudaError_t launchGradient(const OctaveDev& o, int frames, cudaStream_t st, int yBegin = 0, int yEnd = 0,
                           int priority = kNoPriority);

// detect.cu
cudaError_t launchExtremaMask(const EngineParams& P, int octave, uint32_t* mask, int frames,
                              cudaStream_t st, int yBegin = 0, int yEnd = 0, int priority = kNoPriority,
                              const CUtensorMap* dogMap = nullptr);
cudaError_t makeExtremaTmaMap(CUtensorMap* map, const float* base, int pitch, int h, int nz, size_t planeFloats);
// mask blocks [blockBegin, blockBegin + nBlocks) → ordered candidates; segStart[nSegs + 1] receives
// the per-(frame, octave) list offsets
cudaError_t launchCandidateCompaction(const EngineParams& P, const uint32_t* mask,
                                      int* blockSums, Candidate* cands, int capCandidates,
                                      int blockBegin, int nBlocks, int* segStart, int nSegs,
                                      Counters* counters, cudaStream_t st);
cudaError_t launchRefine(const EngineParams& P, const Candidate* cands, int capCandidates,
                         SiftKeypoint* kpTmp, uint32_t* flagWords, int* blockSums,
                         SiftKeypoint* kps, int* kpSeg, int capKeypoints, const int* segCandStart,
                         int* segKpStart, int nSegs, Counters* counters,
                         const KeypointColumnsDev& hostCols, cudaStream_t st);

// describe.cu
// Descriptor records go to `cols` (device columns, always) and, when hostCols.features is not
// null, to the pinned host columns as well.
cudaError_t launchDescribe(const EngineParams& P, const SiftKeypoint* kps, const int* kpSeg,
                           int capKeypoints, const int* segKpStart, int* nOri, float* oriTmp,
                           int* oriOffset, int* descKp, int* blockSums,
                           const DescriptorColumnsDev& cols, const DescriptorColumnsDev& hostCols,
                           int capDescriptors, int* segDescStart, int nSegs, Counters* counters,
                           int smCount, cudaStream_t stream, cudaEvent_t afterOrientation);

// match.cu
size_t matchScratchInts(int nSource, int nTarget, int smCount);
cudaError_t launchMatch(const uint8_t* source, int nSource, const uint8_t* target, int nTarget,
                        float absThr, float relThr, int* scratch, SiftMatch* rows, int smCount,
                        cudaStream_t st);

// geometry.cu (host stages: SIFTDescriptor.compareGeometry, the trie ANN matcher)
float compareGeometry(const SiftMatch* m, int64_t n, const float* sourceXY, const float* targetXY, int minimumSampleSize);
void approximateMatch(const uint8_t* source, int64_t nSource, const uint8_t* target, int64_t nTarget, float absThr,
                      float relThr, std::vector<SiftMatch>& out);

// math debug (capi.cu → describe.cu)
cudaError_t launchMathDebug(int op, const float* a, const float* b, float* out, int64_t n,
                            cudaStream_t st);

}  // namespace sift
