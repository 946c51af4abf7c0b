// capi.cu — context lifetime, per-call orchestration and the extern "C" surface of
// include/siftcuda.h. Host stages restated here (each cites the reference):
//   DifferenceOfGaussians.init / Octave.init   (DifferenceOfGaussians.swift:233-344, :69-147)
//   GaussianKernel / GaussianSeriesKernel taps (GaussianKernel.swift:20-43, GaussianSeriesKernel.swift:27-51)
//   SIFT.getKeypoints / getDescriptors         (SIFT.swift:147-238)
// One context = one device, one stream; every call is synchronous at return. There is no CPU
// path: without a usable sm_100 device every compute entry point fails.
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <new>
#include <string>
#include <vector>

#include "common.cuh"

using namespace sift;

struct SiftContext {
    SiftConfig cfg{};
    int device = 0;
    int smCount = 0;
    cudaStream_t stream = nullptr;
    // octave o >= 1 runs its blur chain + gradient + extrema mask on its own stream as soon as
    // octave o-1 has produced Gaussian slice 3 (fork / join around the main stream)
    cudaStream_t octStream[kOctaves]{};

    // Large octaves are split into two row bands that run as independent chains of blur launches
    // (each band recomputes the few halo rows the later scales need, so there is no dependency
    // between the bands): the fixed cost of a launch — latency, first wave's exposed tile load,
    // partial last wave — of one chain hides under the other chain's work.
    static constexpr int kMaxBands = 4;
    int nBands = 2;                                   // SIFTCUDA_BANDS overrides (tuning)
    cudaStream_t bandStream[kMaxBands]{};             // [0] unused: band 0 runs on the octave stream
    cudaEvent_t evBandFork = nullptr;
    cudaEvent_t evBandSeeded[kMaxBands]{}, evBandDone[kMaxBands]{};
    cudaEvent_t evSeeded[kOctaves]{};   // octave o's slice 3 (and octave o+1's slice 0) written
    cudaEvent_t evOctDone[kOctaves]{};
    SiftInfo info{};
    EngineParams P{};
    Taps seedTaps{};
    int seedNtaps = 0;
    Taps taps[kGaussians - 1]{};
    int ntaps[kGaussians - 1]{};

    int B = 1;       // max_batch
    int nSegs = 0;   // B * 7
    int capCand = 0, capKp = 0, capDesc = 0;  // totals over the batch

    // device memory
    std::vector<void*> allocations;
    int64_t deviceBytes = 0;
    uint8_t* dInput = nullptr;
    float* dGray = nullptr;
    float* dScaled = nullptr;
    int pitch2 = 0;
    uint32_t* dMask = nullptr;
    // Device lists + their bookkeeping. Two sets: on a single large frame octave 0 is compacted,
    // refined and described from set 0 as soon as its extrema mask exists, while the deeper
    // octaves (a long latency-bound chain of small launches) are still running; they follow from
    // set 1. Batches and the two-step API use set 0 alone.
    struct ListSet {
        int* dBlockSums = nullptr;
        Candidate* dCands = nullptr;
        SiftKeypoint* dKpTmp = nullptr;
        uint32_t* dFlagWords = nullptr;
        SiftKeypoint* dKps = nullptr;
        int* dKpSeg = nullptr;
        int* dSegStarts = nullptr;  // [3][nSegs + 1]: candidates, keypoints, descriptors
        int* dNOri = nullptr;
        float* dOriTmp = nullptr;
        int* dOriOffset = nullptr;
        int* dDescKp = nullptr;
        SiftDescriptor* dDesc = nullptr;
        Counters* dCounters = nullptr;
        Counters* hCounters = nullptr;   // pinned
        int* hSegStarts = nullptr;       // pinned
    } L[2];
    // Host-resident results (the fused host-buffer entry points): the descriptor kernel stores its
    // 136-byte records straight into the pinned result array over PCIe while it runs, and the
    // keypoints leave on a copy stream as soon as refinement has counted them — no D2H pass after
    // the last kernel.
    // A large single frame is uploaded in two row chunks on the copy stream; the seed stage and
    // octave 0's first row band run on chunk A while chunk B is still crossing PCIe.
    // One upload chunk per row band: chunk k ends with the last input row band k's chain needs.
    // grayRows / upRows / seedRows[k] = input rows uploaded / upsampled rows / seed rows complete once
    // chunks 0..k are in (cumulative; the last chunk completes the planes). n = 0: whole frame at once.
    struct SeedSplit { int n = 0; int grayRows[kMaxBands] = {}, upRows[kMaxBands] = {}, seedRows[kMaxBands] = {}; } upSplit;
    cudaEvent_t evBandBlurEnd[kMaxBands]{};   // timing: end of each band's blur chain
    cudaEvent_t evUp[kMaxBands]{};
    cudaEvent_t evSeedDone[kMaxBands]{};
    bool countersClean = false;      // both sets' device counters were zeroed after the last call
    bool wantHostOut = false;        // request for the next runDetect / describe
    bool descOnHost = false;         // last describe wrote c->hDesc directly
    bool kpsOnHost = false;          // last detect's keypoints were already copied to c->hKps
    cudaStream_t copyStream = nullptr;
    cudaEvent_t evRefined = nullptr, evKpCopied = nullptr;
    bool split = false;              // results of the last call live in both sets
    bool candSplit = false;          // the last detect put octaves >= 1 in set 1 (debug taps)
    cudaEvent_t evB[5]{};            // stage boundaries of set 1's pass

    // pinned host memory
    SiftKeypoint* hKps = nullptr;
    SiftDescriptor* hDesc = nullptr;
    int* hKpSeg = nullptr;
    std::vector<int32_t> kpCounts, descCounts, candCounts;

    // current input
    const uint8_t* curInput = nullptr;
    int curPitch = 0;
    int64_t curFrameStride = 0;
    int curFrames = 0;
    bool executed = false;   // pyramid + gradients of curFrames frames are on the device
    bool described = false;

    // timing
    bool stageTiming = true;
    cudaEvent_t ev[SIFT_STAGE_COUNT + 1]{};
    cudaEvent_t evBlur0[kGaussians]{};
    SiftTimings timings{};
    int launches = 0;
    bool bandedOctave0 = false;

    std::string lastError;
};

namespace {

int fail(SiftContext* c, int status, const char* what, cudaError_t e = cudaSuccess) {
    if (c) {
        char buf[512];
        if (e != cudaSuccess)
            snprintf(buf, sizeof buf, "%s: %s (%s)", what, cudaGetErrorName(e), cudaGetErrorString(e));
        else
            snprintf(buf, sizeof buf, "%s", what);
        c->lastError = buf;
    }
    return status;
}

#define CTX_TRY(c, expr)                                                   \
    do {                                                                   \
        cudaError_t _e = (expr);                                           \
        if (_e != cudaSuccess) return fail((c), SIFT_ERR_CUDA, #expr, _e); \
    } while (0)

// GaussianKernel.swift:20-43 ≡ GaussianSeriesKernel.swift:27-51 — host float arithmetic.
int gaussianWeights(float s, float* out) {
    const int radius = (int)ceilf(4 * s);
    const int size = radius * 2 + 1;
    if (size > kMaxTaps) return -1;
    float t = 0;
    const float ss = s * s;
    for (int k = -radius; k <= radius; k++) {
        const float kk = (float)(k * k);
        const float w = expf(-0.5f * (kk / ss));
        out[k + radius] = w;
        t += w;
    }
    for (int i = 0; i < size; i++) out[i] = out[i] / t;
    return size;
}

template <class T>
cudaError_t devAlloc(SiftContext* c, T** p, size_t count) {
    void* q = nullptr;
    const size_t bytes = std::max<size_t>(count * sizeof(T), 256);
    cudaError_t e = cudaMalloc(&q, bytes);
    if (e != cudaSuccess) return e;
    c->allocations.push_back(q);
    c->deviceBytes += (int64_t)bytes;
    *p = (T*)q;
    return cudaSuccess;
}

// Octave 0 of a large single frame runs as independent row-band blur chains (runDetect).
bool bandedOctave0(const SiftContext* c, int frames) {
    const OctaveDev& q = c->P.oct[0];
    const long tiles = (long)((q.w + 63) / 64) * ((q.h + 63) / 64) * frames;
    const int nb = c->nBands;
    return frames == 1 && nb > 1 && tiles >= 8L * c->smCount && q.h >= 256 * nb;
}
// Gradient field and extrema mask of octave 0 per row band, right behind that band's blur chain
// (each band then carries one more halo row: the mask of its edge rows reads the DoG row beyond).
bool bandTails() {
    // measured at 1080p: 1065 vs 1058 frames/s device-resident, 864 vs 872 end to end (the e2e
    // critical path is upload -> last band's first three blurs -> octave 1 .. 6) — opt-in
    static const bool on = getenv("SIFTCUDA_BAND_TAILS") && atoi(getenv("SIFTCUDA_BAND_TAILS")) != 0;
    return on;
}
int bandBoundary(const SiftContext* c, int band) {   // first row of `band` (multiple of the tile height)
    const OctaveDev& q = c->P.oct[0];
    if (band >= c->nBands) return q.h;
    return (int)(((long)q.h * band / c->nBands + 63) / 64 * 64);
}

// Rows of the input / upsampled / seed image that band 0's chain depends on: everything the
// first upload chunk must deliver.
SiftContext::SeedSplit seedSplitFor(const SiftContext* c, int frames) {
    SiftContext::SeedSplit sp;
    static const bool enabled = !(getenv("SIFTCUDA_UPLOAD_SPLIT") && atoi(getenv("SIFTCUDA_UPLOAD_SPLIT")) == 0);
    if (!enabled || !bandedOctave0(c, frames)) return sp;
    const OctaveDev& q = c->P.oct[0];
    const int nb = c->nBands, H = c->cfg.height;
    int sumR = 0;
    for (int t = 0; t < kGaussians - 1; t++) sumR += c->ntaps[t] / 2;
    for (int k = 0; k < nb; k++) {
        if (k + 1 == nb) {
            sp.seedRows[k] = q.h; sp.upRows[k] = q.h; sp.grayRows[k] = H;
        } else {
            sp.seedRows[k] = bandBoundary(c, k + 1) + sumR + (bandTails() ? 1 : 0);
            sp.upRows[k] = (sp.seedRows[k] + c->seedNtaps / 2 + 1) & ~1;   // even: whole gray rows (fused kernel)
            sp.grayRows[k] = sp.upRows[k] / 2 + 1;                         // input rows 0 .. upRows / 2
            const bool grows = k == 0 || (sp.grayRows[k] > sp.grayRows[k - 1] && sp.seedRows[k] > sp.seedRows[k - 1]);
            if (!grows || sp.seedRows[k] >= q.h || sp.upRows[k] >= q.h || sp.grayRows[k] >= H) return SiftContext::SeedSplit{};
        }
    }
    sp.n = nb;
    return sp;
}

void destroy(SiftContext* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    // drain every stream of the context (a chunked upload may still be in flight on the copy stream)
    if (c->stream) cudaStreamSynchronize(c->stream);
    if (c->copyStream) cudaStreamSynchronize(c->copyStream);
    for (int o = 1; o < kOctaves; o++)
        if (c->octStream[o]) cudaStreamSynchronize(c->octStream[o]);
    for (int b = 1; b < SiftContext::kMaxBands; b++)
        if (c->bandStream[b]) cudaStreamSynchronize(c->bandStream[b]);
    for (void* p : c->allocations) cudaFree(p);
    for (int k = 0; k < 2; k++) {
        if (c->L[k].hCounters) cudaFreeHost(c->L[k].hCounters);
        if (c->L[k].hSegStarts) cudaFreeHost(c->L[k].hSegStarts);
    }
    for (auto& e : c->evB)
        if (e) cudaEventDestroy(e);
    if (c->hKps) cudaFreeHost(c->hKps);
    if (c->hDesc) cudaFreeHost(c->hDesc);
    if (c->hKpSeg) cudaFreeHost(c->hKpSeg);
    for (auto& e : c->ev)
        if (e) cudaEventDestroy(e);
    for (auto& e : c->evBlur0)
        if (e) cudaEventDestroy(e);
    for (int o = 0; o < kOctaves; o++) {
        if (c->evSeeded[o]) cudaEventDestroy(c->evSeeded[o]);
        if (c->evOctDone[o]) cudaEventDestroy(c->evOctDone[o]);
        if (o > 0 && c->octStream[o]) cudaStreamDestroy(c->octStream[o]);

    }
    for (int b = 1; b < SiftContext::kMaxBands; b++) {
        if (c->bandStream[b]) cudaStreamDestroy(c->bandStream[b]);
        if (c->evBandSeeded[b]) cudaEventDestroy(c->evBandSeeded[b]);
        if (c->evBandDone[b]) cudaEventDestroy(c->evBandDone[b]);
    }
    if (c->evBandFork) cudaEventDestroy(c->evBandFork);
    for (auto& e : c->evUp)
        if (e) cudaEventDestroy(e);
    for (auto& e : c->evSeedDone)
        if (e) cudaEventDestroy(e);
    for (auto& e : c->evBandBlurEnd)
        if (e) cudaEventDestroy(e);
    if (c->evRefined) cudaEventDestroy(c->evRefined);
    if (c->evKpCopied) cudaEventDestroy(c->evKpCopied);
    if (c->copyStream) cudaStreamDestroy(c->copyStream);
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
}

}  // namespace

extern "C" {

int sift_config_default(SiftConfig* cfg, int32_t width, int32_t height) {
    if (!cfg) return SIFT_ERR_INVALID_ARGUMENT;
    memset(cfg, 0, sizeof *cfg);
    cfg->width = width;
    cfg->height = height;
    cfg->max_batch = 1;
    cfg->dog_threshold = 0.0133f;
    cfg->edge_threshold = 10.0f;
    cfg->max_interpolation_iterations = 5;
    cfg->max_offset = 0.6f;
    cfg->image_border = 5;
    cfg->lambda_orientation = 1.5f;
    cfg->orientation_threshold = 0.8f;
    cfg->orientation_smoothing_iterations = 6;
    return SIFT_OK;
}

const char* sift_status_string(int status) {
    switch (status) {
        case SIFT_OK: return "ok";
        case SIFT_ERR_INVALID_ARGUMENT: return "invalid argument";
        case SIFT_ERR_NO_DEVICE: return "no usable sm_100 CUDA device";
        case SIFT_ERR_CUDA: return "CUDA runtime error";
        case SIFT_ERR_CAPACITY: return "device list capacity exceeded (results truncated)";
        case SIFT_ERR_NOT_DETECTED: return "describe called before detect";
        case SIFT_ERR_OUT_OF_MEMORY: return "out of device memory";
        default: return "unknown status";
    }
}

const char* sift_last_error_string(const SiftContext* c) { return c ? c->lastError.c_str() : ""; }

int sift_create(const SiftConfig* cfg, int device, SiftContext** out) {
    if (!cfg || !out) return SIFT_ERR_INVALID_ARGUMENT;
    *out = nullptr;
    if (cfg->width < 8 || cfg->height < 8 || cfg->width > 16384 || cfg->height > 16384 ||
        cfg->max_batch < 1)
        return SIFT_ERR_INVALID_ARGUMENT;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || device < 0 || device >= count)
        return SIFT_ERR_NO_DEVICE;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return SIFT_ERR_NO_DEVICE;
    if (prop.major != 10) return SIFT_ERR_NO_DEVICE;  // the library carries sm_100a code only
    if (cudaSetDevice(device) != cudaSuccess) return SIFT_ERR_NO_DEVICE;

    SiftContext* c = new (std::nothrow) SiftContext();
    if (!c) return SIFT_ERR_OUT_OF_MEMORY;
    c->cfg = *cfg;
    c->device = device;
    c->smCount = prop.multiProcessorCount;
    c->B = cfg->max_batch;
    c->nSegs = c->B * kOctaves;
    const int W = cfg->width, H = cfg->height;

    // ---- schedule: DifferenceOfGaussians.init (:233-344) ------------------------------------
    const float sigmaMinimum = 0.8f, deltaMinimum = 0.5f, sigmaInput = 0.5f;
    SiftInfo& I = c->info;
    I.width = W;
    I.height = H;
    I.max_batch = c->B;
    I.sm_count = c->smCount;
    {
        const float i = sigmaMinimum * sigmaMinimum;
        const float j = sigmaInput * sigmaInput;
        I.seed_sigma = sqrtf(i - j) / deltaMinimum;
        c->seedNtaps = gaussianWeights(I.seed_sigma, c->seedTaps.w);
        I.seed_taps = c->seedNtaps;
        memcpy(I.seed_weights, c->seedTaps.w, sizeof I.seed_weights);
    }
    int blockStart = 0;
    for (int o = 0; o < kOctaves; o++) {
        OctaveDev& q = c->P.oct[o];
        q.delta = deltaMinimum * powf(2, (float)o);
        q.w = (int)((float)W / q.delta);
        q.h = (int)((float)H / q.delta);
        q.pitch = (q.w + 31) / 32 * 32;
        q.plane = (size_t)q.pitch * q.h;
        for (int s = 0; s < kGaussians; s++) {
            const float h = q.delta / deltaMinimum;
            const float i = (float)s / (float)kScales;
            const float j = powf(2, i);
            q.sigmas[s] = (h * sigmaMinimum) * j;
            I.sigmas[o][s] = q.sigmas[s];
        }
        q.log2SigmaRatio = log2f(q.sigmas[1] / q.sigmas[0]);  // SIFTOctave.swift:211
        q.maskRowWords = (q.w + 31) / 32;
        q.maskWords = kScales * q.h * q.maskRowWords;
        q.maskBlocks = (q.maskWords + kScanChunk - 1) / kScanChunk;
        q.maskBlockStart = blockStart;
        blockStart += q.maskBlocks;
        I.octave_width[o] = q.w;
        I.octave_height[o] = q.h;
        I.octave_pitch[o] = q.pitch;
        I.octave_delta[o] = q.delta;
    }
    c->P.blocksPerFrame = blockStart;
    for (int s = 1; s < kGaussians; s++) {
        // Octave.init (:91-110): rho is identical for every octave; octave 0's values are used
        const float sa = c->P.oct[0].sigmas[s - 1], sb = c->P.oct[0].sigmas[s];
        const float rho = sqrtf((sb * sb) - (sa * sa)) / c->P.oct[0].delta;
        I.rho[s - 1] = rho;
        c->ntaps[s - 1] = gaussianWeights(rho, c->taps[s - 1].w);
        I.taps[s - 1] = c->ntaps[s - 1];
        memcpy(I.weights[s - 1], c->taps[s - 1].w, sizeof I.weights[s - 1]);
    }
    if (c->seedNtaps != 11 || c->ntaps[0] != 11 || c->ntaps[1] != 15 || c->ntaps[2] != 17 ||
        c->ntaps[3] != 21 || c->ntaps[4] != 27) {
        delete c;
        return SIFT_ERR_INVALID_ARGUMENT;  // blur kernels are instantiated for the fixed schedule
    }
    c->P.dogThreshold = cfg->dog_threshold;
    c->P.edgeThreshold = cfg->edge_threshold;
    c->P.maxOffset = cfg->max_offset;
    c->P.maxIterations = cfg->max_interpolation_iterations;
    c->P.border = cfg->image_border;
    c->P.lambdaOri = cfg->lambda_orientation;
    c->P.oriThreshold = cfg->orientation_threshold;
    c->P.oriSmoothIterations = cfg->orientation_smoothing_iterations;

    // ---- capacities -------------------------------------------------------------------------
    const double N = (double)W * H;
    auto cap = [&](int32_t given, double perPixel, int floor_) -> int {
        const double perFrame = given > 0 ? (double)given : std::max((double)floor_, N * perPixel);
        const double total = perFrame * c->B;
        return (int)std::min(total, 1.0e9);
    };
    c->capCand = cap(cfg->max_candidates_per_frame, 0.08, 16384);
    c->capKp = cap(cfg->max_keypoints_per_frame, 0.04, 8192);
    c->capDesc = cap(cfg->max_descriptors_per_frame, 0.06, 8192);
    I.max_candidates_per_frame = c->capCand / c->B;
    I.max_keypoints_per_frame = c->capKp / c->B;
    I.max_descriptors_per_frame = c->capDesc / c->B;

    // ---- allocation (all at create, as SIFT.init; nothing is allocated per frame) -----------
    cudaError_t e = cudaSuccess;
    auto A = [&](cudaError_t r) { if (e == cudaSuccess) e = r; };
    A(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    const size_t B = (size_t)c->B;
    A(devAlloc(c, &c->dInput, B * (size_t)W * H * 4));
    A(devAlloc(c, &c->dGray, B * (size_t)W * H));
    c->pitch2 = c->P.oct[0].pitch;
    A(devAlloc(c, &c->dScaled, B * c->P.oct[0].plane));
    for (int o = 0; o < kOctaves; o++) {
        OctaveDev& q = c->P.oct[o];
        A(devAlloc(c, &q.G, B * kGaussians * q.plane));
        A(devAlloc(c, &q.D, B * kDogs * q.plane));
        A(devAlloc(c, &q.grad, B * kScales * q.plane));
    }
    const size_t maskWords = B * (size_t)c->P.blocksPerFrame * kScanChunk;
    A(devAlloc(c, &c->dMask, maskWords));
    const size_t flagWords = (size_t)c->capCand / 32 + 64;
    const size_t nBlockSums = std::max({B * (size_t)c->P.blocksPerFrame, (size_t)c->capCand / 256 + 2,
                                        (size_t)c->capKp / kScanChunk + 2}) + 16;
    for (int k = 0; k < 2; k++) {   // set 1 only ever holds one frame's deeper octaves
        A(devAlloc(c, &c->L[k].dBlockSums, nBlockSums));
        A(devAlloc(c, &c->L[k].dCands, (size_t)c->capCand));
        A(devAlloc(c, &c->L[k].dKpTmp, (size_t)c->capCand));
        A(devAlloc(c, &c->L[k].dFlagWords, flagWords));
        A(devAlloc(c, &c->L[k].dKps, (size_t)c->capKp));
        A(devAlloc(c, &c->L[k].dKpSeg, (size_t)c->capKp));
        A(devAlloc(c, &c->L[k].dSegStarts, 3 * (size_t)(c->nSegs + 1)));
        A(devAlloc(c, &c->L[k].dNOri, (size_t)c->capKp + kScanChunk));
        A(devAlloc(c, &c->L[k].dOriTmp, (size_t)c->capKp * kOriBins));
        A(devAlloc(c, &c->L[k].dOriOffset, (size_t)c->capKp + kScanChunk + 1));
        A(devAlloc(c, &c->L[k].dDesc, (size_t)c->capDesc));
        A(devAlloc(c, &c->L[k].dDescKp, (size_t)c->capDesc));
        A(devAlloc(c, &c->L[k].dCounters, 1));
    }
    if (e == cudaSuccess) A(cudaMemset(c->dMask, 0, maskWords * sizeof(uint32_t)));
    // row padding (columns w..pitch) is read by the extrema kernel's full-warp loads and masked
    // afterwards: give it defined contents once
    for (int o = 0; o < kOctaves && e == cudaSuccess; o++) {
        OctaveDev& q = c->P.oct[o];
        A(cudaMemset(q.G, 0, B * kGaussians * q.plane * sizeof(float)));
        A(cudaMemset(q.D, 0, B * kDogs * q.plane * sizeof(float)));
        A(cudaMemset(q.grad, 0, B * kScales * q.plane * sizeof(float2)));
    }
    if (e == cudaSuccess) A(cudaMemset(c->dScaled, 0, B * c->P.oct[0].plane * sizeof(float)));
    for (int k = 0; k < 2 && e == cudaSuccess; k++)
        A(cudaMemset(c->L[k].dSegStarts, 0, 3 * (size_t)(c->nSegs + 1) * sizeof(int)));
    for (int k = 0; k < 2; k++) {
        A(cudaMallocHost(&c->L[k].hCounters, sizeof(Counters)));
        A(cudaMallocHost(&c->L[k].hSegStarts, 3 * (size_t)(c->nSegs + 1) * sizeof(int)));
        if (e == cudaSuccess) {
            memset(c->L[k].hCounters, 0, sizeof(Counters));
            memset(c->L[k].hSegStarts, 0, 3 * (size_t)(c->nSegs + 1) * sizeof(int));
        }
    }
    for (auto& evn : c->evB) A(cudaEventCreate(&evn));
    A(cudaMallocHost(&c->hKps, std::max<size_t>((size_t)c->capKp * sizeof(SiftKeypoint), 64)));
    A(cudaMallocHost(&c->hDesc, std::max<size_t>((size_t)c->capDesc * sizeof(SiftDescriptor), 64)));
    A(cudaMallocHost(&c->hKpSeg, std::max<size_t>((size_t)c->capKp * sizeof(int), 64)));
    for (auto& evn : c->ev) A(cudaEventCreate(&evn));
    for (auto& evn : c->evBlur0) A(cudaEventCreate(&evn));
    int prioLow = 0, prioHigh = 0;
    cudaDeviceGetStreamPriorityRange(&prioLow, &prioHigh);
    c->octStream[0] = c->stream;
    A(cudaEventCreateWithFlags(&c->evBandFork, cudaEventDisableTiming));
    A(cudaStreamCreateWithFlags(&c->copyStream, cudaStreamNonBlocking));
    A(cudaEventCreateWithFlags(&c->evRefined, cudaEventDisableTiming));
    for (auto& evn : c->evUp) A(cudaEventCreateWithFlags(&evn, cudaEventDisableTiming));
    for (auto& evn : c->evSeedDone) A(cudaEventCreateWithFlags(&evn, cudaEventDisableTiming));
    for (auto& evn : c->evBandBlurEnd) A(cudaEventCreate(&evn));
    A(cudaEventCreateWithFlags(&c->evKpCopied, cudaEventDisableTiming));
    for (int b = 1; b < SiftContext::kMaxBands; b++) {
        A(cudaStreamCreateWithFlags(&c->bandStream[b], cudaStreamNonBlocking));
        A(cudaEventCreateWithFlags(&c->evBandSeeded[b], cudaEventDisableTiming));
        A(cudaEventCreateWithFlags(&c->evBandDone[b], cudaEventDisableTiming));
    }
    if (const char* nb = getenv("SIFTCUDA_BANDS")) c->nBands = std::max(1, std::min(SiftContext::kMaxBands, atoi(nb)));
    for (int o = 0; o < kOctaves; o++) {
        // smaller octaves form a long dependent chain of tiny launches: give their streams a higher
        // priority so that their CTAs are placed ahead of the big octave-0 kernels' when SM slots
        // free up (the chain is latency-critical, octave 0 is throughput-bound)
        if (o > 0) A(cudaStreamCreateWithPriority(&c->octStream[o], cudaStreamNonBlocking, std::max(prioHigh, prioLow - o)));
        A(cudaEventCreateWithFlags(&c->evSeeded[o], cudaEventDisableTiming));
        A(cudaEventCreateWithFlags(&c->evOctDone[o], cudaEventDisableTiming));

    }
    if (e != cudaSuccess) {
        const bool oom = (e == cudaErrorMemoryAllocation);
        destroy(c);
        cudaGetLastError();
        return oom ? SIFT_ERR_OUT_OF_MEMORY : SIFT_ERR_CUDA;
    }
    c->kpCounts.assign(c->nSegs, 0);
    c->descCounts.assign(c->nSegs, 0);
    c->candCounts.assign(c->nSegs, 0);
    I.device_bytes = c->deviceBytes;
    *out = c;
    return SIFT_OK;
}

void sift_destroy(SiftContext* c) { destroy(c); }

int sift_get_info(const SiftContext* c, SiftInfo* out) {
    if (!c || !out) return SIFT_ERR_INVALID_ARGUMENT;
    *out = c->info;
    return SIFT_OK;
}

int sift_set_stage_timing(SiftContext* c, int32_t enabled) {
    if (!c) return SIFT_ERR_INVALID_ARGUMENT;
    c->stageTiming = enabled != 0;
    return SIFT_OK;
}

int sift_last_timings(const SiftContext* c, SiftTimings* out) {
    if (!c || !out) return SIFT_ERR_INVALID_ARGUMENT;
    *out = c->timings;
    return SIFT_OK;
}

int sift_batch_upload(SiftContext* c, const void* const* images, int32_t n, int32_t pitchBytes) {
    if (!c || !images || n < 1 || n > c->B || pitchBytes < c->cfg.width * 4)
        return fail(c, SIFT_ERR_INVALID_ARGUMENT, "sift_batch_upload: bad arguments");
    CTX_TRY(c, cudaSetDevice(c->device));
    const size_t rowBytes = (size_t)c->cfg.width * 4;
    const size_t frameBytes = rowBytes * c->cfg.height;
    for (int f = 0; f < n; f++)
        if (!images[f]) return fail(c, SIFT_ERR_INVALID_ARGUMENT, "sift_batch_upload: null image");
    c->upSplit = seedSplitFor(c, n);
    if (c->upSplit.n > 0) {
        // one row chunk per band on the copy stream, an event behind each
        const uint8_t* src = (const uint8_t*)images[0];
        int g0 = 0;
        for (int k = 0; k < c->upSplit.n; k++) {
            const int g1 = c->upSplit.grayRows[k];
            CTX_TRY(c, cudaMemcpy2DAsync(c->dInput + (size_t)g0 * rowBytes, rowBytes, src + (size_t)g0 * pitchBytes,
                                         pitchBytes, rowBytes, g1 - g0, cudaMemcpyHostToDevice, c->copyStream));
            CTX_TRY(c, cudaEventRecord(c->evUp[k], c->copyStream));
            g0 = g1;
        }
    } else {
        for (int f = 0; f < n; f++)
            CTX_TRY(c, cudaMemcpy2DAsync(c->dInput + f * frameBytes, rowBytes, images[f], pitchBytes,
                                         rowBytes, c->cfg.height, cudaMemcpyHostToDevice, c->stream));
    }
    c->curInput = c->dInput;
    c->curPitch = (int)rowBytes;
    c->curFrameStride = (int64_t)frameBytes;
    c->curFrames = n;
    c->executed = false;
    return SIFT_OK;
}

int sift_batch_set_device_input(SiftContext* c, const void* dev, int32_t n, int32_t pitchBytes,
                                int64_t frameStrideBytes) {
    if (!c || !dev || n < 1 || n > c->B || pitchBytes < c->cfg.width * 4 ||
        (n > 1 && frameStrideBytes < (int64_t)pitchBytes * c->cfg.height) || (pitchBytes & 3) ||
        (frameStrideBytes & 3) || ((uintptr_t)dev & 3))
        return fail(c, SIFT_ERR_INVALID_ARGUMENT, "sift_batch_set_device_input: bad arguments");
    c->upSplit = SiftContext::SeedSplit{};
    c->curInput = (const uint8_t*)dev;
    c->curPitch = pitchBytes;
    c->curFrameStride = frameStrideBytes;
    c->curFrames = n;
    c->executed = false;
    return SIFT_OK;
}

}  // extern "C"

namespace {

// Mask blocks [blockBegin, blockBegin + nBlocks) → ordered candidates → refined keypoints of
// list set k (SIFTOctave.getKeypoints :198-203, interpolateKeypoints :205-288).
int postDetect(SiftContext* c, int k, int blockBegin, int nBlocks, int nSegs, cudaEvent_t afterCompaction,
               cudaEvent_t afterRefine) {
    SiftContext::ListSet& L = c->L[k];
    cudaStream_t st = c->stream;
    CTX_TRY(c, launchCandidateCompaction(c->P, c->dMask, L.dBlockSums, L.dCands, c->capCand, blockBegin,
                                         nBlocks, L.dSegStarts, nSegs, L.dCounters, st));
    c->launches += 3;
    if (afterCompaction) CTX_TRY(c, cudaEventRecord(afterCompaction, st));
    CTX_TRY(c, launchRefine(c->P, L.dCands, c->capCand, L.dKpTmp, L.dFlagWords, L.dBlockSums, L.dKps, L.dKpSeg,
                            c->capKp, L.dSegStarts, L.dSegStarts + (c->nSegs + 1), nSegs, L.dCounters, st));
    c->launches += 3;
    if (afterRefine) CTX_TRY(c, cudaEventRecord(afterRefine, st));
    return SIFT_OK;
}

// getDescriptors (SIFT.swift:207-238) over the keypoints of list set k.
int describeSet(SiftContext* c, int k, int nSegs, const int* kpIndexBase, cudaEvent_t afterOrientation,
                cudaEvent_t afterDescriptor) {
    SiftContext::ListSet& L = c->L[k];
    cudaStream_t st = c->stream;
    c->descOnHost = c->wantHostOut && !c->split;   // pinned memory is device-addressable (UVA)
    CTX_TRY(c, launchDescribe(c->P, L.dKps, L.dKpSeg, c->capKp, L.dSegStarts + (c->nSegs + 1), L.dNOri, L.dOriTmp,
                              L.dOriOffset, L.dDescKp, L.dBlockSums, c->descOnHost ? c->hDesc : L.dDesc, c->capDesc,
                              L.dSegStarts + 2 * (c->nSegs + 1), nSegs, L.dCounters, kpIndexBase, c->smCount, st,
                              afterOrientation));
    c->launches += 5;
    if (afterDescriptor) CTX_TRY(c, cudaEventRecord(afterDescriptor, st));
    return SIFT_OK;
}

// findKeypoints + getKeypointsFromOctaves + interpolateKeypoints (SIFT.swift:154-202), all
// frames of the batch at once, no host synchronisation inside.
int runDetect(SiftContext* c, bool withDescribe) {
    const int F = c->curFrames;
    cudaStream_t st = c->stream;
    const bool T = c->stageTiming;
    c->launches = 0;
    c->kpsOnHost = c->descOnHost = false;
    if (!c->countersClean) {   // normally zeroed behind the previous call's read-back (finish)
        CTX_TRY(c, cudaMemsetAsync(c->L[0].dCounters, 0, sizeof(Counters), st));
        CTX_TRY(c, cudaMemsetAsync(c->L[1].dCounters, 0, sizeof(Counters), st));
    }
    c->countersClean = false;
    if (T) CTX_TRY(c, cudaEventRecord(c->ev[0], st));
    // DifferenceOfGaussians.encodeSeedTexture (:357-389)
    const OctaveDev& o0 = c->P.oct[0];
    BlurArgs seed{};
    seed.in = c->dScaled;
    seed.out = o0.G;  // octave 0 slice 0 (the reference blits seed → slice 0, :176-188)
    seed.w = o0.w; seed.h = o0.h; seed.pitch = o0.pitch;
    seed.inFrameStride = o0.plane;
    seed.outFrameStride = kGaussians * o0.plane;
    seed.frames = F;
    const SiftContext::SeedSplit sp = (F == 1 && c->curInput == c->dInput) ? c->upSplit : SiftContext::SeedSplit{};
    if (sp.n > 0) {
        // chunk k → everything band k of octave 0 still lacks, on band k's stream (main stream for
        // band 0), behind chunk k's arrival and chunk k - 1's seed rows
        for (int k = 0; k < sp.n; k++) {
            cudaStream_t sk = k == 0 ? st : c->bandStream[k];
            CTX_TRY(c, cudaStreamWaitEvent(sk, c->evUp[k], 0));
            if (k > 0) CTX_TRY(c, cudaStreamWaitEvent(sk, c->evSeedDone[k - 1], 0));
            const int up0 = k ? sp.upRows[k - 1] : 0, seed0 = k ? sp.seedRows[k - 1] : 0;
            CTX_TRY(c, launchGrayUpsample(c->curInput, c->curPitch, c->curFrameStride, c->dGray, c->cfg.width,
                                          c->cfg.height, c->dScaled, o0.w, o0.h, o0.pitch, o0.plane, F, sk,
                                          k ? sp.grayRows[k - 1] : 0, sp.grayRows[k], up0, sp.upRows[k]));
            seed.yBegin = seed0; seed.yEnd = sp.seedRows[k];
            CTX_TRY(c, launchBlur(seed, c->seedTaps, c->seedNtaps, sk));
            CTX_TRY(c, cudaEventRecord(c->evSeedDone[k], sk));
            if (k > 0) c->launches += 2;
        }
    } else {
        CTX_TRY(c, launchGrayUpsample(c->curInput, c->curPitch, c->curFrameStride, c->dGray,
                                      c->cfg.width, c->cfg.height, c->dScaled, o0.w, o0.h, o0.pitch,
                                      o0.plane, F, st));
        CTX_TRY(c, launchBlur(seed, c->seedTaps, c->seedNtaps, st));
    }
    c->launches += 2;
    if (T) CTX_TRY(c, cudaEventRecord(c->ev[1], st));
    // encodeOctaves (:391-406): Gaussian series + DoG; octave o+1 slice 0 = octave o slice 3
    // decimated (fused into the blur that produces slice 3); SIFTOctave.encodeGradients (:190-196)
    // and encodeExtrema (:183-189). Octaves form a fork/join DAG: octave o+1 starts on its own
    // stream once octave o has written slice 3, so the small octaves (launch-latency bound) run
    // under the large kernels of octaves 0 and 1 instead of after them.
    bool forked[kOctaves] = {};
    static const int dbgMaxOctave = getenv("SIFTCUDA_DEBUG_MAX_OCTAVE") ? atoi(getenv("SIFTCUDA_DEBUG_MAX_OCTAVE")) : kOctaves;
    static const int dbgSkip = getenv("SIFTCUDA_DEBUG_SKIP") ? atoi(getenv("SIFTCUDA_DEBUG_SKIP")) : 0;  // 1 gradient, 2 extrema
    for (int o = 0; o < kOctaves; o++) {
        const OctaveDev& q = c->P.oct[o];
        if (q.w < 1 || q.h < 1 || o > dbgMaxOctave) continue;
        cudaStream_t so = c->octStream[o];
        if (o > 0) {
            CTX_TRY(c, cudaStreamWaitEvent(so, c->evSeeded[o - 1], 0));
            if (o == 1 && c->bandedOctave0)
                for (int b = 1; b < c->nBands; b++) CTX_TRY(c, cudaStreamWaitEvent(so, c->evBandSeeded[b], 0));
            forked[o] = true;
        }
        // two row bands for a plane with many tiles (octave 0 of a large single frame); batches
        // already have enough independent work per launch
        const int nb = c->nBands;
        const bool banded = (o == 0) && bandedOctave0(c, F);
        if (T && o == 0) CTX_TRY(c, cudaEventRecord(c->evBlur0[0], so));
        if (banded) {
            CTX_TRY(c, cudaEventRecord(c->evBandFork, so));
            for (int b = 1; b < nb; b++) CTX_TRY(c, cudaStreamWaitEvent(c->bandStream[b], c->evBandFork, 0));
        }
        for (int band = 0; band < (banded ? nb : 1); band++) {
            cudaStream_t sb = band == 0 ? so : c->bandStream[band];
            // band rows [r0, r1), boundaries on multiples of the tile height
            const int r0 = banded ? bandBoundary(c, band) : 0;
            const int r1 = banded ? bandBoundary(c, band + 1) : q.h;
            for (int s = 0; s < kGaussians - 1; s++) {
                if (T && o == 0 && !banded && s > 0) CTX_TRY(c, cudaEventRecord(c->evBlur0[s], sb));
                BlurArgs a{};
                a.in = q.G + (size_t)s * q.plane;
                a.out = q.G + (size_t)(s + 1) * q.plane;
                a.dog = q.D + (size_t)s * q.plane;
                a.w = q.w; a.h = q.h; a.pitch = q.pitch;
                a.inFrameStride = a.outFrameStride = kGaussians * q.plane;
                a.dogFrameStride = kDogs * q.plane;
                if (banded) {
                    // rows the later scales of this band still need: sum of their radii
                    int halo = bandTails() ? 1 : 0;
                    for (int t = s + 1; t < kGaussians - 1; t++) halo += c->ntaps[t] / 2;
                    a.yBegin = std::max(0, r0 - halo);
                    a.yEnd = std::min(q.h, r1 + halo);
                }
                if (s + 1 == kScales && o + 1 < kOctaves && c->P.oct[o + 1].w >= 1 && c->P.oct[o + 1].h >= 1) {
                    const OctaveDev& nx = c->P.oct[o + 1];
                    a.half = nx.G;
                    a.halfW = nx.w; a.halfH = nx.h; a.halfPitch = nx.pitch;
                    a.halfFrameStride = kGaussians * nx.plane;
                }
                a.frames = F;
                // small planes: scale s + 1 launches under scale s (programmatic dependent launch)
                static const int pdlMaxTiles = getenv("SIFTCUDA_PDL_TILES") ? atoi(getenv("SIFTCUDA_PDL_TILES")) : 160;
                a.pdl = (s > 0 && !banded && (long)((q.w + 31) / 32) * ((q.h + 31) / 32) * F <= pdlMaxTiles) ? 1 : 0;
                CTX_TRY(c, launchBlur(a, c->taps[s], c->ntaps[s], sb));
                c->launches++;
                if (s + 1 == kScales) CTX_TRY(c, cudaEventRecord(band == 0 ? c->evSeeded[o] : c->evBandSeeded[band], sb));
            }
            if (banded && bandTails()) {
                if (T) CTX_TRY(c, cudaEventRecord(c->evBandBlurEnd[band], sb));
                if (!(dbgSkip & 1)) CTX_TRY(c, launchGradient(q, F, sb, r0, r1));
                if (!(dbgSkip & 2)) CTX_TRY(c, launchExtremaMask(c->P, o, c->dMask, F, sb, r0, r1));
                c->launches += 2;
            }
        }
        if (banded) {   // join: (without band tails) gradient and extrema need every row
            for (int b = 1; b < nb; b++) {
                CTX_TRY(c, cudaEventRecord(c->evBandDone[b], c->bandStream[b]));
                CTX_TRY(c, cudaStreamWaitEvent(so, c->evBandDone[b], 0));
            }
        }
        if (o == 0) c->bandedOctave0 = banded;
        if (T && o == 0) CTX_TRY(c, cudaEventRecord(c->evBlur0[kGaussians - 1], so));
        // (the gradient beside the extrema mask on a second stream, or beside blurs s = 3, 4: both
        // measured, no gain — the stage is throughput-bound)
        if (!(banded && bandTails())) {
            if (!(dbgSkip & 1)) CTX_TRY(c, launchGradient(q, F, so));
            c->launches++;
            if (q.w >= 3 && q.h >= 3 && !(dbgSkip & 2)) {
                CTX_TRY(c, launchExtremaMask(c->P, o, c->dMask, F, so));
                c->launches++;
            }
        }
        if (o > 0) CTX_TRY(c, cudaEventRecord(c->evOctDone[o], so));
        if (o == 0) {
            // Single large frame: octave 0 holds most of the keypoints and is complete long before
            // the chain of deeper octaves. Compact / refine / describe it now from list set 0; the
            // (throughput-bound) orientation and descriptor kernels then run over the
            // (latency-bound) tail of the small octaves, which follow from set 1 after the join.
            // Measured at 1080p: with prioritised octave streams octave 0 is itself the critical
            // path (0.47 ms), so the split only adds a second set of scan launches (788 vs 846
            // frames/s). Off unless SIFTCUDA_SPLIT=1.
            static const bool wantSplit = getenv("SIFTCUDA_SPLIT") && atoi(getenv("SIFTCUDA_SPLIT")) != 0;
            c->split = wantSplit && c->bandedOctave0 && c->P.oct[1].maskBlockStart > 0;
            c->candSplit = c->split;
            if (c->split) {
                if (T) CTX_TRY(c, cudaEventRecord(c->ev[2], st));
                int r = postDetect(c, 0, 0, c->P.oct[1].maskBlockStart, kOctaves, T ? c->ev[3] : nullptr,
                                   T ? c->ev[4] : nullptr);
                if (r != SIFT_OK) return r;
                if (withDescribe) {
                    r = describeSet(c, 0, kOctaves, nullptr, T ? c->ev[5] : nullptr, T ? c->ev[6] : nullptr);
                    if (r != SIFT_OK) return r;
                }
            }
        }
    }
    for (int o = 1; o < kOctaves; o++)
        if (forked[o]) CTX_TRY(c, cudaStreamWaitEvent(st, c->evOctDone[o], 0));
    const int nSegs = F * kOctaves;
    if (c->split) {
        if (T) CTX_TRY(c, cudaEventRecord(c->evB[0], st));
        const int b1 = c->P.oct[1].maskBlockStart;
        int r = postDetect(c, 1, b1, c->P.blocksPerFrame - b1, kOctaves, T ? c->evB[1] : nullptr,
                           T ? c->evB[2] : nullptr);
        if (r != SIFT_OK) return r;
        if (withDescribe) {
            r = describeSet(c, 1, kOctaves, &c->L[0].dCounters->nKeypoints, T ? c->evB[3] : nullptr,
                            T ? c->evB[4] : nullptr);
            if (r != SIFT_OK) return r;
        }
    } else {
        if (T) CTX_TRY(c, cudaEventRecord(c->ev[2], st));
        int r = postDetect(c, 0, 0, c->P.blocksPerFrame * F, nSegs, T ? c->ev[3] : nullptr, T ? c->ev[4] : nullptr);
        if (r != SIFT_OK) return r;
        if (c->wantHostOut) {
            // keypoint count for the early copy (earlyKeypointCopy), read back on the copy stream so
            // that the main stream goes straight on to the orientation kernel
            CTX_TRY(c, cudaEventRecord(c->evKpCopied, st));
            CTX_TRY(c, cudaStreamWaitEvent(c->copyStream, c->evKpCopied, 0));
            CTX_TRY(c, cudaMemcpyAsync(c->L[0].hCounters, c->L[0].dCounters, sizeof(Counters), cudaMemcpyDeviceToHost,
                                       c->copyStream));
            CTX_TRY(c, cudaEventRecord(c->evRefined, c->copyStream));
            c->kpsOnHost = true;
        }
        if (withDescribe) {
            r = describeSet(c, 0, nSegs, nullptr, T ? c->ev[5] : nullptr, T ? c->ev[6] : nullptr);
            if (r != SIFT_OK) return r;
        }
    }
    c->executed = true;
    c->described = withDescribe;
    return SIFT_OK;
}

// Host-output mode: while the orientation and descriptor kernels (already queued) run, wait for
// the refined-keypoint count and send the keypoints home on the copy stream.
int earlyKeypointCopy(SiftContext* c) {
    if (!c->kpsOnHost) return SIFT_OK;
    CTX_TRY(c, cudaEventSynchronize(c->evRefined));
    const int64_t nk = std::min(c->L[0].hCounters->nKeypoints, c->capKp);
    if (nk > 0) {
        CTX_TRY(c, cudaMemcpyAsync(c->hKps, c->L[0].dKps, (size_t)nk * sizeof(SiftKeypoint), cudaMemcpyDeviceToHost,
                                   c->copyStream));
    }
    CTX_TRY(c, cudaEventRecord(c->evKpCopied, c->copyStream));
    return SIFT_OK;
}

// D2H of the small bookkeeping blocks, stream drain, timing read-out, overflow check.
int finish(SiftContext* c, bool withDescribe) {
    cudaStream_t st = c->stream;
    const int nSets = c->split ? 2 : 1;
    const size_t segInts = 3 * (size_t)(c->nSegs + 1);
    for (int k = 0; k < nSets; k++) {
        CTX_TRY(c, cudaMemcpyAsync(c->L[k].hCounters, c->L[k].dCounters, sizeof(Counters), cudaMemcpyDeviceToHost, st));
        CTX_TRY(c, cudaMemcpyAsync(c->L[k].hSegStarts, c->L[k].dSegStarts, segInts * sizeof(int),
                                   cudaMemcpyDeviceToHost, st));
    }
    CTX_TRY(c, cudaStreamSynchronize(st));
    if (c->kpsOnHost) CTX_TRY(c, cudaEventSynchronize(c->evKpCopied));
    // zero the counters for the next frame now, off its critical path (overflow bits accumulate by
    // atomicOr; the counts themselves are rewritten by every scan)
    if (cudaMemsetAsync(c->L[0].dCounters, 0, sizeof(Counters), st) == cudaSuccess &&
        cudaMemsetAsync(c->L[1].dCounters, 0, sizeof(Counters), st) == cudaSuccess)
        c->countersClean = true;
    const int nSegs = c->curFrames * kOctaves;
    int overflow = 0;
    for (int s = 0; s < nSegs; s++) c->candCounts[s] = c->kpCounts[s] = c->descCounts[s] = 0;
    for (int k = 0; k < nSets; k++) {
        const int* candStart = c->L[k].hSegStarts;
        const int* kpStart = c->L[k].hSegStarts + (c->nSegs + 1);
        const int* descStart = c->L[k].hSegStarts + 2 * (c->nSegs + 1);
        const int nDesc = c->L[k].hCounters->nDescriptors;
        for (int s = 0; s < nSegs; s++) {
            c->candCounts[s] += candStart[s + 1] - candStart[s];
            c->kpCounts[s] += kpStart[s + 1] - kpStart[s];
            if (withDescribe) c->descCounts[s] += std::min(descStart[s + 1], nDesc) - std::min(descStart[s], nDesc);
        }
        overflow |= c->L[k].hCounters->overflow;
    }
    SiftTimings& t = c->timings;
    memset(&t, 0, sizeof t);
    t.kernel_launches = c->launches;
    t.stage_timing_enabled = c->stageTiming;
    if (c->stageTiming) {
        const int last = withDescribe ? SIFT_STAGE_COUNT : 4;
        for (int i = 0; i < last; i++) cudaEventElapsedTime(&t.stage_ms[i], c->ev[i], c->ev[i + 1]);
        cudaEventElapsedTime(&t.total_ms, c->ev[0], c->ev[last]);
        if (c->split) {
            // second pass (octaves >= 1): its stages are added to the first pass's; the wait for
            // the deeper octaves between the passes, if any, counts as pyramid time
            const int lastB = withDescribe ? 4 : 2;
            float ms = 0;
            cudaEventElapsedTime(&ms, c->ev[last], c->evB[0]);
            t.stage_ms[SIFT_STAGE_PYRAMID] += ms;
            for (int i = 0; i < lastB; i++) {
                cudaEventElapsedTime(&ms, c->evB[i], c->evB[i + 1]);
                t.stage_ms[SIFT_STAGE_EXTREMA + i] += ms;
            }
            cudaEventElapsedTime(&t.total_ms, c->ev[0], c->evB[lastB]);
        }
        cudaEventElapsedTime(&t.blur_octave0_ms, c->evBlur0[0], c->evBlur0[kGaussians - 1]);
        if (c->bandedOctave0 && bandTails()) {   // the section ends with the last band's last blur
            t.blur_octave0_ms = 0;
            for (int b = 0; b < c->nBands; b++) {
                float ms = 0;
                cudaEventElapsedTime(&ms, c->evBlur0[0], c->evBandBlurEnd[b]);
                t.blur_octave0_ms = std::max(t.blur_octave0_ms, ms);
            }
        }
        for (int s = 0; s < kGaussians - 1; s++)   // per-scale split only without row bands
            t.blur_octave0_launch_ms[s] = c->bandedOctave0 ? t.blur_octave0_ms / (kGaussians - 1) : 0.0f;
        if (!c->bandedOctave0)
            for (int s = 0; s < kGaussians - 1; s++)
                cudaEventElapsedTime(&t.blur_octave0_launch_ms[s], c->evBlur0[s], c->evBlur0[s + 1]);
        t.blur_octave0_launches = kGaussians - 1;
        cudaGetLastError();   // an unrecorded timing event must not poison the next launch check
    }
    if (overflow) {
        char buf[160];
        snprintf(buf, sizeof buf, "list capacity exceeded (mask %d: 1 candidates, 2 keypoints, 4 descriptors)",
                 overflow);
        return fail(c, SIFT_ERR_CAPACITY, buf);
    }
    return SIFT_OK;
}

}  // namespace

extern "C" {

int sift_batch_execute(SiftContext* c) {
    if (!c) return SIFT_ERR_INVALID_ARGUMENT;
    if (!c->curInput || c->curFrames < 1) return fail(c, SIFT_ERR_INVALID_ARGUMENT, "no input set");
    CTX_TRY(c, cudaSetDevice(c->device));
    c->wantHostOut = false;   // staged path: results stay in HBM until sift_batch_download
    const int r = runDetect(c, true);
    if (r != SIFT_OK) return r;
    return finish(c, true);
}

int sift_batch_download(SiftContext* c, SiftBatchResult* out) {
    if (!c || !out) return SIFT_ERR_INVALID_ARGUMENT;
    if (!c->executed) return fail(c, SIFT_ERR_NOT_DETECTED, "download before execute");
    CTX_TRY(c, cudaSetDevice(c->device));
    // list sets are concatenated in order: set 0, then (split mode) set 1 = the deeper octaves
    int64_t nKp = 0, nDesc = 0;
    for (int k = 0; k < (c->split ? 2 : 1); k++) {
        const int64_t nk = std::min(c->L[k].hCounters->nKeypoints, c->capKp);
        const int64_t nd = c->described ? std::min(c->L[k].hCounters->nDescriptors, c->capDesc) : 0;
        if (nKp + nk > c->capKp || nDesc + nd > c->capDesc) return fail(c, SIFT_ERR_CAPACITY, "result arrays too small");
        if (nk > 0 && !c->kpsOnHost)
            CTX_TRY(c, cudaMemcpyAsync(c->hKps + nKp, c->L[k].dKps, (size_t)nk * sizeof(SiftKeypoint),
                                       cudaMemcpyDeviceToHost, c->stream));
        if (nd > 0 && !c->descOnHost)
            CTX_TRY(c, cudaMemcpyAsync(c->hDesc + nDesc, c->L[k].dDesc, (size_t)nd * sizeof(SiftDescriptor),
                                       cudaMemcpyDeviceToHost, c->stream));
        nKp += nk;
        nDesc += nd;
    }
    CTX_TRY(c, cudaStreamSynchronize(c->stream));
    out->n_frames = c->curFrames;
    out->keypoint_counts = c->kpCounts.data();
    out->descriptor_counts = c->descCounts.data();
    out->candidate_counts = c->candCounts.data();
    out->keypoints = c->hKps;
    out->descriptors = c->hDesc;
    out->total_keypoints = nKp;
    out->total_descriptors = nDesc;
    return SIFT_OK;
}

int sift_detect_and_describe_batch(SiftContext* c, const void* const* images, int32_t n,
                                   int32_t pitchBytes, SiftBatchResult* out) {
    if (!out) return SIFT_ERR_INVALID_ARGUMENT;
    int r = sift_batch_upload(c, images, n, pitchBytes);
    if (r != SIFT_OK) return r;
    static const bool hostOut = !(getenv("SIFTCUDA_HOST_OUT") && atoi(getenv("SIFTCUDA_HOST_OUT")) == 0);
    c->wantHostOut = hostOut;
    r = runDetect(c, true);
    c->wantHostOut = false;
    if (r != SIFT_OK) return r;
    r = earlyKeypointCopy(c);
    if (r != SIFT_OK) return r;
    const int re = finish(c, true);
    if (re != SIFT_OK && re != SIFT_ERR_CAPACITY) return re;
    r = sift_batch_download(c, out);
    return r != SIFT_OK ? r : re;
}

int sift_detect(SiftContext* c, const void* bgra8, int32_t pitchBytes,
                const SiftKeypoint** outKps, int32_t counts[SIFT_NUM_OCTAVES]) {
    if (!c || !bgra8 || !outKps || !counts) return SIFT_ERR_INVALID_ARGUMENT;
    const void* imgs[1] = {bgra8};
    int r = sift_batch_upload(c, imgs, 1, pitchBytes);
    if (r != SIFT_OK) return r;
    c->wantHostOut = false;
    r = runDetect(c, false);
    if (r != SIFT_OK) return r;
    const int re = finish(c, false);
    if (re != SIFT_OK && re != SIFT_ERR_CAPACITY) return re;
    SiftBatchResult res;
    r = sift_batch_download(c, &res);
    if (r != SIFT_OK) return r;
    *outKps = res.keypoints;
    for (int o = 0; o < kOctaves; o++) counts[o] = res.keypoint_counts[o];
    return re;
}

int sift_describe(SiftContext* c, const SiftKeypoint* kps, const int32_t counts[SIFT_NUM_OCTAVES],
                  const SiftDescriptor** outDesc, int32_t descCounts[SIFT_NUM_OCTAVES]) {
    if (!c || !counts || !outDesc || !descCounts) return SIFT_ERR_INVALID_ARGUMENT;
    if (!c->executed) return fail(c, SIFT_ERR_NOT_DETECTED, "sift_describe before sift_detect");
    CTX_TRY(c, cudaSetDevice(c->device));
    int64_t n = 0;
    for (int o = 0; o < kOctaves; o++) {
        if (counts[o] < 0) return fail(c, SIFT_ERR_INVALID_ARGUMENT, "negative count");
        n += counts[o];
    }
    if (n > c->capKp) return fail(c, SIFT_ERR_CAPACITY, "more keypoints than max_keypoints_per_frame");
    if (n > 0 && !kps) return SIFT_ERR_INVALID_ARGUMENT;
    // keypoints may alias our own pinned result buffer (the usual detect → describe flow)
    if (n > 0 && kps != c->hKps) memcpy(c->hKps, kps, (size_t)n * sizeof(SiftKeypoint));
    int* kpStart = c->L[0].hSegStarts + (c->nSegs + 1);
    int k = 0;
    for (int o = 0; o < kOctaves; o++) {
        kpStart[o] = k;
        for (int i = 0; i < counts[o]; i++) {
            if (c->hKps[k].octave != o) return fail(c, SIFT_ERR_INVALID_ARGUMENT, "keypoint octave does not match its group");
            c->hKpSeg[k++] = o;
        }
    }
    for (int s = kOctaves; s <= c->nSegs; s++) kpStart[s] = k;
    cudaStream_t st = c->stream;
    memset(c->L[0].hCounters, 0, sizeof(Counters));   // incl. the kernels' work-queue counters
    c->L[0].hCounters->nKeypoints = (int)n;
    CTX_TRY(c, cudaMemcpyAsync(c->L[0].dCounters, c->L[0].hCounters, sizeof(Counters), cudaMemcpyHostToDevice, st));
    if (n > 0) {
        CTX_TRY(c, cudaMemcpyAsync(c->L[0].dKps, c->hKps, (size_t)n * sizeof(SiftKeypoint), cudaMemcpyHostToDevice, st));
        CTX_TRY(c, cudaMemcpyAsync(c->L[0].dKpSeg, c->hKpSeg, (size_t)n * sizeof(int), cudaMemcpyHostToDevice, st));
    }
    CTX_TRY(c, cudaMemcpyAsync(c->L[0].dSegStarts + (c->nSegs + 1), kpStart, (size_t)(c->nSegs + 1) * sizeof(int),
                               cudaMemcpyHostToDevice, st));
    const bool T = c->stageTiming;
    c->launches = 0;
    if (T) CTX_TRY(c, cudaEventRecord(c->ev[4], st));
    const int savedFrames = c->curFrames;
    c->curFrames = 1;
    c->split = false;   // caller-supplied keypoints all live in list set 0
    c->kpsOnHost = true;   // they are the caller's (already in hKps); nothing to copy back
    c->wantHostOut = true;
    int r = describeSet(c, 0, kOctaves, nullptr, T ? c->ev[5] : nullptr, T ? c->ev[6] : nullptr);
    c->wantHostOut = false;
    c->described = true;
    if (r != SIFT_OK) { c->curFrames = savedFrames; return r; }
    CTX_TRY(c, cudaEventRecord(c->evKpCopied, c->copyStream));   // finish() waits on it
    // bookkeeping without touching the detect-stage events
    const bool savedT = c->stageTiming;
    c->stageTiming = false;
    const int re = finish(c, true);
    c->stageTiming = savedT;
    if (T) {
        cudaEventElapsedTime(&c->timings.stage_ms[4], c->ev[4], c->ev[5]);
        cudaEventElapsedTime(&c->timings.stage_ms[5], c->ev[5], c->ev[6]);
        cudaEventElapsedTime(&c->timings.total_ms, c->ev[4], c->ev[6]);
    }
    if (re != SIFT_OK && re != SIFT_ERR_CAPACITY) { c->curFrames = savedFrames; return re; }
    SiftBatchResult res;
    r = sift_batch_download(c, &res);
    c->curFrames = savedFrames;
    if (r != SIFT_OK) return r;
    *outDesc = res.descriptors;
    for (int o = 0; o < kOctaves; o++) descCounts[o] = res.descriptor_counts[o];
    return re;
}

int sift_debug_download(SiftContext* c, int32_t what, int32_t frame, int32_t octave, int32_t slice,
                        float* dst, int64_t dstFloats) {
    if (!c || !dst) return SIFT_ERR_INVALID_ARGUMENT;
    if (!c->executed) return fail(c, SIFT_ERR_NOT_DETECTED, "debug download before execute");
    if (frame < 0 || frame >= c->curFrames) return fail(c, SIFT_ERR_INVALID_ARGUMENT, "bad frame");
    CTX_TRY(c, cudaSetDevice(c->device));
    CTX_TRY(c, cudaStreamSynchronize(c->stream));
    const float* src = nullptr;
    int w = 0, h = 0, pitch = 0, comps = 1;
    if (what == SIFT_PLANE_GRAY) {
        w = pitch = c->cfg.width; h = c->cfg.height;
        src = c->dGray + (size_t)frame * w * h;
    } else {
        if (what == SIFT_PLANE_SEED) { octave = 0; slice = 0; what = SIFT_PLANE_GAUSSIAN; }
        if (octave < 0 || octave >= kOctaves) return fail(c, SIFT_ERR_INVALID_ARGUMENT, "bad octave");
        const OctaveDev& q = c->P.oct[octave];
        w = q.w; h = q.h; pitch = q.pitch;
        if (what == SIFT_PLANE_GAUSSIAN && slice >= 0 && slice < kGaussians)
            src = q.G + ((size_t)frame * kGaussians + slice) * q.plane;
        else if (what == SIFT_PLANE_DOG && slice >= 0 && slice < kDogs)
            src = q.D + ((size_t)frame * kDogs + slice) * q.plane;
        else if (what == SIFT_PLANE_GRADIENT && slice >= 1 && slice <= kScales) {
            src = (const float*)(q.grad + ((size_t)frame * kScales + (slice - 1)) * q.plane);
            comps = 2;
        } else
            return fail(c, SIFT_ERR_INVALID_ARGUMENT, "bad plane / slice");
    }
    if (dstFloats < (int64_t)w * h * comps) return fail(c, SIFT_ERR_INVALID_ARGUMENT, "dst too small");
    if (w > 0 && h > 0)
        CTX_TRY(c, cudaMemcpy2D(dst, (size_t)w * comps * 4, src, (size_t)pitch * comps * 4,
                                (size_t)w * comps * 4, h, cudaMemcpyDeviceToHost));
    return SIFT_OK;
}

int64_t sift_debug_candidates(SiftContext* c, int32_t frame, int32_t octave, int32_t* dst,
                              int64_t capTriples) {
    if (!c || !c->executed || frame < 0 || frame >= c->curFrames || octave < 0 || octave >= kOctaves)
        return -(int64_t)SIFT_ERR_INVALID_ARGUMENT;
    if (cudaSetDevice(c->device) != cudaSuccess) return -(int64_t)SIFT_ERR_CUDA;
    const int seg = frame * kOctaves + octave;
    const SiftContext::ListSet& L = c->L[(c->candSplit && octave >= 1) ? 1 : 0];
    const int nAll = std::min(L.hCounters->nCandidates, c->capCand);
    const int a = std::min(L.hSegStarts[seg], nAll), b = std::min(L.hSegStarts[seg + 1], nAll);
    const int n = b - a;
    if (!dst || n <= 0) return n;
    std::vector<Candidate> tmp((size_t)n);
    if (cudaMemcpy(tmp.data(), L.dCands + a, (size_t)n * sizeof(Candidate), cudaMemcpyDeviceToHost) != cudaSuccess)
        return -(int64_t)SIFT_ERR_CUDA;
    for (int i = 0; i < n && i < capTriples; i++) {
        dst[3 * i + 0] = (int32_t)(tmp[i].xys & 0x7fff);
        dst[3 * i + 1] = (int32_t)((tmp[i].xys >> 15) & 0x7fff);
        dst[3 * i + 2] = (int32_t)(tmp[i].xys >> 30);
    }
    return n;
}

int sift_debug_blur_bench(SiftContext* c, int32_t scale, int32_t mode, int32_t iters, float* outMs) {
    if (!c || !outMs || scale < 0 || scale >= kGaussians - 1 || iters < 1) return SIFT_ERR_INVALID_ARGUMENT;
    if (!c->executed) return fail(c, SIFT_ERR_NOT_DETECTED, "blur bench before execute");
    CTX_TRY(c, cudaSetDevice(c->device));
    const OctaveDev& q = c->P.oct[0];
    BlurArgs a{};
    a.in = q.G + (size_t)scale * q.plane;
    a.out = q.G + (size_t)(scale + 1) * q.plane;
    a.dog = q.D + (size_t)scale * q.plane;
    a.w = q.w; a.h = q.h; a.pitch = q.pitch;
    a.inFrameStride = a.outFrameStride = kGaussians * q.plane;
    a.dogFrameStride = kDogs * q.plane;
    a.frames = c->curFrames;
    a.debugMode = mode;
    const bool dual = (mode & 8) != 0;   // tuning: the same launches split over two streams
    a.debugMode = mode & 7;
    cudaStream_t s2 = c->octStream[1];   // any second stream
    CTX_TRY(c, launchBlur(a, c->taps[scale], c->ntaps[scale], c->stream));
    CTX_TRY(c, cudaStreamSynchronize(c->stream));
    CTX_TRY(c, cudaEventRecord(c->evBlur0[0], c->stream));
    if (dual) {
        CTX_TRY(c, cudaStreamWaitEvent(s2, c->evBlur0[0], 0));
        for (int i = 0; i < iters; i++)
            CTX_TRY(c, launchBlur(a, c->taps[scale], c->ntaps[scale], (i & 1) ? s2 : c->stream));
        CTX_TRY(c, cudaEventRecord(c->evSeeded[0], s2));
        CTX_TRY(c, cudaStreamWaitEvent(c->stream, c->evSeeded[0], 0));
    } else {
        for (int i = 0; i < iters; i++) CTX_TRY(c, launchBlur(a, c->taps[scale], c->ntaps[scale], c->stream));
    }
    CTX_TRY(c, cudaEventRecord(c->evBlur0[1], c->stream));
    CTX_TRY(c, cudaStreamSynchronize(c->stream));
    float ms = 0;
    CTX_TRY(c, cudaEventElapsedTime(&ms, c->evBlur0[0], c->evBlur0[1]));
    *outMs = ms / iters;
    c->executed = false;   // planes were overwritten in debug modes: force a fresh execute
    return SIFT_OK;
}

int sift_debug_math(int device, int32_t op, const float* a, const float* b, float* out, int64_t n) {
    if (!a || !out || n < 0) return SIFT_ERR_INVALID_ARGUMENT;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || device < 0 || device >= count) return SIFT_ERR_NO_DEVICE;
    if (cudaSetDevice(device) != cudaSuccess) return SIFT_ERR_NO_DEVICE;
    if (n == 0) return SIFT_OK;
    float *da = nullptr, *db = nullptr, *dout = nullptr;
    cudaError_t e = cudaMalloc(&da, n * sizeof(float));
    if (e == cudaSuccess) e = cudaMalloc(&db, n * sizeof(float));
    if (e == cudaSuccess) e = cudaMalloc(&dout, n * sizeof(float));
    if (e == cudaSuccess) e = cudaMemcpy(da, a, n * sizeof(float), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = b ? cudaMemcpy(db, b, n * sizeof(float), cudaMemcpyHostToDevice)
                                : cudaMemset(db, 0, n * sizeof(float));
    if (e == cudaSuccess) e = launchMathDebug(op, da, db, dout, n, 0);
    if (e == cudaSuccess) e = cudaMemcpy(out, dout, n * sizeof(float), cudaMemcpyDeviceToHost);
    cudaFree(da); cudaFree(db); cudaFree(dout);
    return e == cudaSuccess ? SIFT_OK : SIFT_ERR_CUDA;
}

}  // extern "C"
