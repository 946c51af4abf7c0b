// capi.cu — context lifetime, per-call orchestration and the extern "C" surface of
// include/siftcuda.h. Host stages restated here (each cites the reference):
//   DifferenceOfGaussians.init / Octave.init   (DifferenceOfGaussians.swift:233-344, :69-147)
//   GaussianKernel / GaussianSeriesKernel taps (GaussianKernel.swift:20-43, GaussianSeriesKernel.swift:27-51)
//   SIFT.getKeypoints / getDescriptors         (SIFT.swift:147-238)
// One context = one device, one compute stream (+ forked octave / band streams) and one copy
// stream. A call's whole device work — ~70 kernels on a fork / join DAG of streams — can be
// recorded once per (slot, batch size, input) as a CUDA graph and replayed with a single launch,
// the way the reference encodes its 102 dispatches into one command buffer (SIFT.swift:157-172);
// measured on B200 the replay costs the host ~10x less but runs ~4 % longer on the device than
// the same launches with programmatic dependent launch and prioritised streams, so it is opt-in
// (sift_set_graph_replay) for hosts that drive many GPUs. Results
// leave the device through the kernels' own stores into pinned host columns, so nothing but a
// 24-byte counter block and the segment starts is copied after the last kernel. Two in-flight
// slots (input arena + result columns each) let the upload of call i + 1 run under the kernels of
// call i. There is no CPU path: without a usable sm_100 device every compute entry point fails.
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <new>
#include <string>
#include <vector>

#include <nvtx3/nvToolsExt.h>

#include "common.cuh"

using namespace sift;

namespace {

// NVTX range named like the reference's measure(name:) sites (Utilities/Performance.swift:12-20;
// SIFT.swift:155,179,192,212,226). Host-side: the range covers the enqueue of that stage.
struct NvtxRange {
    explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
};

struct ResultColumns {
    KeypointColumnsDev kp{};
    DescriptorColumnsDev desc{};
};

// One in-flight call: its input arena, its pinned result columns and bookkeeping.
struct Slot {
    uint8_t* dInput = nullptr;          // device input arena (max_batch frames)
    bool allocated = false;
    ResultColumns host{};               // pinned host memory, device-addressable (UVA)
    ResultColumns dev{};                // the slot's result columns in HBM (slot 0: the context's)
    void* hostBlock = nullptr;          // start of the block the columns are carved from
    bool hostBlockExternal = false;     // caller memory (sift_bind_result_memory), not ours to free
    Counters* hCounters = nullptr;      // pinned
    int* hSegStarts = nullptr;          // pinned: [3][nSegs + 1] candidates, keypoints, descriptors
    cudaEvent_t evUploaded = nullptr, evStart = nullptr, evEnd = nullptr;
    int frames = 0;
    bool pending = false, described = false, hostOut = false, graphReplay = false, staged = false;
    bool copyOut = false;               // the columns reach host memory by the copy engine at sift_wait
    int launches = 0;
    std::vector<int32_t> kpCounts, descCounts, candCounts;
    int64_t nKp = 0, nDesc = 0;
    int status = SIFT_OK;
};

struct GraphEntry {
    int slot = 0, frames = 0, pitch = 0;
    const void* input = nullptr;
    int64_t frameStride = 0;
    bool describe = false, hostOut = false, copyOut = false;
    int seen = 0;                       // calls with this key so far (the first one runs eagerly)
    cudaGraphExec_t exec = nullptr;
    int launches = 0;
    uint64_t lastUse = 0;
};

}  // namespace

struct SiftContext {
    SiftConfig cfg{};
    int device = 0;
    int smCount = 0;
    int bytesPerPixel = 4;
    cudaStream_t stream = nullptr;
    cudaStream_t copyStream = nullptr;
    // Result delivery. Default: the scatter / descriptor kernels store the columns straight into
    // the slot's host memory (costs a 1080p call 24 µs against leaving them in HBM). Alternative
    // (SIFTCUDA_RESULT_COPY, tuning): the kernels write the slot's HBM columns and sift_wait moves
    // exactly the filled part with the copy engine on this stream, under the kernels of the next
    // submitted call — measured level with the direct stores launch by launch (0.909 ms per call
    // either way) and 11 µs ahead under graph replay; a copy-out kernel instead of the copy engine
    // was slower than both (SM stores at PCIe rate back up into the next call's kernels). See
    // profiles/r2/EXPERIMENTS.md.
    cudaStream_t downStream = nullptr;
    int resultCopy = 0;            // SIFTCUDA_RESULT_COPY (tuning): 0 never, 1 pipelined calls, 2 every host call
    // octave o >= 1 runs its blur chain + gradient + extrema mask on its own stream as soon as
    // octave o-1 has produced Gaussian slice 3 (fork / join around the main stream)
    cudaStream_t octStream[kOctaves]{};
    int octPriority[kOctaves]{};

    // Large octaves are split into row bands that run as independent chains of blur launches
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
    BlurTmaSet seedTma{};              // tensor maps of the upsampled plane (seed blur input)
    BlurTmaSet octTma[kOctaves]{};     // ... and of every octave's Gaussian stack
    CUtensorMap dogTma[kOctaves]{};    // DoG stacks (extrema mask kernel of large planes)
    bool dogTmaValid[kOctaves]{};

    int B = 1;       // max_batch
    int nSegs = 0;   // B * 7
    int capCand = 0, capKp = 0, capDesc = 0;  // totals over the batch

    // device memory
    std::vector<void*> allocations;
    int64_t deviceBytes = 0;
    float* dGray = nullptr;
    float* dScaled = nullptr;
    uint32_t* dMask = nullptr;
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
    Counters* dCounters = nullptr;
    ResultColumns dev{};        // device result columns (staged path, device-side matching)

    Slot slot[2];
    // Pipelined calls alternate between two lanes. Small contexts (twinMode) give lane 1 a whole
    // second pipeline — its own scratch planes, streams and slot, created at the first overlapping
    // submit — so that the kernels of two consecutive calls run beside each other instead of in
    // stream order: the descriptor stage of frame i (14 warps per SM) and the launch-bound
    // compaction chain leave SMs idle that the pyramid of frame i + 1 fills. Large contexts (more
    // than 8 frames per call or more than 6 GB) keep
    // both lanes as slots of the one pipeline (a batch fills the machine by itself, and the
    // scratch memory is what bounds the batch).
    SiftContext* twin = nullptr;
    bool twinMode = false, isTwin = false;
    SiftContext* lastOwner = nullptr;   // context holding the last completed call's slot (this or twin)
    int head = 0, next = 0, nPending = 0;
    std::vector<std::pair<char*, size_t>> registered;   // caller memory pinned by sift_register_host_memory
    Slot* last = nullptr;       // slot holding the results of the last completed call

    // staged path input
    const uint8_t* curInput = nullptr;
    int curPitch = 0;
    int64_t curFrameStride = 0;
    int curFrames = 0;
    bool executed = false;      // pyramid + gradients of the last call are on the device

    // CUDA graphs (opt-in: sift_set_graph_replay / SIFTCUDA_GRAPH=1)
    bool graphsEnabled = false;
    std::vector<GraphEntry> graphs;
    uint64_t useCounter = 0;

    // record-shaped views for the reference's two-step API (sift_detect / sift_describe)
    std::vector<SiftKeypoint> kpRecords;
    std::vector<SiftDescriptor> descRecords;
    std::vector<int> kpSegHost;

    // matching
    void* matcher = nullptr;

    // timing
    bool stageTiming = false;
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

size_t alignUp(size_t v, size_t a) { return (v + a - 1) / a * a; }

// Result columns carved out of one block (device or pinned host): every column 256-byte aligned.
size_t columnsBytes(size_t capKp, size_t capDesc) {
    size_t b = 0;
    b += 5 * alignUp(capKp * sizeof(float), 256);
    b += alignUp(capKp * sizeof(short2), 256);
    b += alignUp(capKp * sizeof(uchar2), 256);
    b += alignUp(capDesc * 128, 256);
    b += 2 * alignUp(capDesc * 4, 256);
    return std::max<size_t>(b, 256);
}
ResultColumns carveColumns(void* block, size_t capKp, size_t capDesc) {
    ResultColumns r;
    char* p = (char*)block;
    auto take = [&](size_t bytes) { char* q = p; p += alignUp(bytes, 256); return q; };
    r.kp.absX = (float*)take(capKp * 4);
    r.kp.absY = (float*)take(capKp * 4);
    r.kp.sigma = (float*)take(capKp * 4);
    r.kp.value = (float*)take(capKp * 4);
    r.kp.subScale = (float*)take(capKp * 4);
    r.kp.scaledXY = (short2*)take(capKp * sizeof(short2));
    r.kp.octaveScale = (uchar2*)take(capKp * sizeof(uchar2));
    r.desc.features = (uint8_t*)take(capDesc * 128);
    r.desc.theta = (float*)take(capDesc * 4);
    r.desc.keypoint = (int32_t*)take(capDesc * 4);
    return r;
}

// Slot resources: slot 0 at create, slot 1 at the first call that needs two calls in flight.
int ensureSlot(SiftContext* c, int s) {
    Slot& S = c->slot[s];
    if (S.allocated) return SIFT_OK;
    const size_t frameBytes = (size_t)c->cfg.width * c->cfg.height * c->bytesPerPixel;
    cudaError_t e = cudaSuccess;
    auto A = [&](cudaError_t r) { if (e == cudaSuccess) e = r; };
    A(devAlloc(c, &S.dInput, (size_t)c->B * frameBytes));
    if (s == 0) {
        S.dev = c->dev;
    } else {
        char* dblock = nullptr;
        A(devAlloc(c, &dblock, columnsBytes((size_t)c->capKp, (size_t)c->capDesc)));
        if (e == cudaSuccess) S.dev = carveColumns(dblock, (size_t)c->capKp, (size_t)c->capDesc);
    }
    void* block = nullptr;
    A(cudaMallocHost(&block, columnsBytes((size_t)c->capKp, (size_t)c->capDesc)));
    if (e == cudaSuccess) {
        S.host = carveColumns(block, (size_t)c->capKp, (size_t)c->capDesc);
        S.hostBlock = block;
        S.hostBlockExternal = false;
    }
    A(cudaMallocHost(&S.hCounters, sizeof(Counters)));
    A(cudaMallocHost(&S.hSegStarts, 3 * (size_t)(c->nSegs + 1) * sizeof(int)));
    A(cudaEventCreateWithFlags(&S.evUploaded, cudaEventDisableTiming));
    A(cudaEventCreate(&S.evStart));
    A(cudaEventCreate(&S.evEnd));
    if (e != cudaSuccess)
        return fail(c, e == cudaErrorMemoryAllocation ? SIFT_ERR_OUT_OF_MEMORY : SIFT_ERR_CUDA, "slot allocation", e);
    memset(S.hCounters, 0, sizeof(Counters));
    memset(S.hSegStarts, 0, 3 * (size_t)(c->nSegs + 1) * sizeof(int));
    S.kpCounts.assign(c->nSegs, 0);
    S.descCounts.assign(c->nSegs, 0);
    S.candCounts.assign(c->nSegs, 0);
    S.allocated = true;
    c->info.device_bytes = c->deviceBytes;
    return SIFT_OK;
}

// Octave 0 of a large single frame runs as independent row-band blur chains (enqueuePipeline).
bool bandedOctave0(const SiftContext* c, int frames) {
    const OctaveDev& q = c->P.oct[0];
    const long tiles = (long)((q.w + 63) / 64) * ((q.h + 63) / 64) * frames;
    const int nb = c->nBands;
    return frames == 1 && nb > 1 && tiles >= 8L * c->smCount && q.h >= 256 * nb;
}
int bandBoundary(const SiftContext* c, int band) {   // first row of `band` (multiple of the tile height)
    const OctaveDev& q = c->P.oct[0];
    if (band >= c->nBands) return q.h;
    return (int)(((long)q.h * band / c->nBands + 63) / 64 * 64);
}

void destroyMatcher(SiftContext* c);

void destroy(SiftContext* c) {
    if (!c) return;
    if (c->twin) destroy(c->twin);
    c->twin = nullptr;
    cudaSetDevice(c->device);
    // drain every stream of the context (an upload may still be in flight on the copy stream)
    if (c->stream) cudaStreamSynchronize(c->stream);
    if (c->copyStream) cudaStreamSynchronize(c->copyStream);
    if (c->downStream) cudaStreamSynchronize(c->downStream);
    for (int o = 1; o < kOctaves; o++)
        if (c->octStream[o]) cudaStreamSynchronize(c->octStream[o]);
    for (int b = 1; b < SiftContext::kMaxBands; b++)
        if (c->bandStream[b]) cudaStreamSynchronize(c->bandStream[b]);
    destroyMatcher(c);
    for (auto& g : c->graphs)
        if (g.exec) cudaGraphExecDestroy(g.exec);
    for (void* p : c->allocations) cudaFree(p);
    for (auto& r : c->registered) cudaHostUnregister(r.first);
    for (auto& S : c->slot) {
        if (S.hostBlock && !S.hostBlockExternal) cudaFreeHost(S.hostBlock);
        if (S.hCounters) cudaFreeHost(S.hCounters);
        if (S.hSegStarts) cudaFreeHost(S.hSegStarts);
        if (S.evUploaded) cudaEventDestroy(S.evUploaded);
        if (S.evStart) cudaEventDestroy(S.evStart);
        if (S.evEnd) cudaEventDestroy(S.evEnd);
    }
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
    if (c->copyStream) cudaStreamDestroy(c->copyStream);
    if (c->downStream) cudaStreamDestroy(c->downStream);
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
}

bool finiteF(float v) { return v == v && v <= 3.0e38f && v >= -3.0e38f; }

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
    cfg->input_format = SIFT_INPUT_BGRA8;
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
        case SIFT_ERR_BUSY: return "no free in-flight slot / nothing submitted";
        default: return "unknown status";
    }
}

const char* sift_last_error_string(const SiftContext* c) { return c ? c->lastError.c_str() : ""; }

}  // extern "C"

static int createContext(const SiftConfig* cfg, int device, SiftContext** out, bool isTwin) {
    if (!cfg || !out) return SIFT_ERR_INVALID_ARGUMENT;
    *out = nullptr;
    if (cfg->width < 8 || cfg->height < 8 || cfg->width > 16384 || cfg->height > 16384 ||
        cfg->max_batch < 1)
        return SIFT_ERR_INVALID_ARGUMENT;
    // the 3x3x3 stencils of refinement read x - 1 .. x + 1 around every accepted position: a
    // border below 1 would let them leave the plane (SIFTInterpolate.metal:180-190 uses 5)
    if (cfg->image_border < 1 || cfg->image_border > 16384) return SIFT_ERR_INVALID_ARGUMENT;
    if (!finiteF(cfg->dog_threshold) || !finiteF(cfg->edge_threshold) || !finiteF(cfg->max_offset) ||
        !finiteF(cfg->lambda_orientation) || !finiteF(cfg->orientation_threshold) ||
        cfg->edge_threshold <= 0.0f || cfg->lambda_orientation <= 0.0f || cfg->max_offset < 0.0f)
        return SIFT_ERR_INVALID_ARGUMENT;
    if (cfg->max_interpolation_iterations < 0 || cfg->max_interpolation_iterations > 1000 ||
        cfg->orientation_smoothing_iterations < 0 || cfg->orientation_smoothing_iterations > 1000 ||
        cfg->max_candidates_per_frame < 0 || cfg->max_keypoints_per_frame < 0 ||
        cfg->max_descriptors_per_frame < 0 || cfg->reserved != 0)
        return SIFT_ERR_INVALID_ARGUMENT;
    if (cfg->input_format != SIFT_INPUT_BGRA8 && cfg->input_format != SIFT_INPUT_GRAY8 &&
        cfg->input_format != SIFT_INPUT_NV12)
        return SIFT_ERR_INVALID_ARGUMENT;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || device < 0 || device >= count)
        return SIFT_ERR_NO_DEVICE;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return SIFT_ERR_NO_DEVICE;
    // the library carries arch-specific sm_100a SASS only (no PTX): other 10.x parts cannot run it
    if (prop.major != 10 || prop.minor != 0) return SIFT_ERR_NO_DEVICE;
    if (cudaSetDevice(device) != cudaSuccess) return SIFT_ERR_NO_DEVICE;

    SiftContext* c = new (std::nothrow) SiftContext();
    if (!c) return SIFT_ERR_OUT_OF_MEMORY;
    c->cfg = *cfg;
    c->device = device;
    c->smCount = prop.multiProcessorCount;
    c->B = cfg->max_batch;
    c->nSegs = c->B * kOctaves;
    c->bytesPerPixel = cfg->input_format == SIFT_INPUT_BGRA8 ? 4 : 1;
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
    A(devAlloc(c, &c->dGray, B * (size_t)W * H));
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
    A(devAlloc(c, &c->dBlockSums, nBlockSums));
    A(devAlloc(c, &c->dCands, (size_t)c->capCand));
    A(devAlloc(c, &c->dKpTmp, (size_t)c->capCand));
    A(devAlloc(c, &c->dFlagWords, flagWords));
    A(devAlloc(c, &c->dKps, (size_t)c->capKp));
    A(devAlloc(c, &c->dKpSeg, (size_t)c->capKp));
    A(devAlloc(c, &c->dSegStarts, 3 * (size_t)(c->nSegs + 1)));
    A(devAlloc(c, &c->dNOri, (size_t)c->capKp + kScanChunk));
    A(devAlloc(c, &c->dOriTmp, (size_t)c->capKp * kOriBins));
    A(devAlloc(c, &c->dOriOffset, (size_t)c->capKp + kScanChunk + 1));
    A(devAlloc(c, &c->dDescKp, (size_t)c->capDesc));
    A(devAlloc(c, &c->dCounters, 1));
    {
        char* block = nullptr;
        A(devAlloc(c, &block, columnsBytes((size_t)c->capKp, (size_t)c->capDesc)));
        if (e == cudaSuccess) c->dev = carveColumns(block, (size_t)c->capKp, (size_t)c->capDesc);
    }
    if (e == cudaSuccess) {
        // TMA descriptors for the blur's interior tiles (cuTensorMapEncodeTiled through the runtime's
        // driver entry point); a plane set too small for any interior tile just keeps cp.async loads
        A(makeBlurTmaSet(&c->seedTma, c->dScaled, c->P.oct[0].pitch, c->P.oct[0].h, (int)B, c->P.oct[0].plane));
        for (int o = 0; o < kOctaves; o++) {
            const OctaveDev& q = c->P.oct[o];
            if (q.w >= 1 && q.h >= 1) A(makeBlurTmaSet(&c->octTma[o], q.G, q.pitch, q.h, (int)B * kGaussians, q.plane));
            if (q.w >= 512 && q.h >= 128 && e == cudaSuccess) {
                A(makeExtremaTmaMap(&c->dogTma[o], q.D, q.pitch, q.h, (int)B * kDogs, q.plane));
                c->dogTmaValid[o] = (e == cudaSuccess);
            }
        }
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
    if (e == cudaSuccess) A(cudaMemset(c->dSegStarts, 0, 3 * (size_t)(c->nSegs + 1) * sizeof(int)));
    if (e == cudaSuccess) A(cudaMemset(c->dCounters, 0, sizeof(Counters)));
    for (auto& evn : c->ev) A(cudaEventCreate(&evn));
    for (auto& evn : c->evBlur0) A(cudaEventCreate(&evn));
    int prioLow = 0, prioHigh = 0;
    cudaDeviceGetStreamPriorityRange(&prioLow, &prioHigh);
    c->octStream[0] = c->stream;
    A(cudaEventCreateWithFlags(&c->evBandFork, cudaEventDisableTiming));
    A(cudaStreamCreateWithFlags(&c->copyStream, cudaStreamNonBlocking));
    A(cudaStreamCreateWithPriority(&c->downStream, cudaStreamNonBlocking, prioHigh));
    if (const char* rc = getenv("SIFTCUDA_RESULT_COPY")) c->resultCopy = atoi(rc);
    for (int b = 1; b < SiftContext::kMaxBands; b++) {
        A(cudaStreamCreateWithFlags(&c->bandStream[b], cudaStreamNonBlocking));
        A(cudaEventCreateWithFlags(&c->evBandSeeded[b], cudaEventDisableTiming));
        A(cudaEventCreateWithFlags(&c->evBandDone[b], cudaEventDisableTiming));
    }
    if (const char* nb = getenv("SIFTCUDA_BANDS")) c->nBands = std::max(1, std::min(SiftContext::kMaxBands, atoi(nb)));
    if (const char* g = getenv("SIFTCUDA_GRAPH")) c->graphsEnabled = atoi(g) != 0;
    for (int o = 0; o < kOctaves; o++) {
        // smaller octaves form a long dependent chain of tiny launches: give their streams a higher
        // priority so that their CTAs are placed ahead of the big octave-0 kernels' when SM slots
        // free up (the chain is latency-critical, octave 0 is throughput-bound)
        // (0 = "no explicit priority" in BlurArgs: octave 0 keeps the default)
        c->octPriority[o] = std::max(prioHigh, prioLow - o);
        if (o > 0) A(cudaStreamCreateWithPriority(&c->octStream[o], cudaStreamNonBlocking, c->octPriority[o]));
        A(cudaEventCreateWithFlags(&c->evSeeded[o], cudaEventDisableTiming));
        A(cudaEventCreateWithFlags(&c->evOctDone[o], cudaEventDisableTiming));
    }
    if (e != cudaSuccess) {
        const bool oom = (e == cudaErrorMemoryAllocation);
        destroy(c);
        cudaGetLastError();
        return oom ? SIFT_ERR_OUT_OF_MEMORY : SIFT_ERR_CUDA;
    }
    const int rs = ensureSlot(c, 0);
    if (rs != SIFT_OK) {
        destroy(c);
        cudaGetLastError();
        return rs;
    }
    I.device_bytes = c->deviceBytes;
    c->isTwin = isTwin;
    // SIFTCUDA_TWIN=0: both lanes in the one pipeline (tuning); SIFTCUDA_TWIN_MB: the size limit
    static const bool twinEnabled = !(getenv("SIFTCUDA_TWIN") && atoi(getenv("SIFTCUDA_TWIN")) == 0);
    static const long twinLimitMb = getenv("SIFTCUDA_TWIN_MB") ? atol(getenv("SIFTCUDA_TWIN_MB")) : 6144;
    // batches of more than 8 frames fill the machine by themselves (and were measured that way)
    c->twinMode = !isTwin && twinEnabled && c->B <= 8 && c->deviceBytes <= (int64_t)twinLimitMb * (1 << 20);
    *out = c;
    return SIFT_OK;
}

// The second pipeline of a small context (lane 1 of the pipelined calls).
static int ensureTwin(SiftContext* c) {
    if (c->twin) return SIFT_OK;
    SiftContext* t = nullptr;
    const int r = createContext(&c->cfg, c->device, &t, true);
    if (r != SIFT_OK) {
        // no room (or no luck): lane 1 becomes slot 1 of the one pipeline, as in a large context
        cudaGetLastError();
        c->twinMode = false;
        return SIFT_OK;
    }
    t->graphsEnabled = c->graphsEnabled;
    t->stageTiming = c->stageTiming;
    c->twin = t;
    c->info.device_bytes = c->deviceBytes + t->deviceBytes;
    return SIFT_OK;
}
struct LaneRef {
    SiftContext* x;
    int s;
};
// Lane (0 / 1, what the ABI calls a slot) → the pipeline and slot behind it.
static LaneRef laneOf(SiftContext* c, int lane) {
    if (c->twinMode && lane == 1) return {c->twin, 0};
    return {c, lane};
}

extern "C" {

int sift_create(const SiftConfig* cfg, int device, SiftContext** out) { return createContext(cfg, device, out, false); }

void sift_destroy(SiftContext* c) { destroy(c); }

int sift_get_info(const SiftContext* c, SiftInfo* out) {
    if (!c || !out) return SIFT_ERR_INVALID_ARGUMENT;
    *out = c->info;
    return SIFT_OK;
}

int sift_set_stage_timing(SiftContext* c, int32_t enabled) {
    if (!c) return SIFT_ERR_INVALID_ARGUMENT;
    c->stageTiming = enabled != 0;
    if (c->twin) c->twin->stageTiming = c->stageTiming;
    return SIFT_OK;
}

int sift_set_graph_replay(SiftContext* c, int32_t enabled) {
    if (!c) return SIFT_ERR_INVALID_ARGUMENT;
    c->graphsEnabled = enabled != 0;
    if (c->twin) c->twin->graphsEnabled = c->graphsEnabled;
    return SIFT_OK;
}

int sift_last_timings(const SiftContext* c, SiftTimings* out) {
    if (!c || !out) return SIFT_ERR_INVALID_ARGUMENT;
    *out = c->timings;
    return SIFT_OK;
}

int sift_pending(const SiftContext* c) { return c ? c->nPending : 0; }

int sift_next_slot(const SiftContext* c) { return c ? c->next : 0; }

int sift_result_layout(const SiftContext* c, SiftResultLayout* out) {
    if (!c || !out) return SIFT_ERR_INVALID_ARGUMENT;
    memset(out, 0, sizeof *out);
    out->bytes = (int64_t)columnsBytes((size_t)c->capKp, (size_t)c->capDesc);
    out->capacity_keypoints = c->capKp;
    out->capacity_descriptors = c->capDesc;
    const ResultColumns r = carveColumns(nullptr, (size_t)c->capKp, (size_t)c->capDesc);
    const void* cols[10] = {r.kp.absX, r.kp.absY, r.kp.sigma, r.kp.value, r.kp.subScale, r.kp.scaledXY,
                            r.kp.octaveScale, r.desc.features, r.desc.theta, r.desc.keypoint};
    for (int i = 0; i < 10; i++) out->offset[i] = (int64_t)(uintptr_t)cols[i];
    return SIFT_OK;
}

int sift_register_host_memory(SiftContext* c, void* base, int64_t bytes) {
    if (!c || !base || bytes < 1) return SIFT_ERR_INVALID_ARGUMENT;
    CTX_TRY(c, cudaSetDevice(c->device));
    // the kernels store result columns through the device mapping of this memory
    CTX_TRY(c, cudaHostRegister(base, (size_t)bytes, cudaHostRegisterPortable | cudaHostRegisterMapped));
    void* dptr = nullptr;
    const cudaError_t e = cudaHostGetDevicePointer(&dptr, base, 0);
    if (e != cudaSuccess || dptr != base) {   // unified addressing: the device uses the host pointer
        cudaHostUnregister(base);
        return fail(c, SIFT_ERR_CUDA, "registered memory is not addressable by its host pointer", e);
    }
    c->registered.emplace_back((char*)base, (size_t)bytes);
    return SIFT_OK;
}

int sift_bind_result_memory(SiftContext* c, int32_t slot, void* base, int64_t bytes) {
    if (!c || slot < 0 || slot > 1 || !base || ((uintptr_t)base & 255)) return SIFT_ERR_INVALID_ARGUMENT;
    const size_t need = columnsBytes((size_t)c->capKp, (size_t)c->capDesc);
    if (bytes < (int64_t)need)
        return fail(c, SIFT_ERR_INVALID_ARGUMENT, "sift_bind_result_memory: block smaller than sift_result_layout().bytes");
    bool inside = false;
    for (auto& r : c->registered)
        inside = inside || ((char*)base >= r.first && (char*)base + need <= r.first + r.second);
    if (!inside)
        return fail(c, SIFT_ERR_INVALID_ARGUMENT, "sift_bind_result_memory: not inside memory given to sift_register_host_memory");
    CTX_TRY(c, cudaSetDevice(c->device));
    if (c->twinMode && slot == 1) {
        const int rt = ensureTwin(c);
        if (rt != SIFT_OK) return rt;
    }
    const LaneRef L = laneOf(c, slot);
    SiftContext* x = L.x;
    if (x->slot[L.s].pending) return fail(c, SIFT_ERR_BUSY, "sift_bind_result_memory: the slot has a call in flight");
    const int r = ensureSlot(x, L.s);
    if (r != SIFT_OK) {
        if (x != c) c->lastError = x->lastError;
        return r;
    }
    Slot& S = x->slot[L.s];
    if (S.hostBlock && !S.hostBlockExternal) cudaFreeHost(S.hostBlock);
    S.hostBlock = base;
    S.hostBlockExternal = true;
    S.host = carveColumns(base, (size_t)c->capKp, (size_t)c->capDesc);
    // recorded graphs of this slot carry the old pointers
    for (size_t i = 0; i < x->graphs.size();) {
        if (x->graphs[i].slot == L.s) {
            if (x->graphs[i].exec) cudaGraphExecDestroy(x->graphs[i].exec);
            x->graphs.erase(x->graphs.begin() + (long)i);
        } else {
            i++;
        }
    }
    if (x->last == &S) { x->last = nullptr; x->executed = false; }
    return SIFT_OK;
}

}  // extern "C"

namespace {

struct RunArgs {
    const uint8_t* input = nullptr;
    int pitch = 0;
    int64_t frameStride = 0;
    int frames = 0;
    bool withDescribe = true;
    bool hostOut = false;        // keypoint / descriptor columns go to the slot's pinned arrays
    bool copyOut = false;        // ... through the slot's HBM columns and the copy engine at sift_wait
    Slot* slot = nullptr;
};

// Exactly the filled part of the slot's device columns → its host columns (pinned or registered
// memory), by the copy engine on `st`.
int enqueueColumnDownload(SiftContext* c, Slot& S, cudaStream_t st) {
    const size_t nk = (size_t)S.nKp, nd = (size_t)S.nDesc;
    const ResultColumns &d = S.dev, &h = S.host;
    if (nk) {
        CTX_TRY(c, cudaMemcpyAsync(h.kp.absX, d.kp.absX, nk * 4, cudaMemcpyDeviceToHost, st));
        CTX_TRY(c, cudaMemcpyAsync(h.kp.absY, d.kp.absY, nk * 4, cudaMemcpyDeviceToHost, st));
        CTX_TRY(c, cudaMemcpyAsync(h.kp.sigma, d.kp.sigma, nk * 4, cudaMemcpyDeviceToHost, st));
        CTX_TRY(c, cudaMemcpyAsync(h.kp.value, d.kp.value, nk * 4, cudaMemcpyDeviceToHost, st));
        CTX_TRY(c, cudaMemcpyAsync(h.kp.subScale, d.kp.subScale, nk * 4, cudaMemcpyDeviceToHost, st));
        CTX_TRY(c, cudaMemcpyAsync(h.kp.scaledXY, d.kp.scaledXY, nk * sizeof(short2), cudaMemcpyDeviceToHost, st));
        CTX_TRY(c, cudaMemcpyAsync(h.kp.octaveScale, d.kp.octaveScale, nk * sizeof(uchar2), cudaMemcpyDeviceToHost, st));
    }
    if (nd) {
        CTX_TRY(c, cudaMemcpyAsync(h.desc.features, d.desc.features, nd * 128, cudaMemcpyDeviceToHost, st));
        CTX_TRY(c, cudaMemcpyAsync(h.desc.theta, d.desc.theta, nd * 4, cudaMemcpyDeviceToHost, st));
        CTX_TRY(c, cudaMemcpyAsync(h.desc.keypoint, d.desc.keypoint, nd * 4, cudaMemcpyDeviceToHost, st));
    }
    return SIFT_OK;
}

// getDescriptors (SIFT.swift:207-238) over the keypoints in c->dKps.
int enqueueDescribe(SiftContext* c, const RunArgs& r, int nSegs, bool T) {
    cudaStream_t st = c->stream;
    NvtxRange range("getDescriptors(orientations) + getDescriptors(descriptors)");
    const DescriptorColumnsDev none{};
    CTX_TRY(c, launchDescribe(c->P, c->dKps, c->dKpSeg, c->capKp, c->dSegStarts + (c->nSegs + 1), c->dNOri, c->dOriTmp,
                              c->dOriOffset, c->dDescKp, c->dBlockSums, r.slot->dev.desc,
                              (r.hostOut && !r.copyOut) ? r.slot->host.desc : none, c->capDesc, c->dSegStarts + 2 * (c->nSegs + 1),
                              nSegs, c->dCounters, c->smCount, st, T ? c->ev[5] : nullptr));
    c->launches += 5;
    if (T) CTX_TRY(c, cudaEventRecord(c->ev[6], st));
    return SIFT_OK;
}

// The bookkeeping blocks follow the last kernel into the slot's pinned memory.
int enqueueReadback(SiftContext* c, const RunArgs& r) {
    cudaStream_t st = c->stream;
    const size_t segInts = 3 * (size_t)(c->nSegs + 1);
    CTX_TRY(c, cudaMemcpyAsync(r.slot->hCounters, c->dCounters, sizeof(Counters), cudaMemcpyDeviceToHost, st));
    CTX_TRY(c, cudaMemcpyAsync(r.slot->hSegStarts, c->dSegStarts, segInts * sizeof(int), cudaMemcpyDeviceToHost, st));
    return SIFT_OK;
}

// findKeypoints + getKeypointsFromOctaves + interpolateKeypoints (SIFT.swift:154-202) and
// getDescriptors (:207-238), all frames of the batch at once, no host synchronisation inside.
// Everything is enqueued on c->stream and streams forked from / joined to it by events, so the
// same code runs launch by launch or under stream capture.
int enqueuePipeline(SiftContext* c, const RunArgs& r, bool T) {
    const int F = r.frames;
    cudaStream_t st = c->stream;
    c->launches = 0;
    CTX_TRY(c, cudaMemsetAsync(c->dCounters, 0, sizeof(Counters), st));
    if (T) CTX_TRY(c, cudaEventRecord(c->ev[0], st));
    {
        NvtxRange range("findKeypoints");
        // DifferenceOfGaussians.encodeSeedTexture (:357-389)
        const OctaveDev& o0 = c->P.oct[0];
        CTX_TRY(c, launchGrayUpsample(r.input, c->bytesPerPixel, r.pitch, r.frameStride, c->dGray, c->cfg.width,
                                      c->cfg.height, c->dScaled, o0.w, o0.h, o0.pitch, o0.plane, F, st));
        BlurArgs seed{};
        seed.in = c->dScaled;
        seed.out = o0.G;  // octave 0 slice 0 (the reference blits seed → slice 0, :176-188)
        seed.w = o0.w; seed.h = o0.h; seed.pitch = o0.pitch;
        seed.inFrameStride = o0.plane;
        seed.outFrameStride = kGaussians * o0.plane;
        seed.frames = F;
        seed.tmaSet = &c->seedTma;
        seed.tmaZ0 = 0;
        seed.tmaZStride = 1;
        CTX_TRY(c, launchBlur(seed, c->seedTaps, c->seedNtaps, st));
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
        const int tailStart = tailStartOctave(c->P);
        for (int o = 0; o < kOctaves; o++) {
            const OctaveDev& q = c->P.oct[o];
            if (q.w < 1 || q.h < 1 || o > dbgMaxOctave) continue;
            cudaStream_t so = c->octStream[o];
            if (o == tailStart) {
                // this octave and every deeper one: one launch, planes resident in shared memory
                CTX_TRY(c, cudaStreamWaitEvent(so, c->evSeeded[o - 1], 0));
                if (o == 1 && c->bandedOctave0)
                    for (int b = 1; b < c->nBands; b++) CTX_TRY(c, cudaStreamWaitEvent(so, c->evBandSeeded[b], 0));
                forked[o] = true;
                CTX_TRY(c, launchTailOctaves(c->P, o, c->taps, c->ntaps, c->dMask, F, so, c->octPriority[o] ? c->octPriority[o] : kNoPriority));
                c->launches++;
                CTX_TRY(c, cudaEventRecord(c->evOctDone[o], so));
                break;
            }
            if (o > 0) {
                CTX_TRY(c, cudaStreamWaitEvent(so, c->evSeeded[o - 1], 0));
                if (o == 1 && c->bandedOctave0)
                    for (int b = 1; b < c->nBands; b++) CTX_TRY(c, cudaStreamWaitEvent(so, c->evBandSeeded[b], 0));
                forked[o] = true;
            }
            // row bands for a plane with many tiles (octave 0 of a large single frame); batches
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
                        int halo = 0;
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
                    a.tmaSet = &c->octTma[o];
                    a.tmaZ0 = s;
                    a.tmaZStride = kGaussians;
                    a.priority = c->octPriority[o];
                    // small planes: scale s + 1 launches under scale s (programmatic dependent launch)
                    static const int pdlMaxTiles = getenv("SIFTCUDA_PDL_TILES") ? atoi(getenv("SIFTCUDA_PDL_TILES")) : 160;
                    a.pdl = (s > 0 && !banded && (long)((q.w + 31) / 32) * ((q.h + 31) / 32) * F <= pdlMaxTiles) ? 1 : 0;
                    CTX_TRY(c, launchBlur(a, c->taps[s], c->ntaps[s], sb));
                    c->launches++;
                    if (s + 1 == kScales) CTX_TRY(c, cudaEventRecord(band == 0 ? c->evSeeded[o] : c->evBandSeeded[band], sb));
                }
            }
            if (banded) {   // join: gradient and extrema need every row
                for (int b = 1; b < nb; b++) {
                    CTX_TRY(c, cudaEventRecord(c->evBandDone[b], c->bandStream[b]));
                    CTX_TRY(c, cudaStreamWaitEvent(so, c->evBandDone[b], 0));
                }
            }
            if (o == 0) c->bandedOctave0 = banded;
            if (T && o == 0) CTX_TRY(c, cudaEventRecord(c->evBlur0[kGaussians - 1], so));
            const int prio = c->octPriority[o] ? c->octPriority[o] : kNoPriority;
            if (!(dbgSkip & 1)) CTX_TRY(c, launchGradient(q, F, so, 0, 0, prio));
            c->launches++;
            if (q.w >= 3 && q.h >= 3 && !(dbgSkip & 2)) {
                CTX_TRY(c, launchExtremaMask(c->P, o, c->dMask, F, so, 0, 0, prio, c->dogTmaValid[o] ? &c->dogTma[o] : nullptr));
                c->launches++;
            }
            if (o > 0) CTX_TRY(c, cudaEventRecord(c->evOctDone[o], so));
        }
        for (int o = 1; o < kOctaves; o++)
            if (forked[o]) CTX_TRY(c, cudaStreamWaitEvent(st, c->evOctDone[o], 0));
    }
    const int nSegs = F * kOctaves;
    if (T) CTX_TRY(c, cudaEventRecord(c->ev[2], st));
    {
        // mask → ordered candidates (SIFTOctave.getKeypoints :198-203)
        NvtxRange range("getKeypointsFromOctaves");
        CTX_TRY(c, launchCandidateCompaction(c->P, c->dMask, c->dBlockSums, c->dCands, c->capCand, 0,
                                             c->P.blocksPerFrame * F, c->dSegStarts, nSegs, c->dCounters, st));
        c->launches += 3;
        if (T) CTX_TRY(c, cudaEventRecord(c->ev[3], st));
    }
    {
        // refined keypoints (interpolateKeypoints :205-288); the result columns of the keypoints are
        // written by the scatter kernel (pinned host columns for the host-buffer calls)
        NvtxRange range("interpolateKeypoints");
        CTX_TRY(c, launchRefine(c->P, c->dCands, c->capCand, c->dKpTmp, c->dFlagWords, c->dBlockSums, c->dKps,
                                c->dKpSeg, c->capKp, c->dSegStarts, c->dSegStarts + (c->nSegs + 1), nSegs,
                                c->dCounters, (r.hostOut && !r.copyOut) ? r.slot->host.kp : r.slot->dev.kp, st));
        c->launches += 3;
        if (T) CTX_TRY(c, cudaEventRecord(c->ev[4], st));
    }
    if (r.withDescribe) {
        const int rd = enqueueDescribe(c, r, nSegs, T);
        if (rd != SIFT_OK) return rd;
    }
    return enqueueReadback(c, r);
}

GraphEntry* findGraph(SiftContext* c, int slot, const RunArgs& r) {
    for (auto& g : c->graphs)
        if (g.slot == slot && g.frames == r.frames && g.input == (const void*)r.input && g.pitch == r.pitch &&
            g.frameStride == r.frameStride && g.describe == r.withDescribe && g.hostOut == r.hostOut &&
            g.copyOut == r.copyOut)
            return &g;
    if (c->graphs.size() >= 16) {   // evict the least recently used entry
        size_t victim = 0;
        for (size_t i = 1; i < c->graphs.size(); i++)
            if (c->graphs[i].lastUse < c->graphs[victim].lastUse) victim = i;
        if (c->graphs[victim].exec) cudaGraphExecDestroy(c->graphs[victim].exec);
        c->graphs.erase(c->graphs.begin() + (long)victim);
    }
    GraphEntry g;
    g.slot = slot; g.frames = r.frames; g.input = r.input; g.pitch = r.pitch; g.frameStride = r.frameStride;
    g.describe = r.withDescribe; g.hostOut = r.hostOut; g.copyOut = r.copyOut;
    c->graphs.push_back(g);
    return &c->graphs.back();
}

// Runs the pipeline for one slot: as one CUDA graph launch when this (slot, input, batch size)
// has been seen before, launch by launch otherwise (first call; per-stage timing requested;
// SIFTCUDA_GRAPH=0). evStart / evEnd bracket the device work either way.
int runSlot(SiftContext* c, int slotIndex, RunArgs r) {
    Slot& S = c->slot[slotIndex];
    r.slot = &S;
    cudaStream_t st = c->stream;
    const bool T = c->stageTiming;
    S.frames = r.frames;
    S.described = r.withDescribe;
    S.hostOut = r.hostOut;
    S.graphReplay = false;
    S.copyOut = r.copyOut;
    CTX_TRY(c, cudaEventRecord(S.evStart, st));
    GraphEntry* g = (c->graphsEnabled && !T) ? findGraph(c, slotIndex, r) : nullptr;
    if (g) {
        g->lastUse = ++c->useCounter;
        g->seen++;
        if (g->seen == 2 && !g->exec) {
            // second call with this key: record the DAG once
            cudaGraph_t graph = nullptr;
            cudaError_t e = cudaStreamBeginCapture(st, cudaStreamCaptureModeRelaxed);
            int rp = SIFT_OK;
            if (e == cudaSuccess) {
                rp = enqueuePipeline(c, r, false);
                e = cudaStreamEndCapture(st, &graph);
            }
            if (e == cudaSuccess && rp == SIFT_OK && graph) e = cudaGraphInstantiate(&g->exec, graph, 0);
            if (graph) cudaGraphDestroy(graph);
            if (e != cudaSuccess || rp != SIFT_OK || !g->exec) {
                // capture is an optimisation: fall back to launching the same kernels one by one
                cudaGetLastError();
                g->exec = nullptr;
                c->graphsEnabled = false;
                fail(c, SIFT_ERR_CUDA, "CUDA graph capture failed; running launch by launch", e);
            } else {
                g->launches = c->launches;
            }
        }
        if (g->exec) {
            NvtxRange range("findKeypoints + getDescriptors (graph replay)");
            CTX_TRY(c, cudaGraphLaunch(g->exec, st));
            S.graphReplay = true;
            S.launches = g->launches;
            CTX_TRY(c, cudaEventRecord(S.evEnd, st));
            return SIFT_OK;
        }
    }
    const int rp = enqueuePipeline(c, r, T);
    if (rp != SIFT_OK) return rp;
    S.launches = c->launches;
    CTX_TRY(c, cudaEventRecord(S.evEnd, st));
    return SIFT_OK;
}

// Waits for the slot's device work; per-(frame, octave) counts, timings, overflow status.
int finishSlot(SiftContext* c, Slot& S) {
    CTX_TRY(c, cudaEventSynchronize(S.evEnd));
    const int nSegs = S.frames * kOctaves;
    const int* candStart = S.hSegStarts;
    const int* kpStart = S.hSegStarts + (c->nSegs + 1);
    const int* descStart = S.hSegStarts + 2 * (c->nSegs + 1);
    const int nDesc = S.hCounters->nDescriptors;
    for (int s = 0; s < nSegs; s++) {
        S.candCounts[s] = candStart[s + 1] - candStart[s];
        S.kpCounts[s] = kpStart[s + 1] - kpStart[s];
        S.descCounts[s] = S.described ? std::min(descStart[s + 1], nDesc) - std::min(descStart[s], nDesc) : 0;
    }
    S.nKp = std::min(S.hCounters->nKeypoints, c->capKp);
    S.nDesc = S.described ? std::min(S.hCounters->nDescriptors, c->capDesc) : 0;
    if (S.copyOut) {
        // the counts are known now: the copy engine moves the columns while the kernels of the next
        // submitted call run (a slot's device columns are its own, nothing overwrites them)
        const int rc = enqueueColumnDownload(c, S, c->downStream);
        if (rc != SIFT_OK) return rc;
        CTX_TRY(c, cudaStreamSynchronize(c->downStream));
    }
    SiftTimings& t = c->timings;
    memset(&t, 0, sizeof t);
    t.kernel_launches = S.launches;
    t.graph_replay = S.graphReplay ? 1 : 0;
    cudaEventElapsedTime(&t.total_ms, S.evStart, S.evEnd);
    if (c->stageTiming && !S.graphReplay) {
        t.stage_timing_enabled = 1;
        const int lastStage = S.described ? SIFT_STAGE_COUNT : 4;
        for (int i = 0; i < lastStage; i++) cudaEventElapsedTime(&t.stage_ms[i], c->ev[i], c->ev[i + 1]);
        cudaEventElapsedTime(&t.blur_octave0_ms, c->evBlur0[0], c->evBlur0[kGaussians - 1]);
        for (int s = 0; s < kGaussians - 1; s++)   // per-scale split only without row bands
            t.blur_octave0_launch_ms[s] = c->bandedOctave0 ? t.blur_octave0_ms / (kGaussians - 1) : 0.0f;
        if (!c->bandedOctave0)
            for (int s = 0; s < kGaussians - 1; s++)
                cudaEventElapsedTime(&t.blur_octave0_launch_ms[s], c->evBlur0[s], c->evBlur0[s + 1]);
        t.blur_octave0_launches = kGaussians - 1;
    }
    cudaGetLastError();   // an unrecorded timing event must not poison the next launch check
    S.status = SIFT_OK;
    if (S.hCounters->overflow) {
        char buf[160];
        snprintf(buf, sizeof buf, "list capacity exceeded (mask %d: 1 candidates, 2 keypoints, 4 descriptors)",
                 S.hCounters->overflow);
        S.status = fail(c, SIFT_ERR_CAPACITY, buf);
    }
    c->last = &S;
    c->lastOwner = c;
    c->executed = true;
    return S.status;
}

void fillResult(const SiftContext* c, const Slot& S, SiftBatchResult* out) {
    out->n_frames = S.frames;
    out->status = S.status;
    out->slot = (int32_t)(&S - c->slot);
    out->reserved = 0;
    out->keypoint_counts = S.kpCounts.data();
    out->descriptor_counts = S.descCounts.data();
    out->candidate_counts = S.candCounts.data();
    out->total_keypoints = S.nKp;
    out->total_descriptors = S.nDesc;
    out->keypoints.absolute_x = S.host.kp.absX;
    out->keypoints.absolute_y = S.host.kp.absY;
    out->keypoints.sigma = S.host.kp.sigma;
    out->keypoints.value = S.host.kp.value;
    out->keypoints.sub_scale = S.host.kp.subScale;
    out->keypoints.scaled_xy = reinterpret_cast<const int16_t*>(S.host.kp.scaledXY);
    out->keypoints.octave_scale = reinterpret_cast<const uint8_t*>(S.host.kp.octaveScale);
    out->descriptors.features = S.host.desc.features;
    out->descriptors.theta = S.host.desc.theta;
    out->descriptors.keypoint = S.host.desc.keypoint;
}

// Uploads n host frames into the slot's input arena on the copy stream; the compute stream waits
// for the last of them.
int enqueueUpload(SiftContext* c, Slot& S, const void* const* images, int n, int pitchBytes) {
    const size_t rowBytes = (size_t)c->cfg.width * c->bytesPerPixel;
    const size_t frameBytes = rowBytes * c->cfg.height;
    for (int f = 0; f < n; f++)
        CTX_TRY(c, cudaMemcpy2DAsync(S.dInput + f * frameBytes, rowBytes, images[f], (size_t)pitchBytes, rowBytes,
                                     c->cfg.height, cudaMemcpyHostToDevice, c->copyStream));
    CTX_TRY(c, cudaEventRecord(S.evUploaded, c->copyStream));
    return SIFT_OK;
}

int checkImages(SiftContext* c, const void* const* images, int n, int pitchBytes, const char* who) {
    if (!c) return SIFT_ERR_INVALID_ARGUMENT;
    if (!images || n < 1 || n > c->B || pitchBytes < c->cfg.width * c->bytesPerPixel)
        return fail(c, SIFT_ERR_INVALID_ARGUMENT, who);
    for (int f = 0; f < n; f++)
        if (!images[f]) return fail(c, SIFT_ERR_INVALID_ARGUMENT, "null image");
    return SIFT_OK;
}

// Upload + pipeline of one call on slot s of pipeline x.
int submitOn(SiftContext* x, int s, const void* const* images, int n, int pitchBytes, bool withDescribe,
             bool synchronous) {
    int r = ensureSlot(x, s);
    if (r != SIFT_OK) return r;
    Slot& S = x->slot[s];
    // tuning only: 1 = results stay in HBM (nothing delivered), 2 = no upload (stale input)
    static const int dbgE2E = getenv("SIFTCUDA_DEBUG_E2E") ? atoi(getenv("SIFTCUDA_DEBUG_E2E")) : 0;
    if (!(dbgE2E & 2)) {
        r = enqueueUpload(x, S, images, n, pitchBytes);
        if (r != SIFT_OK) return r;
        CTX_TRY(x, cudaStreamWaitEvent(x->stream, S.evUploaded, 0));
    }
    RunArgs a;
    a.input = S.dInput;
    a.pitch = x->cfg.width * x->bytesPerPixel;
    a.frameStride = (int64_t)a.pitch * x->cfg.height;
    a.frames = n;
    a.withDescribe = withDescribe;
    a.hostOut = !(dbgE2E & 1);
    a.copyOut = a.hostOut && (x->resultCopy == 2 || (x->resultCopy == 1 && !synchronous));
    r = runSlot(x, s, a);
    if (r != SIFT_OK) return r;
    S.pending = true;
    S.staged = false;
    x->curInput = nullptr;
    return SIFT_OK;
}

int submitHost(SiftContext* c, const void* const* images, int n, int pitchBytes, bool withDescribe, bool synchronous) {
    int r = checkImages(c, images, n, pitchBytes, "submit: bad arguments");
    if (r != SIFT_OK) return r;
    if (c->nPending >= 2) return fail(c, SIFT_ERR_BUSY, "both in-flight slots are taken: call sift_wait first");
    CTX_TRY(c, cudaSetDevice(c->device));
    // synchronous callers only ever use lane 0 (lane 1 — the second pipeline of a small context,
    // slot 1 of a large one — is allocated by the first overlapping submit); pipelined callers
    // alternate, so a result stays valid across the next submit
    if (synchronous) c->next = 0;
    if (c->nPending == 0) c->head = c->next;
    const int lane = c->next;
    if (c->twinMode && lane == 1) {
        r = ensureTwin(c);
        if (r != SIFT_OK) return r;
    }
    const LaneRef L = laneOf(c, lane);
    r = submitOn(L.x, L.s, images, n, pitchBytes, withDescribe, synchronous);
    if (r != SIFT_OK) {
        if (L.x != c) c->lastError = L.x->lastError;
        return r;
    }
    c->nPending++;
    c->next ^= 1;
    return SIFT_OK;
}

int waitOldest(SiftContext* c, SiftBatchResult* out) {
    if (c->nPending < 1) return fail(c, SIFT_ERR_BUSY, "sift_wait: nothing submitted");
    CTX_TRY(c, cudaSetDevice(c->device));
    const int lane = c->head;
    const LaneRef L = laneOf(c, lane);
    Slot& S = L.x->slot[L.s];
    const int r = finishSlot(L.x, S);
    S.pending = false;
    c->nPending--;
    c->head ^= 1;
    c->lastOwner = L.x;
    if (L.x != c) {
        // the planes of this call live in the second pipeline: the debug taps and the staged
        // download of the first one have nothing current to show
        c->timings = L.x->timings;
        c->lastError = L.x->lastError;
        c->last = nullptr;
        c->executed = false;
    }
    if (r != SIFT_OK && r != SIFT_ERR_CAPACITY) return r;
    if (out) {
        fillResult(L.x, S, out);
        out->slot = lane;
    }
    return r;
}

}  // namespace

extern "C" {

int sift_submit(SiftContext* c, const void* const* images, int32_t n, int32_t pitchBytes) {
    if (!c) return SIFT_ERR_INVALID_ARGUMENT;
    return submitHost(c, images, n, pitchBytes, true, false);
}

int sift_wait(SiftContext* c, SiftBatchResult* out) {
    if (!c || !out) return SIFT_ERR_INVALID_ARGUMENT;
    return waitOldest(c, out);
}

int sift_detect_and_describe_batch(SiftContext* c, const void* const* images, int32_t n,
                                   int32_t pitchBytes, SiftBatchResult* out) {
    if (!c || !out) return SIFT_ERR_INVALID_ARGUMENT;
    if (c->nPending) return fail(c, SIFT_ERR_BUSY, "synchronous call with submitted work in flight");
    const int r = submitHost(c, images, n, pitchBytes, true, true);
    if (r != SIFT_OK) return r;
    return waitOldest(c, out);
}

int sift_batch_upload(SiftContext* c, const void* const* images, int32_t n, int32_t pitchBytes) {
    int r = checkImages(c, images, n, pitchBytes, "sift_batch_upload: bad arguments");
    if (r != SIFT_OK) return r;
    if (c->nPending) return fail(c, SIFT_ERR_BUSY, "staged call with submitted work in flight");
    CTX_TRY(c, cudaSetDevice(c->device));
    Slot& S = c->slot[0];
    r = enqueueUpload(c, S, images, n, pitchBytes);
    if (r != SIFT_OK) return r;
    // the host frames belong to the caller again when this returns
    CTX_TRY(c, cudaEventSynchronize(S.evUploaded));
    c->curInput = S.dInput;
    c->curPitch = c->cfg.width * c->bytesPerPixel;
    c->curFrameStride = (int64_t)c->curPitch * c->cfg.height;
    c->curFrames = n;
    c->executed = false;
    return SIFT_OK;
}

int sift_batch_set_device_input(SiftContext* c, const void* dev, int32_t n, int32_t pitchBytes,
                                int64_t frameStrideBytes) {
    if (!c) return SIFT_ERR_INVALID_ARGUMENT;
    const int align = c->bytesPerPixel == 4 ? 3 : 0;   // BGRA pixels are read as 32-bit words
    if (!dev || n < 1 || n > c->B || pitchBytes < c->cfg.width * c->bytesPerPixel ||
        (n > 1 && frameStrideBytes < (int64_t)pitchBytes * c->cfg.height) || (pitchBytes & align) ||
        (frameStrideBytes & align) || ((uintptr_t)dev & (uintptr_t)align))
        return fail(c, SIFT_ERR_INVALID_ARGUMENT, "sift_batch_set_device_input: bad arguments");
    c->curInput = (const uint8_t*)dev;
    c->curPitch = pitchBytes;
    c->curFrameStride = frameStrideBytes;
    c->curFrames = n;
    c->executed = false;
    return SIFT_OK;
}

int sift_batch_execute(SiftContext* c) {
    if (!c) return SIFT_ERR_INVALID_ARGUMENT;
    if (!c->curInput || c->curFrames < 1) return fail(c, SIFT_ERR_INVALID_ARGUMENT, "no input set");
    if (c->nPending) return fail(c, SIFT_ERR_BUSY, "staged call with submitted work in flight");
    CTX_TRY(c, cudaSetDevice(c->device));
    RunArgs a;
    a.input = c->curInput;
    a.pitch = c->curPitch;
    a.frameStride = c->curFrameStride;
    a.frames = c->curFrames;
    a.withDescribe = true;
    a.hostOut = false;   // staged path: results stay in HBM until sift_batch_download
    const int r = runSlot(c, 0, a);
    if (r != SIFT_OK) return r;
    c->slot[0].staged = true;
    return finishSlot(c, c->slot[0]);
}

int sift_batch_download(SiftContext* c, SiftBatchResult* out) {
    if (!c || !out) return SIFT_ERR_INVALID_ARGUMENT;
    if (!c->executed || !c->last) return fail(c, SIFT_ERR_NOT_DETECTED, "download before execute");
    CTX_TRY(c, cudaSetDevice(c->device));
    Slot& S = *c->last;
    if (S.staged) {
        // device columns → the slot's pinned columns
        cudaStream_t st = c->stream;
        const int rc = enqueueColumnDownload(c, S, st);
        if (rc != SIFT_OK) return rc;
        CTX_TRY(c, cudaStreamSynchronize(st));
    }
    fillResult(c, S, out);
    return SIFT_OK;
}

int sift_materialize_keypoints(const SiftContext* c, const SiftBatchResult* r, int64_t first, int64_t count,
                               SiftKeypoint* dst) {
    if (!c || !r || first < 0 || count < 0 || first + count > r->total_keypoints || (count > 0 && !dst))
        return SIFT_ERR_INVALID_ARGUMENT;
    const SiftKeypointColumns& k = r->keypoints;
    for (int64_t i = 0; i < count; i++) {
        const int64_t j = first + i;
        SiftKeypoint& d = dst[i];
        d.octave = k.octave_scale[2 * j];
        d.scale = k.octave_scale[2 * j + 1];
        d.subScale = k.sub_scale[j];
        d.scaledX = k.scaled_xy[2 * j];
        d.scaledY = k.scaled_xy[2 * j + 1];
        d.absoluteX = k.absolute_x[j];
        d.absoluteY = k.absolute_y[j];
        const OctaveDev& o = c->P.oct[d.octave < kOctaves ? d.octave : 0];
        d.normalizedX = (float)d.scaledX / (float)o.w;   // SIFTOctave.swift:278-281
        d.normalizedY = (float)d.scaledY / (float)o.h;
        d.sigma = k.sigma[j];
        d.value = k.value[j];
    }
    return SIFT_OK;
}

int sift_materialize_descriptors(const SiftBatchResult* r, int64_t first, int64_t count, SiftDescriptor* dst) {
    if (!r || first < 0 || count < 0 || first + count > r->total_descriptors || (count > 0 && !dst))
        return SIFT_ERR_INVALID_ARGUMENT;
    for (int64_t i = 0; i < count; i++) {
        const int64_t j = first + i;
        dst[i].keypoint = r->descriptors.keypoint[j];
        dst[i].theta = r->descriptors.theta[j];
        memcpy(dst[i].features, r->descriptors.features + j * 128, 128);
    }
    return SIFT_OK;
}

int sift_detect(SiftContext* c, const void* pixels, int32_t pitchBytes,
                const SiftKeypoint** outKps, int32_t counts[SIFT_NUM_OCTAVES]) {
    if (!c || !pixels || !outKps || !counts) return SIFT_ERR_INVALID_ARGUMENT;
    if (c->nPending) return fail(c, SIFT_ERR_BUSY, "synchronous call with submitted work in flight");
    const void* imgs[1] = {pixels};
    int r = submitHost(c, imgs, 1, pitchBytes, false, true);
    if (r != SIFT_OK) return r;
    SiftBatchResult res;
    const int re = waitOldest(c, &res);
    if (re != SIFT_OK && re != SIFT_ERR_CAPACITY) return re;
    c->kpRecords.resize((size_t)std::max<int64_t>(res.total_keypoints, 1));
    r = sift_materialize_keypoints(c, &res, 0, res.total_keypoints, c->kpRecords.data());
    if (r != SIFT_OK) return r;
    *outKps = c->kpRecords.data();
    for (int o = 0; o < kOctaves; o++) counts[o] = res.keypoint_counts[o];
    return re;
}

int sift_describe(SiftContext* c, const SiftKeypoint* kps, const int32_t counts[SIFT_NUM_OCTAVES],
                  const SiftDescriptor** outDesc, int32_t descCounts[SIFT_NUM_OCTAVES]) {
    if (!c || !counts || !outDesc || !descCounts) return SIFT_ERR_INVALID_ARGUMENT;
    if (!c->executed) return fail(c, SIFT_ERR_NOT_DETECTED, "sift_describe before sift_detect");
    if (c->nPending) return fail(c, SIFT_ERR_BUSY, "synchronous call with submitted work in flight");
    CTX_TRY(c, cudaSetDevice(c->device));
    int64_t n = 0;
    for (int o = 0; o < kOctaves; o++) {
        if (counts[o] < 0) return fail(c, SIFT_ERR_INVALID_ARGUMENT, "negative count");
        n += counts[o];
    }
    if (n > c->capKp) return fail(c, SIFT_ERR_CAPACITY, "more keypoints than max_keypoints_per_frame");
    if (n > 0 && !kps) return SIFT_ERR_INVALID_ARGUMENT;
    Slot& S = c->slot[0];
    // group check + segment table on the host; the keypoints may alias (or overlap) our own
    // record array from sift_detect, so they are only read here, never copied over themselves
    int* kpStart = S.hSegStarts + (c->nSegs + 1);
    c->kpSegHost.resize((size_t)std::max<int64_t>(n, 1));
    int k = 0;
    for (int o = 0; o < kOctaves; o++) {
        kpStart[o] = k;
        for (int i = 0; i < counts[o]; i++) {
            if (kps[k].octave != o) return fail(c, SIFT_ERR_INVALID_ARGUMENT, "keypoint octave does not match its group");
            c->kpSegHost[(size_t)k++] = o;
        }
    }
    for (int s = kOctaves; s <= c->nSegs; s++) kpStart[s] = k;
    cudaStream_t st = c->stream;
    memset(S.hCounters, 0, sizeof(Counters));   // incl. the kernels' work-queue counters
    S.hCounters->nKeypoints = (int)n;
    CTX_TRY(c, cudaEventRecord(S.evStart, st));
    CTX_TRY(c, cudaMemcpyAsync(c->dCounters, S.hCounters, sizeof(Counters), cudaMemcpyHostToDevice, st));
    if (n > 0) {
        CTX_TRY(c, cudaMemcpyAsync(c->dKps, kps, (size_t)n * sizeof(SiftKeypoint), cudaMemcpyHostToDevice, st));
        CTX_TRY(c, cudaMemcpyAsync(c->dKpSeg, c->kpSegHost.data(), (size_t)n * sizeof(int), cudaMemcpyHostToDevice, st));
    }
    CTX_TRY(c, cudaMemcpyAsync(c->dSegStarts + (c->nSegs + 1), kpStart, (size_t)(c->nSegs + 1) * sizeof(int),
                               cudaMemcpyHostToDevice, st));
    // the pageable sources above must have been consumed before the caller's array is touched again
    CTX_TRY(c, cudaStreamSynchronize(st));
    const bool T = c->stageTiming;
    c->launches = 0;
    RunArgs a;
    a.frames = 1;
    a.withDescribe = true;
    a.hostOut = true;
    a.slot = &S;
    if (T) CTX_TRY(c, cudaEventRecord(c->ev[4], st));
    int r = enqueueDescribe(c, a, kOctaves, T);
    if (r != SIFT_OK) return r;
    r = enqueueReadback(c, a);
    if (r != SIFT_OK) return r;
    CTX_TRY(c, cudaEventRecord(S.evEnd, st));
    CTX_TRY(c, cudaEventSynchronize(S.evEnd));
    // bookkeeping: only the descriptor segment table changed
    const int* descStart = S.hSegStarts + 2 * (c->nSegs + 1);
    const int nDesc = std::min(S.hCounters->nDescriptors, c->capDesc);
    memset(&c->timings, 0, sizeof c->timings);
    c->timings.kernel_launches = c->launches;
    cudaEventElapsedTime(&c->timings.total_ms, S.evStart, S.evEnd);
    if (T) {
        c->timings.stage_timing_enabled = 1;
        cudaEventElapsedTime(&c->timings.stage_ms[4], c->ev[4], c->ev[5]);
        cudaEventElapsedTime(&c->timings.stage_ms[5], c->ev[5], c->ev[6]);
    }
    cudaGetLastError();
    for (int o = 0; o < kOctaves; o++) descCounts[o] = std::min(descStart[o + 1], nDesc) - std::min(descStart[o], nDesc);
    S.frames = 1;
    S.described = true;
    S.hostOut = true;
    S.copyOut = false;
    S.staged = false;
    S.nDesc = nDesc;
    for (int o = 0; o < kOctaves; o++) S.descCounts[o] = descCounts[o];
    c->descRecords.resize((size_t)std::max(nDesc, 1));
    SiftBatchResult res;
    fillResult(c, S, &res);
    res.total_descriptors = nDesc;
    r = sift_materialize_descriptors(&res, 0, nDesc, c->descRecords.data());
    if (r != SIFT_OK) return r;
    *outDesc = c->descRecords.data();
    if (S.hCounters->overflow & 4) return fail(c, SIFT_ERR_CAPACITY, "descriptor capacity exceeded");
    return SIFT_OK;
}

int sift_debug_download(SiftContext* c, int32_t what, int32_t frame, int32_t octave, int32_t slice,
                        float* dst, int64_t dstFloats) {
    if (!c || !dst) return SIFT_ERR_INVALID_ARGUMENT;
    if (!c->executed || !c->last) return fail(c, SIFT_ERR_NOT_DETECTED, "debug download before execute");
    if (frame < 0 || frame >= c->last->frames) return fail(c, SIFT_ERR_INVALID_ARGUMENT, "bad frame");
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
    if (!c || !c->executed || !c->last || frame < 0 || frame >= c->last->frames || octave < 0 || octave >= kOctaves)
        return -(int64_t)SIFT_ERR_INVALID_ARGUMENT;
    if (cudaSetDevice(c->device) != cudaSuccess) return -(int64_t)SIFT_ERR_CUDA;
    const Slot& S = *c->last;
    const int seg = frame * kOctaves + octave;
    const int nAll = std::min(S.hCounters->nCandidates, c->capCand);
    const int a = std::min(S.hSegStarts[seg], nAll), b = std::min(S.hSegStarts[seg + 1], nAll);
    const int n = b - a;
    if (!dst || n <= 0) return n;
    std::vector<Candidate> tmp((size_t)n);
    if (cudaMemcpy(tmp.data(), c->dCands + a, (size_t)n * sizeof(Candidate), cudaMemcpyDeviceToHost) != cudaSuccess)
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
    if (!c->executed || !c->last) return fail(c, SIFT_ERR_NOT_DETECTED, "blur bench before execute");
    CTX_TRY(c, cudaSetDevice(c->device));
    const OctaveDev& q = c->P.oct[0];
    BlurArgs a{};
    a.in = q.G + (size_t)scale * q.plane;
    a.out = q.G + (size_t)(scale + 1) * q.plane;
    a.dog = q.D + (size_t)scale * q.plane;
    a.w = q.w; a.h = q.h; a.pitch = q.pitch;
    a.inFrameStride = a.outFrameStride = kGaussians * q.plane;
    a.dogFrameStride = kDogs * q.plane;
    a.frames = c->last->frames;
    a.tmaSet = &c->octTma[0];
    a.tmaZ0 = scale;
    a.tmaZStride = kGaussians;
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
    if (a.debugMode) c->executed = false;   // planes were overwritten in debug modes: force a fresh execute
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

#include "capi_match.inc"
