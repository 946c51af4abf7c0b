// describe.cu — orientation assignment and the 4x4x8 descriptor. One warp per keypoint /
// per descriptor, lane-private shared-memory histograms (no atomics), fixed-order reductions:
// results are run-to-run deterministic.
//
// Replaces:
//   SIFTOctave.getKeypointOrientations host filter (SIFTOctave.swift:303-337) +
//   SIFTOrientation.metal:16-175 siftOrientation                        → orientationKernel
//   SIFTOctave.getDescriptors expansion (SIFTOctave.swift:410-424) +
//   SIFTDescriptor.metal:15-237 siftDescriptors                         → descriptorKernel
//
// Parity class (north_star): θ within 1e-3 rad, features within ±1. The histogram sums are
// accumulated lane-strided and reduced in a fixed order, not in the reference's serial (j, i)
// order, so they differ from the oracle in the last bits; every quantity that drives a
// *discontinuous* decision (bin index, window radius, sample coordinate, border filter) is
// evaluated with the spec's exact operation sequence.
#include "common.cuh"
#include "dev_math.cuh"
#include "scan.cuh"

namespace sift {

constexpr float kTau = 6.28318530717958647692f;  // 2 * M_PI_F in float

// ------------------------------------------------------------------------------------------
constexpr int kOriWarps = 8;  // warps (keypoints in flight) per CTA

__global__ void __launch_bounds__(kOriWarps * 32)
orientationKernel(const __grid_constant__ EngineParams P, const SiftKeypoint* __restrict__ kps,
                  const int* __restrict__ kpSeg, const Counters* __restrict__ counters,
                  int* __restrict__ nOri, float* __restrict__ oriTmp) {
    __shared__ float sHist[kOriWarps][kOriBins * 32];  // [bin][lane] per warp
    __shared__ float sH[kOriWarps][2][kOriBins];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int n = counters->nKeypoints;
    float* hist = sHist[wid];
    for (int k = blockIdx.x * kOriWarps + wid; k < n; k += gridDim.x * kOriWarps) {
        const SiftKeypoint kp = kps[k];
        const int frame = kpSeg[k] / kOctaves;
        const OctaveDev& o = P.oct[kp.octave];
        // host filter of SIFTOctave.swift:303-329 on the untruncated coordinates
        const float lambda = P.lambdaOri;
        const float fx = __fdiv_rn(kp.absoluteX, o.delta);
        const float fy = __fdiv_rn(kp.absoluteY, o.delta);
        const float sigma = __fdiv_rn(kp.sigma, o.delta);
        const float rf = ceilf(__fmul_rn(__fmul_rn(3.0f, lambda), sigma));
        const bool reject = (floorf(__fsub_rn(fx, rf)) < 1.0f) ||
                            (ceilf(__fadd_rn(fx, rf)) > (float)(o.w - 2)) ||
                            (floorf(__fsub_rn(fy, rf)) < 1.0f) ||
                            (ceilf(__fadd_rn(fy, rf)) > (float)(o.h - 2));
        if (reject || kp.scale < 1 || kp.scale > kScales) {
            if (lane == 0) nOri[k] = 0;
            continue;
        }
        // kernel inputs are packed with Int32(absoluteCoordinate) (SIFTOctave.swift:333-334)
        const int x = (int)roundf(__fdiv_rn((float)(int)kp.absoluteX, o.delta));
        const int y = (int)roundf(__fdiv_rn((float)(int)kp.absoluteY, o.delta));
        const float expDen = __fmul_rn(__fmul_rn(2.0f, lambda), lambda);
        const int r = (int)rf;
        const int side = 2 * r + 1;
        const float2* __restrict__ g =
            o.grad + ((size_t)frame * kScales + (kp.scale - 1)) * o.plane;

#pragma unroll
        for (int b = 0; b < kOriBins; b++) hist[b * 32 + lane] = 0.0f;
        for (int idx = lane; idx < side * side; idx += 32) {
            const int jj = idx / side;
            const int j = jj - r, i = idx - jj * side - r;
            const int sx = x + i, sy = y + j;
            if (sx < 0 || sx >= o.w || sy < 0 || sy >= o.h) continue;
            const float u = __fdiv_rn((float)i, sigma);
            const float v = __fdiv_rn((float)j, sigma);
            const float r2 = __fadd_rn(__fmul_rn(u, u), __fmul_rn(v, v));
            const float w = dm_expf(__fdiv_rn(-r2, expDen));
            const float2 gm = __ldg(g + (size_t)sy * o.pitch + sx);
            const float t = __fdiv_rn(gm.x, kTau);
            int bin = (int)roundf(__fmul_rn(t, (float)kOriBins));
            if (bin < 0) bin += kOriBins;
            if (bin >= kOriBins) bin -= kOriBins;
            hist[bin * 32 + lane] += __fmul_rn(w, gm.y);
        }
        __syncwarp();
        // reduce the 32 lane-private copies of each bin, rotated start → conflict-free
        float* h0 = sH[wid][0];
        float* h1 = sH[wid][1];
        for (int b = lane; b < kOriBins; b += 32) {
            float s = 0.0f;
            for (int l = 0; l < 32; l++) s += hist[b * 32 + ((l + lane) & 31)];
            h0[b] = s;
        }
        __syncwarp();
        // smoothHistogram (:67-85)
        for (int it = 0; it < P.oriSmoothIterations; it++) {
            for (int b = lane; b < kOriBins; b += 32) {
                const float a = h0[(b - 1 + kOriBins) % kOriBins];
                const float c = h0[b];
                const float d = h0[(b + 1) % kOriBins];
                h1[b] = __fdiv_rn(__fadd_rn(__fadd_rn(a, c), d), 3.0f);
            }
            __syncwarp();
            float* tswap = h0; h0 = h1; h1 = tswap;
        }
        // getPrincipalOrientations (:31-64)
        float mx = fmaxf(h0[lane], lane < kOriBins - 32 ? h0[32 + lane] : -2147483648.0f);
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, d));
        const float threshold = __fmul_rn(P.oriThreshold, mx);
        int count = 0;
#pragma unroll
        for (int pass = 0; pass < 2; pass++) {
            const int b = pass * 32 + lane;
            bool peak = false;
            float orientation = 0.0f;
            if (b < kOriBins) {
                const float hm = h0[(b - 1 + kOriBins) % kOriBins];
                const float hc = h0[b];
                const float hp = h0[(b + 1) % kOriBins];
                if ((hc > threshold) && (hc > hm) && (hc > hp)) {
                    peak = true;
                    const float den = __fmul_rn(2.0f, __fsub_rn(__fadd_rn(hm, hp), __fmul_rn(2.0f, hc)));
                    const float offset = __fdiv_rn(__fsub_rn(hm, hp), den);
                    const float t = __fdiv_rn(__fadd_rn((float)b, offset), (float)kOriBins);
                    orientation = __fmul_rn(t, kTau);
                    if (orientation < 0) orientation = __fadd_rn(orientation, kTau);
                    if (orientation >= kTau) orientation = __fsub_rn(orientation, kTau);
                }
            }
            const uint32_t m = __ballot_sync(0xffffffffu, peak);
            if (peak) oriTmp[(size_t)k * kOriBins + count + __popc(m & ((1u << lane) - 1))] = orientation;
            count += __popc(m);
        }
        if (lane == 0) nOri[k] = count;
        __syncwarp();
    }
}

struct OriCount {
    const int* nOri;
    const Counters* counters;
    __device__ int operator()(int i) const { return i < counters->nKeypoints ? nOri[i] : 0; }
};

// Phase C for orientation counts: exclusive offsets per keypoint (+ one past the end).
__global__ void __launch_bounds__(kScanThreads)
oriOffsetsKernel(const int* __restrict__ nOri, const Counters* __restrict__ counters,
                 const int* __restrict__ blockOffsets, int* __restrict__ oriOffset) {
    __shared__ int sh[9];
    const int n = counters->nKeypoints;
    const int i0 = blockIdx.x * kScanChunk + threadIdx.x * kScanItemsPerThread;
    int v[kScanItemsPerThread];
    int s = 0;
#pragma unroll
    for (int k = 0; k < kScanItemsPerThread; k++) {
        v[k] = (i0 + k) < n ? nOri[i0 + k] : 0;
        s += v[k];
    }
    int total;
    int pos = blockOffsets[blockIdx.x] + blockExclusiveScan256(s, sh, &total);
#pragma unroll
    for (int k = 0; k < kScanItemsPerThread; k++) {
        if ((i0 + k) <= n) oriOffset[i0 + k] = pos;
        pos += v[k];
    }
}

// descriptor segment starts from keypoint segment starts
__global__ void descSegmentStartsKernel(const int* __restrict__ segKpStart,
                                        const int* __restrict__ oriOffset,
                                        int* __restrict__ segDescStart, int nSegs) {
    const int seg = blockIdx.x * blockDim.x + threadIdx.x;
    if (seg > nSegs) return;
    segDescStart[seg] = oriOffset[segKpStart[seg]];
}

// ------------------------------------------------------------------------------------------
constexpr int kDescWarps = 4;  // 4 warps x 16 KB of lane-private histograms = 64 KB per CTA

__global__ void __launch_bounds__(kDescWarps * 32)
descriptorKernel(const __grid_constant__ EngineParams P, const SiftKeypoint* __restrict__ kps,
                 const int* __restrict__ kpSeg, const int* __restrict__ segKpStart,
                 const Counters* __restrict__ counters, const int* __restrict__ oriOffset,
                 const float* __restrict__ oriTmp, SiftDescriptor* __restrict__ desc,
                 int capacity) {
    extern __shared__ __align__(16) float sDesc[];  // [warp][128 bins][32 lanes]
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    float* hist = sDesc + wid * (128 * 32);
    const int nKp = counters->nKeypoints;
    int nDesc = oriOffset[nKp];
    if (nDesc > capacity) nDesc = capacity;
    for (int d = blockIdx.x * kDescWarps + wid; d < nDesc; d += gridDim.x * kDescWarps) {
        // keypoint owning descriptor d: last k with oriOffset[k] <= d
        int lo = 0, hi = nKp;
        while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if (oriOffset[mid] <= d) lo = mid;
            else hi = mid;
        }
        const int k = lo;
        const SiftKeypoint kp = kps[k];
        const int seg = kpSeg[k];
        const int frame = seg / kOctaves;
        const float theta = oriTmp[(size_t)k * kOriBins + (d - oriOffset[k])];
        const OctaveDev& o = P.oct[kp.octave];
        const float2* __restrict__ g =
            o.grad + ((size_t)frame * kScales + (kp.scale - 1)) * o.plane;

        // SIFTDescriptor.metal:137-166, inputs truncated as SIFTOctave.swift:417-418
        const float px = __fdiv_rn((float)(int)kp.absoluteX, o.delta);
        const float py = __fdiv_rn((float)(int)kp.absoluteY, o.delta);
        float sinT, cosT;
        dm_sincosf(theta, &sinT, &cosT);
        const float binsPerRadian = __fdiv_rn(8.0f, kTau);
        const float interval = __fadd_rn((float)kp.scale, kp.subScale);
        const float scale = __fmul_rn(1.6f, dm_exp2f(__fdiv_rn(interval, 3.0f)));
        const float hw = __fmul_rn(3.0f, scale);
        const int radius = (int)__fadd_rn(
            __fmul_rn(__fmul_rn(__fmul_rn(hw, sqrtf(2.0f)), 5.0f), 0.5f), 0.5f);
        const int side = 2 * radius + 1;

#pragma unroll 8
        for (int b = 0; b < 128; b++) hist[b * 32 + lane] = 0.0f;

        for (int idx = lane; idx < side * side; idx += 32) {
            const int jj = idx / side;
            const int j = jj - radius;             // x offset (outer loop of the reference)
            const int i = idx - jj * side - radius;  // y offset
            const float fj = (float)j, fi = (float)i;
            const float rx = __fdiv_rn(__fsub_rn(__fmul_rn(fj, cosT), __fmul_rn(fi, sinT)), hw);
            const float ry = __fdiv_rn(__fadd_rn(__fmul_rn(fj, sinT), __fmul_rn(fi, cosT)), hw);
            const float bx = __fsub_rn(__fadd_rn(rx, 2.0f), 0.5f);
            const float by = __fsub_rn(__fadd_rn(ry, 2.0f), 0.5f);
            // addValue drops cells outside [0, 4): nothing lands unless -1 < b < 4 on both axes
            if (!(bx > -1.0f && bx < 4.0f && by > -1.0f && by < 4.0f)) continue;
            const float cxf = __fadd_rn(px, fj), cyf = __fadd_rn(py, fi);
            if (cxf < 0.0f || cyf < 0.0f) continue;
            const int sx = (int)cxf, sy = (int)cyf;
            if (sx >= o.w || sy >= o.h) continue;
            const float2 gm = __ldg(g + (size_t)sy * o.pitch + sx);
            float orientation = __fsub_rn(gm.x, theta);
            while (orientation < 0.0f) orientation = __fadd_rn(orientation, kTau);
            while (orientation >= kTau) orientation = __fsub_rn(orientation, kTau);
            const float bin = __fmul_rn(orientation, binsPerRadian);
            const float en = __fadd_rn(__fmul_rn(rx, rx), __fmul_rn(ry, ry));
            const float w = dm_expf(__fdiv_rn(-en, 8.0f));
            const float value = __fmul_rn(gm.y, w);

            // addFeature (:82-117): trilinear spread over floor/ceil cells and bins
            const float flx = floorf(bx), fly = floorf(by), flb = floorf(bin);
            const int x0 = (int)flx, x1 = (int)ceilf(bx);
            const int y0 = (int)fly, y1 = (int)ceilf(by);
            int b0 = (int)flb, b1 = (int)ceilf(bin);
            if (b0 >= 8) b0 -= 8;
            if (b1 >= 8) b1 -= 8;
            const float iMax = __fsub_rn(bx, flx), iMin = __fsub_rn(1.0f, iMax);
            const float jMax = __fsub_rn(by, fly), jMin = __fsub_rn(1.0f, jMax);
            const float bMax = __fsub_rn(bin, flb), bMin = __fsub_rn(1.0f, bMax);
            const bool vx0 = (x0 >= 0) && (x0 < 4), vx1 = (x1 >= 0) && (x1 < 4);
            const bool vy0 = (y0 >= 0) && (y0 < 4), vy1 = (y1 >= 0) && (y1 < 4);
#define SIFT_ADD(cx, cy, cb, wx, wy, wb)                                                     \
    hist[(((cy) * 4 + (cx)) * 8 + (cb)) * 32 + lane] +=                                      \
        __fmul_rn(__fmul_rn(__fmul_rn(wx, wy), wb), value)
            if (vx0 && vy0) { SIFT_ADD(x0, y0, b0, iMin, jMin, bMin); SIFT_ADD(x0, y0, b1, iMin, jMin, bMax); }
            if (vx1 && vy0) { SIFT_ADD(x1, y0, b0, iMax, jMin, bMin); SIFT_ADD(x1, y0, b1, iMax, jMin, bMax); }
            if (vx1 && vy1) { SIFT_ADD(x1, y1, b0, iMax, jMax, bMin); SIFT_ADD(x1, y1, b1, iMax, jMax, bMax); }
            if (vx0 && vy1) { SIFT_ADD(x0, y1, b0, iMin, jMax, bMin); SIFT_ADD(x0, y1, b1, iMin, jMax, bMax); }
#undef SIFT_ADD
        }
        __syncwarp();
        // reduce lane-private copies: lane owns bins lane, lane+32, lane+64, lane+96
        float f[4];
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const int b = q * 32 + lane;
            float s = 0.0f;
            for (int l = 0; l < 32; l++) s += hist[b * 32 + ((l + lane) & 31)];
            f[q] = s;
        }
        // normalize → clip 0.2 → normalize → quantize (SIFTDescriptor.metal:15-50, 227-230)
#pragma unroll
        for (int rep = 0; rep < 2; rep++) {
            float m = ((f[0] * f[0] + f[1] * f[1]) + f[2] * f[2]) + f[3] * f[3];
#pragma unroll
            for (int dd = 16; dd > 0; dd >>= 1) m += __shfl_xor_sync(0xffffffffu, m, dd);
            if (m != 0.0f) {
                const float inv = __fdiv_rn(1.0f, sqrtf(m));
#pragma unroll
                for (int q = 0; q < 4; q++) f[q] = __fmul_rn(f[q], inv);
            }
            if (rep == 0) {
#pragma unroll
                for (int q = 0; q < 4; q++) f[q] = fminf(f[q], 0.2f);
            }
        }
        __syncwarp();
        uint8_t* bytes = reinterpret_cast<uint8_t*>(hist);  // reuse as 128-byte staging
#pragma unroll
        for (int q = 0; q < 4; q++)
            bytes[q * 32 + lane] = (uint8_t)(int)fminf(255.0f, __fmul_rn(f[q], 512.0f));
        __syncwarp();
        SiftDescriptor* out = desc + d;
        uint32_t* dst = reinterpret_cast<uint32_t*>(out->features);
        dst[lane] = reinterpret_cast<const uint32_t*>(bytes)[lane];
        if (lane == 0) {
            out->keypoint = k - segKpStart[frame * kOctaves];
            out->theta = theta;
        }
        __syncwarp();
    }
}

cudaError_t launchDescribe(const EngineParams& P, const SiftKeypoint* kps, const int* kpSeg,
                           int capKeypoints, const int* segKpStart, int* nOri, float* oriTmp,
                           int* oriOffset, int* blockSums, SiftDescriptor* desc,
                           int capDescriptors, int* segDescStart, int nSegs, Counters* counters,
                           int smCount, cudaStream_t st, cudaEvent_t afterOrientation) {
    orientationKernel<<<smCount * 4, kOriWarps * 32, 0, st>>>(P, kps, kpSeg, counters, nOri, oriTmp);
    SIFT_CUDA_TRY(cudaGetLastError());
    const int nBlocks = (capKeypoints + 1 + kScanChunk - 1) / kScanChunk;
    OriCount v{nOri, counters};
    scanBlockSumsKernel<<<nBlocks, kScanThreads, 0, st>>>(v, blockSums);
    SIFT_CUDA_TRY(cudaGetLastError());
    SIFT_CUDA_TRY(launchScanOffsets(blockSums, nBlocks, &counters->nDescriptors, capDescriptors,
                                    &counters->overflow, 4, st));
    oriOffsetsKernel<<<nBlocks, kScanThreads, 0, st>>>(nOri, counters, blockSums, oriOffset);
    SIFT_CUDA_TRY(cudaGetLastError());
    descSegmentStartsKernel<<<(nSegs + 1 + 127) / 128, 128, 0, st>>>(segKpStart, oriOffset,
                                                                   segDescStart, nSegs);
    SIFT_CUDA_TRY(cudaGetLastError());
    if (afterOrientation) SIFT_CUDA_TRY(cudaEventRecord(afterOrientation, st));

    static unsigned long long configured = 0;
    const int smemBytes = kDescWarps * 128 * 32 * (int)sizeof(float);
    int dev = 0;
    cudaGetDevice(&dev);
    if (!((configured >> (dev & 63)) & 1ull)) {
        SIFT_CUDA_TRY(cudaFuncSetAttribute(descriptorKernel,
                                           cudaFuncAttributeMaxDynamicSharedMemorySize, smemBytes));
        configured |= 1ull << (dev & 63);
    }
    descriptorKernel<<<smCount * 3, kDescWarps * 32, smemBytes, st>>>(
        P, kps, kpSeg, segKpStart, counters, oriOffset, oriTmp, desc, capDescriptors);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------
__global__ void mathDebugKernel(int op, const float* __restrict__ a, const float* __restrict__ b,
                                float* __restrict__ out, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float s, c;
    switch (op) {
        case 0: out[i] = dm_expf(a[i]); break;
        case 1: out[i] = dm_atan2f(a[i], b[i]); break;
        case 2: dm_sincosf(a[i], &s, &c); out[i] = s; break;
        case 3: dm_sincosf(a[i], &s, &c); out[i] = c; break;
        case 4: out[i] = dm_exp2f(a[i]); break;
        default: out[i] = 0.0f;
    }
}

cudaError_t launchMathDebug(int op, const float* a, const float* b, float* out, int64_t n,
                            cudaStream_t st) {
    mathDebugKernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(op, a, b, out, n);
    return cudaGetLastError();
}

}  // namespace sift
