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
#include <algorithm>
#include <atomic>

#include "common.cuh"
#include "dev_math.cuh"
#include "scan.cuh"

namespace sift {

constexpr float kTau = 6.28318530717958647692f;  // 2 * M_PI_F in float

// Small-integer <-> float conversions on the FMA / ALU pipes (the F2I / I2F instructions run on
// the quarter-rate XU pipe with ~4x the latency). Exact for |v| < 2^22.
constexpr float kMagic = 12582912.0f;          // 1.5 * 2^23: ulp 1, so a directed-rounding add is floor / ceil
constexpr int kMagicBits = 0x4B400000;
__device__ __forceinline__ int floorToInt(float v, float& fl) {
    const float t = __fadd_rd(v, kMagic);
    fl = t - kMagic;
    return __float_as_int(t) - kMagicBits;
}
__device__ __forceinline__ int ceilToInt(float v) { return __float_as_int(__fadd_ru(v, kMagic)) - kMagicBits; }
__device__ __forceinline__ int floorToInt(float v) { return __float_as_int(__fadd_rd(v, kMagic)) - kMagicBits; }
__device__ __forceinline__ float smallIntToFloat(int v) { return __int_as_float(kMagicBits + v) - kMagic; }


// ------------------------------------------------------------------------------------------
constexpr int kOriWarps = 8;      // warps (keypoints in flight) per CTA
constexpr int kOriMaxSide = 96;   // window side 2r+1; r = ceil(4.5 sigma') <= 17 for detected keypoints

__global__ void __launch_bounds__(kOriWarps * 32)
orientationKernel(const __grid_constant__ EngineParams P, const SiftKeypoint* __restrict__ kps,
                  const int* __restrict__ kpSeg, Counters* __restrict__ counters,
                  int* __restrict__ nOri, float* __restrict__ oriTmp) {
    pdlPrologue();
    __shared__ __align__(16) float sHist[kOriWarps][kOriBins * 32];  // [bin][lane] per warp
    __shared__ float sH[kOriWarps][kOriBins];
    __shared__ float sW[kOriWarps][kOriMaxSide];       // separable Gaussian window weights
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int n = counters->nKeypoints;
    float* hist = sHist[wid];
    float* wt = sW[wid];
    // Keypoints differ 2-3x in window size: warps pull their next keypoint from a queue counter
    // (requested before the current one is processed, consumed after) instead of a fixed stride,
    // so no warp is left finishing a long tail alone. Results are indexed by k: order-free.
    const int gridWarps = gridDim.x * kOriWarps;
    int kNext = 0;
    // The queue runs from the end of the list: within an octave the list ascends in scale, i.e. in
    // window area, so the cheapest items (octave 0, scale 1) come last and the tail stays short.
    for (int t = blockIdx.x * kOriWarps + wid; t < n; t = __shfl_sync(0xffffffffu, kNext, 0)) {
        if (lane == 0) kNext = gridWarps + atomicAdd(&counters->oriNext, 1);
        const int k = n - 1 - t;
        const SiftKeypoint kp = kps[k];
        const int frame = kpSeg[k] / kOctaves;
        const OctaveDev& o = P.oct[kp.octave];
        // host filter of SIFTOctave.swift:303-329 on the untruncated coordinates (exact sequence)
        const float lambda = P.lambdaOri;
        const float fx = __fdiv_rn(kp.absoluteX, o.delta);
        const float fy = __fdiv_rn(kp.absoluteY, o.delta);
        const float sigma = __fdiv_rn(kp.sigma, o.delta);
        const float rf = ceilf(__fmul_rn(__fmul_rn(3.0f, lambda), sigma));
        const bool reject = (floorf(__fsub_rn(fx, rf)) < 1.0f) ||
                            (ceilf(__fadd_rn(fx, rf)) > (float)(o.w - 2)) ||
                            (floorf(__fsub_rn(fy, rf)) < 1.0f) ||
                            (ceilf(__fadd_rn(fy, rf)) > (float)(o.h - 2));
        const int r = (int)rf;
        const int side = 2 * r + 1;
        if (reject || kp.scale < 1 || kp.scale > kScales || side > kOriMaxSide || !(rf >= 0.0f)) {
            if (lane == 0) nOri[k] = 0;
            continue;
        }
        // kernel inputs are packed with Int32(absoluteCoordinate) (SIFTOctave.swift:333-334)
        const int x = (int)roundf(__fdiv_rn((float)(int)kp.absoluteX, o.delta));
        const int y = (int)roundf(__fdiv_rn((float)(int)kp.absoluteY, o.delta));
        const float expDen = __fmul_rn(__fmul_rn(2.0f, lambda), lambda);
        const float2* __restrict__ g =
            o.grad + ((size_t)frame * kScales + (kp.scale - 1)) * o.plane;

        // w(i, j) = exp(-((i/s)^2 + (j/s)^2) / (2 lambda^2)) evaluated as the product of two 1-D
        // factors (continuous quantity: differs from the oracle's single exp in the last bits)
        for (int t = lane; t < side; t += 32) {
            const float u = __fdiv_rn((float)(t - r), sigma);
            wt[t] = dm_expf(__fdiv_rn(-__fmul_rn(u, u), expDen));
        }
#pragma unroll
        for (int b = 0; b < kOriBins; b++) hist[b * 32 + lane] = 0.0f;
        __syncwarp();
        const float invSide = 1.0f / (float)side;
        const int nSamples = side * side;
        // four gathers in flight per lane: the loop is bound by the latency of the gradient
        // loads, not by arithmetic
        for (int base = 0; base < nSamples; base += 128) {
            float2 gm[4];
            float wgt[4];
            bool ok[4];
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const int idx = base + u * 32 + lane;
                ok[u] = false;
                if (idx < nSamples) {
                    const int jj = floorToInt((smallIntToFloat(idx) + 0.5f) * invSide);
                    const int ii = idx - jj * side;
                    const int sx = x + ii - r, sy = y + jj - r;
                    if (sx >= 0 && sx < o.w && sy >= 0 && sy < o.h) {
                        ok[u] = true;
                        gm[u] = __ldg(g + (size_t)sy * o.pitch + sx);
                        wgt[u] = wt[ii] * wt[jj];
                    }
                }
            }
#pragma unroll
            for (int u = 0; u < 4; u++) {
                if (ok[u]) {
                    // bin index = round((ori / tau) * 36), a discontinuous quantity that must equal
                    // the spec's. One multiply gives it to within 5e-6; only when that estimate
                    // sits within 1e-4 of a rounding boundary (k + 0.5) is the exact sequence
                    // (IEEE divide, multiply, round-half-away) evaluated — same result always.
                    const float est = gm[u].x * ((float)kOriBins / kTau);
                    const float tr = __fadd_rn(est, kMagic);      // round to nearest even = rintf
                    float rb = tr - kMagic;
                    int bin = __float_as_int(tr) - kMagicBits;
                    if (fabsf(fabsf(est - rb) - 0.5f) < 1e-4f) {
                        rb = roundf(__fmul_rn(__fdiv_rn(gm[u].x, kTau), (float)kOriBins));
                        bin = (int)rb;
                    }
                    if (bin < 0) bin += kOriBins;
                    if (bin >= kOriBins) bin -= kOriBins;
                    hist[bin * 32 + lane] += wgt[u] * gm[u].y;
                }
            }
        }
        __syncwarp();
        // reduce the 32 lane-private copies of each bin: 8 lanes per bin, each sums 4 consecutive
        // copies from one LDS.128 (8 lanes x 16 B = one conflict-free 128-byte row), then three
        // butterfly steps; lane group bg handles bins bg, bg + 4, ...
        float* h0 = sH[wid];
        {
            const int sub = lane & 7, bg = lane >> 3;
#pragma unroll
            for (int j = 0; j < kOriBins / 4; j++) {
                const int b = bg + 4 * j;
                const float4 v = *reinterpret_cast<const float4*>(hist + b * 32 + sub * 4);
                float sum = (v.x + v.y) + (v.z + v.w);
                sum += __shfl_xor_sync(0xffffffffu, sum, 1);
                sum += __shfl_xor_sync(0xffffffffu, sum, 2);
                sum += __shfl_xor_sync(0xffffffffu, sum, 4);
                if (sub == 0) h0[b] = sum;
            }
        }
        __syncwarp();
        // smoothHistogram (:67-85): circular box filter, iterated. The 36 bins live in registers
        // (lane l: bin l, and bin 32 + l for l < 4); neighbours come by shuffle, the four wrap
        // positions by broadcast. Same (a + c) + d, IEEE / 3 per bin as the spec.
        {
            float lo = h0[lane];
            float hi = lane < kOriBins - 32 ? h0[32 + lane] : 0.0f;
            for (int it = 0; it < P.oriSmoothIterations; it++) {
                float pl = __shfl_up_sync(0xffffffffu, lo, 1), nl = __shfl_down_sync(0xffffffffu, lo, 1);
                float ph = __shfl_up_sync(0xffffffffu, hi, 1), nh = __shfl_down_sync(0xffffffffu, hi, 1);
                const float b35 = __shfl_sync(0xffffffffu, hi, kOriBins - 33), b32 = __shfl_sync(0xffffffffu, hi, 0);
                const float b31 = __shfl_sync(0xffffffffu, lo, 31), b0 = __shfl_sync(0xffffffffu, lo, 0);
                if (lane == 0) { pl = b35; ph = b31; }
                if (lane == 31) nl = b32;
                if (lane == kOriBins - 33) nh = b0;
                lo = __fdiv_rn(__fadd_rn(__fadd_rn(pl, lo), nl), 3.0f);
                hi = __fdiv_rn(__fadd_rn(__fadd_rn(ph, hi), nh), 3.0f);
            }
            __syncwarp();
            h0[lane] = lo;
            if (lane < kOriBins - 32) h0[32 + lane] = hi;
            __syncwarp();
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

// Phase C for orientation counts: exclusive offsets per keypoint (+ one past the end), the owner
// keypoint of every descriptor, and the descriptor segment starts (the first keypoint of a
// (frame, octave) segment — kpSeg changes there — writes its offset for every segment up to its
// own; the slot one past the end closes the rest).
__global__ void __launch_bounds__(kScanThreads)
oriOffsetsKernel(const int* __restrict__ nOri, const Counters* __restrict__ counters,
                 const int* __restrict__ blockOffsets, int* __restrict__ oriOffset,
                 int* __restrict__ descKp, int capDescriptors, const int* __restrict__ kpSeg,
                 int* __restrict__ segDescStart, int nSegs) {
    pdlPrologue();
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
    int segPrev = (i0 == 0 || i0 > n) ? -1 : kpSeg[i0 - 1];
#pragma unroll
    for (int k = 0; k < kScanItemsPerThread; k++) {
        const int i = i0 + k;
        if (i <= n) {
            oriOffset[i] = pos;
            const int segCur = (i == n) ? nSegs : kpSeg[i];
            for (int t = segPrev + 1; t <= segCur; t++) segDescStart[t] = pos;
            segPrev = segCur;
        }
        for (int t = 0; t < v[k]; t++)   // owner of each descriptor: no search in the descriptor kernel
            if (pos + t < capDescriptors) descKp[pos + t] = i;
        pos += v[k];
    }
}

// ------------------------------------------------------------------------------------------
// Descriptor. One warp per (keypoint, theta). The reference walks the whole (2·radius+1)^2
// window and lets addValue drop what falls outside the 4x4 grid (SIFTDescriptor.metal:53-79);
// only samples inside the rotated square |rx|, |ry| < 2.5 can contribute, so each window row's
// contributing span is computed analytically (widened by < 1 pixel, then culled by the same
// test the reference's cell bounds imply) and the lanes walk the flattened spans densely.
// Window radius, sample coordinates and centre truncation follow the spec's exact sequences;
// per-sample weights use FMA / reciprocal / ex2.approx — continuous quantities within the ±1
// tolerance of the quantised features.
constexpr int kDescCopies = 32;    // lane-private histogram copies (16 KB per warp). Sharing one copy
                                   // between lanes l and l + 16 in two half-warp rounds (8 KB, twice
                                   // the resident warps) measured equal: 472 vs 463 us at 1080p
constexpr int kDescWarps = 7;      // 112 KB of histograms per CTA, two CTAs (14 warps) per SM
constexpr int kDescMaxSide = 128;  // 2·radius+1; radius <= 39 for detected keypoints
constexpr int kDescBins = 128;
constexpr int kDescDefaultParts = 4;

// Trilinear accumulation of one sample into the lane's histogram copy (addFeature,
// SIFTDescriptor.metal:82-117). Two base addresses per sample (one per orientation bin); the
// four cells are immediate offsets; loads / stores of cells outside the 4x4 grid are predicated off.
__device__ __forceinline__ float ex2Approx(float x) {
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

__device__ __forceinline__ void descAccumulate(char* const hl, const float2 gm, const float bx,
                                               const float by, const float r2, const bool ok,
                                               const float theta) {
    // orientation relative to theta in bins: t in (-12, 4); floor and the 3 low bits give the bin
    const float t = (gm.x - theta) * (8.0f / kTau);
    float tfl, bxfl, byfl;
    const int bi = floorToInt(t, tfl);
    const float fb = t - tfl;
    const float val = gm.y * ex2Approx(r2 * (-0.125f * 1.4426950408889634f));   // exp(-r2 / 8)
    const int x0 = floorToInt(bx, bxfl), y0 = floorToInt(by, byfl);   // in [-1, 3] when ok
    const float fxw = bx - bxfl, fyw = by - byfl;
    // ceil = floor + 1 except on exact integers, where the reference adds a zero weight to the
    // floor cell — same sums either way. Each split is one product and one difference.
    const float vx1 = val * fxw, vx0 = val - vx1;
    const float v01 = vx0 * fyw, v00 = vx0 - v01;
    const float v11 = vx1 * fyw, v10 = vx1 - v11;
    const bool okx0 = ok && (x0 >= 0), okx1 = ok && (x0 < 3);
    const bool oky0 = (y0 >= 0), oky1 = (y0 < 3);
    const bool c00 = okx0 && oky0, c10 = okx1 && oky0, c01 = okx0 && oky1, c11 = okx1 && oky1;
    const int cellB = (y0 * 4 + x0) * (8 * kDescCopies * 4);        // bytes: cell (x0, y0), bin 0
    float* const p0 = reinterpret_cast<float*>(hl + cellB + (bi & 7) * (kDescCopies * 4));
    float* const p1 = reinterpret_cast<float*>(hl + cellB + ((bi + 1) & 7) * (kDescCopies * 4));
    constexpr int DX = 8 * kDescCopies, DY = 32 * kDescCopies;      // floats to cell x+1 / y+1
    const float g0 = 1.0f - fb;
    // the eight addresses of a lane are distinct and private: load all, add, store all
    float t0, t1, t2, t3, t4, t5, t6, t7;   // each read only under the predicate that loaded it
    if (c00) t0 = p0[0];
    if (c00) t1 = p1[0];
    if (c10) t2 = p0[DX];
    if (c10) t3 = p1[DX];
    if (c01) t4 = p0[DY];
    if (c01) t5 = p1[DY];
    if (c11) t6 = p0[DY + DX];
    if (c11) t7 = p1[DY + DX];
    if (c00) p0[0] = fmaf(v00, g0, t0);
    if (c00) p1[0] = fmaf(v00, fb, t1);
    if (c10) p0[DX] = fmaf(v10, g0, t2);
    if (c10) p1[DX] = fmaf(v10, fb, t3);
    if (c01) p0[DY] = fmaf(v01, g0, t4);
    if (c01) p1[DY] = fmaf(v01, fb, t5);
    if (c11) p0[DY + DX] = fmaf(v11, g0, t6);
    if (c11) p1[DY + DX] = fmaf(v11, fb, t7);
}

// Lanes walk the flattened spans densely in units of 16-byte aligned sample pairs (x fastest,
// stride 32 pairs): one LDG.128 and one row search per two samples.
// (Measured and dropped: one sample per step, 0.43 vs 0.41 ms; the warp as a (32 / C)-row x
// C-column tile over row groups, C = 4, 8, 16 — uniform control flow, but idle lane slots at the
// ragged span ends pay the full accumulation cost: 0.60 / 0.63 / 0.74 ms against 0.46 ms.)
template <int NPAIRS, int PARTS>
__global__ void __launch_bounds__(kDescWarps * 32, 2)
descriptorKernel(const __grid_constant__ EngineParams P, const SiftKeypoint* __restrict__ kps,
                 const int* __restrict__ kpSeg, const int* __restrict__ segKpStart,
                 Counters* __restrict__ counters, const int* __restrict__ oriOffset,
                 const float* __restrict__ oriTmp, const int* __restrict__ descKp,
                 const DescriptorColumnsDev cols, const DescriptorColumnsDev hostCols, int capacity) {
    pdlPrologue();
    extern __shared__ __align__(16) float sDesc[];  // [warp][128 bins][32 copies]
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    float* hist = sDesc + wid * (kDescBins * kDescCopies);
    const int nKp = counters->nKeypoints;
    int nDesc = oriOffset[nKp];
    if (nDesc > capacity) nDesc = capacity;
    constexpr int warpsPerCta = kDescWarps;
    // dynamic work queue, as in the orientation kernel (descriptor windows differ up to 4x in area)
    const int gridWarps = gridDim.x * warpsPerCta;
    int dNext = 0;
    for (int t = blockIdx.x * warpsPerCta + wid; t < nDesc; t = __shfl_sync(0xffffffffu, dNext, 0)) {
        if (lane == 0) dNext = gridWarps + atomicAdd(&counters->descNext, 1);
        const int d = nDesc - 1 - t;   // from the end: largest windows first (see orientationKernel)
        const int k = descKp[d];   // keypoint owning descriptor d
        const SiftKeypoint kp = kps[k];
        const int frame = kpSeg[k] / kOctaves;
        const float theta = oriTmp[(size_t)k * kOriBins + (d - oriOffset[k])];
        const OctaveDev& o = P.oct[kp.octave];
        const float2* __restrict__ g =
            o.grad + ((size_t)frame * kScales + (kp.scale - 1)) * o.plane;

        // SIFTDescriptor.metal:137-166 (exact sequences), inputs truncated as SIFTOctave.swift:417-418
        const float px = __fdiv_rn((float)(int)kp.absoluteX, o.delta);
        const float py = __fdiv_rn((float)(int)kp.absoluteY, o.delta);
        float sinT, cosT;
        dm_sincosf(theta, &sinT, &cosT);
        const float interval = __fadd_rn((float)kp.scale, kp.subScale);
        const float scale = __fmul_rn(1.6f, dm_exp2f(__fdiv_rn(interval, 3.0f)));
        const float hw = __fmul_rn(3.0f, scale);
        int radius = (int)__fadd_rn(
            __fmul_rn(__fmul_rn(__fmul_rn(hw, sqrtf(2.0f)), 5.0f), 0.5f), 0.5f);
        radius = max(0, min(radius, (kDescMaxSide - 1) / 2));
        const float a = cosT / hw, b = sinT / hw;   // rx = j a - i b, ry = j b + i a

        // Contributing span of window row i (y offset): x offsets j with |rx| < 2.5 and |ry| < 2.5,
        //   |j a - i b| < 2.5  ->  j in (i s1 - c1, i s1 + c1),  s1 =  b / a, c1 = 2.5 / |a|
        //   |j b + i a| < 2.5  ->  j in (i s2 - c2, i s2 + c2),  s2 = -a / b, c2 = 2.5 / |b|
        // intersected with the window and the plane. Both lines are linear in i, so every lane
        // derives the bounds of the row it is in with two FMAs per side — no tables, no dependent
        // shared-memory look-ups. (|a|, |b| are floored at 1e-6 for the span only: it stays a
        // superset; the exact cull below uses the true a, b.) Lanes walk x fastest, so a warp's
        // gathers are contiguous in memory.
        const float ae = copysignf(fmaxf(fabsf(a), 1e-6f), a), be = copysignf(fmaxf(fabsf(b), 1e-6f), b);
        const float s1 = b / ae, c1 = 2.5f / fabsf(ae);
        const float s2 = -a / be, c2 = 2.5f / fabsf(be);
        // sample (x, y) = (trunc(px + j), trunc(py + i)) = (ipx + j, ipy + i) wherever px + j >= 0:
        // px, py are multiples of 1/delta and small, so the float sums of the spec are exact
        const int ipx = (int)floorf(px), ipy = (int)floorf(py);
        const float xlo = fmaxf(-(float)radius, (float)(-ipx)), xhi = fminf((float)radius, (float)(o.w - 1 - ipx));
        // window rows that can meet the rotated square: |i| <= 2.5 hw (|sin| + |cos|) (+ 1 of slack);
        // the window itself is +-radius = 2.5 sqrt(2) hw, so up to 30 % of its rows are empty
        const int iExt = (int)(2.5f * hw * (fabsf(sinT) + fabsf(cosT))) + 2;
        const int iMin = max(max(-radius, -iExt), -ipy);
        const int iMax = min(min(radius, iExt), o.h - 1 - ipy);

#pragma unroll 8
        for (int bb = 0; bb < kDescBins * kDescCopies / 128; bb++)     // 32 x STS.128 per lane
            reinterpret_cast<float4*>(hist)[bb * 32 + lane] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        __syncwarp();
        char* const hl = reinterpret_cast<char*>(hist + lane);

        if constexpr (PARTS > 0) {
            // Walk by UNITS: unit u = part (u % PARTS) of window row (u / PARTS); a lane takes unit
            // lane + 32 k and walks its run of 16-byte aligned sample pairs along the row, one pair
            // per iteration with the next pair's gather already in flight. The row bounds are
            // evaluated once per unit and never inside the walk (the flattened walk below re-derives
            // them, ~26 instructions, every time a lane crosses a row: ~1.6 times per pair), and the
            // rotated coordinates advance by a multiply-add per pair. The 32 lanes of a step cover
            // 32 / PARTS neighbouring rows of almost equal length, so few lane slots idle.
            const int xLast = o.w - 1;
            const int nUnits = (iMax - iMin + 1) * PARTS;
            const float a2 = a + a, b2 = b + b;
            for (int u0 = 0; u0 < nUnits; u0 += 32) {
                const int u = u0 + lane;
                int len = 0, x = 0;
                float rx0 = 0.0f, ry0 = 0.0f;
                const float2* __restrict__ gp = g;
                if (u < nUnits) {
                    const int row = u / PARTS, part = u - row * PARTS;
                    const int i = iMin + row;
                    const float fi = smallIntToFloat(i);
                    const float lo = fmaxf(fmaxf(fmaf(fi, s1, -c1), fmaf(fi, s2, -c2)), xlo);
                    const float hi = fminf(fminf(fmaf(fi, s1, c1), fmaf(fi, s2, c2)), xhi);
                    const int xs = (ipx + max(ceilToInt(lo - 1.0f), -ipx)) & ~1;
                    const int xe = min(ipx + floorToInt(hi + 1.0f), xLast);
                    const int np = max((xe - xs + 2) >> 1, 0);
                    const int seg = (np + PARTS - 1) / PARTS;
                    const int p0 = part * seg;
                    len = max(min(p0 + seg, np) - p0, 0);
                    x = xs + 2 * p0;
                    const float fj = smallIntToFloat(x - ipx);
                    rx0 = fj * a - fi * b;
                    ry0 = fj * b + fi * a;
                    gp = g + (size_t)(ipy + i) * o.pitch + x;
                }
                // two pairs per iteration, the gathers of the following two already in flight (a
                // distance of ~300 issue slots of this warp: the gradient planes mostly come from DRAM)
                auto gather = [&](int k) -> float4 {
                    return k < len ? __ldg(reinterpret_cast<const float4*>(gp + 2 * k)) : make_float4(0.f, 0.f, 0.f, 0.f);
                };
                auto pairAt = [&](const float4 v, int k, float kf) {
                    const bool live = k < len;
                    const float rx = fmaf(kf, a2, rx0), ry = fmaf(kf, b2, ry0);
                    const float rx1 = rx + a, ry1 = ry + b;
                    const bool ok0 = live && fabsf(rx) < 2.5f && fabsf(ry) < 2.5f;
                    const bool ok1 = live && fabsf(rx1) < 2.5f && fabsf(ry1) < 2.5f && (x + 2 * k < xLast);
                    descAccumulate(hl, make_float2(v.x, v.y), rx + 1.5f, ry + 1.5f, rx * rx + ry * ry, ok0, theta);
                    descAccumulate(hl, make_float2(v.z, v.w), rx1 + 1.5f, ry1 + 1.5f, rx1 * rx1 + ry1 * ry1, ok1, theta);
                };
                float4 va = gather(0), vb = gather(1);
                float kf = 0.0f;
                for (int k = 0; __any_sync(0xffffffffu, k < len); k += 2) {
                    const float4 vc = gather(k + 2);
                    pairAt(va, k, kf);
                    const float4 vd = gather(k + 3);
                    pairAt(vb, k + 1, kf + 1.0f);
                    kf += 2.0f;
                    va = vc;
                    vb = vd;
                }
            }
        } else {
            // Flattened walk (PARTS == 0): lanes stride through the concatenated spans of all rows.
            // Sample PAIRS start at even absolute x. Rows are padded to whole pairs; the padding
            // samples fail the exact cull (or the plane check) below.
            int i = iMin, xs = 0, np = 0, q = lane;
            float fi = smallIntToFloat(iMin);
            const float2* __restrict__ grow = g + (size_t)(ipy + iMin) * o.pitch;
            const int xLast = o.w - 1;
            auto rowBounds = [&]() {
                const float lo = fmaxf(fmaxf(fmaf(fi, s1, -c1), fmaf(fi, s2, -c2)), xlo);
                const float hi = fminf(fminf(fmaf(fi, s1, c1), fmaf(fi, s2, c2)), xhi);
                xs = (ipx + max(ceilToInt(lo - 1.0f), -ipx)) & ~1;
                const int xe = min(ipx + floorToInt(hi + 1.0f), xLast);
                np = max((xe - xs + 2) >> 1, 0);
            };
            auto settle = [&]() {
                while (i <= iMax && q >= np) {
                    q -= np;
                    i++;
                    fi += 1.0f;
                    grow += o.pitch;   // dereferenced only while i <= iMax
                    rowBounds();
                }
            };
            rowBounds();
            settle();
            constexpr int NP = NPAIRS;   // pairs per lane per batch (measured: 1: 0.379, 2: 0.350, 3: 0.341, 4: 0.351 ms)
            // One batch = coordinates + gathers of NP pairs per lane, then their accumulation.
            auto fetch = [&](float4 (&gm)[NP], float (&rxs)[NP], float (&rys)[NP], bool (&ok0)[NP],
                             bool (&ok1)[NP]) -> bool {
                const bool any = __any_sync(0xffffffffu, i <= iMax);
#pragma unroll
                for (int u = 0; u < NP; u++) {
                    const int x = xs + 2 * q;
                    const float fj = smallIntToFloat(x - ipx);
                    const float rx = fj * a - fi * b;
                    const float ry = fj * b + fi * a;
                    const bool rowOk = i <= iMax;
                    ok0[u] = rowOk && fabsf(rx) < 2.5f && fabsf(ry) < 2.5f;
                    ok1[u] = rowOk && fabsf(rx + a) < 2.5f && fabsf(ry + b) < 2.5f && (x < xLast);
                    gm[u] = __ldg(reinterpret_cast<const float4*>((ok0[u] || ok1[u]) ? grow + x : g));
                    rxs[u] = rx;
                    rys[u] = ry;
                    q += 32;
                    settle();
                }
                return any;
            };
            auto accumulate = [&](const float4 (&gm)[NP], const float (&rxs)[NP], const float (&rys)[NP],
                                  const bool (&ok0)[NP], const bool (&ok1)[NP]) {
#pragma unroll
                for (int u = 0; u < NP; u++) {
                    const float rx = rxs[u], ry = rys[u], rx1 = rx + a, ry1 = ry + b;
                    descAccumulate(hl, make_float2(gm[u].x, gm[u].y), rx + 1.5f, ry + 1.5f,
                                   rx * rx + ry * ry, ok0[u], theta);
                    descAccumulate(hl, make_float2(gm[u].z, gm[u].w), rx1 + 1.5f, ry1 + 1.5f,
                                   rx1 * rx1 + ry1 * ry1, ok1[u], theta);
                }
            };
            // (issuing the gathers of batch n + 1 ahead of the accumulation of batch n, with two
            // register sets, was measured: no change — the gather latency is already covered)
            float4 gm[NP];
            float rxs[NP], rys[NP];
            bool ok0[NP], ok1[NP];
            while (fetch(gm, rxs, rys, ok0, ok1)) accumulate(gm, rxs, rys, ok0, ok1);
        }
        __syncwarp();
        // reduce lane-private copies: lane owns bins lane, lane+32, lane+64, lane+96
        // (8 x LDS.128 per bin; the start chunk is rotated by the lane so that a quarter-warp's eight
        // 16-byte reads cover all 32 banks)
        float f[4];
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const float4* row = reinterpret_cast<const float4*>(hist + (q * 32 + lane) * kDescCopies);
            float s0 = 0.0f, s1 = 0.0f, s2 = 0.0f, s3 = 0.0f;
#pragma unroll
            for (int l = 0; l < kDescCopies / 4; l++) {
                const float4 v = row[(l + lane) & (kDescCopies / 4 - 1)];
                s0 += v.x; s1 += v.y; s2 += v.z; s3 += v.w;
            }
            f[q] = (s0 + s1) + (s2 + s3);
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
        // result columns: the 128 feature bytes as one coalesced 128-byte row of the dense
        // [n][128] matrix (the operand layout of the matcher), theta and the keypoint's index in its
        // frame beside it. hostCols (pinned host memory, when set) receives the same: those stores
        // leave over PCIe as whole rows while the kernel keeps running — they are the D2H.
        const uint32_t word = reinterpret_cast<const uint32_t*>(bytes)[lane];
        const int kIndex = k - segKpStart[frame * kOctaves];
        reinterpret_cast<uint32_t*>(cols.features + (size_t)d * kDescBins)[lane] = word;
        if (lane == 0) { cols.theta[d] = theta; cols.keypoint[d] = kIndex; }
        if (hostCols.features) {
            reinterpret_cast<uint32_t*>(hostCols.features + (size_t)d * kDescBins)[lane] = word;
            if (lane == 0) { hostCols.theta[d] = theta; hostCols.keypoint[d] = kIndex; }
        }
        __syncwarp();
    }
}

cudaError_t launchDescribe(const EngineParams& P, const SiftKeypoint* kps, const int* kpSeg,
                           int capKeypoints, const int* segKpStart, int* nOri, float* oriTmp,
                           int* oriOffset, int* descKp, int* blockSums,
                           const DescriptorColumnsDev& cols, const DescriptorColumnsDev& hostCols,
                           int capDescriptors, int* segDescStart, int nSegs, Counters* counters,
                           int smCount, cudaStream_t st, cudaEvent_t afterOrientation) {
    SIFT_CUDA_TRY(pdlLaunch(orientationKernel, dim3(smCount * 4), dim3(kOriWarps * 32), 0, st, true, P, kps, kpSeg,
                            counters, nOri, oriTmp));
    const int nBlocks = (capKeypoints + 1 + kScanChunk - 1) / kScanChunk;
    OriCount v{nOri, counters};
    SIFT_CUDA_TRY(pdlLaunch(scanBlockSumsKernel<OriCount>, dim3(nBlocks), dim3(kScanThreads), 0, st, true, v,
                            blockSums));
    SIFT_CUDA_TRY(launchScanOffsets(blockSums, nBlocks, &counters->nDescriptors, capDescriptors,
                                    &counters->overflow, 4, st));
    SIFT_CUDA_TRY(pdlLaunch(oriOffsetsKernel, dim3(nBlocks), dim3(kScanThreads), 0, st, true, (const int*)nOri,
                            (const Counters*)counters, (const int*)blockSums, oriOffset, descKp, capDescriptors,
                            kpSeg, segDescStart, nSegs));
    if (afterOrientation) SIFT_CUDA_TRY(cudaEventRecord(afterOrientation, st));

    const int smemBytes = kDescWarps * kDescBins * kDescCopies * (int)sizeof(float);
    // two CTAs (14 warps) per SM use all of its shared memory; SIFTCUDA_DESC_CTAS=1 leaves half of it
    // to kernels of another context running beside this one (two contexts per device, frames alternating)
    static const int ctasEnv = getenv("SIFTCUDA_DESC_CTAS") ? atoi(getenv("SIFTCUDA_DESC_CTAS")) : 0;
    const int ctasPerSm = ctasEnv > 0 ? std::min(ctasEnv, 2) : (228 * 1024) / (smemBytes + 1024);
    // walk variant: units of 1/PARTS of a window row (default) or the flattened span walk (0)
    static const int parts = getenv("SIFTCUDA_DESC_PARTS") ? atoi(getenv("SIFTCUDA_DESC_PARTS")) : kDescDefaultParts;
    auto launch = [&](auto kernel, int slot) -> cudaError_t {
        static std::atomic<unsigned long long> configured[5];   // per-device bit per variant, as in launchBlurCfg
        int dev = 0;
        cudaGetDevice(&dev);
        if (!((configured[slot].load(std::memory_order_acquire) >> (dev & 63)) & 1ull)) {
            SIFT_CUDA_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smemBytes));
            configured[slot].fetch_or(1ull << (dev & 63), std::memory_order_release);
        }
        return pdlLaunch(kernel, dim3(smCount * ctasPerSm), dim3(kDescWarps * 32), (size_t)smemBytes, st, true, P, kps,
                         kpSeg, segKpStart, counters, (const int*)oriOffset, (const float*)oriTmp,
                         (const int*)descKp, cols, hostCols, capDescriptors);
    };
    switch (parts) {
        case 0: return launch(descriptorKernel<3, 0>, 0);
        case 2: return launch(descriptorKernel<3, 2>, 2);
        case 3: return launch(descriptorKernel<3, 3>, 3);
        case 8: return launch(descriptorKernel<3, 8>, 1);
        default: return launch(descriptorKernel<3, 4>, 4);
    }
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
