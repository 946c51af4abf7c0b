// pyramid.cu — seed image, separable Gaussian series fused with the DoG subtraction, gradient
// field. Hand-written for sm_100a; compiled with -fmad=false: every float expression is
// evaluated exactly as written (the arithmetic spec of DESIGN.md), only explicit fmaf() fuses.
//
// Replaces, with identical per-pixel arithmetic:
//   ConvertSRGBToGrayscale.metal:11-23, BilinearUpScale.metal:12-64       → grayUpsampleKernel
//   Convolution.metal:15-52, ConvolutionSeries.metal:16-53 (X then Y),
//   Subtract.metal:12-21, NearestNeighborDownScale.metal:15-22            → blurKernel
//   SIFTGradient.metal:15-39                                              → gradientKernel
#include <algorithm>
#include <atomic>
#include <cstdlib>

#include "common.cuh"
#include "dev_math.cuh"
#include "tma.cuh"

namespace sift {

// Packed fp32x2 FMA (sm_100 FFMA2): two independent IEEE fused multiply-adds per instruction,
// bit-identical to two fmaf(). Halves the issue slots of the convolution inner loops.
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pack2(float lo, float hi) {
    f32x2 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void unpack2(f32x2 v, float& lo, float& hi) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) {
    f32x2 d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}

// Common.hpp:15-22 symmetrizedCoordinates (floor-mod form, identical for i >= -2l).
__device__ __forceinline__ int symmetrized(int i, int l) {
    int ll = 2 * l;
    i = ((i % ll) + ll) % ll;
    if (i > l - 1) i = ll - 1 - i;
    return i;
}

// ------------------------------------------------------------------------------------------
// Seed stage. byte/255 has only 256 values, so the (exactly rounded) quotients are tabulated once
// per CTA in shared memory instead of IEEE divisions per pixel.
//
// Gray conversion + exact 2x upsample in one pass (w2 = 2W, h2 = 2H: the only ratio the pipeline
// uses). A thread owns gray pixels (2n, 2n + 1) of row m: it converts the 3 x 2 BGRA patch
// (columns 2n .. 2n + 2, rows m, m + 1, mirrored at the edges), writes its two gray pixels and
// the 4 x 2 block of upsampled pixels (columns 4n .. 4n + 3, rows 2m, 2m + 1) that depend on
// nothing else — the same expressions as grayKernel / upsampleKernel with fx, fy in {0, 0.5},
// 0.75 BGRA loads per output pixel instead of a gray plane round trip and 4 gathers.
__global__ void __launch_bounds__(256)
grayUpsample2xKernel(const uint8_t* __restrict__ bgra, int pitchBytes, int64_t frameStrideBytes,
                     float* __restrict__ gray, int W, int H, float* __restrict__ scaled, int pitch2,
                     size_t scaledFrameStride, int mBegin) {
    __shared__ float lut[256];
    lut[threadIdx.x] = (float)threadIdx.x / 255.0f;
    __syncthreads();
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    const int m = mBegin + blockIdx.y, f = blockIdx.z;
    const int c0 = 2 * n;
    if (c0 >= W) return;
    const int w2 = 2 * W;
    const int c1 = min(c0 + 1, W - 1);
    int c2 = c0 + 2;
    if (c2 >= W) c2 = 2 * W - 1 - c2;      // ip = im + 1 mirrored (upsampleKernel)
    c2 = max(c2, 0);
    int mp = m + 1;
    if (mp >= H) mp = 2 * H - 1 - mp;
    const uint8_t* img = bgra + (size_t)f * frameStrideBytes;
    float g[2][3];
#pragma unroll
    for (int r = 0; r < 2; r++) {
        const uint8_t* row = img + (size_t)(r == 0 ? m : mp) * pitchBytes;
        const int cols[3] = {c0, c1, c2};
#pragma unroll
        for (int k = 0; k < 3; k++) {
            const uchar4 p = *reinterpret_cast<const uchar4*>(row + 4 * cols[k]);  // b, g, r, a
            const float bb = lut[p.x], gg = lut[p.y], rr = lut[p.z];
            g[r][k] = ((0.0f + (0.212639005871510f * rr)) + (0.715168678767756f * gg)) +
                      (0.072192315360734f * bb);
        }
    }
    float* grow = gray + ((size_t)f * H + m) * W;
    grow[c0] = g[0][0];
    if (c0 + 1 < W) grow[c0 + 1] = g[0][1];
    // output column 4n + k: im = 2n + (k >> 1), ip = im + 1, fx = (k & 1) / 2
#pragma unroll
    for (int r = 0; r < 2; r++) {
        const float fy = r ? 0.5f : 0.0f;
        float o[4];
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const int im = k >> 1, ip = im + 1;    // indices into g[][]: columns c0, c1, c2
            const float fx = (k & 1) ? 0.5f : 0.0f;
            // c0 = src[jp][ip], c1 = src[jm][ip], c2 = src[jp][im], c3 = src[jm][im]; jm = m, jp = mp
            // when the output column pair belongs to gray column c1 = W - 1 clamped (odd W edge) the
            // values are unused (guarded below)
            const float a = (fy * g[1][ip]) + ((1 - fy) * g[0][ip]);
            const float b = (fy * g[1][im]) + ((1 - fy) * g[0][im]);
            o[k] = (fx * a) + ((1 - fx) * b);
        }
        float* dst = scaled + (size_t)f * scaledFrameStride + (size_t)(2 * m + r) * pitch2 + 4 * n;
        if (4 * n + 3 < w2) {
            *reinterpret_cast<float4*>(dst) = make_float4(o[0], o[1], o[2], o[3]);
        } else {
#pragma unroll
            for (int k = 0; k < 4; k++)
                if (4 * n + k < w2) dst[k] = o[k];
        }
    }
}

// GRAY8 / NV12-luma input (include/siftcuda.h SIFT_INPUT_GRAY8): a gray byte v is converted as
// the BGRA pixel (v, v, v) would be — the same luminance expression on l = v / 255, tabulated per
// CTA — so the planes are bit-identical to the BGRA path on the gray-expanded frame, at 1 byte
// instead of 4 read per pixel. A thread owns gray pixels 4n .. 4n + 3 of row m: it reads the 5 x 2
// byte patch (columns 4n .. 4n + 4, rows m, m + 1, mirrored at the edges) and writes its four
// gray pixels and the 8 x 2 block of upsampled pixels that depend on nothing else.
__global__ void __launch_bounds__(256)
gray8Upsample2xKernel(const uint8_t* __restrict__ luma, int pitchBytes, int64_t frameStrideBytes,
                      float* __restrict__ gray, int W, int H, float* __restrict__ scaled, int pitch2,
                      size_t scaledFrameStride) {
    __shared__ float lut[256];
    {
        const float l = (float)threadIdx.x / 255.0f;
        lut[threadIdx.x] = ((0.0f + (0.212639005871510f * l)) + (0.715168678767756f * l)) +
                           (0.072192315360734f * l);
    }
    __syncthreads();
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    const int m = blockIdx.y, f = blockIdx.z;
    const int c0 = 4 * n;
    if (c0 >= W) return;
    const int w2 = 2 * W;
    int mp = m + 1;
    if (mp >= H) mp = 2 * H - 1 - mp;
    const uint8_t* img = luma + (size_t)f * frameStrideBytes;
    float g[2][5];
#pragma unroll
    for (int r = 0; r < 2; r++) {
        const uint8_t* row = img + (size_t)(r == 0 ? m : mp) * pitchBytes;
#pragma unroll
        for (int k = 0; k < 5; k++) {
            int c = c0 + k;
            if (c >= W) c = 2 * W - 1 - c;      // ip = im + 1 mirrored (BilinearUpScale.metal:36-48)
            c = max(c, 0);
            g[r][k] = lut[__ldg(row + c)];
        }
    }
    float* grow = gray + ((size_t)f * H + m) * W;
#pragma unroll
    for (int k = 0; k < 4; k++)
        if (c0 + k < W) grow[c0 + k] = g[0][k];
#pragma unroll
    for (int r = 0; r < 2; r++) {
        const float fy = r ? 0.5f : 0.0f;
        float o[8];
#pragma unroll
        for (int k = 0; k < 8; k++) {
            const int im = k >> 1, ip = im + 1;
            const float fx = (k & 1) ? 0.5f : 0.0f;
            const float a = (fy * g[1][ip]) + ((1 - fy) * g[0][ip]);
            const float b = (fy * g[1][im]) + ((1 - fy) * g[0][im]);
            o[k] = (fx * a) + ((1 - fx) * b);
        }
        float* dst = scaled + (size_t)f * scaledFrameStride + (size_t)(2 * m + r) * pitch2 + 8 * n;
        if (8 * n + 7 < w2) {
            reinterpret_cast<float4*>(dst)[0] = make_float4(o[0], o[1], o[2], o[3]);
            reinterpret_cast<float4*>(dst)[1] = make_float4(o[4], o[5], o[6], o[7]);
        } else {
#pragma unroll
            for (int k = 0; k < 8; k++)
                if (8 * n + k < w2) dst[k] = o[k];
        }
    }
}

// DifferenceOfGaussians.encodeSeedTexture (:357-389) up to the blur: luminosity + 2x bilinear.
// The pyramid always doubles exactly (w2 = Int(W / 0.5) = 2 W, DifferenceOfGaussians.swift:235-238).
cudaError_t launchGrayUpsample(const uint8_t* pixels, int bytesPerPixel, int pitchBytes,
                               int64_t frameStrideBytes, float* gray, int W, int H, float* scaled,
                               int w2, int h2, int pitch2, size_t scaledFrameStride, int frames,
                               cudaStream_t st) {
    if (w2 != 2 * W || h2 != 2 * H || W < 2 || H < 2) return cudaErrorInvalidValue;
    if (bytesPerPixel == 4) {
        dim3 g((unsigned)(((W + 1) / 2 + 255) / 256), (unsigned)H, (unsigned)frames);
        grayUpsample2xKernel<<<g, 256, 0, st>>>(pixels, pitchBytes, frameStrideBytes, gray, W, H, scaled,
                                                pitch2, scaledFrameStride, 0);
    } else if (bytesPerPixel == 1) {
        dim3 g((unsigned)(((W + 3) / 4 + 255) / 256), (unsigned)H, (unsigned)frames);
        gray8Upsample2xKernel<<<g, 256, 0, st>>>(pixels, pitchBytes, frameStrideBytes, gray, W, H, scaled,
                                                 pitch2, scaledFrameStride);
    } else {
        return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------
// Separable Gaussian blur, X pass then Y pass through shared memory, fused with
//   dog  = out - in            (Subtract.metal)
//   half = out[2y][2x]         (NearestNeighborDownScale.metal, seeds the next octave)
// Per output pixel the accumulation is sum = fma(w[i], c, sum), i ascending, as the oracle.
//
// One CTA per TX x TY output tile (tiles of all frames in one linear grid). The input tile with
// halo (rows TY + 2R, columns TX + 2RP, RP = R rounded up to 4 so that rows stay 16-byte aligned
// with global memory) arrives by cp.async in NG row groups, each its own commit group, and the
// X pass of group g starts as soon as that group has landed — the remaining groups' HBM/L2
// latency hides behind it (a CTA is otherwise bound by the latency of its own tile load). A
// second buffer holds the X-pass result (rows TY + 2R, columns TX). Row pitches are 4 * odd
// floats: 8 lanes on 8 consecutive rows issuing LDS.128 / STS.128 hit 8 distinct 4-bank groups,
// so the row-per-lane X pass is conflict-free; the Y pass reads float2 columns with consecutive
// lanes on consecutive pairs, also conflict-free.
// (Measured alternatives — persistent CTAs with a double-buffered input, march-down strips, a
// strip-streaming form with register-resident Y accumulators (profiles/experiments/),
// 128x64 tiles — were all slower on B200; see profiles/r1/SUMMARY.md.)
template <int NTAPS, int TX, int TY, int NT>
struct BlurCfg {
    static constexpr int RY = TX * TY / (2 * NT);   // Y pass: 2 columns x RY rows per thread
    static constexpr int R = NTAPS / 2;
    static constexpr int RP = (R + 3) / 4 * 4;
    static constexpr int IN_W = TX + 2 * RP;
    static constexpr int IN_H = TY + 2 * R;
    static constexpr int IP = (IN_W % 8 == 4) ? IN_W : IN_W + 4;  // IN_W is a multiple of 4
    static constexpr int TP = (TX % 8 == 4) ? TX : TX + 4;
    static constexpr int SMEM_FLOATS = IN_H * IP + IN_H * TP;
    static constexpr int XSEG = 8;                   // outputs per thread in the X pass
    static constexpr int NG = 3;                     // row groups of the pipelined tile load
    // rows per group: a multiple of 8, so that every group starts on a 128-byte boundary of the
    // tile (row pitch = 16 bytes x odd) — the alignment a TMA destination needs
    static constexpr int GH = (IN_H + 8 * NG - 1) / (8 * NG) * 8;
    static constexpr int GH_LAST = IN_H - (NG - 1) * GH;   // rows of the last group
    static constexpr int SMEM_BYTES = SMEM_FLOATS * 4 + NG * 8;   // + one mbarrier per group
};

template <int N>
__device__ __forceinline__ void cpAsyncWaitGroup() {
    asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory");
}

template <int NTAPS, int TX, int TY, int NT, bool DOG, bool HALF>
__global__ void __launch_bounds__(NT)
blurKernel(const BlurArgs a, const __grid_constant__ Taps taps, const __grid_constant__ CUtensorMap mapFull,
           const __grid_constant__ CUtensorMap mapLast) {
    using C = BlurCfg<NTAPS, TX, TY, NT>;
    constexpr int R = C::R, RP = C::RP, IN_W = C::IN_W, IN_H = C::IN_H, IP = C::IP, TP = C::TP;
    constexpr int NG = C::NG, GH = C::GH;
    extern __shared__ __align__(128) float smem[];
    float* const sIn = smem;
    float* const sTmp = smem + IN_H * IP;
    uint64_t* const bars = reinterpret_cast<uint64_t*>(smem + C::SMEM_FLOATS);   // [NG], TMA path only

    const int tid = threadIdx.x;
    if (a.tma && tid == 0) {
#pragma unroll
        for (int g = 0; g < NG; g++) mbarInit(&bars[g], 1);
        mbarInitFence();
    }
    pdlPrologue();   // small planes are launched behind their predecessor (a.pdl)
    const int w = a.w, h = a.h, pitch = a.pitch;
    // output rows [yB, yE): the whole plane, or one row band of it (bands of one plane run as
    // independent launches on separate streams; the mirror boundary still refers to the plane)
    const int yB = a.yBegin, yE = a.yEnd > 0 ? a.yEnd : h;
    const int tilesX = (w + TX - 1) / TX, tilesY = (yE - yB + TY - 1) / TY;
    const int tilesPerFrame = tilesX * tilesY;
    const int tile = blockIdx.x;
    const int f = tile / tilesPerFrame;
    const int tr = tile - f * tilesPerFrame;
    const int ty = tr / tilesX, tx = tr - ty * tilesX;
    const int x0 = tx * TX, y0 = yB + ty * TY;
    const float* __restrict__ in = a.in + (size_t)f * a.inFrameStride;

    // ---- asynchronous load of the input tile with halo, NG row groups ----------------------
    const bool interior = (x0 - RP >= 0) && (x0 + TX + RP <= w) && (y0 - R >= 0) &&
                          (y0 + TY + R <= h);
    // Interior tiles (the bulk of a large plane): the tile + halo is three TMA box copies
    // (cp.async.bulk.tensor.3d: x, y, slice), issued by one thread, each completing on the mbarrier
    // of its row group — no per-thread address arithmetic, no LDGSTS issue slots. The box is IP
    // floats wide, i.e. it writes the padded row pitch directly (the 4 surplus columns are never
    // read). Edge tiles need the mirror boundary per element and keep the cp.async path.
    const bool useTma = a.tma && interior;   // CTA-uniform
    if (a.tma) __syncthreads();              // barrier initialisation visible before anyone waits
    if (useTma) {
        if (tid == 0) {
            const int z = a.tmaZ0 + f * a.tmaZStride;
#pragma unroll
            for (int g = 0; g < NG; g++) {
                const int rows = g + 1 < NG ? GH : C::GH_LAST;
                mbarExpectTx(&bars[g], (uint32_t)(rows * IP * 4));
                tmaLoad3d(g + 1 < NG ? &mapFull : &mapLast, &bars[g], sIn + g * GH * IP, x0 - RP, y0 - R + g * GH, z);
            }
        }
    } else
#pragma unroll
    for (int g = 0; g < NG; g++) {
        const int rBeg = g * GH, rEnd = min(rBeg + GH, IN_H);
        if (interior) {
            // LPR lanes per tile row (>= the row's 16-byte vectors), whole rows per warp: row and
            // column of a lane are fixed, the loop only adds constants to two addresses
            constexpr int V = IN_W / 4;
            constexpr int LPR = V > 16 ? 32 : (V > 8 ? 16 : 8);
            constexpr int RSTEP = (NT / 32) * (32 / LPR);      // rows per CTA iteration
            const int lane = tid & 31, wid = tid >> 5;
            const int c4 = lane % LPR;
            const int r0 = rBeg + wid * (32 / LPR) + lane / LPR;
            if (c4 < V) {
                const float* gp = in + (size_t)(y0 - R + r0) * pitch + (x0 - RP) + 4 * c4;
                unsigned sa = (unsigned)__cvta_generic_to_shared(sIn + r0 * IP + 4 * c4);
                const size_t gstep = (size_t)RSTEP * pitch;
                for (int r = r0; r < rEnd; r += RSTEP) {
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(sa), "l"(gp));
                    gp += gstep;
                    sa += RSTEP * IP * 4;
                }
            }
        } else {
            // edge tile: mirror boundary. One warp per row (row index reflected once per warp),
            // lanes across columns with a single-reflection fast path.
            const int lane = tid & 31, wid = tid >> 5;
            for (int r = rBeg + wid; r < rEnd; r += NT / 32) {
                const int gy = symmetrized(y0 - R + r, h);
                const float* __restrict__ srow = in + (size_t)gy * pitch;
                for (int c = lane; c < IN_W; c += 32) {
                    int gx = x0 - RP + c;
                    if (gx < 0) gx = -1 - gx;
                    else if (gx >= w) gx = 2 * w - 1 - gx;
                    if (gx < 0 || gx >= w) gx = symmetrized(x0 - RP + c, w);
                    const unsigned s = (unsigned)__cvta_generic_to_shared(sIn + r * IP + c);
                    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(s), "l"(srow + gx));
                }
            }
        }
        asm volatile("cp.async.commit_group;\n" ::: "memory");
    }

    if (a.debugMode & 4) {   // tuning: tile load only
        if (useTma) {
#pragma unroll
            for (int g = 0; g < NG; g++) mbarWait(&bars[g], 0);
        }
        cpAsyncWaitGroup<0>();
        __syncthreads();
        if (sIn[tid] == 1.2345e30f) a.out[0] = 1.0f;
        return;
    }
    // ---- X pass, group by group: one row, 8 consecutive outputs per task --------------------
#pragma unroll
    for (int g = 0; g < NG; g++) {
        if (useTma) {
            mbarWait(&bars[g], 0);   // the group's bytes have landed (and are visible to the waiter)
        } else {
            if (g == 0) cpAsyncWaitGroup<NG - 1>();
            else if (g == 1) cpAsyncWaitGroup<NG - 2>();
            else cpAsyncWaitGroup<0>();
            __syncthreads();
        }
        const int rBeg = g * GH, rows = min(rBeg + GH, IN_H) - rBeg;
        constexpr int SEGS = TX / C::XSEG;
        constexpr int NV = (C::XSEG + 2 * RP) / 4;
        for (int t = tid; t < rows * SEGS; t += NT) {
            const int seg = t / rows, r = rBeg + (t - seg * rows);
            const float4* src = reinterpret_cast<const float4*>(sIn + r * IP + seg * C::XSEG);
            float v[NV * 4];
#pragma unroll
            for (int k = 0; k < NV; k++) {
                const float4 q = src[k];
                v[4 * k + 0] = q.x; v[4 * k + 1] = q.y; v[4 * k + 2] = q.z; v[4 * k + 3] = q.w;
            }
            float acc[C::XSEG];
            if (!(a.debugMode & 1)) {
                // Packed FFMA2 on output pairs (2p, 2p + 1): tap i of the pair reads the input pair
                // starting at m = 2p + (RP - R) + i. Walking m upwards, every input pair is formed
                // once (free when m is even: the halves sit in an aligned register pair of the
                // LDS.128; two moves when m is odd), feeds the up to four output pairs that need it
                // and dies; each output still accumulates its taps in ascending order, and a
                // packed FMA is two independent IEEE FMAs — bit-identical to the scalar chain.
                constexpr int OFF = RP - R;
                f32x2 acc2[C::XSEG / 2];
#pragma unroll
                for (int p = 0; p < C::XSEG / 2; p++) acc2[p] = pack2(0.0f, 0.0f);
#pragma unroll
                for (int m = OFF; m < OFF + NTAPS + C::XSEG - 2; m++) {
                    const f32x2 pm = pack2(v[m], v[m + 1]);
#pragma unroll
                    for (int p = 0; p < C::XSEG / 2; p++) {
                        const int i = m - OFF - 2 * p;
                        if (i >= 0 && i < NTAPS) acc2[p] = fma2(pack2(taps.w[i], taps.w[i]), pm, acc2[p]);
                    }
                }
#pragma unroll
                for (int p = 0; p < C::XSEG / 2; p++) unpack2(acc2[p], acc[2 * p], acc[2 * p + 1]);
            } else {
#pragma unroll
                for (int k = 0; k < C::XSEG; k++) acc[k] = v[k + RP];
            }
            float4* dst = reinterpret_cast<float4*>(sTmp + r * TP + seg * C::XSEG);
            dst[0] = make_float4(acc[0], acc[1], acc[2], acc[3]);
            dst[1] = make_float4(acc[4], acc[5], acc[6], acc[7]);
        }
    }
    __syncthreads();

    // ---- Y pass: 2 adjacent columns x RY rows per thread, streaming over X-pass rows -------
    // Row k of the X-pass result feeds output row q with tap i = k - q, so walking k upwards
    // accumulates every output in ascending tap order (the spec's order) while only the RY
    // accumulator pairs and one float2 of input are live: (RY + 2R) LDS.64 per 2 RY outputs.
    // Packed fp32x2 FMAs (FFMA2) on the column pair.
    {
        constexpr int RY = C::RY;
        constexpr int CGS = TX / 2;
        static_assert(CGS * (TY / RY) == NT, "one Y-pass task per thread");
        float* __restrict__ out = a.out + (size_t)f * a.outFrameStride;
        float* __restrict__ dog = DOG ? a.dog + (size_t)f * a.dogFrameStride : nullptr;
        float* __restrict__ half = HALF ? a.half + (size_t)f * a.halfFrameStride : nullptr;
        const int yb = tid / CGS, cg = tid - yb * CGS;
        const int gx = x0 + 2 * cg;
        f32x2 acc2[RY];
#pragma unroll
        for (int q = 0; q < RY; q++) acc2[q] = pack2(0.0f, 0.0f);
        if (!(a.debugMode & 1)) {
#pragma unroll
            for (int k = 0; k < RY + 2 * R; k++) {
                const float2 v = *reinterpret_cast<const float2*>(sTmp + (yb * RY + k) * TP + 2 * cg);
                const f32x2 v2 = pack2(v.x, v.y);
#pragma unroll
                for (int q = 0; q < RY; q++) {
                    const int i = k - q;
                    if (i >= 0 && i < NTAPS) acc2[q] = fma2(pack2(taps.w[i], taps.w[i]), v2, acc2[q]);
                }
            }
        } else {
#pragma unroll
            for (int q = 0; q < RY; q++) {
                const float2 v = *reinterpret_cast<const float2*>(sTmp + (yb * RY + q + R) * TP + 2 * cg);
                acc2[q] = pack2(v.x, v.y);
            }
        }
        const bool full = (x0 + TX <= w) && (y0 + TY <= yE);   // CTA-uniform
        const bool noStore = (a.debugMode & 2) != 0;
        const int gy0 = y0 + yb * RY;
        const float* cptr = sIn + (yb * RY + R) * IP + RP + 2 * cg;   // centre pixels of the DoG
        if (full && !noStore) {
            // interior tile: two running row pointers, no per-row index arithmetic or guards
            float* po = out + (size_t)gy0 * pitch + gx;
            float* pd = DOG ? dog + (size_t)gy0 * pitch + gx : nullptr;
#pragma unroll
            for (int q = 0; q < RY; q++) {
                float2 r;
                unpack2(acc2[q], r.x, r.y);
                *reinterpret_cast<float2*>(po) = r;
                if (DOG) {
                    const float2 c = *reinterpret_cast<const float2*>(cptr + q * IP);
                    *reinterpret_cast<float2*>(pd) = make_float2(r.x - c.x, r.y - c.y);
                    pd += pitch;
                }
                if (HALF && ((gy0 + q) & 1) == 0 && ((gy0 + q) >> 1) < a.halfH && (gx >> 1) < a.halfW)
                    half[(size_t)((gy0 + q) >> 1) * a.halfPitch + (gx >> 1)] = r.x;   // gx is even
                po += pitch;
            }
        } else {
#pragma unroll
            for (int q = 0; q < RY; q++) {
                float2 r;
                unpack2(acc2[q], r.x, r.y);
                const int gy = gy0 + q;
                const size_t o = (size_t)gy * pitch + gx;
                float2 d2 = make_float2(0.f, 0.f);
                if (DOG) {
                    const float2 c = *reinterpret_cast<const float2*>(cptr + q * IP);
                    d2 = make_float2(r.x - c.x, r.y - c.y);
                }
                if (noStore) {
                    if (r.x == 1.2345e30f) out[o] = d2.x;   // keeps the computation alive
                } else if (gy < yE) {
                    if (gx < w) { out[o] = r.x; if (DOG) dog[o] = d2.x; }
                    if (gx + 1 < w) { out[o + 1] = r.y; if (DOG) dog[o + 1] = d2.y; }
                }
                if (HALF && (gy & 1) == 0 && (gy >> 1) < a.halfH && gy < yE && (gx >> 1) < a.halfW && gx < w)
                    half[(size_t)(gy >> 1) * a.halfPitch + (gx >> 1)] = r.x;   // gx is even
            }
        }
    }
}

constexpr int tapIndex(int ntaps) { return ntaps == 11 ? 0 : ntaps == 15 ? 1 : ntaps == 17 ? 2 : ntaps == 21 ? 3 : 4; }

template <int NTAPS, int TX, int TY, int NT, bool DOG, bool HALF>
static cudaError_t launchBlurCfg(const BlurArgs& a0, const Taps& taps, cudaStream_t st) {
    using C = BlurCfg<NTAPS, TX, TY, NT>;
    static_assert(C::IN_W % 4 == 0 && C::IP % 8 == 4 && C::TP % 8 == 4, "bank layout");
    static_assert(TX % C::XSEG == 0 && C::RY >= 1 && TY % C::RY == 0 && C::NG == 3, "tile shape");
    static_assert(C::GH % 8 == 0 && C::GH_LAST >= 1 && C::GH_LAST <= C::GH && C::SMEM_FLOATS % 2 == 0, "row groups");
    const int smemBytes = C::SMEM_BYTES;
    // per-device bit (the attribute is per device); contexts on separate host threads may race
    // here, so the word is atomic — setting the attribute twice is harmless, losing a bit is not
    static std::atomic<unsigned long long> configured{0};
    int dev = 0;
    cudaGetDevice(&dev);
    if (!((configured.load(std::memory_order_acquire) >> (dev & 63)) & 1ull)) {
        SIFT_CUDA_TRY(cudaFuncSetAttribute(blurKernel<NTAPS, TX, TY, NT, DOG, HALF>,
                                           cudaFuncAttributeMaxDynamicSharedMemorySize, smemBytes));
        configured.fetch_or(1ull << (dev & 63), std::memory_order_release);
    }
    BlurArgs a = a0;
    static const bool tmaEnabled = !(getenv("SIFTCUDA_BLUR_TMA") && atoi(getenv("SIFTCUDA_BLUR_TMA")) == 0);   // tuning switch
    const BlurTmaSet* set = a.tmaSet;
    // planes narrower or lower than one tile + halo have no interior tile at all
    a.tma = (tmaEnabled && set && set->valid && a.w >= TX + 2 * C::RP && a.h >= TY + 2 * C::R) ? 1 : 0;
    static const CUtensorMap noMap{};
    const CUtensorMap& mFull = a.tma ? set->map[tapIndex(NTAPS)][TX == 64 ? 0 : 1][0] : noMap;
    const CUtensorMap& mLast = a.tma ? set->map[tapIndex(NTAPS)][TX == 64 ? 0 : 1][1] : noMap;
    const int rows = (a.yEnd > 0 ? a.yEnd : a.h) - a.yBegin;
    const long nTiles = (long)((a.w + TX - 1) / TX) * ((rows + TY - 1) / TY) * a.frames;
    if (nTiles > 2147483647L || rows < 1) return cudaErrorInvalidValue;
    return launchKernel(blurKernel<NTAPS, TX, TY, NT, DOG, HALF>, dim3((unsigned)nTiles), dim3(NT), (size_t)smemBytes,
                        st, a.pdl != 0, a.priority ? a.priority : kNoPriority, a, taps, mFull, mLast);
}

// Tensor maps of one stack of planes [nz][h][pitch] for every blur configuration: box = (IP
// columns, GH or GH_LAST rows, 1 slice). Columns beyond the plane's pitch / rows beyond h are
// never requested by interior tiles (out-of-bounds elements would be zero-filled).
template <int NTAPS, int TX, int TY, int NT>
static cudaError_t makeBlurMaps(BlurTmaSet* set, const float* base, int pitch, int h, int nz, size_t planeFloats) {
    using C = BlurCfg<NTAPS, TX, TY, NT>;
    auto enc = tensorMapEncoder();
    if (!enc) return cudaErrorNotSupported;
    const cuuint64_t dims[3] = {(cuuint64_t)pitch, (cuuint64_t)h, (cuuint64_t)nz};
    const cuuint64_t strides[2] = {(cuuint64_t)pitch * 4, (cuuint64_t)planeFloats * 4};
    const cuuint32_t elem[3] = {1, 1, 1};
    for (int kind = 0; kind < 2; kind++) {
        const cuuint32_t box[3] = {(cuuint32_t)C::IP, (cuuint32_t)(kind == 0 ? C::GH : C::GH_LAST), 1};
        const CUresult r = enc(&set->map[tapIndex(NTAPS)][TX == 64 ? 0 : 1][kind], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3,
                               const_cast<float*>(base), dims, strides, box, elem, CU_TENSOR_MAP_INTERLEAVE_NONE,
                               CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return cudaErrorInvalidValue;
    }
    return cudaSuccess;
}

cudaError_t makeBlurTmaSet(BlurTmaSet* set, const float* base, int pitch, int h, int nz, size_t planeFloats) {
    set->valid = 0;
    if (pitch < 32 || h < 1 || nz < 1) return cudaSuccess;
#define SIFT_MAPS(T)                                                                     \
    SIFT_CUDA_TRY((makeBlurMaps<T, 64, 64, 256>(set, base, pitch, h, nz, planeFloats))); \
    SIFT_CUDA_TRY((makeBlurMaps<T, 32, 32, 128>(set, base, pitch, h, nz, planeFloats)));
    SIFT_MAPS(11) SIFT_MAPS(15) SIFT_MAPS(17) SIFT_MAPS(21) SIFT_MAPS(27)
#undef SIFT_MAPS
    set->valid = 1;
    return cudaSuccess;
}

template <int NTAPS, int TX, int TY, int NT>
static cudaError_t launchBlurFlags(const BlurArgs& a, const Taps& taps, cudaStream_t st) {
    if (a.dog && a.half) return launchBlurCfg<NTAPS, TX, TY, NT, true, true>(a, taps, st);
    if (a.dog) return launchBlurCfg<NTAPS, TX, TY, NT, true, false>(a, taps, st);
    if (a.half) return cudaErrorInvalidValue;
    return launchBlurCfg<NTAPS, TX, TY, NT, false, false>(a, taps, st);
}

// Large planes: 64x64 tiles, 256 threads (least halo overhead). Planes that would not give every
// SM two such tiles: 32x32 tiles, 128 threads — a launch is then bound by the latency of one
// tile, so smaller tiles on more SMs finish sooner.
template <int NTAPS>
static cudaError_t launchBlurT(const BlurArgs& a, const Taps& taps, cudaStream_t st) {
    const int rows = (a.yEnd > 0 ? a.yEnd : a.h) - a.yBegin;
    const long tiles64 = (long)((a.w + 63) / 64) * ((rows + 63) / 64) * a.frames;
    if (tiles64 >= 2 * 148) return launchBlurFlags<NTAPS, 64, 64, 256>(a, taps, st);
    return launchBlurFlags<NTAPS, 32, 32, 128>(a, taps, st);
}

cudaError_t launchBlur(const BlurArgs& a, const Taps& taps, int ntaps, cudaStream_t st) {
    switch (ntaps) {
        case 11: return launchBlurT<11>(a, taps, st);
        case 15: return launchBlurT<15>(a, taps, st);
        case 17: return launchBlurT<17>(a, taps, st);
        case 21: return launchBlurT<21>(a, taps, st);
        case 27: return launchBlurT<27>(a, taps, st);
        default: return cudaErrorInvalidValue;
    }
}

// ------------------------------------------------------------------------------------------
// SIFTGradient.metal:15-39 for Gaussian slices 1..3 (the only ones ever read downstream:
// refined scale is in [1, 3], SIFTInterpolate.metal:187-189). 4 pixels per thread, float4
// loads of the three rows, two float4 stores of (orientation, magnitude) pairs. The mirror
// boundary of symmetrizedCoordinates reduces to a clamp for offsets of one pixel.
__global__ void __launch_bounds__(256) gradientKernel(const OctaveDev o, int yBegin, int yEnd) {
    // 4 pixels x 2 rows per thread: rows y - 1 .. y + 2 are loaded once for both output rows
    const int x0 = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
    const int y = yBegin + blockIdx.y * 2;   // rows [yBegin, yEnd), yBegin even
    const int s = blockIdx.z % kScales;   // 0..2 → Gaussian slice s + 1
    const int f = blockIdx.z / kScales;
    if (x0 >= o.w) return;
    const float* __restrict__ g = o.G + ((size_t)f * kGaussians + (s + 1)) * o.plane;
    const int hLast = o.h - 1;
    // symmetrized(-1) = 0, symmetrized(h) = h - 1: the mirror reduces to a clamp for one pixel
    const int yr[4] = {max(y - 1, 0), y, min(y + 1, hLast), min(y + 2, hLast)};
    // rows are padded to a multiple of 32 floats, so the float4 reads stay inside the row
    float4 r4[4];
    float lf[2], rt[2];
#pragma unroll
    for (int k = 0; k < 4; k++) r4[k] = __ldg(reinterpret_cast<const float4*>(g + (size_t)yr[k] * o.pitch + x0));
    const int xl = x0 > 0 ? x0 - 1 : 0, xr = x0 + 4 < o.w ? x0 + 4 : o.w - 1;
#pragma unroll
    for (int k = 0; k < 2; k++) {
        lf[k] = __ldg(g + (size_t)yr[1 + k] * o.pitch + xl);
        rt[k] = __ldg(g + (size_t)yr[1 + k] * o.pitch + xr);
    }
    const float v[4][4] = {{r4[0].x, r4[0].y, r4[0].z, r4[0].w}, {r4[1].x, r4[1].y, r4[1].z, r4[1].w},
                           {r4[2].x, r4[2].y, r4[2].z, r4[2].w}, {r4[3].x, r4[3].y, r4[3].z, r4[3].w}};
#pragma unroll
    for (int row = 0; row < 2; row++) {
        if (y + row >= yEnd) break;
        const float* c = v[1 + row];
        const float* up = v[row];
        // the row below output row y + row; at the bottom edge the clamp makes it the row itself
        const float* dn = v[2 + row];
        float r[8];
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const int x = x0 + k;
            const float cmx = (k == 0) ? lf[row] : c[k - 1];
            float cpx = (k == 3) ? rt[row] : c[k + 1];
            if (x + 1 >= o.w) cpx = c[k];              // symmetrized(w) = w - 1 (this pixel)
            const float tx = (cpx - cmx) * 0.5f;
            const float ty = (dn[k] - up[k]) * 0.5f;
            r[2 * k] = dm_atan2f(tx, ty);
            r[2 * k + 1] = sqrtf((tx * tx) + (ty * ty));
        }
        float2* dst = o.grad + ((size_t)f * kScales + s) * o.plane + (size_t)(y + row) * o.pitch + x0;
        float4* d4 = reinterpret_cast<float4*>(dst);
        d4[0] = make_float4(r[0], r[1], r[2], r[3]);
        d4[1] = make_float4(r[4], r[5], r[6], r[7]);
    }
}

cudaError_t launchGradient(const OctaveDev& o, int frames, cudaStream_t st, int yBegin, int yEnd, int priority) {
    if (yEnd <= 0) { yBegin = 0; yEnd = o.h; }
    yBegin &= ~1;
    yEnd = std::min(yEnd, o.h);
    if (yEnd <= yBegin) return cudaSuccess;
    dim3 grid((o.w + 1023) / 1024, (yEnd - yBegin + 1) / 2, kScales * frames);
    return launchKernel(gradientKernel, grid, dim3(256), 0, st, false, priority, o, yBegin, yEnd);
}

// ------------------------------------------------------------------------------------------
// Tail octaves. The deepest octaves (a couple of thousand pixels each) are pure latency when run
// as ~7 dependent launches per octave: at 1080p octaves 3 - 6 end the pyramid stage 47 us after the
// large octaves are done. Here ONE CTA per frame keeps a whole octave in shared memory — the
// current and the next Gaussian plane (rows extended by their mirrored halo), the X-pass plane
// (extended by mirrored halo rows) and a ring of three DoG planes — and walks all five scales of
// every tail octave in one launch: blur X / Y (the same ascending fma chain; the mirror boundary
// is materialised once per pass instead of per tap), DoG, the decimated seed of the next octave,
// the gradient field of slices 1 - 3 and the extrema mask of scales 1 - 3, with block barriers
// where the launches were. Every plane is still written to global memory (refinement,
// orientation, descriptor and the debug taps read them). Arithmetic is expression-for-expression
// that of blurKernel / gradientKernel / extremaMaskSmallKernel: planes and candidates stay
// bit-identical. One SM has 1 / 148 of the machine, so only planes small enough that their launch
// chain (not their arithmetic) is what costs are taken: <= kTailMaxPixels.
struct TailArgs {
    OctaveDev oct[kOctaves];
    int oStart;                 // first tail octave; its slice 0 was seeded by the previous octave's blur
    float softThreshold;
    uint32_t* mask;
    int blocksPerFrame;
    int w0, h0;                 // size of octave oStart (the largest tail plane): sizes the shared-memory planes
    Taps taps[kGaussians - 1];
};

constexpr int kTailThreads = 1024;
constexpr int kTailPad = 16;          // halo columns / rows kept around the planes (>= the largest radius, 13)
constexpr int kTailMaxPixels = 2600;

template <int NT>
__device__ __forceinline__ void tailBlurScale(const Taps& taps, float* A, float* B, float* T, float* Ds, int w, int h,
                                              float* __restrict__ Gout, float* __restrict__ Dout, int pitch) {
    constexpr int R = NT / 2;
    const int tid = threadIdx.x;
    const int wE = w + 2 * kTailPad;
    // mirrored halo columns of the input rows (Common.hpp:15-22)
    for (int p = tid; p < h * 2 * R; p += kTailThreads) {
        const int y = p / (2 * R), k = p - y * (2 * R);
        const int x = k < R ? k - R : w + (k - R);
        A[y * wE + kTailPad + x] = A[y * wE + kTailPad + symmetrized(x, w)];
    }
    __syncthreads();
    // X pass into rows kTailPad .. kTailPad + h of T
    for (int p = tid; p < w * h; p += kTailThreads) {
        const int y = p / w, x = p - y * w;
        const float* row = A + y * wE + kTailPad + x - R;
        float sum = 0.0f;
#pragma unroll
        for (int i = 0; i < NT; i++) sum = fmaf(taps.w[i], row[i], sum);
        T[(y + kTailPad) * w + x] = sum;
    }
    __syncthreads();
    // mirrored halo rows of the X-pass plane
    for (int p = tid; p < 2 * R * w; p += kTailThreads) {
        const int k = p / w, x = p - k * w;
        const int y = k < R ? k - R : h + (k - R);
        T[(y + kTailPad) * w + x] = T[(symmetrized(y, h) + kTailPad) * w + x];
    }
    __syncthreads();
    // Y pass, DoG, stores
    for (int p = tid; p < w * h; p += kTailThreads) {
        const int y = p / w, x = p - y * w;
        const float* col = T + (y + kTailPad - R) * w + x;
        float sum = 0.0f;
#pragma unroll
        for (int i = 0; i < NT; i++) sum = fmaf(taps.w[i], col[i * w], sum);
        const float d = sum - A[y * wE + kTailPad + x];
        B[y * wE + kTailPad + x] = sum;
        Ds[p] = d;
        Gout[(size_t)y * pitch + x] = sum;
        Dout[(size_t)y * pitch + x] = d;
    }
    __syncthreads();
}

__global__ void __launch_bounds__(kTailThreads, 1) tailOctavesKernel(const __grid_constant__ TailArgs a) {
    extern __shared__ __align__(16) float tsm[];
    const int PA = a.h0 * (a.w0 + 2 * kTailPad);      // Gaussian planes: rows extended by the halo
    const int PT = (a.h0 + 2 * kTailPad) * a.w0;      // X-pass plane: extended by halo rows
    const int PD = a.w0 * a.h0;
    float* A = tsm;                  // Gaussian slice s
    float* B = tsm + PA;             // Gaussian slice s + 1
    float* T = tsm + 2 * PA;
    float* Dr = tsm + 2 * PA + PT;   // DoG ring: slice s at Dr + (s % 3) * PD
    const int tid = threadIdx.x, lane = tid & 31, f = blockIdx.x;
    pdlPrologue();
    for (int oc = a.oStart; oc < kOctaves; oc++) {
        const OctaveDev& o = a.oct[oc];
        const int w = o.w, h = o.h, n = w * h;
        if (w < 1 || h < 1) break;
        const int wE = w + 2 * kTailPad;
        float* __restrict__ G = o.G + (size_t)f * kGaussians * o.plane;
        float* __restrict__ D = o.D + (size_t)f * kDogs * o.plane;
        // slice 0: seeded by the previous octave's blur (first tail octave) or by this CTA below
        for (int p = tid; p < n; p += kTailThreads) {
            const int y = p / w, x = p - y * w;
            A[y * wE + kTailPad + x] = __ldcg(G + (size_t)y * o.pitch + x);
        }
        __syncthreads();
        for (int s = 0; s < kGaussians - 1; s++) {
            float* Ds = Dr + (s % 3) * PD;
            float* Gout = G + (size_t)(s + 1) * o.plane;
            float* Dout = D + (size_t)s * o.plane;
            switch (s) {   // the fixed schedule of DifferenceOfGaussians.swift:91-110: 11, 15, 17, 21, 27 taps
                case 0: tailBlurScale<11>(a.taps[0], A, B, T, Ds, w, h, Gout, Dout, o.pitch); break;
                case 1: tailBlurScale<15>(a.taps[1], A, B, T, Ds, w, h, Gout, Dout, o.pitch); break;
                case 2: tailBlurScale<17>(a.taps[2], A, B, T, Ds, w, h, Gout, Dout, o.pitch); break;
                case 3: tailBlurScale<21>(a.taps[3], A, B, T, Ds, w, h, Gout, Dout, o.pitch); break;
                default: tailBlurScale<27>(a.taps[4], A, B, T, Ds, w, h, Gout, Dout, o.pitch); break;
            }
            // gradient field of Gaussian slice s + 1 in 1..3 (SIFTGradient.metal:15-39; clamp = the
            // mirror for one pixel, as gradientKernel)
            if (s + 1 <= kScales) {
                float2* __restrict__ grad = o.grad + ((size_t)f * kScales + s) * o.plane;
                for (int p = tid; p < n; p += kTailThreads) {
                    const int y = p / w, x = p - y * w;
                    const float* row = B + y * wE + kTailPad;
                    const float cpx = row[min(x + 1, w - 1)], cmx = row[max(x - 1, 0)];
                    const float dn = B[min(y + 1, h - 1) * wE + kTailPad + x], up = B[max(y - 1, 0) * wE + kTailPad + x];
                    const float tx = (cpx - cmx) * 0.5f;
                    const float ty = (dn - up) * 0.5f;
                    grad[(size_t)y * o.pitch + x] = make_float2(dm_atan2f(tx, ty), sqrtf((tx * tx) + (ty * ty)));
                }
            }
            // decimated seed of the next octave from slice 3 (NearestNeighborDownScale.metal:15-22):
            // to global memory (its plane is a result like any other); re-read, L2-hot, when this
            // octave is done
            if (s + 1 == kScales && oc + 1 < kOctaves && a.oct[oc + 1].w >= 1 && a.oct[oc + 1].h >= 1) {
                const OctaveDev& nx = a.oct[oc + 1];
                float* __restrict__ G0 = nx.G + (size_t)f * kGaussians * nx.plane;
                for (int p = tid; p < nx.w * nx.h; p += kTailThreads) {
                    const int y = p / nx.w, x = p - y * nx.w;
                    G0[(size_t)y * nx.pitch + x] = B[(2 * y) * wE + kTailPad + 2 * x];
                }
            }
            // extrema mask of scale sc = s - 1 in 1..3 from DoG slices sc - 1, sc, sc + 1 = s
            // (SIFTExtrema.metal:62-110: neighbours 1..25, neighbour 0 = (-1, -1, -1) skipped; the
            // 0.8 C_DoG pre-threshold of SIFTInterpolate.metal:208 fused in, as extremaMaskSmallKernel)
            const int sc = s - 1;
            if (sc >= 1 && sc <= kScales && w >= 3 && h >= 3) {
                const float* Dm = Dr + ((sc - 1) % 3) * PD;
                const float* Dc = Dr + (sc % 3) * PD;
                const float* Dp = Dr + ((sc + 1) % 3) * PD;
                uint32_t* __restrict__ m = a.mask + ((size_t)f * a.blocksPerFrame + o.maskBlockStart) * (size_t)kScanChunk;
                const int units = h * o.maskRowWords;   // one warp per (row, mask word)
                for (int u = tid >> 5; u < units; u += kTailThreads / 32) {
                    const int y = u / o.maskRowWords, xw = u - y * o.maskRowWords;
                    const int x = xw * 32 + lane;
                    bool cand = false;
                    if (x >= 1 && x <= w - 2 && y >= 1 && y <= h - 2) {
                        const int c = y * w + x;
                        const float v = Dc[c];
                        if (!(fabsf(v) <= a.softThreshold)) {
                            float mn = +1000.0f, mx = -1000.0f;
#pragma unroll
                            for (int ds = -1; ds <= 1; ds++) {
                                const float* pl = ds < 0 ? Dm : (ds == 0 ? Dc : Dp);
#pragma unroll
                                for (int dy = -1; dy <= 1; dy++) {
#pragma unroll
                                    for (int dx = -1; dx <= 1; dx++) {
                                        if (ds == 0 && dy == 0 && dx == 0) continue;       // the centre
                                        if (ds == -1 && dy == -1 && dx == -1) continue;    // neighbour 0
                                        const float nv = pl[c + dy * w + dx];
                                        mn = fminf(mn, nv);
                                        mx = fmaxf(mx, nv);
                                    }
                                }
                            }
                            cand = (v < mn) || (v > mx);
                        }
                    }
                    const uint32_t word = __ballot_sync(0xffffffffu, cand);
                    if (lane == 0) m[((size_t)(sc - 1) * h + y) * o.maskRowWords + xw] = word;
                }
            }
            __syncthreads();
            float* t = A; A = B; B = t;   // slice s + 1 becomes the input of the next scale
        }
        __syncthreads();   // the seed of the next octave (global stores of this CTA) is read back above
    }
}

// First octave the tail kernel takes: the largest one of at most kTailMaxPixels pixels, never
// octave 0; kOctaves = none.
int tailStartOctave(const EngineParams& P) {
    static const bool enabled = !(getenv("SIFTCUDA_TAIL") && atoi(getenv("SIFTCUDA_TAIL")) == 0);   // tuning switch
    if (!enabled) return kOctaves;
    static const int maxPixels = getenv("SIFTCUDA_TAIL_PIXELS") ? atoi(getenv("SIFTCUDA_TAIL_PIXELS")) : kTailMaxPixels;
    for (int o = 1; o < kOctaves; o++) {
        const long w = P.oct[o].w, h = P.oct[o].h;
        const long floats = 2 * h * (w + 2 * kTailPad) + (h + 2 * kTailPad) * w + 3 * w * h;
        if (w >= 1 && h >= 1 && w * h <= maxPixels && floats * 4 <= 200 * 1024) return o;
    }
    return kOctaves;
}

cudaError_t launchTailOctaves(const EngineParams& P, int oStart, const Taps* taps, const int* ntaps, uint32_t* mask,
                              int frames, cudaStream_t st, int priority) {
    if (oStart >= kOctaves) return cudaSuccess;
    TailArgs a{};
    for (int o = 0; o < kOctaves; o++) a.oct[o] = P.oct[o];
    a.oStart = oStart;
    a.softThreshold = P.dogThreshold * 0.8f;
    a.mask = mask;
    a.blocksPerFrame = P.blocksPerFrame;
    a.w0 = P.oct[oStart].w;
    a.h0 = P.oct[oStart].h;
    for (int s = 0; s < kGaussians - 1; s++) {
        if (ntaps[s] != (s == 0 ? 11 : s == 1 ? 15 : s == 2 ? 17 : s == 3 ? 21 : 27)) return cudaErrorInvalidValue;
        a.taps[s] = taps[s];
    }
    const int smemBytes = (2 * a.h0 * (a.w0 + 2 * kTailPad) + (a.h0 + 2 * kTailPad) * a.w0 + 3 * a.w0 * a.h0) * (int)sizeof(float);
    static std::atomic<unsigned long long> configured{0};
    int dev = 0;
    cudaGetDevice(&dev);
    if (!((configured.load(std::memory_order_acquire) >> (dev & 63)) & 1ull)) {
        SIFT_CUDA_TRY(cudaFuncSetAttribute(tailOctavesKernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
        configured.fetch_or(1ull << (dev & 63), std::memory_order_release);
    }
    return launchKernel(tailOctavesKernel, dim3((unsigned)frames), dim3(kTailThreads), (size_t)smemBytes, st, false, priority, a);
}

}  // namespace sift

