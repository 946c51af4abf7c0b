// detect.cu — 3x3x3 extremum detection with the contrast pre-threshold, deterministic
// compaction, sub-pixel refinement with contrast / edge rejection. Compiled with -fmad=false:
// every float expression is evaluated exactly as written (arithmetic spec).
//
// Replaces, with identical arithmetic:
//   SIFTExtrema.metal:62-110 siftExtremaList (+ SIFTExtremaListKernel.swift:37-69)  → extremaMaskKernel
//   SIFTOctave.getKeypoints (SIFTOctave.swift:198-203)                              → scan + scatter
//   SIFTInterpolate.metal:17-300 siftInterpolate and helpers, Common.hpp:34-47 invert,
//   SIFTOctave.interpolateKeypoints (SIFTOctave.swift:205-288)                      → refineKernel
#include <algorithm>
#include <atomic>

#include "common.cuh"
#include "dev_math.cuh"
#include "scan.cuh"
#include "tma.cuh"

namespace sift {

// ------------------------------------------------------------------------------------------
// Extrema → bitmask. A warp owns 32 consecutive x (one mask word) and marches down kExtRows
// output rows holding a 3-row window of all 5 DoG slices in registers: per new row and slice
// one coalesced load + two shuffles give (left, centre, right); row-wise partial minima /
// maxima are kept so every 3x3 block needs 2-3 more min/max. Candidate iff strictly below the
// minimum or above the maximum of neighbours 1..25 of the reference's table
// (SIFTExtrema.metal:15-45; neighbour 0 = (-1,-1,-1) is skipped at :84 — hence the `cr`
// partials that leave the left pixel of the row above in the slice below out) AND
// |v| > 0.8·C_DoG (first test of siftInterpolate, SIFTInterpolate.metal:208; fusing it here is
// result-neutral). min/max are exact, so evaluation order is free.
constexpr int kExtRows = 30;    // output rows per warp for large planes (32 rows loaded)
constexpr int kExtWarps = 8;

struct RowPart {
    float c;          // centre
    float h3n, h3x;   // min / max of (left, centre, right)
    float lrn, lrx;   // min / max of (left, right)
    float crn, crx;   // min / max of (centre, right)
};

// Raw values of one row of the five slices: the lane's own pixel and, for lanes 0 / 31, the
// pixel just outside the warp's 32 columns (other lanes re-read their own pixel: same cache
// line, no divergence). All ten loads of a row are independent and issued back to back.
struct RowRaw {
    float c[kDogs], e[kDogs];
};

// pc[t] points at the lane's pixel of slice t in the row to load; dx = -1 / 0 / +1 floats to the
// edge pixel. Afterwards the pointers move one row down unless that row is the plane's last
// (rows past the end are never consumed, the pointers just stay inside the plane).
__device__ __forceinline__ RowRaw loadRowAdvance(const float* (&pc)[kDogs], int dx, int& row, int hLast,
                                                 int pitch) {
    RowRaw r;
#pragma unroll
    for (int t = 0; t < kDogs; t++) {
        r.c[t] = __ldg(pc[t]);
        r.e[t] = __ldg(pc[t] + dx);
    }
    const int step = row < hLast ? pitch : 0;
    row++;
#pragma unroll
    for (int t = 0; t < kDogs; t++) pc[t] += step;
    return r;
}

__device__ __forceinline__ RowPart makeRowPart(float c, float e, int lane) {
    float l = __shfl_up_sync(0xffffffffu, c, 1);
    float r = __shfl_down_sync(0xffffffffu, c, 1);
    if (lane == 0) l = e;
    if (lane == 31) r = e;
    RowPart p;
    p.c = c;
    p.lrn = fminf(l, r);
    p.lrx = fmaxf(l, r);
    p.crn = fminf(c, r);
    p.crx = fmaxf(c, r);
    p.h3n = fminf(l, p.crn);
    p.h3x = fmaxf(l, p.crx);
    return p;
}

__global__ void __launch_bounds__(kExtWarps * 32)
extremaMaskKernel(const OctaveDev o, float softThreshold, uint32_t* __restrict__ mask,
                  int blocksPerFrame, int rowsPerWarp, int yBegin, int yEnd) {
    const int lane = threadIdx.x & 31;
    const int xw = blockIdx.x * kExtWarps + (threadIdx.x >> 5);
    if (xw >= o.maskRowWords) return;  // whole warp
    const int x = xw * 32 + lane;      // < pitch: readable even beyond w (masked below)
    const int ex = (lane == 0) ? max(x - 1, 0) : ((lane == 31) ? min(x + 1, o.w - 1) : x);
    const int yFirst = yBegin + blockIdx.y * rowsPerWarp;  // first output row of this warp (rows [yBegin, yEnd))
    const int f = blockIdx.z;
    const float* __restrict__ D = o.D + (size_t)f * kDogs * o.plane;
    uint32_t* __restrict__ m =
        mask + ((size_t)f * blocksPerFrame + o.maskBlockStart) * (size_t)kScanChunk;
    const bool xInside = (x >= 1) && (x <= o.w - 2);

    RowPart win[kDogs][3];  // [slice][row slot]; slots rotate statically (loop unrolled by 3)
    // running addresses: five slice pointers one row ahead of the last load, three mask indices
    const int hLast = o.h - 1;
    const int dx = ex - x;
    int row = min(yFirst - 1, hLast);
    const float* pc[kDogs];
#pragma unroll
    for (int t = 0; t < kDogs; t++) pc[t] = D + (size_t)t * o.plane + (size_t)row * o.pitch + x;
    uint32_t* mp[kScales];
#pragma unroll
    for (int sc = 0; sc < kScales; sc++) mp[sc] = m + ((size_t)sc * o.h + yFirst) * o.maskRowWords + xw;
    {
        const RowRaw r0 = loadRowAdvance(pc, dx, row, hLast, o.pitch);
        const RowRaw r1 = loadRowAdvance(pc, dx, row, hLast, o.pitch);
#pragma unroll
        for (int t = 0; t < kDogs; t++) {
            win[t][0] = makeRowPart(r0.c[t], r0.e[t], lane);
            win[t][1] = makeRowPart(r1.c[t], r1.e[t], lane);
        }
    }
    // software pipeline, three rows ahead: 30 independent loads in flight per lane (the kernel is
    // bound by memory latency at 16 resident warps per SM, not by issue slots)
    RowRaw ahead0 = loadRowAdvance(pc, dx, row, hLast, o.pitch);
    RowRaw ahead1 = loadRowAdvance(pc, dx, row, hLast, o.pitch);
    RowRaw ahead2 = loadRowAdvance(pc, dx, row, hLast, o.pitch);
    for (int r0 = 0; r0 < rowsPerWarp; r0 += 3) {
#pragma unroll
        for (int k = 0; k < 3; k++) {
            const int y = yFirst + r0 + k;
            const int prev = k % 3, cur = (k + 1) % 3, next = (k + 2) % 3;
            if (y >= yEnd) return;  // warp-uniform
            const RowRaw now = ahead0;
            ahead0 = ahead1;
            ahead1 = ahead2;
            ahead2 = loadRowAdvance(pc, dx, row, hLast, o.pitch);      // row y + 4, consumed three rows later
#pragma unroll
            for (int t = 0; t < kDogs; t++) win[t][next] = makeRowPart(now.c[t], now.e[t], lane);
#pragma unroll
            for (int s = 1; s <= kScales; s++) {
                const float v = win[s][cur].c;
                float mn = fminf(fminf(win[s][prev].h3n, win[s][next].h3n), win[s][cur].lrn);
                float mx = fmaxf(fmaxf(win[s][prev].h3x, win[s][next].h3x), win[s][cur].lrx);
                mn = fminf(mn, fminf(fminf(win[s + 1][prev].h3n, win[s + 1][cur].h3n), win[s + 1][next].h3n));
                mx = fmaxf(mx, fmaxf(fmaxf(win[s + 1][prev].h3x, win[s + 1][cur].h3x), win[s + 1][next].h3x));
                mn = fminf(mn, fminf(fminf(win[s - 1][prev].crn, win[s - 1][cur].h3n), win[s - 1][next].h3n));
                mx = fmaxf(mx, fmaxf(fmaxf(win[s - 1][prev].crx, win[s - 1][cur].h3x), win[s - 1][next].h3x));
                const bool cand = xInside && !(fabsf(v) <= softThreshold) && ((v < mn) || (v > mx));
                const uint32_t word = __ballot_sync(0xffffffffu, cand);
                if (lane == 0) *mp[s - 1] = word;
                mp[s - 1] += o.maskRowWords;
            }
        }
    }
}

// Large planes: the same row march fed by TMA. A CTA owns 7 mask words (224 columns) x a strip of
// rows; a producer warp streams blocks of 3 rows x 5 slices x (224 + 2 x 4 halo) columns into a
// shared-memory ring with cp.async.bulk.tensor.3d (one copy per block, mbarrier full / empty per
// ring slot), the 7 consumer warps read (left, centre, right) of their pixel from the tile and run
// the window logic above. The 30 loads a lane keeps in flight in the kernel above become one TMA
// issue by one thread several blocks ahead, and the loads the window logic waits on are
// shared-memory reads. TMA's zero fill outside the plane is harmless here: border pixels are never
// candidates, and every neighbour of a candidate lies inside the plane.
constexpr int kExtTmaWarps = 7;                       // consumer warps = mask words per CTA
constexpr int kExtTmaCols = kExtTmaWarps * 32 + 8;    // box width: 4 columns of halo on each side
constexpr int kExtTmaRows = 3;                        // rows per block = window rotation period
constexpr int kExtTmaRing = 5;                        // 70 KB per CTA: three CTAs (21 consumer warps) per SM
constexpr int kExtTmaBlockFloats = (kDogs * kExtTmaRows * kExtTmaCols + 31) / 32 * 32;   // 128-byte multiple
constexpr int kExtTmaSmemBytes = kExtTmaRing * kExtTmaBlockFloats * 4 + 2 * kExtTmaRing * 8;

__global__ void __launch_bounds__((kExtTmaWarps + 1) * 32, 3)
extremaMaskTmaKernel(const __grid_constant__ CUtensorMap map, const OctaveDev o, float softThreshold,
                     uint32_t* __restrict__ mask, int blocksPerFrame, int rowsPerCta, int yBegin, int yEnd) {
    extern __shared__ __align__(128) float ring[];
    uint64_t* const full = reinterpret_cast<uint64_t*>(ring + kExtTmaRing * kExtTmaBlockFloats);
    uint64_t* const empty = full + kExtTmaRing;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int f = blockIdx.z;
    const int x0 = blockIdx.x * (kExtTmaWarps * 32);
    const int yA = yBegin + blockIdx.y * rowsPerCta;           // first output row of this strip
    const int yB = min(yA + rowsPerCta, yEnd);                 // one past its last output row
    const int y0 = yA - 1;                                     // first input row
    const int nRowsIn = yB - yA + 2;
    const int nBlocks = (nRowsIn + kExtTmaRows - 1) / kExtTmaRows;
    if (threadIdx.x == 0) {
#pragma unroll
        for (int k = 0; k < kExtTmaRing; k++) {
            mbarInit(&full[k], 1);
            mbarInit(&empty[k], kExtTmaWarps);
        }
        mbarInitFence();
    }
    __syncthreads();
    if (wid == kExtTmaWarps) {
        // ---- producer ---------------------------------------------------------------------------
        if (lane == 0) {
            for (int j = 0; j < nBlocks; j++) {
                const int slot = j % kExtTmaRing;
                if (j >= kExtTmaRing) mbarWait(&empty[slot], ((j / kExtTmaRing) & 1) ^ 1);
                mbarExpectTx(&full[slot], (uint32_t)(kDogs * kExtTmaRows * kExtTmaCols * 4));
                tmaLoad3d(&map, &full[slot], ring + slot * kExtTmaBlockFloats, x0 - 4, y0 + j * kExtTmaRows, f * kDogs);
            }
        }
        return;
    }
    // ---- consumers --------------------------------------------------------------------------------
    const int xw = blockIdx.x * kExtTmaWarps + wid;
    const bool store = xw < o.maskRowWords;
    const int x = x0 + wid * 32 + lane;
    const bool xInside = (x >= 1) && (x <= o.w - 2);
    const int col = 4 + wid * 32 + lane;
    uint32_t* __restrict__ m = mask + ((size_t)f * blocksPerFrame + o.maskBlockStart) * (size_t)kScanChunk;
    // running mask addresses of the three scales (first output row yA), one word per row
    uint32_t* mp[kScales];
#pragma unroll
    for (int sc = 0; sc < kScales; sc++) mp[sc] = m + ((size_t)sc * o.h + yA) * o.maskRowWords + min(xw, o.maskRowWords - 1);
    RowPart win[kDogs][3];
    for (int j = 0; j < nBlocks; j++) {
        const int slot = j % kExtTmaRing;
        mbarWait(&full[slot], (j / kExtTmaRing) & 1);
        const float* blk = ring + slot * kExtTmaBlockFloats + col;
#pragma unroll
        for (int k = 0; k < kExtTmaRows; k++) {
            // input row y0 + 3 j + k goes to window slot k; the output row is the one before it
            const int prev = (k + 1) % 3, cur = (k + 2) % 3, next = k;
#pragma unroll
            for (int t = 0; t < kDogs; t++) {
                const float* p = blk + (t * kExtTmaRows + k) * kExtTmaCols;
                const float c = p[0], l = p[-1], r = p[1];
                RowPart q;
                q.c = c;
                q.lrn = fminf(l, r);
                q.lrx = fmaxf(l, r);
                q.crn = fminf(c, r);
                q.crx = fmaxf(c, r);
                q.h3n = fminf(l, q.crn);
                q.h3x = fmaxf(l, q.crx);
                win[t][next] = q;
            }
            const int y = y0 + j * kExtTmaRows + k - 1;
            if ((j > 0 || k == 2) && y < yB) {   // warp-uniform
#pragma unroll
                for (int s = 1; s <= kScales; s++) {
                    const float v = win[s][cur].c;
                    float mn = fminf(fminf(win[s][prev].h3n, win[s][next].h3n), win[s][cur].lrn);
                    float mx = fmaxf(fmaxf(win[s][prev].h3x, win[s][next].h3x), win[s][cur].lrx);
                    mn = fminf(mn, fminf(fminf(win[s + 1][prev].h3n, win[s + 1][cur].h3n), win[s + 1][next].h3n));
                    mx = fmaxf(mx, fmaxf(fmaxf(win[s + 1][prev].h3x, win[s + 1][cur].h3x), win[s + 1][next].h3x));
                    mn = fminf(mn, fminf(fminf(win[s - 1][prev].crn, win[s - 1][cur].h3n), win[s - 1][next].h3n));
                    mx = fmaxf(mx, fmaxf(fmaxf(win[s - 1][prev].crx, win[s - 1][cur].h3x), win[s - 1][next].h3x));
                    const bool cand = xInside && !(fabsf(v) <= softThreshold) && ((v < mn) || (v > mx));
                    const uint32_t word = __ballot_sync(0xffffffffu, cand);
                    if (lane == 0 && store) *mp[s - 1] = word;
                    mp[s - 1] += o.maskRowWords;
                }
            }
        }
        __syncwarp();
        if (lane == 0) mbarArrive(&empty[slot]);
    }
}

// Small planes (deep octaves): one thread per pixel, every load independent. The marching
// kernel above serialises >= 8 dependent row steps per warp, which is pure latency on a plane of a
// few thousand pixels at the end of the octave chain; this form finishes in one memory round trip.
__global__ void __launch_bounds__(256)
extremaMaskSmallKernel(const OctaveDev o, float softThreshold, uint32_t* __restrict__ mask,
                  int blocksPerFrame) {
    const int x = blockIdx.x * 256 + threadIdx.x;
    const int y = blockIdx.y;
    const int f = blockIdx.z;
    const int lane = threadIdx.x & 31;
    const int xw = x >> 5;
    if (xw >= o.maskRowWords) return;  // whole warp exits together (x is warp-aligned)
    const float* __restrict__ D = o.D + (size_t)f * kDogs * o.plane;
    uint32_t* __restrict__ m =
        mask + ((size_t)f * blocksPerFrame + o.maskBlockStart) * (size_t)kScanChunk;
    const bool inside = (x >= 1) && (x <= o.w - 2) && (y >= 1) && (y <= o.h - 2);
    const size_t c = (size_t)y * o.pitch + x;
#pragma unroll
    for (int s = 1; s <= kScales; s++) {
        bool cand = false;
        if (inside) {
            const float* __restrict__ p = D + (size_t)s * o.plane + c;
            const float v = __ldg(p);
            if (!(fabsf(v) <= softThreshold)) {
                float mn = +1000.0f, mx = -1000.0f;
#pragma unroll
                for (int ds = -1; ds <= 1; ds++) {
#pragma unroll
                    for (int dy = -1; dy <= 1; dy++) {
#pragma unroll
                        for (int dx = -1; dx <= 1; dx++) {
                            if (ds == 0 && dy == 0 && dx == 0) continue;       // the centre
                            if (ds == -1 && dy == -1 && dx == -1) continue;    // neighbour 0
                            const float nv = __ldg(p + (ptrdiff_t)ds * (ptrdiff_t)o.plane +
                                                   (ptrdiff_t)dy * o.pitch + dx);
                            mn = fminf(mn, nv);
                            mx = fmaxf(mx, nv);
                        }
                    }
                }
                cand = (v < mn) || (v > mx);
            }
        }
        const uint32_t word = __ballot_sync(0xffffffffu, cand);
        if (lane == 0) m[((size_t)(s - 1) * o.h + y) * o.maskRowWords + xw] = word;
    }
}

// Mask rows [yBegin, yEnd) ∩ [1, h - 1) of one octave (0, 0 = all rows): a row band of octave 0
// gets its mask as soon as that band's blur chain is done.
// Tensor map of one octave's DoG stack [nz][h][pitch] for extremaMaskTmaKernel.
cudaError_t makeExtremaTmaMap(CUtensorMap* map, const float* base, int pitch, int h, int nz, size_t planeFloats) {
    auto enc = tensorMapEncoder();
    if (!enc) return cudaErrorNotSupported;
    const cuuint64_t dims[3] = {(cuuint64_t)pitch, (cuuint64_t)h, (cuuint64_t)nz};
    const cuuint64_t strides[2] = {(cuuint64_t)pitch * 4, (cuuint64_t)planeFloats * 4};
    const cuuint32_t box[3] = {(cuuint32_t)kExtTmaCols, (cuuint32_t)kExtTmaRows, (cuuint32_t)kDogs};
    const cuuint32_t elem[3] = {1, 1, 1};
    const CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(base), dims, strides, box, elem,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? cudaSuccess : cudaErrorInvalidValue;
}

cudaError_t launchExtremaMask(const EngineParams& P, int octave, uint32_t* mask, int frames,
                              cudaStream_t st, int yBegin, int yEnd, int priority, const CUtensorMap* dogMap) {
    const OctaveDev& o = P.oct[octave];
    if (o.w < 3 || o.h < 3) return cudaSuccess;
    const bool all = yEnd <= 0;
    const int yA = all ? 1 : std::max(yBegin, 1), yB = all ? o.h - 1 : std::min(yEnd, o.h - 1);
    if (yB <= yA) return cudaSuccess;
    if (all && (long)o.w * o.h * frames <= 96 * 1024) {
        dim3 grid((o.maskRowWords * 32 + 255) / 256, o.h, frames);
        return launchKernel(extremaMaskSmallKernel, grid, dim3(256), 0, st, false, priority, o, P.dogThreshold * 0.8f,
                            mask, P.blocksPerFrame);
    }
    static const bool tmaEnabled = !(getenv("SIFTCUDA_EXTREMA_TMA") && atoi(getenv("SIFTCUDA_EXTREMA_TMA")) == 0);   // tuning switch
    // (measured: 58 against 70 us on the 3840 x 2160 plane, 2 - 3 % of the step on the batch and 8K
    // workloads; on planes that cannot fill the SMs twice over with 24-row strips the register
    // march below is as fast or faster)
    const long tmaCtas = (long)((o.maskRowWords + kExtTmaWarps - 1) / kExtTmaWarps) * ((yB - yA + 23) / 24) * frames;
    if (tmaEnabled && dogMap && all && o.w >= 512 && o.h >= 128 && tmaCtas >= 2L * 3 * 148) {
        static std::atomic<unsigned long long> configured{0};
        int dev = 0;
        cudaGetDevice(&dev);
        if (!((configured.load(std::memory_order_acquire) >> (dev & 63)) & 1ull)) {
            SIFT_CUDA_TRY(cudaFuncSetAttribute(extremaMaskTmaKernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kExtTmaSmemBytes));
            configured.fetch_or(1ull << (dev & 63), std::memory_order_release);
        }
        const int colGroups = (o.maskRowWords + kExtTmaWarps - 1) / kExtTmaWarps;
        const int nRows = yB - yA;
        // strips: whole waves of the 3 x 148 resident CTAs when the plane is large enough, rows per
        // strip >= 24 so that the two rows of prologue stay below 8 %
        static const int rowsEnv = getenv("SIFTCUDA_EXTREMA_ROWS") ? atoi(getenv("SIFTCUDA_EXTREMA_ROWS")) : 0;
        int strips = std::max(1, nRows / 60);
        const long resident = 3L * 148;
        const long ctas = (long)colGroups * strips * frames;
        if (ctas > resident) {
            const long waves = (ctas + resident - 1) / resident;
            strips = (int)std::max<long>(1, waves * resident / ((long)colGroups * frames));
        }
        int rows = (nRows + strips - 1) / strips;
        rows = std::max(rows, std::min(nRows, 24));
        if (rowsEnv > 0) rows = rowsEnv;
        dim3 grid(colGroups, (nRows + rows - 1) / rows, frames);
        return launchKernel(extremaMaskTmaKernel, grid, dim3((kExtTmaWarps + 1) * 32), (size_t)kExtTmaSmemBytes, st, false,
                            priority, *dogMap, o, P.dogThreshold * 0.8f, mask, P.blocksPerFrame, rows, yA, yB);
    }
    // rows per warp (multiple of 3): long strips amortise the 2-row halo on large planes; small
    // planes get short strips so that the serial row loop does not bound the launch
    const int gx = (o.maskRowWords + kExtWarps - 1) / kExtWarps;
    const int nRows = yB - yA;
    int rows = kExtRows;
    while (rows > 6 && (long)gx * ((nRows + rows - 1) / rows) * frames < 592) rows -= 6;
    dim3 grid(gx, (nRows + rows - 1) / rows, frames);
    return launchKernel(extremaMaskKernel, grid, dim3(kExtWarps * 32), 0, st, false, priority, o, P.dogThreshold * 0.8f,
                        mask, P.blocksPerFrame, rows, yA, yB);
}

// ------------------------------------------------------------------------------------------
// scan phase B (declared in scan.cuh)
__global__ void __launch_bounds__(1024)
scanOffsetsKernel(int* blockSums, int n, int* totalOut, int capacity, int* overflow,
                  int overflowBit, const MaskSegments segs) {
    pdlPrologue();
    __shared__ int warpSums[33];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int ipt = (n + 1023) / 1024;
    const int begin = min(tid * ipt, n), end = min(begin + ipt, n);
    int s = 0;
    for (int i = begin; i < end; i++) s += blockSums[i];
    int inc = s;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= d) inc += v;
    }
    if (lane == 31) warpSums[wid] = inc;
    __syncthreads();
    if (wid == 0) {
        const int ws = warpSums[lane];
        int winc = ws;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, winc, d);
            if (lane >= d) winc += v;
        }
        warpSums[lane] = winc - ws;
        if (lane == 31) warpSums[32] = winc;
    }
    __syncthreads();
    int run = inc - s + warpSums[wid];
    for (int i = begin; i < end; i++) {
        const int v = blockSums[i];
        blockSums[i] = run;
        run += v;
    }
    const int total = min(warpSums[32], capacity);
    if (tid == 0) {
        if (warpSums[32] > capacity) atomicOr(overflow, overflowBit);
        *totalOut = total;
    }
    if (segs.segStart) {
        // list segment (frame, octave) starts where its first mask block starts
        __syncthreads();   // scanned sums of the other threads
        for (int seg = tid; seg <= segs.nSegs; seg += 1024) {
            int v = total;
            if (seg < segs.nSegs) {
                const int frame = seg / kOctaves, oc = seg - frame * kOctaves;
                const int lb = frame * segs.blocksPerFrame + segs.octaveBlockStart[oc] - segs.blockBegin;
                if (lb <= 0) v = 0;
                else if (lb < n) v = min(blockSums[lb], total);
            }
            segs.segStart[seg] = v;
        }
    }
}

cudaError_t launchScanOffsets(int* blockSums, int n, int* totalOut, int capacity, int* overflow,
                              int overflowBit, cudaStream_t st, const MaskSegments* segs) {
    return pdlLaunch(scanOffsetsKernel, dim3(1), dim3(1024), 0, st, true, blockSums, n, totalOut, capacity,
                     overflow, overflowBit, segs ? *segs : MaskSegments{});
}

struct MaskPopc {
    const uint32_t* mask;
    __device__ int operator()(int i) const { return __popc(__ldg(mask + i)); }
};

// Phase C for the extrema mask: every set bit becomes a Candidate at its scanned position.
__global__ void __launch_bounds__(kScanThreads)
scatterCandidatesKernel(const __grid_constant__ EngineParams P, const uint32_t* __restrict__ mask,
                        const int* __restrict__ blockOffsets, Candidate* __restrict__ cands,
                        int capacity, int blockBegin) {
    pdlPrologue();
    __shared__ int sh[9];
    const int b = blockBegin + blockIdx.x;   // global mask block; blockOffsets is indexed locally
    const int frame = b / P.blocksPerFrame;
    const int fb = b - frame * P.blocksPerFrame;
    int oc = 0;
#pragma unroll
    for (int k = 1; k < kOctaves; k++)
        if (fb >= P.oct[k].maskBlockStart) oc = k;
    const OctaveDev& o = P.oct[oc];
    const int wordInOctave = (fb - o.maskBlockStart) * kScanChunk + threadIdx.x * kScanItemsPerThread;
    const uint32_t* __restrict__ src = mask + (size_t)b * kScanChunk + threadIdx.x * kScanItemsPerThread;
    uint32_t words[kScanItemsPerThread];
    int s = 0;
#pragma unroll
    for (int k = 0; k < kScanItemsPerThread; k++) {
        words[k] = __ldg(src + k);
        s += __popc(words[k]);
    }
    int total;
    int pos = blockOffsets[blockIdx.x] + blockExclusiveScan256(s, sh, &total);
    if (s == 0) return;
    const int seg = frame * kOctaves + oc;
#pragma unroll
    for (int k = 0; k < kScanItemsPerThread; k++) {
        uint32_t wbits = words[k];
        if (wbits == 0) continue;
        const int wi = wordInOctave + k;
        const int row = wi / o.maskRowWords;         // (s - 1) * h + y
        const int xw = wi - row * o.maskRowWords;
        const int sc = row / o.h;
        const int y = row - sc * o.h;
        while (wbits) {
            const int bit = __ffs(wbits) - 1;
            wbits &= wbits - 1;
            if (pos < capacity) {
                Candidate cd;
                cd.xys = packXYS(xw * 32 + bit, y, sc + 1);
                cd.seg = seg;
                cands[pos] = cd;
            }
            pos++;
        }
    }
}

cudaError_t launchCandidateCompaction(const EngineParams& P, const uint32_t* mask,
                                      int* blockSums, Candidate* cands, int capCandidates,
                                      int blockBegin, int nBlocks, int* segStart, int nSegs,
                                      Counters* counters, cudaStream_t st) {
    if (nBlocks < 1) return cudaSuccess;
    MaskPopc v{mask + (size_t)blockBegin * kScanChunk};
    scanBlockSumsKernel<<<nBlocks, kScanThreads, 0, st>>>(v, blockSums);
    SIFT_CUDA_TRY(cudaGetLastError());
    MaskSegments ms;
    ms.segStart = segStart;
    ms.nSegs = nSegs;
    ms.blocksPerFrame = P.blocksPerFrame;
    ms.blockBegin = blockBegin;
    for (int o = 0; o < kOctaves; o++) ms.octaveBlockStart[o] = P.oct[o].maskBlockStart;
    SIFT_CUDA_TRY(launchScanOffsets(blockSums, nBlocks, &counters->nCandidates, capCandidates,
                                    &counters->overflow, 1, st, &ms));
    return pdlLaunch(scatterCandidatesKernel, dim3(nBlocks), dim3(kScanThreads), 0, st, true, P, mask,
                     (const int*)blockSums, cands, capCandidates, blockBegin);
}

// ------------------------------------------------------------------------------------------
// Refinement.
struct V3 {
    float x, y, z;
};
__device__ __forceinline__ V3 cross3(const V3 a, const V3 b) {
    V3 r;
    r.x = (a.y * b.z) - (a.z * b.y);
    r.y = (a.z * b.x) - (a.x * b.z);
    r.z = (a.x * b.y) - (a.y * b.x);
    return r;
}

struct DogView {
    const float* __restrict__ base;  // slice 0 of this frame's DoG stack
    size_t plane;
    int pitch;
    __device__ __forceinline__ float at(int x, int y, int s) const {
        return __ldg(base + (size_t)s * plane + (size_t)y * pitch + x);
    }
};

// SIFTInterpolate.metal:156-177 interpolationStep: alpha = -(H^-1) dD, H from hessian3D
// (:103-164), inverse by cross products (Common.hpp:34-47) with det = dot(x0, cross(x1, x2)).
__device__ __forceinline__ V3 interpolationStep(const DogView& t, int x, int y, int s, V3* dDout) {
    const float zzz = t.at(x, y, s);
    const float pzz = t.at(x + 1, y, s), nzz = t.at(x - 1, y, s);
    const float zpz = t.at(x, y + 1, s), znz = t.at(x, y - 1, s);
    const float zzp = t.at(x, y, s + 1), zzn = t.at(x, y, s - 1);
    const float ppz = t.at(x + 1, y + 1, s), nnz = t.at(x - 1, y - 1, s);
    const float npz = t.at(x - 1, y + 1, s), pnz = t.at(x + 1, y - 1, s);
    const float pzp = t.at(x + 1, y, s + 1), nzp = t.at(x - 1, y, s + 1);
    const float zpp = t.at(x, y + 1, s + 1), znp = t.at(x, y - 1, s + 1);
    const float pzn = t.at(x + 1, y, s - 1), nzn = t.at(x - 1, y, s - 1);
    const float zpn = t.at(x, y + 1, s - 1), znn = t.at(x, y - 1, s - 1);

    const float dxx = (pzz + nzz) - (2 * zzz);
    const float dyy = (zpz + znz) - (2 * zzz);
    const float dss = (zzp + zzn) - (2 * zzz);
    const float dxy = (((ppz - npz) - pnz) + nnz) * 0.25f;
    const float dxs = (((pzp - nzp) - pzn) + nzn) * 0.25f;
    const float dys = (((zpp - znp) - zpn) + znn) * 0.25f;

    const V3 x0 = {dxx, dxy, dxs}, x1 = {dxy, dyy, dys}, x2 = {dxs, dys, dss};
    const V3 c0 = cross3(x1, x2), c1 = cross3(x2, x0), c2 = cross3(x0, x1);
    const float d = ((x0.x * c0.x) + (x0.y * c0.y)) + (x0.z * c0.z);
    const float inv = 1.0f / d;
    const V3 h0 = {-(inv * c0.x), -(inv * c0.y), -(inv * c0.z)};
    const V3 h1 = {-(inv * c1.x), -(inv * c1.y), -(inv * c1.z)};
    const V3 h2 = {-(inv * c2.x), -(inv * c2.y), -(inv * c2.z)};
    V3 dD;
    dD.x = (pzz - nzz) * 0.5f;
    dD.y = (zpz - znz) * 0.5f;
    dD.z = (zzp - zzn) * 0.5f;
    *dDout = dD;
    V3 a;
    a.x = ((h0.x * dD.x) + (h1.x * dD.y)) + (h2.x * dD.z);
    a.y = ((h0.y * dD.x) + (h1.y * dD.y)) + (h2.y * dD.z);
    a.z = ((h0.z * dD.x) + (h1.z * dD.y)) + (h2.z * dD.z);
    return a;
}

// SIFTInterpolate.metal:17-61 isOnEdge.
__device__ __forceinline__ bool isOnEdge(const DogView& t, int x, int y, int s, float edgeThreshold) {
    const float v = t.at(x, y, s);
    const float zn = t.at(x, y - 1, s), zp = t.at(x, y + 1, s);
    const float pz = t.at(x + 1, y, s), nz = t.at(x - 1, y, s);
    const float pp = t.at(x + 1, y + 1, s), np = t.at(x - 1, y + 1, s);
    const float pn = t.at(x + 1, y - 1, s), nn = t.at(x - 1, y - 1, s);
    const float hxx = (zn + zp) - (2 * v);
    const float hyy = (pz + nz) - (2 * v);
    const float hxy = ((pp - np) - (pn - nn)) * 0.25f;
    const float trace = hxx + hyy;
    const float determinant = (hxx * hyy) - (hxy * hxy);
    if (determinant <= 0) return true;
    const float threshold = ((edgeThreshold + 1) * (edgeThreshold + 1)) / edgeThreshold;
    const float curvature = (trace * trace) / determinant;
    return curvature >= threshold;
}

__device__ __forceinline__ bool outOfBounds(int x, int y, int s, int w, int h, int border) {
    return x < border || x > w - border - 1 || y < border || y > h - border - 1 || s < 1 ||
           s > kScales;
}

// One thread per candidate; a CTA of 256 threads owns 256 consecutive candidates, ballots the
// converged flags into 8 words and publishes its count, so the following scan + scatter keep the
// canonical candidate order and every thread copies at most one keypoint.
constexpr int kRefineThreads = 256;

__global__ void __launch_bounds__(kRefineThreads)
refineKernel(const __grid_constant__ EngineParams P, const Candidate* __restrict__ cands,
             const Counters* __restrict__ counters, SiftKeypoint* __restrict__ kpTmp,
             uint32_t* __restrict__ flagWords, int* __restrict__ blockSums) {
    pdlPrologue();
    __shared__ int warpCount[kRefineThreads / 32];
    const int n = counters->nCandidates;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int i = blockIdx.x * kRefineThreads + threadIdx.x;
    bool ok = false;
    if (i < n) {
        const Candidate cd = cands[i];
        const int frame = cd.seg / kOctaves, oc = cd.seg - frame * kOctaves;
        const OctaveDev& o = P.oct[oc];
        DogView t;
        t.base = o.D + (size_t)frame * kDogs * o.plane;
        t.plane = o.plane;
        t.pitch = o.pitch;
        int x = cd.xys & 0x7fff, y = (cd.xys >> 15) & 0x7fff, s = cd.xys >> 30;
        // the 0.8·C_DoG test (SIFTInterpolate.metal:208) already held in the extrema kernel
        if (!outOfBounds(x, y, s, o.w, o.h, P.border)) {
            bool converged = false, alive = true;
            V3 alpha = {0, 0, 0}, dD = {0, 0, 0};
            int it = 0;
            while (it < P.maxIterations) {
                alpha = interpolationStep(t, x, y, s, &dD);
                if ((fabsf(alpha.x) < P.maxOffset) && (fabsf(alpha.y) < P.maxOffset) &&
                    (fabsf(alpha.z) < P.maxOffset)) {
                    converged = true;
                    break;
                }
                if (alpha.x > +P.maxOffset) x += 1;
                if (alpha.x < -P.maxOffset) x -= 1;
                if (alpha.y > +P.maxOffset) y += 1;
                if (alpha.y < -P.maxOffset) y -= 1;
                if (alpha.z > +P.maxOffset) s += 1;
                if (alpha.z < -P.maxOffset) s -= 1;
                if (outOfBounds(x, y, s, o.w, o.h, P.border)) {
                    alive = false;
                    break;
                }
                it += 1;
            }
            if (alive && converged) {
                // interpolateContrast (:90-100): v + 0.5·(dD.x·alpha.x) — x term only
                const float value = t.at(x, y, s) + ((dD.x * alpha.x) * 0.5f);
                if (!(fabsf(value) <= P.dogThreshold) && !isOnEdge(t, x, y, s, P.edgeThreshold)) {
                    SiftKeypoint k;
                    k.octave = oc;
                    k.scale = s;
                    k.subScale = alpha.z;
                    k.scaledX = x;
                    k.scaledY = y;
                    k.absoluteX = ((float)x + alpha.x) * o.delta;
                    k.absoluteY = ((float)y + alpha.y) * o.delta;
                    k.normalizedX = (float)x / (float)o.w;
                    k.normalizedY = (float)y / (float)o.h;
                    k.sigma = o.sigmas[s] * dm_exp2f(alpha.z * o.log2SigmaRatio);
                    k.value = value;
                    kpTmp[i] = k;
                    ok = true;
                }
            }
        }
    }
    const uint32_t word = __ballot_sync(0xffffffffu, ok);
    if (lane == 0) {
        flagWords[i >> 5] = word;
        warpCount[wid] = __popc(word);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        int total = 0;
#pragma unroll
        for (int k = 0; k < kRefineThreads / 32; k++) total += warpCount[k];
        blockSums[blockIdx.x] = total;
    }
}

__global__ void __launch_bounds__(kRefineThreads)
scatterKeypointsKernel(const uint32_t* __restrict__ flags, const Counters* __restrict__ counters,
                       const int* __restrict__ blockOffsets, const Candidate* __restrict__ cands,
                       const SiftKeypoint* __restrict__ kpTmp, SiftKeypoint* __restrict__ kps,
                       int* __restrict__ kpSeg, int capacity, const int* __restrict__ candSegStart,
                       int* __restrict__ kpSegStart, int nSegs, const KeypointColumnsDev hc) {
    pdlPrologue();
    const int n = counters->nCandidates;
    const int i = blockIdx.x * kRefineThreads + threadIdx.x;
    // keypoint segment starts: the scanned position of each segment's first candidate (lists are
    // sorted by segment; no search, no atomics)
    if (i <= nSegs) {
        const int c0 = candSegStart[i];
        int v = counters->nKeypoints;
        if (c0 < n) {
            const int blk = c0 / kRefineThreads, w0 = (c0 % kRefineThreads) >> 5, bit = c0 & 31;
            const uint32_t* __restrict__ fw = flags + blk * (kRefineThreads / 32);
            int pos = blockOffsets[blk] + __popc(fw[w0] & ((1u << bit) - 1u));
            for (int k = 0; k < w0; k++) pos += __popc(fw[k]);
            v = min(pos, v);
        }
        kpSegStart[i] = v;
    }
    if (blockIdx.x * kRefineThreads >= n) return;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const uint32_t* __restrict__ w = flags + blockIdx.x * (kRefineThreads / 32);
    const uint32_t mine = w[wid];
    if (!((mine >> lane) & 1u)) return;
    int pos = blockOffsets[blockIdx.x] + __popc(mine & ((1u << lane) - 1u));
    for (int k = 0; k < wid; k++) pos += __popc(w[k]);
    if (pos < capacity) {
        const SiftKeypoint k = kpTmp[i];
        kps[pos] = k;
        kpSeg[pos] = cands[i].seg;
        if (hc.absX) {
            // result columns (SiftKeypointColumns), usually pinned host memory: these stores are the
            // D2H of the keypoints; consecutive threads write consecutive slots of each column
            hc.absX[pos] = k.absoluteX;
            hc.absY[pos] = k.absoluteY;
            hc.sigma[pos] = k.sigma;
            hc.value[pos] = k.value;
            hc.subScale[pos] = k.subScale;
            hc.scaledXY[pos] = make_short2((short)k.scaledX, (short)k.scaledY);
            hc.octaveScale[pos] = make_uchar2((unsigned char)k.octave, (unsigned char)k.scale);
        }
    }
}

cudaError_t launchRefine(const EngineParams& P, const Candidate* cands, int capCandidates,
                         SiftKeypoint* kpTmp, uint32_t* flagWords, int* blockSums,
                         SiftKeypoint* kps, int* kpSeg, int capKeypoints, const int* segCandStart,
                         int* segKpStart, int nSegs, Counters* counters,
                         const KeypointColumnsDev& hostCols, cudaStream_t st) {
    const int nBlocks = (capCandidates + kRefineThreads - 1) / kRefineThreads;
    SIFT_CUDA_TRY(pdlLaunch(refineKernel, dim3(nBlocks), dim3(kRefineThreads), 0, st, true, P, cands,
                            (const Counters*)counters, kpTmp, flagWords, blockSums));
    SIFT_CUDA_TRY(launchScanOffsets(blockSums, nBlocks, &counters->nKeypoints, capKeypoints,
                                    &counters->overflow, 2, st));
    return pdlLaunch(scatterKeypointsKernel, dim3(nBlocks), dim3(kRefineThreads), 0, st, true,
                     (const uint32_t*)flagWords, (const Counters*)counters, (const int*)blockSums, cands,
                     (const SiftKeypoint*)kpTmp, kps, kpSeg, capKeypoints, segCandStart, segKpStart, nSegs,
                     hostCols);
}

}  // namespace sift
