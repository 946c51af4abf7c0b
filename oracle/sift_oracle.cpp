// sift_oracle.cpp — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// Scalar float32 CPU restatement of lukevanin/SIFTMetal's detect + describe path: the Metal
// kernels of Sources/MetalShaders/Metal/*.metal and the Swift host stages of
// Sources/SIFTMetal/SIFT/{DifferenceOfGaussians,SIFTOctave,SIFT}.swift, one function per
// reference kernel / host stage, each citing the file:line it follows (paths relative to the
// reference root). Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// `--impl reference` legs may load this library; libsiftcuda.so never links or calls it.
//
// Why C++ and not Swift: the build container and the GPU boxes have no Swift toolchain
// (SURVEY.md §8c); the reference itself needs Apple Metal and cannot run here at all.
//
// Pinning: the reference's own XCTest code asserts nothing numeric on this path, but its test
// resources are IPOL "Anatomy of SIFT" dumps for butterfly.png; tests/test_oracle_fixtures.py
// holds this oracle to them (keypoint count, sub-0.01 px positions, θ offset, descriptor match
// rate — bands from SURVEY.md §8c).
//
// Arithmetic spec: binary32 everywhere; each expression is evaluated exactly as
// parenthesised here with no contraction (compile with -ffp-contract=off), EXCEPT the
// convolution accumulation, which is a fused multiply-add chain `sum = fma(w[i], c, sum)` in
// ascending tap order — the contraction Metal's default fast-math applies to
// `sum += w * c` (Convolution.metal:28, ConvolutionSeries.metal:30). exp/atan2/sin/cos/exp2 are
// the fixed sequences of oracle_math.h. Where the reference is undefined (Metal reads outside a
// texture, unordered atomics) the choice made here is stated at the site.

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../include/siftcuda.h"
#include "oracle_math.h"

#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

struct Plane {
    int w = 0, h = 0;
    std::vector<float> d;
    void resize(int w_, int h_) { w = w_; h = h_; d.assign((size_t)w * h, 0.0f); }
    float& at(int x, int y) { return d[(size_t)y * w + x]; }
    float at(int x, int y) const { return d[(size_t)y * w + x]; }
};

struct Candidate { int32_t x, y, s; };

struct Octave {
    int w = 0, h = 0;
    float delta = 0;
    float sigmas[SIFT_NUM_GAUSSIANS];
    Plane G[SIFT_NUM_GAUSSIANS];
    Plane D[SIFT_NUM_DOGS];
    std::vector<float> grad[SIFT_NUM_GAUSSIANS];  // interleaved (orientation, magnitude)
    std::vector<Candidate> candidates;            // canonical order (s, y, x)
};

struct Stats {
    int64_t raw25 = 0, raw26 = 0, soft = 0, interp = 0, contrast = 0, final_ = 0;
};

struct Oracle {
    SiftConfig cfg;
    int W = 0, H = 0;
    int seedW = 0, seedH = 0;
    float seedSigma = 0;
    std::vector<float> seedWeights;
    float rho[SIFT_NUM_GAUSSIANS - 1];
    std::vector<float> weights[SIFT_NUM_GAUSSIANS - 1];
    Plane gray, scaled, seed;
    Octave oct[SIFT_NUM_OCTAVES];
    std::vector<SiftKeypoint> keypoints;           // octave-major
    int32_t keypointCounts[SIFT_NUM_OCTAVES];
    std::vector<SiftDescriptor> descriptors;       // octave-major
    int32_t descriptorCounts[SIFT_NUM_OCTAVES];
    Stats stats[SIFT_NUM_OCTAVES];
    bool collectStats = false;
    bool allGradients = false;  // reference computes 6 slices; only 1..3 are ever read
    bool detected = false;
};

// Common.hpp:15-22 symmetrizedCoordinates. The reference form `(i + 2l) % 2l` is valid for
// i >= -2l only; the floor-mod below is identical there and defined beyond.
inline int symmetrized(int i, int l) {
    int ll = 2 * l;
    i = ((i % ll) + ll) % ll;
    if (i > l - 1) i = ll - 1 - i;
    return i;
}

// GaussianKernel.swift:20-43 ≡ GaussianSeriesKernel.swift:27-51 (host, Float).
std::vector<float> gaussianWeights(float s) {
    int radius = (int)ceilf(4 * s);
    int size = radius * 2 + 1;
    std::vector<float> w;
    float t = 0;
    float ss = s * s;
    for (int k = -radius; k <= radius; k++) {
        float kk = (float)(k * k);
        float v = expf(-0.5f * (kk / ss));
        w.push_back(v);
        t += v;
    }
    for (int i = 0; i < size; i++) w[i] = w[i] / t;
    return w;
}

// ConvertSRGBToGrayscale.metal:11-23. bgra8Unorm read = byte / 255; no gamma linearisation.
void convertSRGBToGrayscale(const uint8_t* bgra, int pitch, Plane& out) {
#pragma omp parallel for schedule(static)
    for (int y = 0; y < out.h; y++) {
        const uint8_t* row = bgra + (size_t)y * pitch;
        for (int x = 0; x < out.w; x++) {
            float b = (float)row[4 * x + 0] / 255.0f;
            float g = (float)row[4 * x + 1] / 255.0f;
            float r = (float)row[4 * x + 2] / 255.0f;
            float i = ((0.0f + (0.212639005871510f * r)) + (0.715168678767756f * g)) +
                      (0.072192315360734f * b);
            out.at(x, y) = i;
        }
    }
}

// BilinearUpScale.metal:12-64.
void bilinearUpScale(const Plane& in, Plane& out) {
    const int wo = out.w, ho = out.h, wi = in.w, hi = in.h;
    const float dx = (float)wi / (float)wo;
    const float dy = (float)hi / (float)ho;
#pragma omp parallel for schedule(static)
    for (int j = 0; j < ho; j++) {
        for (int i = 0; i < wo; i++) {
            const float x = (float)i * dx;
            const float y = (float)j * dy;
            int im = (int)x, jm = (int)y;
            int ip = im + 1, jp = jm + 1;
            if (ip >= wi) ip = 2 * wi - 1 - ip;
            if (im >= wi) im = 2 * wi - 1 - im;
            if (jp >= hi) jp = 2 * hi - 1 - jp;
            if (jm >= hi) jm = 2 * hi - 1 - jm;
            const float fx = x - floorf(x);
            const float fy = y - floorf(y);
            const float c0 = in.at(ip, jp), c1 = in.at(ip, jm);
            const float c2 = in.at(im, jp), c3 = in.at(im, jm);
            const float a = (fy * c0) + ((1 - fy) * c1);
            const float b = (fy * c2) + ((1 - fy) * c3);
            out.at(i, j) = (fx * a) + ((1 - fx) * b);
        }
    }
}

// Convolution.metal:15-32 ≡ ConvolutionSeries.metal:16-33 (X) — per pixel:
// sum = 0; for i in 0..<n: sum = fma(w[i], in[sym(x - n/2 + i)], sum).
// The mirrored row is materialised once so the tap loop is a plain stride; per-pixel operation
// order is unchanged.
void convolutionX(const Plane& in, Plane& out, const std::vector<float>& wts) {
    const int n = (int)wts.size(), r = n / 2, w = in.w, h = in.h;
#pragma omp parallel
    {
        std::vector<float> ext((size_t)w + 2 * r);
#pragma omp for schedule(static)
        for (int y = 0; y < h; y++) {
            for (int x = -r; x < w + r; x++) ext[x + r] = in.at(symmetrized(x, w), y);
            float* o = &out.d[(size_t)y * w];
            for (int x = 0; x < w; x++) o[x] = 0.0f;
            for (int i = 0; i < n; i++) {
                const float wi = wts[i];
                const float* e = ext.data() + i;
                for (int x = 0; x < w; x++) o[x] = fmaf(wi, e[x], o[x]);
            }
        }
    }
}

// Convolution.metal:35-52 ≡ ConvolutionSeries.metal:36-53 (Y).
void convolutionY(const Plane& in, Plane& out, const std::vector<float>& wts) {
    const int n = (int)wts.size(), r = n / 2, w = in.w, h = in.h;
#pragma omp parallel for schedule(static)
    for (int y = 0; y < h; y++) {
        float* o = &out.d[(size_t)y * w];
        for (int x = 0; x < w; x++) o[x] = 0.0f;
        for (int i = 0; i < n; i++) {
            const float wi = wts[i];
            const float* src = &in.d[(size_t)symmetrized(y - r + i, h) * w];
            for (int x = 0; x < w; x++) o[x] = fmaf(wi, src[x], o[x]);
        }
    }
}

// GaussianKernel.encode (GaussianKernel.swift:62-92) / GaussianSeriesKernel.encode (:107-118):
// X pass into a working plane, Y pass out.
void gaussianBlur(const Plane& in, Plane& out, const std::vector<float>& wts) {
    Plane tmp;
    tmp.resize(in.w, in.h);
    convolutionX(in, tmp, wts);
    convolutionY(tmp, out, wts);
}

// NearestNeighborDownScale.metal:15-22.
void nearestNeighborDownScale(const Plane& in, Plane& out) {
#pragma omp parallel for schedule(static)
    for (int y = 0; y < out.h; y++)
        for (int x = 0; x < out.w; x++) out.at(x, y) = in.at(2 * x, 2 * y);
}

// Subtract.metal:12-21.
void subtract(const Plane& a, const Plane& b, Plane& out) {
    const size_t n = out.d.size();
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < n; i++) out.d[i] = a.d[i] - b.d[i];
}

// SIFTGradient.metal:15-39. atan2(dx, dy): x-derivative is the first argument.
void siftGradient(const Plane& g, std::vector<float>& out) {
    const int w = g.w, h = g.h;
    out.resize((size_t)w * h * 2);
#pragma omp parallel for schedule(static)
    for (int y = 0; y < h; y++) {
        const int py = symmetrized(y + 1, h), my = symmetrized(y - 1, h);
        for (int x = 0; x < w; x++) {
            const int px = symmetrized(x + 1, w), mx = symmetrized(x - 1, w);
            const float tx = (g.at(px, y) - g.at(mx, y)) * 0.5f;
            const float ty = (g.at(x, py) - g.at(x, my)) * 0.5f;
            out[((size_t)y * w + x) * 2 + 0] = om_atan2f(tx, ty);
            out[((size_t)y * w + x) * 2 + 1] = sqrtf((tx * tx) + (ty * ty));
        }
    }
}

// SIFTExtrema.metal:15-45 neighbour table, (dx, dy, ds).
const int kNeighborOffsets[26][3] = {
    {-1, -1, -1}, {0, -1, -1}, {+1, -1, -1}, {-1, 0, -1}, {0, 0, -1}, {+1, 0, -1},
    {-1, +1, -1}, {0, +1, -1}, {+1, +1, -1},
    {-1, -1, 0},  {0, -1, 0},  {+1, -1, 0},  {-1, 0, 0},
    {+1, 0, 0},   {-1, +1, 0}, {0, +1, 0},   {+1, +1, 0},
    {-1, -1, +1}, {0, -1, +1}, {+1, -1, +1}, {-1, 0, +1}, {0, 0, +1}, {+1, 0, +1},
    {-1, +1, +1}, {0, +1, +1}, {+1, +1, +1},
};

// SIFTExtrema.metal:62-110 siftExtremaList, grid (w−2)×(h−2)×3 (SIFTExtremaListKernel.swift:
// 52-62): strict min/max against neighbours 1…25 — index 0, (−1,−1,−1), is skipped at :84.
// Fused with the first test of siftInterpolate (SIFTInterpolate.metal:208,
// abs(v) <= 0.8·dogThreshold → discard), which is result-neutral. Output order: the reference's
// is unordered (atomics); canonical order here is (s, y, x).
void siftExtremaList(Octave& o, float dogThreshold, Stats& st, bool collect) {
    const int w = o.w, h = o.h;
    const float soft = dogThreshold * 0.8f;
    o.candidates.clear();
    for (int s = 1; s <= SIFT_SCALES_PER_OCTAVE; s++) {
        std::vector<std::vector<Candidate>> rows(std::max(h, 1));
        int64_t raw25 = 0, raw26 = 0;
#pragma omp parallel for schedule(static) reduction(+ : raw25, raw26)
        for (int y = 1; y <= h - 2; y++) {
            for (int x = 1; x <= w - 2; x++) {
                const float v = o.D[s].at(x, y);
                const bool pass = !(fabsf(v) <= soft);
                if (!pass && !collect) continue;
                float minimum = +1000, maximum = -1000;
                for (int i = 1; i < 26; i++) {
                    const int* n = kNeighborOffsets[i];
                    const float nv = o.D[s + n[2]].at(x + n[0], y + n[1]);
                    minimum = fminf(minimum, nv);
                    maximum = fmaxf(maximum, nv);
                }
                const bool ext = (v < minimum) || (v > maximum);
                if (collect) {
                    const float n0 = o.D[s - 1].at(x - 1, y - 1);
                    if (ext) raw25++;
                    if ((v < fminf(minimum, n0)) || (v > fmaxf(maximum, n0))) raw26++;
                }
                if (ext && pass) rows[y].push_back({x, y, s});
            }
        }
        st.raw25 += raw25;
        st.raw26 += raw26;
        for (int y = 1; y <= h - 2; y++)
            o.candidates.insert(o.candidates.end(), rows[y].begin(), rows[y].end());
    }
    st.soft += (int64_t)o.candidates.size();
}

struct Vec3 { float x, y, z; };

inline Vec3 cross(const Vec3& a, const Vec3& b) {
    return {(a.y * b.z) - (a.z * b.y), (a.z * b.x) - (a.x * b.z), (a.x * b.y) - (a.y * b.x)};
}

// SIFTInterpolate.metal:64-87 derivatives3D.
inline Vec3 derivatives3D(const Octave& o, int x, int y, int s) {
    const float pzz = o.D[s].at(x + 1, y), nzz = o.D[s].at(x - 1, y);
    const float zpz = o.D[s].at(x, y + 1), znz = o.D[s].at(x, y - 1);
    const float zzp = o.D[s + 1].at(x, y), zzn = o.D[s - 1].at(x, y);
    return {(pzz - nzz) * 0.5f, (zpz - znz) * 0.5f, (zzp - zzn) * 0.5f};
}

// SIFTInterpolate.metal:156-177 interpolationStep = -(H^-1)·dD with hessian3D (:103-164) and
// invert (Common.hpp:34-47). Metal's determinant() leaves the operation order open; the spec
// adopts the formula in the reference's own comment (Common.hpp:39): dot(x0, cross(x1, x2)).
inline Vec3 interpolationStep(const Octave& o, int x, int y, int s) {
    const Plane &c = o.D[s], &p = o.D[s + 1], &n = o.D[s - 1];
    const float zzz = c.at(x, y);
    const float pzz = c.at(x + 1, y), nzz = c.at(x - 1, y);
    const float zpz = c.at(x, y + 1), znz = c.at(x, y - 1);
    const float zzp = p.at(x, y), zzn = n.at(x, y);
    const float ppz = c.at(x + 1, y + 1), nnz = c.at(x - 1, y - 1);
    const float npz = c.at(x - 1, y + 1), pnz = c.at(x + 1, y - 1);
    const float pzp = p.at(x + 1, y), nzp = p.at(x - 1, y);
    const float zpp = p.at(x, y + 1), znp = p.at(x, y - 1);
    const float pzn = n.at(x + 1, y), nzn = n.at(x - 1, y);
    const float zpn = n.at(x, y + 1), znn = n.at(x, y - 1);

    const float dxx = (pzz + nzz) - (2 * zzz);
    const float dyy = (zpz + znz) - (2 * zzz);
    const float dss = (zzp + zzn) - (2 * zzz);
    const float dxy = (((ppz - npz) - pnz) + nnz) * 0.25f;
    const float dxs = (((pzp - nzp) - pzn) + nzn) * 0.25f;
    const float dys = (((zpp - znp) - zpn) + znn) * 0.25f;

    const Vec3 x0 = {dxx, dxy, dxs}, x1 = {dxy, dyy, dys}, x2 = {dxs, dys, dss};
    const Vec3 c0 = cross(x1, x2), c1 = cross(x2, x0), c2 = cross(x0, x1);
    const float d = ((x0.x * c0.x) + (x0.y * c0.y)) + (x0.z * c0.z);
    const float inv = 1.0f / d;
    // Hi = -1.0 * ((1/d) * cp), columns c0, c1, c2
    const Vec3 h0 = {-(inv * c0.x), -(inv * c0.y), -(inv * c0.z)};
    const Vec3 h1 = {-(inv * c1.x), -(inv * c1.y), -(inv * c1.z)};
    const Vec3 h2 = {-(inv * c2.x), -(inv * c2.y), -(inv * c2.z)};
    const Vec3 dD = {(pzz - nzz) * 0.5f, (zpz - znz) * 0.5f, (zzp - zzn) * 0.5f};
    // matrix × vector: columns scaled by the vector's components, summed left to right
    return {((h0.x * dD.x) + (h1.x * dD.y)) + (h2.x * dD.z),
            ((h0.y * dD.x) + (h1.y * dD.y)) + (h2.y * dD.z),
            ((h0.z * dD.x) + (h1.z * dD.y)) + (h2.z * dD.z)};
}

// SIFTInterpolate.metal:17-61 isOnEdge.
inline bool isOnEdge(const Plane& t, int x, int y, float edgeThreshold) {
    const float v = t.at(x, y);
    const float zn = t.at(x, y - 1), zp = t.at(x, y + 1);
    const float pz = t.at(x + 1, y), nz = t.at(x - 1, y);
    const float pp = t.at(x + 1, y + 1), np = t.at(x - 1, y + 1);
    const float pn = t.at(x + 1, y - 1), nn = t.at(x - 1, y - 1);
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

// SIFTInterpolate.metal:180-190 outOfBounds.
inline bool outOfBounds(int x, int y, int s, int w, int h, int scales, int border) {
    return x < border || x > w - border - 1 || y < border || y > h - border - 1 || s < 1 ||
           s > scales;
}

// SIFTInterpolate.metal:193-300 siftInterpolate + SIFTOctave.interpolateKeypoints
// (SIFTOctave.swift:205-288) for one candidate. Returns false when the reference leaves
// converged == 0. `stage` reports how far it got (for the IPOL stage-count pins).
bool siftInterpolate(const Octave& o, int octaveIndex, const Candidate& c, const SiftConfig& cfg,
                     float log2SigmaRatio, SiftKeypoint& out, int& stage) {
    stage = 0;
    float value = o.D[c.s].at(c.x, c.y);
    if (fabsf(value) <= cfg.dog_threshold * 0.8f) return false;
    int x = c.x, y = c.y, s = c.s;
    const int w = o.w, h = o.h;
    if (outOfBounds(x, y, s, w, h, SIFT_SCALES_PER_OCTAVE, cfg.image_border)) return false;
    bool converged = false;
    Vec3 alpha = {0, 0, 0};
    int i = 0;
    while (i < cfg.max_interpolation_iterations) {
        alpha = interpolationStep(o, x, y, s);
        if ((fabsf(alpha.x) < cfg.max_offset) && (fabsf(alpha.y) < cfg.max_offset) &&
            (fabsf(alpha.z) < cfg.max_offset)) {
            converged = true;
            break;
        }
        if (alpha.x > +cfg.max_offset) x += 1;
        if (alpha.x < -cfg.max_offset) x -= 1;
        if (alpha.y > +cfg.max_offset) y += 1;
        if (alpha.y < -cfg.max_offset) y -= 1;
        if (alpha.z > +cfg.max_offset) s += 1;
        if (alpha.z < -cfg.max_offset) s -= 1;
        if (outOfBounds(x, y, s, w, h, SIFT_SCALES_PER_OCTAVE, cfg.image_border)) return false;
        i += 1;
    }
    if (!converged) return false;
    stage = 1;
    // interpolateContrast (:90-100): only the x term of dD·alpha is used.
    const Vec3 dD = derivatives3D(o, x, y, s);
    value = o.D[s].at(x, y) + ((dD.x * alpha.x) * 0.5f);
    if (fabsf(value) <= cfg.dog_threshold) return false;
    stage = 2;
    if (isOnEdge(o.D[s], x, y, cfg.edge_threshold)) return false;
    stage = 3;
    out.octave = octaveIndex;
    out.scale = s;
    out.subScale = alpha.z;
    out.scaledX = x;
    out.scaledY = y;
    out.absoluteX = ((float)x + alpha.x) * o.delta;
    out.absoluteY = ((float)y + alpha.y) * o.delta;
    out.normalizedX = (float)x / (float)w;
    out.normalizedY = (float)y / (float)h;
    // SIFTOctave.swift:282 sigma = sigmas[scale] * pow(sigmaRatio, subScale); the spec evaluates
    // the power as exp2(subScale · log2(sigmaRatio)), log2 taken once on the host.
    out.sigma = o.sigmas[s] * om_exp2f(alpha.z * log2SigmaRatio);
    out.value = value;
    return true;
}

// SIFTOctave.getKeypointOrientations host filter (SIFTOctave.swift:303-329).
inline bool orientationBorderReject(const SiftKeypoint& k, const Octave& o, float lambda) {
    const float minX = 1.0f, minY = 1.0f;
    const float maxX = (float)(o.w - 2), maxY = (float)(o.h - 2);
    const float x = k.absoluteX / o.delta;
    const float y = k.absoluteY / o.delta;
    const float sigma = k.sigma / o.delta;
    const float r = ceilf((3 * lambda) * sigma);
    if (floorf(x - r) < minX) return true;
    if (ceilf(x + r) > maxX) return true;
    if (floorf(y - r) < minY) return true;
    if (ceilf(y + r) > maxY) return true;
    return false;
}

// SIFTOrientation.metal:140-175 siftOrientation for one keypoint: histogram (:88-136), six
// smoothing passes (:67-85), principal orientations (:31-64). Inputs are packed as
// SIFTOctave.swift:331-337 does: Int32(absoluteCoordinate) — truncation in input pixels.
// A sample outside the plane cannot occur after the host border filter; if it did it would
// contribute nothing (oracle-defined).
int siftOrientation(const Octave& o, const SiftKeypoint& k, const SiftConfig& cfg,
                    float* orientations) {
    const int bins = SIFT_ORIENTATION_HISTOGRAM_BINS;
    const float tau = 2 * 3.14159265358979323846f;
    const float lambda = cfg.lambda_orientation;
    const int absoluteX = (int32_t)k.absoluteX, absoluteY = (int32_t)k.absoluteY;
    const std::vector<float>& g = o.grad[k.scale];
    float histogram[bins];
    for (int i = 0; i < bins; i++) histogram[i] = 0;

    const int x = (int)roundf((float)absoluteX / o.delta);
    const int y = (int)roundf((float)absoluteY / o.delta);
    const float sigma = k.sigma / o.delta;
    const float exponentDenominator = (2.0f * lambda) * lambda;
    const int r = (int)ceilf((3 * lambda) * sigma);
    for (int j = -r; j <= r; j++) {
        for (int i = -r; i <= r; i++) {
            const float u = (float)i / sigma;
            const float v = (float)j / sigma;
            const float r2 = (u * u) + (v * v);
            const float w = om_expf(-r2 / exponentDenominator);
            const int sx = x + i, sy = y + j;
            if (sx < 0 || sx >= o.w || sy < 0 || sy >= o.h) continue;
            const float orientation = g[((size_t)sy * o.w + sx) * 2 + 0];
            const float magnitude = g[((size_t)sy * o.w + sx) * 2 + 1];
            const float t = orientation / tau;
            int bin = (int)roundf(t * (float)bins);
            if (bin < 0) bin += bins;
            if (bin >= bins) bin -= bins;
            histogram[bin] += w * magnitude;
        }
    }
    // smoothHistogram
    for (int it = 0; it < cfg.orientation_smoothing_iterations; it++) {
        float temp[bins];
        for (int i = 0; i < bins; i++) temp[i] = histogram[i];
        for (int i = 0; i < bins; i++) {
            const float h0 = temp[((i - 1) + bins) % bins];
            const float h1 = temp[i];
            const float h2 = temp[(i + 1) % bins];
            histogram[i] = ((h0 + h1) + h2) / 3.0f;
        }
    }
    // getPrincipalOrientations
    float maximum = (float)INT32_MIN;
    for (int i = 0; i < bins; i++) maximum = fmaxf(maximum, histogram[i]);
    const float threshold = cfg.orientation_threshold * maximum;
    int count = 0;
    for (int i = 0; i < bins; i++) {
        const float hm = histogram[((i - 1) + bins) % bins];
        const float h0 = histogram[i];
        const float hp = histogram[(i + 1) % bins];
        if ((h0 > threshold) && (h0 > hm) && (h0 > hp)) {
            // interpolatePeak (:31-33) and orientationFromBin (:16-28)
            const float offset = (hm - hp) / (2 * ((hm + hp) - (2 * h0)));
            const float t = ((float)i + offset) / (float)bins;
            float orientation = t * tau;
            if (orientation < 0) orientation += tau;
            if (orientation >= tau) orientation -= tau;
            orientations[count++] = orientation;
        }
    }
    return count;
}

// SIFTDescriptor.metal:53-79 addValue / :82-117 addFeature.
inline void addValue(float* patch, int x, int y, int b, float value) {
    const int side = 4, bins = SIFT_DESCRIPTOR_ORIENTATION_BINS;
    if ((x < 0) || (x >= side) || (y < 0) || (y >= side)) return;
    if (b < 0) b += bins;
    if (b >= bins) b -= bins;
    patch[(y * side * bins) + (x * bins) + b] += value;
}

inline void addFeature(float* patch, float x, float y, float b, float value) {
    const int fx = (int)floorf(x), cx = (int)ceilf(x);
    const int fy = (int)floorf(y), cy = (int)ceilf(y);
    const int ba = (int)floorf(b), bb = (int)ceilf(b);
    const float iMax = x - floorf(x), iMin = 1 - iMax;
    const float jMax = y - floorf(y), jMin = 1 - jMax;
    const float bMax = b - floorf(b), bMin = 1 - bMax;
    addValue(patch, fx, fy, ba, ((iMin * jMin) * bMin) * value);
    addValue(patch, fx, fy, bb, ((iMin * jMin) * bMax) * value);
    addValue(patch, cx, fy, ba, ((iMax * jMin) * bMin) * value);
    addValue(patch, cx, fy, bb, ((iMax * jMin) * bMax) * value);
    addValue(patch, cx, cy, ba, ((iMax * jMax) * bMin) * value);
    addValue(patch, cx, cy, bb, ((iMax * jMax) * bMax) * value);
    addValue(patch, fx, cy, ba, ((iMin * jMax) * bMin) * value);
    addValue(patch, fx, cy, bb, ((iMin * jMax) * bMax) * value);
}

// SIFTDescriptor.metal:15-27 normalizeFeatures. An all-zero patch (1/sqrt(0)) is undefined in
// the reference; oracle-defined: stays all-zero.
inline void normalizeFeatures(float* f) {
    float magnitude = 0;
    for (int i = 0; i < 128; i++) magnitude += (f[i] * f[i]);
    if (magnitude == 0) return;
    const float d = 1.0f / sqrtf(magnitude);
    for (int i = 0; i < 128; i++) f[i] *= d;
}

// SIFTDescriptor.metal:120-237 siftDescriptors for one (keypoint, theta), inputs packed as
// SIFTOctave.swift:410-424 (Int32 truncation of the absolute coordinate). The reference reads
// the gradient texture at ushort2(px + j, py + i) with no bounds check (:202; the guard at
// :168-184 is commented out) — undefined in Metal. Oracle-defined: a sample whose float
// coordinate is < 0 or whose truncation is >= the plane size contributes nothing.
void siftDescriptor(const Octave& o, const SiftKeypoint& k, float theta, uint8_t* out) {
    const int d = 4, bins = SIFT_DESCRIPTOR_ORIENTATION_BINS;
    const std::vector<float>& g = o.grad[k.scale];
    const float px = (float)(int32_t)k.absoluteX / o.delta;
    const float py = (float)(int32_t)k.absoluteY / o.delta;
    const float tau = 2 * 3.14159265358979323846f;
    float cosT, sinT;
    om_sincosf(theta, &sinT, &cosT);
    const float binsPerRadian = (float)bins / tau;
    const float exponentDenominator = (float)(d * d) * 0.5f;
    const float interval = (float)k.scale + k.subScale;
    const float intervals = (float)SIFT_SCALES_PER_OCTAVE;
    const float sigma = 1.6f;
    const float scale = sigma * om_exp2f(interval / intervals);
    const float histogramWidth = 3.0f * scale;
    const int radius = (int)((((histogramWidth * sqrtf(2.0f)) * ((float)d + 1.0f)) * 0.5f) + 0.5f);

    float features[128];
    for (int i = 0; i < 128; i++) features[i] = 0;
    for (int j = -radius; j <= +radius; j++) {
        for (int i = -radius; i <= +radius; i++) {
            const float rx = (((float)j * cosT) - ((float)i * sinT)) / histogramWidth;
            const float ry = (((float)j * sinT) + ((float)i * cosT)) / histogramWidth;
            const float bx = (rx + (float)(d / 2)) - 0.5f;
            const float by = (ry + (float)(d / 2)) - 0.5f;
            const float cxf = px + (float)j, cyf = py + (float)i;
            if (cxf < 0 || cyf < 0) continue;
            const int sx = (int)cxf, sy = (int)cyf;
            if (sx >= o.w || sy >= o.h) continue;
            float orientation = g[((size_t)sy * o.w + sx) * 2 + 0] - theta;
            const float magnitude = g[((size_t)sy * o.w + sx) * 2 + 1];
            while (orientation < 0) orientation += tau;
            while (orientation >= tau) orientation -= tau;
            const float bin = orientation * binsPerRadian;
            const float exponentNumerator = (rx * rx) + (ry * ry);
            const float w = om_expf(-exponentNumerator / exponentDenominator);
            const float value = magnitude * w;
            addFeature(features, bx, by, bin, value);
        }
    }
    normalizeFeatures(features);
    for (int i = 0; i < 128; i++) features[i] = fminf(features[i], 0.2f);  // thresholdFeatures
    normalizeFeatures(features);
    for (int i = 0; i < 128; i++)                                           // quantizeFeatures
        out[i] = (uint8_t)(int)fminf(255.0f, features[i] * 512.0f);
}

// DifferenceOfGaussians.init (DifferenceOfGaussians.swift:233-344) + Octave.init (:69-147).
void setup(Oracle& o, const SiftConfig& cfg) {
    o.cfg = cfg;
    o.W = cfg.width;
    o.H = cfg.height;
    const float sigmaMinimum = 0.8f, deltaMinimum = 0.5f, sigmaInput = 0.5f;
    o.seedW = (int)((float)o.W / deltaMinimum);
    o.seedH = (int)((float)o.H / deltaMinimum);
    {
        const float i = sigmaMinimum * sigmaMinimum;
        const float j = sigmaInput * sigmaInput;
        o.seedSigma = sqrtf(i - j) / deltaMinimum;
        o.seedWeights = gaussianWeights(o.seedSigma);
    }
    o.gray.resize(o.W, o.H);
    o.scaled.resize(o.seedW, o.seedH);
    o.seed.resize(o.seedW, o.seedH);
    for (int oc = 0; oc < SIFT_NUM_OCTAVES; oc++) {
        Octave& q = o.oct[oc];
        q.delta = deltaMinimum * powf(2, (float)oc);
        q.w = (int)((float)o.W / q.delta);
        q.h = (int)((float)o.H / q.delta);
        for (int s = 0; s < SIFT_NUM_GAUSSIANS; s++) {
            const float h = q.delta / deltaMinimum;
            const float i = (float)s / (float)SIFT_SCALES_PER_OCTAVE;
            const float j = powf(2, i);
            q.sigmas[s] = (h * sigmaMinimum) * j;
        }
        for (int s = 0; s < SIFT_NUM_GAUSSIANS; s++) q.G[s].resize(q.w, q.h);
        for (int s = 0; s < SIFT_NUM_DOGS; s++) q.D[s].resize(q.w, q.h);
        if (oc == 0) {
            for (int s = 1; s < SIFT_NUM_GAUSSIANS; s++) {
                const float sa = q.sigmas[s - 1], sb = q.sigmas[s];
                o.rho[s - 1] = sqrtf((sb * sb) - (sa * sa)) / q.delta;
                o.weights[s - 1] = gaussianWeights(o.rho[s - 1]);
            }
        }
    }
}

// SIFT.getKeypoints (SIFT.swift:147-202): findKeypoints → getKeypointsFromOctaves →
// interpolateKeypoints.
void detect(Oracle& o, const uint8_t* bgra, int pitch) {
    // DifferenceOfGaussians.encodeSeedTexture (:357-389)
    convertSRGBToGrayscale(bgra, pitch, o.gray);
    bilinearUpScale(o.gray, o.scaled);
    gaussianBlur(o.scaled, o.seed, o.seedWeights);
    // encodeOctaves (:391-406) + Octave.encode (:149-157)
    for (int oc = 0; oc < SIFT_NUM_OCTAVES; oc++) {
        Octave& q = o.oct[oc];
        if (q.w < 1 || q.h < 1) continue;
        if (oc == 0) q.G[0].d = o.seed.d;  // blit copy (:176-188)
        else nearestNeighborDownScale(o.oct[oc - 1].G[SIFT_SCALES_PER_OCTAVE], q.G[0]);  // :192-199
        for (int s = 0; s < SIFT_NUM_GAUSSIANS - 1; s++) {
            // rho is the same for every octave (sigma and delta both double), so the octave-0
            // tables serve all of them, exactly as each reference octave recomputes equal values.
            gaussianBlur(q.G[s], q.G[s + 1], o.weights[s]);
        }
        for (int s = 0; s < SIFT_NUM_DOGS; s++) subtract(q.G[s + 1], q.G[s], q.D[s]);
    }
    // SIFTOctave.encode (:177-196): extrema + gradients
    o.keypoints.clear();
    for (int oc = 0; oc < SIFT_NUM_OCTAVES; oc++) {
        Octave& q = o.oct[oc];
        o.stats[oc] = Stats();
        q.candidates.clear();
        o.keypointCounts[oc] = 0;
        if (q.w < 3 || q.h < 3) continue;
        siftExtremaList(q, o.cfg.dog_threshold, o.stats[oc], o.collectStats);
        for (int s = 0; s < SIFT_NUM_GAUSSIANS; s++) {
            if (o.allGradients || (s >= 1 && s <= SIFT_SCALES_PER_OCTAVE)) siftGradient(q.G[s], q.grad[s]);
        }
        // interpolateKeypoints (SIFTOctave.swift:205-288)
        const float sigmaRatio = q.sigmas[1] / q.sigmas[0];
        const float log2SigmaRatio = log2f(sigmaRatio);
        const int n = (int)q.candidates.size();
        std::vector<SiftKeypoint> kp(n);
        std::vector<int> stage(n);
        std::vector<uint8_t> ok(n);
#pragma omp parallel for schedule(dynamic, 64)
        for (int i = 0; i < n; i++)
            ok[i] = siftInterpolate(q, oc, q.candidates[i], o.cfg, log2SigmaRatio, kp[i], stage[i]);
        for (int i = 0; i < n; i++) {
            if (stage[i] >= 1) o.stats[oc].interp++;
            if (stage[i] >= 2) o.stats[oc].contrast++;
            if (ok[i]) {
                o.stats[oc].final_++;
                o.keypoints.push_back(kp[i]);
                o.keypointCounts[oc]++;
            }
        }
    }
    o.detected = true;
}

// SIFT.getDescriptors (SIFT.swift:207-238) over caller-supplied keypoints grouped by octave.
void describe(Oracle& o, const SiftKeypoint* kps, const int32_t* counts) {
    o.descriptors.clear();
    int base = 0;
    for (int oc = 0; oc < SIFT_NUM_OCTAVES; oc++) {
        Octave& q = o.oct[oc];
        const int n = counts[oc];
        o.descriptorCounts[oc] = 0;
        // getKeypointOrientations (SIFTOctave.swift:290-382)
        std::vector<int> nori(n, 0);
        std::vector<float> ori((size_t)n * SIFT_ORIENTATION_HISTOGRAM_BINS);
#pragma omp parallel for schedule(dynamic, 16)
        for (int i = 0; i < n; i++) {
            const SiftKeypoint& k = kps[base + i];
            if (orientationBorderReject(k, q, o.cfg.lambda_orientation)) continue;
            nori[i] = siftOrientation(q, k, o.cfg, &ori[(size_t)i * SIFT_ORIENTATION_HISTOGRAM_BINS]);
        }
        // getDescriptors (SIFTOctave.swift:384-492): one input per (keypoint, orientation)
        std::vector<int> first(n + 1, 0);
        for (int i = 0; i < n; i++) first[i + 1] = first[i] + nori[i];
        const int nd = first[n];
        const size_t d0 = o.descriptors.size();
        o.descriptors.resize(d0 + nd);
#pragma omp parallel for schedule(dynamic, 8)
        for (int i = 0; i < n; i++) {
            for (int t = 0; t < nori[i]; t++) {
                SiftDescriptor& dsc = o.descriptors[d0 + first[i] + t];
                dsc.keypoint = base + i;
                dsc.theta = ori[(size_t)i * SIFT_ORIENTATION_HISTOGRAM_BINS + t];
                siftDescriptor(q, kps[base + i], dsc.theta, dsc.features);
            }
        }
        o.descriptorCounts[oc] = nd;
        base += n;
    }
}

}  // namespace

// ---------------------------------------------------------------------------------------------
// C interface for tests / bench (ctypes). Not part of the product ABI.
extern "C" {

void* oracle_create(const SiftConfig* cfg) {
    if (!cfg || cfg->width < 2 || cfg->height < 2) return nullptr;
    Oracle* o = new Oracle();
    setup(*o, *cfg);
    return o;
}

void oracle_destroy(void* h) { delete (Oracle*)h; }

void oracle_set_options(void* h, int collect_stats, int all_gradients) {
    Oracle* o = (Oracle*)h;
    o->collectStats = collect_stats != 0;
    o->allGradients = all_gradients != 0;
}

void oracle_set_threads(int n) {
#ifdef _OPENMP
    omp_set_num_threads(n > 0 ? n : omp_get_num_procs());
#else
    (void)n;
#endif
}

int oracle_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

int oracle_detect(void* h, const uint8_t* bgra, int pitch) {
    Oracle* o = (Oracle*)h;
    detect(*o, bgra, pitch);
    return (int)o->keypoints.size();
}

int oracle_describe(void* h) {
    Oracle* o = (Oracle*)h;
    if (!o->detected) return -1;
    describe(*o, o->keypoints.data(), o->keypointCounts);
    return (int)o->descriptors.size();
}

int oracle_describe_keypoints(void* h, const SiftKeypoint* kps, const int32_t* counts) {
    Oracle* o = (Oracle*)h;
    if (!o->detected) return -1;
    describe(*o, kps, counts);
    return (int)o->descriptors.size();
}

int oracle_get_keypoints(void* h, SiftKeypoint* dst, int32_t* counts) {
    Oracle* o = (Oracle*)h;
    if (dst) memcpy(dst, o->keypoints.data(), o->keypoints.size() * sizeof(SiftKeypoint));
    if (counts) memcpy(counts, o->keypointCounts, sizeof(o->keypointCounts));
    return (int)o->keypoints.size();
}

int oracle_get_descriptors(void* h, SiftDescriptor* dst, int32_t* counts) {
    Oracle* o = (Oracle*)h;
    if (dst) memcpy(dst, o->descriptors.data(), o->descriptors.size() * sizeof(SiftDescriptor));
    if (counts) memcpy(counts, o->descriptorCounts, sizeof(o->descriptorCounts));
    return (int)o->descriptors.size();
}

int oracle_get_candidates(void* h, int octave, int32_t* dst_xyz, int capacity) {
    Oracle* o = (Oracle*)h;
    const auto& c = o->oct[octave].candidates;
    const int n = (int)c.size();
    for (int i = 0; i < n && i < capacity; i++) {
        dst_xyz[3 * i + 0] = c[i].x;
        dst_xyz[3 * i + 1] = c[i].y;
        dst_xyz[3 * i + 2] = c[i].s;
    }
    return n;
}

// stats per octave: raw25, raw26, soft, interp, contrast, final
void oracle_get_stats(void* h, int64_t* dst) {
    Oracle* o = (Oracle*)h;
    for (int oc = 0; oc < SIFT_NUM_OCTAVES; oc++) {
        const Stats& s = o->stats[oc];
        int64_t v[6] = {s.raw25, s.raw26, s.soft, s.interp, s.contrast, s.final_};
        memcpy(dst + 6 * oc, v, sizeof(v));
    }
}

// same `what` codes as sift_debug_download
int oracle_get_plane(void* h, int what, int octave, int slice, float* dst) {
    Oracle* o = (Oracle*)h;
    const std::vector<float>* src = nullptr;
    switch (what) {
        case SIFT_PLANE_GRAY: src = &o->gray.d; break;
        case SIFT_PLANE_SEED: src = &o->seed.d; break;
        case SIFT_PLANE_GAUSSIAN: src = &o->oct[octave].G[slice].d; break;
        case SIFT_PLANE_DOG: src = &o->oct[octave].D[slice].d; break;
        case SIFT_PLANE_GRADIENT: src = &o->oct[octave].grad[slice]; break;
        default: return -1;
    }
    memcpy(dst, src->data(), src->size() * sizeof(float));
    return (int)src->size();
}

void oracle_get_info(void* h, SiftInfo* info) {
    Oracle* o = (Oracle*)h;
    memset(info, 0, sizeof(*info));
    info->width = o->W;
    info->height = o->H;
    info->max_batch = 1;
    for (int oc = 0; oc < SIFT_NUM_OCTAVES; oc++) {
        info->octave_width[oc] = o->oct[oc].w;
        info->octave_height[oc] = o->oct[oc].h;
        info->octave_pitch[oc] = o->oct[oc].w;
        info->octave_delta[oc] = o->oct[oc].delta;
        for (int s = 0; s < SIFT_NUM_GAUSSIANS; s++) info->sigmas[oc][s] = o->oct[oc].sigmas[s];
    }
    info->seed_sigma = o->seedSigma;
    info->seed_taps = (int)o->seedWeights.size();
    for (size_t i = 0; i < o->seedWeights.size(); i++) info->seed_weights[i] = o->seedWeights[i];
    for (int s = 0; s < SIFT_NUM_GAUSSIANS - 1; s++) {
        info->rho[s] = o->rho[s];
        info->taps[s] = (int)o->weights[s].size();
        for (size_t i = 0; i < o->weights[s].size(); i++) info->weights[s][i] = o->weights[s][i];
    }
}

// op codes as sift_debug_math
void oracle_math(int op, const float* a, const float* b, float* out, int64_t n) {
    for (int64_t i = 0; i < n; i++) {
        float s, c;
        switch (op) {
            case 0: out[i] = om_expf(a[i]); break;
            case 1: out[i] = om_atan2f(a[i], b[i]); break;
            case 2: om_sincosf(a[i], &s, &c); out[i] = s; break;
            case 3: om_sincosf(a[i], &s, &c); out[i] = c; break;
            case 4: out[i] = om_exp2f(a[i]); break;
            default: out[i] = 0;
        }
    }
}

// SIFTDescriptor.match(source:target:absoluteThreshold:relativeThreshold:)
// (SIFTDescriptor.swift:298-361): for every source descriptor a linear scan over the targets in
// order, `if distance < best { second = best; best = distance; match = t }` (:339-343) — so
// `second` is the best seen BEFORE the last improvement, not the second smallest distance, and it
// starts as .greatestFiniteMagnitude (:332-333); kept iff best < absoluteThreshold and
// best < second * relativeThreshold (:353-359). distance = sqrt(distanceSquared) (Vector.swift:
// 237-239) of indexValue = features / 255 re-ordered cell by cell (SIFTDescriptor.swift:37-80; a
// permutation, it does not change a distance).
//
// Oracle-defined where the reference is unspecified: distanceSquared is vDSP.distanceSquared
// (Accelerate, closed source, summation order and vector width unspecified, Vector.swift:234) on
// floats that are themselves rounded quotients f / 255. Every feature is an integer 0...255, so
// the exact value is sum((b - a)^2) / 255^2 with an integer numerator below 2^24: the oracle
// compares candidates by that exact integer (the order any correctly rounded evaluation gives
// except for ties in the last bit) and evaluates the two threshold tests on
// distance = sqrtf(float(numerator)) / 255.0f, single IEEE operations.
// Output: kept correspondences in source order; returns their number.
int64_t oracle_match(const uint8_t* source, int64_t nSource, const uint8_t* target, int64_t nTarget,
                     float absoluteThreshold, float relativeThreshold, SiftMatch* out) {
    std::vector<SiftMatch> rows((size_t)std::max<int64_t>(nSource, 1));
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < nSource; i++) {
        const uint8_t* a = source + i * 128;
        const int32_t none = INT32_MAX;
        int32_t best = none, second = none;
        int64_t match = -1;
        for (int64_t j = 0; j < nTarget; j++) {
            const uint8_t* b = target + j * 128;
            int32_t d2 = 0;
            for (int k = 0; k < 128; k++) {
                const int32_t d = (int32_t)b[k] - (int32_t)a[k];
                d2 += d * d;
            }
            if (d2 < best) {
                second = best;
                best = d2;
                match = j;
            }
        }
        SiftMatch r;
        r.source = (int32_t)i;
        r.target = -1;
        r.distance = 0.0f;
        if (match >= 0) {   // `guard let bestMatch`, `guard let secondBestMatchDistance` (:346-352)
            const float dBest = std::sqrt((float)best) / 255.0f;
            const float dSecond = second == none ? 3.402823466e+38f : std::sqrt((float)second) / 255.0f;
            r.distance = dBest;
            if (dBest < absoluteThreshold && dBest < (dSecond * relativeThreshold)) r.target = (int32_t)match;
        }
        rows[(size_t)i] = r;
    }
    int64_t n = 0;
    for (int64_t i = 0; i < nSource; i++)
        if (rows[(size_t)i].target >= 0) out[n++] = rows[(size_t)i];
    return n;
}

// ---- SIFTDescriptor.matchGeometry / compareGeometry (SIFTDescriptor.swift:104-296) --------------
// Geometric consistency score of the first 80 brute-force matches: for consecutive quadruples of
// matches the vector m0->m1 is compared with m2->m3 in the source and in the target image (length
// ratio and angle), the squared similarities are averaged after dropping outliers beyond two
// standard deviations. Coordinates are keypoint.absoluteCoordinate (makeCoordinate, :145-159).
// Oracle-defined: simd_length = sqrtf(x*x + y*y), simd_normalize = v / length, simd_dot =
// a.x*b.x + a.y*b.y, each a single IEEE operation in this order (Apple's simd library leaves the
// rounding of its fast paths unspecified).
namespace {
inline float geoClamp01(float v) { return v < 0.0f ? 0.0f : (v > 1.0f ? 1.0f : v); }
}  // namespace

float oracle_compare_geometry(const SiftMatch* matches, int64_t nMatches, const float* sourceXY,
                              const float* targetXY, int minimumSampleSize) {
    const float minimumLength = 2;
    float sum = 0;
    int count = 0;
    std::vector<float> scores;
    auto coord = [](const float* xy, int32_t i, float& x, float& y) { x = xy[2 * i]; y = xy[2 * i + 1]; };
    for (int64_t i = 0; i < nMatches - 3; i++) {   // stride(from: 0, to: matches.count - 3, by: 1)
        const SiftMatch &m0 = matches[i], &m1 = matches[i + 1], &m2 = matches[i + 2], &m3 = matches[i + 3];
        float ax, ay, bx, by;
        coord(sourceXY, m1.source, ax, ay); coord(sourceXY, m0.source, bx, by);
        const float sbx = ax - bx, sby = ay - by;
        coord(targetXY, m1.target, ax, ay); coord(targetXY, m0.target, bx, by);
        const float tbx = ax - bx, tby = ay - by;
        const float sourceBaseLength = std::sqrt((sbx * sbx) + (sby * sby));
        const float targetBaseLength = std::sqrt((tbx * tbx) + (tby * tby));
        if (!(sourceBaseLength >= minimumLength)) continue;
        if (!(targetBaseLength >= minimumLength)) continue;
        const float sbnx = sbx / sourceBaseLength, sbny = sby / sourceBaseLength;
        const float tbnx = tbx / targetBaseLength, tbny = tby / targetBaseLength;
        coord(sourceXY, m3.source, ax, ay); coord(sourceXY, m2.source, bx, by);
        const float stx = ax - bx, sty = ay - by;
        coord(targetXY, m3.target, ax, ay); coord(targetXY, m2.target, bx, by);
        const float ttx = ax - bx, tty = ay - by;
        const float sourceTestLength = std::sqrt((stx * stx) + (sty * sty));
        const float targetTestLength = std::sqrt((ttx * ttx) + (tty * tty));
        if (!(sourceTestLength >= minimumLength)) continue;
        if (!(targetTestLength >= minimumLength)) continue;
        const float stnx = stx / sourceTestLength, stny = sty / sourceTestLength;
        const float ttnx = ttx / targetTestLength, ttny = tty / targetTestLength;
        const float sourceRatio = sourceTestLength / sourceBaseLength;
        const float targetRatio = targetTestLength / targetBaseLength;
        // dotProduct (:161-163): clamp(dot * 0.5 + 0.5, 0, 1)
        const float sourceDot = geoClamp01((((stnx * sbnx) + (stny * sbny)) * 0.5f) + 0.5f);
        const float targetDot = geoClamp01((((ttnx * tbnx) + (ttny * tbny)) * 0.5f) + 0.5f);
        const float orientationSimilarity = 1.0f - std::fabs(sourceDot - targetDot);
        float scaleSimilarity;
        if (sourceRatio < targetRatio) scaleSimilarity = geoClamp01(sourceRatio / targetRatio);
        else scaleSimilarity = geoClamp01(targetRatio / sourceRatio);
        const float similarity = orientationSimilarity * scaleSimilarity;
        const float score = similarity * similarity;
        scores.push_back(score);
        sum += score;
        count += 1;
    }
    if (count < minimumSampleSize) return 0;
    const float mean = sum / (float)count;
    float error = 0;
    for (float score : scores) {
        const float delta = score - mean;
        error += (delta * delta);
    }
    const float variance = error / (float)(count - 1);
    const float standardDeviation = std::sqrt(variance);
    float fairMeanSum = 0, fairMeanCount = 0;
    for (float score : scores) {
        const float zscore = std::fabs((score - mean) / standardDeviation);
        if (zscore <= 2) {   // NaN (all scores equal: 0 / 0) compares false, as in Swift
            fairMeanSum += score;
            fairMeanCount += 1;
        }
    }
    return fairMeanSum / fairMeanCount;
}

// matchGeometry (:104-143): brute-force matches (defaults 1.176 / 0.6), at least 7 of them, the
// first 80 go to compareGeometry.
float oracle_match_geometry(const uint8_t* source, const float* sourceXY, int64_t nSource, const uint8_t* target,
                            const float* targetXY, int64_t nTarget, float absoluteThreshold,
                            float relativeThreshold) {
    const int minimumSampleSize = 7;
    const int maximumSampleSize = std::max(minimumSampleSize, 80);
    std::vector<SiftMatch> m((size_t)std::max<int64_t>(nSource, 1));
    const int64_t n = oracle_match(source, nSource, target, nTarget, absoluteThreshold, relativeThreshold, m.data());
    if (n < minimumSampleSize) return 0;
    return oracle_compare_geometry(m.data(), std::min<int64_t>(n, maximumSampleSize), sourceXY, targetXY,
                                   minimumSampleSize);
}

// ---- Trie ANN (Utilities/Trie.swift) + SIFTDescriptor.approximateMatch (SIFTDescriptor.swift:362-417) ----
// Restated as the reference builds it: a recursive 8-ary trie keyed by the 16 components of
// indexKey (one per histogram cell, in the reference's re-ordered cell sequence
// SIFTDescriptor.swift:55-77), values in the leaves in insertion order, leaves linked left /
// right in depth-first bin order (Trie.swift:119-128,142-157), nearest(key:query:radius:k:)
// visiting the key's leaf, then `radius` leaves to the left, then `radius` to the right
// (:229-253) with a FiniteQueue of capacity k that a value enters only when it beats the current
// best (:287-300).
// Oracle-defined: indexKey[c] = mean of the cell's 8 raw features f / 255 (vDSP.mean, rounding
// unspecified) only matters through binIndex = Int((value * 7).rounded()) (:313-320); with S the
// integer sum of the cell's 8 features that is round-half-away(7 S / 2040), evaluated exactly in
// integers. Distances are compared as exact integers as in oracle_match.
namespace {

const int kTrieCellOrder[16] = {5, 6, 9, 10, 0, 3, 12, 15, 1, 2, 4, 7, 8, 11, 13, 14};   // SIFTDescriptor.swift:57-77
const int kTrieBins = 8;

void trieKey(const uint8_t* f, int key[16]) {
    for (int c = 0; c < 16; c++) {
        int S = 0;
        for (int b = 0; b < 8; b++) S += f[kTrieCellOrder[c] * 8 + b];
        key[c] = (14 * S + 2040) / 4080;   // round-half-up of 7 S / 2040, S >= 0
    }
}

struct TrieNode {
    TrieNode* child[kTrieBins] = {};
    bool hasNodes = false;
    std::vector<int32_t> values;
    TrieNode *left = nullptr, *right = nullptr;
    ~TrieNode() { for (auto* c : child) delete c; }
};

void trieInsert(TrieNode* n, const int* key, int depth, int32_t value) {
    if (depth == 16) { n->values.push_back(value); return; }
    TrieNode*& c = n->child[key[depth]];
    if (!c) { c = new TrieNode(); n->hasNodes = true; }
    trieInsert(c, key, depth + 1, value);
}

void trieLeaves(TrieNode* n, std::vector<TrieNode*>& out) {
    if (n->hasNodes) { for (auto* c : n->child) if (c) trieLeaves(c, out); }
    else out.push_back(n);
}

int trieWrap(int input) {   // wrapBinIndex (:338-350)
    int output = input;
    const int n = kTrieBins - 1;
    if (output < 0) output += n;
    else if (output >= n) output -= n;
    return output;
}

TrieNode* trieClosest(TrieNode* n, int bin) {   // closestNode (:270-285)
    if (n->child[bin]) return n->child[bin];
    int bestDistance = INT32_MAX;
    TrieNode* best = nullptr;
    for (int j = 0; j < kTrieBins; j++) {
        if (!n->child[j]) continue;
        const int distance = trieWrap(std::abs(j - bin));   // binDifference (:303-309)
        if (distance < bestDistance) { bestDistance = distance; best = n->child[j]; }
    }
    return best;
}

struct TrieQueue {   // FiniteQueue<Match>(capacity: 2) (:201-226): newest first
    int32_t value[2];
    int32_t d2[2];
    int count = 0;
    void insert(int32_t v, int32_t d) {
        value[1] = value[0]; d2[1] = d2[0];
        value[0] = v; d2[0] = d;
        if (count < 2) count++;
    }
};

void trieNearestValue(const TrieNode* n, const uint8_t* query, const uint8_t* target, TrieQueue& q) {   // (:287-300)
    int32_t best = q.count ? q.d2[0] : INT32_MAX;
    for (int32_t v : n->values) {
        const uint8_t* b = target + (int64_t)v * 128;
        int32_t d2 = 0;
        for (int k = 0; k < 128; k++) { const int32_t d = (int32_t)b[k] - (int32_t)query[k]; d2 += d * d; }
        if (d2 < best) { best = d2; q.insert(v, d2); }
    }
}

}  // namespace

int64_t oracle_approximate_match(const uint8_t* source, int64_t nSource, const uint8_t* target, int64_t nTarget,
                                 float absoluteThreshold, float relativeThreshold, SiftMatch* out) {
    if (nTarget < 1) return 0;   // link() of an empty trie leaves the root unlinked; nothing can match
    TrieNode root;
    int key[16];
    for (int64_t j = 0; j < nTarget; j++) {
        trieKey(target + j * 128, key);
        trieInsert(&root, key, 0, (int32_t)j);
    }
    std::vector<TrieNode*> leaves;
    trieLeaves(&root, leaves);
    for (size_t i = 0; i < leaves.size(); i++) {   // link (:119-128)
        TrieNode* a = leaves[i];
        TrieNode* b = leaves[(i + 1) % leaves.size()];
        a->right = b;
        b->left = a;
    }
    int64_t n = 0;
    const int radius = 10;
    for (int64_t i = 0; i < nSource; i++) {
        const uint8_t* q = source + i * 128;
        trieKey(q, key);
        TrieNode* bin = &root;   // nearestNode (:255-268)
        for (int d = 0; d < 16; d++) {
            if (!bin->hasNodes) break;
            TrieNode* next = trieClosest(bin, key[d]);
            if (!next) break;
            bin = next;
        }
        TrieQueue queue;
        trieNearestValue(bin, q, target, queue);
        TrieNode* node = bin;
        for (int r = 0; r < radius; r++) { node = node->left; trieNearestValue(node, q, target, queue); }
        node = bin;
        for (int r = 0; r < radius; r++) { node = node->right; trieNearestValue(node, q, target, queue); }
        if (queue.count != 2) continue;   // guard matches.count == 2 (:394-396)
        const float dBest = std::sqrt((float)queue.d2[0]) / 255.0f;
        const float dSecond = std::sqrt((float)queue.d2[1]) / 255.0f;
        if (!(dBest < absoluteThreshold)) continue;
        if (!(dBest < (dSecond * relativeThreshold))) continue;
        SiftMatch m;
        m.source = (int32_t)i;
        m.target = queue.value[0];
        m.distance = dBest;
        out[n++] = m;
    }
    return n;
}

}  // extern "C"
