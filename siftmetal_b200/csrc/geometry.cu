// geometry.cu — host stages downstream of the matcher. In the reference these
// are CPU code as well (Swift): they consume a handful of correspondences or walk a pointer
// structure, so they stay on the host here; the distance work they sit on is the device matcher.
//
//   SIFTDescriptor.matchGeometry / compareGeometry   SIFTDescriptor.swift:104-296
//   Trie (approximate nearest neighbour)             Utilities/Trie.swift:76-416
//   SIFTDescriptor.approximateMatch                  SIFTDescriptor.swift:362-417
//
// Compiled with -ffp-contract=off: every float expression is evaluated as written.
#include <math.h>
#include <stdint.h>

#include <algorithm>
#include <numeric>
#include <vector>

#include "../../include/siftcuda.h"

namespace sift {

namespace {
inline float clamp01(float v) { return v < 0.0f ? 0.0f : (v > 1.0f ? 1.0f : v); }

struct Vec2 {
    float x, y;
};
inline Vec2 sub(const float* xy, int32_t a, int32_t b) { return Vec2{xy[2 * a] - xy[2 * b], xy[2 * a + 1] - xy[2 * b + 1]}; }
inline float length(Vec2 v) { return sqrtf((v.x * v.x) + (v.y * v.y)); }                  // simd_length
inline Vec2 normalize(Vec2 v, float len) { return Vec2{v.x / len, v.y / len}; }            // simd_normalize
inline float halfDot(Vec2 a, Vec2 b) { return clamp01((((a.x * b.x) + (a.y * b.y)) * 0.5f) + 0.5f); }   // dotProduct :161-163
}  // namespace

// compareGeometry (SIFTDescriptor.swift:165-296).
float compareGeometry(const SiftMatch* m, int64_t n, const float* sourceXY, const float* targetXY, int minimumSampleSize) {
    const float minimumLength = 2;
    std::vector<float> scores;
    float sum = 0;
    for (int64_t i = 0; i + 3 < n; i++) {
        const Vec2 sBase = sub(sourceXY, m[i + 1].source, m[i].source);
        const Vec2 tBase = sub(targetXY, m[i + 1].target, m[i].target);
        const float sBaseLen = length(sBase), tBaseLen = length(tBase);
        if (!(sBaseLen >= minimumLength) || !(tBaseLen >= minimumLength)) continue;
        const Vec2 sTest = sub(sourceXY, m[i + 3].source, m[i + 2].source);
        const Vec2 tTest = sub(targetXY, m[i + 3].target, m[i + 2].target);
        const float sTestLen = length(sTest), tTestLen = length(tTest);
        if (!(sTestLen >= minimumLength) || !(tTestLen >= minimumLength)) continue;
        const float sRatio = sTestLen / sBaseLen, tRatio = tTestLen / tBaseLen;
        const float sDot = halfDot(normalize(sTest, sTestLen), normalize(sBase, sBaseLen));
        const float tDot = halfDot(normalize(tTest, tTestLen), normalize(tBase, tBaseLen));
        const float orientationSimilarity = 1.0f - fabsf(sDot - tDot);
        const float scaleSimilarity = sRatio < tRatio ? clamp01(sRatio / tRatio) : clamp01(tRatio / sRatio);
        const float similarity = orientationSimilarity * scaleSimilarity;
        const float score = similarity * similarity;
        scores.push_back(score);
        sum += score;
    }
    const int count = (int)scores.size();
    if (count < minimumSampleSize) return 0;
    const float mean = sum / (float)count;
    float error = 0;
    for (float s : scores) {
        const float delta = s - mean;
        error += (delta * delta);
    }
    const float standardDeviation = sqrtf(error / (float)(count - 1));
    float fairSum = 0, fairCount = 0;
    for (float s : scores)
        if (fabsf((s - mean) / standardDeviation) <= 2) {   // z-score filter (:262-276)
            fairSum += s;
            fairCount += 1;
        }
    return fairSum / fairCount;
}

// ---- Trie over sorted keys ---------------------------------------------------------------------
// The reference's trie has constant height 16 and 8 bins per level; its leaves, linked in
// depth-first bin order, are exactly the distinct keys in lexicographic order, each holding its
// values in insertion order. So the structure is kept flat: target indices stably sorted by key,
// one Leaf per distinct key; a node of the reference = the range of leaves sharing a key prefix,
// its children = the sub-ranges by next digit (found by binary search).
namespace {

constexpr int kCellOrder[16] = {5, 6, 9, 10, 0, 3, 12, 15, 1, 2, 4, 7, 8, 11, 13, 14};   // SIFTDescriptor.swift:57-77
constexpr int kBins = 8, kDepth = 16;

// 16 digits of 3 bits, most significant first: integer order = lexicographic order of the keys.
// digit = binIndex(for: mean of the cell's raw features) = round-half-away(7 S / 2040), S = the
// integer sum of the cell's 8 features (Trie.swift:313-320; exact, see the oracle's note).
uint64_t packedKey(const uint8_t* f) {
    uint64_t k = 0;
    for (int c = 0; c < kDepth; c++) {
        int S = 0;
        for (int b = 0; b < 8; b++) S += f[kCellOrder[c] * 8 + b];
        k = (k << 3) | (uint64_t)((14 * S + 2040) / 4080);
    }
    return k;
}
inline int digit(uint64_t key, int depth) { return (int)((key >> (3 * (kDepth - 1 - depth))) & 7u); }

struct Leaf {
    uint64_t key;
    int32_t first, count;   // range in the sorted value array
};

inline int wrapBin(int v) {   // wrapBinIndex (Trie.swift:338-350)
    const int n = kBins - 1;
    return v < 0 ? v + n : (v >= n ? v - n : v);
}

inline int32_t distanceSquared(const uint8_t* a, const uint8_t* b) {
    int32_t d2 = 0;
    for (int k = 0; k < 128; k++) {
        const int32_t d = (int32_t)b[k] - (int32_t)a[k];
        d2 += d * d;
    }
    return d2;
}

}  // namespace

// SIFTDescriptor.approximateMatch(source:target:absoluteThreshold:relativeThreshold:): radius 10, k 2.
void approximateMatch(const uint8_t* source, int64_t nSource, const uint8_t* target, int64_t nTarget, float absThr,
                      float relThr, std::vector<SiftMatch>& out) {
    out.clear();
    if (nSource < 1 || nTarget < 1) return;
    std::vector<uint64_t> keys((size_t)nTarget);
    for (int64_t j = 0; j < nTarget; j++) keys[(size_t)j] = packedKey(target + j * 128);
    std::vector<int32_t> order((size_t)nTarget);
    std::iota(order.begin(), order.end(), 0);
    std::stable_sort(order.begin(), order.end(), [&](int32_t a, int32_t b) { return keys[(size_t)a] < keys[(size_t)b]; });
    std::vector<Leaf> leaves;
    for (int32_t i = 0; i < (int32_t)nTarget; i++) {
        const uint64_t k = keys[(size_t)order[(size_t)i]];
        if (leaves.empty() || leaves.back().key != k) leaves.push_back(Leaf{k, i, 0});
        leaves.back().count++;
    }
    const int nLeaves = (int)leaves.size();
    const int radius = 10;
    for (int64_t i = 0; i < nSource; i++) {
        const uint8_t* q = source + i * 128;
        const uint64_t qk = packedKey(q);
        // nearestNode (Trie.swift:255-268): at every level the child in the key's bin, else the
        // existing child with the smallest binDifference (first such in bin order)
        int lo = 0, hi = nLeaves;   // leaves sharing the prefix chosen so far
        for (int d = 0; d < kDepth; d++) {
            const int shift = 3 * (kDepth - 1 - d);
            int childLo[kBins + 1];
            int p = lo;
            for (int b = 0; b < kBins; b++) {   // sub-range of digit b: leaves are sorted, so a lower_bound per digit
                childLo[b] = p;
                int a = p, e = hi;
                while (a < e) {
                    const int mid = (a + e) / 2;
                    if ((int)((leaves[(size_t)mid].key >> shift) & 7u) <= b) a = mid + 1;
                    else e = mid;
                }
                p = a;
            }
            childLo[kBins] = hi;
            const int want = digit(qk, d);
            int pick = -1;
            if (childLo[want + 1] > childLo[want]) pick = want;
            else {
                int bestDistance = INT32_MAX;
                for (int b = 0; b < kBins; b++) {
                    if (childLo[b + 1] == childLo[b]) continue;
                    const int dist = wrapBin(std::abs(b - want));   // binDifference (:303-309)
                    if (dist < bestDistance) { bestDistance = dist; pick = b; }
                }
            }
            lo = childLo[pick];
            hi = childLo[pick + 1];
        }
        const int bin = lo;   // hi == lo + 1: one leaf
        // nearest (Trie.swift:229-253) with FiniteQueue(capacity: 2): newest first
        int32_t qv[2] = {-1, -1}, qd[2] = {0, 0};
        int qn = 0;
        auto visit = [&](int leaf) {   // nearestValue (:287-300)
            int32_t best = qn ? qd[0] : INT32_MAX;
            const Leaf& L = leaves[(size_t)leaf];
            for (int32_t t = 0; t < L.count; t++) {
                const int32_t v = order[(size_t)(L.first + t)];
                const int32_t d2 = distanceSquared(q, target + (int64_t)v * 128);
                if (d2 < best) {
                    best = d2;
                    qv[1] = qv[0]; qd[1] = qd[0];
                    qv[0] = v; qd[0] = d2;
                    if (qn < 2) qn++;
                }
            }
        };
        visit(bin);
        for (int r = 1; r <= radius; r++) visit(((bin - r) % nLeaves + nLeaves) % nLeaves);   // leftNode chain (circular)
        for (int r = 1; r <= radius; r++) visit((bin + r) % nLeaves);                          // rightNode chain
        if (qn != 2) continue;   // guard matches.count == 2 (SIFTDescriptor.swift:394-396)
        const float dBest = sqrtf((float)qd[0]) / 255.0f, dSecond = sqrtf((float)qd[1]) / 255.0f;
        if (!(dBest < absThr) || !(dBest < (dSecond * relThr))) continue;
        out.push_back(SiftMatch{(int32_t)i, qv[0], dBest});
    }
}

}  // namespace sift
