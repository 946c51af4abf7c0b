// SIFT.hpp — C++17 host-side mirror of lukevanin/SIFTMetal's public Swift API, header-only over
// the C ABI of include/siftcuda.h.
//
// The reference's host is Swift (Sources/SIFTMetal/SIFT/SIFT.swift); neither the build container
// nor the GPU boxes have a Swift toolchain, so the host side above the C ABI is offered in C++
// (this file), in Python (siftmetal_b200/api.py, used by the tests and the bench) and as an
// unverified Swift veneer (swift/). Names, argument meaning and error behaviour follow the
// reference:
//
//   SIFT::Configuration(IntegralSize)                         SIFT.swift:57-103
//   SIFT(device, configuration)                               SIFT.swift:112-143
//   getKeypoints(bgra8, bytesPerRow)  -> [[SIFTKeypoint]]     SIFT.swift:147-152 (7 octave lists)
//   getDescriptors(keypointOctaves)   -> [[SIFTDescriptor]]   SIFT.swift:207-238
//   match(source, target, abs, rel)   -> [SiftMatch]          SIFTDescriptor.swift:298-361
// plus the batch / pipelined calls and lazy views over the column wire format (siftcuda.h).
//
// Where the reference aborts (try!, precondition, fatalError), this mirror throws
// siftcuda::Error carrying the C status; there is no CPU fallback anywhere.
#pragma once

#include <array>
#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

#include "siftcuda.h"

namespace siftcuda {

struct Error : std::runtime_error {
    int status;
    Error(int s, const std::string& what) : std::runtime_error(what), status(s) {}
};

// Utilities/Math.swift:11-19
struct IntegralSize {
    int width;
    int height;
};

// SIFTKeypoint.swift:11-57 (SIMD2 members as std::array)
struct SIFTKeypoint {
    int octave;
    int scale;
    float subScale;
    std::array<int, 2> scaledCoordinate;
    std::array<float, 2> absoluteCoordinate;
    std::array<float, 2> normalizedCoordinate;
    float sigma;
    float value;
};

// Utilities/Vector.swift:12-60
struct IntVector {
    std::vector<int> components;
    int count() const { return (int)components.size(); }
    int operator[](int i) const { return components[(size_t)i]; }
    float distanceSquared(const IntVector& other) const {
        long k = 0;
        for (size_t i = 0; i < components.size(); i++) {
            const long d = other.components[i] - components[i];
            k += d * d;
        }
        return (float)k;
    }
};

// SIFTDescriptor.swift:12-35 (stored properties; the index keys of :43-89 are a permutation of
// the features for the trie matcher and do not change a distance)
struct SIFTDescriptor {
    SIFTKeypoint keypoint;
    float theta;
    IntVector features;
    std::vector<float> rawFeatures() const {  // SIFTDescriptor.swift:37-41
        std::vector<float> r(features.components.size());
        for (size_t i = 0; i < r.size(); i++) r[i] = (float)features.components[i] / 255.0f;
        return r;
    }
};

class SIFT {
public:
    // SIFT.swift:57-103 — only inputSize can be set from outside, as in the reference.
    struct Configuration {
        IntegralSize inputSize;
        explicit Configuration(IntegralSize size) : inputSize(size) {}
    };

    // `device` is a CUDA ordinal (was: MTLDevice).
    SIFT(int device, const Configuration& configuration) : configuration_(configuration) {
        SiftConfig cfg;
        check(sift_config_default(&cfg, configuration.inputSize.width, configuration.inputSize.height), nullptr);
        check(sift_create(&cfg, device, &ctx_), nullptr);
    }
    ~SIFT() { sift_destroy(ctx_); }
    SIFT(const SIFT&) = delete;
    SIFT& operator=(const SIFT&) = delete;

    const Configuration& configuration() const { return configuration_; }

    // `bgra8`: inputSize.height rows of `bytesPerRow` bytes of BGRA8 pixels (was: a bgra8Unorm
    // MTLTexture, ConvertSRGBToGrayscaleKernel.swift:34).
    std::vector<std::vector<SIFTKeypoint>> getKeypoints(const void* bgra8, int bytesPerRow) {
        const SiftKeypoint* kps = nullptr;
        int32_t counts[SIFT_NUM_OCTAVES];
        check(sift_detect(ctx_, bgra8, bytesPerRow, &kps, counts), ctx_);
        std::vector<std::vector<SIFTKeypoint>> out(SIFT_NUM_OCTAVES);
        size_t k = 0;
        for (int o = 0; o < SIFT_NUM_OCTAVES; o++) {
            out[o].reserve((size_t)counts[o]);
            for (int i = 0; i < counts[o]; i++, k++) out[o].push_back(fromPod(kps[k]));
        }
        return out;
    }

    // Depends on the pyramid left on the device by the preceding getKeypoints on this instance,
    // exactly like the reference (SIFTOctave.swift:354,459).
    std::vector<std::vector<SIFTDescriptor>> getDescriptors(
        const std::vector<std::vector<SIFTKeypoint>>& keypointOctaves) {
        if (keypointOctaves.size() != SIFT_NUM_OCTAVES)  // precondition, SIFT.swift:208
            throw Error(SIFT_ERR_INVALID_ARGUMENT, "keypointOctaves.count must equal the number of octaves");
        std::vector<SiftKeypoint> flat;
        std::vector<const SIFTKeypoint*> owner;
        int32_t counts[SIFT_NUM_OCTAVES];
        for (int o = 0; o < SIFT_NUM_OCTAVES; o++) {
            counts[o] = (int32_t)keypointOctaves[o].size();
            for (const auto& k : keypointOctaves[o]) {
                flat.push_back(toPod(k));
                owner.push_back(&k);
            }
        }
        const SiftDescriptor* desc = nullptr;
        int32_t dcounts[SIFT_NUM_OCTAVES];
        check(sift_describe(ctx_, flat.data(), counts, &desc, dcounts), ctx_);
        std::vector<std::vector<SIFTDescriptor>> out(SIFT_NUM_OCTAVES);
        size_t d = 0;
        for (int o = 0; o < SIFT_NUM_OCTAVES; o++) {
            out[o].reserve((size_t)dcounts[o]);
            for (int i = 0; i < dcounts[o]; i++, d++) {
                SIFTDescriptor r;
                r.keypoint = *owner[(size_t)desc[d].keypoint];
                r.theta = desc[d].theta;
                r.features.components.assign(desc[d].features, desc[d].features + SIFT_DESCRIPTOR_FEATURE_COUNT);
                out[o].push_back(std::move(r));
            }
        }
        return out;
    }

    SiftContext* context() { return ctx_; }

    // ---- beyond the reference's two calls: batches, the pipelined form, matching --------------
    // Lazy views over the column wire format (SiftBatchResult): nothing is converted until an
    // element is indexed — the reference builds one object per keypoint / descriptor up front
    // (SIFTOctave.swift:257-286, :470-489; SIFTDescriptor.swift:36-89). Views borrow the
    // context-owned pinned columns: valid until the slot is reused (see siftcuda.h).
    class KeypointView {
    public:
        KeypointView() = default;
        KeypointView(SiftContext* ctx, const SiftBatchResult& r, int64_t first, int64_t count)
            : ctx_(ctx), r_(r), first_(first), count_(count) {}
        int64_t size() const { return count_; }
        SIFTKeypoint operator[](int64_t i) const {
            SiftKeypoint p;
            check(sift_materialize_keypoints(ctx_, &r_, first_ + i, 1, &p), nullptr);
            return fromPod(p);
        }
        // raw columns of this view
        const float* absoluteX() const { return r_.keypoints.absolute_x + first_; }
        const float* absoluteY() const { return r_.keypoints.absolute_y + first_; }
        const float* sigma() const { return r_.keypoints.sigma + first_; }
    private:
        SiftContext* ctx_ = nullptr;
        SiftBatchResult r_{};
        int64_t first_ = 0, count_ = 0;
    };

    class DescriptorView {
    public:
        DescriptorView() = default;
        DescriptorView(const SiftBatchResult& r, int64_t first, int64_t count, KeypointView keypoints)
            : r_(r), first_(first), count_(count), keypoints_(keypoints) {}
        int64_t size() const { return count_; }
        SIFTDescriptor operator[](int64_t i) const {
            SiftDescriptor d;
            check(sift_materialize_descriptors(&r_, first_ + i, 1, &d), nullptr);
            SIFTDescriptor out;
            out.keypoint = keypoints_[d.keypoint];
            out.theta = d.theta;
            out.features.components.assign(d.features, d.features + SIFT_DESCRIPTOR_FEATURE_COUNT);
            return out;
        }
        // dense [size()][128] uint8 feature matrix: the operand layout of match()
        const uint8_t* features() const { return r_.descriptors.features + first_ * SIFT_DESCRIPTOR_FEATURE_COUNT; }
        const float* theta() const { return r_.descriptors.theta + first_; }
    private:
        SiftBatchResult r_{};
        int64_t first_ = 0, count_ = 0;
        KeypointView keypoints_;
    };

    struct FrameResult {
        KeypointView keypoints;
        DescriptorView descriptors;
        std::array<int32_t, SIFT_NUM_OCTAVES> keypointCounts{}, descriptorCounts{};
    };

    // getKeypoints + getDescriptors for a batch of frames in one call (no host round trip).
    std::vector<FrameResult> detectAndDescribe(const std::vector<const void*>& frames, int bytesPerRow) {
        SiftBatchResult r;
        check(sift_detect_and_describe_batch(ctx_, frames.data(), (int32_t)frames.size(), bytesPerRow, &r), ctx_);
        return split(r);
    }
    // Pipelined form: up to two calls in flight; wait() returns them in submission order.
    void submit(const std::vector<const void*>& frames, int bytesPerRow) {
        check(sift_submit(ctx_, frames.data(), (int32_t)frames.size(), bytesPerRow), ctx_);
    }
    std::vector<FrameResult> wait() {
        SiftBatchResult r;
        check(sift_wait(ctx_, &r), ctx_);
        return split(r);
    }

    // SIFTDescriptor.match(source:target:absoluteThreshold:relativeThreshold:)
    // (SIFTDescriptor.swift:298-361) on feature matrices; correspondences in source order.
    std::vector<SiftMatch> match(const uint8_t* source, int64_t nSource, const uint8_t* target, int64_t nTarget,
                                 float absoluteThreshold = 300.0f, float relativeThreshold = 0.6f) {
        const SiftMatch* m = nullptr;
        int64_t n = 0;
        check(sift_match(ctx_, source, nSource, target, nTarget, absoluteThreshold, relativeThreshold, &m, &n), ctx_);
        return std::vector<SiftMatch>(m, m + n);
    }
    std::vector<SiftMatch> match(const DescriptorView& source, const DescriptorView& target,
                                 float absoluteThreshold = 300.0f, float relativeThreshold = 0.6f) {
        return match(source.features(), source.size(), target.features(), target.size(), absoluteThreshold,
                     relativeThreshold);
    }

private:
    std::vector<FrameResult> split(const SiftBatchResult& r) {
        std::vector<FrameResult> out((size_t)r.n_frames);
        int64_t k = 0, d = 0;
        for (int f = 0; f < r.n_frames; f++) {
            int64_t nk = 0, nd = 0;
            for (int o = 0; o < SIFT_NUM_OCTAVES; o++) {
                out[(size_t)f].keypointCounts[(size_t)o] = r.keypoint_counts[f * SIFT_NUM_OCTAVES + o];
                out[(size_t)f].descriptorCounts[(size_t)o] = r.descriptor_counts[f * SIFT_NUM_OCTAVES + o];
                nk += r.keypoint_counts[f * SIFT_NUM_OCTAVES + o];
                nd += r.descriptor_counts[f * SIFT_NUM_OCTAVES + o];
            }
            out[(size_t)f].keypoints = KeypointView(ctx_, r, k, nk);
            out[(size_t)f].descriptors = DescriptorView(r, d, nd, out[(size_t)f].keypoints);
            k += nk;
            d += nd;
        }
        return out;
    }

    static SIFTKeypoint fromPod(const SiftKeypoint& p) {
        return SIFTKeypoint{p.octave, p.scale, p.subScale, {p.scaledX, p.scaledY},
                            {p.absoluteX, p.absoluteY}, {p.normalizedX, p.normalizedY}, p.sigma, p.value};
    }
    static SiftKeypoint toPod(const SIFTKeypoint& k) {
        SiftKeypoint p;
        p.octave = k.octave;
        p.scale = k.scale;
        p.subScale = k.subScale;
        p.scaledX = k.scaledCoordinate[0];
        p.scaledY = k.scaledCoordinate[1];
        p.absoluteX = k.absoluteCoordinate[0];
        p.absoluteY = k.absoluteCoordinate[1];
        p.normalizedX = k.normalizedCoordinate[0];
        p.normalizedY = k.normalizedCoordinate[1];
        p.sigma = k.sigma;
        p.value = k.value;
        return p;
    }
    static void check(int status, SiftContext* ctx) {
        if (status == SIFT_OK) return;
        std::string msg = sift_status_string(status);
        if (ctx) {
            const char* detail = sift_last_error_string(ctx);
            if (detail && *detail) msg += std::string(": ") + detail;
        }
        throw Error(status, msg);
    }

    Configuration configuration_;
    SiftContext* ctx_ = nullptr;
};

}  // namespace siftcuda
