/*
 * siftcuda.h — C ABI of libsiftcuda.so, the B200 (sm_100a) SIFT detect + describe engine.
 *
 * This header is the drop-in seam for the hot path of lukevanin/SIFTMetal. In the reference the
 * seam is the C module `MetalShaders` (Sources/MetalShaders/module.modulemap:1-4, umbrella
 * Sources/MetalShaders/include/MetalShaders.h:6-11) whose POD structs cross between the Swift
 * host (Sources/SIFTMetal/SIFT/{SIFT,SIFTOctave,DifferenceOfGaussians}.swift) and the Metal
 * kernels (Sources/MetalShaders/Metal/ *.metal). Here the whole per-frame pipeline lives behind
 * the entry points below, so the Swift (or C++/Python) host only marshals pixels in and
 * keypoints / descriptors out. Every declaration cites the reference interface it replaces.
 *
 * Plain C, plain pointers and sizes. No torch types. No CPU fallback: every compute entry point
 * returns SIFT_ERR_NO_DEVICE / SIFT_ERR_CUDA when no sm_100 device is usable.
 */
#ifndef SIFTCUDA_H
#define SIFTCUDA_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SIFTCUDA_ABI_VERSION 2

/* DifferenceOfGaussians.Configuration (DifferenceOfGaussians.swift:23-51): 7 octaves, 3 scales
 * per octave, hence 6 Gaussian + 5 DoG slices per octave (:80-81). Fixed, as in the reference. */
#define SIFT_NUM_OCTAVES 7
#define SIFT_SCALES_PER_OCTAVE 3
#define SIFT_NUM_GAUSSIANS 6
#define SIFT_NUM_DOGS 5
/* include/SIFTOrientation.h:12, include/SIFTDescriptor.h:14-16 */
#define SIFT_ORIENTATION_HISTOGRAM_BINS 36
#define SIFT_DESCRIPTOR_HISTOGRAM_WIDTH 4
#define SIFT_DESCRIPTOR_ORIENTATION_BINS 8
#define SIFT_DESCRIPTOR_FEATURE_COUNT 128
/* include/ConvolutionSeries.h:13 */
#define SIFT_CONVOLUTION_WEIGHTS_LENGTH 32

/* ---- status codes (the reference aborts via try!/precondition/fatalError; we return) ------- */
enum {
    SIFT_OK = 0,
    SIFT_ERR_INVALID_ARGUMENT = 1, /* null pointer, bad size, batch larger than max_batch        */
    SIFT_ERR_NO_DEVICE = 2,        /* no CUDA device / not an sm_100 part                          */
    SIFT_ERR_CUDA = 3,             /* a CUDA runtime call failed (see sift_last_error_string)      */
    SIFT_ERR_CAPACITY = 4,         /* a device list overflowed its capacity; results truncated —  */
                                   /* replaces the precondition crash of Buffer.swift:35-39        */
    SIFT_ERR_NOT_DETECTED = 5,     /* sift_describe before sift_detect (SIFTOctave.swift:354,459)  */
    SIFT_ERR_OUT_OF_MEMORY = 6,
    SIFT_ERR_BUSY = 7,             /* sift_submit with every in-flight slot taken / wait on none   */
};

/* ---- input pixel formats ----------------------------------------------------------------------
 * The reference accepts one format: a bgra8Unorm texture (precondition of
 * ConvertSRGBToGrayscaleKernel.swift:34), which CoreVideoMetalCache.swift:23-31 makes from a camera
 * CVPixelBuffer. GRAY8 and NV12 are the ingestion formats of the video-tracking configuration
 * (SURVEY.md §8f-2): a gray byte v is converted exactly as the BGRA pixel (v, v, v, 255) would be
 * (same luminance expression, ConvertSRGBToGrayscale.metal:11-23), so results are identical to the
 * reference's on the gray-expanded frame while a quarter of the bytes cross PCIe / HBM. NV12: the
 * pointer passed is the Y (luma) plane, `pitch_bytes` its row pitch; the interleaved CbCr plane is
 * never read. */
#define SIFT_INPUT_BGRA8 0
#define SIFT_INPUT_GRAY8 1
#define SIFT_INPUT_NV12 2

/* ---- configuration ---------------------------------------------------------------------------
 * SIFT.Configuration (SIFT.swift:57-103). In the reference only `inputSize` can be set and the
 * thresholds are literals at the call sites (SIFTOctave.swift:217-226, :296-300, :396-401);
 * sift_config_default() fills exactly those literals. */
typedef struct SiftConfig {
    int32_t width;                  /* inputSize.width  (pixels of the input)                     */
    int32_t height;                 /* inputSize.height                                           */
    int32_t max_batch;              /* frames resident per call; 1 = the reference's behaviour    */
    float   dog_threshold;          /* 0.0133  (SIFTOctave.swift:218)                             */
    float   edge_threshold;         /* 10.0    (SIFTOctave.swift:224)                             */
    int32_t max_interpolation_iterations; /* 5 (SIFTOctave.swift:219)                             */
    float   max_offset;             /* 0.6     (SIFTOctave.swift:220)                             */
    int32_t image_border;           /* 5       (SIFTInterpolate.metal:182); must be >= 1          */
    float   lambda_orientation;     /* 1.5     (SIFTOctave.swift:298)                             */
    float   orientation_threshold;  /* 0.8     (SIFTOctave.swift:299)                             */
    int32_t orientation_smoothing_iterations; /* 6 (SIFTOrientation.metal:167)                    */
    /* Per-frame list capacities. The reference hard-codes 4096/4096/2048 per octave and aborts
     * beyond (SIFTOctave.swift:22-26); 0 = size from the image.                   */
    int32_t max_candidates_per_frame;
    int32_t max_keypoints_per_frame;
    int32_t max_descriptors_per_frame;
    int32_t input_format;           /* SIFT_INPUT_*; 0 = BGRA8, the reference's only format       */
    int32_t reserved;               /* must be 0                                                  */
} SiftConfig;

/* ---- record types of the reference (array-of-structs) ----------------------------------------- */

/* SIFTKeypoint (SIFTKeypoint.swift:11-35), field for field; SIMD2 members flattened. It is what
 * SIFTOctave.interpolateKeypoints builds from SIFTInterpolateOutputKeypoint
 * (include/SIFTInterpolate.h:34-46, SIFTOctave.swift:257-286). 44 bytes. */
typedef struct SiftKeypoint {
    int32_t octave;
    int32_t scale;
    float   subScale;
    int32_t scaledX, scaledY;         /* scaledCoordinate                                         */
    float   absoluteX, absoluteY;     /* absoluteCoordinate (input-image pixels)                  */
    float   normalizedX, normalizedY; /* normalizedCoordinate                                     */
    float   sigma;
    float   value;
} SiftKeypoint;

/* SIFTDescriptorResult (include/SIFTDescriptor.h:37-42) with the 128 features packed to bytes
 * (they are 0…255 by construction, SIFTDescriptor.metal:41-49) — 136 bytes instead of 524.
 * `keypoint` indexes the frame's keypoint array (octave-major, as returned by the same call). */
typedef struct SiftDescriptor {
    int32_t keypoint;
    float   theta;
    uint8_t features[SIFT_DESCRIPTOR_FEATURE_COUNT];
} SiftDescriptor;

/* ---- result wire format: packed columns (structure-of-arrays) -----------------------------------
 * What a batch call returns. The reference builds one Swift object per keypoint / descriptor on
 * the CPU (SIFTOctave.swift:257-286, :470-489; SIFTDescriptor.init does float copies and a
 * re-ordering per descriptor, SIFTDescriptor.swift:36-89) — at 10^5 descriptors per frame that
 * would dominate. Here the kernels write the columns below directly (into pinned host memory for
 * the host-buffer calls) and hosts materialise SiftKeypoint / SiftDescriptor records lazily
 * (sift_materialize_*; the Python / C++ / Swift mirrors index the columns on demand).
 * 26 bytes per keypoint, 136 per descriptor; the feature matrix is dense [n][128] uint8, which is
 * also the operand layout of sift_match. */
typedef struct SiftKeypointColumns {
    const float*   absolute_x;     /* absoluteCoordinate.x, input-image pixels                    */
    const float*   absolute_y;
    const float*   sigma;
    const float*   value;
    const float*   sub_scale;
    const int16_t* scaled_xy;      /* scaledCoordinate: x, y interleaved (octave planes <= 16384) */
    const uint8_t* octave_scale;   /* octave, scale interleaved                                   */
} SiftKeypointColumns;

typedef struct SiftDescriptorColumns {
    const uint8_t* features;       /* [n][128]                                                    */
    const float*   theta;
    const int32_t* keypoint;       /* index into the frame's keypoints                            */
} SiftDescriptorColumns;

/* Result of a batch call. All arrays are context-owned pinned host memory, valid until the slot
 * is reused: the next call on the context for the synchronous entry points, the second following
 * sift_submit for the pipelined ones. Keypoints / descriptors are concatenated frame-major,
 * octave-major, in canonical order (scale, y, x of the originating extremum; orientation bin
 * ascending). normalizedCoordinate = scaledCoordinate / octave size is not stored (derived). */
typedef struct SiftBatchResult {
    int32_t n_frames;
    int32_t status;                     /* SIFT_OK or SIFT_ERR_CAPACITY (results truncated)       */
    const int32_t* keypoint_counts;     /* [n_frames][SIFT_NUM_OCTAVES]                           */
    const int32_t* descriptor_counts;   /* [n_frames][SIFT_NUM_OCTAVES]                           */
    const int32_t* candidate_counts;    /* [n_frames][SIFT_NUM_OCTAVES] raw 25-neighbour extrema  */
                                        /* above the 0.8·C_DoG pre-threshold                      */
    int64_t total_keypoints;
    int64_t total_descriptors;
    int32_t slot;                       /* in-flight slot (0 / 1) that holds these columns         */
    int32_t reserved;
    SiftKeypointColumns keypoints;
    SiftDescriptorColumns descriptors;
} SiftBatchResult;

/* Geometry and schedule of one context: DifferenceOfGaussians.init (:233-344) restated. */
typedef struct SiftInfo {
    int32_t width, height, max_batch;
    int32_t octave_width[SIFT_NUM_OCTAVES];
    int32_t octave_height[SIFT_NUM_OCTAVES];
    int32_t octave_pitch[SIFT_NUM_OCTAVES];          /* floats per row in device memory           */
    float   octave_delta[SIFT_NUM_OCTAVES];
    float   sigmas[SIFT_NUM_OCTAVES][SIFT_NUM_GAUSSIANS];
    float   seed_sigma;
    int32_t seed_taps;
    float   seed_weights[SIFT_CONVOLUTION_WEIGHTS_LENGTH];
    float   rho[SIFT_NUM_GAUSSIANS - 1];
    int32_t taps[SIFT_NUM_GAUSSIANS - 1];
    float   weights[SIFT_NUM_GAUSSIANS - 1][SIFT_CONVOLUTION_WEIGHTS_LENGTH];
    int32_t max_candidates_per_frame, max_keypoints_per_frame, max_descriptors_per_frame;
    int64_t device_bytes;                            /* HBM held by the context                   */
    int32_t sm_count;
} SiftInfo;

/* Per-stage device times of the last completed call, from CUDA events recorded on the context's
 * own stream (replaces measure(name:) of Utilities/Performance.swift:12-20; the same sites also
 * carry NVTX ranges named as the reference's: findKeypoints, getKeypointsFromOctaves,
 * interpolateKeypoints, getDescriptors(orientations), getDescriptors(descriptors)). ms.
 * total_ms is always measured (two events around the call's work). The per-stage split is opt-in
 * (sift_set_stage_timing): it records ~20 more events and is not available under graph replay. */
#define SIFT_STAGE_SEED 0
#define SIFT_STAGE_PYRAMID 1     /* blur + DoG + gradient + extrema mask, octaves on forked streams */
#define SIFT_STAGE_EXTREMA 2     /* mask -> ordered candidate list (scan + scatter)                */
#define SIFT_STAGE_REFINE 3
#define SIFT_STAGE_ORIENTATION 4
#define SIFT_STAGE_DESCRIPTOR 5
#define SIFT_STAGE_COUNT 6
typedef struct SiftTimings {
    float total_ms;                      /* first launch → last kernel                           */
    float stage_ms[SIFT_STAGE_COUNT];    /* valid when stage timing was enabled                   */
    float blur_octave0_ms;               /* octave 0: the 5 blur + DoG scales, summed             */
    float blur_octave0_launch_ms[SIFT_NUM_GAUSSIANS - 1]; /* each scale (11,15,17,21,27 taps)     */
    int32_t blur_octave0_launches;
    int32_t kernel_launches;             /* kernels launched (or replayed from the graph)         */
    int32_t stage_timing_enabled;
    int32_t graph_replay;                /* 1: the call ran as one CUDA graph launch              */
} SiftTimings;

typedef struct SiftContext SiftContext;

/* ---- lifetime ----------------------------------------------------------------------------- */

/* Fills the literals the reference uses. Replaces SIFT.Configuration.init(inputSize:)
 * (SIFT.swift:100-102). */
int sift_config_default(SiftConfig* config, int32_t width, int32_t height);

/* Replaces SIFT.init(device:configuration:) (SIFT.swift:112-143): allocates every device plane
 * and list once, computes the Gaussian tap tables on the host in float
 * (GaussianKernel.swift:20-43, GaussianSeriesKernel.swift:27-51). `device` is a CUDA ordinal
 * (was: MTLDevice). SIFT_ERR_INVALID_ARGUMENT for sizes outside [8, 16384], image_border < 1,
 * non-finite thresholds, negative capacities / iteration counts, an unknown input format. */
int sift_create(const SiftConfig* config, int device, SiftContext** out_context);
void sift_destroy(SiftContext* context);
int sift_get_info(const SiftContext* context, SiftInfo* out_info);

/* ---- the reference's two entry points ------------------------------------------------------- */

/* Replaces SIFT.getKeypoints(_:) (SIFT.swift:147-152). `pixels` is a host pointer to height rows
 * of `pitch_bytes` bytes in the context's input format (was: a bgra8Unorm MTLTexture,
 * ConvertSRGBToGrayscaleKernel.swift:34). On return *out_keypoints points at
 * sum(counts_per_octave) keypoint records grouped by octave (context-owned, valid until the next
 * call); the Gaussian-gradient planes stay on the device for a following sift_describe, exactly
 * like the reference's gradientTextures (SIFTOctave.swift:354,459). */
int sift_detect(SiftContext* context, const void* pixels, int32_t pitch_bytes,
                const SiftKeypoint** out_keypoints, int32_t counts_per_octave[SIFT_NUM_OCTAVES]);

/* Replaces SIFT.getDescriptors(keypointOctaves:) (SIFT.swift:207-238): orientation assignment
 * (SIFTOctave.swift:290-382) then one descriptor per (keypoint, orientation)
 * (SIFTOctave.swift:384-492). `keypoints` are caller-supplied (they may have been filtered; they
 * may alias or overlap the array sift_detect returned), grouped by octave with
 * `counts_per_octave`; SiftDescriptor.keypoint indexes that array. */
int sift_describe(SiftContext* context, const SiftKeypoint* keypoints,
                  const int32_t counts_per_octave[SIFT_NUM_OCTAVES],
                  const SiftDescriptor** out_descriptors,
                  int32_t descriptor_counts_per_octave[SIFT_NUM_OCTAVES]);

/* ---- batch path (frames are independent units; they shard across GPUs by context) --------- */

/* getKeypoints + getDescriptors for n ≤ max_batch frames in one go, no host round trip between
 * the two. images[i] is a host frame in the context's input format. Synchronous:
 * = sift_submit + sift_wait. */
int sift_detect_and_describe_batch(SiftContext* context, const void* const* images, int32_t n,
                                   int32_t pitch_bytes, SiftBatchResult* out_result);

/* Pipelined form of the same call (video ingestion; the reference's seam is
 * CoreVideoMetalCache.swift:23-31, one texture per camera frame). A context has two in-flight
 * slots: sift_submit queues the upload of the frames (copy stream, into the slot's own input
 * arena) and all the kernels behind it and returns at once; sift_wait blocks until the OLDEST
 * submitted call has finished and hands out its result. With two submits in flight the upload of
 * call i+1 crosses PCIe under the kernels of call i and the results of call i (written into the
 * slot's pinned arrays by the kernels themselves) are complete when its last kernel retires.
 * The host frames must stay valid and unmodified until the matching sift_wait returns.
 * A small context (<= 6 GB of device memory and max_batch <= 8, e.g. single 1080p / 4K frames) gives slot 1 a second
 * pipeline of its own (scratch planes, streams), created by the first overlapping submit: the
 * kernels of two calls in flight then run beside each other on the device, and sift_get_info()
 * reports the doubled device_bytes. After a call that ran on slot 1 the debug taps and
 * sift_batch_download of the context have nothing current (SIFT_ERR_NOT_DETECTED);
 * sift_match_frames follows the slot the last completed call ran on.
 * SIFT_ERR_BUSY: submit with both slots in flight, or wait with none. */
int sift_submit(SiftContext* context, const void* const* images, int32_t n, int32_t pitch_bytes);
int sift_wait(SiftContext* context, SiftBatchResult* out_result);
int sift_pending(const SiftContext* context);   /* number of submitted, not yet waited calls   */
int sift_next_slot(const SiftContext* context); /* slot the next sift_submit will use (0 / 1)  */

/* Where a slot's result columns live. By default each slot owns a pinned block; a caller that
 * gathers results across processes (one process per GPU) can bind a slot to its own
 * memory instead — e.g. a POSIX shared-memory mapping another process reads — so that the kernels'
 * stores ARE the gather: nothing is copied on the host. sift_register_host_memory pins a range
 * once (cudaHostRegister; unregistered by sift_destroy; it must outlive the context);
 * sift_bind_result_memory points a slot with no call in flight at a 256-byte aligned block of at
 * least sift_result_layout().bytes inside a registered range — cheap, so a caller may rotate a
 * slot through several blocks to keep older results alive. offset[] locates the ten columns in a
 * block, in the order of SiftKeypointColumns then SiftDescriptorColumns. */
typedef struct SiftResultLayout {
    int64_t bytes;
    int64_t capacity_keypoints, capacity_descriptors;   /* rows each column can hold (whole batch) */
    int64_t offset[10];
} SiftResultLayout;
int sift_result_layout(const SiftContext* context, SiftResultLayout* out_layout);
int sift_register_host_memory(SiftContext* context, void* base, int64_t bytes);
int sift_bind_result_memory(SiftContext* context, int32_t slot, void* base, int64_t bytes);

/* The same work split so that device-resident inputs can be timed apart from PCIe:
 *   upload (H2D into the context's input arena; the copy has completed when the call returns, so
 *   the host frames may be reused at once)  or  set_device_input (caller's device memory,
 *   n frames `frame_stride_bytes` apart — not copied, must stay valid until execute returns),
 *   execute (every kernel of the path; returns after the stream drained; SIFT_ERR_CAPACITY if a
 *   list overflowed), download (D2H of the result columns). */
int sift_batch_upload(SiftContext* context, const void* const* images, int32_t n,
                      int32_t pitch_bytes);
int sift_batch_set_device_input(SiftContext* context, const void* device_pixels, int32_t n,
                                int32_t pitch_bytes, int64_t frame_stride_bytes);
int sift_batch_execute(SiftContext* context);
int sift_batch_download(SiftContext* context, SiftBatchResult* out_result);

/* Lazy views: records [first, first + count) of a result, built from its columns on the host.
 * SIFTKeypoint.swift:37-46 / SIFTDescriptor.swift:26-34 initialisers, one record at a time. */
int sift_materialize_keypoints(const SiftContext* context, const SiftBatchResult* result,
                               int64_t first, int64_t count, SiftKeypoint* dst);
int sift_materialize_descriptors(const SiftBatchResult* result, int64_t first, int64_t count,
                                 SiftDescriptor* dst);

/* ---- descriptor matching ------------------------------------------------------
 * SIFTDescriptor.match(source:target:absoluteThreshold:relativeThreshold:)
 * (SIFTDescriptor.swift:298-361): for every source descriptor a linear scan of the targets in
 * order keeping `best` (strict <) and `second` = the best before the last improvement (NOT the
 * true second smallest, :339-343); kept iff best < absoluteThreshold and best < second ·
 * relativeThreshold. distance = sqrt(distanceSquared) over features / 255
 * (SIFTDescriptor.swift:37-41, Utilities/Vector.swift:226-235). On the device the n_source ×
 * n_target squared distances are one uint8 GEMM on tcgen05 (‖a‖² + ‖b‖² − 2 a·b, exact in
 * int32) with the scan fused into its epilogue. Matches come back in source order. */
typedef struct SiftMatch {
    int32_t source;          /* row of the source feature matrix                                  */
    int32_t target;          /* row of the target feature matrix                                  */
    float   distance;        /* featureDistance (SIFTCorrespondence.swift)                        */
} SiftMatch;

/* Host feature matrices [n][128] uint8 (SiftDescriptorColumns.features of any result). */
int sift_match(SiftContext* context, const uint8_t* source_features, int64_t n_source,
               const uint8_t* target_features, int64_t n_target, float absolute_threshold,
               float relative_threshold, const SiftMatch** out_matches, int64_t* out_count);
/* Frames of the last completed batch of this context, matched from their device-resident
 * feature columns (no descriptor bytes cross PCIe). */
int sift_match_frames(SiftContext* context, int32_t source_frame, int32_t target_frame,
                      float absolute_threshold, float relative_threshold,
                      const SiftMatch** out_matches, int64_t* out_count);

/* SIFTDescriptor.matchGeometry(source:target:absoluteThreshold:relativeThreshold:)
 * (SIFTDescriptor.swift:104-296; defaults 1.176 / 0.6): brute-force matches on the device, then —
 * on the host, as in the reference — the geometric-consistency score of the first 80 of them
 * (0 when fewer than 7 matched). `*_xy` are the absoluteCoordinate (x, y) pairs of the keypoint
 * each descriptor row belongs to, interleaved. */
int sift_match_geometry(SiftContext* context, const uint8_t* source_features, const float* source_xy,
                        int64_t n_source, const uint8_t* target_features, const float* target_xy,
                        int64_t n_target, float absolute_threshold, float relative_threshold,
                        float* out_score);

/* SIFTDescriptor.approximateMatch(source:target:absoluteThreshold:relativeThreshold:)
 * (SIFTDescriptor.swift:362-417) over the reference's approximate-nearest-neighbour trie
 * (Utilities/Trie.swift: 8 bins, keys = the 16 per-cell feature means, radius 10, k 2). A host
 * stage in the reference and here (it visits ~21 leaves per query); kept as a flat sorted-key
 * table instead of a pointer trie. Same leaf order, same visiting order, same results. */
int sift_approximate_match(SiftContext* context, const uint8_t* source_features, int64_t n_source,
                           const uint8_t* target_features, int64_t n_target,
                           float absolute_threshold, float relative_threshold,
                           const SiftMatch** out_matches, int64_t* out_count);

/* ---- diagnostics -------------------------------------------------------------------------- */

const char* sift_status_string(int status);
const char* sift_last_error_string(const SiftContext* context);
int sift_set_stage_timing(SiftContext* context, int32_t enabled);   /* default: off            */
/* Replay each call as one CUDA graph launch (recorded at the second call of a given slot / batch
 * size / input, as the reference encodes its 102 dispatches into one command buffer,
 * SIFT.swift:157-172) instead of ~70 stream launches. Cuts the host cost of a call by ~10x; on
 * the device the replay measured ~4 % slower than the launches with programmatic dependent launch
 * on prioritised streams, so it is off unless enabled here or by SIFTCUDA_GRAPH=1 — meant for
 * hosts that feed many GPUs. Results are identical either way. */
int sift_set_graph_replay(SiftContext* context, int32_t enabled);
int sift_last_timings(const SiftContext* context, SiftTimings* out_timings);

/* Debug taps into the pyramid of the last execute. `what`: */
#define SIFT_PLANE_GRAY 0        /* luminosity, W×H            (slice, octave ignored)             */
#define SIFT_PLANE_SEED 1        /* blurred 2× seed = octave 0 Gaussian slice 0                    */
#define SIFT_PLANE_GAUSSIAN 2    /* slice 0…5                                                      */
#define SIFT_PLANE_DOG 3         /* slice 0…4                                                      */
#define SIFT_PLANE_GRADIENT 4    /* slice 1…3, interleaved (orientation, magnitude), 2 floats/px   */
int sift_debug_download(SiftContext* context, int32_t what, int32_t frame, int32_t octave,
                        int32_t slice, float* dst, int64_t dst_floats);

/* Raw candidates (x, y, scale triples = SIFTExtremaResult, include/SIFTExtrema.h:14-18) of one
 * frame and octave of the last execute, canonical order. Returns the count (or <0: -status). */
int64_t sift_debug_candidates(SiftContext* context, int32_t frame, int32_t octave,
                              int32_t* dst_xyz, int64_t dst_capacity_triples);

/* Evaluates the engine's device math primitives on n inputs (parity tests against the oracle's
 * restatement): op 0 expf(a) · 1 atan2f(a, b) · 2 sinf(a) · 3 cosf(a) · 4 exp2f(a). */
int sift_debug_math(int device, int32_t op, const float* a, const float* b, float* out,
                    int64_t n);

/* Tuning aid: times `iters` back-to-back launches of the octave-0 blur of scale `scale` (0..4 =
 * 11,15,17,21,27 taps) on the planes left by the last execute, with CUDA events; `mode` 0 normal,
 * 1 without the FMA loops, 2 without the stores, 3 neither. Mean ms per launch in *out_ms. */
int sift_debug_blur_bench(SiftContext* context, int32_t scale, int32_t mode, int32_t iters,
                          float* out_ms);

#ifdef __cplusplus
}
#endif
#endif /* SIFTCUDA_H */
