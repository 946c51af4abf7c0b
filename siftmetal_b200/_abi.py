"""ctypes / numpy mirror of include/siftcuda.h (struct layouts only — no logic).

Each class restates one POD of the C ABI; the C header cites the reference struct it replaces
(Sources/MetalShaders/include/*.h, Sources/SIFTMetal/SIFT/SIFTKeypoint.swift:11-35).
"""
import ctypes as C

import numpy as np

NUM_OCTAVES = 7
SCALES_PER_OCTAVE = 3
NUM_GAUSSIANS = 6
NUM_DOGS = 5
ORIENTATION_BINS = 36
FEATURE_COUNT = 128
WEIGHTS_LENGTH = 32

SIFT_OK = 0
SIFT_ERR_INVALID_ARGUMENT = 1
SIFT_ERR_NO_DEVICE = 2
SIFT_ERR_CUDA = 3
SIFT_ERR_CAPACITY = 4
SIFT_ERR_NOT_DETECTED = 5
SIFT_ERR_OUT_OF_MEMORY = 6
SIFT_ERR_BUSY = 7

INPUT_BGRA8, INPUT_GRAY8, INPUT_NV12 = 0, 1, 2
INPUT_BYTES_PER_PIXEL = {INPUT_BGRA8: 4, INPUT_GRAY8: 1, INPUT_NV12: 1}

PLANE_GRAY, PLANE_SEED, PLANE_GAUSSIAN, PLANE_DOG, PLANE_GRADIENT = range(5)

STAGE_NAMES = ("seed", "pyramid", "extrema", "refine", "orientation", "descriptor")


class SiftConfig(C.Structure):
    _fields_ = [
        ("width", C.c_int32),
        ("height", C.c_int32),
        ("max_batch", C.c_int32),
        ("dog_threshold", C.c_float),
        ("edge_threshold", C.c_float),
        ("max_interpolation_iterations", C.c_int32),
        ("max_offset", C.c_float),
        ("image_border", C.c_int32),
        ("lambda_orientation", C.c_float),
        ("orientation_threshold", C.c_float),
        ("orientation_smoothing_iterations", C.c_int32),
        ("max_candidates_per_frame", C.c_int32),
        ("max_keypoints_per_frame", C.c_int32),
        ("max_descriptors_per_frame", C.c_int32),
        ("input_format", C.c_int32),
        ("reserved", C.c_int32),
    ]


class SiftInfo(C.Structure):
    _fields_ = [
        ("width", C.c_int32),
        ("height", C.c_int32),
        ("max_batch", C.c_int32),
        ("octave_width", C.c_int32 * NUM_OCTAVES),
        ("octave_height", C.c_int32 * NUM_OCTAVES),
        ("octave_pitch", C.c_int32 * NUM_OCTAVES),
        ("octave_delta", C.c_float * NUM_OCTAVES),
        ("sigmas", (C.c_float * NUM_GAUSSIANS) * NUM_OCTAVES),
        ("seed_sigma", C.c_float),
        ("seed_taps", C.c_int32),
        ("seed_weights", C.c_float * WEIGHTS_LENGTH),
        ("rho", C.c_float * (NUM_GAUSSIANS - 1)),
        ("taps", C.c_int32 * (NUM_GAUSSIANS - 1)),
        ("weights", (C.c_float * WEIGHTS_LENGTH) * (NUM_GAUSSIANS - 1)),
        ("max_candidates_per_frame", C.c_int32),
        ("max_keypoints_per_frame", C.c_int32),
        ("max_descriptors_per_frame", C.c_int32),
        ("device_bytes", C.c_int64),
        ("sm_count", C.c_int32),
    ]


class SiftTimings(C.Structure):
    _fields_ = [
        ("total_ms", C.c_float),
        ("stage_ms", C.c_float * 6),
        ("blur_octave0_ms", C.c_float),
        ("blur_octave0_launch_ms", C.c_float * (NUM_GAUSSIANS - 1)),
        ("blur_octave0_launches", C.c_int32),
        ("kernel_launches", C.c_int32),
        ("stage_timing_enabled", C.c_int32),
        ("graph_replay", C.c_int32),
    ]


class SiftKeypointColumns(C.Structure):
    _fields_ = [
        ("absolute_x", C.c_void_p),
        ("absolute_y", C.c_void_p),
        ("sigma", C.c_void_p),
        ("value", C.c_void_p),
        ("sub_scale", C.c_void_p),
        ("scaled_xy", C.c_void_p),
        ("octave_scale", C.c_void_p),
    ]


class SiftDescriptorColumns(C.Structure):
    _fields_ = [
        ("features", C.c_void_p),
        ("theta", C.c_void_p),
        ("keypoint", C.c_void_p),
    ]


class SiftBatchResult(C.Structure):
    _fields_ = [
        ("n_frames", C.c_int32),
        ("status", C.c_int32),
        ("keypoint_counts", C.POINTER(C.c_int32)),
        ("descriptor_counts", C.POINTER(C.c_int32)),
        ("candidate_counts", C.POINTER(C.c_int32)),
        ("total_keypoints", C.c_int64),
        ("total_descriptors", C.c_int64),
        ("slot", C.c_int32),
        ("reserved", C.c_int32),
        ("keypoints", SiftKeypointColumns),
        ("descriptors", SiftDescriptorColumns),
    ]


class SiftResultLayout(C.Structure):
    _fields_ = [("bytes", C.c_int64), ("capacity_keypoints", C.c_int64), ("capacity_descriptors", C.c_int64),
                ("offset", C.c_int64 * 10)]


class SiftMatch(C.Structure):
    _fields_ = [("source", C.c_int32), ("target", C.c_int32), ("distance", C.c_float)]


MATCH_DTYPE = np.dtype([("source", "<i4"), ("target", "<i4"), ("distance", "<f4")])

# numpy views of the result PODs
KEYPOINT_DTYPE = np.dtype(
    [
        ("octave", "<i4"),
        ("scale", "<i4"),
        ("subScale", "<f4"),
        ("scaledX", "<i4"),
        ("scaledY", "<i4"),
        ("absoluteX", "<f4"),
        ("absoluteY", "<f4"),
        ("normalizedX", "<f4"),
        ("normalizedY", "<f4"),
        ("sigma", "<f4"),
        ("value", "<f4"),
    ]
)
DESCRIPTOR_DTYPE = np.dtype([("keypoint", "<i4"), ("theta", "<f4"), ("features", "u1", (FEATURE_COUNT,))])
assert KEYPOINT_DTYPE.itemsize == 44
assert DESCRIPTOR_DTYPE.itemsize == 136
