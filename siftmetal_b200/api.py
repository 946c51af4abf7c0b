"""Host-side mirror of the reference's public SIFT API over the C ABI (include/siftcuda.h).

The reference's host is Swift (Sources/SIFTMetal/SIFT/SIFT.swift); this image has no Swift
toolchain, so the same surface is offered here for the tests and the bench (and in C++ in
include/SIFT.hpp, in Swift — unverified — under swift/). Names, argument meaning and error
behaviour follow the reference:

    SIFT.Configuration(inputSize: IntegralSize)         SIFT.swift:57-103
    SIFT(device:configuration:)                          SIFT.swift:112-143
    getKeypoints(_:) -> [[SIFTKeypoint]]                 SIFT.swift:147-152  (7 octave lists)
    getDescriptors(keypointOctaves:) -> [[SIFTDescriptor]]   SIFT.swift:207-238

Everything numeric happens in libsiftcuda.so on the GPU. There is no CPU fallback: if the
library is missing or no sm_100 device is usable, construction raises.
"""
import ctypes as C
import os
from dataclasses import dataclass
from typing import List, Optional, Sequence

import numpy as np

from . import _abi
from ._abi import (  # noqa: F401  (re-exported)
    DESCRIPTOR_DTYPE,
    KEYPOINT_DTYPE,
    NUM_OCTAVES,
    SiftBatchResult,
    SiftConfig,
    SiftInfo,
    SiftTimings,
)

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libsiftcuda.so")

_lib = None


class SiftError(RuntimeError):
    def __init__(self, status, message):
        super().__init__(f"siftcuda status {status}: {message}")
        self.status = status


def load_library():
    """dlopen libsiftcuda.so and declare its prototypes. Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found — build it with `python -m siftmetal_b200.build` "
            "(there is no CPU fallback)"
        )
    L = C.CDLL(LIB_PATH)
    vp, i32, i64 = C.c_void_p, C.c_int32, C.c_int64
    L.sift_config_default.argtypes = [C.POINTER(SiftConfig), i32, i32]
    L.sift_create.argtypes = [C.POINTER(SiftConfig), C.c_int, C.POINTER(vp)]
    L.sift_destroy.argtypes = [vp]
    L.sift_destroy.restype = None
    L.sift_get_info.argtypes = [vp, C.POINTER(SiftInfo)]
    L.sift_detect.argtypes = [vp, vp, i32, C.POINTER(vp), C.POINTER(i32)]
    L.sift_describe.argtypes = [vp, vp, C.POINTER(i32), C.POINTER(vp), C.POINTER(i32)]
    L.sift_detect_and_describe_batch.argtypes = [vp, C.POINTER(vp), i32, i32, C.POINTER(SiftBatchResult)]
    L.sift_batch_upload.argtypes = [vp, C.POINTER(vp), i32, i32]
    L.sift_batch_set_device_input.argtypes = [vp, vp, i32, i32, i64]
    L.sift_batch_execute.argtypes = [vp]
    L.sift_batch_download.argtypes = [vp, C.POINTER(SiftBatchResult)]
    L.sift_status_string.argtypes = [C.c_int]
    L.sift_status_string.restype = C.c_char_p
    L.sift_last_error_string.argtypes = [vp]
    L.sift_last_error_string.restype = C.c_char_p
    L.sift_set_stage_timing.argtypes = [vp, i32]
    L.sift_last_timings.argtypes = [vp, C.POINTER(SiftTimings)]
    L.sift_debug_download.argtypes = [vp, i32, i32, i32, i32, vp, i64]
    L.sift_debug_candidates.argtypes = [vp, i32, i32, vp, i64]
    L.sift_debug_candidates.restype = i64
    L.sift_debug_math.argtypes = [C.c_int, i32, vp, vp, vp, i64]
    L.sift_debug_blur_bench.argtypes = [vp, i32, i32, i32, C.POINTER(C.c_float)]
    _lib = L
    return L


EXPORTED_SYMBOLS = (
    "sift_config_default sift_create sift_destroy sift_get_info sift_detect sift_describe "
    "sift_detect_and_describe_batch sift_batch_upload sift_batch_set_device_input "
    "sift_batch_execute sift_batch_download sift_status_string sift_last_error_string "
    "sift_set_stage_timing sift_last_timings sift_debug_download sift_debug_candidates "
    "sift_debug_math sift_debug_blur_bench"
).split()


@dataclass
class BatchResult:
    """Flat result of one batch: structured arrays + per-(frame, octave) counts."""

    keypoints: np.ndarray          # KEYPOINT_DTYPE, frame-major, octave-major
    descriptors: np.ndarray        # DESCRIPTOR_DTYPE
    keypoint_counts: np.ndarray    # [n_frames, 7]
    descriptor_counts: np.ndarray  # [n_frames, 7]
    candidate_counts: np.ndarray   # [n_frames, 7]
    status: int = 0

    def frame(self, f):
        k0 = int(self.keypoint_counts[:f].sum())
        k1 = k0 + int(self.keypoint_counts[f].sum())
        d0 = int(self.descriptor_counts[:f].sum())
        d1 = d0 + int(self.descriptor_counts[f].sum())
        return self.keypoints[k0:k1], self.descriptors[d0:d1]


class Engine:
    """Thin owner of one SiftContext (one GPU, one stream). Not thread-safe, like the
    reference's SIFT instance (one MTLCommandQueue, shared scratch buffers)."""

    def __init__(self, width, height, device=0, max_batch=1, **overrides):
        self.L = load_library()
        self.cfg = SiftConfig()
        self._check(self.L.sift_config_default(C.byref(self.cfg), width, height), ctx=False)
        self.cfg.max_batch = max_batch
        for k, v in overrides.items():
            if not hasattr(self.cfg, k):
                raise TypeError(f"unknown SiftConfig field {k}")
            setattr(self.cfg, k, v)
        self.ctx = C.c_void_p()
        st = self.L.sift_create(C.byref(self.cfg), device, C.byref(self.ctx))
        if st != 0:
            self.ctx = None
            raise SiftError(st, self.L.sift_status_string(st).decode())
        self.device = device
        self.width, self.height, self.max_batch = width, height, max_batch
        self.info = SiftInfo()
        self.L.sift_get_info(self.ctx, C.byref(self.info))

    def close(self):
        if getattr(self, "ctx", None):
            self.L.sift_destroy(self.ctx)
            self.ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _check(self, st, ctx=True, allow_capacity=False):
        if st == 0 or (allow_capacity and st == _abi.SIFT_ERR_CAPACITY):
            return st
        msg = self.L.sift_status_string(st).decode()
        if ctx and self.ctx:
            detail = self.L.sift_last_error_string(self.ctx).decode()
            if detail:
                msg = f"{msg}: {detail}"
        raise SiftError(st, msg)

    # -- reference-shaped single-frame calls -------------------------------------------------
    def detect(self, bgra):
        bgra = self._as_bgra(bgra)
        out = C.c_void_p()
        counts = (C.c_int32 * NUM_OCTAVES)()
        self._check(self.L.sift_detect(self.ctx, bgra.ctypes.data, bgra.strides[0], C.byref(out), counts))
        counts = np.array(counts, dtype=np.int32)
        n = int(counts.sum())
        kps = self._view(out.value, n, KEYPOINT_DTYPE).copy()
        return kps, counts

    def describe(self, keypoints, counts):
        keypoints = np.ascontiguousarray(keypoints, dtype=KEYPOINT_DTYPE)
        cin = (C.c_int32 * NUM_OCTAVES)(*[int(c) for c in counts])
        out = C.c_void_p()
        cout = (C.c_int32 * NUM_OCTAVES)()
        self._check(self.L.sift_describe(self.ctx, keypoints.ctypes.data, cin, C.byref(out), cout))
        cout = np.array(cout, dtype=np.int32)
        return self._view(out.value, int(cout.sum()), DESCRIPTOR_DTYPE).copy(), cout

    # -- batch path ----------------------------------------------------------------------------
    def upload(self, frames: Sequence[np.ndarray]):
        frames = [self._as_bgra(f) for f in frames]
        ptrs = (C.c_void_p * len(frames))(*[f.ctypes.data for f in frames])
        self._keepalive = frames
        self._check(self.L.sift_batch_upload(self.ctx, ptrs, len(frames), frames[0].strides[0]))

    def set_device_input(self, device_ptr, n, pitch_bytes, frame_stride_bytes):
        self._check(self.L.sift_batch_set_device_input(self.ctx, device_ptr, n, pitch_bytes, frame_stride_bytes))

    def execute(self, allow_capacity=False):
        return self._check(self.L.sift_batch_execute(self.ctx), allow_capacity=allow_capacity)

    def download(self, copy=True) -> BatchResult:
        r = SiftBatchResult()
        self._check(self.L.sift_batch_download(self.ctx, C.byref(r)))
        return self._wrap(r, copy)

    def detect_and_describe(self, frames: Sequence[np.ndarray], copy=True, allow_capacity=False) -> BatchResult:
        frames = [self._as_bgra(f) for f in frames]
        ptrs = (C.c_void_p * len(frames))(*[f.ctypes.data for f in frames])
        r = SiftBatchResult()
        st = self._check(
            self.L.sift_detect_and_describe_batch(self.ctx, ptrs, len(frames), frames[0].strides[0], C.byref(r)),
            allow_capacity=allow_capacity,
        )
        out = self._wrap(r, copy)
        out.status = st
        return out

    def detect_and_describe_ptrs(self, ptr_array, n, pitch_bytes):
        """Hot-loop variant for the bench: pre-built (c_void_p * n) of pinned host frames;
        returns (total_keypoints, total_descriptors) without copying results again."""
        r = SiftBatchResult()
        self._check(self.L.sift_detect_and_describe_batch(self.ctx, ptr_array, n, pitch_bytes, C.byref(r)))
        return int(r.total_keypoints), int(r.total_descriptors)

    # -- diagnostics -----------------------------------------------------------------------------
    def timings(self) -> dict:
        t = SiftTimings()
        self._check(self.L.sift_last_timings(self.ctx, C.byref(t)))
        d = {"total_ms": t.total_ms, "kernel_launches": t.kernel_launches,
             "blur_octave0_ms": t.blur_octave0_ms, "blur_octave0_launches": t.blur_octave0_launches,
             "blur_octave0_launch_ms": list(t.blur_octave0_launch_ms)}
        for i, name in enumerate(_abi.STAGE_NAMES):
            d[name + "_ms"] = t.stage_ms[i]
        return d

    def set_stage_timing(self, enabled):
        self._check(self.L.sift_set_stage_timing(self.ctx, int(enabled)))

    def blur_bench(self, scale, mode=0, iters=20):
        ms = C.c_float()
        self._check(self.L.sift_debug_blur_bench(self.ctx, scale, mode, iters, C.byref(ms)))
        return ms.value

    def plane(self, what, octave=0, slice=0, frame=0):
        if what == _abi.PLANE_GRAY:
            shape = (self.height, self.width)
        elif what == _abi.PLANE_SEED:
            shape = (self.info.octave_height[0], self.info.octave_width[0])
        elif what == _abi.PLANE_GRADIENT:
            shape = (self.info.octave_height[octave], self.info.octave_width[octave], 2)
        else:
            shape = (self.info.octave_height[octave], self.info.octave_width[octave])
        out = np.zeros(shape, dtype=np.float32)
        self._check(self.L.sift_debug_download(self.ctx, what, frame, octave, slice, out.ctypes.data, out.size))
        return out

    def candidates(self, octave, frame=0):
        n = self.L.sift_debug_candidates(self.ctx, frame, octave, None, 0)
        if n < 0:
            raise SiftError(-n, "sift_debug_candidates")
        out = np.zeros((n, 3), dtype=np.int32)
        if n:
            self.L.sift_debug_candidates(self.ctx, frame, octave, out.ctypes.data, n)
        return out

    # -- helpers ---------------------------------------------------------------------------------
    def _as_bgra(self, a):
        a = np.asarray(a)
        if a.dtype != np.uint8 or a.ndim != 3 or a.shape[2] != 4 or a.shape[:2] != (self.height, self.width):
            # precondition of ConvertSRGBToGrayscaleKernel.swift:34 (bgra8Unorm, matching size)
            raise ValueError(f"expected uint8 BGRA8 frame of shape ({self.height}, {self.width}, 4), got {a.dtype} {a.shape}")
        if a.strides[2] != 1 or a.strides[1] != 4:
            a = np.ascontiguousarray(a)
        return a

    @staticmethod
    def _view(addr, n, dtype):
        if n == 0 or not addr:
            return np.zeros(0, dtype=dtype)
        buf = (C.c_char * (n * dtype.itemsize)).from_address(addr)
        return np.frombuffer(buf, dtype=dtype, count=n)

    def _wrap(self, r: SiftBatchResult, copy) -> BatchResult:
        nf = r.n_frames
        shape = (nf, NUM_OCTAVES)
        kc = np.ctypeslib.as_array(r.keypoint_counts, shape=shape).copy()
        dc = np.ctypeslib.as_array(r.descriptor_counts, shape=shape).copy()
        cc = np.ctypeslib.as_array(r.candidate_counts, shape=shape).copy()
        k = self._view(r.keypoints, int(r.total_keypoints), KEYPOINT_DTYPE)
        d = self._view(r.descriptors, int(r.total_descriptors), DESCRIPTOR_DTYPE)
        if copy:
            k, d = k.copy(), d.copy()
        return BatchResult(k, d, kc, dc, cc)


def device_math(op, a, b=None, device=0):
    L = load_library()
    a = np.ascontiguousarray(a, dtype=np.float32)
    bb = np.ascontiguousarray(b if b is not None else np.zeros_like(a), dtype=np.float32)
    out = np.zeros_like(a)
    st = L.sift_debug_math(device, op, a.ctypes.data, bb.ctypes.data, out.ctypes.data, a.size)
    if st != 0:
        raise SiftError(st, L.sift_status_string(st).decode())
    return out


# ------------------------------------------------------------------------------------------------
# Reference-shaped object API


@dataclass
class IntegralSize:  # Utilities/Math.swift:11-19
    width: int
    height: int


@dataclass
class SIFTKeypoint:  # SIFTKeypoint.swift:11-57
    octave: int
    scale: int
    subScale: float
    scaledCoordinate: tuple
    absoluteCoordinate: tuple
    normalizedCoordinate: tuple
    sigma: float
    value: float

    @staticmethod
    def from_record(r):
        return SIFTKeypoint(int(r["octave"]), int(r["scale"]), float(r["subScale"]),
                            (int(r["scaledX"]), int(r["scaledY"])),
                            (float(r["absoluteX"]), float(r["absoluteY"])),
                            (float(r["normalizedX"]), float(r["normalizedY"])),
                            float(r["sigma"]), float(r["value"]))

    def to_record(self):
        r = np.zeros((), dtype=KEYPOINT_DTYPE)
        r["octave"], r["scale"], r["subScale"] = self.octave, self.scale, np.float32(self.subScale)
        r["scaledX"], r["scaledY"] = self.scaledCoordinate
        r["absoluteX"], r["absoluteY"] = self.absoluteCoordinate
        r["normalizedX"], r["normalizedY"] = self.normalizedCoordinate
        r["sigma"], r["value"] = self.sigma, self.value
        return r


class IntVector:  # Utilities/Vector.swift:12-60
    def __init__(self, components):
        components = [int(c) for c in components]
        if not components:
            raise ValueError("IntVector must not be empty")  # precondition(!components.isEmpty)
        self.components = components
        self.count = len(components)

    def __getitem__(self, i):
        return self.components[i]

    def __eq__(self, other):
        return isinstance(other, IntVector) and self.components == other.components

    def distanceSquared(self, other):
        assert self.count == other.count
        return float(sum((b - a) * (b - a) for a, b in zip(self.components, other.components)))

    def distance(self, other):
        return float(np.sqrt(np.float32(self.distanceSquared(other))))


class SIFTDescriptor:  # SIFTDescriptor.swift:12-90 (stored properties; index keys built lazily)
    def __init__(self, keypoint: SIFTKeypoint, theta: float, features: IntVector):
        if features.count <= 0:
            raise ValueError("features must not be empty")
        self.keypoint = keypoint
        self.theta = theta
        self.features = features

    @property
    def rawFeatures(self):  # SIFTDescriptor.swift:37-41
        return [np.float32(c) / np.float32(255) for c in self.features.components]


class SIFT:
    """Drop-in for the reference's `SIFT` class (SIFT.swift:53-239) on a B200."""

    @dataclass
    class Configuration:  # SIFT.swift:57-103 — only inputSize is settable, as in the reference
        inputSize: IntegralSize

    def __init__(self, device: int, configuration: "SIFT.Configuration"):
        self.configuration = configuration
        self._engine = Engine(configuration.inputSize.width, configuration.inputSize.height, device=device)
        self._last = None

    @property
    def engine(self) -> Engine:
        return self._engine

    def getKeypoints(self, inputTexture: np.ndarray) -> List[List[SIFTKeypoint]]:
        """`inputTexture`: H×W×4 uint8 BGRA8 (was: a bgra8Unorm MTLTexture)."""
        kps, counts = self._engine.detect(inputTexture)
        out, k = [], 0
        for o in range(NUM_OCTAVES):
            out.append([SIFTKeypoint.from_record(kps[k + i]) for i in range(int(counts[o]))])
            k += int(counts[o])
        return out

    def getDescriptors(self, keypointOctaves: List[List[SIFTKeypoint]]) -> List[List[SIFTDescriptor]]:
        if len(keypointOctaves) != NUM_OCTAVES:
            raise ValueError("keypointOctaves.count must equal the number of octaves")  # SIFT.swift:208
        counts = [len(o) for o in keypointOctaves]
        flat = [k for o in keypointOctaves for k in o]
        recs = np.zeros(len(flat), dtype=KEYPOINT_DTYPE)
        for i, k in enumerate(flat):
            recs[i] = k.to_record()
        desc, dcounts = self._engine.describe(recs, counts)
        out, d = [], 0
        for o in range(NUM_OCTAVES):
            lst = []
            for i in range(int(dcounts[o])):
                r = desc[d + i]
                lst.append(SIFTDescriptor(flat[int(r["keypoint"])], float(r["theta"]), IntVector(r["features"].tolist())))
            out.append(lst)
            d += int(dcounts[o])
        return out

    def close(self):
        self._engine.close()
