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
    vp, i32, i64, f32 = C.c_void_p, C.c_int32, C.c_int64, C.c_float
    L.sift_config_default.argtypes = [C.POINTER(SiftConfig), i32, i32]
    L.sift_create.argtypes = [C.POINTER(SiftConfig), C.c_int, C.POINTER(vp)]
    L.sift_destroy.argtypes = [vp]
    L.sift_destroy.restype = None
    L.sift_get_info.argtypes = [vp, C.POINTER(SiftInfo)]
    L.sift_detect.argtypes = [vp, vp, i32, C.POINTER(vp), C.POINTER(i32)]
    L.sift_describe.argtypes = [vp, vp, C.POINTER(i32), C.POINTER(vp), C.POINTER(i32)]
    L.sift_detect_and_describe_batch.argtypes = [vp, C.POINTER(vp), i32, i32, C.POINTER(SiftBatchResult)]
    L.sift_submit.argtypes = [vp, C.POINTER(vp), i32, i32]
    L.sift_wait.argtypes = [vp, C.POINTER(SiftBatchResult)]
    L.sift_pending.argtypes = [vp]
    L.sift_next_slot.argtypes = [vp]
    L.sift_result_layout.argtypes = [vp, C.POINTER(_abi.SiftResultLayout)]
    L.sift_register_host_memory.argtypes = [vp, vp, i64]
    L.sift_bind_result_memory.argtypes = [vp, i32, vp, i64]
    L.sift_batch_upload.argtypes = [vp, C.POINTER(vp), i32, i32]
    L.sift_batch_set_device_input.argtypes = [vp, vp, i32, i32, i64]
    L.sift_batch_execute.argtypes = [vp]
    L.sift_batch_download.argtypes = [vp, C.POINTER(SiftBatchResult)]
    L.sift_materialize_keypoints.argtypes = [vp, C.POINTER(SiftBatchResult), i64, i64, vp]
    L.sift_materialize_descriptors.argtypes = [C.POINTER(SiftBatchResult), i64, i64, vp]
    L.sift_match.argtypes = [vp, vp, i64, vp, i64, f32, f32, C.POINTER(vp), C.POINTER(i64)]
    L.sift_match_frames.argtypes = [vp, i32, i32, f32, f32, C.POINTER(vp), C.POINTER(i64)]
    L.sift_match_geometry.argtypes = [vp, vp, vp, i64, vp, vp, i64, f32, f32, C.POINTER(f32)]
    L.sift_approximate_match.argtypes = [vp, vp, i64, vp, i64, f32, f32, C.POINTER(vp), C.POINTER(i64)]
    L.sift_status_string.argtypes = [C.c_int]
    L.sift_status_string.restype = C.c_char_p
    L.sift_last_error_string.argtypes = [vp]
    L.sift_last_error_string.restype = C.c_char_p
    L.sift_set_stage_timing.argtypes = [vp, i32]
    L.sift_set_graph_replay.argtypes = [vp, i32]
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
    "sift_detect_and_describe_batch sift_submit sift_wait sift_pending sift_next_slot sift_result_layout "
    "sift_register_host_memory sift_bind_result_memory sift_batch_upload "
    "sift_batch_set_device_input sift_batch_execute sift_batch_download sift_materialize_keypoints "
    "sift_materialize_descriptors sift_match sift_match_frames sift_match_geometry "
    "sift_approximate_match sift_status_string "
    "sift_last_error_string sift_set_stage_timing sift_set_graph_replay sift_last_timings sift_debug_download "
    "sift_debug_candidates sift_debug_math sift_debug_blur_bench"
).split()


def _view(addr, n, dtype, shape=None):
    """Zero-copy numpy view of `n` items of `dtype` at a C address (context-owned pinned memory)."""
    dtype = np.dtype(dtype)
    if n <= 0 or not addr:
        return np.zeros((0,) + tuple(shape[1:]) if shape else 0, dtype=dtype)
    buf = (C.c_char * (n * dtype.itemsize)).from_address(addr)
    a = np.frombuffer(buf, dtype=dtype, count=n)
    return a.reshape(shape) if shape else a


class KeypointColumns:
    """Lazy view of the keypoint columns of a result (SiftKeypointColumns): the arrays are the
    wire format; a `SIFTKeypoint` object is only built when one is indexed, a record array only
    when `.records()` is asked for. The reference materialises every keypoint eagerly
    (SIFTOctave.swift:257-286)."""

    def __init__(self, absolute_x, absolute_y, sigma, value, sub_scale, scaled_xy, octave_scale, octave_sizes):
        self.absolute_x, self.absolute_y = absolute_x, absolute_y
        self.sigma, self.value, self.sub_scale = sigma, value, sub_scale
        self.scaled_xy, self.octave_scale = scaled_xy, octave_scale   # [n, 2] int16 / uint8
        self._sizes = octave_sizes                                     # [7, 2] float32 (w, h)

    def __len__(self):
        return len(self.sigma)

    def copy(self):
        return KeypointColumns(self.absolute_x.copy(), self.absolute_y.copy(), self.sigma.copy(), self.value.copy(),
                               self.sub_scale.copy(), self.scaled_xy.copy(), self.octave_scale.copy(), self._sizes)

    def slice(self, a, b):
        return KeypointColumns(self.absolute_x[a:b], self.absolute_y[a:b], self.sigma[a:b], self.value[a:b],
                               self.sub_scale[a:b], self.scaled_xy[a:b], self.octave_scale[a:b], self._sizes)

    def __getitem__(self, i):
        if isinstance(i, slice):
            a, b, st = i.indices(len(self))
            if st != 1:
                raise IndexError("contiguous slices only")
            return self.slice(a, b)
        o, s = int(self.octave_scale[i, 0]), int(self.octave_scale[i, 1])
        x, y = int(self.scaled_xy[i, 0]), int(self.scaled_xy[i, 1])
        w, h = self._sizes[o]
        return SIFTKeypoint(o, s, float(self.sub_scale[i]), (x, y),
                            (float(self.absolute_x[i]), float(self.absolute_y[i])),
                            (float(np.float32(x) / w), float(np.float32(y) / h)),
                            float(self.sigma[i]), float(self.value[i]))

    def records(self):
        """KEYPOINT_DTYPE record array (the reference's SIFTKeypoint fields), vectorised."""
        r = np.zeros(len(self), dtype=KEYPOINT_DTYPE)
        o = self.octave_scale[:, 0].astype(np.int32)
        r["octave"], r["scale"] = o, self.octave_scale[:, 1]
        r["subScale"] = self.sub_scale
        r["scaledX"], r["scaledY"] = self.scaled_xy[:, 0], self.scaled_xy[:, 1]
        r["absoluteX"], r["absoluteY"] = self.absolute_x, self.absolute_y
        if len(self):
            r["normalizedX"] = r["scaledX"].astype(np.float32) / self._sizes[o, 0]   # SIFTOctave.swift:278-281
            r["normalizedY"] = r["scaledY"].astype(np.float32) / self._sizes[o, 1]
        r["sigma"], r["value"] = self.sigma, self.value
        return r


class DescriptorColumns:
    """Lazy view of the descriptor columns (SiftDescriptorColumns): dense [n, 128] uint8 feature
    matrix + theta + keypoint index. `SIFTDescriptor` objects are built on indexing only (the
    reference's initialiser does float copies and a re-ordering per descriptor,
    SIFTDescriptor.swift:36-89)."""

    def __init__(self, features, theta, keypoint, keypoints: Optional[KeypointColumns] = None):
        self.features, self.theta, self.keypoint = features, theta, keypoint
        self._kps = keypoints

    def __len__(self):
        return len(self.theta)

    def copy(self):
        return DescriptorColumns(self.features.copy(), self.theta.copy(), self.keypoint.copy(), self._kps)

    def slice(self, a, b, keypoints=None):
        return DescriptorColumns(self.features[a:b], self.theta[a:b], self.keypoint[a:b],
                                 keypoints if keypoints is not None else self._kps)

    def __getitem__(self, i):
        kp = self._kps[int(self.keypoint[i])] if self._kps is not None else None
        return SIFTDescriptor(kp, float(self.theta[i]), IntVector(self.features[i].tolist()))

    def records(self):
        r = np.zeros(len(self), dtype=DESCRIPTOR_DTYPE)
        r["keypoint"], r["theta"], r["features"] = self.keypoint, self.theta, self.features
        return r


class BatchResult:
    """Result of one batch: the column views + per-(frame, octave) counts. `.keypoints` /
    `.descriptors` give record arrays (KEYPOINT_DTYPE / DESCRIPTOR_DTYPE), built on first use."""

    def __init__(self, keypoint_columns, descriptor_columns, keypoint_counts, descriptor_counts, candidate_counts,
                 status=0, slot=0):
        self.keypoint_columns = keypoint_columns
        self.descriptor_columns = descriptor_columns
        self.keypoint_counts = keypoint_counts        # [n_frames, 7]
        self.descriptor_counts = descriptor_counts    # [n_frames, 7]
        self.candidate_counts = candidate_counts      # [n_frames, 7]
        self.status = status
        self.slot = slot                              # in-flight slot that holds the (uncopied) columns
        self._kp_records = self._desc_records = None

    @property
    def keypoints(self):
        if self._kp_records is None:
            self._kp_records = self.keypoint_columns.records()
        return self._kp_records

    @property
    def descriptors(self):
        if self._desc_records is None:
            self._desc_records = self.descriptor_columns.records()
        return self._desc_records

    def _range(self, f):
        k0 = int(self.keypoint_counts[:f].sum())
        k1 = k0 + int(self.keypoint_counts[f].sum())
        d0 = int(self.descriptor_counts[:f].sum())
        d1 = d0 + int(self.descriptor_counts[f].sum())
        return k0, k1, d0, d1

    def frame(self, f):
        """Record arrays of frame f."""
        k0, k1, d0, d1 = self._range(f)
        return self.keypoints[k0:k1], self.descriptors[d0:d1]

    def frame_view(self, f):
        """Lazy column views of frame f (no records are built)."""
        k0, k1, d0, d1 = self._range(f)
        kv = self.keypoint_columns.slice(k0, k1)
        return kv, self.descriptor_columns.slice(d0, d1, keypoints=kv)


class Engine:
    """Thin owner of one SiftContext (one GPU, one stream). Not thread-safe, like the
    reference's SIFT instance (one MTLCommandQueue, shared scratch buffers)."""

    def __init__(self, width, height, device=0, max_batch=1, input_format=_abi.INPUT_BGRA8, **overrides):
        self.L = load_library()
        self.cfg = SiftConfig()
        self._check(self.L.sift_config_default(C.byref(self.cfg), width, height), ctx=False)
        self.cfg.max_batch = max_batch
        self.cfg.input_format = input_format
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
        self.input_format = input_format
        self.bytes_per_pixel = _abi.INPUT_BYTES_PER_PIXEL.get(input_format, 4)
        self.info = SiftInfo()
        self.L.sift_get_info(self.ctx, C.byref(self.info))
        self._octave_sizes = np.array([[self.info.octave_width[o], self.info.octave_height[o]] for o in range(NUM_OCTAVES)],
                                      dtype=np.float32)
        self._inflight = []   # host frames of submitted calls, kept alive until their wait

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
    def detect(self, frame):
        frame = self._as_frame(frame)
        out = C.c_void_p()
        counts = (C.c_int32 * NUM_OCTAVES)()
        self._check(self.L.sift_detect(self.ctx, frame.ctypes.data, frame.strides[0], C.byref(out), counts))
        counts = np.array(counts, dtype=np.int32)
        n = int(counts.sum())
        kps = _view(out.value, n, KEYPOINT_DTYPE).copy()
        return kps, counts

    def describe(self, keypoints, counts):
        keypoints = np.ascontiguousarray(keypoints, dtype=KEYPOINT_DTYPE)
        counts = [int(c) for c in counts]
        if len(counts) != NUM_OCTAVES or sum(counts) != len(keypoints):
            raise ValueError("counts must hold 7 per-octave counts that add up to len(keypoints)")
        cin = (C.c_int32 * NUM_OCTAVES)(*counts)
        out = C.c_void_p()
        cout = (C.c_int32 * NUM_OCTAVES)()
        self._check(self.L.sift_describe(self.ctx, keypoints.ctypes.data, cin, C.byref(out), cout))
        cout = np.array(cout, dtype=np.int32)
        return _view(out.value, int(cout.sum()), DESCRIPTOR_DTYPE).copy(), cout

    # -- batch path ----------------------------------------------------------------------------
    def _frame_ptrs(self, frames):
        """Host frames of one call: all C-contiguous (one pitch for the whole batch)."""
        frames = [self._as_frame(f) for f in frames]
        ptrs = (C.c_void_p * len(frames))(*[f.ctypes.data for f in frames])
        return frames, ptrs, self.width * self.bytes_per_pixel

    def upload(self, frames: Sequence[np.ndarray]):
        frames, ptrs, pitch = self._frame_ptrs(frames)
        self._check(self.L.sift_batch_upload(self.ctx, ptrs, len(frames), pitch))   # copies done at return

    def set_device_input(self, device_ptr, n, pitch_bytes, frame_stride_bytes):
        self._check(self.L.sift_batch_set_device_input(self.ctx, device_ptr, n, pitch_bytes, frame_stride_bytes))

    def execute(self, allow_capacity=False):
        return self._check(self.L.sift_batch_execute(self.ctx), allow_capacity=allow_capacity)

    def download(self, copy=True) -> BatchResult:
        r = SiftBatchResult()
        self._check(self.L.sift_batch_download(self.ctx, C.byref(r)))
        return self._wrap(r, copy)

    def detect_and_describe(self, frames: Sequence[np.ndarray], copy=True, allow_capacity=False) -> BatchResult:
        frames, ptrs, pitch = self._frame_ptrs(frames)
        r = SiftBatchResult()
        st = self._check(
            self.L.sift_detect_and_describe_batch(self.ctx, ptrs, len(frames), pitch, C.byref(r)),
            allow_capacity=allow_capacity,
        )
        out = self._wrap(r, copy)
        out.status = st
        return out

    def detect_and_describe_ptrs(self, ptr_array, n, pitch_bytes):
        """Hot-loop variant for the bench: pre-built (c_void_p * n) of pinned host frames;
        returns (total_keypoints, total_descriptors) without touching the result columns."""
        r = SiftBatchResult()
        self._check(self.L.sift_detect_and_describe_batch(self.ctx, ptr_array, n, pitch_bytes, C.byref(r)))
        return int(r.total_keypoints), int(r.total_descriptors)

    # pipelined form: up to two calls in flight (upload of call i + 1 under the kernels of call i)
    def submit(self, frames: Sequence[np.ndarray]):
        frames, ptrs, pitch = self._frame_ptrs(frames)
        self._check(self.L.sift_submit(self.ctx, ptrs, len(frames), pitch))
        self._inflight.append((frames, ptrs))   # must stay valid and unmodified until wait()

    def submit_ptrs(self, ptr_array, n, pitch_bytes):
        self._check(self.L.sift_submit(self.ctx, ptr_array, n, pitch_bytes))
        self._inflight.append(None)

    def wait(self, copy=True, allow_capacity=False) -> BatchResult:
        r = SiftBatchResult()
        st = self._check(self.L.sift_wait(self.ctx, C.byref(r)), allow_capacity=allow_capacity)
        if self._inflight:
            self._inflight.pop(0)
        out = self._wrap(r, copy)
        out.status = st
        return out

    def wait_counts(self):
        """Bench variant of wait(): totals only."""
        r = SiftBatchResult()
        self._check(self.L.sift_wait(self.ctx, C.byref(r)))
        if self._inflight:
            self._inflight.pop(0)
        return int(r.total_keypoints), int(r.total_descriptors)

    def pending(self):
        return int(self.L.sift_pending(self.ctx))

    def device_bytes(self):
        """Device memory the context holds now (grows when the second pipeline / slot is created)."""
        info = SiftInfo()
        self.L.sift_get_info(self.ctx, C.byref(info))
        return int(info.device_bytes)

    def next_slot(self):
        return int(self.L.sift_next_slot(self.ctx))

    def result_layout(self):
        lay = _abi.SiftResultLayout()
        self._check(self.L.sift_result_layout(self.ctx, C.byref(lay)))
        return lay

    def register_host_memory(self, address, nbytes):
        self._check(self.L.sift_register_host_memory(self.ctx, address, nbytes))

    def bind_result_memory(self, slot, address, nbytes):
        """Slot `slot` (0 / 1) writes its result columns into caller memory (e.g. a shared-memory
        mapping read by another process) instead of its own pinned block."""
        self._check(self.L.sift_bind_result_memory(self.ctx, slot, address, nbytes))

    # -- matching (SIFTDescriptor.match, SIFTDescriptor.swift:298-361) ----------------------------
    def match(self, source_features, target_features, absolute_threshold=300.0, relative_threshold=0.6):
        a = np.ascontiguousarray(source_features, dtype=np.uint8).reshape(-1, 128)
        b = np.ascontiguousarray(target_features, dtype=np.uint8).reshape(-1, 128)
        out, n = C.c_void_p(), C.c_int64()
        self._check(self.L.sift_match(self.ctx, a.ctypes.data, len(a), b.ctypes.data, len(b),
                                      absolute_threshold, relative_threshold, C.byref(out), C.byref(n)))
        return _view(out.value, n.value, _abi.MATCH_DTYPE).copy()

    def match_frames(self, source_frame, target_frame, absolute_threshold=300.0, relative_threshold=0.6):
        out, n = C.c_void_p(), C.c_int64()
        self._check(self.L.sift_match_frames(self.ctx, source_frame, target_frame, absolute_threshold,
                                             relative_threshold, C.byref(out), C.byref(n)))
        return _view(out.value, n.value, _abi.MATCH_DTYPE).copy()

    def match_geometry(self, source_features, source_xy, target_features, target_xy, absolute_threshold=1.176,
                       relative_threshold=0.6):
        """SIFTDescriptor.matchGeometry (SIFTDescriptor.swift:104-296): geometric-consistency score."""
        a = np.ascontiguousarray(source_features, dtype=np.uint8).reshape(-1, 128)
        b = np.ascontiguousarray(target_features, dtype=np.uint8).reshape(-1, 128)
        axy = np.ascontiguousarray(source_xy, dtype=np.float32).reshape(-1, 2)
        bxy = np.ascontiguousarray(target_xy, dtype=np.float32).reshape(-1, 2)
        if len(axy) != len(a) or len(bxy) != len(b):
            raise ValueError("one (x, y) pair per descriptor row")
        score = C.c_float()
        self._check(self.L.sift_match_geometry(self.ctx, a.ctypes.data, axy.ctypes.data, len(a), b.ctypes.data,
                                               bxy.ctypes.data, len(b), absolute_threshold, relative_threshold,
                                               C.byref(score)))
        return score.value

    def approximate_match(self, source_features, target_features, absolute_threshold=300.0, relative_threshold=0.6):
        """SIFTDescriptor.approximateMatch over the trie ANN (SIFTDescriptor.swift:362-417, Utilities/Trie.swift)."""
        a = np.ascontiguousarray(source_features, dtype=np.uint8).reshape(-1, 128)
        b = np.ascontiguousarray(target_features, dtype=np.uint8).reshape(-1, 128)
        out, n = C.c_void_p(), C.c_int64()
        self._check(self.L.sift_approximate_match(self.ctx, a.ctypes.data, len(a), b.ctypes.data, len(b),
                                                  absolute_threshold, relative_threshold, C.byref(out), C.byref(n)))
        return _view(out.value, n.value, _abi.MATCH_DTYPE).copy()

    # -- diagnostics -----------------------------------------------------------------------------
    def timings(self) -> dict:
        t = SiftTimings()
        self._check(self.L.sift_last_timings(self.ctx, C.byref(t)))
        d = {"total_ms": t.total_ms, "kernel_launches": t.kernel_launches,
             "blur_octave0_ms": t.blur_octave0_ms, "blur_octave0_launches": t.blur_octave0_launches,
             "blur_octave0_launch_ms": list(t.blur_octave0_launch_ms),
             "stage_timing_enabled": bool(t.stage_timing_enabled), "graph_replay": bool(t.graph_replay)}
        for i, name in enumerate(_abi.STAGE_NAMES):
            d[name + "_ms"] = t.stage_ms[i]
        return d

    def set_stage_timing(self, enabled):
        self._check(self.L.sift_set_stage_timing(self.ctx, int(enabled)))

    def set_graph_replay(self, enabled):
        self._check(self.L.sift_set_graph_replay(self.ctx, int(enabled)))

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
    def _as_frame(self, a):
        """One host frame in the context's input format, C-contiguous (every frame of a batch is
        passed with the same pitch, so row-strided views are copied)."""
        a = np.asarray(a)
        want = (self.height, self.width, 4) if self.bytes_per_pixel == 4 else (self.height, self.width)
        if a.dtype != np.uint8 or a.shape != want:
            # precondition of ConvertSRGBToGrayscaleKernel.swift:34 (pixel format, matching size)
            raise ValueError(f"expected uint8 frame of shape {want}, got {a.dtype} {a.shape}")
        return np.ascontiguousarray(a)

    def _wrap(self, r: SiftBatchResult, copy) -> BatchResult:
        nf = r.n_frames
        shape = (nf, NUM_OCTAVES)
        kc = np.ctypeslib.as_array(r.keypoint_counts, shape=shape).copy()
        dc = np.ctypeslib.as_array(r.descriptor_counts, shape=shape).copy()
        cc = np.ctypeslib.as_array(r.candidate_counts, shape=shape).copy()
        nk, nd = int(r.total_keypoints), int(r.total_descriptors)
        k = r.keypoints
        kv = KeypointColumns(_view(k.absolute_x, nk, "<f4"), _view(k.absolute_y, nk, "<f4"), _view(k.sigma, nk, "<f4"),
                             _view(k.value, nk, "<f4"), _view(k.sub_scale, nk, "<f4"),
                             _view(k.scaled_xy, 2 * nk, "<i2").reshape(-1, 2), _view(k.octave_scale, 2 * nk, "u1").reshape(-1, 2),
                             self._octave_sizes)
        d = r.descriptors
        dv = DescriptorColumns(_view(d.features, 128 * nd, "u1").reshape(-1, 128), _view(d.theta, nd, "<f4"),
                               _view(d.keypoint, nd, "<i4"))
        if copy:   # the pinned columns are reused by a later call
            kv, dv = kv.copy(), dv.copy()
        return BatchResult(kv, dv, kc, dc, cc, status=int(r.status), slot=int(r.slot))


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


class LazyKeypointList:
    """Sequence of SIFTKeypoint over packed KEYPOINT_DTYPE records; elements are built on access
    (and cached, so that descriptors can refer to the very object the caller holds)."""

    def __init__(self, records):
        self.records = records
        self._cache = {}

    def __len__(self):
        return len(self.records)

    def __getitem__(self, i):
        if isinstance(i, slice):
            return [self[j] for j in range(*i.indices(len(self)))]
        if i < 0:
            i += len(self)
        k = self._cache.get(i)
        if k is None:
            k = self._cache[i] = SIFTKeypoint.from_record(self.records[i])
        return k

    def __iter__(self):
        for i in range(len(self)):
            yield self[i]


class _Concat:
    """Flat index over per-octave lazy lists."""

    def __init__(self, lists):
        self._lists = lists
        self._starts = np.cumsum([0] + [len(l) for l in lists])

    def __len__(self):
        return int(self._starts[-1])

    def __getitem__(self, i):
        o = int(np.searchsorted(self._starts, i, side="right")) - 1
        return self._lists[o][i - int(self._starts[o])]


class LazyDescriptorList:
    """Sequence of SIFTDescriptor over packed records; elements are materialised on access."""

    def __init__(self, records, keypoints):
        self.records = records          # DESCRIPTOR_DTYPE slice
        self._keypoints = keypoints     # the caller's flat keypoint list (records index into it)

    def __len__(self):
        return len(self.records)

    def __getitem__(self, i):
        if isinstance(i, slice):
            return [self[j] for j in range(*i.indices(len(self)))]
        r = self.records[i]
        return SIFTDescriptor(self._keypoints[int(r["keypoint"])], float(r["theta"]), IntVector(r["features"].tolist()))

    def __iter__(self):
        for i in range(len(self)):
            yield self[i]

    @property
    def thetas(self):
        return self.records["theta"]

    @property
    def features(self):
        return self.records["features"]


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

    def getKeypoints(self, inputTexture: np.ndarray) -> List[LazyKeypointList]:
        """`inputTexture`: H×W×4 uint8 BGRA8 (was: a bgra8Unorm MTLTexture). Returns the 7 per-octave
        keypoint lists (SIFT.swift:147-152) as lazy sequences over the packed records."""
        kps, counts = self._engine.detect(inputTexture)
        out, k = [], 0
        for o in range(NUM_OCTAVES):
            out.append(LazyKeypointList(kps[k:k + int(counts[o])]))
            k += int(counts[o])
        return out

    def getDescriptors(self, keypointOctaves: List[List[SIFTKeypoint]]) -> List["LazyDescriptorList"]:
        """[[SIFTDescriptor]] indexed by octave. Each inner list is a lazy sequence over the packed
        descriptor records: a `SIFTDescriptor` (with its IntVector of 128 ints) is only built when an
        element is accessed — SIFTDescriptor.init's per-descriptor work (SIFTDescriptor.swift:36-89)
        is not paid for descriptors nobody looks at."""
        if len(keypointOctaves) != NUM_OCTAVES:
            raise ValueError("keypointOctaves.count must equal the number of octaves")  # SIFT.swift:208
        counts = [len(o) for o in keypointOctaves]
        if all(isinstance(o, LazyKeypointList) for o in keypointOctaves):
            recs = np.concatenate([o.records for o in keypointOctaves])      # no per-keypoint objects
            flat = _Concat(keypointOctaves)
        else:
            flat = [k for o in keypointOctaves for k in o]
            recs = np.zeros(len(flat), dtype=KEYPOINT_DTYPE)
            for i, k in enumerate(flat):
                recs[i] = k.to_record()
        desc, dcounts = self._engine.describe(recs, counts)
        out, d = [], 0
        for o in range(NUM_OCTAVES):
            out.append(LazyDescriptorList(desc[d:d + int(dcounts[o])], flat))
            d += int(dcounts[o])
        return out

    def close(self):
        self._engine.close()
