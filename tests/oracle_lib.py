"""ctypes loader for the CPU oracle (oracle/libsiftoracle.so) — test infrastructure only."""
import ctypes as C
import os
import subprocess

import numpy as np

from siftmetal_b200 import _abi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
ORACLE_SO = os.path.join(ORACLE_DIR, "libsiftoracle.so")

_lib = None


def build_oracle(force=False):
    srcs = [os.path.join(ORACLE_DIR, f) for f in ("sift_oracle.cpp", "oracle_math.h")]
    srcs.append(os.path.join(ROOT, "include", "siftcuda.h"))
    stale = not os.path.exists(ORACLE_SO) or any(
        os.path.getmtime(s) > os.path.getmtime(ORACLE_SO) for s in srcs if os.path.exists(s)
    )
    if force or stale:
        subprocess.run(["make", "-C", ORACLE_DIR, "-B"], check=True, capture_output=True)
    return ORACLE_SO


def lib():
    global _lib
    if _lib is None:
        build_oracle()
        L = C.CDLL(ORACLE_SO)
        L.oracle_create.restype = C.c_void_p
        L.oracle_create.argtypes = [C.POINTER(_abi.SiftConfig)]
        L.oracle_destroy.argtypes = [C.c_void_p]
        L.oracle_set_options.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.oracle_set_threads.argtypes = [C.c_int]
        L.oracle_max_threads.restype = C.c_int
        L.oracle_detect.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        L.oracle_describe.argtypes = [C.c_void_p]
        L.oracle_describe_keypoints.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.oracle_get_keypoints.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.oracle_get_descriptors.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.oracle_get_candidates.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int]
        L.oracle_get_stats.argtypes = [C.c_void_p, C.c_void_p]
        L.oracle_get_plane.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]
        L.oracle_get_info.argtypes = [C.c_void_p, C.POINTER(_abi.SiftInfo)]
        L.oracle_math.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64]
        L.oracle_match.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_float, C.c_float, C.c_void_p]
        L.oracle_match.restype = C.c_int64
        L.oracle_approximate_match.argtypes = L.oracle_match.argtypes
        L.oracle_approximate_match.restype = C.c_int64
        L.oracle_match_geometry.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int64,
                                            C.c_float, C.c_float]
        L.oracle_match_geometry.restype = C.c_float
        _lib = L
    return _lib


def default_config(width, height):
    """The literals of the reference (SIFT.swift:57-103, SIFTOctave.swift:217-226,296-300)."""
    c = _abi.SiftConfig()
    c.width, c.height, c.max_batch = width, height, 1
    c.dog_threshold = 0.0133
    c.edge_threshold = 10.0
    c.max_interpolation_iterations = 5
    c.max_offset = 0.6
    c.image_border = 5
    c.lambda_orientation = 1.5
    c.orientation_threshold = 0.8
    c.orientation_smoothing_iterations = 6
    return c


class Oracle:
    """One reference-equivalent SIFT instance on the CPU (SIFT.swift:112-143)."""

    def __init__(self, width, height, collect_stats=False, all_gradients=False, threads=0):
        self.L = lib()
        self.width, self.height = width, height
        self.cfg = default_config(width, height)
        self.h = self.L.oracle_create(C.byref(self.cfg))
        if not self.h:
            raise ValueError("oracle_create failed")
        self.L.oracle_set_options(self.h, int(collect_stats), int(all_gradients))
        self.L.oracle_set_threads(threads)
        self.info = _abi.SiftInfo()
        self.L.oracle_get_info(self.h, C.byref(self.info))

    def close(self):
        if self.h:
            self.L.oracle_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def detect(self, bgra):
        bgra = np.ascontiguousarray(bgra, dtype=np.uint8)
        assert bgra.shape == (self.height, self.width, 4)
        n = self.L.oracle_detect(self.h, bgra.ctypes.data, bgra.strides[0])
        kps = np.zeros(n, dtype=_abi.KEYPOINT_DTYPE)
        counts = np.zeros(_abi.NUM_OCTAVES, dtype=np.int32)
        self.L.oracle_get_keypoints(self.h, kps.ctypes.data, counts.ctypes.data)
        return kps, counts

    def describe(self, keypoints=None, counts=None):
        if keypoints is None:
            n = self.L.oracle_describe(self.h)
        else:
            keypoints = np.ascontiguousarray(keypoints, dtype=_abi.KEYPOINT_DTYPE)
            counts = np.ascontiguousarray(counts, dtype=np.int32)
            n = self.L.oracle_describe_keypoints(self.h, keypoints.ctypes.data, counts.ctypes.data)
        if n < 0:
            raise RuntimeError("describe before detect")
        d = np.zeros(n, dtype=_abi.DESCRIPTOR_DTYPE)
        dc = np.zeros(_abi.NUM_OCTAVES, dtype=np.int32)
        self.L.oracle_get_descriptors(self.h, d.ctypes.data, dc.ctypes.data)
        return d, dc

    def candidates(self, octave):
        n = self.L.oracle_get_candidates(self.h, octave, None, 0)
        out = np.zeros((n, 3), dtype=np.int32)
        if n:
            self.L.oracle_get_candidates(self.h, octave, out.ctypes.data, n)
        return out

    def stats(self):
        s = np.zeros((_abi.NUM_OCTAVES, 6), dtype=np.int64)
        self.L.oracle_get_stats(self.h, s.ctypes.data)
        return s  # columns: raw25, raw26, soft, interp, contrast, final

    def plane(self, what, octave=0, slice=0):
        if what == _abi.PLANE_GRAY:
            shape = (self.height, self.width)
        elif what == _abi.PLANE_SEED:
            shape = (self.info.octave_height[0], self.info.octave_width[0])
        elif what == _abi.PLANE_GRADIENT:
            shape = (self.info.octave_height[octave], self.info.octave_width[octave], 2)
        else:
            shape = (self.info.octave_height[octave], self.info.octave_width[octave])
        out = np.zeros(shape, dtype=np.float32)
        n = self.L.oracle_get_plane(self.h, what, octave, slice, out.ctypes.data)
        assert n == out.size, (n, out.size)
        return out


def oracle_math(op, a, b=None):
    a = np.ascontiguousarray(a, dtype=np.float32)
    b = np.ascontiguousarray(b if b is not None else np.zeros_like(a), dtype=np.float32)
    out = np.zeros_like(a)
    lib().oracle_math(op, a.ctypes.data, b.ctypes.data, out.ctypes.data, a.size)
    return out


def oracle_match(source, target, absolute_threshold=300.0, relative_threshold=0.6):
    """SIFTDescriptor.match (SIFTDescriptor.swift:298-361) on [n, 128] uint8 feature matrices."""
    a = np.ascontiguousarray(source, dtype=np.uint8).reshape(-1, 128)
    b = np.ascontiguousarray(target, dtype=np.uint8).reshape(-1, 128)
    out = np.zeros(max(len(a), 1), dtype=_abi.MATCH_DTYPE)
    n = lib().oracle_match(a.ctypes.data, len(a), b.ctypes.data, len(b), absolute_threshold, relative_threshold,
                           out.ctypes.data)
    return out[:n].copy()


def oracle_approximate_match(source, target, absolute_threshold=300.0, relative_threshold=0.6):
    """SIFTDescriptor.approximateMatch over the reference's trie (SIFTDescriptor.swift:362-417, Trie.swift)."""
    a = np.ascontiguousarray(source, dtype=np.uint8).reshape(-1, 128)
    b = np.ascontiguousarray(target, dtype=np.uint8).reshape(-1, 128)
    out = np.zeros(max(len(a), 1), dtype=_abi.MATCH_DTYPE)
    n = lib().oracle_approximate_match(a.ctypes.data, len(a), b.ctypes.data, len(b), absolute_threshold,
                                       relative_threshold, out.ctypes.data)
    return out[:n].copy()


def oracle_match_geometry(source, source_xy, target, target_xy, absolute_threshold=1.176, relative_threshold=0.6):
    """SIFTDescriptor.matchGeometry (SIFTDescriptor.swift:104-296)."""
    a = np.ascontiguousarray(source, dtype=np.uint8).reshape(-1, 128)
    b = np.ascontiguousarray(target, dtype=np.uint8).reshape(-1, 128)
    axy = np.ascontiguousarray(source_xy, dtype=np.float32).reshape(-1, 2)
    bxy = np.ascontiguousarray(target_xy, dtype=np.float32).reshape(-1, 2)
    return float(lib().oracle_match_geometry(a.ctypes.data, axy.ctypes.data, len(a), b.ctypes.data, bxy.ctypes.data,
                                             len(b), absolute_threshold, relative_threshold))
