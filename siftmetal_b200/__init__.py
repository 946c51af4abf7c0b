"""siftmetal_b200 — B200-native (sm_100a) SIFT detect + describe behind the reference's API.

The product is libsiftcuda.so (hand-written CUDA, C ABI in include/siftcuda.h); this package is
its Python host binding, mirroring lukevanin/SIFTMetal's public Swift surface. Importing the
package does not load the library; constructing `SIFT` / `Engine` does, and raises if it is
missing or no sm_100 device is usable — there is no CPU fallback.
"""
from . import _abi  # noqa: F401
from .api import (  # noqa: F401
    BatchResult,
    DescriptorColumns,
    Engine,
    KeypointColumns,
    LazyDescriptorList,
    LazyKeypointList,
    IntegralSize,
    IntVector,
    SIFT,
    SIFTDescriptor,
    SIFTKeypoint,
    SiftError,
    device_math,
    load_library,
)

__all__ = [
    "SIFT", "SIFTKeypoint", "SIFTDescriptor", "IntVector", "IntegralSize", "Engine", "BatchResult",
    "KeypointColumns", "DescriptorColumns", "LazyKeypointList", "LazyDescriptorList",
    "SiftError", "load_library", "device_math",
]
