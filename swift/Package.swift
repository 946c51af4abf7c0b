// swift-tools-version: 5.7
// SIFTCUDA — Swift veneer over libsiftcuda.so (B200 / sm_100a), keeping the public API of
// lukevanin/SIFTMetal (Package.swift:1-84 of the reference declares targets MetalShaders + SIFTMetal;
// here the C module `MetalShaders` is replaced by the system-library target CSIFTCUDA).
//
// UNVERIFIED: neither the build container nor the GPU boxes have a Swift toolchain, so this
// package has never been compiled. It is kept logic-free on purpose; the same surface is tested
// through the C ABI from Python (siftmetal_b200/api.py) and C++ (include/SIFT.hpp).
//
// Build (on a Linux host with Swift and a B200):
//   swift build -Xcc -I../include -Xlinker -L../siftmetal_b200 -Xlinker -rpath -Xlinker $PWD/../siftmetal_b200
import PackageDescription

let package = Package(
    name: "SIFTCUDA",
    products: [
        .library(name: "SIFTCUDA", targets: ["SIFTCUDA"]),
    ],
    targets: [
        .systemLibrary(name: "CSIFTCUDA", path: "Sources/CSIFTCUDA"),
        .target(name: "SIFTCUDA", dependencies: ["CSIFTCUDA"], path: "Sources/SIFTCUDA"),
    ]
)
