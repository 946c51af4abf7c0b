//  SIFT.swift — drop-in for lukevanin/SIFTMetal's `SIFT` class on NVIDIA B200.
//
//  UNVERIFIED (no Swift toolchain in the build environment). Logic-free marshalling over the C ABI
//  of include/siftcuda.h; every numeric step runs in libsiftcuda.so.
//
//  Reference surface kept (Sources/SIFTMetal/SIFT/SIFT.swift):
//    SIFT.Configuration(inputSize:)                         :57-103
//    SIFT(device:configuration:)                            :112-143   device: MTLDevice -> CUDA ordinal
//    getKeypoints(_:) -> [[SIFTKeypoint]]                   :147-152   MTLTexture -> BGRA8 pixel buffer
//    getDescriptors(keypointOctaves:) -> [[SIFTDescriptor]] :207-238
import CSIFTCUDA

public struct IntegralSize {  // Utilities/Math.swift:11-19
    public var width: Int
    public var height: Int
    public init(width: Int, height: Int) {
        self.width = width
        self.height = height
    }
}

public struct SIFTKeypoint {  // SIFTKeypoint.swift:11-57
    public var octave: Int
    public var scale: Int
    public var subScale: Float
    public var scaledCoordinate: SIMD2<Int>
    public var absoluteCoordinate: SIMD2<Float>
    public var normalizedCoordinate: SIMD2<Float>
    public var sigma: Float
    public var value: Float

    init(_ p: SiftKeypoint) {
        octave = Int(p.octave)
        scale = Int(p.scale)
        subScale = p.subScale
        scaledCoordinate = SIMD2<Int>(Int(p.scaledX), Int(p.scaledY))
        absoluteCoordinate = SIMD2<Float>(p.absoluteX, p.absoluteY)
        normalizedCoordinate = SIMD2<Float>(p.normalizedX, p.normalizedY)
        sigma = p.sigma
        value = p.value
    }

    var pod: SiftKeypoint {
        SiftKeypoint(
            octave: Int32(octave), scale: Int32(scale), subScale: subScale,
            scaledX: Int32(scaledCoordinate.x), scaledY: Int32(scaledCoordinate.y),
            absoluteX: absoluteCoordinate.x, absoluteY: absoluteCoordinate.y,
            normalizedX: normalizedCoordinate.x, normalizedY: normalizedCoordinate.y,
            sigma: sigma, value: value)
    }
}

public struct IntVector: Equatable {  // Utilities/Vector.swift:12-60
    public let count: Int
    public private(set) var components: [Int]
    public init(_ components: [Int]) {
        precondition(!components.isEmpty)
        self.components = components
        self.count = components.count
    }
    public subscript(index: Int) -> Int { components[index] }
}

public struct SIFTDescriptor {  // SIFTDescriptor.swift:12-35 (index keys of :43-89 belong to the matcher)
    public let keypoint: SIFTKeypoint
    public let theta: Float
    public let features: IntVector
    public var rawFeatures: [Float] { features.components.map { Float($0) / Float(255) } }
}

public final class SIFT {

    public struct Configuration {  // only inputSize is settable, as in the reference
        var inputSize: IntegralSize
        public init(inputSize: IntegralSize) {
            self.inputSize = inputSize
        }
    }

    let configuration: Configuration
    private var context: OpaquePointer?

    public init(device: Int32, configuration: Configuration) {
        self.configuration = configuration
        var cfg = SiftConfig()
        sift_config_default(&cfg, Int32(configuration.inputSize.width), Int32(configuration.inputSize.height))
        let status = sift_create(&cfg, device, &context)
        // the reference aborts on any set-up failure (try!, fatalError); keep the non-throwing init
        precondition(status == SIFT_OK, String(cString: sift_status_string(status)))
    }

    deinit {
        sift_destroy(context)
    }

    /// `pixels`: `inputSize.height` rows of `bytesPerRow` bytes of BGRA8 (was: a bgra8Unorm MTLTexture).
    public func getKeypoints(_ pixels: UnsafeRawPointer, bytesPerRow: Int) -> [[SIFTKeypoint]] {
        var out: UnsafePointer<SiftKeypoint>?
        var counts = [Int32](repeating: 0, count: Int(SIFT_NUM_OCTAVES))
        let status = sift_detect(context, pixels, Int32(bytesPerRow), &out, &counts)
        precondition(status == SIFT_OK, String(cString: sift_last_error_string(context)))
        var result = [[SIFTKeypoint]]()
        var k = 0
        for o in 0 ..< Int(SIFT_NUM_OCTAVES) {
            var octave = [SIFTKeypoint]()
            octave.reserveCapacity(Int(counts[o]))
            for _ in 0 ..< Int(counts[o]) {
                octave.append(SIFTKeypoint(out![k]))
                k += 1
            }
            result.append(octave)
        }
        return result
    }

    public func getDescriptors(keypointOctaves: [[SIFTKeypoint]]) -> [[SIFTDescriptor]] {
        precondition(keypointOctaves.count == Int(SIFT_NUM_OCTAVES))  // SIFT.swift:208
        let flat = keypointOctaves.joined().map { $0.pod }
        let owners = Array(keypointOctaves.joined())
        var counts = keypointOctaves.map { Int32($0.count) }
        var out: UnsafePointer<SiftDescriptor>?
        var dcounts = [Int32](repeating: 0, count: Int(SIFT_NUM_OCTAVES))
        let status = flat.withUnsafeBufferPointer {
            sift_describe(context, $0.baseAddress, &counts, &out, &dcounts)
        }
        precondition(status == SIFT_OK, String(cString: sift_last_error_string(context)))
        var result = [[SIFTDescriptor]]()
        var d = 0
        for o in 0 ..< Int(SIFT_NUM_OCTAVES) {
            var octave = [SIFTDescriptor]()
            for _ in 0 ..< Int(dcounts[o]) {
                let r = out![d]
                let features = withUnsafeBytes(of: r.features) { $0.map { Int($0) } }
                octave.append(SIFTDescriptor(keypoint: owners[Int(r.keypoint)], theta: r.theta, features: IntVector(features)))
                d += 1
            }
            result.append(octave)
        }
        return result
    }
}
