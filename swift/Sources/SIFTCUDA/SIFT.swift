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

    // MARK: - Beyond the reference's two calls (ABI v2): batches, pipelining, lazy views, matching

    /// getKeypoints + getDescriptors for several frames in one call; results as lazy views over the
    /// column wire format (`SiftBatchResult`): a `SIFTKeypoint` / `SIFTDescriptor` is only built when
    /// an element is indexed (the reference pays SIFTDescriptor.init per descriptor, SIFTDescriptor.swift:36-89).
    public func detectAndDescribe(frames: [UnsafeRawPointer?], bytesPerRow: Int) -> [FrameResult] {
        var r = SiftBatchResult()
        let status = frames.withUnsafeBufferPointer {
            sift_detect_and_describe_batch(context, $0.baseAddress, Int32(frames.count), Int32(bytesPerRow), &r)
        }
        precondition(status == SIFT_OK, String(cString: sift_last_error_string(context)))
        return split(r)
    }

    /// Pipelined form (CoreVideoMetalCache.swift:23-31 hands over one texture per camera frame): up to two
    /// calls in flight, `wait()` returns them in submission order. The frames must stay valid until then.
    public func submit(frames: [UnsafeRawPointer?], bytesPerRow: Int) {
        let status = frames.withUnsafeBufferPointer {
            sift_submit(context, $0.baseAddress, Int32(frames.count), Int32(bytesPerRow))
        }
        precondition(status == SIFT_OK, String(cString: sift_last_error_string(context)))
    }

    public func wait() -> [FrameResult] {
        var r = SiftBatchResult()
        let status = sift_wait(context, &r)
        precondition(status == SIFT_OK, String(cString: sift_last_error_string(context)))
        return split(r)
    }

    /// SIFTDescriptor.match(source:target:absoluteThreshold:relativeThreshold:) (SIFTDescriptor.swift:298-361)
    /// on the dense feature matrices of two descriptor views; (source row, target row, featureDistance).
    public func match(source: DescriptorView, target: DescriptorView,
                      absoluteThreshold: Float = 300, relativeThreshold: Float = 0.6) -> [SiftMatch] {
        var out: UnsafePointer<SiftMatch>?
        var n: Int64 = 0
        let status = sift_match(context, source.features, Int64(source.count), target.features, Int64(target.count),
                                absoluteThreshold, relativeThreshold, &out, &n)
        precondition(status == SIFT_OK, String(cString: sift_last_error_string(context)))
        return Array(UnsafeBufferPointer(start: out, count: Int(n)))
    }

    private func split(_ r: SiftBatchResult) -> [FrameResult] {
        var out = [FrameResult]()
        var k: Int64 = 0
        var d: Int64 = 0
        for f in 0 ..< Int(r.n_frames) {
            var nk: Int64 = 0
            var nd: Int64 = 0
            for o in 0 ..< Int(SIFT_NUM_OCTAVES) {
                nk += Int64(r.keypoint_counts[f * Int(SIFT_NUM_OCTAVES) + o])
                nd += Int64(r.descriptor_counts[f * Int(SIFT_NUM_OCTAVES) + o])
            }
            let kv = KeypointView(context: context, result: r, first: k, count: Int(nk))
            out.append(FrameResult(keypoints: kv, descriptors: DescriptorView(result: r, first: d, count: Int(nd), keypoints: kv)))
            k += nk
            d += nd
        }
        return out
    }
}

/// Lazy random-access view of the keypoint columns of one frame.
public struct KeypointView: RandomAccessCollection {
    let context: OpaquePointer?
    let result: SiftBatchResult
    let first: Int64
    public let count: Int
    public var startIndex: Int { 0 }
    public var endIndex: Int { count }
    public subscript(index: Int) -> SIFTKeypoint {
        var p = SiftKeypoint()
        var r = result
        let status = sift_materialize_keypoints(context, &r, first + Int64(index), 1, &p)
        precondition(status == SIFT_OK)
        return SIFTKeypoint(p)
    }
}

/// Lazy random-access view of the descriptor columns of one frame.
public struct DescriptorView: RandomAccessCollection {
    let result: SiftBatchResult
    let first: Int64
    public let count: Int
    let keypoints: KeypointView
    public var startIndex: Int { 0 }
    public var endIndex: Int { count }
    /// dense [count][128] uint8 feature matrix
    public var features: UnsafePointer<UInt8>? { result.descriptors.features.map { $0 + Int(first) * 128 } }
    public subscript(index: Int) -> SIFTDescriptor {
        var d = SiftDescriptor()
        var r = result
        let status = sift_materialize_descriptors(&r, first + Int64(index), 1, &d)
        precondition(status == SIFT_OK)
        let features = withUnsafeBytes(of: d.features) { $0.map { Int($0) } }
        return SIFTDescriptor(keypoint: keypoints[Int(d.keypoint)], theta: d.theta, features: IntVector(features))
    }
}

public struct FrameResult {
    public let keypoints: KeypointView
    public let descriptors: DescriptorView
}

