"""GPU tests of the boundary beyond the two reference calls: ingestion formats, the pipelined
submit / wait form, CUDA-graph replay, the column wire format and its lazy views, the staged path."""
import ctypes as C

import numpy as np
import pytest

from siftmetal_b200 import _abi

pytestmark = pytest.mark.gpu


def _same(a, b):
    return (np.array_equal(a.keypoints, b.keypoints) and np.array_equal(a.descriptors, b.descriptors) and
            np.array_equal(a.keypoint_counts, b.keypoint_counts) and
            np.array_equal(a.descriptor_counts, b.descriptor_counts) and
            np.array_equal(a.candidate_counts, b.candidate_counts))


@pytest.mark.parametrize("w,h", [(640, 480), (333, 77), (1367, 911)])
def test_gray8_and_nv12_equal_bgra_on_the_gray_expanded_frame(w, h):
    """SIFT_INPUT_GRAY8 / NV12: a gray byte is converted exactly like the BGRA pixel (v, v, v), so
    every plane and every result equals the reference path on the expanded frame (incl. odd widths)."""
    from siftmetal_b200 import Engine
    from siftmetal_b200.synth import pink_noise_gray

    g = pink_noise_gray(w, h, 7)
    bgra = np.repeat(g[:, :, None], 4, axis=2).copy()
    bgra[..., 3] = 255
    e4 = Engine(w, h)
    r4 = e4.detect_and_describe([bgra])
    for fmt in (_abi.INPUT_GRAY8, _abi.INPUT_NV12):
        e1 = Engine(w, h, input_format=fmt)
        if fmt == _abi.INPUT_NV12:
            # Y plane followed by the interleaved CbCr plane (never read): pass the luma rows
            nv12 = np.concatenate([g, np.full((h // 2 + 1, w), 128, np.uint8)], axis=0)
            frame = nv12[:h]
        else:
            frame = g
        r1 = e1.detect_and_describe([frame])
        assert np.array_equal(e1.plane(_abi.PLANE_GRAY), e4.plane(_abi.PLANE_GRAY))
        assert np.array_equal(e1.plane(_abi.PLANE_SEED), e4.plane(_abi.PLANE_SEED))
        assert _same(r1, r4)
        e1.close()
    e4.close()


def test_pipelined_submit_wait_equals_synchronous_calls():
    """Two calls in flight (upload of call i + 1 under the kernels of call i, results written by
    the kernels into each slot's pinned columns): same results, in submission order."""
    from siftmetal_b200 import Engine, SiftError
    from siftmetal_b200.synth import pink_noise_bgra

    w, h, n = 640, 480, 3
    batches = [[pink_noise_bgra(w, h, 10 * b + i) for i in range(n)] for b in range(5)]
    eng = Engine(w, h, max_batch=n)
    ref = [eng.detect_and_describe(b) for b in batches]
    got = []
    eng.submit(batches[0])
    for b in batches[1:]:
        eng.submit(b)
        assert eng.pending() == 2
        with pytest.raises(SiftError) as ei:
            eng.submit(b)                          # both slots taken
        assert ei.value.status == _abi.SIFT_ERR_BUSY
        got.append(eng.wait())
    got.append(eng.wait())
    assert eng.pending() == 0
    for a, b in zip(ref, got):
        assert _same(a, b)
    # zero-copy views stay valid until the slot is reused: results of call i survive submit(i + 1)
    eng.submit(batches[0])
    v0 = eng.wait(copy=False)
    eng.submit(batches[1])
    assert np.array_equal(v0.keypoints, ref[0].keypoints) and np.array_equal(v0.descriptors, ref[0].descriptors)
    eng.wait()
    eng.close()


def test_graph_replay_equals_launch_by_launch():
    """First call of a shape runs launch by launch, the second records the CUDA graph, later ones
    replay it: identical results; per-stage timing (opt-in) forces the eager path."""
    from siftmetal_b200 import Engine
    from siftmetal_b200.synth import pink_noise_bgra

    w, h = 960, 540
    imgs = [pink_noise_bgra(w, h, i) for i in range(2)]
    eng = Engine(w, h)
    base = [eng.detect_and_describe([imgs[i]]) for i in range(2)]
    assert not eng.timings()["graph_replay"]              # opt-in
    eng.set_graph_replay(True)
    r = [eng.detect_and_describe([imgs[i % 2]]) for i in range(6)]
    assert _same(r[0], base[0]) and _same(r[1], base[1])
    t = eng.timings()
    assert t["graph_replay"] and not t["stage_timing_enabled"] and t["total_ms"] > 0
    assert t["kernel_launches"] > 40
    assert _same(r[0], r[2]) and _same(r[2], r[4]) and _same(r[1], r[3]) and _same(r[3], r[5])
    assert not _same(r[0], r[1])
    eng.set_stage_timing(True)
    rt = eng.detect_and_describe([imgs[0]])
    t = eng.timings()
    assert t["stage_timing_enabled"] and not t["graph_replay"]
    assert all(t[k + "_ms"] > 0 for k in _abi.STAGE_NAMES) and t["blur_octave0_ms"] > 0
    assert abs(sum(t[k + "_ms"] for k in _abi.STAGE_NAMES) - t["total_ms"]) < 0.05 * t["total_ms"] + 0.02
    assert _same(rt, r[0])
    eng.set_stage_timing(False)
    assert _same(eng.detect_and_describe([imgs[0]]), r[0])
    assert eng.timings()["graph_replay"]
    eng.close()


def test_column_wire_format_and_lazy_views(butterfly_bgra):
    """SiftBatchResult columns (26 B per keypoint, dense [n, 128] features) against the record
    API of the reference-shaped calls; lazy views build objects only on access; the C
    materialisers give the same records."""
    from siftmetal_b200 import Engine, SIFT, IntegralSize, SIFTKeypoint, SIFTDescriptor

    h, w = butterfly_bgra.shape[:2]
    eng = Engine(w, h)
    res = eng.detect_and_describe([butterfly_bgra])
    kc, dc = res.keypoint_columns, res.descriptor_columns
    assert kc.scaled_xy.dtype == np.int16 and kc.octave_scale.dtype == np.uint8
    assert dc.features.shape == (len(dc), 128) and dc.features.dtype == np.uint8 and dc.features.flags.c_contiguous
    per_kp = sum(a.itemsize * (a.shape[1] if a.ndim == 2 else 1) for a in
                 (kc.absolute_x, kc.absolute_y, kc.sigma, kc.value, kc.sub_scale, kc.scaled_xy, kc.octave_scale))
    assert per_kp == 26
    # records from the two-step API == records materialised from the columns
    kps, counts = eng.detect(butterfly_bgra)
    assert np.array_equal(kps, res.keypoints) and np.array_equal(counts, res.keypoint_counts[0])
    desc, dcounts = eng.describe(kps, counts)
    assert np.array_equal(desc, res.descriptors) and np.array_equal(dcounts, res.descriptor_counts[0])
    # lazy element access
    kv, dv = res.frame_view(0)
    k5 = kv[5]
    assert isinstance(k5, SIFTKeypoint) and k5.scaledCoordinate == (int(kps["scaledX"][5]), int(kps["scaledY"][5]))
    assert np.float32(k5.normalizedCoordinate[0]) == kps["normalizedX"][5]
    d7 = dv[7]
    assert isinstance(d7, SIFTDescriptor) and d7.features.components == desc["features"][7].tolist()
    assert d7.keypoint.sigma == float(kps["sigma"][desc["keypoint"][7]])
    # C materialisers
    r = _abi.SiftBatchResult()
    assert eng.L.sift_detect_and_describe_batch(
        eng.ctx, (C.c_void_p * 1)(butterfly_bgra.ctypes.data), 1, w * 4, C.byref(r)) == 0
    out_k = np.zeros(10, _abi.KEYPOINT_DTYPE)
    assert eng.L.sift_materialize_keypoints(eng.ctx, C.byref(r), 20, 10, out_k.ctypes.data) == 0
    assert np.array_equal(out_k, kps[20:30])
    out_d = np.zeros(4, _abi.DESCRIPTOR_DTYPE)
    assert eng.L.sift_materialize_descriptors(C.byref(r), 100, 4, out_d.ctypes.data) == 0
    assert np.array_equal(out_d, desc[100:104])
    assert eng.L.sift_materialize_descriptors(C.byref(r), r.total_descriptors - 1, 2, out_d.ctypes.data) == \
        _abi.SIFT_ERR_INVALID_ARGUMENT
    eng.close()
    # the reference-shaped object API hands out lazy per-octave lists
    sift = SIFT(device=0, configuration=SIFT.Configuration(inputSize=IntegralSize(w, h)))
    octs = sift.getKeypoints(butterfly_bgra)
    dl = sift.getDescriptors(octs)
    assert [len(o) for o in dl] == list(dcounts)
    first = dl[0][0]
    assert isinstance(first, SIFTDescriptor) and first.features.components == desc["features"][0].tolist()
    assert first.keypoint is octs[0][int(desc["keypoint"][0])] or first.keypoint == octs[0][int(desc["keypoint"][0])]
    sift.close()


def test_staged_path_and_partial_batches():
    """upload / set_device_input → execute → download equals the host-buffer call; a context
    accepts any n <= max_batch."""
    import torch

    from siftmetal_b200 import Engine
    from siftmetal_b200.synth import pink_noise_bgra

    w, h, n = 320, 240, 4
    frames = [pink_noise_bgra(w, h, 40 + i) for i in range(n)]
    eng = Engine(w, h, max_batch=n)
    ref = eng.detect_and_describe(frames)
    eng.upload(frames)
    frames_copy = [f.copy() for f in frames]
    for f in frames:
        f[:] = 0                      # the copies completed inside upload(): the caller may reuse its buffers
    eng.execute()
    assert _same(eng.download(), ref)
    dev = torch.from_numpy(np.stack(frames_copy)).cuda()
    eng.set_device_input(dev.data_ptr(), n, w * 4, w * h * 4)
    eng.execute()
    eng.execute()                     # second execute of the same input: graph capture + replay
    assert _same(eng.download(), ref)
    eng.set_device_input(dev.data_ptr() + w * h * 4, 2, w * 4, w * h * 4)
    eng.execute()
    part = eng.download()
    k1, d1 = ref.frame(1)
    k2, d2 = part.frame(0)
    assert np.array_equal(k1, k2) and np.array_equal(d1, d2) and part.keypoint_counts.shape == (2, 7)
    eng.close()
