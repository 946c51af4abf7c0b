"""Host logic of the column wire format (no GPU): lazy views over the result columns build the same
records / objects as the reference's eager SIFTKeypoint / SIFTDescriptor lists
(SIFTOctave.swift:257-286, SIFTDescriptor.swift:36-89), frames are cut by the per-octave counts."""
import numpy as np
import pytest

from siftmetal_b200._abi import DESCRIPTOR_DTYPE, KEYPOINT_DTYPE
from siftmetal_b200.api import (BatchResult, DescriptorColumns, KeypointColumns, LazyDescriptorList,
                                LazyKeypointList, SIFTKeypoint, _Concat)


def _columns(n, nd, seed=0):
    rng = np.random.default_rng(seed)
    sizes = np.array([[3840, 2160], [1920, 1080], [960, 540], [480, 270], [240, 135], [120, 67], [60, 33]], np.float32)
    octave = np.sort(rng.integers(0, 7, n)).astype(np.uint8)
    scale = rng.integers(1, 4, n).astype(np.uint8)
    xy = np.stack([rng.integers(1, 60, n), rng.integers(1, 33, n)], 1).astype(np.int16)
    kc = KeypointColumns(rng.random(n, np.float32) * 1920, rng.random(n, np.float32) * 1080,
                         rng.random(n, np.float32) * 8, rng.standard_normal(n).astype(np.float32),
                         rng.random(n, np.float32) - 0.5, xy, np.stack([octave, scale], 1), sizes)
    owner = np.sort(rng.integers(0, max(n, 1), nd)).astype(np.int32)
    dc = DescriptorColumns(rng.integers(0, 256, (nd, 128)).astype(np.uint8),
                           (rng.random(nd, np.float32) * 6.28).astype(np.float32), owner, kc)
    return kc, dc, sizes


def test_keypoint_records_and_objects_agree():
    kc, _, sizes = _columns(50, 0)
    r = kc.records()
    assert r.dtype == KEYPOINT_DTYPE and len(r) == len(kc) == 50
    for i in (0, 7, 49):
        k = kc[i]
        assert isinstance(k, SIFTKeypoint)
        assert (k.octave, k.scale) == (int(r["octave"][i]), int(r["scale"][i]))
        assert k.scaledCoordinate == (int(r["scaledX"][i]), int(r["scaledY"][i]))
        assert k.absoluteCoordinate == (float(r["absoluteX"][i]), float(r["absoluteY"][i]))
        # normalizedCoordinate = scaled / octave size, in float32 (SIFTOctave.swift:278-281)
        w, h = sizes[k.octave]
        assert k.normalizedCoordinate == (float(np.float32(r["scaledX"][i]) / w), float(np.float32(r["scaledY"][i]) / h))
        assert np.float32(k.normalizedCoordinate[0]) == r["normalizedX"][i]
        assert (k.sigma, k.value, k.subScale) == (float(r["sigma"][i]), float(r["value"][i]), float(r["subScale"][i]))
    part = kc[10:20]
    assert len(part) == 10 and np.array_equal(part.records(), r[10:20])
    with pytest.raises(IndexError):
        kc[0:10:2]
    assert len(KeypointColumns(*[a[:0] for a in (kc.absolute_x, kc.absolute_y, kc.sigma, kc.value, kc.sub_scale,
                                                  kc.scaled_xy, kc.octave_scale)], sizes).records()) == 0


def test_descriptor_views_refer_to_their_keypoints():
    kc, dc, _ = _columns(40, 55, seed=1)
    r = dc.records()
    assert r.dtype == DESCRIPTOR_DTYPE and len(r) == 55
    for i in (0, 13, 54):
        d = dc[i]
        owner = kc[int(r["keypoint"][i])]
        assert d.keypoint.absoluteCoordinate == owner.absoluteCoordinate and d.keypoint.sigma == owner.sigma
        assert d.theta == float(r["theta"][i])
        assert list(d.features.components) == r["features"][i].tolist()
    c = dc.copy()
    c.features[0, 0] ^= 0xFF
    assert dc.features[0, 0] != c.features[0, 0]          # a copy, not a view


def test_batch_result_cuts_frames_by_counts():
    kc, dc, _ = _columns(30, 36, seed=2)
    kcounts = np.array([[4, 3, 2, 1, 0, 0, 0], [0, 0, 0, 0, 0, 0, 0], [9, 6, 3, 1, 1, 0, 0]], np.int32)
    # descriptors per frame: frame 0 owns keypoints 0..9, frame 2 keypoints 0..19 (indices are per frame)
    dcounts = np.array([[5, 4, 2, 1, 0, 0, 0], [0, 0, 0, 0, 0, 0, 0], [10, 8, 4, 1, 1, 0, 0]], np.int32)
    assert kcounts.sum() == 30 and dcounts.sum() == 36
    res = BatchResult(kc, dc, kcounts, dcounts, kcounts * 2)
    k0, d0 = res.frame(0)
    k1, d1 = res.frame(1)
    k2, d2 = res.frame(2)
    assert (len(k0), len(k1), len(k2)) == (10, 0, 20) and (len(d0), len(d1), len(d2)) == (12, 0, 24)
    assert np.array_equal(np.concatenate([k0, k1, k2]), res.keypoints)
    assert np.array_equal(np.concatenate([d0, d1, d2]), res.descriptors)
    kv, dv = res.frame_view(2)
    assert np.array_equal(kv.records(), k2) and np.array_equal(dv.records(), d2)
    assert res.keypoints is res.keypoints                  # built once


def test_lazy_lists_build_on_access_and_cache():
    kc, dc, _ = _columns(12, 9, seed=3)
    recs = kc.records()
    per_octave = [LazyKeypointList(recs[recs["octave"] == o]) for o in range(7)]
    flat = _Concat(per_octave)
    assert len(flat) == 12
    for i in range(12):
        assert flat[i].absoluteCoordinate == (float(recs["absoluteX"][i]), float(recs["absoluteY"][i]))
    one = per_octave[int(recs["octave"][0])]
    assert one[0] is one[0] and one[-1] is one[len(one) - 1]                     # cached objects
    drecs = dc.records()
    drecs["keypoint"] = np.minimum(drecs["keypoint"], 11)
    lst = LazyDescriptorList(drecs, flat)
    assert len(lst) == 9 and len(lst[2:5]) == 3
    d = lst[4]
    assert d.keypoint is flat[int(drecs["keypoint"][4])]
    assert np.array_equal(lst.thetas, drecs["theta"]) and lst.features.shape == (9, 128)
    assert [x.theta for x in lst] == [float(t) for t in drecs["theta"]]
