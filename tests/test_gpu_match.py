"""GPU parity of the device matcher (tcgen05 uint8 GEMM + fused ratio-test scan, csrc/match.cu)
against the oracle's restatement of SIFTDescriptor.match (SIFTDescriptor.swift:298-361): identical
correspondence index pairs, bit-identical distances."""
import os

import numpy as np
import pytest

from conftest import GOLDEN
from siftmetal_b200 import _abi

pytestmark = pytest.mark.gpu


def _features(rng, n, spread=255):
    return rng.integers(0, spread + 1, (n, 128), dtype=np.uint8)


def _check(eng, a, b, abs_thr=300.0, rel_thr=0.6):
    from oracle_lib import oracle_match

    g = eng.match(a, b, abs_thr, rel_thr)
    o = oracle_match(a, b, abs_thr, rel_thr)
    assert len(g) == len(o), (len(g), len(o))
    assert np.array_equal(g["source"], o["source"]) and np.array_equal(g["target"], o["target"])
    assert np.array_equal(g["distance"], o["distance"])
    return g


@pytest.fixture(scope="module")
def eng():
    from siftmetal_b200 import Engine

    e = Engine(64, 48)
    yield e
    e.close()


@pytest.mark.parametrize("ns,nt", [(1, 1), (1, 2), (3, 700), (300, 1000), (1000, 5), (129, 257), (128, 256),
                                   (2000, 3000), (5000, 9000)])
def test_random_feature_matrices(eng, ns, nt):
    """Ragged sizes around the 128-row / 256-column tiles; near-duplicate targets so that a fair
    share of the rows passes the ratio test, exact duplicates so that ties occur."""
    rng = np.random.default_rng(ns * 7919 + nt)
    b = _features(rng, nt)
    a = _features(rng, ns)
    k = min(ns, nt)
    pick = rng.permutation(nt)[:k]
    noise = rng.integers(-6, 7, (k, 128))
    a[:k] = np.clip(b[pick].astype(np.int32) + noise, 0, 255).astype(np.uint8)
    if nt >= 8:
        b[nt // 2] = b[1]            # exact duplicate rows: the first one must win (strict <)
        b[nt - 1] = b[1]
    g = _check(eng, a, b)
    if k >= 100:
        assert len(g) > 0.5 * k


def test_second_is_the_best_before_the_last_improvement(eng):
    """SIFTDescriptor.swift:339-343: `second` is only updated when `best` improves, so a closer
    runner-up that comes AFTER the best match does not veto it."""
    base = np.full(128, 100, np.uint8)
    src = base[None, :].copy()
    far = base.copy(); far[:64] += 40          # d2 = 64 * 1600
    best = base.copy(); best[0] += 1            # d2 = 1
    near = base.copy(); near[0] += 2            # d2 = 4: true second smallest
    # order far, best, near: second = far -> ratio test passes although near is very close
    g = _check(eng, src, np.stack([far, best, near]))
    assert len(g) == 1 and g["target"][0] == 1
    # order near, best, far: second = near -> 1 < 2 * 0.6 passes as well, 0.6 * sqrt(4) = 1.2
    g = _check(eng, src, np.stack([near, best, far]))
    assert len(g) == 1 and g["target"][0] == 1
    # best first: second stays .greatestFiniteMagnitude -> kept
    g = _check(eng, src, np.stack([best, near, far]))
    assert len(g) == 1 and g["target"][0] == 0
    # runner-up just inside the ratio: near2 with d2 = 2 -> sqrt(1) < 0.6 * sqrt(2) = 0.848 fails
    near2 = base.copy(); near2[0] += 1; near2[1] += 1
    g = _check(eng, src, np.stack([near2, best, far]))
    assert len(g) == 0
    # absolute threshold (SIFTDescriptor.match(source:target:) defaults to 1.176)
    g = _check(eng, src, np.stack([far, best]), abs_thr=1.0 / 255.0)
    assert len(g) == 0
    # empty target: `guard let bestMatch` fails
    assert len(eng.match(src, np.zeros((0, 128), np.uint8))) == 0
    assert len(eng.match(np.zeros((0, 128), np.uint8), src)) == 0


def test_butterfly_against_ipol_descriptors(butterfly_bgra):
    """DescriptorTests.swift:120-125: descriptors found on the fixture image matched against the
    IPOL reference descriptors with absoluteThreshold 300, relativeThreshold 0.6."""
    from oracle_lib import oracle_match
    from siftmetal_b200 import Engine

    ipol = np.loadtxt(os.path.join(GOLDEN, "butterfly-descriptors.txt"), usecols=range(132))
    target = ipol[:, 4:132].astype(np.uint8)
    h, w = butterfly_bgra.shape[:2]
    e = Engine(w, h)
    res = e.detect_and_describe([butterfly_bgra])
    src = res.descriptor_columns.features
    g = _check(e, src, target)
    assert len(g) >= 0.60 * len(src)                      # SURVEY.md §8c band (measured 63.8 %)
    owner = res.keypoints[res.descriptors["keypoint"][g["source"]]]
    px = np.hypot(owner["absoluteX"] - ipol[g["target"], 1], owner["absoluteY"] - ipol[g["target"], 0])
    # the reference's `second` (best before the last improvement) is laxer than a true second-best
    # ratio test, so a few correspondences are wrong ones; the bulk lands on the same keypoint
    assert np.mean(px < 2.0) > 0.95
    o = oracle_match(src, target)
    assert np.array_equal(g, o)
    e.close()


def test_match_frames_uses_the_device_resident_columns():
    """sift_match_frames: two frames of the last batch matched from the device feature columns —
    same correspondences as matching their downloaded feature matrices."""
    from siftmetal_b200 import Engine
    from siftmetal_b200.synth import pink_noise_bgra

    w, h = 640, 480
    f0 = pink_noise_bgra(w, h, 3)
    f1 = np.roll(f0, (3, 5), axis=(0, 1))                 # the same scene shifted by (5, 3) pixels
    f2 = pink_noise_bgra(w, h, 4)
    e = Engine(w, h, max_batch=3)
    res = e.detect_and_describe([f0, f1, f2])
    feats = [res.frame_view(f)[1].features for f in range(3)]
    for a, b in ((0, 1), (1, 0), (0, 2), (2, 2)):
        g = e.match_frames(a, b)
        o = _check(e, feats[a], feats[b])
        assert np.array_equal(g, o)
    m = e.match_frames(0, 1)
    assert len(m) > 0.3 * len(feats[0])
    k0, _ = res.frame(0)
    k1, _ = res.frame(1)
    d0 = res.frame(0)[1]["keypoint"][m["source"]]
    d1 = res.frame(1)[1]["keypoint"][m["target"]]
    dx = k1["absoluteX"][d1] - k0["absoluteX"][d0]
    dy = k1["absoluteY"][d1] - k0["absoluteY"][d0]
    assert abs(np.median(dx) - 5) < 0.1 and abs(np.median(dy) - 3) < 0.1
    assert len(e.match_frames(0, 2)) < 0.05 * len(feats[0])
    e.close()
