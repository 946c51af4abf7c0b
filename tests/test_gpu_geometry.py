"""Host stages downstream of the matcher against the oracle's restatement:
SIFTDescriptor.matchGeometry (SIFTDescriptor.swift:104-296) and approximateMatch over the trie
(SIFTDescriptor.swift:362-417, Utilities/Trie.swift)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def scene():
    from siftmetal_b200 import Engine
    from siftmetal_b200.synth import pink_noise_bgra

    w, h = 640, 480
    f0 = pink_noise_bgra(w, h, 3)
    frames = [f0, np.roll(f0, (3, 5), axis=(0, 1)), pink_noise_bgra(w, h, 4)]
    e = Engine(w, h, max_batch=3)
    res = e.detect_and_describe(frames)
    out = []
    for f in range(3):
        kv, dv = res.frame_view(f)
        xy = np.stack([kv.absolute_x[dv.keypoint], kv.absolute_y[dv.keypoint]], 1).astype(np.float32)
        out.append((dv.features.copy(), xy))
    yield e, out
    e.close()


def test_approximate_match_equals_the_reference_trie(scene):
    from oracle_lib import oracle_approximate_match, oracle_match

    e, frames = scene
    rng = np.random.default_rng(5)
    cases = [(frames[0][0], frames[1][0]), (frames[1][0], frames[0][0]), (frames[0][0], frames[2][0]),
             (frames[0][0][:50], frames[1][0][:7]),                     # fewer leaves than 2 * radius + 1: the ring wraps
             (rng.integers(0, 256, (200, 128), dtype=np.uint8), rng.integers(0, 256, (300, 128), dtype=np.uint8)),
             (frames[0][0][:10], frames[0][0][:1])]                     # a single target: the queue never holds two
    for a, b in cases:
        g = e.approximate_match(a, b)
        o = oracle_approximate_match(a, b)
        assert len(g) == len(o)
        assert np.array_equal(g, o)
    # on the shifted frame the trie finds most of what brute force finds
    g = e.approximate_match(frames[0][0], frames[1][0])
    exact = oracle_match(frames[0][0], frames[1][0])
    assert len(g) > 0.5 * len(exact)
    assert len(e.approximate_match(frames[0][0], np.zeros((0, 128), np.uint8))) == 0


def test_match_geometry_score(scene):
    from oracle_lib import oracle_match_geometry

    e, frames = scene
    (fa, xa), (fb, xb), (fc, xc) = frames
    same = e.match_geometry(fa, xa, fb, xb, 300.0, 0.6)
    assert same == np.float32(oracle_match_geometry(fa, xa, fb, xb, 300.0, 0.6))
    assert same > 0.9                                                   # a pure translation keeps every length and angle
    other = e.match_geometry(fa, xa, fc, xc, 300.0, 0.6)
    assert other == np.float32(oracle_match_geometry(fa, xa, fc, xc, 300.0, 0.6))
    assert other < same
    # the reference's default absolute threshold (1.176) on feature distances over f / 255
    d = e.match_geometry(fa, xa, fb, xb)
    assert d == np.float32(oracle_match_geometry(fa, xa, fb, xb))
    # fewer than 7 matches: 0 (SIFTDescriptor.swift:129-132)
    assert e.match_geometry(fa[:5], xa[:5], fb, xb, 300.0, 0.6) == 0.0
    # a scaled + rotated copy of the coordinates keeps the score (ratios and angles are invariant)
    th = 0.3
    R = np.array([[np.cos(th), -np.sin(th)], [np.sin(th), np.cos(th)]], np.float32) * 1.7
    assert abs(e.match_geometry(fa, xa, fb, (xb @ R.T).astype(np.float32), 300.0, 0.6) - same) < 1e-3
