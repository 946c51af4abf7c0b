"""CPU pins of the oracle's matching stages against independent numpy
restatements of the reference's loops and against the cases the reference's semantics decide:
SIFTDescriptor.match (SIFTDescriptor.swift:298-361), compareGeometry (:165-296), approximateMatch
over the trie (:362-417, Utilities/Trie.swift)."""
import numpy as np
import pytest

import oracle_lib as ol

F32 = np.float32
FMAX = F32(3.402823466e38)


def _numpy_match(a, b, abs_thr, rel_thr):
    """The reference's loop, literally: `second` only moves when `best` improves (:339-343)."""
    out = []
    for i in range(len(a)):
        if len(b) == 0:
            continue
        d2 = ((b.astype(np.int64) - a[i].astype(np.int64)) ** 2).sum(1)
        best = second = None
        idx = -1
        for j, d in enumerate(d2):
            if best is None or d < best:
                second, best, idx = best, int(d), j
        db = np.sqrt(F32(best)) / F32(255.0)
        ds = FMAX if second is None else np.sqrt(F32(second)) / F32(255.0)
        if db < F32(abs_thr) and db < ds * F32(rel_thr):
            out.append((i, idx, db))
    return out


@pytest.mark.parametrize("ns,nt,spread", [(1, 1, 255), (17, 33, 255), (64, 50, 12), (40, 200, 4), (5, 0, 255), (0, 5, 255)])
def test_match_equals_the_reference_loop(ns, nt, spread):
    rng = np.random.default_rng(ns * 1000 + nt)
    a = rng.integers(0, spread + 1, (ns, 128)).astype(np.uint8)
    b = rng.integers(0, spread + 1, (nt, 128)).astype(np.uint8)
    if ns and nt:
        b[rng.integers(0, nt, max(1, nt // 3))] = a[rng.integers(0, ns, max(1, nt // 3))]    # exact duplicates, ties
    for abs_thr, rel_thr in ((300.0, 0.6), (1.176, 0.6), (0.5, 0.9)):
        got = ol.oracle_match(a, b, abs_thr, rel_thr)
        want = _numpy_match(a, b, abs_thr, rel_thr)
        assert [(int(s), int(t)) for s, t in zip(got["source"], got["target"])] == [(s, t) for s, t, _ in want]
        assert np.array_equal(got["distance"], np.array([d for _, _, d in want], dtype=F32))


def test_second_is_the_best_before_the_last_improvement():
    base = np.full(128, 100, np.uint8)
    src = base[None, :].copy()
    far = base.copy(); far[:64] += 40
    best = base.copy(); best[0] += 1
    near = base.copy(); near[0] += 2
    near2 = base.copy(); near2[0] += 1; near2[1] += 1
    m = ol.oracle_match(src, np.stack([far, best, near]))
    assert len(m) == 1 and m["target"][0] == 1            # the closer runner-up comes too late to veto
    m = ol.oracle_match(src, np.stack([near, best, far]))
    assert len(m) == 1 and m["target"][0] == 1            # 1 < 0.6 * 2
    m = ol.oracle_match(src, np.stack([best, near, far]))
    assert len(m) == 1 and m["target"][0] == 0            # second stays greatestFiniteMagnitude
    assert len(ol.oracle_match(src, np.stack([near2, best, far]))) == 0      # 1 < 0.6 * sqrt(2) fails
    assert len(ol.oracle_match(src, np.stack([far, best]), 1.0 / 255.0, 0.6)) == 0   # absolute threshold is strict


def _numpy_compare_geometry(matches, sxy, txy, minimum=7):
    """compareGeometry (:165-296) in float32, one IEEE operation per step, in the oracle's order."""
    def length(x, y):
        return np.sqrt(F32(x * x) + F32(y * y))

    def clamp01(v):
        return F32(min(max(v, F32(0)), F32(1)))

    scores = []
    for i in range(len(matches) - 3):
        m0, m1, m2, m3 = matches[i:i + 4]
        sb = sxy[m1["source"]] - sxy[m0["source"]]
        tb = txy[m1["target"]] - txy[m0["target"]]
        sbl, tbl = length(*sb), length(*tb)
        if not (sbl >= 2) or not (tbl >= 2):
            continue
        st = sxy[m3["source"]] - sxy[m2["source"]]
        tt = txy[m3["target"]] - txy[m2["target"]]
        stl, ttl = length(*st), length(*tt)
        if not (stl >= 2) or not (ttl >= 2):
            continue
        sbn, tbn, stn, ttn = sb / sbl, tb / tbl, st / stl, tt / ttl
        sr, tr = stl / sbl, ttl / tbl
        sdot = clamp01(F32(F32(F32(stn[0] * sbn[0]) + F32(stn[1] * sbn[1])) * F32(0.5)) + F32(0.5))
        tdot = clamp01(F32(F32(F32(ttn[0] * tbn[0]) + F32(ttn[1] * tbn[1])) * F32(0.5)) + F32(0.5))
        ori = F32(1) - abs(F32(sdot - tdot))
        sca = clamp01(sr / tr) if sr < tr else clamp01(tr / sr)
        sim = F32(ori * sca)
        scores.append(F32(sim * sim))
    if len(scores) < minimum:
        return F32(0)
    total = F32(0)
    for s in scores:
        total = F32(total + s)
    mean = F32(total / F32(len(scores)))
    err = F32(0)
    for s in scores:
        d = F32(s - mean)
        err = F32(err + F32(d * d))
    sd = np.sqrt(F32(err / F32(len(scores) - 1)))
    fs = fc = F32(0)
    with np.errstate(invalid="ignore", divide="ignore"):
        for s in scores:
            if abs(F32(F32(s - mean) / sd)) <= 2:
                fs, fc = F32(fs + s), F32(fc + F32(1))
        return F32(fs / fc)


@pytest.mark.parametrize("n,noise", [(6, 0.0), (12, 3.0), (60, 8.0), (150, 25.0)])
def test_match_geometry_equals_the_numpy_restatement(n, noise):
    rng = np.random.default_rng(n)
    feats = rng.integers(0, 256, (n, 128)).astype(np.uint8)
    sxy = (rng.random((n, 2)) * 1000).astype(F32)
    perm = rng.permutation(n)
    c, s = np.cos(0.3), np.sin(0.3)
    txy = ((sxy @ np.array([[c, -s], [s, c]]).T) * 1.7 + 40 + rng.standard_normal((n, 2)) * noise).astype(F32)[perm]
    tfeats = feats[perm]
    got = ol.oracle_match_geometry(feats, sxy, tfeats, txy)
    m = ol.oracle_match(feats, tfeats, 1.176, 0.6)
    assert len(m) == n and np.array_equal(perm[m["target"]], np.arange(n))       # every descriptor finds itself
    want = F32(0) if len(m) < 7 else _numpy_compare_geometry(m[:80], sxy, txy)
    assert (np.isnan(got) and np.isnan(want)) or got == float(want), (got, want)
    if n >= 12 and noise > 0:
        assert 0.0 < got <= 1.0


def test_trie_keeps_a_match_only_with_two_improvements():
    """approximateMatch (:362-417): the FiniteQueue of two only admits a value that beats the current
    best (Trie.swift:287-300) and `guard matches.count == 2` (:394-396) drops a query whose queue saw
    a single improvement — e.g. an exact duplicate met first. Insertion order inside a leaf decides."""
    q = np.full((1, 128), 100, np.uint8)                 # every cell: S = 800, key digit 3
    t0 = q[0].copy(); t0[0] += 1                          # same key (S = 801), d2 = 1
    t1 = q[0].copy()                                      # the duplicate, d2 = 0
    m = ol.oracle_approximate_match(q, np.stack([t0, t1]))
    assert len(m) == 1 and m["target"][0] == 1 and m["distance"][0] == 0      # t0 enters, t1 improves: two entries
    assert len(ol.oracle_approximate_match(q, np.stack([t1, t0]))) == 0       # t1 first: t0 never enters
    # random descriptors sit in singleton leaves (2^16 keys of digits 3 / 4): a duplicate is met
    # first in its own leaf, nothing in the 20 neighbouring leaves beats distance 0 -> dropped
    rng = np.random.default_rng(7)
    target = rng.integers(0, 256, (300, 128)).astype(np.uint8)
    perm = rng.permutation(300)[:120]
    assert len(ol.oracle_approximate_match(target[perm], target)) == 0
    # against the exhaustive matcher the trie can only lose matches or settle for a farther neighbour
    query = np.clip(target[perm].astype(np.int16) + rng.integers(-40, 41, (120, 128)), 0, 255).astype(np.uint8)
    approx = ol.oracle_approximate_match(query, target, 300.0, 0.999)
    exact = ol.oracle_match(query, target, 300.0, 0.999)
    assert len(exact) > 0
    e = dict(zip(exact["source"].tolist(), exact["distance"].tolist()))
    for s_, d in zip(approx["source"].tolist(), approx["distance"].tolist()):
        if s_ in e:
            assert d >= e[s_]
    assert len(ol.oracle_approximate_match(query, np.zeros((0, 128), np.uint8))) == 0
