"""GPU parity: libsiftcuda.so (through the C ABI) against the CPU oracle on the same inputs.

Bars (BASELINE.json north_star / SURVEY.md §8c):
  * pyramid planes, DoG, gradient field, candidate set (octave, scale, x, y), refined integer
    (scale, x, y): BIT-EXACT
  * refined position / sigma within 1e-3 px, orientation within 1e-3 rad
  * uint8 descriptor features within ±1
"""
import numpy as np
import pytest

from siftmetal_b200 import _abi

pytestmark = pytest.mark.gpu

POS_TOL = 1e-3    # px   (north_star)
THETA_TOL = 1e-3  # rad  (north_star)
FEAT_TOL = 1      # per uint8 element (north_star)


def _engine(w, h, **kw):
    from siftmetal_b200 import Engine

    return Engine(w, h, device=0, **kw)


def _oracle(w, h):
    from oracle_lib import Oracle

    return Oracle(w, h)


def _check_frame(eng, ora, bgra, kps, desc, kcounts, dcounts, ccounts=None, planes=True, frame=0,
                 max_structural=0.0):
    """Full comparison of one frame's GPU results with a fresh oracle run."""
    okps, ocounts = ora.detect(bgra)
    odesc, odcounts = ora.describe()
    info = ora.info
    if planes:
        assert np.array_equal(eng.plane(_abi.PLANE_GRAY, frame=frame), ora.plane(_abi.PLANE_GRAY))
        assert np.array_equal(eng.plane(_abi.PLANE_SEED, frame=frame), ora.plane(_abi.PLANE_SEED))
        for o in range(7):
            if info.octave_width[o] < 1 or info.octave_height[o] < 1:
                continue
            for s in range(6):
                assert np.array_equal(eng.plane(_abi.PLANE_GAUSSIAN, o, s, frame), ora.plane(_abi.PLANE_GAUSSIAN, o, s)), (o, s)
            for s in range(5):
                assert np.array_equal(eng.plane(_abi.PLANE_DOG, o, s, frame), ora.plane(_abi.PLANE_DOG, o, s)), (o, s)
            if info.octave_width[o] >= 3 and info.octave_height[o] >= 3:
                for s in (1, 2, 3):
                    g, r = eng.plane(_abi.PLANE_GRADIENT, o, s, frame), ora.plane(_abi.PLANE_GRADIENT, o, s)
                    assert np.array_equal(g, r), (o, s, np.abs(g - r).max())
    # candidate set: identical, in canonical order
    for o in range(7):
        oc = ora.candidates(o)
        gc = eng.candidates(o, frame)
        assert np.array_equal(gc, oc), (o, len(gc), len(oc))
        if ccounts is not None:
            assert ccounts[o] == len(oc)
    # keypoints
    assert np.array_equal(kcounts, ocounts), (kcounts, ocounts)
    for f in ("octave", "scale", "scaledX", "scaledY"):
        assert np.array_equal(kps[f], okps[f]), f
    for f in ("absoluteX", "absoluteY", "sigma"):
        assert np.all(np.abs(kps[f] - okps[f]) <= POS_TOL), f
    for f in ("subScale", "value", "normalizedX", "normalizedY"):
        assert np.allclose(kps[f], okps[f], rtol=0, atol=1e-6), f
    exact = all(np.array_equal(kps[f], okps[f]) for f in kps.dtype.names)
    # descriptors: same (keypoint, orientation) structure, θ and features within tolerance.
    # Structural mismatches (a keypoint with a missing / extra orientation: the 36-bin histogram is
    # summed in a different order on the GPU, so a peak test can flip in the last bit) must be 0 on
    # the fixture and <= 1e-4 of the keypoints on synthetic sets; the keypoints
    # concerned are left out of the element-wise comparison.
    nk = len(kps)
    gpk = np.bincount(desc["keypoint"], minlength=nk)
    opk = np.bincount(odesc["keypoint"], minlength=nk)
    bad = np.nonzero(gpk != opk)[0]
    assert len(bad) <= max_structural * nk, (len(bad), nk, dcounts, odcounts)
    if len(bad):
        desc = desc[~np.isin(desc["keypoint"], bad)]
        odesc = odesc[~np.isin(odesc["keypoint"], bad)]
    else:
        assert np.array_equal(dcounts, odcounts), (dcounts, odcounts)
    assert np.array_equal(desc["keypoint"], odesc["keypoint"])
    dth = np.abs(desc["theta"] - odesc["theta"])
    dth = np.minimum(dth, 2 * np.pi - dth)
    assert np.all(dth <= THETA_TOL), dth.max()
    df = np.abs(desc["features"].astype(np.int16) - odesc["features"].astype(np.int16))
    assert df.max(initial=0) <= FEAT_TOL, df.max()
    return {"keypoints_bit_exact": exact, "max_dtheta": float(dth.max(initial=0)), "structural": int(len(bad)),
            "feat_mismatch_frac": float((df > 0).mean()) if df.size else 0.0}


def test_device_math_bit_exact():
    """dev_math.cuh evaluates the same operation sequences as oracle/oracle_math.h."""
    from oracle_lib import oracle_math
    from siftmetal_b200 import device_math

    rng = np.random.default_rng(7)
    x = np.concatenate([np.linspace(-90, 1, 300001), -rng.random(100000) * 20]).astype(np.float32)
    assert np.array_equal(device_math(0, x), oracle_math(0, x))
    y = rng.standard_normal(400000).astype(np.float32) * 0.1
    z = rng.standard_normal(400000).astype(np.float32) * 0.1
    y[:500] = 0
    z[250:750] = 0
    assert np.array_equal(device_math(1, y, z), oracle_math(1, y, z))
    t = np.linspace(-7, 7, 300001).astype(np.float32)
    assert np.array_equal(device_math(2, t), oracle_math(2, t))
    assert np.array_equal(device_math(3, t), oracle_math(3, t))
    e = np.linspace(-2, 3, 100001).astype(np.float32)
    assert np.array_equal(device_math(4, e), oracle_math(4, e))


def test_schedule_matches_oracle():
    eng, ora = _engine(512, 340), _oracle(512, 340)
    gi, oi = eng.info, ora.info
    for name in ("octave_width", "octave_height", "octave_delta", "rho", "taps", "seed_weights"):
        assert list(getattr(gi, name)) == list(getattr(oi, name)), name
    assert gi.seed_sigma == oi.seed_sigma and gi.seed_taps == oi.seed_taps
    for s in range(5):
        assert list(gi.weights[s]) == list(oi.weights[s])
    for o in range(7):
        assert list(gi.sigmas[o]) == list(oi.sigmas[o])
    eng.close()


def test_butterfly_fixture_full_parity(butterfly_bgra):
    """config[0]: the reference's own fixture image, every stage against the oracle."""
    h, w = butterfly_bgra.shape[:2]
    eng, ora = _engine(w, h), _oracle(w, h)
    res = eng.detect_and_describe([butterfly_bgra])
    kps, desc = res.frame(0)
    rep = _check_frame(eng, ora, butterfly_bgra, kps, desc, res.keypoint_counts[0], res.descriptor_counts[0],
                       res.candidate_counts[0])
    assert abs(len(kps) - 1310) <= 10          # same band the oracle is pinned to
    assert rep["keypoints_bit_exact"]
    eng.close()


@pytest.mark.parametrize("w,h", [(203, 157), (64, 48), (640, 480), (40, 24), (333, 77), (1367, 911)])
def test_synthetic_sizes(w, h):
    """Ragged sizes: odd widths, octaves narrower than the 27-tap kernel, empty top octaves, and
    one large odd frame (row-banded octave 0, chunked upload, partial edge tiles in both axes)."""
    from siftmetal_b200.synth import pink_noise_bgra

    img = pink_noise_bgra(w, h, frame_index=w + h)
    eng, ora = _engine(w, h), _oracle(w, h)
    res = eng.detect_and_describe([img])
    kps, desc = res.frame(0)
    _check_frame(eng, ora, img, kps, desc, res.keypoint_counts[0], res.descriptor_counts[0], res.candidate_counts[0])
    eng.close()


def test_flat_and_saturated_images_yield_nothing():
    """Empty input case: no extrema above the contrast pre-threshold anywhere."""
    eng = _engine(96, 64)
    for v in (0, 255, 77):
        img = np.full((64, 96, 4), v, np.uint8)
        res = eng.detect_and_describe([img])
        assert res.keypoint_counts.sum() == 0 and res.descriptor_counts.sum() == 0
        assert len(res.keypoints) == 0 and len(res.descriptors) == 0
    eng.close()


def test_reference_api_two_step_flow(butterfly_bgra):
    """getKeypoints then getDescriptors(keypointOctaves:) as SIFT.swift:147-238, including a
    caller-filtered keypoint set (every third keypoint)."""
    from siftmetal_b200 import SIFT, IntegralSize

    h, w = butterfly_bgra.shape[:2]
    sift = SIFT(device=0, configuration=SIFT.Configuration(inputSize=IntegralSize(w, h)))
    octs = sift.getKeypoints(butterfly_bgra)
    assert len(octs) == 7
    ora = _oracle(w, h)
    okps, ocounts = ora.detect(butterfly_bgra)
    assert [len(o) for o in octs] == list(ocounts)
    # full set
    d_all = sift.getDescriptors(octs)
    odesc, odcounts = ora.describe()
    assert [len(o) for o in d_all] == list(odcounts)
    flat = [d for o in d_all for d in o]
    assert all(len(d.features.components) == 128 for d in flat)
    th = np.array([d.theta for d in flat], np.float32)
    assert np.all(np.abs(th - odesc["theta"]) <= THETA_TOL)
    # filtered subset through the engine (array form)
    eng = sift.engine
    kps, counts = eng.detect(butterfly_bgra)
    keep = np.arange(len(kps)) % 3 == 0
    starts = np.concatenate([[0], np.cumsum(counts)])
    sub_counts = np.array([keep[starts[i]:starts[i + 1]].sum() for i in range(7)], np.int32)
    gdesc, gdc = eng.describe(kps[keep], sub_counts)
    sdesc, sdc = ora.describe(okps[keep], sub_counts)
    assert np.array_equal(gdc, sdc)
    assert np.array_equal(gdesc["keypoint"], sdesc["keypoint"])
    assert np.all(np.abs(gdesc["theta"] - sdesc["theta"]) <= THETA_TOL)
    assert np.abs(gdesc["features"].astype(np.int16) - sdesc["features"].astype(np.int16)).max() <= FEAT_TOL
    sift.close()


def test_batch_matches_per_frame_oracle():
    """Frames of a batch are independent units: each equals its own oracle run."""
    from siftmetal_b200.synth import pink_noise_bgra

    w, h, n = 320, 240, 5
    frames = [pink_noise_bgra(w, h, i) for i in range(n)]
    eng, ora = _engine(w, h, max_batch=n), _oracle(w, h)
    res = eng.detect_and_describe(frames)
    assert res.keypoint_counts.shape == (n, 7)
    for f in range(n):
        kps, desc = res.frame(f)
        _check_frame(eng, ora, frames[f], kps, desc, res.keypoint_counts[f], res.descriptor_counts[f],
                     res.candidate_counts[f], planes=(f in (0, n - 1)), frame=f)
    # partial batch on the same context
    res2 = eng.detect_and_describe(frames[1:3])
    k1, d1 = res.frame(1)
    k2, d2 = res2.frame(0)
    assert np.array_equal(k1, k2) and np.array_equal(d1, d2)
    eng.close()


def test_error_behaviour(butterfly_bgra):
    from siftmetal_b200 import Engine, SiftError

    h, w = butterfly_bgra.shape[:2]
    eng = Engine(w, h)
    with pytest.raises(ValueError):
        eng.detect(butterfly_bgra[:, :-1])                      # size precondition
    with pytest.raises(SiftError) as ei:
        eng.describe(np.zeros(0, _abi.KEYPOINT_DTYPE), [0] * 7)   # describe before detect
    assert ei.value.status == _abi.SIFT_ERR_NOT_DETECTED
    eng.close()
    with pytest.raises(SiftError) as ei:
        Engine(4, 4)
    assert ei.value.status == _abi.SIFT_ERR_INVALID_ARGUMENT
    # configuration fields that would let the 3x3x3 stencils leave the plane, or make no sense
    for bad in (dict(image_border=0), dict(image_border=-3), dict(max_keypoints_per_frame=-1),
                dict(max_interpolation_iterations=-1), dict(dog_threshold=float("nan")),
                dict(edge_threshold=float("inf")), dict(input_format=9), dict(reserved=1)):
        with pytest.raises(SiftError) as ei:
            Engine(w, h, **bad)
        assert ei.value.status == _abi.SIFT_ERR_INVALID_ARGUMENT, bad
    eng = Engine(w, h)
    kps, counts = eng.detect(butterfly_bgra)
    with pytest.raises(ValueError):
        eng.describe(kps, [int(c) + 1 for c in counts])       # counts must add up to len(keypoints)
    with pytest.raises(SiftError) as ei:
        eng.wait()                                               # nothing submitted
    assert ei.value.status == _abi.SIFT_ERR_BUSY
    eng.close()
    # capacity overflow is reported, never silent (reference: precondition crash, Buffer.swift:35-39)
    small = Engine(w, h, max_keypoints_per_frame=100)
    res = small.detect_and_describe([butterfly_bgra], allow_capacity=True)
    assert res.status == _abi.SIFT_ERR_CAPACITY
    assert res.keypoint_counts.sum() <= 100
    small.close()


def test_full_size_properties_1080p():
    """BASELINE config[1] size, checked through size-independent properties: determinism,
    D = G[s+1] − G[s], octave seeding by decimation, batch ≡ single, canonical ordering."""
    from siftmetal_b200.synth import pink_noise_bgra

    w, h = 1920, 1080
    img = pink_noise_bgra(w, h, 0)
    eng = _engine(w, h)
    a = eng.detect_and_describe([img])
    b = eng.detect_and_describe([img])
    assert np.array_equal(a.keypoints, b.keypoints) and np.array_equal(a.descriptors, b.descriptors)
    assert 15000 < len(a.keypoints) < 45000, len(a.keypoints)   # ≈ 13.7 per 1000 px
    for o in (0, 3):
        g = [eng.plane(_abi.PLANE_GAUSSIAN, o, s) for s in range(6)]
        for s in range(5):
            assert np.array_equal(eng.plane(_abi.PLANE_DOG, o, s), g[s + 1] - g[s])
        nxt = eng.plane(_abi.PLANE_GAUSSIAN, o + 1, 0)
        assert np.array_equal(nxt, g[3][::2, ::2][: nxt.shape[0], : nxt.shape[1]])
    k = a.keypoints
    # canonical order: octave-major
    assert np.all(np.diff(k["octave"]) >= 0)
    # every descriptor points at a keypoint of its own frame, θ in [0, 2π)
    d = a.descriptors
    assert d["keypoint"].min() >= 0 and d["keypoint"].max() < len(k)
    assert np.all(np.diff(d["keypoint"]) >= 0)
    assert np.all((d["theta"] >= 0) & (d["theta"] < 2 * np.pi + 1e-6))
    # descriptor features: clipped-normalised vectors ⇒ L2 norm of features/512 close to 1
    nrm = np.linalg.norm(d["features"].astype(np.float32) / 512.0, axis=1)
    assert np.all(nrm < 1.02) and np.median(nrm) > 0.9
    eng.close()
    # spot parity against the oracle on a crop-sized problem is covered above; here compare the
    # 1080p candidate COUNT per octave with the oracle (cheap: detection only)
    ora = _oracle(w, h)
    okps, ocounts = ora.detect(img)
    assert np.array_equal(a.keypoint_counts[0], ocounts)
    # ... and the full result of the fused host-buffer call (two-chunk upload, seed stage and
    # band 0 started on the first chunk, host-resident outputs) against the oracle at full size
    eng = _engine(w, h)
    res = eng.detect_and_describe([img])
    assert np.array_equal(eng.plane(_abi.PLANE_SEED), ora.plane(_abi.PLANE_SEED))
    for s in (1, 5):
        assert np.array_equal(eng.plane(_abi.PLANE_GAUSSIAN, 0, s), ora.plane(_abi.PLANE_GAUSSIAN, 0, s))
    kps, desc = res.frame(0)
    _check_frame(eng, ora, img, kps, desc, res.keypoint_counts[0], res.descriptor_counts[0],
                 res.candidate_counts[0], planes=False)
    # the staged path with a device-resident input (no upload chunks) gives the same arrays
    import torch
    dev = torch.from_numpy(img).cuda()
    eng.set_device_input(dev.data_ptr(), 1, w * 4, w * h * 4)
    eng.execute()
    res2 = eng.download()
    assert np.array_equal(res.keypoints, res2.keypoints) and np.array_equal(res.descriptors, res2.descriptors)
    eng.close()
    assert np.array_equal(k["scaledX"], okps["scaledX"]) and np.array_equal(k["scaledY"], okps["scaledY"])
    assert np.all(np.abs(k["absoluteX"] - okps["absoluteX"]) <= POS_TOL)


def _free_host_gb():
    try:
        with open("/proc/meminfo") as f:
            for line in f:
                if line.startswith("MemAvailable"):
                    return int(line.split()[1]) / 1e6
    except OSError:
        pass
    return 0.0


def test_config3_4k_frame_against_oracle():
    """BASELINE configs[3] frame size (3840x2160): candidates, keypoints, orientations and
    descriptors of one full-size frame against the oracle (planes are covered at 1080p)."""
    from siftmetal_b200.synth import pink_noise_bgra

    w, h = 3840, 2160
    img = pink_noise_bgra(w, h, 11)
    eng, ora = _engine(w, h), _oracle(w, h)
    res = eng.detect_and_describe([img])
    kps, desc = res.frame(0)
    rep = _check_frame(eng, ora, img, kps, desc, res.keypoint_counts[0], res.descriptor_counts[0],
                       res.candidate_counts[0], planes=False, max_structural=1e-4)
    assert rep["keypoints_bit_exact"]
    assert len(kps) > 60000          # ~ 13.7 keypoints per 1000 input pixels
    # seed plane and the last Gaussian of octave 0 (largest planes of the config) bit-exact
    assert np.array_equal(eng.plane(_abi.PLANE_SEED), ora.plane(_abi.PLANE_SEED))
    assert np.array_equal(eng.plane(_abi.PLANE_GAUSSIAN, 0, 5), ora.plane(_abi.PLANE_GAUSSIAN, 0, 5))
    eng.close()


def test_config4_8k_tile_against_oracle():
    """BASELINE configs[4]: one 8192x8192 tile (octave 0 is 16384^2: 15-bit packed coordinates,
    the largest mask / list capacities) against the oracle at full size."""
    from siftmetal_b200.synth import pink_noise_bgra

    if _free_host_gb() < 45:
        pytest.skip("the oracle needs ~30 GB of host memory at 8192x8192")
    w, h = 8192, 8192
    img = pink_noise_bgra(w, h, 21)
    eng, ora = _engine(w, h), _oracle(w, h)
    res = eng.detect_and_describe([img])
    kps, desc = res.frame(0)
    rep = _check_frame(eng, ora, img, kps, desc, res.keypoint_counts[0], res.descriptor_counts[0],
                       res.candidate_counts[0], planes=False, max_structural=1e-4)
    assert rep["keypoints_bit_exact"]
    assert len(kps) > 500000
    assert kps["scaledX"].max() > 16000 and kps["scaledY"].max() > 16000   # the far corner is reached
    for o in (0, 6):
        assert np.array_equal(eng.plane(_abi.PLANE_DOG, o, 4), ora.plane(_abi.PLANE_DOG, o, 4))
    eng.close()


def test_config2_vga256_batch_against_oracle():
    """BASELINE configs[2]: 256 frames of 640x480 in one resident batch. Every frame's per-octave
    candidate / keypoint / descriptor counts against the oracle; frames 0, 127 and 255 in full."""
    from siftmetal_b200.synth import pink_noise_bgra

    w, h, n = 640, 480, 256
    uniq = [pink_noise_bgra(w, h, 100 + i) for i in range(32)]
    order = [(7 * i + i // 32) % 32 for i in range(n)]          # every unique frame at 8 batch positions
    frames = [uniq[k] for k in order]
    eng, ora = _engine(w, h, max_batch=n), _oracle(w, h)
    res = eng.detect_and_describe(frames)
    assert res.keypoint_counts.shape == (n, 7)
    ref = {}
    for k in range(32):
        okps, oc = ora.detect(uniq[k])
        odesc, odc = ora.describe()
        cc = np.array([len(ora.candidates(o)) for o in range(7)])
        ref[k] = (oc.copy(), odc.copy(), cc, okps, odesc)
    structural = 0
    for f in range(n):
        oc, odc, cc, okps, odesc = ref[order[f]]
        assert np.array_equal(res.keypoint_counts[f], oc), f
        assert np.array_equal(res.candidate_counts[f], cc), f
        structural += int(np.abs(res.descriptor_counts[f] - odc).sum())
    assert structural <= 1e-4 * res.keypoint_counts.sum(), structural
    for f in (0, 127, 255):
        kps, desc = res.frame(f)
        _check_frame(eng, ora, frames[f], kps, desc, res.keypoint_counts[f], res.descriptor_counts[f],
                     res.candidate_counts[f], planes=(f == 255), frame=f, max_structural=1e-3)
    # same unique frame at different batch positions: identical records
    a, b = [i for i in range(n) if order[i] == 5][:2]
    ka, da = res.frame(a)
    kb, db = res.frame(b)
    assert np.array_equal(ka, kb) and np.array_equal(da["features"], db["features"])
    eng.close()


@pytest.mark.parametrize("switch", ["SIFTCUDA_GRAPH=1", "SIFTCUDA_BANDS=1", "SIFTCUDA_BANDS=3", "SIFTCUDA_PDL=0",
                                    "SIFTCUDA_BLUR_TMA=0", "SIFTCUDA_EXTREMA_TMA=0", "SIFTCUDA_TAIL=0",
                                    "SIFTCUDA_RESULT_COPY=2"])
def test_tuning_switches_do_not_change_results(switch):
    """Alternative schedules (CUDA-graph replay instead of stream launches; no row bands; three row
    bands; no programmatic dependent launch; cp.async instead of TMA tile loads; the register-march
    extrema kernel instead of the TMA-fed one; the deepest octave launch by launch instead of in the
    one-CTA tail kernel; result columns through HBM and the copy-out kernel instead of direct
    stores to host memory) must give the same
    result arrays as the default, call after call (graph: first call eager, second captured, third replayed)."""
    import os
    import subprocess
    import sys

    code = (
        "import sys, numpy as np; sys.path.insert(0, '.');"
        "from siftmetal_b200 import Engine; from siftmetal_b200.synth import pink_noise_bgra;"
        "img = pink_noise_bgra(1920, 1080, 5); e = Engine(1920, 1080);"
        "rs = [e.detect_and_describe([img]) for _ in range(3)];"
        "assert all(np.array_equal(rs[0].keypoints, r.keypoints) and np.array_equal(rs[0].descriptors, r.descriptors) for r in rs);"
        "r = rs[-1]; c = sum(len(e.candidates(o)) for o in range(7));"
        "np.savez(sys.argv[1], k=r.keypoints, d=r.descriptors, kc=r.keypoint_counts, dc=r.descriptor_counts, c=c,"
        " g=int(e.timings()['graph_replay']))"
    )
    outs = []
    for flag in ("0", "1"):
        path = f"/tmp/sift_switch_{flag}.npz"
        env = dict(os.environ)
        if flag == "1":
            env.update(kv.split("=") for kv in switch.split())
        subprocess.run([sys.executable, "-c", code, path], check=True, env=env,
                       cwd=os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
        outs.append(np.load(path))
    a, b = outs
    assert int(a["g"]) == 0                                   # the default launches on streams
    assert int(b["g"]) == (1 if "GRAPH=1" in switch else 0)
    assert int(a["c"]) == int(b["c"])
    for key in ("k", "d", "kc", "dc"):
        assert np.array_equal(a[key], b[key]), key


@pytest.mark.parametrize("parts", ["0", "3"])
def test_descriptor_walk_variants_stay_within_tolerance(parts):
    """SIFTCUDA_DESC_PARTS selects how a warp walks a descriptor window (0: flattened spans, n:
    units of 1/n of a window row; 4 ships). The walks add the same samples in a different order,
    so keypoints / orientations are identical and features agree within the ±1 of the parity bar —
    each variant is also within ±1 of the oracle."""
    import os
    import subprocess
    import sys

    from siftmetal_b200.synth import pink_noise_bgra

    code = (
        "import sys, numpy as np; sys.path.insert(0, '.');"
        "from siftmetal_b200 import Engine; from siftmetal_b200.synth import pink_noise_bgra;"
        "img = pink_noise_bgra(640, 480, 9); e = Engine(640, 480); r = e.detect_and_describe([img]);"
        "np.savez(sys.argv[1], k=r.keypoints, d=r.descriptors)"
    )
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    path = f"/tmp/sift_parts_{parts}.npz"
    subprocess.run([sys.executable, "-c", code, path], check=True, env={**os.environ, "SIFTCUDA_DESC_PARTS": parts}, cwd=root)
    v = np.load(path)
    eng, ora = _engine(640, 480), _oracle(640, 480)
    img = pink_noise_bgra(640, 480, 9)
    res = eng.detect_and_describe([img])
    assert np.array_equal(v["k"], res.keypoints)
    assert np.array_equal(v["d"]["keypoint"], res.descriptors["keypoint"])
    assert np.array_equal(v["d"]["theta"], res.descriptors["theta"])
    df = np.abs(v["d"]["features"].astype(np.int16) - res.descriptors["features"].astype(np.int16))
    assert df.max(initial=0) <= FEAT_TOL
    ora.detect(img)
    odesc, _ = ora.describe()
    assert np.abs(v["d"]["features"].astype(np.int16) - odesc["features"].astype(np.int16)).max(initial=0) <= FEAT_TOL
    eng.close()
