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


def test_pipelined_copy_out_under_load():
    """Result delivery of pipelined calls, both ways: direct stores of the kernels into the slot's
    host columns (default) and the slot's HBM columns + copy engine at sift_wait
    (SIFTCUDA_RESULT_COPY=1, read at create). 24 back-to-back 1080p calls over three alternating
    frames, every result equal to the synchronous call's."""
    import os

    from siftmetal_b200 import Engine
    from siftmetal_b200.synth import pink_noise_bgra

    w, h = 1920, 1080
    imgs = [pink_noise_bgra(w, h, 20 + i) for i in range(3)]
    eng = Engine(w, h)
    ref = [eng.detect_and_describe([im]) for im in imgs]

    def run(e):
        got, inflight = [], 0
        for i in range(24):
            if inflight == 2:
                got.append(e.wait())
                inflight -= 1
            e.submit([imgs[i % 3]])
            inflight += 1
        while inflight:
            got.append(e.wait())
            inflight -= 1
        return got

    for i, r in enumerate(run(eng)):
        assert _same(r, ref[i % 3]), i
    eng.close()
    os.environ["SIFTCUDA_RESULT_COPY"] = "1"
    try:
        eng = Engine(w, h)
        for i, r in enumerate(run(eng)):
            assert _same(r, ref[i % 3]), i
        eng.close()
    finally:
        del os.environ["SIFTCUDA_RESULT_COPY"]


def test_second_pipeline_of_small_contexts():
    """A small context runs lane 1 of the pipelined calls on a second pipeline of its own (scratch
    planes, streams, slot), created by the first overlapping submit, so that consecutive calls
    overlap on the device. Same results as the synchronous call and as the one-pipeline form
    (SIFTCUDA_TWIN=0, other process); device matching follows the lane the last call ran on."""
    import os
    import subprocess
    import sys

    from siftmetal_b200 import Engine
    from siftmetal_b200.synth import pink_noise_bgra

    w, h = 960, 540
    imgs = [pink_noise_bgra(w, h, 40 + i) for i in range(3)]
    eng = Engine(w, h, max_batch=2)
    ref = [eng.detect_and_describe([imgs[i], imgs[(i + 1) % 3]]) for i in range(3)]
    bytes_one = eng.device_bytes()
    got, inflight = [], 0
    for i in range(9):
        if inflight == 2:
            got.append(eng.wait())
            inflight -= 1
        eng.submit([imgs[i % 3], imgs[(i + 1) % 3]])
        inflight += 1
    while inflight:
        got.append(eng.wait())
        inflight -= 1
    slots = [r.slot for r in got]
    assert all(a != b for a, b in zip(slots, slots[1:])) and set(slots) == {0, 1}    # lanes alternate
    for i, r in enumerate(got):
        assert _same(r, ref[i % 3]), i
    assert eng.device_bytes() > 1.9 * bytes_one            # the second pipeline exists now

    def check_device_match(r):
        m_dev = eng.match_frames(0, 1)
        d = r.descriptors
        k = int(r.descriptor_counts[0].sum())
        assert np.array_equal(m_dev, eng.match(d["features"][:k], d["features"][k:]))

    # device matching reads the descriptor column of the lane the last completed call ran on
    eng.submit([imgs[0], imgs[1]])
    eng.submit([imgs[1], imgs[2]])
    eng.wait()
    rb = eng.wait()
    check_device_match(rb)
    eng.submit([imgs[2], imgs[0]])
    rc = eng.wait()
    assert rc.slot != rb.slot
    check_device_match(rc)
    eng.close()

    code = (
        "import sys, numpy as np; sys.path.insert(0, '.');"
        "from siftmetal_b200 import Engine; from siftmetal_b200.synth import pink_noise_bgra;"
        "imgs = [pink_noise_bgra(960, 540, 40 + i) for i in range(3)]; e = Engine(960, 540, max_batch=2);"
        "b0 = e.device_bytes();"
        "e.submit([imgs[0], imgs[1]]); e.submit([imgs[1], imgs[2]]); a = e.wait(); b = e.wait();"
        "assert e.device_bytes() < 1.5 * b0;"
        "np.savez(sys.argv[1], k0=a.keypoints, d0=a.descriptors, k1=b.keypoints, d1=b.descriptors)"
    )
    path = "/tmp/sift_twin_0.npz"
    subprocess.run([sys.executable, "-c", code, path], check=True, env={**os.environ, "SIFTCUDA_TWIN": "0"},
                   cwd=os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    v = np.load(path)
    assert np.array_equal(v["k0"], ref[0].keypoints) and np.array_equal(v["d0"], ref[0].descriptors)
    assert np.array_equal(v["k1"], ref[1].keypoints) and np.array_equal(v["d1"], ref[1].descriptors)


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


def test_cpp_host_mirror_runs(tmp_path):
    """include/SIFT.hpp on the device: batch call, pipelined submit / wait, lazy views, match."""
    import os
    import subprocess

    from conftest import ROOT
    from siftmetal_b200 import api
    from siftmetal_b200.synth import pink_noise_bgra

    img = pink_noise_bgra(320, 240, 2)
    raw = tmp_path / "frame.bgra"
    raw.write_bytes(img.tobytes())
    prog = tmp_path / "run.cpp"
    prog.write_text(
        '#include "SIFT.hpp"\n#include <cstdio>\n#include <vector>\n'
        "int main(int argc, char** argv){ std::vector<unsigned char> px(320 * 240 * 4);\n"
        " FILE* f = std::fopen(argv[1], \"rb\"); if (!f || std::fread(px.data(), 1, px.size(), f) != px.size()) return 2;\n"
        " siftcuda::SIFT s(0, siftcuda::SIFT::Configuration(siftcuda::IntegralSize{320, 240}));\n"
        " std::vector<const void*> fr{px.data()};\n"
        " auto a = s.detectAndDescribe(fr, 320 * 4); s.submit(fr, 320 * 4); auto b = s.wait();\n"
        " auto m = s.match(a[0].descriptors, b[0].descriptors);\n"
        " auto k = a[0].keypoints[3]; auto d = a[0].descriptors[5];\n"
        " auto octs = s.getKeypoints(px.data(), 320 * 4); size_t nk = 0; for (auto& o : octs) nk += o.size();\n"
        " std::printf(\"%lld %lld %zu %zu %d %d\\n\", (long long)a[0].keypoints.size(), (long long)a[0].descriptors.size(),\n"
        "   m.size(), nk, k.scaledCoordinate[0], (int)d.features.components.size()); return 0; }\n")
    exe = tmp_path / "run"
    libdir = os.path.dirname(api.LIB_PATH)
    subprocess.run(["g++", "-std=c++17", "-O1", "-I", os.path.join(ROOT, "include"), str(prog), "-o", str(exe),
                    "-L", libdir, "-lsiftcuda", f"-Wl,-rpath,{libdir}"], check=True)
    r = subprocess.run([str(exe), str(raw)], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    nk, nd, nm, nk2, x, nf = [int(v) for v in r.stdout.split()]
    from siftmetal_b200 import Engine

    e = Engine(320, 240)
    ref = e.detect_and_describe([img])
    assert nk == len(ref.keypoints) == nk2 and nd == len(ref.descriptors) and nf == 128
    assert x == int(ref.keypoints["scaledX"][3])
    # a descriptor set matched against itself: every descriptor whose nearest neighbour is itself at
    # distance 0 and passes 0 < 0.6 * second is kept unless an earlier duplicate exists
    assert nm == len(e.match(ref.descriptor_columns.features, ref.descriptor_columns.features))
    e.close()


def test_results_stored_into_caller_memory():
    """sift_register_host_memory / sift_bind_result_memory: the slots' result columns live in
    caller memory (here a POSIX shared-memory segment, as the multi-process gather uses) and the
    kernels store into it directly; rotating a slot through several blocks keeps older results."""
    import ctypes as C
    from multiprocessing import shared_memory

    from siftmetal_b200 import Engine, SiftError
    from siftmetal_b200.sharding import _COLS, block_layout
    from siftmetal_b200.synth import pink_noise_bgra

    w, h = 320, 240
    imgs = [pink_noise_bgra(w, h, 60 + i) for i in range(4)]
    eng = Engine(w, h)
    ref = [eng.detect_and_describe([im]) for im in imgs]
    lay = eng.result_layout()
    offsets, nbytes = block_layout(int(lay.capacity_keypoints), int(lay.capacity_descriptors))
    assert offsets == list(lay.offset) and nbytes == int(lay.bytes)
    blocks = 4
    stride = (nbytes + 4095) // 4096 * 4096
    shm = shared_memory.SharedMemory(create=True, size=blocks * stride + 4096)
    try:
        base = C.addressof(C.c_char.from_buffer(shm.buf))
        pad = (-base) % 4096
        with pytest.raises(SiftError):
            eng.bind_result_memory(0, base + pad, nbytes)            # not registered yet
        eng.register_host_memory(base + pad, blocks * stride)
        for i, im in enumerate(imgs):                                # four pipelined calls, a block each
            if eng.pending() == 2:
                eng.wait(copy=False)
            eng.bind_result_memory(eng.next_slot(), base + pad + i * stride, nbytes)
            eng.submit([im])
        while eng.pending():
            eng.wait(copy=False)
        for i in range(blocks):                                      # every block still holds its call's columns
            r = ref[i]
            nk, nd = len(r.keypoints), len(r.descriptors)
            got = {}
            for (group, name, dtype, per), off in zip(_COLS, offsets):
                rows = nk if group == "kp" else nd
                got[name] = np.ndarray((rows, per) if per > 1 else (rows,), dtype, shm.buf, pad + i * stride + off)
            assert np.array_equal(got["absolute_x"], r.keypoint_columns.absolute_x)
            assert np.array_equal(got["scaled_xy"], r.keypoint_columns.scaled_xy)
            assert np.array_equal(got["octave_scale"], r.keypoint_columns.octave_scale)
            assert np.array_equal(got["features"], r.descriptor_columns.features)
            assert np.array_equal(got["theta"], r.descriptor_columns.theta)
            assert np.array_equal(got["keypoint"], r.descriptor_columns.keypoint)
            del got
        eng.close()
    finally:
        try:
            shm.close()
        except BufferError:
            pass
        shm.unlink()
