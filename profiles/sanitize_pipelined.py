"""compute-sanitizer target: pipelined sift_submit / sift_wait over both slots, small frames.

    compute-sanitizer --tool memcheck python profiles/sanitize_pipelined.py
    SIFTCUDA_RESULT_COPY=1 compute-sanitizer --tool memcheck python profiles/sanitize_pipelined.py
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from siftmetal_b200 import Engine  # noqa: E402
from siftmetal_b200.synth import pink_noise_bgra  # noqa: E402

w, h, n = 320, 240, 2
eng = Engine(w, h, max_batch=n)
batches = [[pink_noise_bgra(w, h, 3 * b + i) for i in range(n)] for b in range(3)]
ref = [eng.detect_and_describe(b) for b in batches]
got, inflight = [], 0
for i in range(6):
    if inflight == 2:
        got.append(eng.wait())
        inflight -= 1
    eng.submit(batches[i % 3])
    inflight += 1
while inflight:
    got.append(eng.wait())
    inflight -= 1
for i, r in enumerate(got):
    a = ref[i % 3]
    assert np.array_equal(a.keypoints, r.keypoints) and np.array_equal(a.descriptors, r.descriptors), i
m = eng.match_frames(0, 1)
eng.close()
print("pipelined ok:", sum(len(r.keypoints) for r in got), "keypoints,", len(m), "matches, result copy",
      os.environ.get("SIFTCUDA_RESULT_COPY", "0"))
