"""One-off parity check of large odd-sized frames (banded octave 0 + chunked upload + odd widths)
against the oracle: python profiles/odd_sizes_check.py"""
import sys
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import numpy as np
from siftmetal_b200 import Engine, _abi
from siftmetal_b200.synth import pink_noise_bgra
from oracle_lib import Oracle
import test_gpu_parity as T
for (w, h) in [(2001, 1203), (1367, 911)]:
    img = pink_noise_bgra(w, h, 7)
    eng, ora = Engine(w, h), Oracle(w, h)
    res = eng.detect_and_describe([img])
    kps, desc = res.frame(0)
    rep = T._check_frame(eng, ora, img, kps, desc, res.keypoint_counts[0], res.descriptor_counts[0], res.candidate_counts[0], planes=True)
    print(w, h, "ok", len(kps), len(desc), rep)
    eng.close()
