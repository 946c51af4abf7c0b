"""Octave-0 blur launch time per tap count: tiled kernel vs strip-streaming kernel (B200).
30 back-to-back launches of each scale, CUDA events (sift_debug_blur_bench)."""
import os
import sys
sys.path.insert(0, '.')
from siftmetal_b200 import Engine
from siftmetal_b200.synth import pink_noise_bgra
w, h = 1920, 1080
eng = Engine(w, h)
img = pink_noise_bgra(w, h, 0)
variants = [("tiled 64x64", {"SIFTCUDA_BLUR_STREAM": "0"})]
for ctas in sys.argv[1:] or ["3", "2"]:
    variants.append((f"stream, {ctas} CTAs/SM", {"SIFTCUDA_BLUR_STREAM": "1", "SIFTCUDA_STREAM_CTAS": ctas}))
for name, env in variants:
    os.environ.update(env)
    row = []
    for scale in range(5):
        eng.detect_and_describe([img])
        row.append(eng.blur_bench(scale, 0, 30) * 1000)
    print(f"{name:22s}: " + "  ".join(f"{t:6.1f}us" for t in row) + f"   sum {sum(row):6.1f}us")
