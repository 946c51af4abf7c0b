"""Tuning aid: octave-0 blur launch time per tap count in the four debug modes (B200)."""
import sys
sys.path.insert(0, '.')
from siftmetal_b200 import Engine
from siftmetal_b200.synth import pink_noise_bgra
w, h = 1920, 1080
eng = Engine(w, h)
img = pink_noise_bgra(w, h, 0)
names = {0: "normal", 1: "no FMA loops", 2: "no stores", 3: "no FMA, no stores", 4: "tile load only", 8: "normal, two streams"}
for mode in (0, 1, 2, 3, 4, 8):
    row = []
    for scale in range(5):
        eng.detect_and_describe([img])
        row.append(eng.blur_bench(scale, mode, 30) * 1000)
    print(f"mode {mode} ({names.get(mode, 'stagger %d ns' % ((mode >> 8) * 100)):18s}): " + "  ".join(f"{t:6.1f}us" for t in row))
