"""One 1080p frame through the fused host-buffer call (two-chunk upload, row bands, PDL chains,
host-resident results) — run under compute-sanitizer by profiles/sanitize.sh."""
import sys
sys.path.insert(0, '.')
from siftmetal_b200 import Engine
from siftmetal_b200.synth import pink_noise_bgra
eng = Engine(1920, 1080)
res = eng.detect_and_describe([pink_noise_bgra(1920, 1080, 3)])
print("1080p ok", len(res.keypoints), len(res.descriptors))
eng.close()
