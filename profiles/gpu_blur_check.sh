timeout 800 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python - <<'PY'
import sys
sys.path.insert(0, '.')
from siftmetal_b200 import Engine
from siftmetal_b200.synth import pink_noise_bgra
w, h = 1920, 1080
eng = Engine(w, h)
img = pink_noise_bgra(w, h, 0)
row = []
for scale in range(5):
    eng.detect_and_describe([img])
    row.append(eng.blur_bench(scale, 0, 30) * 1000)
print("blur us per launch (11/15/17/21/27 taps): " + "  ".join(f"{t:6.1f}" for t in row), " sum %.1f" % sum(row))
PY
bash profiles/gpu_ab.sh
