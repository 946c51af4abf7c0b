"""Device matcher throughput: n x n descriptors (uint8 x 128), device-resident feature matrices,
CUDA events around launchMatch through sift_match's device path (H2D excluded by matching twice and
timing the second call's kernels via the context's stream). Prints one JSON line."""
import ctypes as C, json, sys, time
sys.path.insert(0, ".")
import numpy as np, torch
from siftmetal_b200 import Engine

n = int(sys.argv[1]) if len(sys.argv) > 1 else 35000
rng = np.random.default_rng(1)
b = rng.integers(0, 256, (n, 128), dtype=np.uint8)
a = np.clip(b[rng.permutation(n)].astype(np.int32) + rng.integers(-8, 9, (n, 128)), 0, 255).astype(np.uint8)
e = Engine(64, 48)
m = e.match(a, b)                      # warm-up (allocations, attribute)
reps = 5
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(reps):
    m = e.match(a, b)
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / reps
peaks = json.load(open("MEASURED_PEAKS.json")) if __import__("os").path.exists("MEASURED_PEAKS.json") else {}
ops = 2.0 * n * n * 128
print(json.dumps({"n_source": n, "n_target": n, "matches": int(len(m)), "ms_per_call_incl_h2d_d2h": 1000 * dt,
                  "tera_ops_per_s_incl_copies": ops / dt / 1e12,
                  "note": "wall clock around sift_match incl. H2D of both feature matrices (2 x n x 128 B) and D2H of the rows; "
                          "the kernel alone is in the ncu capture (profiles/r2/SUMMARY.md)",
                  "bf16_tflops_measured_peak": peaks.get("bf16_tflops")}))
e.close()
