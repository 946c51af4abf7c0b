"""Two contexts on one device, frames alternating between them: does the descriptor stage of one
frame overlap the pyramid stage of the next? Pipelined e2e throughput, 1080p."""
import ctypes as C, sys, time, os
sys.path.insert(0, ".")
import numpy as np, torch
from siftmetal_b200 import Engine
from siftmetal_b200.synth import pink_noise_bgra

w, h = 1920, 1080
lanes = int(sys.argv[1]) if len(sys.argv) > 1 else 2
steps = 400
frames = [pink_noise_bgra(w, h, i) for i in range(4)]
pinned = torch.empty((4, h, w, 4), dtype=torch.uint8, pin_memory=True)
pn = pinned.numpy()
for i, f in enumerate(frames):
    pn[i] = f
ptrs = [(C.c_void_p * 1)(pn[i].ctypes.data) for i in range(4)]
engs = [Engine(w, h) for _ in range(lanes)]
def run(n):
    q = []
    depth = 2 * lanes
    for i in range(n):
        if len(q) == depth:
            q.pop(0).wait_counts()
        e = engs[i % lanes]
        e.submit_ptrs(ptrs[i % 4], 1, w * 4)
        q.append(e)
    while q:
        q.pop(0).wait_counts()
run(20)
torch.cuda.synchronize()
t0 = time.perf_counter(); run(steps); torch.cuda.synchronize(); dt = time.perf_counter() - t0
print(f"lanes {lanes} DESC_CTAS={os.environ.get('SIFTCUDA_DESC_CTAS','-')}: {steps / dt:.1f} frames/s e2e pipelined ({1000 * dt / steps:.4f} ms per frame)")
