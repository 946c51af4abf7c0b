import ctypes as C, sys, time
sys.path.insert(0, ".")
import numpy as np, torch
from siftmetal_b200 import Engine
from siftmetal_b200.synth import pink_noise_bgra
w, h = 1920, 1080
img = pink_noise_bgra(w, h, 0)
pinned = torch.empty((h, w, 4), dtype=torch.uint8, pin_memory=True); pinned.numpy()[:] = img
ptr = (C.c_void_p * 1)(pinned.numpy().ctypes.data)
e = Engine(w, h); e.set_graph_replay(True)
def run(n, depth):
    replays = 0; q = 0; t0 = time.perf_counter(); lat = []
    for i in range(n):
        if q == depth:
            ta = time.perf_counter(); e.wait_counts(); lat.append(time.perf_counter() - ta); q -= 1
            replays += e.timings()["graph_replay"]
        tb = time.perf_counter(); e.submit_ptrs(ptr, 1, w * 4); lat.append(-(time.perf_counter() - tb)); q += 1
    while q:
        e.wait_counts(); q -= 1; replays += e.timings()["graph_replay"]
    dt = time.perf_counter() - t0
    sub = [-x for x in lat if x < 0]
    print(f"depth {depth}: {n/dt:.1f} frames/s, replays {replays}/{n}, submit mean {1e3*np.mean(sub):.3f} ms max {1e3*np.max(sub):.3f} ms")
for depth in (1, 2, 2, 1, 2):
    run(100, depth)
