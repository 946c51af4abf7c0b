"""Where the gap between the device-resident rate and the pipelined end-to-end rate comes from.

Runs the pipelined sift_submit / sift_wait loop of bench.py (1080p, one frame per call, two calls in
flight) three ways: direct stores of the result columns to host memory (SIFTCUDA_RESULT_COPY=0), the
slot's HBM columns + copy engine at sift_wait (=1, the default for pipelined calls), and with the columns left in HBM
(SIFTCUDA_DEBUG_E2E=1: nothing delivered — the floor). One process per variant. Usage: python profiles/exp_e2e_gap.py [calls]
"""
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def child(calls):
    sys.path.insert(0, ROOT)
    import numpy as np
    import torch
    from siftmetal_b200 import Engine
    from siftmetal_b200.synth import pink_noise_bgra

    w, h = 1920, 1080
    eng = Engine(w, h, device=0, max_batch=1)
    host = torch.empty((h, w, 4), dtype=torch.uint8).pin_memory()
    host.numpy()[:] = pink_noise_bgra(w, h, 0)
    import ctypes
    ptrs = (ctypes.c_void_p * 1)(host.data_ptr())

    host = {"submit": 0.0, "wait": 0.0}

    def loop(n):
        inflight = 0
        for _ in range(n):
            if inflight == 2:
                t = time.perf_counter()
                eng.wait(copy=False)
                host["wait"] += time.perf_counter() - t
                inflight -= 1
            t = time.perf_counter()
            eng.submit_ptrs(ptrs, 1, w * 4)
            host["submit"] += time.perf_counter() - t
            inflight += 1
        while inflight:
            eng.wait(copy=False)
            inflight -= 1

    loop(10)
    torch.cuda.synchronize()
    out = []
    host["submit"] = host["wait"] = 0.0
    for rep in range(3):
        t0 = time.perf_counter()
        loop(calls)
        torch.cuda.synchronize()
        out.append((time.perf_counter() - t0) / calls * 1e3)
    dev_ms = None
    if os.environ.get("SIFTCUDA_DEBUG_E2E", "0") == "0":
        # the device-resident step of bench.py, here without the L2 flush between steps
        dev = torch.from_numpy(pink_noise_bgra(w, h, 0)).cuda()
        eng.set_device_input(dev.data_ptr(), 1, w * 4, w * h * 4)
        for _ in range(5):
            eng.execute()
        dev_ms = 0.0
        for _ in range(calls):
            eng.execute()
            dev_ms += eng.timings()["total_ms"]
        dev_ms /= calls
    print(json.dumps({"device_ms_no_flush": dev_ms, "variant": os.environ.get("SIFTCUDA_DEBUG_E2E", "0"),
                      "result_copy": os.environ.get("SIFTCUDA_RESULT_COPY", "1"), "graph": os.environ.get("SIFTCUDA_GRAPH", "0"), "ms_per_call": out,
                      "host_submit_ms": host["submit"] / (3 * calls) * 1e3, "host_wait_ms": host["wait"] / (3 * calls) * 1e3}))


if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[1] == "child":
        child(int(sys.argv[2]))
    else:
        calls = sys.argv[1] if len(sys.argv) > 1 else "300"
        for graph in ("0", "1"):
            for copy, v in (("0", "0"), ("1", "0"), ("0", "1")):
                env = dict(os.environ, SIFTCUDA_DEBUG_E2E=v, SIFTCUDA_RESULT_COPY=copy, SIFTCUDA_GRAPH=graph)
                subprocess.run([sys.executable, os.path.abspath(__file__), "child", calls], env=env, check=False)
