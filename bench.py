#!/usr/bin/env python
"""bench.py — SIFT detect + describe throughput on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload 1080p|vga256|4k64|8k]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference ...       # the reference algorithm on the host CPU cores

A step = one pass of the whole hot path (seed → pyramid/DoG → extrema → refine → gradient →
orientation → descriptor) over one batch of synthetic 1/f-noise frames.
Default workload = BASELINE.json configs[1]: a single 1920×1080 frame per step per GPU.

  value      frames/s with the input already resident in HBM: device time between two CUDA events the
             library records on its own stream around each call's work, summed over the K steps, L2
             flushed between steps; max over ranks
  e2e        frames/s through the public pipelined C-ABI calls (sift_submit / sift_wait, two calls
             in flight) with HOST (pinned) frames: every step's H2D of its frames and the arrival of
             its keypoint + descriptor columns in host memory are inside the wall-clock timed region
             (the kernels store the result columns straight into pinned memory; the upload of step
             i + 1 crosses PCIe under the kernels of step i; a small context — the single-frame
             workloads — runs the two calls in flight on two pipelines of its own, so their kernels
             also overlap on the device, which the one-step-at-a-time `value` does not). For a sharded workload on N > 1 GPUs
             the host gather (every rank's result columns visible to rank 0 in frame order: the
             slots are bound to shared-memory segments, one barrier per call) is inside the region too. `sync_call` is the same through the
             synchronous sift_detect_and_describe_batch (no overlap between calls).
  roofline   the dominant streaming kernel (octave-0 Gaussian blur + DoG, 5 scales per step):
             algorithmic bytes (12 B per octave-0 pixel per scale: read G[s], write G[s+1], write
             DoG[s]) ÷ its mean time per scale from CUDA events over a second region of the same
             steps with per-stage timing switched on (it costs ~20 event records per call, so the
             timed region runs without it)
  cpu_baseline   the C++ oracle (a port of the reference's kernels + host stages; the Swift/Metal
             reference cannot run on Linux) on the host cores, bounded sample

The reference arm (--impl reference) times that same oracle with all host threads on the same
workload; it is the only place (with cpu_baseline) where bench.py executes anything in oracle/.
"""
import argparse
import ctypes as C
import json
import os
import platform
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (width, height, frames in the whole job, sharded across ranks?)
    "1080p": (1920, 1080, 1, False),     # configs[1]: single image per GPU (replicas when N > 1)
    "vga256": (640, 480, 256, True),     # configs[2]
    "4k64": (3840, 2160, 64, True),      # configs[3]
    "8k": (8192, 8192, 1, False),        # configs[4]
}
METRIC = "1080p SIFT frames/sec (detect+describe)"
STAGES = ("seed", "pyramid", "extrema", "refine", "orientation", "descriptor")
NOMINAL_HBM_GBS = 8000.0   # north_star's "about 8 TB/s"; fractions are quoted against both peaks


def workload_string(name):
    """config.workload — the same string in both arms (the driver compares them)."""
    w, h, total, sharded = WORKLOADS[name]
    if sharded:
        return f"{name}: {total} x {w}x{h} synthetic 1/f-noise BGRA8 frames per step, sharded across the GPUs"
    return f"{name}: one {w}x{h} synthetic 1/f-noise BGRA8 frame per step per GPU"


def parse():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=200)
    p.add_argument("--warmup", type=int, default=5)
    p.add_argument("--impl", default="ours", choices=["ours", "reference"])
    p.add_argument("--workload", default="1080p", choices=sorted(WORKLOADS))
    p.add_argument("--chunk", type=int, default=0, help="frames resident per execute (0 = auto)")
    p.add_argument("--input-format", default="bgra8", choices=["bgra8", "gray8"],
                   help="ingestion format of the e2e / value paths (bgra8 = the reference's)")
    p.add_argument("--no-cpu-baseline", action="store_true")
    p.add_argument("--no-flush", action="store_true", help="skip the L2 flush between steps")
    p.add_argument("--quick", action="store_true", help="profiling runs under ncu: 1 warm-up, no e2e, no CPU baseline")
    return p.parse_args()


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def measured_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the octave-0 blur, from the ncu
    --set full capture summarised by profiles/summarize.py (None when no capture has been summarised)."""
    try:
        with open(os.path.join(ROOT, "profiles", "r2", "traffic.json")) as f:
            d = json.load(f)
        return float(d["blur_octave0_dram_bytes_per_launch"]), d.get("source", "profiles/r2/traffic.json")
    except Exception:
        return None, None


def cpu_model():
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("model name"):
                    return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return platform.processor() or "unknown"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                 "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.lines:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for n, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def make_frames(w, h, n, first_index=0, gray=False):
    from siftmetal_b200.synth import pink_noise_bgra, pink_noise_gray

    gen = pink_noise_gray if gray else pink_noise_bgra
    # 8 distinct frames; global frame g of a batch is frame g % 8 whatever the number of shards,
    # so a sharded batch has the same content at every N (generation is FFT-bound)
    base = {}
    out = []
    for i in range(n):
        j = (first_index + i) % 8
        if j not in base:
            base[j] = gen(w, h, j)
        out.append(base[j])
    return out


def load_oracle():
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib

    return oracle_lib


def time_oracle(w, h, frames, threads, repeats):
    """Seconds per frame of the CPU oracle (detect + describe) with `threads` OpenMP threads."""
    ol = load_oracle()
    ora = ol.Oracle(w, h, threads=threads)
    used = ol.lib().oracle_max_threads()
    ora.detect(frames[0]); ora.describe()       # warm-up (page faults, OpenMP team)
    t0 = time.perf_counter()
    n = 0
    for r in range(repeats):
        f = frames[r % len(frames)]
        ora.detect(f)
        ora.describe()
        n += 1
    dt = time.perf_counter() - t0
    return dt / n, used, n


def run_reference(args, rank, world):
    """Reference arm: the reference's algorithm (C++ oracle port) on the host cores."""
    if rank != 0:
        return
    w, h, total, _ = WORKLOADS[args.workload]
    frames = make_frames(w, h, 1)
    # each step = one frame of the workload (bounded sample; 1080p ≈ 0.3 s on 16 cores)
    ol = load_oracle()
    ora = ol.Oracle(w, h, threads=0)
    cores = ol.lib().oracle_max_threads()
    for _ in range(max(1, min(args.warmup, 2))):
        ora.detect(frames[0]); ora.describe()
    steps = max(1, min(args.steps, 20 if w * h <= 1920 * 1080 else 3))
    t0 = time.perf_counter()
    for _ in range(steps):
        ora.detect(frames[0]); ora.describe()
    dt = time.perf_counter() - t0
    fps = steps / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": fps, "unit": "frames/s", "n_gpus": args.gpus,
        "steps": steps, "warmup": args.warmup, "ms_per_step": 1000 * dt / steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_string(args.workload),
                   "note": "CPU restatement (oracle/) of the reference's Metal kernels + Swift host stages; "
                           "the Swift/Metal reference itself cannot run on Linux; each step = one frame of the workload"},
        "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": cores, "kind": "port", "cpu": cpu_model(),
                         "sample": f"{steps} x one {w}x{h} frame, OpenMP {cores} threads"},
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist

    from siftmetal_b200 import Engine, _abi
    from siftmetal_b200.sharding import ShmGather, shard_range

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: there is no CPU fallback for the product path")
    torch.cuda.set_device(local)
    distributed = world > 1
    if distributed:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    w, h, total, sharded = WORKLOADS[args.workload]
    gray = args.input_format == "gray8"
    bpp = 1 if gray else 4
    if sharded:   # contiguous frame blocks per rank, no collective on the data path
        first, last = shard_range(total, rank, world)
        n_local = max(1, last - first)
        frames_per_step_job = total
    else:         # single-image workloads: one replica per GPU
        first, n_local = 0, total
        frames_per_step_job = total * world
    scaling = "strong" if sharded else "weak"
    # frames resident per execute: bounded by memory (≈ 70 B per octave pixel, 5.33 octave px / px)
    per_frame_bytes = w * h * 5.34 * 76 + w * h * 8
    chunk = args.chunk or max(1, min(n_local, int(60e9 // per_frame_bytes)))
    eng = Engine(w, h, device=local, max_batch=chunk,
                 input_format=_abi.INPUT_GRAY8 if gray else _abi.INPUT_BGRA8)
    frames = make_frames(w, h, n_local, first_index=first, gray=gray)

    # pinned host staging of this rank's frames (e2e path) — torch only as the pinned allocator
    pinned = torch.empty((n_local, h, w, 4) if not gray else (n_local, h, w), dtype=torch.uint8, pin_memory=True)
    pin_np = pinned.numpy()
    for i, f in enumerate(frames):
        pin_np[i] = f
    chunks = [(s, min(chunk, n_local - s)) for s in range(0, n_local, chunk)]
    ptr_arrays = []
    for s, n in chunks:
        ptr_arrays.append((C.c_void_p * n)(*[pin_np[s + i].ctypes.data for i in range(n)]))
    # device-resident copy of the inputs (value path)
    dev_in = pinned.to(f"cuda:{local}")
    torch.cuda.synchronize()
    frame_bytes = w * h * bpp
    flush = None if args.no_flush else torch.empty(256 << 20, dtype=torch.uint8, device=f"cuda:{local}")

    def barrier():
        if distributed:
            dist.barrier()
        torch.cuda.synchronize()

    def step_device(collect=None):
        """One step with device-resident input; returns (device ms, kernel launches)."""
        ms, launches = 0.0, 0
        for s, n in chunks:
            eng.set_device_input(dev_in.data_ptr() + s * frame_bytes, n, w * bpp, frame_bytes)
            eng.execute()
            t = eng.timings()
            ms += t["total_ms"]
            launches += t["kernel_launches"]
            if collect is not None:
                collect["stage"] += np.array([t[k + "_ms"] for k in STAGES])
                collect["blur"] += np.array(t["blur_octave0_launch_ms"])
                collect["graph"] += int(t["graph_replay"])
        return ms, launches

    # ---- warm-up (the second call of every chunk shape records its CUDA graph) ---------------------
    n_warm = 2 if args.quick else max(3, args.warmup)
    for _ in range(n_warm):
        step_device()
    res = eng.download()
    kp_per_step = int(res.keypoint_counts.sum())       # of the last chunk
    desc_per_step = int(res.descriptor_counts.sum())

    # ---- timed region: K steps, device-resident input ----------------------------------------------
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    barrier()
    t_wall0 = time.perf_counter()
    dev_ms, launches = 0.0, 0
    graph_steps = {"stage": np.zeros(6), "blur": np.zeros(5), "graph": 0}
    for _ in range(args.steps):
        if flush is not None:
            flush.zero_()                  # evict L2 (126 MB) between steps; not in the device time
            torch.cuda.synchronize()
        ms, ln = step_device(graph_steps)
        dev_ms += ms
        launches += ln
    barrier()
    wall_s = time.perf_counter() - t_wall0
    clocks = sampler.stop() if rank == 0 else None

    # ---- e2e: host buffers through the public pipelined calls, results gathered on rank 0 ----------
    graph_mode = os.environ.get("SIFTCUDA_GRAPH", "0") not in ("", "0")
    gather = None
    if distributed and sharded:
        # a recorded graph carries the result pointers: no region rotation under graph replay
        gather = ShmGather(rank, world, eng, max_frames_per_call=chunk, calls_per_step=len(chunks),
                           regions=2 if graph_mode else 4)
    e2e_counts = [0, 0]

    def finish_call(pci):
        r = eng.wait(copy=False)
        e2e_counts[0] += len(r.keypoint_columns)
        e2e_counts[1] += len(r.descriptor_columns)
        if gather is not None:
            gather.publish(r, chunks[pci][0])
            gather.call_done()                             # barrier: this call of every rank is visible on rank 0

    def pad_calls(pci):
        # ranks whose shard needs fewer calls per step still meet the others at every barrier
        if gather is not None and pci == len(chunks) - 1:
            for _ in range(gather.calls_per_step - len(chunks)):
                gather.publish_empty()
                gather.call_done()

    def step_e2e_pipelined(steps):
        """`steps` steps with two calls in flight across chunk and step boundaries."""
        seq = [(st, ci) for st in range(steps) for ci in range(len(chunks))]
        inflight = []
        for st, ci in seq:
            if len(inflight) == 2:
                pst, pci = inflight.pop(0)
                finish_call(pci)
                pad_calls(pci)
            s, n = chunks[ci]
            if gather is not None:
                gather.before_submit()                     # the slot's columns go straight into shared memory
            eng.submit_ptrs(ptr_arrays[ci], n, w * bpp)
            inflight.append((st, ci))
        while inflight:
            pst, pci = inflight.pop(0)
            finish_call(pci)
            pad_calls(pci)

    def step_e2e_sync():
        nk = nd = 0
        for (s, n), pa in zip(chunks, ptr_arrays):
            k, d = eng.detect_and_describe_ptrs(pa, n, w * bpp)
            nk += k
            nd += d
        return nk, nd

    e2e_steps = 2 if args.quick else max(5, args.steps // 2)
    sync_steps = 1 if args.quick else max(3, min(e2e_steps, 50))
    step_e2e_pipelined(2)                                    # warm-up: slot 1 allocation, its graph
    if gather is None:
        step_e2e_sync()
    barrier()
    e2e_counts[:] = [0, 0]
    t0 = time.perf_counter()
    step_e2e_pipelined(e2e_steps)
    barrier()
    e2e_s = time.perf_counter() - t0
    nk_total, nd_total = e2e_counts
    gathered = gather.summary() if gather is not None else None
    if gather is not None:
        # give the slots their own pinned blocks back for the synchronous leg
        gather.close()
        gather = None
        eng.close()
        eng = Engine(w, h, device=local, max_batch=chunk,
                     input_format=_abi.INPUT_GRAY8 if gray else _abi.INPUT_BGRA8)
        step_device()
        step_e2e_sync()
        barrier()
    t0 = time.perf_counter()
    for _ in range(sync_steps):
        step_e2e_sync()
    barrier()
    sync_s = time.perf_counter() - t0

    # ---- diagnostic region: the same steps launch by launch with per-stage events --------------------
    diag = {"stage": np.zeros(6), "blur": np.zeros(5), "graph": 0}
    diag_steps = 2 if args.quick else max(5, min(args.steps, 30))
    eng.set_stage_timing(True)
    step_device()
    diag_ms = 0.0
    for _ in range(diag_steps):
        if flush is not None:
            flush.zero_()
            torch.cuda.synchronize()
        ms, _ = step_device(diag)
        diag_ms += ms
    eng.set_stage_timing(False)
    # ... and the octave-0 blur kernels alone, back to back (20 launches each)
    isolated = None
    if not args.quick:
        step_device()
        isolated = [eng.blur_bench(s, 0, 20) for s in range(5)]

    # ---- reduce over ranks (max time) -----------------------------------------------------------------
    times = torch.tensor([dev_ms, e2e_s, wall_s, sync_s], dtype=torch.float64, device=f"cuda:{local}")
    if distributed:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    dev_ms_max, e2e_s_max, wall_s_max, sync_s_max = [float(x) for x in times.tolist()]

    if rank == 0:
        K = args.steps
        value = frames_per_step_job * K / (dev_ms_max / 1000.0)
        e2e_value = frames_per_step_job * e2e_steps / e2e_s_max
        info = eng.info
        p0 = info.octave_width[0] * info.octave_height[0]
        sum_p = sum(info.octave_width[o] * info.octave_height[o] for o in range(7))
        peak, peak_src = measured_peaks()
        # dominant streaming kernel: octave-0 blur+DoG, algorithmic bytes per launch = 12 B x P0 x frames
        launches_timed = 5 * diag_steps * len(chunks)
        blur_avg_s = float(diag["blur"].sum()) / 1000.0 / launches_timed
        bytes_per_launch = 12.0 * p0 * (n_local / len(chunks))
        achieved = bytes_per_launch / blur_avg_s / 1e9
        model_b_bytes = (bpp + 16) * w * h + 36 * sum_p   # SURVEY §8d model B, per frame
        model_bg_bytes = model_b_bytes + 36 * sum_p       # + precomputed gradients
        traffic, traffic_src = measured_traffic()
        frame_gbps = model_b_bytes * frames_per_step_job / world / (dev_ms_max / K / 1000) / 1e9
        line = {
            "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": K,
            "warmup": n_warm, "ms_per_step": dev_ms_max / K, "higher_is_better": True,
            "scaling": scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {
                "workload": workload_string(args.workload),
                "input_format": args.input_format,
                "frames_per_gpu_per_step": n_local, "resident_chunk": chunk,
                "l2": "explicit 256 MiB flush between steps" if flush is not None else
                      "no flush (per-step working set of ~0.75 GB per 1080p frame exceeds the 126 MB L2)",
                "keypoints_per_step_per_gpu": kp_per_step, "descriptors_per_step_per_gpu": desc_per_step,
                "timing": "CUDA events recorded by libsiftcuda on its own stream around each call's work "
                          f"(CUDA-graph replay in {graph_steps['graph']} of {K * len(chunks)} calls), summed",
            },
            "wall_ms_per_step": 1000 * wall_s_max / K,
            "stage_ms_per_step": {k: float(v) / diag_steps for k, v in zip(STAGES, diag["stage"])},
            "stage_timing_note": f"second region with per-stage events, {diag_steps} steps, "
                                 f"{diag_ms / diag_steps:.4f} ms per step (timed region: {dev_ms_max / K:.4f})",
            "frame_roofline": {
                "model_B_GBps": frame_gbps,
                "model_B_plus_grad_GBps": frame_gbps * model_bg_bytes / model_b_bytes,
                "model_B_bytes_per_frame": model_b_bytes, "peak_GBps": peak,
                "frac_of_measured": frame_gbps / peak, "frac_of_nominal_8TBps": frame_gbps / NOMINAL_HBM_GBS,
            },
            "roofline": {
                "bound": "hbm",
                "kernel": "blurKernel<NTAPS,64,64,256> octave 0: 5 scales (11,15,17,21,27 taps) per step; on a large "
                          "single frame each scale runs as 2 concurrent row-band launches; avg_launch_ms is the "
                          "CUDA-event time of the whole 5-scale section / 5 (one scale of the full plane) in the "
                          "launch-by-launch region; isolated_launch_ms is each kernel alone, 20 launches back to back",
                "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "frac_of_nominal_8TBps": achieved / NOMINAL_HBM_GBS,
                "traffic": traffic * (n_local / len(chunks)) if (traffic and (w, h) == (1920, 1080)) else None,
                "traffic_source": traffic_src,
                "peak_source": peak_src,
                "bytes_per_launch": bytes_per_launch, "avg_launch_ms": blur_avg_s * 1000,
                "per_tap_launch_ms": [float(x) / (diag_steps * len(chunks)) for x in diag["blur"]],
                "isolated_launch_ms": isolated,
                "isolated_frac": [bytes_per_launch / (ms / 1000.0) / 1e9 / peak for ms in isolated] if isolated else None,
            },
            "e2e": {"value": e2e_value, "unit": "frames/s",
                    "h2d_bytes_per_step": n_local * frame_bytes,
                    "d2h_bytes_per_step": int(nk_total // max(e2e_steps, 1)) * 26 + int(nd_total // max(e2e_steps, 1)) * 136
                                          + len(chunks) * (24 + 3 * 4 * (7 * chunk + 1)),
                    "steps": e2e_steps,
                    "timing": "wall clock around sift_submit / sift_wait with two calls in flight (on two pipelines of "
                              "the context when it is small: see sift_get_info device_bytes), pinned host frames, "
                              "result columns in pinned host memory" + (", every call's shards visible on rank 0 in frame order "
                                                                        "(kernels store into shared-memory segments)"
                                                                        if gathered is not None else ""),
                    "ms_per_step": 1000.0 * e2e_s_max / e2e_steps,
                    "sync_call": {"value": frames_per_step_job * sync_steps / sync_s_max, "steps": sync_steps,
                                  "ms_per_step": 1000.0 * sync_s_max / sync_steps,
                                  "timing": "wall clock around sift_detect_and_describe_batch, one call at a time"},
                    "gather": gathered},
            "gpu_launches": int(launches),
            "clocks": clocks,
        }
        if not (args.no_cpu_baseline or args.quick):
            reps = 8 if w * h <= 1920 * 1080 else 1
            bgra = frames if not gray else make_frames(w, h, min(n_local, 4), first_index=first)
            spf, cores, n = time_oracle(w, h, bgra[:4], 0, reps)
            line["cpu_baseline"] = {"value": 1.0 / spf, "unit": "frames/s", "cores": cores, "kind": "port",
                                    "cpu": cpu_model(),
                                    "sample": f"{n} x one {w}x{h} frame of the same workload, OpenMP {cores} threads"}
            if w * h <= 1920 * 1080:
                spf1, _, n1 = time_oracle(w, h, bgra[:2], 1, 2)
                line["cpu_baseline"]["single_thread"] = {"value": 1.0 / spf1, "cores": 1,
                                                         "sample": f"{n1} x one {w}x{h} frame, 1 thread"}
        print(json.dumps(line))
    if gather is not None:
        gather.close()
    eng.close()
    if distributed:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
