"""Builds profiles/<round>/SUMMARY.md from gpurun_out/: launch list shares, ncu key metrics of
the hot kernels (incl. DRAM traffic per launch), bench JSON."""
import collections, csv, io, json, os, re, subprocess, sys
if len(sys.argv) not in (2, 3) or sys.argv[1].startswith("-"):
    sys.exit("usage: python profiles/summarize.py <output dir, e.g. profiles/r2> [input dir, default gpurun_out]")
out_dir = sys.argv[1]
g = sys.argv[2] if len(sys.argv) == 3 else "gpurun_out"
lines = []
P = lines.append
bench = json.loads([l for l in open(f"{g}/bench_1080p.json") if l.startswith("{")][-1])
P(f"# Profile summary ({out_dir})\n")
P("Workload: 1 x 1920x1080 1/f-noise frame per step, 1 B200. All ncu numbers are cold-cache, serialised launches "
  "(`--clock-control none`); compare SHARES with the event-timed bench, not absolutes.\n")
P("## bench.py (CUDA events, warm, L2 flushed between steps)\n")
P("```json\n" + json.dumps({k: bench[k] for k in ("value", "unit", "ms_per_step", "stage_ms_per_step", "stage_timing_note", "frame_roofline", "roofline", "e2e", "cpu_baseline", "gpu_launches", "clocks") if k in bench}, indent=1) + "\n```\n")
# launch list
rows = list(csv.DictReader([l for l in open(f"{g}/launches.csv") if not l.startswith("==")]))
names = [(r["Kernel Name"], float(r["Metric Value"])) for r in rows]
idx = [i for i, n in enumerate(names) if "grayKernel" in n[0] or "grayUpsample2xKernel" in n[0]]
step = names[idx[1]:idx[2]] if len(idx) > 2 else names[idx[-1]:]
agg = collections.OrderedDict()
for n, t in step:
    if "at::" in n:
        continue   # torch's L2-flush fill between steps: not part of the device-timed region
    nm = re.sub(r"\(.*", "", n).replace("void ", "").replace("sift::", "")
    agg.setdefault(nm, []).append(t / 1000)
tot = sum(sum(v) for v in agg.values())
P("## ncu launch list of one step (`--metrics gpu__time_duration.sum`)\n")
P("| kernel | launches | total µs | share |\n|---|---|---|---|")
for k, v in agg.items():
    P(f"| `{k}` | {len(v)} | {sum(v):.1f} | {100 * sum(v) / tot:.1f} % |")
P(f"| **sum** | {sum(len(v) for v in agg.values())} | {tot:.1f} | 100 % |\n")
KEYS = [("time µs", "gpu__time_duration.sum"), ("DRAM read MB", "dram__bytes_read.sum"), ("DRAM write MB", "dram__bytes_write.sum"),
        ("DRAM % of peak", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"), ("issue active %", "smsp__issue_active.avg.pct_of_peak_sustained_active"),
        ("warps active %", "sm__warps_active.avg.pct_of_peak_sustained_active"), ("FMA pipe %", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active"),
        ("regs", "launch__registers_per_thread"), ("warp inst", "smsp__inst_executed.sum"), ("L2 hit %", "lts__t_sector_hit_rate.pct")]
P("## ncu --set full, per launch\n")
P("| kernel | " + " | ".join(k for k, _ in KEYS) + " |\n|---|" + "---|" * len(KEYS))
blur_traffic = []
for rep in ("prof_grayUpsample2xKernel", "prof_blur", "prof_gradientKernel", "prof_extremaMaskTmaKernel", "prof_extremaMaskKernel",
            "prof_tailOctavesKernel", "prof_orientationKernel", "prof_descriptorKernel", "prof_matchKernel"):
    path = f"{g}/{rep}.ncu-rep"
    if not os.path.exists(path):
        continue
    txt = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rr = list(csv.reader(io.StringIO(txt)))
    hdr = rr[0]
    ix = {h: i for i, h in enumerate(hdr)}
    for r in rr[2:]:
        name = re.sub(r"\(.*", "", r[ix["Kernel Name"]]).replace("void ", "")
        vals = []
        for _, key in KEYS:
            v = r[ix[key]] if key in ix else ""
            try:
                v = f"{float(v):.4g}"
            except ValueError:
                pass
            vals.append(v)
        P(f"| `{name[:44]}` grid {r[ix['Grid Size']]} | " + " | ".join(vals) + " |")
        if rep == "prof_blur":
            def mb(key):
                v, u = float(r[ix[key]]), rr[1][ix[key]]
                return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
            blur_traffic.append(mb("dram__bytes_read.sum") + mb("dram__bytes_write.sum"))
P("")
if blur_traffic:
    # sidecar read by bench.py: measured DRAM traffic per octave-0 blur launch (mean of the 5 scales)
    with open(os.path.join(out_dir, "traffic.json"), "w") as f:
        json.dump({"blur_octave0_dram_bytes_per_launch": sum(blur_traffic) / len(blur_traffic),
                   "per_scale": blur_traffic,
                   "source": f"{out_dir}/SUMMARY.md: ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum, "
                             "mean of the 5 octave-0 launches (1080p, SIFTCUDA_BANDS=1, cold cache)"}, f, indent=1)
os.makedirs(out_dir, exist_ok=True)
open(os.path.join(out_dir, "SUMMARY.md"), "w").write("\n".join(lines) + "\n")
print("\n".join(lines)[:6000])
