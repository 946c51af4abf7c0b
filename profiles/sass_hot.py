"""Dynamic instruction breakdown of a kernel from an ncu report captured with --import-source on:
share of executed warp instructions per region between barriers, per opcode, and the hottest
SASS lines. Usage: python profiles/sass_hot.py report.ncu-rep [kernel-index] [top-n]"""
import csv, subprocess, sys
rep = sys.argv[1]
which = int(sys.argv[2]) if len(sys.argv) > 2 else 0
topn = int(sys.argv[3]) if len(sys.argv) > 3 else 25
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
kern, cur = [], None
for r in csv.reader(txt.splitlines()):
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "rows": []}
        kern.append(cur)
    elif r and r[0] == "Address":
        cur["hdr"] = r
    elif cur is not None and len(r) > 5:
        cur["rows"].append(r)
k = kern[which]
h = k["hdr"]
iE, iS, iSrc = h.index("Instructions Executed"), h.index("# Samples"), h.index("Source")
tot = sum(int(r[iE]) for r in k["rows"])
stot = sum(int(r[iS]) for r in k["rows"])
print(k["name"][:90], "| warp instr", tot, "| sass lines", len(k["rows"]), "| samples", stot)
mix = {}
for r in k["rows"]:
    t = r[iSrc].split()
    op = (t[1] if t[0].startswith("@") else t[0]).split(".")[0]
    mix[op] = mix.get(op, 0) + int(r[iE])
print("opcode mix:", ", ".join(f"{op} {100 * e / tot:.1f}%" for op, e in sorted(mix.items(), key=lambda x: -x[1])[:16]))
print(f"hottest {topn} lines by stall samples:")
for i, r in sorted(enumerate(k["rows"]), key=lambda x: -int(x[1][iS]))[:topn]:
    print(f"  {i:5d} samples {int(r[iS]):6d} ({100 * int(r[iS]) / max(stot, 1):4.1f}%) exec {int(r[iE]):9d}  {r[iSrc].strip()[:80]}")
