"""Per-source-line hot spots of an .ncu-rep captured with --import-source on (read here, no GPU):
    python tools/ncu_lines.py gpurun_out/prof.ncu-rep [top]
Prints, per source line, executed warp instructions and stall samples (share of the kernel)."""
import csv, subprocess, sys
from collections import defaultdict
rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
agg = defaultdict(lambda: [0, 0, ""])
fname, hdr = "", None
cur = None
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        fname = r[1].split("/")[-1]; continue
    if r[0] == "Line No":
        hdr = {n: i for i, n in enumerate(r)}; continue
    if hdr is None or len(r) < 10:
        continue
    if r[0] != "":
        cur = (fname, int(r[0])); agg[cur][2] = r[1].strip()[:110]
    else:
        try:
            agg[cur][0] += int(r[hdr["Instructions Executed"]]); agg[cur][1] += int(r[hdr["# Samples"]])
        except Exception:
            pass
ti = sum(v[0] for v in agg.values()) or 1; ts = sum(v[1] for v in agg.values()) or 1
print(f"total warp-instr {ti/1e6:.1f}M samples {ts}")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    print(f"{k[0]}:{k[1]:4d} instr {100*v[0]/ti:5.1f}% samples {100*v[1]/ts:5.1f}%  {v[2]}")
