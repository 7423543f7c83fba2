"""Summarises an .ncu-rep (read here, no GPU): key raw metrics, stall reasons, SASS op mix and
per-segment hot spots.   python tools/ncu_summary.py gpurun_out/prof.ncu-rep [segment_size]"""
import csv, subprocess, sys
from collections import Counter
rep = sys.argv[1]
seg = int(sys.argv[2]) if len(sys.argv) > 2 else 400
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
KEYS = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "launch__grid_size", "launch__block_size",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]
for d in rows[2:]:
    print("==", d[hdr.index("Kernel Name")][:90] if "Kernel Name" in hdr else "")
    for k in KEYS:
        if k in hdr:
            print(f"  {k:75s} {d[hdr.index(k)]:>16s} {units[hdr.index(k)]}")
    st = [(float(d[i]), h) for i, h in enumerate(hdr) if h.startswith("smsp__pcsamp_warps_issue_stalled") and "not_issued" not in h and d[i] not in ("", "n/a")]
    tot = sum(v for v, _ in st) or 1
    print("  stalls:", ", ".join(f"{h.replace('smsp__pcsamp_warps_issue_stalled_', '')} {100 * v / tot:.0f}%" for v, h in sorted(st, reverse=True)[:9]))
sass = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(sass.splitlines()))
hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
for k, h0 in enumerate(hi[:1]):
    h = rows[h0]; c = {n: i for i, n in enumerate(h)}
    end = hi[k + 1] - 1 if k + 1 < len(hi) else len(rows)
    items = []
    for r in rows[h0 + 1:end]:
        try:
            items.append((r[c["Source"]], int(r[c["# Samples"]]), int(r[c["Instructions Executed"]])))
        except Exception:
            pass
    ti = sum(i[2] for i in items) or 1; ts = sum(i[1] for i in items) or 1
    mix = Counter(); smp = Counter()
    for src, s, ie in items:
        t = src.split(); op = (t[1] if t[0].startswith("@") else t[0]).split(".")[0]
        mix[op] += ie; smp[op] += s
    print(f"SASS instructions {len(items)}, executed warp-instr {ti / 1e6:.1f}M, samples {ts}")
    print("  op mix:", ", ".join(f"{o} {100 * v / ti:.1f}%/{100 * smp[o] / ts:.0f}%s" for o, v in mix.most_common(16)))
    for a in range(0, len(items), seg):
        part = items[a:a + seg]
        ie = sum(x[2] for x in part); s = sum(x[1] for x in part)
        m = Counter()
        for src, ss, ii in part:
            t = src.split(); op = (t[1] if t[0].startswith("@") else t[0]).split(".")[0]; m[op] += ii
        print(f"  [{a:5d}] exec {100 * ie / ti:5.1f}% samples {100 * s / ts:5.1f}%  " + ", ".join(f"{o}:{100 * v / max(ie, 1):.0f}" for o, v in m.most_common(6)))
