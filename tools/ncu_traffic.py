"""Writes profiles/ncu_traffic.json from ncu --set full captures (read here, no GPU needed):
dram__bytes_read.sum + dram__bytes_write.sum per launch / zones per launch for every kernel stage bench.py reports,
with the kernel options and the git hash of the build the captures were taken from.

    python tools/ncu_traffic.py TAG [zones_per_launch]      # reads gpurun_out/TAG_prof_<kernel>.ncu-rep
"""
import csv, json, os, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
STAGES = {"k_gram": "k_gram_mma", "k_tridiag": "k_tridiag_warp", "k_tql+k_tvec": "k_tvec", "k_apply": "k_apply_tma"}


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, d = rows[0], rows[1], rows[2]
    g = lambda k: (float(d[hdr.index(k)].replace(",", "")), units[hdr.index(k)])
    def bytes_of(k):
        v, u = g(k)
        return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
    return {"grid": g("launch__grid_size")[0], "bytes": bytes_of("dram__bytes_read.sum") + bytes_of("dram__bytes_write.sum"),
            "us": g("gpu__time_duration.sum")[0], "kernel": d[hdr.index("Kernel Name")][:80]}


res, detail = {}, {}
for stage, kn in STAGES.items():
    rep = os.path.join(ROOT, "gpurun_out", f"{tag}_prof_{kn}.ncu-rep")
    if not os.path.exists(rep):
        continue
    r = raw(rep)
    zones = r["grid"] * (32 if kn == "k_tql" else 1)
    res[stage] = r["bytes"] / zones
    detail[stage] = dict(r, zones_per_launch=zones)
head = subprocess.run(["git", "rev-parse", "--short", "HEAD"], capture_output=True, text=True, cwd=ROOT).stdout.strip()
json.dump({"git": head, "captures": f"gpurun_out/{tag}_prof_*.ncu-rep (ncu --set full, bench.py --nx 300 --ny 300 --nobs 90000)",
           "options": {"N": 64, "eig_kernel": 4, "gram_kernel": 1, "fuse_apply": 0, "tvec_split": 0, "apply_kernel": 0},
           "bytes_per_zone": res, "detail": detail}, open(os.path.join(ROOT, "profiles", "ncu_traffic.json"), "w"), indent=1)
print(json.dumps(res))
