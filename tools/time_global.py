"""Times the global scheme (oakb200_global_analysis_dev) on a B200: state resident in HBM, synthetic data of the
C3 shape by default (n = 3e7 rows, N = 64, m = 1e6).  The apply is the dominant kernel and is HBM-bound (reads and
writes the state once: 16 n N bytes); prints its achieved GB/s against the measured copy bandwidth in
MEASURED_PEAKS.json.  Round-2 tooling; run on the GPU box:  python tools/time_global.py [--n 30000000 --N 64 --m 1000000]"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import oak_b200
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=30_000_000)
    ap.add_argument("--N", type=int, default=64)
    ap.add_argument("--m", type=int, default=1_000_000)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--apply-kernel", type=int, default=0)
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    g = torch.Generator(device=dev).manual_seed(20261017)
    Sf = torch.randn((a.N, a.n), dtype=torch.float64, device=dev, generator=g)
    Sf -= Sf.mean(dim=0, keepdim=True)            # anomalies: the members sum to zero in every row
    HSf = torch.randn((a.N, a.m), dtype=torch.float64, device=dev, generator=g)
    HSf -= HSf.mean(dim=0, keepdim=True)
    xf = torch.randn(a.n, dtype=torch.float64, device=dev, generator=g)
    Hxf = torch.randn(a.m, dtype=torch.float64, device=dev, generator=g)
    yo = Hxf + 0.1 * torch.randn(a.m, dtype=torch.float64, device=dev, generator=g)
    R = torch.full((a.m,), 0.25, dtype=torch.float64, device=dev)
    xa = torch.empty_like(xf)
    Sa = torch.empty_like(Sf)
    with oak_b200.Handle(0, apply_kernel=a.apply_kernel) as h:
        ms = []
        for it in range(a.steps + 2):
            st = h.global_analysis_dev(xf, Hxf, yo, Sf, HSf, R, xa, Sa)
            if it >= 2:
                ms.append(st["ms_total"])
    t = sorted(ms)[len(ms) // 2]
    byt = 16.0 * a.n * a.N + 16.0 * a.n + 8.0 * a.m * a.N
    peak = None
    try:
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    print(json.dumps({"metric": "global analysis, state resident", "ms": t, "rows": a.n, "N": a.N, "m": a.m, "apply_kernel": a.apply_kernel,
                      "algorithmic_GB": byt / 1e9, "achieved_GBps": byt / 1e9 / (t * 1e-3), "measured_peaks": peak,
                      "colsum_check": float(Sa.sum(dim=0).abs().max())}))


if __name__ == "__main__":
    main()
