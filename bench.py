#!/usr/bin/env python
"""bench.py — local-analysis throughput (grid columns analysed per second) on 1/2/4/8 B200.

Workload (BASELINE.json configs[2], "C3"): synthetic 3-D ocean 1000 x 1000 x 30, N = 64 members,
1e6 observations with diagonal R, Gaussian localisation corrLen 4 km / cut-off 8 km, Cartesian metric.
A step = one complete local analysis (locAnalysis, rrsqrt.F90:433) of every water column.
Multi-GPU: strong scaling — the 1e6 columns are split in contiguous zone ranges (parall.F90:176-177),
each rank analyses its slab with its observation halo, one NCCL all-gather reassembles the analysed
anomalies; `value` = all columns / max-over-ranks time.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--nx .. --ny .. --nz .. --N .. --nobs ..]
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

SEED = 20261017


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="c3", choices=["c3", "c4", "c5", "hgen"],
                    help="c3 = the headline workload (BASELINE configs[2]); c4 = 2000x2000x50, N=128, obs at every surface "
                         "point, analysed slab by slab in place (configs[3]); c5 = 4-D 256x256x10x4, N=64, inflation + log "
                         "anamorphosis through the ensemble entry point (configs[4]); hgen = interpolation weights of the observation "
                         "operator (batched cinterp) for 1e6 observations on the C3 grid (SURVEY 8f rank 3)")
    ap.add_argument("--slab-rows", type=int, default=50, help="c4: grid rows (of nx zones) per resident slab")
    ap.add_argument("--max-slabs", type=int, default=0, help="c4: analyse only this many slabs per rank (0 = all) and say so")
    ap.add_argument("--nx", type=int, default=1000)
    ap.add_argument("--ny", type=int, default=1000)
    ap.add_argument("--nz", type=int, default=30)
    ap.add_argument("--N", type=int, default=64)
    ap.add_argument("--nobs", dest="m", type=int, default=1000000)
    ap.add_argument("--corr", type=float, default=4000.0)
    ap.add_argument("--maxlen", type=float, default=8000.0)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-pageable", action="store_true", help="skip the e2e legs on pageable host arrays")
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--cpu-seconds", type=float, default=15.0)
    ap.add_argument("--eig-kernel", type=int, default=4)
    ap.add_argument("--gram-kernel", type=int, default=1, help="1 = mma.m8n8k4 tiles, 4 warps per zone (default), 2 = 2 warps per zone, 3 / 4 = same with 32-candidate chunks, 0 = DFMA register tiles")
    ap.add_argument("--fuse-apply", type=int, default=0, help="1 = the transform kernel updates the zone rows from the factored transform (no T, no k_apply)")
    ap.add_argument("--apply-kernel", type=int, default=0, help="1 = k_apply on mma.m8n8k4 tiles (zones the fused transform kernel leaves over; all zones when --fuse-apply 0)")
    ap.add_argument("--tvec-split", type=int, default=0, help="1 = eigenvector kernel as two kernels (vectors of T | back-transformation and the rest)")
    ap.add_argument("--jacobi-tol", type=float, default=0.0, help="experiment: override the Jacobi stopping tolerance")
    ap.add_argument("--sync-phases", action="store_true", help="N>1: blocking library calls instead of the asynchronous pipeline")
    ap.add_argument("--phases", type=int, default=0, help="pipeline phases for N>1 (0: 1 with the fused gather, 4 with --nccl-gather)")
    ap.add_argument("--peer-mode", type=int, default=1, help="fused gather: 1 = copy engines push each finished batch, 0 = stores of the apply kernel")
    ap.add_argument("--push-pieces", type=int, default=1, help="fused gather, peer mode 1: apply + push of every batch in this many pieces")
    ap.add_argument("--gather", default="multicast", choices=["multicast", "peer"],
                    help="N>1, fused gather: multicast = one multimem store per row, replicated by NVSwitch into every rank's "
                         "result array (falls back to peer when unsupported); peer = one store / copy per destination")
    ap.add_argument("--push-kernel", type=int, default=1, help="fused gather: 1 = rows pushed by a small kernel on a side stream (SM stores over NVLink) instead of copy-engine copies")
    ap.add_argument("--push-ctas", type=int, default=24)
    ap.add_argument("--nccl-gather", action="store_true", help="N>1: reassemble with NCCL all-gathers instead of the fused peer stores of the apply kernel")
    return ap.parse_args()


def workload_name(a):
    return (f"synthetic 3D ocean {a.nx}x{a.ny}x{a.nz}, N={a.N}, {a.m} obs, diagonal R, local ETKF "
            f"(gaussian corrLen {a.corr:g} m, cut-off {a.maxlen:g} m, cartesian)")


# --------------------------------------------------------------------------------------------------
# clocks sampled during the timed region
# --------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.index)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); pw.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


# --------------------------------------------------------------------------------------------------
# synthetic data of one rank, generated on the device
# --------------------------------------------------------------------------------------------------
def build_rank_data(a, rank, world, dev, first=None, obs_np=None):
    """Synthetic slab of one rank (in one phase when `first` is given), generated on the device."""
    import torch
    from oak_b200 import synthetic as S
    from oak_b200.dist import ShardPlan
    g = S.Grid(a.nx, a.ny, a.nz)
    if obs_np is None:
        obs_np = S.observations(np, g, a.m, SEED)
    zones = np.arange(g.nzones, dtype=np.int64)
    zx, zy = g.zone_xy(np, zones)
    zs = np.full(g.nzones, a.nz, dtype=np.int32)
    plan = ShardPlan(zs, zx, zy, a.corr, a.maxlen, obs_np["ox"], obs_np["oy"], rank, world, first=first)
    n_loc = plan.r1 - plan.r0
    Sf = torch.empty((a.N, n_loc), dtype=torch.float64, device=dev)
    xf = torch.empty(n_loc, dtype=torch.float64, device=dev)
    step = 1 << 20
    for s in range(0, n_loc, step):
        e = min(n_loc, s + step)
        rows = torch.arange(plan.r0 + s, plan.r0 + e, dtype=torch.int64, device=dev)
        E = S.ensemble_rows(torch, g, rows, a.N, SEED)
        mean, anom = S.anomalies(torch, E)
        xf[s:e] = mean
        Sf[:, s:e] = anom
        del E, anom, mean, rows
    oi = torch.from_numpy(plan.obs_idx).to(dev)
    obs_t = S.observations(torch, g, a.m, SEED) if dev.type == "cpu" else {
        k: torch.from_numpy(np.ascontiguousarray(v)).to(dev) for k, v in obs_np.items()}
    sub = {k: (v[:, oi] if v.dim() > 1 else v[oi]) for k, v in obs_t.items()}
    mh = int(oi.numel())
    HE = torch.empty((a.N, mh), dtype=torch.float64, device=dev)
    yo = torch.empty(mh, dtype=torch.float64, device=dev)
    ostep = 1 << 18
    for s in range(0, mh, ostep):
        e = min(mh, s + ostep)
        part = {k: (v[:, s:e] if v.dim() > 1 else v[s:e]) for k, v in sub.items()}
        he, y = S.obs_space(torch, g, part, a.N, SEED)
        HE[:, s:e] = he
        yo[s:e] = y
    Hxf, HSf = S.anomalies(torch, HE)
    del HE
    return dict(grid=g, plan=plan, Sf=Sf, xf=xf, HSf=HSf.contiguous(), Hxf=Hxf.contiguous(), yo=yo,
                var=sub["var"].contiguous(), ox=obs_np["ox"][plan.obs_idx], oy=obs_np["oy"][plan.obs_idx])


def load_traffic(a):
    """{stage: DRAM bytes per zone} from profiles/ncu_traffic.json, for the kernel options of this run only."""
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as fh:
            t = json.load(fh)
        opt = t.get("options", {})
        same = (opt.get("N") == a.N and opt.get("gram_kernel") == a.gram_kernel and opt.get("fuse_apply") == a.fuse_apply
                and opt.get("tvec_split") == a.tvec_split and opt.get("apply_kernel") == a.apply_kernel
                and opt.get("eig_kernel") == a.eig_kernel)
        return {k: float(v) for k, v in t.get("bytes_per_zone", {}).items()} if same else {}
    except Exception:
        return {}


def flops_per_zone(N, nz, mloc_mean, cand_mean):
    """algorithmic work per zone: Gram + eigendecomposition (LAPACK count) + transform +
    amplitudes + apply + selection"""
    return (2 * N * N * mloc_mean + 9 * N ** 3 + 2 * N ** 3 + 2 * mloc_mean * N + 4 * N * N +
            2 * nz * N * N + 2 * nz * N + 25 * cand_mean)


# --------------------------------------------------------------------------------------------------
# CPU baseline: the oracle (port of the reference's CPU path: O(m) scan per zone, dgemm, dsyev) on a sample
# --------------------------------------------------------------------------------------------------
def cpu_sample_problem(a, d, nsample):
    import torch
    g, plan = d["grid"], d["plan"]
    rng = np.random.default_rng(1)
    zl = np.sort(rng.choice(plan.z1 - plan.z0, size=min(nsample, plan.z1 - plan.z0), replace=False))
    rows = (zl[:, None] * a.nz + np.arange(a.nz)[None, :]).ravel()
    rt = torch.from_numpy(rows).to(d["Sf"].device)
    Sf = np.asfortranarray(d["Sf"][:, rt].cpu().numpy().T)
    xf = d["xf"][rt].cpu().numpy()
    zx, zy = g.zone_xy(np, (plan.z0 + zl).astype(np.int64))
    return dict(zl=zl, Sf=Sf, xf=xf, zx=zx, zy=zy, zs=np.full(zl.size, a.nz, np.int32))


def run_oracle_sample(a, host, sp, count, cellgrid=False, keep=None):
    """The oracle on the first `count` sampled columns; `keep` (a dict) receives its xa, Sa (rows of those columns)."""
    import oracle
    obs = oracle.make_obs(host["m"], obsx=host["ox"], obsy=host["oy"])
    k = count * a.nz
    t0 = time.perf_counter()
    r = (oracle.loc_analysis_cellgrid if cellgrid else oracle.loc_analysis)(sp["zs"][:count], dict(x=sp["zx"][:count], y=sp["zy"][:count]), a.corr, a.maxlen, obs,
                        sp["xf"][:k], host["Hxf"], host["yo"], sp["Sf"][:k], host["HSf"], host["var"])
    t = time.perf_counter() - t0
    if keep is not None:
        keep.update(count=count, xa=r[0], Sa=r[1], scan=not cellgrid)
    return t


def parity_vs_oracle(a, d, sp, kept):
    """Relative error (max norm over the sampled columns, as the parity tests measure it) of the GPU result of THIS
    run (d["xa"], d["Sa"], left there by the timed steps) against what the oracle computed for the same columns."""
    import torch
    count = kept["count"]
    rows = (sp["zl"][:count, None] * a.nz + np.arange(a.nz)[None, :]).ravel()
    rt = torch.from_numpy(rows).to(d["Sa"].device)
    Sg = d["Sa"][:, rt].cpu().numpy().T
    xg = d["xa"][rt].cpu().numpy()
    eS = float(np.abs(Sg - kept["Sa"]).max() / np.abs(kept["Sa"]).max())
    ex = float(np.abs(xg - kept["xa"]).max() / np.abs(kept["xa"]).max())
    return {"cols": int(count), "max_rel_Sa": eS, "max_rel_xa": ex, "tol": 1e-9, "ok": bool(eS < 1e-9 and ex < 1e-9),
            "against": "oracle port (dsyev/dgemm), " + ("O(m) scan" if kept["scan"] else "cell-grid selection") +
                       " on randomly sampled columns of this run's workload"}


def host_obs_arrays(d):
    return dict(m=int(d["yo"].numel()), ox=d["ox"], oy=d["oy"], Hxf=d["Hxf"].cpu().numpy(), yo=d["yo"].cpu().numpy(),
                var=d["var"].cpu().numpy(), HSf=np.asfortranarray(d["HSf"].cpu().numpy().T))


def cpu_baseline(a, d, seconds):
    import oracle
    host = host_obs_arrays(d)
    cores = oracle.set_threads(0)   # every online processor, whatever OMP_NUM_THREADS the launcher exported
    sp = cpu_sample_problem(a, d, 200000)
    probe = max(cores * 2, 16)
    t = run_oracle_sample(a, host, sp, probe)
    count = int(min(sp["zl"].size, max(probe, probe * seconds / max(t, 1e-6))))
    kept = {}
    t = run_oracle_sample(a, host, sp, count, keep=kept)
    out = {"value": count / t, "unit": "columns/s", "cores": cores, "kind": "port",
           "sample": f"{count} random columns of the same workload (all {host['m']} observations scanned per "
                     f"column as assimilation.F90:3745-3757 does, OpenBLAS dgemm/dsyev, OpenMP dynamic over "
                     f"columns), {t:.1f} s"}
    # the "fair" figure of SURVEY 8d(ii): the same port with a CPU cell grid in front of the exact predicate
    # (what the GPU path does), about a third of the time budget
    try:
        tf = run_oracle_sample(a, host, sp, probe, cellgrid=True)
        cf = int(min(sp["zl"].size, max(probe, probe * (seconds / 3.0) / max(tf, 1e-6))))
        tf = run_oracle_sample(a, host, sp, cf, cellgrid=True)
        out["with_cell_grid"] = {"value": cf / tf, "unit": "columns/s", "cores": cores,
                                 "sample": f"{cf} random columns, cell-grid selection instead of the O(m) scan, {tf:.1f} s"}
    except Exception as e:  # reported, never fatal for the bench line
        out["with_cell_grid"] = {"value": None, "note": str(e)[:120]}
    return out, (host, sp, kept)


def quick_parity(a, d, ncols=2000):
    """Parity of this run against the oracle when the CPU baseline leg is off (N > 1, --no-cpu): the cell-grid
    variant of the port (equal to the scan, tests/test_oracle_golden.py) on a small random sample."""
    import oracle
    oracle.set_threads(0)
    host = host_obs_arrays(d)
    sp = cpu_sample_problem(a, d, ncols)
    kept = {}
    run_oracle_sample(a, host, sp, sp["zl"].size, cellgrid=True, keep=kept)
    return parity_vs_oracle(a, d, sp, kept)


# --------------------------------------------------------------------------------------------------
def main():
    a = parse()
    import torch
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))

    if a.impl == "reference":
        return reference_arm(a, rank, world)
    if a.config == "c4":
        return run_c4(a, rank, world, local)
    if a.config == "c5":
        return run_c5(a, rank, world, local)
    if a.config == "hgen":
        return run_hgen(a, rank, world, local)

    import torch.distributed as dist
    import oak_b200
    from oak_b200.dist import allgather_slabs
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # NCCL_DEBUG is the caller's (the driver reads NCCL's rank lines from it); NCCL logs to stderr / its own file,
        # the one JSON line goes to stdout
        dist.init_process_group("nccl", device_id=dev)
    from oak_b200 import synthetic as S
    from oak_b200.dist import phase_ranges
    g = S.Grid(a.nx, a.ny, a.nz)
    fused = world > 1 and not a.nccl_gather
    nphase = a.phases if a.phases > 0 else (4 if (world > 1 and not fused) else 1)
    obs_np = S.observations(np, g, a.m, SEED)
    # phase j of rank p = a contiguous zone range; the all-gather of phase j overlaps the analysis of j+1
    phases = []
    ranges = phase_ranges(g.nzones, world, nphase)
    use_async = world > 1 and not a.sync_phases and nphase > 1
    for j, first in enumerate(ranges):
        d = build_rank_data(a, rank, world, dev, first=first, obs_np=obs_np)
        plan = d["plan"]
        h = oak_b200.Handle(local, eig_kernel=a.eig_kernel, gram_kernel=a.gram_kernel, fuse_apply=a.fuse_apply, tvec_split=a.tvec_split, apply_kernel=a.apply_kernel)
        if os.environ.get("OAK_B200_FIXED_SWEEPS"):  # kernel timing experiments only (tools/ab.py)
            h.set_option("fixed_sweeps", float(os.environ["OAK_B200_FIXED_SWEEPS"]))
        if os.environ.get("OAK_B200_ZB") and os.environ["OAK_B200_ZB"] != "0":  # pipeline experiments (tools/ab.py)
            h.set_option("zones_per_batch", float(os.environ["OAK_B200_ZB"]))
        for kv in os.environ.get("OAK_B200_OPTIONS", "").split(","):   # experiments: key=value,... library options
            if "=" in kv:
                h.set_option(kv.split("=")[0], float(kv.split("=")[1]))
        if a.jacobi_tol > 0:
            h.set_option("jacobi_tol", a.jacobi_tol)
        h.set_zones(plan.zoneSize, zone_x=plan.zx, zone_y=plan.zy, corrLen=plan.corrLen, maxLen=plan.maxLen,
                    loctype=1, metrictype=0, weightfun=0)
        h.set_observations(obs_x=d["ox"], obs_y=d["oy"])
        if use_async:
            # all phases are enqueued at once; the block scheduler serves the earlier phase first, later phases
            # fill its tail; the caller's stream waits for each phase, so the all-gathers are stream-ordered
            h.set_option("stream_priority", -(len(ranges) - 1 - j))
            h.set_option("order_after_caller", 0)
            h.set_option("async", 1)
        d["h"] = h
        d["xa"] = torch.empty(plan.r1 - plan.r0, dtype=torch.float64, device=dev)
        d["Sa"] = torch.empty_like(d["Sf"])
        phases.append(d)
    d, plan, h = phases[0], phases[0]["plan"], phases[0]["h"]
    n_loc = sum(p["plan"].r1 - p["plan"].r0 for p in phases)
    z_rank = sum(p["plan"].z1 - p["plan"].z0 for p in phases)
    Sa_full, peer, fused_note, use_mc = None, None, None, False
    if fused:
        # fused all-gather: every finished batch of this rank goes straight into the result arrays of all ranks
        # (CUDA-IPC mappings over NVLink).  The set-up is collective: if ANY rank cannot map its peers, all ranks
        # fall back together to the NCCL all-gather (same kernels otherwise), so nobody waits in a collective alone.
        from oak_b200.dist import PeerResult, MulticastResult
        err = None
        use_mc = False
        if a.gather == "multicast":
            # collective attempt: all ranks succeed or all fall back to the peer flavour
            try:
                mcres = MulticastResult(dist, a.N, plan.n, rank, world, dev)
                okm = 1.0
            except Exception as e:
                mcres, okm, mc_err = None, 0.0, str(e)[:80]
            okt = torch.tensor([okm], device=dev)
            dist.all_reduce(okt, op=dist.ReduceOp.MIN)
            use_mc = okt.item() == 1.0
            if not use_mc and mcres is not None:
                mcres.close()
                mcres = None
        if use_mc:
            peer = mcres
        else:
          try:
            peer = PeerResult(dist, h, a.N, plan.n, rank, world, dev)
            err = peer.failed
          except Exception as e:
            err, peer = str(e)[:80], None
        okf = torch.tensor([0.0 if err else 1.0], device=dev)
        dist.all_reduce(okf, op=dist.ReduceOp.MIN)
        if okf.item() == 1.0 and use_mc:
            mS, mx = peer.multicast_pointers()
            for p in phases:
                p["h"].set_option("push_ctas", a.push_ctas)
                p["h"].set_multicast_output(mS, mx, plan.n, p["plan"].r0)
            Sa_full = peer.Sa
        elif okf.item() == 1.0:
            Sp, xp = peer.destinations()
            for p in phases:
                p["h"].set_option("peer_mode", a.peer_mode)
                p["h"].set_option("push_pieces", a.push_pieces)
                p["h"].set_option("push_kernel", a.push_kernel)
                p["h"].set_option("push_ctas", a.push_ctas)
                p["h"].set_peer_outputs(Sp, xp, plan.n, p["plan"].r0)
            Sa_full = peer.Sa
        else:
            fused, fused_note = False, "peer mapping failed on some rank (%s): NCCL all-gather, single phase" % (err or "another rank",)
            if peer is not None:
                peer.Sa = peer.xa = None
                peer.close_mappings()
            dist.barrier()
            if peer is not None:
                peer.free()
                peer = None
    if world > 1 and Sa_full is None:
        Sa_full = torch.empty((a.N, plan.n), dtype=torch.float64, device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def analyse(p, **kw):
        return p["h"].local_analysis_dev(p["xf"], p["Hxf"], p["yo"], p["Sf"], p["HSf"], p["var"], p["xa"], p["Sa"])

    def add_stats(tot, st):
        for k, v in st.items():
            tot[k] = tot.get(k, 0) + v
        return tot

    def step():
        tot, works = {}, []
        for p in phases:
            st = analyse(p)
            if not use_async:
                add_stats(tot, st)
            if world > 1 and not fused:   # asynchronous: overlaps the next phase's kernels
                works += allgather_slabs(dist, p["Sa"], p["plan"], out=Sa_full, wait=False)
        for w in works:
            w.wait()
        if use_async:
            for p in phases:
                add_stats(tot, p["h"].synchronize())
        if fused:
            peer.fence()   # readers after the writers of all ranks
        return tot

    if fused and use_mc:
        pass
    if fused and a.peer_mode == 2 and not use_mc:
        fused_note = "EXPERIMENT: results not pushed to the peers (compute-only timing, not a valid bench line)"
    elif fused:   # once: the fused result equals the NCCL all-gather of the slabs, bit for bit
        step()
        torch.cuda.synchronize()
        ref = torch.empty_like(Sa_full)
        for p in phases:
            allgather_slabs(dist, p["Sa"], p["plan"], out=ref)
        torch.cuda.synchronize()
        ok = torch.tensor([1.0 if torch.equal(ref, Sa_full) else 0.0], device=dev)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        fused_note = "fused peer stores checked against the NCCL all-gather: " + ("identical" if ok.item() == 1.0 else "MISMATCH")
        if ok.item() != 1.0:
            raise SystemExit("fused gather mismatch")
        del ref

    for _ in range(a.warmup):
        st = step()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    launches = 0
    for _ in range(a.steps):
        st = step()
        launches += st["launches"]
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop() if rank == 0 else None
    tms = torch.tensor([ms], dtype=torch.float64, device=dev)
    agg = torch.tensor([float(st["obs_relevant_sum"]), float(st["obs_candidate_sum"]), float(st["jacobi_sweeps_sum"]),
                        float(st["zones_skipped"])], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
        dist.all_reduce(agg, op=dist.ReduceOp.SUM)
    ms = float(tms.item())
    ms_per_step = ms / a.steps
    nzones = g.nzones
    value = nzones / (ms_per_step * 1e-3)

    # ---- per-kernel times (separate profiled pass: batches serialised, CUDA events around each kernel family)
    stp = {}
    for p in phases:
        p["h"].set_option("async", 0)
        p["h"].set_option("profile", 1)
        add_stats(stp, analyse(p))
        p["h"].set_option("profile", 0)
    peak_dfma = h.fp64_peak(0)
    peak_dmma = h.fp64_peak(1)

    out = None
    if rank == 0:
        analysed = nzones - agg[3].item()
        mloc_mean = agg[0].item() / max(analysed, 1)
        cand_mean = agg[1].item() / max(nzones, 1)
        sweeps_mean = agg[2].item() / max(analysed, 1)
        tri = a.eig_kernel == 4 and a.N <= 64
        per_batch = (6 + (1 if a.tvec_split else 0)) if tri else 3  # kernels per batch: gram, (tridiag, tql, tvec [x2 when split], fallback | eig), apply
        nb = max(1, (stp["launches"] - len(phases)) // per_batch)  # batches of this rank in one step
        N3 = a.N ** 3
        fz = flops_per_zone(a.N, a.nz, mloc_mean, cand_mean)
        # Algorithmic flops per zone credited to each kernel (SURVEY.md §8d split): the LAPACK
        # count 9 N^3 of dsyev = 4/3 N^3 (dsytrd, reduction to tridiagonal form) + the rest (QL iteration
        # with vectors + back-transformation), which together with T = U f(L) U^T (2 N^3) and ampl (4 N^2) is
        # what k_tql + k_tvec replace.
        stages = {
            "k_gram": (stp["ms_gram"], 2 * a.N * a.N * mloc_mean + 2 * mloc_mean * a.N + 25 * cand_mean),
            "k_apply": (stp["ms_apply"], 2 * a.nz * a.N * a.N + 2 * a.nz * a.N),
        }
        if tri:
            stages["k_tridiag"] = (stp["ms_tridiag"], 4.0 / 3.0 * N3)
            stages["k_tql+k_tvec"] = (stp["ms_tql"] + stp["ms_tvec"], (9 - 4.0 / 3.0) * N3 + 2 * N3 + 4 * a.N * a.N)
            if a.fuse_apply:
                # the apply runs inside k_tvec (k_apply only serves the zones it did not finish): one stage, the
                # credited flops of both
                ms_a, fl_a = stages.pop("k_apply")
                ms_t, fl_t = stages.pop("k_tql+k_tvec")
                stages["k_tql+k_tvec+apply"] = (ms_t + ms_a, fl_t + fl_a)
        else:
            stages["k_eig_fast"] = (stp["ms_eig"], 9 * N3 + 2 * N3 + 4 * a.N * a.N)
        # dram__bytes_read.sum + dram__bytes_write.sum per zone and kernel: parsed by tools/ncu_traffic.py from the
        # ncu --set full captures of the build named in that file (profiles/ncu_traffic.json: kernel options, git
        # hash, zones per launch); null when there is no capture for the kernel or the options differ
        traffic_zone = load_traffic(a)
        stage_out = {}
        for name, (ms_k, fl) in stages.items():
            ach = fl * z_rank / (ms_k * 1e-3) / 1e12 if ms_k > 0 else 0.0
            stage_out[name] = {"ms_per_step": ms_k, "algorithmic_flops_per_zone": fl, "achieved_tflops": ach,
                               "frac": ach / peak_dfma}
        dom = max((k for k in stages if k != "k_apply"), key=lambda k: stages[k][0])
        dom_ms, dom_fl = stages[dom]
        dom_ms_launch = dom_ms / nb
        achieved = dom_fl * (z_rank / nb) / (dom_ms_launch * 1e-3) / 1e12
        roof = {"bound": "fp64", "kernel": dom + " (largest share of the step among: " + ", ".join(stages) + ")",
                "achieved": achieved, "peak": peak_dfma, "unit": "TFLOP/s", "frac": achieved / peak_dfma,
                "traffic": traffic_zone[dom] * (z_rank / nb) if dom in traffic_zone else None,
                "traffic_source": "profiles/ncu_traffic.json (ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum per launch / zones per launch)" if dom in traffic_zone else None,
                "peak_source": "DFMA micro-kernel measured in this run (oakb200_fp64_peak); MEASURED_PEAKS.json has no "
                               "fp64 figure; DMMA m8n8k4 measured %.1f TFLOP/s" % peak_dmma,
                "algorithmic_flops_per_zone_kernel": dom_fl, "launches_per_step": nb,
                "ms_per_launch": dom_ms_launch, "zones_per_launch": z_rank / nb,
                "kernel_ms_per_step": {"pack": stp["ms_pack"], "gram": stp["ms_gram"], "eig": stp["ms_eig"],
                                       "apply": stp["ms_apply"], "tridiag": stp.get("ms_tridiag", 0.0),
                                       "tql": stp.get("ms_tql", 0.0), "tvec": stp.get("ms_tvec", 0.0)},
                "stages": stage_out,
                "zones_fallback_to_jacobi": int(stp.get("zones_fallback", 0)),
                "whole_step": {"algorithmic_flops_per_zone": fz, "achieved": value * fz / 1e12 / world,
                               "frac_of_fp64_peak_per_gpu": value * fz / 1e12 / world / peak_dfma},
                "hbm_view": {"algorithmic_bytes_per_zone": 2 * a.nz * a.N * 8 + 2 * a.nz * 8 + a.m * a.N * 8 / nzones,
                             "note": "HBM is not the binding roof (arithmetic intensity >100 flop/B)"}}
        out = {"metric": "local-analysis grid columns/sec", "value": value, "unit": "columns/s", "n_gpus": world,
               "steps": a.steps, "warmup": a.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
               "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
               "config": {"workload": workload_name(a), "zones": nzones, "parallelism": f"zone-range x{world}" + (((", all-gather replaced by NVSwitch multicast: every finished batch is stored once (multimem.st, k_push_mc on a side stream) and replicated by the switch into every rank's result array (cuMulticast object over the ranks' arrays); " if use_mc else (", all-gather replaced by stores of every finished batch into every rank's result array (k_push on a side stream over NVLink, CUDA IPC mappings); " if (a.peer_mode == 1 and a.push_kernel) else (", all-gather replaced by peer copies of every finished batch into every rank's result array (copy engines over NVLink, CUDA IPC mappings); " if a.peer_mode == 1 else ", all-gather fused into the apply kernel (stores into every rank's result array over NVLink, CUDA IPC); "))) + str(fused_note)) if fused else (f", {nphase} pipelined phases (all-gather of phase j overlaps analysis of j+1; " + ("asynchronous, stream priorities" if use_async else "blocking calls") + ")" + (("; " + fused_note) if fused_note else "") if world > 1 else "")),
                          "l2": "inputs (>= 15 GB state) exceed L2; no explicit flush",
                          "mean_relevant_obs_per_column": mloc_mean, "mean_candidates_per_column": cand_mean,
                          "mean_jacobi_sweeps": sweeps_mean, "eig_kernel": a.eig_kernel,
                          "gram_kernel": a.gram_kernel, "fuse_apply": a.fuse_apply, "tvec_split": a.tvec_split, "apply_kernel": a.apply_kernel},
               "gpu_launches": int(launches), "clocks": clocks, "roofline": roof}

    # ---- end to end through the host-buffer entry point (pinned host memory, H2D + D2H inside the timed region)
    e2e = None
    if not a.no_e2e:
        try:
            pin = lambda t: t.cpu().pin_memory()
            hb = []
            for p in phases:
                nl = p["plan"].r1 - p["plan"].r0
                q = dict(Sf=torch.empty((a.N, nl), dtype=torch.float64, pin_memory=True),
                         Sa=torch.empty((a.N, nl), dtype=torch.float64, pin_memory=True),
                         xa=torch.empty(nl, dtype=torch.float64, pin_memory=True),
                         xf=pin(p["xf"]), Hxf=pin(p["Hxf"]), yo=pin(p["yo"]), HSf=pin(p["HSf"]), var=pin(p["var"]))
                q["Sf"].copy_(p["Sf"])
                hb.append(q)
            torch.cuda.synchronize()

            def e2e_step():
                tot = {}
                for p, q in zip(phases, hb):
                    # what the Fortran shim does on every Assim: positions to the device + cell-grid build
                    p["h"].set_observations(obs_x=p["ox"], obs_y=p["oy"])
                    add_stats(tot, p["h"].local_analysis_pinned(q["xf"], q["Hxf"], q["yo"], q["Sf"], q["HSf"], q["var"],
                                                                q["xa"], q["Sa"]))
                return tot

            e2e_step()  # warm-up
            if os.environ.get("OAK_B200_CHUNK_MB"):   # experiment: chunk size of the host-streaming path
                for cm in os.environ["OAK_B200_CHUNK_MB"].split(","):
                    for p in phases:
                        p["h"].set_option("chunk_mb", float(cm))
                    e2e_step()
                    torch.cuda.synchronize()
                    t0 = time.perf_counter()
                    e2e_step(); e2e_step()
                    torch.cuda.synchronize()
                    print("chunk_mb", cm, "e2e columns/s %.0f" % (nzones / ((time.perf_counter() - t0) / 2)), file=sys.stderr, flush=True)
            barrier()
            t0 = time.perf_counter()
            for _ in range(a.e2e_steps):
                ste = e2e_step()
            torch.cuda.synchronize()
            te = (time.perf_counter() - t0) / a.e2e_steps
            tt = torch.tensor([te], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            ok = bool(torch.equal(hb[0]["Sa"][:, :1000].to(dev), phases[0]["Sa"][:, :1000]))
            e2e = {"value": nzones / float(tt.item()), "unit": "columns/s", "h2d_bytes_per_step": ste["h2d_bytes"],
                   "d2h_bytes_per_step": ste["d2h_bytes"], "steps": a.e2e_steps,
                   "note": "oakb200_set_observations + oakb200_local_analysis on pinned host buffers; state streamed in zone chunks; "
                           "bytes are per rank; no all-gather of host buffers" + ("" if ok else "; MISMATCH vs resident run")}
            # the same call on PAGEABLE host arrays (what a Fortran caller's allocatables are): through the library's pinned
            # staging ring filled by host threads (option host_stage, the default for large calls), with direct asynchronous
            # copies from the pageable arrays (the driver stages them), and with the arrays page-locked for the call
            if world == 1 and not a.no_pageable:
                pg = {}
                q = hb[0]
                pb = {k: torch.empty_like(v, pin_memory=False).copy_(v) for k, v in q.items()}

                def timed_pageable():
                    torch.cuda.synchronize()
                    t0 = time.perf_counter()
                    h.set_observations(obs_x=phases[0]["ox"], obs_y=phases[0]["oy"])
                    h.local_analysis_pinned(pb["xf"], pb["Hxf"], pb["yo"], pb["Sf"], pb["HSf"], pb["var"], pb["xa"], pb["Sa"])
                    torch.cuda.synchronize()
                    return nzones / (time.perf_counter() - t0)
                same = True
                for thr in [int(x) for x in os.environ.get("OAK_B200_STAGE_THREADS", "0").split(",")]:
                    h.set_option("host_stage", 1); h.set_option("stage_threads", thr)
                    timed_pageable()                       # first call allocates the pinned staging buffers
                    pb["Sa"].zero_()
                    pg["host_stage=1,threads=%s" % (thr or "default")] = timed_pageable()
                    same = same and bool(torch.equal(pb["Sa"], q["Sa"]))
                h.set_option("stage_threads", 0)
                h.set_option("host_stage", 0)
                pb["Sa"].zero_()
                pg["host_stage=0"] = timed_pageable()
                same = same and bool(torch.equal(pb["Sa"], q["Sa"]))
                if os.environ.get("OAK_B200_BENCH_HOST_REGISTER"):   # measured in round 2: 0.09 - 0.20 M columns/s, kept out of the default run
                    h.set_option("host_register", 1)
                    pg["host_register=1"] = timed_pageable()
                    h.set_option("host_register", 0)
                h.set_option("host_stage", -1)
                pg["identical_to_pinned_run"] = same
                e2e["pageable_columns_per_s"] = pg
                del pb
            del hb
        except Exception as ex:  # e.g. not enough pinnable host memory
            e2e = {"value": None, "unit": "columns/s", "h2d_bytes_per_step": None, "d2h_bytes_per_step": None,
                   "error": repr(ex)[:300]}
    if rank == 0:
        out["e2e"] = e2e
        # parity of THIS run's device result (left in d["xa"], d["Sa"] by the timed steps and the profile pass, which
        # analyse the same inputs) against the oracle, on the columns the CPU baseline analyses anyway
        try:
            if world == 1 and not a.no_cpu:
                out["cpu_baseline"], (_, sp_, kept_) = cpu_baseline(a, d, a.cpu_seconds)
                out["parity"] = parity_vs_oracle(a, d, sp_, kept_)
            else:
                out["parity"] = quick_parity(a, d)
        except Exception as ex:
            out.setdefault("cpu_baseline", {"value": None, "unit": "columns/s", "cores": None, "kind": "port",
                                            "sample": "failed: " + repr(ex)[:200]})
            out["parity"] = {"cols": 0, "ok": False, "error": repr(ex)[:200]}
        print(json.dumps(out))
    if peer is not None:
        Sa_full = None
        peer.Sa = peer.xa = None
        torch.cuda.synchronize()
        peer.close()
    h.close()
    if world > 1:
        dist.destroy_process_group()


# --------------------------------------------------------------------------------------------------
# BASELINE configs[3] ("C4"): 2000 x 2000 x 50, N = 128, an observation at every surface point, corrLen 10 km /
# cut-off 20 km (m_loc ~ 1257).  The state is 205 GB: it never exists as a whole.  Every rank walks over its zone
# range in slabs of `--slab-rows` grid rows; a slab (forecast anomalies, mean, the observations within one search
# radius) is generated on the device, analysed IN PLACE (Sa = Sf) through oakb200_local_analysis_dev and dropped.
# Timed: the analysis calls (CUDA events), summed over the slabs of a rank, max over ranks.  No all-gather: the
# analysed state stays distributed, as in the MPI reference until output (parall.F90:507).
# --------------------------------------------------------------------------------------------------
def run_c4(a, rank, world, local):
    import torch
    import torch.distributed as dist
    import oak_b200
    from oak_b200 import synthetic as S
    nx, ny, nz, N = (2000, 2000, 50, 128) if a.nx == 1000 else (a.nx, a.ny, a.nz, a.N)
    corr, maxlen = (10000.0, 20000.0) if a.corr == 4000.0 else (a.corr, a.maxlen)
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    g = S.Grid(nx, ny, nz)
    halo = int(math.ceil(maxlen / g.dx)) + 1
    j_first = [ny * p // world for p in range(world + 1)]           # parall.F90:176-177 on whole grid rows
    slabs = [(j, min(j + a.slab_rows, j_first[rank + 1])) for j in range(j_first[rank], j_first[rank + 1], a.slab_rows)]
    if a.max_slabs > 0:
        slabs = slabs[:a.max_slabs]
    h = oak_b200.Handle(local)
    peak = h.fp64_peak(0)

    def make_slab(j0, j1):
        zones = torch.arange(j0 * nx, j1 * nx, dtype=torch.int64, device=dev)
        nzs = zones.numel()
        n_loc = nzs * nz
        Sf = torch.empty((N, n_loc), dtype=torch.float64, device=dev)
        xf = torch.empty(n_loc, dtype=torch.float64, device=dev)
        step = 1 << 19
        for s0 in range(0, n_loc, step):
            e0 = min(n_loc, s0 + step)
            rows = torch.arange(j0 * nx * nz + s0, j0 * nx * nz + e0, dtype=torch.int64, device=dev)
            mean, anom = S.anomalies(torch, S.ensemble_rows(torch, g, rows, N, SEED))
            xf[s0:e0] = mean
            Sf[:, s0:e0] = anom
        # observations: one at every surface point of the grid rows within the halo (H = identity on the surface row)
        jo0, jo1 = max(0, j0 - halo), min(ny, j1 + halo)
        oz = torch.arange(jo0 * nx, jo1 * nx, dtype=torch.int64, device=dev)
        ox, oy = g.zone_xy(torch, oz)
        HE = torch.empty((N, oz.numel()), dtype=torch.float64, device=dev)
        for s0 in range(0, oz.numel(), step):
            e0 = min(oz.numel(), s0 + step)
            HE[:, s0:e0] = S.ensemble_rows(torch, g, oz[s0:e0] * nz, N, SEED)
        Hxf, HSf = S.anomalies(torch, HE)
        del HE
        yo = g.mu_rows(torch, oz * nz) + 0.05 * S.normal(torch, oz, 5, SEED)
        rm = 0.05 * (1.0 + 0.5 * S.uniform(torch, oz, 4, SEED))
        zx, zy = g.zone_xy(torch, zones)
        return dict(nzs=nzs, Sf=Sf, xf=xf, xa=torch.empty_like(xf), HSf=HSf.contiguous(), Hxf=Hxf.contiguous(), yo=yo,
                    var=(rm * rm).contiguous(), zx=zx.cpu().numpy(), zy=zy.cpu().numpy(), ox=ox.cpu().numpy(),
                    oy=oy.cpu().numpy())

    def analyse(d):
        h.set_zones(np.full(d["nzs"], nz, np.int32), zone_x=d["zx"], zone_y=d["zy"], corrLen=corr, maxLen=maxlen,
                    loctype=1, metrictype=0, weightfun=0)
        h.set_observations(obs_x=d["ox"], obs_y=d["oy"])
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        st = h.local_analysis_dev(d["xf"], d["Hxf"], d["yo"], d["Sf"], d["HSf"], d["var"], d["xa"], d["Sf"])  # in place
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1), st

    # warm-up: the first slab three times on copies (kernels, workspaces, clocks)
    d = make_slab(*slabs[0])
    keep = d["Sf"].clone()
    for _ in range(max(a.warmup, 1)):
        d["Sf"].copy_(keep)
        analyse(d)
    # parity of the first slab against the oracle on sampled columns (interior of the slab)
    parity = None
    if rank == 0:
        try:
            import oracle
            oracle.set_threads(0)
            d["Sf"].copy_(keep)
            analyse(d)
            rng = np.random.default_rng(1)
            zl = np.sort(rng.choice(d["nzs"], size=min(96, d["nzs"]), replace=False))
            rows = (zl[:, None] * nz + np.arange(nz)[None, :]).ravel()
            rt = torch.from_numpy(rows).to(dev)
            obs = oracle.make_obs(int(d["yo"].numel()), obsx=d["ox"], obsy=d["oy"])
            xo, So, _, _ = oracle.loc_analysis_cellgrid(np.full(zl.size, nz, np.int32), dict(x=d["zx"][zl], y=d["zy"][zl]), corr, maxlen, obs,
                                                        d["xf"][rt].cpu().numpy(), d["Hxf"].cpu().numpy(), d["yo"].cpu().numpy(),
                                                        np.asfortranarray(keep[:, rt].cpu().numpy().T),
                                                        np.asfortranarray(d["HSf"].cpu().numpy().T), d["var"].cpu().numpy())
            Sg, xg = d["Sf"][:, rt].cpu().numpy().T, d["xa"][rt].cpu().numpy()
            eS, ex = float(np.abs(Sg - So).max() / np.abs(So).max()), float(np.abs(xg - xo).max() / np.abs(xo).max())
            parity = {"cols": int(zl.size), "max_rel_Sa": eS, "max_rel_xa": ex, "tol": 1e-9, "ok": bool(eS < 1e-9 and ex < 1e-9),
                      "against": "oracle port (dsyev/dgemm, cell-grid selection) on sampled columns of the first slab"}
        except Exception as ex:
            parity = {"cols": 0, "ok": False, "error": repr(ex)[:200]}
    # per-kernel times of the first slab (profile pass: batches serialised)
    d["Sf"].copy_(keep)
    h.set_option("profile", 1)
    _, stp = analyse(d)
    h.set_option("profile", 0)
    kernel_ms = {k: stp.get("ms_" + k, 0.0) for k in ("pack", "gram", "eig", "apply")}
    del keep
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms_total, zones_done, launches, agg = 0.0, 0, 0, np.zeros(3)
    t_wall = time.perf_counter()
    for si, (j0, j1) in enumerate(slabs):
        if si > 0:
            del d
            d = make_slab(j0, j1)
        else:
            d = make_slab(j0, j1)
        ms, st = analyse(d)
        ms_total += ms
        zones_done += d["nzs"]
        launches += st["launches"]
        agg += [st["obs_relevant_sum"], st["obs_candidate_sum"], st["zones_skipped"]]
    wall = time.perf_counter() - t_wall
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([ms_total, float(zones_done), agg[0], agg[1]], dtype=torch.float64, device=dev)
    tmax = t.clone()
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    if rank == 0:
        zones_all = float(t[1].item())
        ms_step = float(tmax[0].item())
        value = zones_all / (ms_step * 1e-3)
        mloc = t[2].item() / max(zones_all, 1)
        cand = t[3].item() / max(zones_all, 1)
        fz = flops_per_zone(N, nz, mloc, cand)
        partial = a.max_slabs > 0
        out = {"metric": "local-analysis grid columns/sec", "value": value, "unit": "columns/s", "n_gpus": world,
               "steps": 1, "warmup": max(a.warmup, 1), "ms_per_step": ms_step, "higher_is_better": True,
               "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
               "config": {"workload": f"synthetic 3D ocean {nx}x{ny}x{nz}, N={N}, {nx * ny} obs (one per surface point), diagonal R, "
                                      f"local ETKF (gaussian corrLen {corr:g} m, cut-off {maxlen:g} m, cartesian)",
                          "zones": int(zones_all), "zones_in_full_grid": nx * ny,
                          "coverage": ("PARTIAL: %d slabs per rank" % a.max_slabs) if partial else "whole grid",
                          "parallelism": f"zone-range x{world}; state ({nx * ny * nz * N * 8 / 1e9:.0f} GB) never resident as a whole: slabs of "
                                         f"{a.slab_rows} grid rows generated on the device, analysed in place, dropped; no gather",
                          "mean_relevant_obs_per_column": mloc, "mean_candidates_per_column": cand,
                          "transform": "block Jacobi kernel (the tridiagonal route covers N <= 64)",
                          "timed": "sum of the oakb200_local_analysis_dev calls (CUDA events), max over ranks; slab generation "
                                   "and oakb200_set_zones / set_observations outside", "wall_s_with_generation": wall},
               "gpu_launches": int(launches), "clocks": clocks, "parity": parity,
               "roofline": {"bound": "fp64", "kernel": "whole step", "peak": peak, "unit": "TFLOP/s",
                            "achieved": value * fz / 1e12 / world, "frac": value * fz / 1e12 / world / peak,
                            "algorithmic_flops_per_zone": fz, "traffic": None,
                            "kernel_ms_first_slab": kernel_ms, "zones_first_slab": (slabs[0][1] - slabs[0][0]) * nx,
                            "peak_source": "DFMA micro-kernel measured in this run"},
               "e2e": None, "cpu_baseline": None}
        print(json.dumps(out))
    h.close()
    if world > 1:
        dist.destroy_process_group()


# --------------------------------------------------------------------------------------------------
# BASELINE configs[4] ("C5"): 4-D grid 256 x 256 x 10 levels x 4 times (a zone = the 40 elements of a column), N = 64,
# 2e5 observations, inflation.mult = 1.05, log anamorphosis, through the ENSEMBLE entry point (Assim's ensemble
# branch: H E, forward anamorphosis, mean / anomalies, local analysis, inflation, Ea, inverse anamorphosis).
# --------------------------------------------------------------------------------------------------
def run_c5(a, rank, world, local):
    import torch
    import oak_b200
    from oak_b200 import synthetic as S
    if rank != 0:
        return
    nx, ny, nz, N, m = (256, 256, 40, 64, 200000) if a.nx == 1000 else (a.nx, a.ny, a.nz, a.N, a.m)
    corr, maxlen = a.corr, a.maxlen
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    g = S.Grid(nx, ny, nz)
    rows = torch.arange(g.n, dtype=torch.int64, device=dev)
    E = torch.exp(0.3 * S.ensemble_rows(torch, g, rows, N, SEED)).contiguous()       # strictly positive, (N, n)
    obs = S.observations(np, g, m, SEED)
    Hi, Hj, Hs = S.coo_operator(g, obs)
    yo = np.log(1.0 + 0.2 * np.abs(np.asarray(g.mu_rows(np, obs["rows"][0])))) + 1.0 + np.asarray(obs["noise"])
    zx, zy = g.zone_xy(np, np.arange(g.nzones, dtype=np.int64))
    zs = np.full(g.nzones, nz, np.int32)
    t = lambda x, dt=torch.float64: torch.from_numpy(np.ascontiguousarray(x)).to(dev).to(dt)
    dHi, dHj, dHs, dyo, dvar = t(Hi, torch.int32), t(Hj, torch.int32), t(Hs), t(yo), t(obs["var"])
    h = oak_b200.Handle(local)
    h.set_zones(zs, zone_x=zx, zone_y=zy, corrLen=corr, maxLen=maxlen, loctype=1, metrictype=0, weightfun=0)
    h.set_observations(obs_x=obs["ox"], obs_y=obs["oy"])
    Ea = torch.empty_like(E)
    xf = torch.empty(g.n, dtype=torch.float64, device=dev)
    xa = torch.empty_like(xf)
    infl = 1.05

    def step():
        return h.assim_ensemble_dev(E, dHi, dHj, dHs, None, dyo, dvar, Ea, anamtype=2, inflation=infl, xf_out=xf, xa_out=xa)

    for _ in range(max(a.warmup, 3)):
        st = step()
    torch.cuda.synchronize()
    sampler = ClockSampler(local)
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    launches = 0
    for _ in range(a.steps):
        st = step()
        launches += st["launches"]
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.steps
    clocks = sampler.stop()
    value = g.nzones / (ms * 1e-3)
    # the ensemble branch in its two forms, timed the same way: fused (prologue / epilogue inside the apply kernel: 2 passes over
    # the state) and three-pass (k_mean_anom over E, analysis in place, k_epilogue over Ea: 6 passes); with the log anamorphosis
    # of this configuration and with none.  Both forms must give identical bits.  The library's default (ens_fuse = -1)
    # follows this measurement: fused only without anamorphosis.
    def timed(anam, fuse):
        h.set_option("ens_fuse", fuse)
        f = lambda: h.assim_ensemble_dev(E, dHi, dHj, dHs, None, dyo, dvar, Ea, anamtype=anam, inflation=infl, xf_out=xf, xa_out=xa)
        f(); f()
        torch.cuda.synchronize()
        e0.record()
        for _ in range(a.steps):
            f()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / a.steps, Ea.clone(), xa.clone()
    try:
        hbm_peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        hbm_peak = None
    saved_bytes = 8.0 * g.n * N * 4
    ens_fuse = {"passes_over_the_state": {"fused": 2, "three_pass": 6}, "saved_algorithmic_bytes": saved_bytes,
                "hbm_peak_GBps": hbm_peak, "peak_source": "MEASURED_PEAKS.json (driver-written copy bandwidth)" if hbm_peak else None}
    for name, anam in (("log_anamorphosis", 2), ("no_anamorphosis", 1)):
        t1, E1, x1 = timed(anam, 1)
        t0, E0, x0 = timed(anam, 0)
        ens_fuse[name] = {"fused_ms_per_step": t1, "three_pass_ms_per_step": t0,
                          "identical_bits": bool(torch.equal(E1, E0) and torch.equal(x1, x0)),
                          "saved_passes_GBps": saved_bytes / ((t0 - t1) * 1e-3) / 1e9 if t0 > t1 else None}
        del E1, E0, x1, x0
    h.set_option("ens_fuse", -1)
    step()
    analysed = g.nzones - st["zones_skipped"]
    mloc = st["obs_relevant_sum"] / max(analysed, 1)
    cand = st["obs_candidate_sum"] / max(g.nzones, 1)
    fz = flops_per_zone(N, nz, mloc, cand)
    peak = h.fp64_peak(0)
    # streaming passes of the ensemble branch (H E, mean / anomalies in, epilogue out): algorithmic bytes
    stream_bytes = 8.0 * g.n * N * 2 + 8.0 * m * N * 3
    # end to end through the host-buffer entry point (pageable numpy arrays in, analysed ensemble out)
    e2e = None
    if not a.no_e2e:
        Eh = np.asfortranarray(E.cpu().numpy().T)
        h.assim_ensemble(Eh, Hi, Hj, Hs, None, yo, oak_b200.DiagCovar(obs["var"]), 2, infl, None)
        t0 = time.perf_counter()
        for _ in range(a.e2e_steps):
            h.set_observations(obs_x=obs["ox"], obs_y=obs["oy"])
            Eah, _, _, ste = h.assim_ensemble(Eh, Hi, Hj, Hs, None, yo, oak_b200.DiagCovar(obs["var"]), 2, infl, None)
        te = (time.perf_counter() - t0) / a.e2e_steps
        e2e = {"value": g.nzones / te, "unit": "columns/s", "h2d_bytes_per_step": ste["h2d_bytes"],
               "d2h_bytes_per_step": ste["d2h_bytes"], "steps": a.e2e_steps,
               "note": "oakb200_set_observations + oakb200_assim_ensemble on host arrays (E in, Ea out)",
               "identical_to_resident": bool(np.array_equal(Eah, Ea.cpu().numpy().T))}
    # parity: the whole configuration through the oracle's ensemble branch (brute-force scan per column)
    parity, cpu = None, None
    if not a.no_cpu:
        try:
            import oracle
            cores = oracle.set_threads(0)
            oo = oracle.make_obs(m, obsx=obs["ox"], obsy=obs["oy"])
            t0 = time.perf_counter()
            Eo, xfo, xao = oracle.assim_ensemble(zs, dict(x=zx, y=zy), corr, maxlen, oo, np.asfortranarray(E.cpu().numpy().T), Hi, Hj,
                                                 Hs, np.zeros(m), yo, obs["var"], anamtype=2, inflation=infl)
            tc = time.perf_counter() - t0
            Eg = Ea.cpu().numpy().T
            eE = float(np.abs(Eg - Eo).max() / np.abs(Eo).max())
            ex = float(np.abs(xa.cpu().numpy() - xao).max() / np.abs(xao).max())
            parity = {"cols": int(g.nzones), "max_rel_Ea": eE, "max_rel_xa": ex, "tol": 1e-9, "ok": bool(eE < 1e-9 and ex < 1e-9),
                      "against": "oracle port of the ensemble branch (O(m) scan, dsyev/dgemm) on EVERY column of this workload"}
            cpu = {"value": g.nzones / tc, "unit": "columns/s", "cores": cores, "kind": "port",
                   "sample": f"the whole configuration ({g.nzones} columns, {m} observations scanned per column), {tc:.1f} s"}
        except Exception as ex:
            parity = {"cols": 0, "ok": False, "error": repr(ex)[:200]}
    out = {"metric": "local-analysis grid columns/sec", "value": value, "unit": "columns/s", "n_gpus": 1, "steps": a.steps,
           "warmup": max(a.warmup, 3), "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
           "dtype": "f64", "data": "synthetic",
           "config": {"workload": f"synthetic 4-D {nx}x{ny} columns x {nz} elements (10 levels x 4 times), N={N}, {m} obs, diagonal R, "
                                  f"ensemble entry point: H E, log anamorphosis, local ETKF (gaussian corrLen {corr:g} m, cut-off "
                                  f"{maxlen:g} m), inflation {infl}, inverse anamorphosis",
                      "zones": g.nzones, "parallelism": "1 GPU", "mean_relevant_obs_per_column": mloc,
                      "l2": "ensemble 1.3 GB per pass exceeds L2; no explicit flush"},
           "gpu_launches": int(launches), "clocks": clocks, "parity": parity,
           "roofline": {"bound": "fp64", "kernel": "whole step", "peak": peak, "unit": "TFLOP/s",
                        "achieved": value * fz / 1e12, "frac": value * fz / 1e12 / peak, "algorithmic_flops_per_zone": fz,
                        "traffic": None, "streaming_passes_algorithmic_bytes": stream_bytes, "ens_fuse": ens_fuse},
           "e2e": e2e, "cpu_baseline": cpu}
    print(json.dumps(out))
    h.close()


# --------------------------------------------------------------------------------------------------
# interpolation weights of the observation operator (genObservationOper -> cinterp,
# assimilation.F90:2471-2656, ndgrid.F90:1183-1257) for the observations of the C3 workload on its 3-D grid.
# metric: observations per second; the kernel is HBM / latency bound: algorithmic bytes per observation =
# 3 coordinates in + 8 corners x 3 subscripts (int32) + 8 weights + nbp out.
# --------------------------------------------------------------------------------------------------
def run_hgen(a, rank, world, local):
    import torch
    import oak_b200
    from oak_b200 import synthetic as S
    if rank != 0:
        return
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    g = S.Grid(a.nx, a.ny, a.nz)
    m = a.m
    gshape = (a.nx, a.ny, a.nz)
    axes = [g.dx * np.arange(1, a.nx + 1), g.dx * np.arange(1, a.ny + 1), -5.0 * np.arange(a.nz) ** 1.3]   # depth: descending, stretched
    rng = np.random.default_rng(SEED)
    lo = np.array([ax.min() for ax in axes]); hi = np.array([ax.max() for ax in axes])
    xi = lo + (hi - lo) * rng.uniform(-0.01, 1.01, (m, 3))                                   # ~4 % outside the grid
    masked = None
    h = oak_b200.Handle(local)
    t = lambda x, dt: torch.from_numpy(np.ascontiguousarray(x)).to(dev).to(dt)
    d_ax = t(np.concatenate(axes), torch.float64)
    d_xi = t(xi, torch.float64)
    d_idx = torch.empty((m, 8, 3), dtype=torch.int32, device=dev)
    d_co = torch.empty((m, 8), dtype=torch.float64, device=dev)
    d_nbp = torch.empty(m, dtype=torch.int32, device=dev)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)                     # > L2 (126 MB)
    for _ in range(max(a.warmup, 3)):
        h.cinterp_dev(gshape, d_ax, None, d_xi, d_idx, d_co, d_nbp)
    torch.cuda.synchronize()
    sampler = ClockSampler(local)
    sampler.start()
    tot = 0.0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(a.steps):
        flush.zero_()
        e0.record()
        h.cinterp_dev(gshape, d_ax, None, d_xi, d_idx, d_co, d_nbp)
        e1.record()
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    ms = tot / a.steps
    clocks = sampler.stop()
    value = m / (ms * 1e-3)
    bytes_obs = 3 * 8 + 8 * 3 * 4 + 8 * 8 + 4
    try:
        hbm_peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
        peak_src = "MEASURED_PEAKS.json (driver-written copy bandwidth)"
    except Exception:
        hbm_peak, peak_src = 6548.0, "fallback: the copy bandwidth measured on this pool's B200s in round 1"
    ach = bytes_obs * m / (ms * 1e-3) / 1e9
    # end to end: host arrays in, host arrays out (what the Fortran driver calls once per model variable)
    e2e = None
    if not a.no_e2e:
        h.cinterp(gshape, axes, xi[:1000])
        t0 = time.perf_counter()
        for _ in range(a.e2e_steps):
            idx_h, co_h, nbp_h = h.cinterp(gshape, axes, xi)
        te = (time.perf_counter() - t0) / a.e2e_steps
        e2e = {"value": m / te, "unit": "observations/s", "h2d_bytes_per_step": int(xi.nbytes + sum(ax.nbytes for ax in axes)),
               "d2h_bytes_per_step": int(idx_h.nbytes + co_h.nbytes + nbp_h.nbytes), "steps": a.e2e_steps,
               "note": "oakb200_cinterp on pageable host arrays (positions in; corner subscripts, weights, nbp out)",
               "identical_to_resident": bool(np.array_equal(idx_h, d_idx.cpu().numpy()) and np.array_equal(co_h, d_co.cpu().numpy()))}
    parity, cpu = None, None
    if not a.no_cpu:
        try:
            import oracle
            ns = 20000
            sel = rng.choice(m, ns, replace=False)
            coord = oracle.ndgrid_full_coords(gshape, axes=axes)
            og = oracle.NdGrid(gshape, coord)
            og.cinterp(xi[sel])                       # builds the part of the databox tree these points need
            t0 = time.perf_counter()
            i0, c0, n0 = og.cinterp(xi[sel])
            tc = time.perf_counter() - t0
            og.close()
            gi, gc, gn = d_idx[sel].cpu().numpy(), d_co[sel].cpu().numpy(), d_nbp[sel].cpu().numpy()
            ins = n0 > 0
            err = float(np.abs(gc[ins] - c0[ins]).max())
            parity = {"obs": ns, "nbp_equal": bool(np.array_equal(gn, n0)), "cells_equal": bool(np.array_equal(gi[ins], i0[ins])),
                      "max_abs_weight_error": err, "tol": 1e-12, "ok": bool(np.array_equal(gn, n0) and np.array_equal(gi[ins], i0[ins]) and err < 1e-12),
                      "against": "oracle restatement of cinterp (databox tree, simplices, dgetrf-rule elimination) on randomly sampled observations of this run"}
            cpu = {"value": ns / tc, "unit": "observations/s", "cores": 1, "kind": "port",
                   "sample": f"{ns} observations of this workload, databox tree already built (the reference runs genObservationOper serially, one observation after the other), {tc:.2f} s"}
        except Exception as ex:
            parity = {"obs": 0, "ok": False, "error": repr(ex)[:200]}
    out = {"metric": "observation-operator rows generated per second (interpolation weights, batched cinterp)", "value": value,
           "unit": "observations/s", "n_gpus": 1, "steps": a.steps, "warmup": max(a.warmup, 3), "ms_per_step": ms,
           "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
           "config": {"workload": f"interpolation weights of {m} observations on the {a.nx}x{a.ny}x{a.nz} rectilinear grid of the C3 workload "
                                  "(stretched descending depth axis), 3-D simplex interpolation as ndgrid.F90:464-665",
                      "l2": "256 MB written between timed iterations (L2 flush)"},
           "gpu_launches": int(a.steps), "clocks": clocks, "parity": parity,
           "roofline": {"bound": "hbm", "kernel": "k_cinterp<3>", "achieved": ach, "peak": hbm_peak, "unit": "GB/s", "frac": ach / hbm_peak,
                        "traffic": None, "algorithmic_bytes_per_observation": bytes_obs, "peak_source": peak_src,
                        "note": "one thread per observation: binary searches on the axes, up to 24 simplices x a 4x4 elimination; latency bound"},
           "e2e": e2e, "cpu_baseline": cpu}
    print(json.dumps(out))
    h.close()


def reference_arm(a, rank, world):
    """The reference's CPU implementation of the path (the oracle port: the Fortran reference cannot be
    built in this image) on a bounded sample of the same workload, all host threads."""
    if rank != 0:
        return
    import torch
    dev = torch.device("cuda", 0) if torch.cuda.is_available() else torch.device("cpu")
    d = build_rank_data(a, 0, 1, dev)
    import oracle
    host = host_obs_arrays(d)
    # torch.distributed.run exports OMP_NUM_THREADS=1 to its children: the baseline uses every online processor
    cores = oracle.set_threads(0)
    sp = cpu_sample_problem(a, d, 100000)
    probe = max(cores * 2, 16)
    t = run_oracle_sample(a, host, sp, probe)
    count = int(min(sp["zl"].size, max(probe, probe * 4.0 / max(t, 1e-6))))  # ~4 s per step
    for _ in range(a.warmup):
        run_oracle_sample(a, host, sp, count)
    t0 = time.perf_counter()
    for _ in range(a.steps):
        run_oracle_sample(a, host, sp, count)
    tt = (time.perf_counter() - t0) / a.steps
    v = count / tt
    sample = (f"{count} random columns per step of the same workload; every column scans all {host['m']} observations "
              f"(assimilation.F90:3745-3757), dgemm + dsyev per column, OpenMP dynamic over columns")
    print(json.dumps({"impl": "reference", "metric": "local-analysis grid columns/sec", "value": v,
                      "unit": "columns/s", "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup,
                      "ms_per_step": tt * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                      "dtype": "f64", "data": "synthetic",
                      "config": {"workload": workload_name(a), "zones": d["grid"].nzones},
                      "cpu_baseline": {"value": v, "unit": "columns/s", "cores": cores, "kind": "port",
                                       "sample": sample},
                      "e2e": {"value": v, "unit": "columns/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


if __name__ == "__main__":
    main()
