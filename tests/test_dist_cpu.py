"""Host-side multi-process logic (partition, observation halo, all-gather of slabs) under gloo,
world_size 2 and 3, on CPUs.  The per-rank analysis is the oracle here (stand-in for the device
call): the sharded pipeline must reproduce the single-process answer exactly."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, uneven, nphase, ret):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import oracle
        from oak_b200 import synthetic
        from oak_b200.dist import ShardPlan, allgather_slabs, phase_ranges
        c = synthetic.small_case(nx=20, ny=36, nz=3, N=8, m=260, corr=2000.0, maxlen=4000.0)
        zs = c["zoneSize"].copy()
        if uneven:
            zs = (1 + (np.arange(zs.size) % 3)).astype(np.int32)
        n = int(zs.sum())
        Sf, xf = c["Sf"][:n], c["xf"][:n]
        Sa = torch.zeros((c["N"], n), dtype=torch.float64)
        mloc_parts, halo_max, works = [], 0, []
        for first in phase_ranges(zs.size, world, nphase):
            plan = ShardPlan(zs, c["zx"], c["zy"], c["corr"], c["maxlen"], c["obs"]["ox"], c["obs"]["oy"], rank, world,
                             first=first)
            oi = plan.obs_idx
            halo_max = max(halo_max, len(oi))
            obs = oracle.make_obs(len(oi), obsx=c["obs"]["ox"][oi], obsy=c["obs"]["oy"][oi])
            xa_l, Sa_l, _, mloc_l = oracle.loc_analysis(plan.zoneSize, dict(x=plan.zx, y=plan.zy), plan.corrLen,
                                                        plan.maxLen, obs, xf[plan.r0:plan.r1], c["Hxf"][oi],
                                                        c["yo"][oi], Sf[plan.r0:plan.r1], c["HSf"][oi], c["var"][oi])
            mloc_parts.append((plan.z0, plan.z1, mloc_l))
            works += allgather_slabs(dist, torch.from_numpy(np.ascontiguousarray(Sa_l.T)), plan, out=Sa, wait=False)
        for w in works:
            w.wait()
        if rank == 0:
            obs_all = oracle.make_obs(c["m"], obsx=c["obs"]["ox"], obsy=c["obs"]["oy"])
            xa_g, Sa_g, _, mloc_g = oracle.loc_analysis(zs, dict(x=c["zx"], y=c["zy"]), c["corr"], c["maxlen"],
                                                        obs_all, xf, c["Hxf"], c["yo"], Sf, c["HSf"], c["var"])
            ok = np.array_equal(Sa.numpy().T, Sa_g) and all(np.array_equal(ml, mloc_g[a:b]) for a, b, ml in mloc_parts)
            ok = ok and halo_max < c["m"] and (mloc_g > 0).any()
            ret.put(bool(ok))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,uneven,nphase", [(2, False, 1), (3, True, 1), (2, False, 3)])
def test_sharded_pipeline_equals_single_process(world, uneven, nphase):
    ctx = mp.get_context("spawn")
    ret = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, uneven, nphase, ret)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=300)
        assert p.exitcode == 0
    assert ret.get(timeout=10) is True


def test_halo_is_superset_of_relevant_observations():
    sys.path.insert(0, ROOT)
    import oracle
    from oak_b200 import synthetic
    from oak_b200.dist import ShardPlan
    c = synthetic.small_case(nx=16, ny=16, nz=2, N=4, m=400, corr=1500.0, maxlen=3000.0)
    obs = oracle.make_obs(c["m"], obsx=c["obs"]["ox"], obsy=c["obs"]["oy"])
    for world in (2, 4, 8):
        seen = 0
        for rank in range(world):
            plan = ShardPlan(c["zoneSize"], c["zx"], c["zy"], c["corr"], c["maxlen"], c["obs"]["ox"],
                             c["obs"]["oy"], rank, world)
            halo = set(plan.obs_idx.tolist())
            for z in range(plan.z0, plan.z1, 7):
                _, rel = oracle.select_observations(obs, (c["zx"][z], c["zy"][z]), c["corr"], c["maxlen"])
                assert set(np.nonzero(rel)[0].tolist()) <= halo
            seen += plan.z1 - plan.z0
        assert seen == c["grid"].nzones
