"""Parity of the CUDA path (through the C ABI) with the CPU oracle and with the reference's own
known-answer tests.  Tolerances: observation index sets bit-exact; analysed mean / anomalies within
1e-9 relative (BASELINE.json north_star), written as RTOL below."""
import numpy as np
import pytest

import oracle
from refcases import assim_case, kalman_check, rrsqrt_case

pytestmark = pytest.mark.gpu

RTOL = 1e-9  # north star: relative tolerance on the analysed ensemble (fp64)
TOL_REF = 1e-8  # test/test_rrsqrt.F90:20


def rel(a, b):
    return np.abs(np.asarray(a) - np.asarray(b)).max() / max(np.abs(b).max(), 1e-300)


@pytest.fixture(scope="module")
def ob():
    import oak_b200
    return oak_b200


@pytest.fixture(scope="module", params=[4, 0, 1],
                ids=["eig_tridiag", "eig_fast", "eig_simple"])
def handle(request, ob):
    # pad_to=64 sends even the small golden cases through the production register-resident kernels
    # (4: tridiagonal route, the default; 0: register-resident block Jacobi; 1: simple shared-memory cross-check kernel)
    h = ob.Handle(0, eig_kernel=request.param, pad_to=64 if request.param != 1 else 0)
    yield h
    h.close()


# ------------------------------------------------------------------------------------------------
# reference known-answer tests (test/test_rrsqrt.F90)
# ------------------------------------------------------------------------------------------------
def _sel(ob, c, zoneSize, weightfun, corr, maxlen):
    starts = np.concatenate([[0], np.cumsum(zoneSize)[:-1]])
    return ob.Selector(zone_x=c["xmod"][starts], zone_y=np.zeros(len(zoneSize)), corrLen=corr, maxLen=maxlen,
                       obs_x=c["xobs"], obs_y=np.zeros(c["m"]), loctype=1, metrictype=0, weightfun=weightfun)


def test_rrsqrt_single_zone_and_zone_per_point_equal_global(ob, handle):
    # test/test_rrsqrt.F90:142-159
    c = rrsqrt_case()
    xa_check, Pa_check = kalman_check(c["xf"], c["Sf"], c["H"], c["y"], np.diag(c["var"]))
    for zs in ([c["n"]], [1] * c["n"]):
        xa, Sa, ampl = ob.locanalysis(zs, _sel(ob, c, zs, 2, 1.0, 1e30), c["xf"], c["Hxf"], c["y"], c["Sf"],
                                      c["HSf"], ob.DiagCovar(c["var"]), handle=handle)
        assert np.abs(xa - xa_check).max() < TOL_REF
        assert np.abs(Sa @ Sa.T - Pa_check).max() < TOL_REF
        assert np.abs(Sa.sum(axis=1)).max() < 1e-12
        assert (ampl == 0).all()  # rrsqrt.F90:324
        xo, So, _ = oracle.analysis(c["xf"], c["Hxf"], c["y"], c["Sf"], c["HSf"], c["var"])
        assert rel(xa, xo) < RTOL and rel(Sa, So) < RTOL


def test_rrsqrt_gaspari_cohn(ob, handle):
    # test/test_rrsqrt.F90:162-233, callback :254-271
    c = rrsqrt_case()
    n, m = c["n"], c["m"]
    zs = [1] * n
    xa, Sa, _ = ob.locanalysis(zs, _sel(ob, c, zs, 1, c["length"], 1e30), c["xf"], c["Hxf"], c["y"], c["Sf"],
                               c["HSf"], ob.DiagCovar(c["var"]), handle=handle)
    Pf = c["Sf"] @ c["Sf"].T
    R = np.diag(c["var"])
    xa_check = np.zeros(n)
    for i in range(n):
        w = np.array([oracle.locfun(abs(c["xmod"][i] - xo) / c["length"]) for xo in c["xobs"]])
        iloc = np.where(w != 0)[0]
        if len(iloc) == 0:
            xa_check[i] = c["xf"][i]
            continue
        invR = np.linalg.inv(R[np.ix_(iloc, iloc)]) * np.outer(w[iloc], w[iloc])
        Hl = c["H"][iloc]
        Pa = np.linalg.inv(np.linalg.inv(Pf) + Hl.T @ invR @ Hl)
        xa_check[i] = c["xf"][i] + Pa[i] @ (Hl.T @ (invR @ (c["y"][iloc] - Hl @ c["xf"])))
    assert np.abs(xa - xa_check).max() < TOL_REF
    obs = oracle.make_obs(m, obsx=c["xobs"], obsy=np.zeros(m), weightfun=1)
    xo, So, _, mloc = oracle.loc_analysis(zs, dict(x=c["xmod"], y=np.zeros(n)), c["length"], 1e30, obs, c["xf"],
                                          c["Hxf"], c["y"], c["Sf"], c["HSf"], c["var"])
    assert rel(xa, xo) < RTOL and rel(Sa, So) < RTOL
    # index sets, bit exact
    off, idx, w = handle.select_observations()
    assert (np.diff(off) == mloc).all()
    for i in range(n):
        wo, ro = oracle.select_observations(obs, (c["xmod"][i], 0.0), c["length"], 1e30)
        assert list(idx[off[i]:off[i + 1]] - 1) == list(np.nonzero(ro)[0])
        assert np.abs(w[off[i]:off[i + 1]] - wo[ro]).max() < 1e-15


def test_assim_case_ensemble_in_ensemble_out(ob, handle):
    # test/test_assim.F90:96-172 (tol 1e-5) through the ensemble entry point
    c = assim_case()
    n, N = c["n"], c["N"]
    xf = c["Ef"].sum(axis=1) / N
    xa_check, Pa_check = kalman_check(xf, (c["Ef"] - xf[:, None]) / np.sqrt(N - 1.0), c["H"], c["yo"],
                                      np.diag(c["var"]))
    sel = ob.Selector(zone_x=c["x"][:1], zone_y=c["y"][:1], obs_x=c["obsx"], obs_y=c["obsy"], metrictype=0,
                      weightfun=2)
    Ea, xf_o, xa_o = ob.assim_ensemble([n], sel, c["Ef"], [1], [5], [1.0], np.zeros(1), c["yo"],
                                       ob.DiagCovar(c["var"]), handle=handle)
    xa = Ea.sum(axis=1) / N
    Eap = Ea - xa[:, None]
    assert np.abs(xa - xa_check).max() < 1e-5
    assert np.abs(Eap @ Eap.T / (N - 1.0) - Pa_check).max() < 1e-5
    obs = oracle.make_obs(1, obsx=c["obsx"], obsy=c["obsy"], weightfun=2)
    Eo, xfo, xao = oracle.assim_ensemble([n], dict(x=c["x"][:1], y=c["y"][:1]), 1.0, 1e30, obs, c["Ef"],
                                         np.array([1], np.int32), np.array([5], np.int32), np.array([1.0]),
                                         np.zeros(1), c["yo"], c["var"])
    assert rel(Ea, Eo) < RTOL and rel(xf_o, xfo) < 1e-14 and rel(xa_o, xao) < RTOL


# ------------------------------------------------------------------------------------------------
# observation selection: bit-exact index sets for every metric / loctype / weight function
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("metric,loctype,weightfun", [(0, 1, 0), (1, 1, 0), (2, 1, 0), (0, 1, 1), (1, 1, 1),
                                                      (0, 2, 0), (0, 3, 0), (0, 1, 2)])
def test_selection_index_sets_bit_exact(ob, metric, loctype, weightfun):
    rng = np.random.default_rng(100 * metric + 10 * loctype + weightfun)
    nzones, m = 400, 3000
    if metric == 0:
        zx, zy = rng.uniform(0, 1e5, nzones), rng.uniform(0, 1e5, nzones)
        ox, oy = rng.uniform(-1e3, 1.01e5, m), rng.uniform(-1e3, 1.01e5, m)
        corr, maxl = rng.uniform(2e3, 6e3, nzones), rng.uniform(4e3, 2e4, nzones)
        # lattice points at exactly the cut-off distance exercise the `<=`
        zx[:20] = 1000.0 * np.arange(20); zy[:20] = 0.0
        ox[:20] = 1000.0 * np.arange(20) + 3000.0; oy[:20] = 4000.0
        maxl[:20] = 5000.0
    else:
        zx, zy = rng.uniform(-180, 180, nzones), rng.uniform(-89.5, 89.5, nzones)
        ox, oy = rng.uniform(-180, 360, m), rng.uniform(-90, 90, m)
        zx[:5] = [179.9, -179.9, 0.05, 359.9, 10.0]; zy[:5] = [0, 10, -20, 60, 89.9]
        corr, maxl = rng.uniform(2e5, 6e5, nzones), rng.uniform(3e5, 3e6, nzones)
        maxl[5:8] = [1.5e7, 2.1e7, 30.0]
    zz, oz = rng.uniform(0, 100, nzones), rng.uniform(0, 100, m)
    if loctype != 1:
        corr, maxl = rng.uniform(2, 6, nzones), rng.uniform(4, 20, nzones)
    h = ob.Handle(0)
    h.set_zones(np.ones(nzones, np.int32), zone_x=zx, zone_y=zy, zone_z=zz, zone_t=zz, corrLen=corr, maxLen=maxl,
                loctype=loctype, metrictype=metric, weightfun=weightfun)
    h.set_observations(obs_x=ox, obs_y=oy, obs_z=oz, obs_t=oz)
    off, idx, w = h.select_observations()
    h.close()
    obs = oracle.make_obs(m, obsx=ox, obsy=oy, obsz=oz, obst=oz, loctype=loctype, metrictype=metric,
                          weightfun=weightfun, trig=1)
    total = 0
    for z in range(nzones):
        wo, ro = oracle.select_observations(obs, (zx[z], zy[z], zz[z], zz[z]), corr[z], maxl[z])
        want = np.nonzero(ro)[0]
        got = idx[off[z]:off[z + 1]] - 1
        assert got.size == want.size and (got == want).all(), (z, got.size, want.size)
        if want.size:
            assert np.abs(w[off[z]:off[z + 1]] - wo[ro]).max() <= 4e-16 * max(1.0, np.abs(wo[ro]).max())
        total += want.size
    assert total > 0
    if metric == 0 and loctype == 1 and weightfun == 0:
        for z in range(20):  # the 3-4-5 triangles: d == maxLen exactly must be relevant
            assert z in (idx[off[z]:off[z + 1]] - 1)


def test_selection_matches_cellgrid_contract(ob):
    # test/test_cellgrid.F90:6-8,:64-83: integer lattices, query (2,2), every point with d < maxdist found
    for side, maxdist in [(4, 2.0), (20, 3.0), (300, 5.0)]:
        ii, jj = np.meshgrid(np.arange(1, side + 1.0), np.arange(1, side + 1.0), indexing="ij")
        ox, oy = ii.ravel(order="F"), jj.ravel(order="F")
        h = ob.Handle(0)
        h.set_zones([1], zone_x=[2.0], zone_y=[2.0], corrLen=1.0, maxLen=maxdist, metrictype=0)
        h.set_observations(obs_x=ox, obs_y=oy)
        off, idx, _ = h.select_observations()
        h.close()
        d = np.sqrt((ox - 2) ** 2 + (oy - 2) ** 2)
        assert set(np.nonzero(d < maxdist)[0]) <= set(idx - 1)
        assert list(idx - 1) == list(np.nonzero(d <= maxdist)[0])


# ------------------------------------------------------------------------------------------------
# local analysis on synthetic grids vs the oracle (LAPACK dsyev / BLAS dgemm)
# ------------------------------------------------------------------------------------------------
def _oracle_loc(c, zoneSize=None, e01=None, zone_list=None):
    obs = oracle.make_obs(c["m"], obsx=c["obs"]["ox"], obsy=c["obs"]["oy"])
    zs = c["zoneSize"] if zoneSize is None else zoneSize
    return oracle.loc_analysis(zs, dict(x=c["zx"], y=c["zy"]), c["corr"], c["maxlen"], obs, c["xf"], c["Hxf"],
                               c["yo"], c["Sf"], c["HSf"], c["var"], e01=e01, zone_list=zone_list)


def _configure(ob, h, c, zoneSize=None):
    sel = ob.Selector(zone_x=c["zx"], zone_y=c["zy"], corrLen=c["corr"], maxLen=c["maxlen"], obs_x=c["obs"]["ox"],
                      obs_y=c["obs"]["oy"], metrictype=0)
    h.configure(c["zoneSize"] if zoneSize is None else zoneSize, sel)


@pytest.mark.parametrize("N,m,nx,ny,nz", [(16, 300, 24, 20, 3), (64, 900, 20, 16, 5), (40, 500, 16, 12, 4),
                                          (100, 1200, 12, 10, 2), (128, 1500, 10, 8, 3), (20, 5, 30, 30, 1)])
def test_local_analysis_matches_oracle(ob, handle, N, m, nx, ny, nz):
    from oak_b200 import synthetic
    c = synthetic.small_case(nx=nx, ny=ny, nz=nz, N=N, m=m, corr=3000.0, maxlen=6000.0, seed=N + m)
    _configure(ob, handle, c)
    xa, Sa, ampl, st = handle.local_analysis(c["xf"], c["Hxf"], c["yo"], c["Sf"], c["HSf"], ob.DiagCovar(c["var"]))
    xo, So, _, mloc = _oracle_loc(c)
    assert st["obs_relevant_sum"] == mloc.sum()
    assert st["zones_skipped"] == (mloc == 0).sum()
    assert rel(xa, xo) < RTOL, rel(xa, xo)
    assert rel(Sa, So) < RTOL, rel(Sa, So)
    # columns of Sa keep summing to zero (ensemble input => Omega = I, rrsqrt.F90:166-176)
    assert np.abs(Sa.sum(axis=1)).max() < 1e-12 * max(1.0, np.abs(Sa).max()) * N


@pytest.mark.parametrize("N,m", [(24, 60), (64, 150)])
def test_localise_obs_false_all_observations_and_amplitudes(ob, N, m):
    """locAnalysis(..., localise_obs=.false.) (rrsqrt.F90:374-385, exercised by test/test_rrsqrt.F90:162-183): a zone
    with at least one relevant observation is analysed with ALL observations (Gaussian weights, no cut-off) and its
    amplitudes are returned; zones without a relevant observation are skipped."""
    from oak_b200 import synthetic
    c = synthetic.small_case(nx=12, ny=10, nz=3, N=N, m=m, corr=3000.0, maxlen=4000.0, seed=N + m)
    keep = c["obs"]["ox"] < 6000.0     # zones on the right have no relevant observation
    for k in ("Hxf", "yo", "var"):
        c[k] = c[k][keep]
    c["HSf"] = np.asfortranarray(c["HSf"][keep])
    c["obs"] = {k: (v[..., keep] if v.ndim > 1 else v[keep]) for k, v in c["obs"].items()}
    c["m"] = int(keep.sum())
    obs = oracle.make_obs(c["m"], obsx=c["obs"]["ox"], obsy=c["obs"]["oy"])
    xo, So, ao, mloc = oracle.loc_analysis(c["zoneSize"], dict(x=c["zx"], y=c["zy"]), c["corr"], c["maxlen"], obs,
                                           c["xf"], c["Hxf"], c["yo"], c["Sf"], c["HSf"], c["var"], local_obs=False,
                                           want_ampl=True)
    for gram_kernel in (1, 0):
        with ob.Handle(0, localise_obs=0, gram_kernel=gram_kernel) as h:
            _configure(ob, h, c)
            xa, Sa, ampl, st = h.local_analysis(c["xf"], c["Hxf"], c["yo"], c["Sf"], c["HSf"], ob.DiagCovar(c["var"]),
                                                want_amplitudes=True)
        assert 0 < st["zones_skipped"] < c["grid"].nzones
        assert rel(xa, xo) < RTOL and rel(Sa, So) < RTOL, (rel(xa, xo), rel(Sa, So))
        assert rel(ampl, ao) < RTOL, rel(ampl, ao)
        assert (ampl[:, np.abs(ao).sum(axis=0) == 0] == 0).all()


def test_edge_cases_empty_zones_all_relevant_excluded_obs_ragged_zones(ob, handle):
    from oak_b200 import synthetic
    c = synthetic.small_case(nx=30, ny=10, nz=4, N=24, m=200, corr=2500.0, maxlen=5000.0, seed=3)
    # observations only in the left third: most zones on the right have none
    keep = c["obs"]["ox"] < 10000.0
    for k in ("Hxf", "yo", "var"):
        c[k] = c[k][keep]
    c["HSf"] = np.asfortranarray(c["HSf"][keep])
    c["obs"] = {k: (v[..., keep] if v.ndim > 1 else v[keep]) for k, v in c["obs"].items()}
    c["m"] = int(keep.sum())
    nzones = c["grid"].nzones
    # ragged zones: sizes 1..7, total = n
    zs = []
    left = c["Sf"].shape[0]
    k = 0
    while left > 0:
        s = min(left, 1 + (k % 7)); zs.append(s); left -= s; k += 1
    zs = np.array(zs, np.int32)
    rng = np.random.default_rng(0)
    c["zx"] = rng.uniform(0, 30000, zs.size); c["zy"] = rng.uniform(0, 10000, zs.size)
    maxl = np.full(zs.size, 5000.0); maxl[0] = 1e9      # zone 0 sees every observation
    c["maxlen"] = maxl
    c["corr"] = np.full(zs.size, 2500.0)
    e01 = (rng.uniform(size=c["m"]) > 0.2).astype(np.float64)  # excluded observations (assimilation.F90:3086-3092)
    _configure(ob, handle, c, zs)
    R = ob.DCDCovar(e01, ob.DiagCovar(c["var"]))
    xa, Sa, _, st = handle.local_analysis(c["xf"], c["Hxf"], c["yo"], c["Sf"], c["HSf"], R)
    xo, So, _, mloc = _oracle_loc(c, zs, e01=e01)
    assert mloc[0] == c["m"] and (mloc == 0).sum() > 5 and st["zones_skipped"] == (mloc == 0).sum()
    assert rel(xa, xo) < RTOL and rel(Sa, So) < RTOL
    start = np.concatenate([[0], np.cumsum(zs)])
    for z in np.nonzero(mloc == 0)[0]:  # untouched zones are bit-identical to the forecast
        assert (Sa[start[z]:start[z + 1]] == c["Sf"][start[z]:start[z + 1]]).all()
        assert (xa[start[z]:start[z + 1]] == c["xf"][start[z]:start[z + 1]]).all()
    # m = 0: nothing to do
    handle.set_observations(obs_x=np.zeros(0), obs_y=np.zeros(0))
    xa0, Sa0, _, st0 = handle.local_analysis(c["xf"], np.zeros(0), np.zeros(0), c["Sf"], np.zeros((0, 24)),
                                             ob.DiagCovar(np.zeros(0)))
    assert (xa0 == c["xf"]).all() and (Sa0 == c["Sf"]).all() and st0["zones_skipped"] == zs.size


@pytest.mark.needs_torch_cuda
def test_in_place_chunked_and_device_resident_paths_agree(ob):
    import torch
    from oak_b200 import synthetic
    c = synthetic.small_case(nx=40, ny=30, nz=6, N=64, m=2500, corr=2500.0, maxlen=5000.0, seed=11)
    h = ob.Handle(0)
    _configure(ob, h, c)
    xa1, Sa1, _, st1 = h.local_analysis(c["xf"], c["Hxf"], c["yo"], c["Sf"], c["HSf"], ob.DiagCovar(c["var"]))
    # many small chunks / batches, in place
    h.set_option("chunk_mb", 0.05)
    h.set_option("zones_per_batch", 37)
    S = c["Sf"].copy(order="F")
    xa2, Sa2, _, st2 = h.local_analysis(c["xf"], c["Hxf"], c["yo"], S, c["HSf"], ob.DiagCovar(c["var"]), out_Sa=S)
    assert Sa2 is S and (Sa2 == Sa1).all() and (xa2 == xa1).all()
    assert st2["launches"] > st1["launches"] and st2["h2d_bytes"] == st1["h2d_bytes"]
    # device-resident entry point on torch tensors (member-major = column-major n x N)
    dev = torch.device("cuda:0")
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    Sf_d, HSf_d = t(c["Sf"].T), t(c["HSf"].T)
    xa_d, Sa_d = torch.empty(c["Sf"].shape[0], dtype=torch.float64, device=dev), torch.empty_like(Sf_d)
    st3 = h.local_analysis_dev(t(c["xf"]), t(c["Hxf"]), t(c["yo"]), Sf_d, HSf_d, t(c["var"]), xa_d, Sa_d)
    torch.cuda.synchronize()
    assert (Sa_d.cpu().numpy().T == Sa1).all() and (xa_d.cpu().numpy() == xa1).all()
    assert st3["h2d_bytes"] == 0 and st3["launches"] >= 4
    xo, So, _, _ = _oracle_loc(c)
    assert rel(xa1, xo) < RTOL and rel(Sa1, So) < RTOL
    h.close()


@pytest.mark.parametrize("in_place", [False, True])
def test_pageable_arrays_through_the_pinned_staging_ring(ob, in_place):
    # option host_stage (default for calls of 32 MB and more on pageable arrays, forced here): the chunks of the state
    # travel through the stream slots' pinned buffers, filled and drained by several host threads; same bits as the
    # direct asynchronous copies, with many chunks per slot, ragged zones, in place or not
    from oak_b200 import synthetic
    c = synthetic.small_case(nx=40, ny=30, nz=6, N=64, m=2500, corr=2500.0, maxlen=5000.0, seed=11)
    out = {}
    for stage in (0, 1):
        h = ob.Handle(0, host_stage=stage, stage_threads=3, chunk_mb=0.3, zones_per_batch=64)
        _configure(ob, h, c)
        S = c["Sf"].copy(order="F")
        xa, Sa, _, st = h.local_analysis(c["xf"], c["Hxf"], c["yo"], S, c["HSf"], ob.DiagCovar(c["var"]),
                                         **({"out_Sa": S} if in_place else {}))
        assert (Sa is S) == in_place
        out[stage] = (xa.copy(), Sa.copy(), st["h2d_bytes"], st["d2h_bytes"])
        h.close()
    assert (out[0][0] == out[1][0]).all() and (out[0][1] == out[1][1]).all() and out[0][2:] == out[1][2:]
    xo, So, _, _ = _oracle_loc(c)
    assert rel(out[1][0], xo) < RTOL and rel(out[1][1], So) < RTOL


def test_ensemble_entry_point_through_the_staging_buffers(ob):
    # oakb200_assim_ensemble on pageable arrays with host_stage = 1: E up and Ea down in pieces through the pinned
    # buffers (staged_copy), HSf of the local analysis likewise; identical to the plain copies
    from oak_b200 import synthetic
    g = synthetic.Grid(14, 11, 4)
    N, m = 24, 90
    rows = np.arange(g.n, dtype=np.int64)
    E = np.exp(0.3 * synthetic.ensemble_rows(np, g, rows, N, 5)).T.copy(order="F")
    obs = synthetic.observations(np, g, m, 5)
    Hi, Hj, Hs = synthetic.coo_operator(g, obs)
    yo = 1.0 + 0.1 * synthetic.normal(np, np.arange(m, dtype=np.int64), 8, 5)
    zx, zy = g.zone_xy(np, np.arange(g.nzones, dtype=np.int64))
    zs = np.full(g.nzones, 4, np.int32)
    sel = ob.Selector(zone_x=zx, zone_y=zy, corrLen=3000.0, maxLen=6000.0, obs_x=obs["ox"], obs_y=obs["oy"], metrictype=0)
    out = []
    for stage in (0, 1):
        h = ob.Handle(0, host_stage=stage, stage_threads=3)
        h.set_option("scheme", 1)
        h.configure(zs, sel)
        Ea, xf, xa, _ = h.assim_ensemble(E, Hi, Hj, Hs, None, yo, ob.DiagCovar(obs["var"]), anamtype=2, inflation=1.02)
        out.append((Ea, xf, xa))
        h.close()
    for a, b in zip(*out):
        assert np.array_equal(a, b)


@pytest.mark.needs_torch_cuda
def test_tapered_batches_give_the_same_bits(ob):
    # option taper (device-resident entry point): the last batches of a call halve so that the chain of kernels that
    # ends the call is short; zones are independent, so any partition into batches must give identical results
    import torch
    from oak_b200 import synthetic
    c = synthetic.small_case(nx=80, ny=64, nz=2, N=16, m=4000, corr=2500.0, maxlen=5000.0, seed=13)   # 5120 zones
    dev = torch.device("cuda:0")
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    out = []
    for taper in (0, 2):
        h = ob.Handle(0, zones_per_batch=2368, taper=taper)     # batches 2368, 2368, 384 against 2368, 1776, 976
        _configure(ob, h, c)
        Sf_d, HSf_d = t(c["Sf"].T), t(c["HSf"].T)
        xa_d, Sa_d = torch.empty(c["Sf"].shape[0], dtype=torch.float64, device=dev), torch.empty_like(Sf_d)
        st = h.local_analysis_dev(t(c["xf"]), t(c["Hxf"]), t(c["yo"]), Sf_d, HSf_d, t(c["var"]), xa_d, Sa_d)
        torch.cuda.synchronize()
        out.append((xa_d.cpu().numpy(), Sa_d.cpu().numpy(), st["launches"]))
        h.close()
    assert (out[0][0] == out[1][0]).all() and (out[0][1] == out[1][1]).all()


def test_assim_ensemble_with_inflation_anamorphosis_and_saturation(ob, handle):
    from oak_b200 import synthetic
    g = synthetic.Grid(18, 14, 3)
    N, m = 32, 150
    rows = np.arange(g.n, dtype=np.int64)
    E = np.exp(0.3 * synthetic.ensemble_rows(np, g, rows, N, 5)).T.copy(order="F")   # strictly positive
    obs = synthetic.observations(np, g, m, 5)
    Hi, Hj, Hs = synthetic.coo_operator(g, obs)
    Hj[:3] = 0; Hs[:3] = 0.0      # out-of-grid rows (assimilation.F90:2597-2611)
    Hshift = 0.01 * np.arange(m)
    yo = 1.0 + 0.1 * synthetic.normal(np, np.arange(m, dtype=np.int64), 8, 5)
    zx, zy = g.zone_xy(np, np.arange(g.nzones, dtype=np.int64))
    zs = np.full(g.nzones, 3, np.int32)
    sel = ob.Selector(zone_x=zx, zone_y=zy, corrLen=3000.0, maxLen=6000.0, obs_x=obs["ox"], obs_y=obs["oy"],
                      metrictype=0)
    maxc = np.full(g.n, 0.05)
    Ea, xf, xa = ob.assim_ensemble(zs, sel, E, Hi, Hj, Hs, Hshift, yo, ob.DiagCovar(obs["var"]), anamtype=2,
                                   inflation=1.05, maxCorrection=maxc, handle=handle)
    oo = oracle.make_obs(m, obsx=obs["ox"], obsy=obs["oy"])
    Eo, xfo, xao = oracle.assim_ensemble(zs, dict(x=zx, y=zy), 3000.0, 6000.0, oo, E, Hi, Hj, Hs, Hshift, yo,
                                         obs["var"], anamtype=2, inflation=1.05, maxCorrection=maxc)
    assert rel(xf, xfo) < 1e-14 and rel(xa, xao) < RTOL and rel(Ea, Eo) < RTOL
    # xa is the mean of the back-transformed analysis ensemble (assimilation.F90:3343-3349); the saturation of the
    # correction (the two `where` statements of :3311-3312) was active
    assert rel(xa, Ea.mean(axis=1)) < 1e-14
    assert (np.abs(np.log(Ea).mean(axis=1) - xf) > 0.0499).any()


@pytest.mark.parametrize("zone_rows", [3, 4, "ragged", 40])
def test_assim_ensemble_fused_in_the_apply_kernel_equals_the_three_pass_form(ob, zone_rows):
    # option ens_fuse (default 1): prologue (anamorphosis, mean, anomalies) and epilogue (inflation, saturation, Ea,
    # inverse anamorphosis, mean) of assimilation.F90:3123-3131,:3301-3349 run inside k_apply / k_apply_tma on the staged
    # rows; same operations in the same order as k_mean_anom / k_epilogue, so the results are IDENTICAL, bit for bit,
    # to ens_fuse = 0 — for zones of odd size (k_apply), even size (k_apply_tma on the device), unequal sizes, more
    # rows than one chunk (40 > 32), and zones without any observation (which keep the forecast but are still inflated
    # and back-transformed).  Both are compared with the oracle as well.
    from oak_b200 import synthetic
    nzg = {3: 3, 4: 4, "ragged": 3, 40: 40}[zone_rows]
    g = synthetic.Grid(10, 8, nzg)
    N, m = 24, 60
    rows = np.arange(g.n, dtype=np.int64)
    E = np.exp(0.3 * synthetic.ensemble_rows(np, g, rows, N, 5)).T.copy(order="F")
    obs = synthetic.observations(np, g, m, 5)
    obs["ox"] = 0.45 * obs["ox"]; obs["oy"] = 0.45 * obs["oy"]       # observations in one corner: far zones have none
    Hi, Hj, Hs = synthetic.coo_operator(g, obs)
    Hshift = 0.01 * np.arange(m)
    yo = 1.0 + 0.1 * synthetic.normal(np, np.arange(m, dtype=np.int64), 8, 5)
    zx, zy = g.zone_xy(np, np.arange(g.nzones, dtype=np.int64))
    if zone_rows == "ragged":      # merge pairs of columns: sizes 6, 3, 6, 3, ... (first element rule for the position)
        keep = np.ones(g.nzones, bool); keep[1::3] = False
        zs = np.where(np.roll(~keep, -1), 2 * nzg, nzg)[keep].astype(np.int32)
        zx, zy = zx[keep], zy[keep]
        assert zs.sum() == g.n
    else:
        zs = np.full(g.nzones, nzg, np.int32)
    sel = ob.Selector(zone_x=zx, zone_y=zy, corrLen=1500.0, maxLen=3000.0, obs_x=obs["ox"], obs_y=obs["oy"], metrictype=0)
    maxc = np.full(g.n, 0.05)
    out = {}
    for fuse in (0, 1):
        h = ob.Handle(0)
        h.set_option("ens_fuse", fuse)
        h.set_option("scheme", 1)
        h.configure(zs, sel)
        *out[fuse], st = h.assim_ensemble(E, Hi, Hj, Hs, Hshift, yo, ob.DiagCovar(obs["var"]), anamtype=2,
                                          inflation=1.05, maxCorrection=maxc)
        assert 0 < st["zones_skipped"] < st["zones_total"]
        out[fuse].append(st["launches"])
        h.close()
    assert out[1].pop() == out[0].pop() - 3     # k_mean_anom (state), k_epilogue and their pass over the state are gone
    for a, b in zip(out[0], out[1]):
        assert np.array_equal(a, b)
    oo = oracle.make_obs(m, obsx=obs["ox"], obsy=obs["oy"])
    Eo, xfo, xao = oracle.assim_ensemble(zs, dict(x=zx, y=zy), 1500.0, 3000.0, oo, E, Hi, Hj, Hs, Hshift, yo,
                                         obs["var"], anamtype=2, inflation=1.05, maxCorrection=maxc)
    Ea, xf, xa = out[1]
    assert rel(xf, xfo) < 1e-14 and rel(xa, xao) < RTOL and rel(Ea, Eo) < RTOL


@pytest.mark.parametrize("metric,weightfun", [(0, 0), (1, 0), (2, 1)])
def test_production_selection_counts_per_zone_equal_the_index_sets(ob, metric, weightfun):
    # the selection that matters runs inside the Gram kernel; its per-zone counts (oakb200_zone_counts) must equal the
    # sizes of the index sets of k_select (oakb200_select_observations, compared bit for bit with the oracle in
    # test_selection_index_sets_bit_exact) and the oracle's own counts: Cartesian, both spherical metrics, Gaussian and
    # Gaspari-Cohn weights
    from oak_b200 import synthetic
    c = synthetic.small_case(nx=30, ny=24, nz=3, N=32, m=1800, corr=2500.0, maxlen=5000.0, seed=17)
    zx, zy, ox, oy, corr, maxlen = c["zx"], c["zy"], c["obs"]["ox"], c["obs"]["oy"], c["corr"], c["maxlen"]
    if metric != 0:      # metres -> degrees around 45 N, lengths stay in metres
        zx, zy, ox, oy = zx / 80e3, 45.0 + zy / 111e3, ox / 80e3, 45.0 + oy / 111e3
    sel = ob.Selector(zone_x=zx, zone_y=zy, corrLen=corr, maxLen=maxlen, obs_x=ox, obs_y=oy, metrictype=metric, weightfun=weightfun)
    with ob.Handle(0) as h:
        h.configure(c["zoneSize"], sel)
        xa, Sa, _, st = h.local_analysis(c["xf"], c["Hxf"], c["yo"], c["Sf"], c["HSf"], ob.DiagCovar(c["var"]))
        counts = h.zone_counts()
        off, idx, _ = h.select_observations(0, len(c["zoneSize"]))
    assert (counts == np.diff(off)).all() and counts.sum() == st["obs_relevant_sum"]
    obs = oracle.make_obs(c["m"], obsx=ox, obsy=oy, metrictype=metric, weightfun=weightfun, trig=1)
    xo, So, _, mloc = oracle.loc_analysis(c["zoneSize"], dict(x=zx, y=zy), corr, maxlen, obs, c["xf"], c["Hxf"], c["yo"],
                                          c["Sf"], c["HSf"], c["var"])
    assert (counts == mloc).all() and mloc.max() > 20
    assert rel(xa, xo) < RTOL and rel(Sa, So) < RTOL


def test_error_behaviour(ob):
    h = ob.Handle(0)
    with pytest.raises(ob.OakB200Error):      # analysis before configuration
        h.local_analysis(np.zeros(2), np.zeros(1), np.zeros(1), np.zeros((2, 4)), np.zeros((1, 4)),
                         ob.DiagCovar(np.ones(1)))
    h.set_zones([2], zone_x=[0.0], zone_y=[0.0], corrLen=1.0, maxLen=10.0, metrictype=0)
    h.set_observations(obs_x=[0.0], obs_y=[0.0])
    with pytest.raises(ob.OakB200Error):      # n mismatch
        h.local_analysis(np.zeros(3), np.zeros(1), np.zeros(1), np.zeros((3, 4)), np.zeros((1, 4)),
                         ob.DiagCovar(np.ones(1)))
    with pytest.raises(ob.OakB200Error):      # N too large
        h.local_analysis(np.zeros(2), np.zeros(1), np.zeros(1), np.zeros((2, 200)), np.zeros((1, 200)),
                         ob.DiagCovar(np.ones(1)))
    with pytest.raises(ob.OakB200Error) as e:  # NaN in the amplitudes is fatal (rrsqrt.F90:145-149)
        h.local_analysis(np.zeros(2), np.zeros(1), np.array([np.nan]), np.ones((2, 4)), np.ones((1, 4)),
                         ob.DiagCovar(np.ones(1)))
    assert e.value.code == -7
    with pytest.raises(ob.OakB200Error):      # unsupported metric (assimilation.F90:3666-3669)
        h.set_zones([2], zone_x=[0.0], zone_y=[0.0], corrLen=1.0, maxLen=10.0, metrictype=5)
    h.close()


def test_fp64_peak_microbenchmarks(ob):
    h = ob.Handle(0)
    dfma, dmma = h.fp64_peak(0), h.fp64_peak(1)
    h.close()
    assert 5.0 < dfma < 100.0 and 1.0 < dmma < 200.0


# ------------------------------------------------------------------------------------------------
# BASELINE.json configs at reduced size (parity-test cases, not bench lines)
# ------------------------------------------------------------------------------------------------
def test_config2_shallow_water_2d_three_staggered_variables(ob, handle):
    """configs[1]: test/shallow_water2d_ens.F90 local ETKF — zeta (nx x ny), ubar ((nx-1) x ny),
    vbar (nx x (ny-1)) on a C-grid with land points removed, zone = horizontal cell holding its
    surviving {zeta,ubar,vbar} points (sizes 1..3), N = 20, 5 observations of zeta, rmse 0.05,
    Cartesian metric, corrLength 10e3, maxLength 40e3 (test/shallow_water2d.init with schemetype = 1)."""
    nx, ny, N, dx = 30, 24, 20, 5e3
    rng = np.random.default_rng(42)
    ii, jj = np.meshgrid(np.arange(nx), np.arange(ny), indexing="ij")
    sea = ((ii - nx / 2) ** 2 / (nx / 2) ** 2 + (jj - ny / 2) ** 2 / (ny / 2) ** 2) < 0.92   # basin with closed boundary
    mask_u = sea[:-1, :] & sea[1:, :]
    mask_v = sea[:, :-1] & sea[:, 1:]
    part = (ii + nx * jj)
    labels = np.concatenate([part[sea], part[:-1, :][mask_u], part[:, :-1][mask_v]])   # Zones.partition of the 3 variables
    xs = np.concatenate([((ii + 1) * dx)[sea], ((ii[:-1] + 1.5) * dx)[mask_u], ((ii[:, :-1] + 1) * dx)[mask_v]])
    ys = np.concatenate([((jj + 1) * dx)[sea], ((jj[:-1] + 1) * dx)[mask_u], ((jj[:, :-1] + 1.5) * dx)[mask_v]])
    # initPartition: gap-free relabel + stable counting sort (assimilation.F90:405-430,:578-641)
    uniq, relabel = np.unique(labels, return_inverse=True)
    zs, zoneIndex, _ = oracle.init_partition((relabel + 1).astype(np.int32), uniq.size)
    perm = zoneIndex - 1
    n = perm.size
    starts = np.concatenate([[0], np.cumsum(zs)[:-1]])
    zx, zy = xs[perm][starts], ys[perm][starts]          # position of each zone's first element
    assert set(np.unique(zs)) <= {1, 2, 3} and len(set(np.unique(zs))) > 1
    # ensemble of Gaussian bumps (shallow_water2d_ens.F90:191-209), observations of zeta at 5 sea points
    gz = rng.normal(size=(3, N))
    xc, yc, zc = 20e3 + 10e3 * gz[0], 50e3 + 30e3 * gz[1], gz[2]
    E_full = zc[None, :] * np.exp(-((xs[:, None] - xc) / 30e3) ** 2 - ((ys[:, None] - yc) / 60e3) ** 2)
    E_full[sea.sum():] = 0.01 * rng.normal(size=(n - sea.sum(), N))   # U, V
    E = np.asfortranarray(E_full[perm])
    sea_idx = np.nonzero(sea.ravel(order="F"))[0]
    obs_pts = rng.choice(sea.sum(), size=5, replace=False)
    inv = np.empty(n, np.int64); inv[perm] = np.arange(n)
    Hj = (inv[obs_pts] + 1).astype(np.int32); Hi = np.arange(1, 6, dtype=np.int32); Hs = np.ones(5)
    ox, oy = xs[obs_pts], ys[obs_pts]
    yo = 0.3 * rng.normal(size=5)
    var = np.full(5, 0.05 ** 2)
    sel = ob.Selector(zone_x=zx, zone_y=zy, corrLen=10e3, maxLen=40e3, obs_x=ox, obs_y=oy, metrictype=0)
    Ea, xf, xa = ob.assim_ensemble(zs, sel, E, Hi, Hj, Hs, None, yo, ob.DiagCovar(var), handle=handle)
    oo = oracle.make_obs(5, obsx=ox, obsy=oy)
    Eo, xfo, xao = oracle.assim_ensemble(zs, dict(x=zx, y=zy), 10e3, 40e3, oo, E, Hi, Hj, Hs, np.zeros(5), yo, var)
    assert rel(Ea, Eo) < RTOL and rel(xa, xao) < RTOL and rel(xf, xfo) < 1e-14
    assert np.abs(Ea - E).max() > 1e-3     # the analysis did something


def test_config4_dense_observations_large_radius_n128(ob, handle):
    """configs[3] at reduced size: N = 128, an observation at every surface grid point, large radius."""
    from oak_b200 import synthetic
    g = synthetic.Grid(26, 22, 3)
    N = 128
    rows = np.arange(g.n, dtype=np.int64)
    E = synthetic.ensemble_rows(np, g, rows, N, 9)
    zones = np.arange(g.nzones, dtype=np.int64)
    zx, zy = g.zone_xy(np, zones)
    surf = zones * g.nz
    HE = E[:, surf]                                   # H = identity on the surface level
    yo = g.mu_rows(np, surf) + 0.05 * synthetic.normal(np, zones, 6, 9)
    xf, Sf = synthetic.anomalies(np, E)
    Hxf, HSf = synthetic.anomalies(np, HE)
    var = np.full(g.nzones, 0.05 ** 2)
    c = dict(m=g.nzones, obs=dict(ox=zx, oy=zy), zx=zx, zy=zy, corr=5000.0, maxlen=10000.0, xf=xf, Hxf=Hxf, yo=yo,
             Sf=np.asfortranarray(Sf.T), HSf=np.asfortranarray(HSf.T), var=var,
             zoneSize=np.full(g.nzones, g.nz, np.int32))
    _configure(ob, handle, c)
    xa, Sa, _, st = handle.local_analysis(c["xf"], c["Hxf"], c["yo"], c["Sf"], c["HSf"], ob.DiagCovar(var))
    xo, So, _, mloc = _oracle_loc(c)
    assert mloc.max() > 250 and st["obs_relevant_sum"] == mloc.sum()
    assert rel(xa, xo) < RTOL and rel(Sa, So) < RTOL


def test_config5_time_localisation_inflation_anamorphosis(ob, handle):
    """configs[4] at reduced size: 4-D state (zone = column over z and t), localisation in time
    (loctype = 3: |obsT - t|, assimilation.F90:3752-3753), N = 64, inflation.mult = 1.05, log anamorphosis."""
    from oak_b200 import synthetic
    nxy, nzt, N, m = 90, 8, 64, 240
    n = nxy * nzt
    rng = np.random.default_rng(5)
    E = np.exp(0.2 * rng.normal(size=(n, N)) + 0.1 * np.sin(np.arange(n))[:, None])
    zs = np.full(nxy, nzt, np.int32)
    zt = rng.uniform(0, 10, nxy)                     # time coordinate of each zone's first element
    ot = rng.uniform(-1, 11, m)
    rows = rng.integers(0, n, size=(2, m))
    Hi = np.tile(np.arange(1, m + 1, dtype=np.int32), 2)
    Hj = (rows.reshape(-1) + 1).astype(np.int32)
    Hs = np.tile([0.6, 0.4], (m, 1)).T.reshape(-1).copy()
    yo = 1.0 + 0.1 * rng.normal(size=m)
    var = rng.uniform(0.01, 0.04, m)
    sel = ob.Selector(zone_x=np.zeros(nxy), zone_y=np.zeros(nxy), zone_t=zt, corrLen=0.8, maxLen=1.6,
                      obs_x=np.zeros(m), obs_y=np.zeros(m), obs_t=ot, loctype=3, metrictype=0)
    Ea, xf, xa = ob.assim_ensemble(zs, sel, E, Hi, Hj, Hs, None, yo, ob.DiagCovar(var), anamtype=2, inflation=1.05,
                                   handle=handle)
    oo = oracle.make_obs(m, obst=ot, loctype=3)
    Eo, xfo, xao = oracle.assim_ensemble(zs, dict(t=zt), 0.8, 1.6, oo, E, Hi, Hj, Hs, np.zeros(m), yo, var,
                                         anamtype=2, inflation=1.05)
    assert rel(Ea, Eo) < RTOL and rel(xa, xao) < RTOL


@pytest.mark.parametrize("scale", [1e-4, 1.0, 40.0])
def test_dynamic_range_of_the_spectrum(ob, handle, scale):
    """lambda_max from ~1e-6 to ~1e6: the fp32-steered rotations and the deferred column scales must not
    lose the 1e-9 parity at either end (dsyev itself is good to ~eps*lambda_max here)."""
    from oak_b200 import synthetic
    c = synthetic.small_case(nx=14, ny=12, nz=3, N=64, m=700, corr=3000.0, maxlen=6000.0, seed=77)
    c["HSf"] = np.asfortranarray(c["HSf"] * scale)
    c["Hxf"] = c["Hxf"] * scale
    c["yo"] = c["yo"] * scale
    _configure(ob, handle, c)
    xa, Sa, _, st = handle.local_analysis(c["xf"], c["Hxf"], c["yo"], c["Sf"], c["HSf"], ob.DiagCovar(c["var"]))
    xo, So, _, mloc = _oracle_loc(c)
    assert rel(xa, xo) < RTOL and rel(Sa, So) < RTOL, (rel(xa, xo), rel(Sa, So))


@pytest.mark.parametrize("N", [24, 64])
def test_degenerate_spectrum_falls_back_to_jacobi(ob, N):
    """Repeated non-zero eigenvalues of G (orthogonal observation rows of equal norm, weight and variance):
    independent factorisations of T - lambda I cannot give an orthonormal basis of the eigenspace, so the
    tridiagonal route (eig_kernel 4) must hand these zones to the Jacobi kernel; close-but-distinct eigenvalues
    are orthogonalised in place.  Either way the result matches dsyev within RTOL."""
    from oak_b200 import synthetic
    c = synthetic.small_case(nx=6, ny=5, nz=2, N=N, m=6, corr=1e9, maxlen=1e12, seed=5)  # every zone sees all 6 obs
    rng = np.random.default_rng(3)
    Q, _ = np.linalg.qr(rng.normal(size=(N, 6)))
    rows = Q.T.copy()                      # 6 orthonormal rows
    rows[:3] *= 2.0                        # eigenvalue 4 w^2/r three times
    rows[3:5] *= 1.0 + np.array([0.0, 3e-7])[:, None]  # two eigenvalues 6e-7 apart (relative)
    rows[5] *= 0.5
    c["HSf"] = np.asfortranarray(rows)
    c["var"] = np.full(6, 0.25)
    xo, So, _, mloc = _oracle_loc(c)
    assert (mloc == 6).all()
    with ob.Handle(0, eig_kernel=4) as h:
        _configure(ob, h, c)
        xa, Sa, _, st = h.local_analysis(c["xf"], c["Hxf"], c["yo"], c["Sf"], c["HSf"], ob.DiagCovar(c["var"]))
        assert rel(xa, xo) < RTOL and rel(Sa, So) < RTOL, (rel(xa, xo), rel(Sa, So), st["zones_fallback"])
        # no in-place orthogonalisation allowed: every zone with a close pair goes to the Jacobi kernel
        h.set_option("tri_maxgroup", 0)
        xa, Sa, _, st = h.local_analysis(c["xf"], c["Hxf"], c["yo"], c["Sf"], c["HSf"], ob.DiagCovar(c["var"]))
        assert st["zones_fallback"] == len(mloc) and st["jacobi_sweeps_sum"] > 0
        assert rel(xa, xo) < RTOL and rel(Sa, So) < RTOL, (rel(xa, xo), rel(Sa, So))


@pytest.mark.needs_torch_cuda
def test_fused_gather_peer_outputs_receive_the_slab(ob):
    """oakb200_set_peer_outputs: the apply kernel stores this rank's rows (analysed and untouched zones) into
    every destination array at row0 (here two arrays on the same device, allocated through oakb200_ipc_alloc;
    across GPUs the destinations are CUDA-IPC mappings, exercised by bench.py --gpus N)."""
    import torch
    from oak_b200 import synthetic
    c = synthetic.small_case(nx=24, ny=18, nz=4, N=24, m=30, corr=1500.0, maxlen=3000.0, seed=21)
    dev = torch.device("cuda", 0)
    t = lambda x: torch.as_tensor(np.ascontiguousarray(x), dtype=torch.float64, device=dev)
    N, n = c["Sf"].shape[1], c["Sf"].shape[0]
    row0, ntot = 7, n + 19                      # the slab sits in the middle of a larger "global" array
    with ob.Handle(0) as h:
        _configure(ob, h, c)
        bufs = [h.ipc_alloc(8 * (N * ntot + ntot)) for _ in range(2)]

        class Raw:
            def __init__(self, ptr, count):
                self.__cuda_array_interface__ = {"shape": (count,), "typestr": "<f8", "data": (ptr, False), "version": 2}
        flats = [torch.as_tensor(Raw(ptr, N * ntot + ntot), device=dev) for ptr, _ in bufs]
        for f in flats:
            f.fill_(-7.0)
        h.set_peer_outputs([ptr for ptr, _ in bufs], [ptr + 8 * N * ntot for ptr, _ in bufs], ntot, row0)
        Sf = t(c["Sf"].T)
        xa, Sa = torch.empty(n, dtype=torch.float64, device=dev), torch.empty_like(Sf)
        st = h.local_analysis_dev(t(c["xf"]), t(c["Hxf"]), t(c["yo"]), Sf, t(c["HSf"].T), t(c["var"]), xa, Sa)
        torch.cuda.synchronize()
        assert st["zones_skipped"] > 0           # untouched zones are forwarded too
        for f in flats:
            S = f[:N * ntot].view(N, ntot)
            x = f[N * ntot:]
            assert torch.equal(S[:, row0:row0 + n], Sa) and torch.equal(x[row0:row0 + n], xa)
            assert (S[:, :row0] == -7.0).all() and (S[:, row0 + n:] == -7.0).all() and (x[:row0] == -7.0).all()
        h.set_peer_outputs([], [], 0, 0)
        del flats
        for ptr, _ in bufs:
            h.ipc_free(ptr)
    xo, So, _, _ = _oracle_loc(c)
    assert rel(xa.cpu().numpy(), xo) < RTOL and rel(Sa.cpu().numpy().T, So) < RTOL


@pytest.mark.parametrize("monotone", [True, False])
def test_assim_ensemble_tabulated_anamorphosis(ob, monotone):
    """Anamorphosis type 3 (assimilation.F90:4539-4567, interp1 anamorphosis.F90:304-339): piecewise-linear table,
    values outside the table (forward AND inverse direction) take the reference's clamping rule; a non-monotone
    table goes through the linear first-bracket scan."""
    from oak_b200 import synthetic
    g = synthetic.Grid(16, 12, 3)
    N, m = 24, 120
    rows = np.arange(g.n, dtype=np.int64)
    E = np.exp(0.5 * synthetic.ensemble_rows(np, g, rows, N, 9)).T.copy(order="F")   # ~0.2 .. 5
    xs = np.array([0.5, 0.7, 1.0, 1.3, 1.7, 2.0])          # ~2 % of the values fall outside on either side
    tab = np.column_stack([xs, np.log(xs) * 2.0])
    if not monotone:
        tab = tab[[0, 1, 3, 2, 4, 5]]
    obs = synthetic.observations(np, g, m, 9)
    Hi, Hj, Hs = synthetic.coo_operator(g, obs)
    yo = 1.2 + 0.2 * synthetic.normal(np, np.arange(m, dtype=np.int64), 8, 9)
    zx, zy = g.zone_xy(np, np.arange(g.nzones, dtype=np.int64))
    zs = np.full(g.nzones, 3, np.int32)
    sel = ob.Selector(zone_x=zx, zone_y=zy, corrLen=3000.0, maxLen=6000.0, obs_x=obs["ox"], obs_y=obs["oy"],
                      metrictype=0)
    with ob.Handle(0) as h:
        with pytest.raises(ob.OakB200Error):       # type 3 without a table
            ob.assim_ensemble(zs, sel, E, Hi, Hj, Hs, None, yo, ob.DiagCovar(obs["var"]), anamtype=3, handle=h)
        Ea, xf, xa = ob.assim_ensemble(zs, sel, E, Hi, Hj, Hs, None, yo, ob.DiagCovar(obs["var"]), anamtype=3,
                                       inflation=1.02, handle=h, anamtable=tab)
    oo = oracle.make_obs(m, obsx=obs["ox"], obsy=obs["oy"])
    Eo, xfo, xao = oracle.assim_ensemble(zs, dict(x=zx, y=zy), 3000.0, 6000.0, oo, E, Hi, Hj, Hs, np.zeros(m), yo,
                                         obs["var"], anamtype=3, inflation=1.02, anamtable=tab)
    assert ((E < xs[0]).any() and (E > xs[-1]).any())      # the clamping rule is exercised
    assert rel(xf, xfo) < 1e-14 and rel(xa, xao) < RTOL and rel(Ea, Eo) < RTOL


def test_assim_ensemble_per_variable_anamorphosis(ob):
    """anamtransform looks the transform up per element through the element's variable (assimilation.F90:4531-4567):
    zones holding three variables with different transforms (identity, log, tabulated), maxCorrection active."""
    from oak_b200 import synthetic
    g = synthetic.Grid(14, 10, 3)
    N, m = 20, 100
    rows = np.arange(g.n, dtype=np.int64)
    E = np.exp(0.4 * synthetic.ensemble_rows(np, g, rows, N, 11)).T.copy(order="F")
    xs = np.array([0.4, 0.8, 1.0, 1.5, 2.5])
    tab = np.column_stack([xs, np.sqrt(xs)])
    rowvar = 1 + (np.arange(g.n) % 3)          # the three rows of a zone belong to variables 1, 2, 3
    specs = [(1, None), (2, None), (3, tab)]
    obs = synthetic.observations(np, g, m, 11)
    Hi, Hj, Hs = synthetic.coo_operator(g, obs)
    yo = 1.1 + 0.2 * synthetic.normal(np, np.arange(m, dtype=np.int64), 8, 11)
    zx, zy = g.zone_xy(np, np.arange(g.nzones, dtype=np.int64))
    zs = np.full(g.nzones, 3, np.int32)
    sel = ob.Selector(zone_x=zx, zone_y=zy, corrLen=3000.0, maxLen=6000.0, obs_x=obs["ox"], obs_y=obs["oy"],
                      metrictype=0)
    maxc = np.full(g.n, 0.08)
    with ob.Handle(0) as h:
        with pytest.raises(ob.OakB200Error):       # anamtype 0 without the per-variable description
            h.configure(zs, sel)
            h.assim_ensemble(E, Hi, Hj, Hs, None, yo, ob.DiagCovar(obs["var"]), 0, 1.0, None)
        Ea, xf, xa = ob.assim_ensemble(zs, sel, E, Hi, Hj, Hs, None, yo, ob.DiagCovar(obs["var"]), inflation=1.03,
                                       maxCorrection=maxc, handle=h, anamvars=(rowvar, specs))
    oo = oracle.make_obs(m, obsx=obs["ox"], obsy=obs["oy"])
    Eo, xfo, xao = oracle.assim_ensemble(zs, dict(x=zx, y=zy), 3000.0, 6000.0, oo, E, Hi, Hj, Hs, np.zeros(m), yo,
                                         obs["var"], inflation=1.03, maxCorrection=maxc, anamvars=(rowvar - 1, specs))
    assert rel(xf, xfo) < 1e-14 and rel(xa, xao) < RTOL and rel(Ea, Eo) < RTOL
    assert not np.allclose(xf[1::3], E[1::3].mean(axis=1))     # variable 2 was transformed ...
    assert rel(xf[0::3], E[0::3].mean(axis=1)) < 1e-14          # ... variable 1 was not


@pytest.mark.parametrize("N", [2, 3, 5, 9, 33, 63])
def test_tiny_and_odd_ensemble_sizes(ob, N):
    """N = 2 (no Householder step at all), 3 (one), odd sizes and sizes just above / below the padded widths"""
    from oak_b200 import synthetic
    c = synthetic.small_case(nx=10, ny=8, nz=2, N=N, m=120, corr=3000.0, maxlen=6000.0, seed=N)
    with ob.Handle(0) as h:
        _configure(ob, h, c)
        xa, Sa, _, st = h.local_analysis(c["xf"], c["Hxf"], c["yo"], c["Sf"], c["HSf"], ob.DiagCovar(c["var"]))
    xo, So, _, mloc = _oracle_loc(c)
    assert rel(xa, xo) < RTOL and rel(Sa, So) < RTOL, (rel(xa, xo), rel(Sa, So), st["zones_fallback"])


def test_committed_golden_fixture(ob, handle):
    """the CUDA path against tests/golden/rrsqrt_known_answers.npz (closed-form known answers of
    test/test_rrsqrt.F90 + the oracle's Sa for the Gaspari-Cohn local case; tools/make_golden.py)"""
    import os
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "rrsqrt_known_answers.npz"))
    c = rrsqrt_case()
    zs = [1] * c["n"]
    xa, Sa, _ = ob.locanalysis(zs, _sel(ob, c, zs, 1, c["length"], 1e30), c["xf"], c["Hxf"], c["y"], c["Sf"],
                               c["HSf"], ob.DiagCovar(c["var"]), handle=handle)
    assert np.abs(xa - g["xa_gc_local"]).max() < TOL_REF
    assert rel(xa, g["xa_gc_local_oracle"]) < RTOL and rel(Sa, g["Sa_gc_local_oracle"]) < RTOL
    zs = [c["n"]]
    xa, Sa, _ = ob.locanalysis(zs, _sel(ob, c, zs, 2, 1.0, 1e30), c["xf"], c["Hxf"], c["y"], c["Sf"],
                               c["HSf"], ob.DiagCovar(c["var"]), handle=handle)
    assert np.abs(xa - g["xa_global"]).max() < TOL_REF and np.abs(Sa @ Sa.T - g["Pa_global"]).max() < TOL_REF
    assert rel(Sa, g["Sa_global_oracle"]) < RTOL
