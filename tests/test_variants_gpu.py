"""GPU parity tests of the variants prepared at the end of round 1 WITHOUT a GPU (options gram_kernel, fuse_apply,
tvec_split, push_pieces; all default off).  They passed under the CPU emulation of the kernel sources
(tools/cuemu, tests/test_emulated_kernels.py) but had not run on a device when they were written, so they live in
their own file, collected after the device-verified tests of test_gpu_parity.py.  Same tolerances."""
import numpy as np
import pytest

from test_gpu_parity import RTOL, _configure, _oracle_loc, ob, rel  # noqa: F401  (ob is a fixture)

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("variant", [1, 2, 3, 4], ids=["mma_4warps", "mma_2warps", "mma_4warps_ch32", "mma_2warps_ch32"])
@pytest.mark.parametrize("N,m,maxlen", [(64, 900, 6000.0), (40, 700, 9000.0), (64, 60, 4000.0)])
def test_gram_tensor_core_variants_match_the_register_tile_kernel(ob, variant, N, m, maxlen):
    # option gram_kernel = 1 / 2: G and c accumulated by mma.m8n8k4.f64 on the lower 8 x 8 tiles (gram_mma.cu).
    # Segments of every length 0..32 occur (k-steps of 4 rows padded with coef = 0), several chunks per zone at
    # the larger radius, zones without observations at the smaller one; N = 40 exercises the zero padding to 64.
    from oak_b200 import synthetic
    c = synthetic.small_case(nx=14, ny=12, nz=3, N=N, m=m, corr=maxlen / 2, maxlen=maxlen, seed=11 * N + m)
    xo, So, _, mloc = _oracle_loc(c)
    out = {}
    for gk in (0, variant):
        with ob.Handle(0, gram_kernel=gk, pad_to=64) as h:
            _configure(ob, h, c)
            xa, Sa, _, st = h.local_analysis(c["xf"], c["Hxf"], c["yo"], c["Sf"], c["HSf"], ob.DiagCovar(c["var"]))
        assert st["obs_relevant_sum"] == mloc.sum() and st["zones_skipped"] == (mloc == 0).sum()
        assert rel(xa, xo) < RTOL and rel(Sa, So) < RTOL, (gk, rel(xa, xo), rel(Sa, So))
        out[gk] = (xa, Sa)
    assert rel(out[variant][0], out[0][0]) < 1e-11 and rel(out[variant][1], out[0][1]) < 1e-11


@pytest.mark.parametrize("N,nz,m,gram", [(64, 30, 500, 0), (40, 13, 400, 1), (20, 5, 300, 0), (64, 3, 6, 0)],
                         ids=["N64_30rows", "N40_13rows_mma_gram", "N20_np32", "degenerate_fallback"])
def test_fused_apply_from_the_factored_transform(ob, N, nz, m, gram):
    """Option fuse_apply: k_tvec updates the zone rows itself, Sa_z = ((Sf_z - (Sf_z Y) Y^T) - a1 u_v^T) D - a2 u_w^T
    on mma.m8n8k4 tiles, without forming T; k_apply only serves the zones k_tvec did not finish (no observation,
    or handed to the Jacobi kernel).  Chunks of 8 rows: zone sizes 30 (3 full + 6), 13, 5, 3 and ragged 1..11;
    in place (Sa aliases Sf) as the Fortran caller does."""
    from oak_b200 import synthetic
    degenerate = m == 6
    if degenerate:
        c = synthetic.small_case(nx=6, ny=5, nz=nz, N=N, m=6, corr=1e9, maxlen=1e12, seed=5)
        Q, _ = np.linalg.qr(np.random.default_rng(3).normal(size=(N, 6)))
        rows = Q.T.copy()
        rows[:3] *= 2.0
        rows[5] *= 0.5
        c["HSf"] = np.asfortranarray(rows)
        c["var"] = np.full(6, 0.25)
        zs = None
    else:
        c = synthetic.small_case(nx=14, ny=6, nz=nz, N=N, m=m, corr=2500.0, maxlen=5000.0, seed=N + nz)
        # keep the observations of the left part only: zones on the right have none and keep the forecast
        keep = c["obs"]["ox"] < 5500.0
        for k in ("Hxf", "yo", "var"):
            c[k] = c[k][keep]
        c["HSf"] = np.asfortranarray(c["HSf"][keep])
        c["obs"] = {k: (v[..., keep] if v.ndim > 1 else v[keep]) for k, v in c["obs"].items()}
        c["m"] = int(keep.sum())
        zs = None
        if N == 40:  # ragged zones 1..11 rows with their own positions
            zs, left, k = [], c["Sf"].shape[0], 0
            while left > 0:
                sz = min(left, 1 + (k % 11)); zs.append(sz); left -= sz; k += 1
            zs = np.array(zs, np.int32)
            rng = np.random.default_rng(1)
            c["zx"] = rng.uniform(0, 14000, zs.size); c["zy"] = rng.uniform(0, 6000, zs.size)
            c["corr"] = np.full(zs.size, 2500.0); c["maxlen"] = np.full(zs.size, 5000.0)
    xo, So, _, mloc = _oracle_loc(c, zs)
    assert degenerate or ((mloc == 0).any() and (mloc > 0).any())
    with ob.Handle(0, eig_kernel=4, fuse_apply=1, gram_kernel=gram, pad_to=64 if gram else 0) as h:
        _configure(ob, h, c, zs)
        if degenerate:
            h.set_option("tri_maxgroup", 0)  # every zone goes to the Jacobi kernel and then through k_apply
        buf = np.asfortranarray(c["Sf"].copy())
        xa, Sa, _, st = h.local_analysis(c["xf"], c["Hxf"], c["yo"], buf, c["HSf"], ob.DiagCovar(c["var"]), out_Sa=buf)
    assert Sa is buf
    if degenerate:
        assert st["zones_fallback"] == len(mloc)
    assert rel(xa, xo) < RTOL and rel(Sa, So) < RTOL, (rel(xa, xo), rel(Sa, So))
    start = np.concatenate([[0], np.cumsum(c["zoneSize"] if zs is None else zs)])
    for z in np.nonzero(mloc == 0)[0]:
        assert (Sa[start[z]:start[z + 1]] == c["Sf"][start[z]:start[z + 1]]).all()


@pytest.mark.parametrize("N,fuse", [(64, 0), (40, 1), (24, 0)])
def test_eigenvector_kernel_split_in_two(ob, N, fuse):
    """Option tvec_split: the eigenvectors of T (twisted factorisations, grouping of close eigenvalues) come from a
    kernel of their own that builds W in global memory; the back-transformation and everything after it start from
    that W.  Same results as the single kernel, including zones the first half hands to the Jacobi kernel."""
    from oak_b200 import synthetic
    c = synthetic.small_case(nx=12, ny=6, nz=5, N=N, m=260, corr=2500.0, maxlen=5000.0, seed=3 * N)
    xo, So, _, mloc = _oracle_loc(c)
    out = []
    for split in (0, 1):
        with ob.Handle(0, eig_kernel=4, tvec_split=split, fuse_apply=fuse) as h:
            _configure(ob, h, c)
            xa, Sa, _, st = h.local_analysis(c["xf"], c["Hxf"], c["yo"], c["Sf"], c["HSf"], ob.DiagCovar(c["var"]))
        assert rel(xa, xo) < RTOL and rel(Sa, So) < RTOL, (split, rel(xa, xo), rel(Sa, So))
        out.append((xa, Sa, st["zones_fallback"]))
    assert rel(out[1][0], out[0][0]) < 1e-12 and rel(out[1][1], out[0][1]) < 1e-12 and out[0][2] == out[1][2]
    # degenerate spectrum: every zone is flagged by the first half and recomputed by the Jacobi kernel
    d = synthetic.small_case(nx=5, ny=4, nz=2, N=N, m=6, corr=1e9, maxlen=1e12, seed=5)
    Q, _ = np.linalg.qr(np.random.default_rng(3).normal(size=(N, 6)))
    rows = Q.T.copy(); rows[:3] *= 2.0; rows[5] *= 0.5
    d["HSf"] = np.asfortranarray(rows); d["var"] = np.full(6, 0.25)
    xo, So, _, mloc = _oracle_loc(d)
    with ob.Handle(0, eig_kernel=4, tvec_split=1, fuse_apply=fuse, tri_maxgroup=0) as h:
        _configure(ob, h, d)
        xa, Sa, _, st = h.local_analysis(d["xf"], d["Hxf"], d["yo"], d["Sf"], d["HSf"], ob.DiagCovar(d["var"]))
    assert st["zones_fallback"] == len(mloc) and rel(xa, xo) < RTOL and rel(Sa, So) < RTOL


@pytest.mark.emu_only
@pytest.mark.parametrize("pieces,fuse", [(1, 0), (3, 0), (4, 1)])
def test_pushes_in_pieces_under_emulation(ob, pieces, fuse):
    """Copy-engine flavour of the fused gather with the apply of a batch launched in pieces (option push_pieces),
    also combined with fuse_apply.  Host logic (piece ranges, pointer offsets into T / ampl / flags, the rows each
    push covers): checked under the CPU emulation, where device pointers are host pointers, so that numpy arrays
    can stand for the peers' result arrays.  The device version of this test is
    test_fused_gather_peer_outputs_receive_the_slab."""
    import ctypes as C
    from oak_b200 import synthetic, _lib
    c = synthetic.small_case(nx=14, ny=5, nz=4, N=24, m=150, corr=2500.0, maxlen=5000.0, seed=9)
    keep = c["obs"]["ox"] < 5500.0
    for k in ("Hxf", "yo", "var"):
        c[k] = c[k][keep]
    c["HSf"] = np.asfortranarray(c["HSf"][keep])
    c["obs"] = {k: (v[..., keep] if v.ndim > 1 else v[keep]) for k, v in c["obs"].items()}
    c["m"] = int(keep.sum())
    xo, So, _, mloc = _oracle_loc(c)
    n, N, m = c["Sf"].shape[0], 24, c["m"]
    row0, ntot = 17, n + 40                       # this rank's rows start at row 17 of the assembled arrays
    peers = [np.full((ntot, N), 7.0, order="F") for _ in range(2)]
    peers_x = [np.full(ntot, 7.0) for _ in range(2)]
    with ob.Handle(0, eig_kernel=4, fuse_apply=fuse, push_pieces=pieces, zones_per_batch=20) as h:
        _configure(ob, h, c)
        h.set_peer_outputs([p.ctypes.data for p in peers], [p.ctypes.data for p in peers_x], ntot, row0)
        Sf = np.asfortranarray(c["Sf"].copy())
        xa, Sa = np.empty(n), np.empty((n, N), order="F")
        ptr = lambda a: C.c_void_p(np.ascontiguousarray(a).ctypes.data) if not a.flags.f_contiguous else C.c_void_p(a.ctypes.data)
        arrs = [np.ascontiguousarray(c[k], dtype=np.float64) for k in ("xf", "Hxf", "yo")]
        HSf, var = np.asfortranarray(c["HSf"]), np.ascontiguousarray(c["var"])
        st = _lib.Stats()
        rc = h._L.oakb200_local_analysis_dev(h._h, n, N, m, ptr(arrs[0]), ptr(arrs[1]), ptr(arrs[2]), ptr(Sf), n, ptr(HSf),
                                             max(m, 1), ptr(var), None, ptr(xa), ptr(Sa), n, None, None, C.byref(st))
        assert rc == 0, h._L.oakb200_last_error()
    assert (mloc == 0).any() and rel(xa, xo) < RTOL and rel(Sa, So) < RTOL
    for P, px in zip(peers, peers_x):
        assert (P[row0:row0 + n] == Sa).all() and (px[row0:row0 + n] == xa).all()
        assert (P[:row0] == 7.0).all() and (P[row0 + n:] == 7.0).all() and (px[:row0] == 7.0).all()




@pytest.mark.parametrize("eig_kernel", [4, 0])
def test_global_scheme_known_answers_and_oracle(ob, eig_kernel):
    """analysis (rrsqrt.F90:196-208) through oakb200_global_analysis: the reference's own known answers
    (test/test_rrsqrt.F90:57-74: Kalman gain form of xa and Pa, tol 1e-8) and the oracle at 1e-9; then a larger case
    with excluded observations (DCDCovar), several partial Gram matrices, several row blocks and several host chunks."""
    import oracle
    from refcases import kalman_check, rrsqrt_case
    from test_gpu_parity import TOL_REF
    c = rrsqrt_case()
    xa_check, Pa_check = kalman_check(c["xf"], c["Sf"], c["H"], c["y"], np.diag(c["var"]))
    with ob.Handle(0, eig_kernel=eig_kernel, pad_to=64) as h:
        xa, Sa, ampl, st = h.global_analysis(c["xf"], c["Hxf"], c["y"], c["Sf"], c["HSf"], ob.DiagCovar(c["var"]))
    assert np.abs(xa - xa_check).max() < TOL_REF and np.abs(Sa @ Sa.T - Pa_check).max() < TOL_REF
    xo, So, ao = oracle.analysis(c["xf"], c["Hxf"], c["y"], c["Sf"], c["HSf"], c["var"])
    assert rel(xa, xo) < RTOL and rel(Sa, So) < RTOL and rel(ampl, ao) < 1e-8
    assert st["obs_relevant_sum"] == c["m"]

    from oak_b200 import synthetic
    for N in (40, 100):
        d = synthetic.small_case(nx=21, ny=17, nz=4, N=N, m=700, corr=3000.0, maxlen=6000.0, seed=N)
        e01 = (np.random.default_rng(1).uniform(size=d["m"]) > 0.25).astype(np.float64)
        var = np.where(e01 > 0, d["var"], 1e300)    # an excluded observation = infinite variance for the oracle
        xo, So, ao = oracle.analysis(d["xf"], d["Hxf"], d["yo"], d["Sf"], d["HSf"], var)
        with ob.Handle(0, eig_kernel=eig_kernel, chunk_mb=0.2) as h:
            buf = np.asfortranarray(d["Sf"].copy())
            xa, Sa, ampl, st = h.global_analysis(d["xf"], d["Hxf"], d["yo"], buf, d["HSf"],
                                                 ob.DCDCovar(e01, ob.DiagCovar(d["var"])), out_Sa=buf)
            assert rel(xa, xo) < RTOL and rel(Sa, So) < RTOL, (N, rel(xa, xo), rel(Sa, So))
            # no observation: the forecast comes back unchanged
            xa0, Sa0, a0, st0 = h.global_analysis(d["xf"], np.zeros(0), np.zeros(0), d["Sf"], np.zeros((0, N)),
                                                  ob.DiagCovar(np.zeros(0)))
            assert (xa0 == d["xf"]).all() and (Sa0 == d["Sf"]).all() and (a0 == 0).all()
    # the module-level mirror of the reference call
    xa, Sa, ampl = ob.analysis(c["xf"], c["Hxf"], c["y"], c["Sf"], c["HSf"], ob.DiagCovar(c["var"]))
    assert np.abs(xa - xa_check).max() < TOL_REF


def test_assim_case_through_the_global_scheme(ob):
    # BASELINE config 1, test/test_assim.F90:96-172 (tol 1e-5): ensemble in, ensemble out with schemetype = 0
    import oracle
    from refcases import assim_case, kalman_check
    c = assim_case()
    n, N = c["n"], c["N"]
    xf = c["Ef"].sum(axis=1) / N
    xa_check, Pa_check = kalman_check(xf, (c["Ef"] - xf[:, None]) / np.sqrt(N - 1.0), c["H"], c["yo"],
                                      np.diag(c["var"]))
    Ea, xf_o, xa_o = ob.assim_ensemble(None, None, c["Ef"], [1], [5], [1.0], np.zeros(1), c["yo"],
                                       ob.DiagCovar(c["var"]))
    xa = Ea.sum(axis=1) / N
    Eap = Ea - xa[:, None]
    assert np.abs(xa - xa_check).max() < 1e-5
    assert np.abs(Eap @ Eap.T / (N - 1.0) - Pa_check).max() < 1e-5
    obs = oracle.make_obs(1, obsx=c["obsx"], obsy=c["obsy"], weightfun=2)
    Eo, xfo, xao = oracle.assim_ensemble([n], dict(x=c["x"][:1], y=c["y"][:1]), 1.0, 1e30, obs, c["Ef"],
                                         np.array([1], np.int32), np.array([5], np.int32), np.array([1.0]),
                                         np.zeros(1), c["yo"], c["var"])
    assert rel(Ea, Eo) < RTOL and rel(xf_o, xfo) < 1e-14 and rel(xa_o, xao) < RTOL


def test_apply_on_tensor_core_tiles(ob):
    """Option apply_kernel = 1 (k_apply_mma): local scheme with ragged zones of 1..75 rows (chunks of 32: partial,
    exact, several), N = 64 and N = 40 (padding), in place, zones without observations; and the global scheme,
    whose row blocks share one transform."""
    import oracle
    from oak_b200 import synthetic
    for N in (64, 40):
        c = synthetic.small_case(nx=14, ny=6, nz=25, N=N, m=300, corr=2500.0, maxlen=5000.0, seed=N + 1)
        keep = c["obs"]["ox"] < 5500.0
        for k in ("Hxf", "yo", "var"):
            c[k] = c[k][keep]
        c["HSf"] = np.asfortranarray(c["HSf"][keep])
        c["obs"] = {k: (v[..., keep] if v.ndim > 1 else v[keep]) for k, v in c["obs"].items()}
        c["m"] = int(keep.sum())
        zs, left, k = [], c["Sf"].shape[0], 0
        sizes = [1, 7, 31, 32, 33, 64, 75, 12]
        while left > 0:
            sz = min(left, sizes[k % len(sizes)]); zs.append(sz); left -= sz; k += 1
        zs = np.array(zs, np.int32)
        rng = np.random.default_rng(2)
        c["zx"] = rng.uniform(0, 14000, zs.size); c["zy"] = rng.uniform(0, 6000, zs.size)
        c["corr"] = np.full(zs.size, 2500.0); c["maxlen"] = np.full(zs.size, 5000.0)
        xo, So, _, mloc = _oracle_loc(c, zs)
        assert (mloc == 0).any() and (mloc > 0).any()
        with ob.Handle(0, eig_kernel=4, apply_kernel=1, pad_to=64) as h:
            _configure(ob, h, c, zs)
            buf = np.asfortranarray(c["Sf"].copy())
            xa, Sa, _, st = h.local_analysis(c["xf"], c["Hxf"], c["yo"], buf, c["HSf"], ob.DiagCovar(c["var"]), out_Sa=buf)
        assert rel(xa, xo) < RTOL and rel(Sa, So) < RTOL, (N, rel(xa, xo), rel(Sa, So))
        start = np.concatenate([[0], np.cumsum(zs)])
        for z in np.nonzero(mloc == 0)[0]:
            assert (Sa[start[z]:start[z + 1]] == c["Sf"][start[z]:start[z + 1]]).all()
        # global scheme on the same arrays
        xo, So, ao = oracle.analysis(c["xf"], c["Hxf"], c["yo"], c["Sf"], c["HSf"], c["var"])
        with ob.Handle(0, apply_kernel=1, pad_to=64, chunk_mb=0.3) as h:
            xa, Sa, ampl, _ = h.global_analysis(c["xf"], c["Hxf"], c["yo"], c["Sf"], c["HSf"], ob.DiagCovar(c["var"]))
        assert rel(xa, xo) < RTOL and rel(Sa, So) < RTOL, (N, rel(xa, xo), rel(Sa, So))
