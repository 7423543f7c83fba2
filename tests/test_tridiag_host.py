"""The scalar routines of the tridiagonal transform route (oak_b200/csrc/tridiag_math.cuh: QL / Pal-Walker-Kahan
eigenvalues, twisted-factorisation eigenvectors) are __host__ __device__: compiled here with g++
(tools/tridiag_host.cpp) and checked against numpy on the CPU; the full numpy prototype of the route
(tools/proto_tridiag.py) is run with them on a few matrices.  The CUDA kernels that call them are covered by
the GPU parity tests."""
import ctypes
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))


@pytest.fixture(scope="module")
def hostlib(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("tridiag") / "libtridiag_host.so")
    subprocess.check_call(["g++", "-O2", "-shared", "-fPIC", os.path.join(ROOT, "tools", "tridiag_host.cpp"), "-o", out])
    lib = ctypes.CDLL(out)
    dp = ctypes.POINTER(ctypes.c_double)
    for f in (lib.host_tql, lib.host_pwk):
        f.argtypes = [ctypes.c_int, dp, dp, ctypes.c_int, ctypes.c_double]
    for f in (lib.host_twisted, lib.host_twisted2):
        f.restype = ctypes.c_double
        f.argtypes = [ctypes.c_int, dp, dp, ctypes.c_int, ctypes.c_double, ctypes.c_double, dp, ctypes.c_int, dp]
    return lib


def _tridiag(n, seed, rank=None):
    rng = np.random.default_rng(seed)
    A = rng.normal(size=(rank or 3 * n, n))
    G = A.T @ A
    import proto_tridiag as p
    d, e, V, tau = p.householder_tridiag(G)
    return d, e, G


def _eigvals(f, d, e, stride=3):
    dp = ctypes.POINTER(ctypes.c_double)
    n = len(d)
    db = np.zeros(n * stride); eb = np.zeros(n * stride)
    db[::stride] = d; eb[:(n - 1) * stride:stride] = e
    tn = max(np.abs(d).max(), np.abs(e).max() if n > 1 else 0.0)
    rot = f(n, db.ctypes.data_as(dp), eb.ctypes.data_as(dp), stride, tn)
    return db[::stride].copy(), rot


@pytest.mark.parametrize("n,rank", [(2, None), (3, None), (16, None), (64, None), (64, 5), (40, 39)])
def test_ql_and_pwk_eigenvalues(hostlib, n, rank):
    d, e, G = _tridiag(n, 100 + n, rank)
    ref = np.linalg.eigvalsh(G)
    for f in (hostlib.host_tql, hostlib.host_pwk):
        lam, rot = _eigvals(f, d, e)
        assert rot >= 0
        assert (np.diff(lam) >= 0).all()
        assert np.abs(lam - ref).max() <= 2e-14 * np.abs(ref).max()


def test_pwk_handles_zero_diagonal_and_split_matrices(hostlib):
    lam, rot = _eigvals(hostlib.host_pwk, np.zeros(8), np.zeros(7))
    assert rot == 0 and (lam == 0).all()
    d = np.array([3.0, 1.0, 2.0, 0.0, 0.0, 5.0]); e = np.array([0.5, 0.0, 0.25, 0.0, 0.0])
    T = np.diag(d) + np.diag(e, 1) + np.diag(e, -1)
    lam, _ = _eigvals(hostlib.host_pwk, d, e)
    assert np.abs(lam - np.linalg.eigvalsh(T)).max() < 1e-14


def test_twisted_factorisation_vectors(hostlib):
    dp = ctypes.POINTER(ctypes.c_double)
    d, e, G = _tridiag(48, 7)
    T = np.diag(d) + np.diag(e, 1) + np.diag(e, -1)
    lam = np.linalg.eigvalsh(T)
    tn = max(np.abs(d).max(), np.abs(e).max())
    n, sw = len(d), 5
    W = np.zeros((n, n))
    for j in range(n):
        w = np.zeros(n * sw); gam = ctypes.c_double()
        zz = hostlib.host_twisted(n, np.ascontiguousarray(d).ctypes.data_as(dp), np.ascontiguousarray(e).ctypes.data_as(dp), 1,
                                  lam[j], tn * 1e-150, w.ctypes.data_as(dp), sw, ctypes.byref(gam))
        z = w[::sw] / np.sqrt(zz)
        assert np.abs(T @ z - lam[j] * z).max() <= 1e-13 * tn      # residual = |gamma_r| / |z|
        assert abs(gam.value) / np.sqrt(zz) <= 1e-13 * tn
        W[:, j] = z
    gap = np.diff(lam).min() / tn
    assert np.abs(W.T @ W - np.eye(n)).max() <= 1e-14 / gap        # orthogonality ~ eps |T| / gap


def test_numpy_prototype_of_the_route(hostlib, monkeypatch):
    """(I+G)^-1/2 from Householder + PWK + twisted vectors + grouped Gram-Schmidt equals the eigh-based one"""
    import proto_tridiag as p
    from proto_jacobi import make_G
    worst = 0.0
    for N, mloc, ws in [(64, 200, 1.0), (64, 30, 1.0), (20, 5, 1.0), (64, 200, 10.0)]:
        G = make_G(N, mloc, 3, ws)
        G = 0.5 * (G + G.T)
        M, lam, info = p.transform_tridiag(G, passes=1)
        Mr = p.ref_M(G)
        worst = max(worst, np.abs(M - Mr).max() / np.abs(Mr).max())
    assert worst < 1e-11


@pytest.mark.parametrize("n", [1, 2, 3, 5, 16, 33, 64])
def test_twisted_vector2_equals_twisted_vector(hostlib, n):
    """the variant with interleaved pivot recurrences and stored reciprocals (prepared for k_tvec,
    -DTVEC_TWISTED2=1) gives the same vectors and residuals"""
    dp = ctypes.POINTER(ctypes.c_double)
    if n == 1:
        d, e = np.array([2.5]), np.zeros(0)
    else:
        d, e, _ = _tridiag(n, 31 + n)
    T = np.diag(d) + (np.diag(e, 1) + np.diag(e, -1) if n > 1 else 0.0)
    lam = np.linalg.eigvalsh(T)
    tn = max(np.abs(d).max(), np.abs(e).max() if n > 1 else 0.0)
    eb = np.ascontiguousarray(np.concatenate([e, [0.0]]))
    for j in range(n):
        out = []
        for f in (hostlib.host_twisted, hostlib.host_twisted2):
            w = np.zeros(n * 2); gam = ctypes.c_double()
            zz = f(n, np.ascontiguousarray(d).ctypes.data_as(dp), eb.ctypes.data_as(dp), 1, lam[j], tn * 1e-150,
                   w.ctypes.data_as(dp), 2, ctypes.byref(gam))
            out.append((w[::2] / np.sqrt(zz), abs(gam.value) / np.sqrt(zz)))
        (z1, r1), (z2, r2) = out
        assert r2 <= 1e-13 * tn and np.abs(T @ z2 - lam[j] * z2).max() <= 1e-13 * tn
        gapj = min([abs(lam[j] - lam[k]) for k in range(n) if k != j] + [tn]) / tn
        assert min(np.abs(z1 - z2).max(), np.abs(z1 + z2).max()) <= 1e-13 / max(gapj, 1e-12)


def test_pwk_loop_forms_give_identical_bits(tmp_path):
    """k_tql's QL iteration in its forms — loop nest (PWK_FLAT=0) / one flat loop over sweeps (default), direct reads /
    register prefetch queue (pwk_eigenvalues_t<6>, TQL_GLOBAL) — performs the same operations on the same values per
    zone: eigenvalues and rotation counts must be IDENTICAL, also on matrices that split (tracked block ends, the
    bookkeeping that replaced the rescan at every l)."""
    dp = ctypes.POINTER(ctypes.c_double)
    fns = []
    for k, flags in enumerate((["-DPWK_FLAT=0"], ["-DPWK_FLAT=1"])):
        out = str(tmp_path / f"libtri{k}.so")
        subprocess.check_call(["g++", "-O2", "-shared", "-fPIC", "-ffp-contract=off"] + flags +
                              [os.path.join(ROOT, "tools", "tridiag_host.cpp"), "-o", out])
        lib = ctypes.CDLL(out)
        for f in (lib.host_pwk, lib.host_pwk_pf):
            f.argtypes = [ctypes.c_int, dp, dp, ctypes.c_int, ctypes.c_double]
            fns.append(f)
    rng = np.random.default_rng(5)
    for t in range(400):
        n = int(rng.integers(2, 65))
        A = rng.normal(size=(int(rng.integers(1, 3 * n)), n)) * 10 ** rng.uniform(-3, 3)
        G = A.T @ A
        if t % 5 == 0:
            G = np.diag(rng.uniform(0, 1, n))
        if t % 11 == 0:     # repeated diagonal blocks: many splits
            k = n // 4 + 1
            B = rng.normal(size=(k, k))
            G = np.kron(np.eye(4), B @ B.T)[:n, :n].copy()
        import scipy.linalg as sl
        H = sl.hessenberg(G)
        d = np.diag(H).copy(); e = np.append(np.diag(H, -1), 0.0)
        if t % 7 == 0 and n > 4:
            e[n // 2] = 0.0; e[n // 3] = 0.0
        if t % 19 == 0:
            d = rng.normal(size=n); e = np.append(rng.normal(size=n - 1) * np.where(rng.uniform(size=n - 1) < 0.3, 1e-18, 1.0), 0.0)
        tn = max(np.abs(d).max(), np.abs(e).max())
        res = []
        for f in fns:
            dd, ee = d.copy(), e.copy()
            rc = f(n, dd.ctypes.data_as(dp), ee.ctypes.data_as(dp), 1, tn)
            res.append((rc, dd))
        for rc, dd in res[1:]:
            assert rc == res[0][0] and np.array_equal(dd, res[0][1])
        T = np.diag(d) + np.diag(e[:-1], 1) + np.diag(e[:-1], -1)
        assert np.abs(res[0][1] - np.linalg.eigvalsh(T)).max() <= 3e-14 * max(tn, 1e-300)
