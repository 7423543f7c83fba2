"""Prototype (numpy, scalar loops where the CUDA code is scalar) of the tridiagonal route of the
per-zone transform (eig_kernel = 4):
   G = Q T Q^T (Householder) ; eigenvalues of T by implicit QL (no vectors) ;
   eigenvectors of T by the twisted (double) factorisation of T - lambda I with one Rayleigh
   correction ; Gram-Schmidt inside groups of close eigenvalues ; U = Q W ;
   (I+G)^-1/2 = I + sum_j g_j u_j u_j^T , g_j = (1+lambda_j)^-1/2 - 1.
Decides thresholds (null eigenvalues, close groups, fallback to the Jacobi kernel)."""
import numpy as np, sys
from proto_jacobi import make_G

EPS = 2.0 ** -52


def householder_tridiag(G):
    """thread-per-row style (full square updates), lower form: reduces column k below the subdiagonal.
    Returns d, e, V (V[:,k] = reflector k, zero for rows <= k, V[k+1,k] = 1 convention not used: plain v), tau"""
    A = G.copy()
    n = A.shape[0]
    V = np.zeros((n, n)); tau = np.zeros(n)
    for k in range(n - 2):
        x = A[k + 1:, k].copy()
        alpha = x[0]
        xn2 = x[1:] @ x[1:]
        if xn2 == 0.0:
            continue  # already tridiagonal in this column (tau = 0)
        nrm = np.sqrt(alpha * alpha + xn2)
        beta = -np.copysign(nrm, alpha)
        t = (beta - alpha) / beta
        v = x / (alpha - beta); v[0] = 1.0
        V[k + 1:, k] = v; tau[k] = t
        # A22 <- H A22 H,  H = I - t v v^T
        A22 = A[k + 1:, k + 1:]
        p = t * (A22 @ v)
        w = p - (0.5 * t * (p @ v)) * v
        A22 -= np.outer(v, w) + np.outer(w, v)
        A[k + 1, k] = beta; A[k, k + 1] = beta
        A[k + 2:, k] = 0; A[k, k + 2:] = 0
    d = np.diag(A).copy(); e = np.diag(A, -1).copy()
    return d, e, V, tau


def tql_eigenvalues(d, e):
    """implicit QL without vectors (EISPACK tql1 / NR tqli); returns ascending eigenvalues"""
    d = d.copy(); n = len(d)
    e = np.concatenate([e, [0.0]])
    steps = 0
    tn = max(np.abs(d).max(), np.abs(e).max()) if n > 1 else abs(d[0])
    for l in range(n):
        it = 0
        while True:
            m = l
            while m < n - 1:
                dd = abs(d[m]) + abs(d[m + 1])
                if abs(e[m]) <= EPS * dd or abs(e[m]) <= 0.5 * EPS * tn:
                    break
                m += 1
            if m == l:
                break
            it += 1
            if it > 60:
                raise RuntimeError("tql no convergence")
            g = (d[l + 1] - d[l]) / (2.0 * e[l])
            r = np.hypot(g, 1.0)
            g = d[m] - d[l] + e[l] / (g + np.copysign(r, g))
            s = c = 1.0; p = 0.0
            i = m - 1
            brk = False
            while i >= l:
                steps += 1
                f = s * e[i]; b = c * e[i]
                r = np.hypot(f, g)
                e[i + 1] = r
                if r == 0.0:
                    d[i + 1] -= p; e[m] = 0.0; brk = True
                    break
                s = f / r; c = g / r
                g = d[i + 1] - p
                r = (d[i] - g) * s + 2.0 * c * b
                p = s * r
                d[i + 1] = g + p
                g = c * r - b
                i -= 1
            if brk:
                continue
            d[l] -= p; e[l] = g; e[m] = 0.0
    return np.sort(d), steps


def twisted_vector(d, e, lam, pivmin, passes=2):
    """eigenvector of tridiag(d,e) for eigenvalue lam by the double factorisation; returns z (unit), lam, resid"""
    n = len(d)
    e2 = e * e
    for ps in range(passes):
        dp = np.empty(n)  # forward pivots delta+
        dp[0] = d[0] - lam
        for i in range(n - 1):
            piv = dp[i]
            if abs(piv) < pivmin: piv = -pivmin
            dp[i] = piv
            dp[i + 1] = (d[i + 1] - lam) - e2[i] / piv
        dm = np.empty(n)
        dm[n - 1] = d[n - 1] - lam
        for i in range(n - 2, -1, -1):
            piv = dm[i + 1]
            if abs(piv) < pivmin: piv = -pivmin
            dm[i + 1] = piv
            dm[i] = (d[i] - lam) - e2[i] / piv
        gam = dp + dm - (d - lam)
        r = int(np.argmin(np.abs(gam)))
        z = np.zeros(n); z[r] = 1.0
        for i in range(r, 0, -1):
            z[i - 1] = -e[i - 1] * z[i] / dp[i - 1]
        for i in range(r, n - 1):
            z[i + 1] = -e[i] * z[i] / dm[i + 1]
        zz = z @ z
        corr = gam[r] / zz
        resid = abs(gam[r]) / np.sqrt(zz)
        if ps < passes - 1:
            lam = lam + corr
    return z / np.sqrt(zz), lam, resid


def transform_tridiag(G, gtol=1e-3, verbose=False, passes=2):
    n = G.shape[0]
    d, e, V, tau = householder_tridiag(G)
    lam, steps = tql_eigenvalues(d, e)
    tn = max(np.abs(d).max(), np.abs(e).max() if n > 1 else 0.0)  # scale
    pivmin = max(tn * 1e-290, 1e-300) if tn > 0 else 1e-300
    pivmin = max(np.finfo(float).tiny * max(1.0, (e * e).max() if n > 1 else 1.0), 1e-300)
    g = 1.0 / np.sqrt(1.0 + np.maximum(lam, 0.0)) - 1.0
    W = np.zeros((n, n))
    null = np.abs(g) < 1e-13
    res = np.zeros(n)
    lam2 = lam.copy()
    for j in range(n):
        if null[j]:
            continue
        W[:, j], lam2[j], res[j] = twisted_vector(d, e, lam[j], pivmin, passes)
    # groups of close eigenvalues (ascending): Gram-Schmidt inside a group
    ngs = 0; worst = 0.0
    j = 0
    nfall = 0
    while j < n:
        k = j
        while k + 1 < n and (not null[k + 1]) and (not null[k]) and lam[k + 1] - lam[k] <= gtol * tn:
            k += 1
        if k > j:
            for a in range(j + 1, k + 1):
                for b in range(j, a):
                    dot = W[:, b] @ W[:, a]
                    W[:, a] -= dot * W[:, b]
                    ngs += 1
                nr = np.sqrt(W[:, a] @ W[:, a])
                worst = max(worst, 1.0 - nr)
                if nr < 1e-3:
                    nfall += 1
                W[:, a] /= nr
        j = k + 1
    # back-transform U = H_0 H_1 ... H_{n-3} W
    U = W.copy()
    for k in range(n - 3, -1, -1):
        if tau[k] == 0.0: continue
        v = V[:, k]
        U -= np.outer(tau[k] * v, v @ U)
    Y = U * np.sqrt(-g)
    M = np.eye(n) - Y @ Y.T
    info = dict(steps=steps, ngs=ngs, nfall=nfall, maxres=res.max() / max(tn, 1e-300), tn=tn,
                orth=np.abs(U[:, ~null].T @ U[:, ~null] - np.eye((~null).sum())).max() if (~null).any() else 0.0)
    return M, lam, info


def ref_M(G):
    lam, U = np.linalg.eigh(G)
    lam = np.maximum(lam, 0)
    return (U / np.sqrt(1 + lam)) @ U.T


def degenerate_G(N, vals, seed):
    rng = np.random.default_rng(seed)
    Q, _ = np.linalg.qr(rng.normal(size=(N, N)))
    lam = np.zeros(N); lam[:len(vals)] = vals
    return (Q * lam) @ Q.T


if __name__ == "__main__":
    cases = []
    for N, mloc, ws in [(64, 200, 1), (64, 200, 10), (64, 200, 0.01), (64, 30, 1), (64, 3, 1), (64, 1000, 3), (128, 1257, 1), (20, 5, 1), (12, 5, 1), (16, 40, 1)]:
        for s in range(4 if N <= 64 else 2):
            cases.append((f"N={N} mloc={mloc} ws={ws} s={s}", make_G(N, mloc, s, ws)))
    cases.append(("deg 5,5,5,2,2", degenerate_G(64, [5, 5, 5, 2, 2], 1)))
    cases.append(("deg 63x7", degenerate_G(64, [7.0] * 63, 2)))
    cases.append(("deg close 1e-6", degenerate_G(64, [5, 5 + 5e-6, 3, 3 + 3e-9, 1, 1 + 1e-12], 3)))
    cases.append(("zero", np.zeros((64, 64))))
    cases.append(("diag", np.diag(np.arange(64.0))))
    cases.append(("identity*3", 3 * np.eye(32)))
    worst = 0
    for name, G in cases:
        G = 0.5 * (G + G.T)
        M, lam, info = transform_tridiag(G)
        Mr = ref_M(G)
        err = np.abs(M - Mr).max() / np.abs(Mr).max()
        worst = max(worst, err)
        print(f"{name:32s} err={err:.2e} steps={info['steps']} gs={info['ngs']} fall={info['nfall']} res={info['maxres']:.1e} orth={info['orth']:.1e}")
    print("worst", worst)


def host_check():
    """the C++ routines of oak_b200/csrc/tridiag_math.cuh compiled for the host, inside the same pipeline"""
    import ctypes
    lib = ctypes.CDLL("/tmp/libtridiag_host.so")
    dp = ctypes.POINTER(ctypes.c_double)
    lib.host_twisted.restype = ctypes.c_double
    lib.host_twisted.argtypes = [ctypes.c_int, dp, dp, ctypes.c_int, ctypes.c_double, ctypes.c_double, dp, ctypes.c_int, dp]
    lib.host_tql.argtypes = [ctypes.c_int, dp, dp, ctypes.c_int, ctypes.c_double]
    global tql_eigenvalues, twisted_vector

    def tql_c(d, e):
        n = len(d); s = 3
        db = np.zeros(n * s); eb = np.zeros(n * s)
        db[::s] = d; eb[:(n - 1) * s:s] = e
        tn = max(np.abs(d).max(), np.abs(e).max()) if n > 1 else abs(d[0])
        rot = lib.host_tql(n, db.ctypes.data_as(dp), eb.ctypes.data_as(dp), s, tn)
        assert rot >= 0
        return db[::s].copy(), rot

    def tw_c(d, e, lam, pivmin, passes=1):
        n = len(d); sw = 5
        w = np.zeros(n * sw); gam = ctypes.c_double()
        d = np.ascontiguousarray(d); e = np.ascontiguousarray(e)
        zz = lib.host_twisted(n, d.ctypes.data_as(dp), e.ctypes.data_as(dp), 1, lam, pivmin, w.ctypes.data_as(dp), sw, ctypes.byref(gam))
        z = w[::sw].copy()
        return z / np.sqrt(zz), lam, abs(gam.value) / np.sqrt(zz)
    tql_eigenvalues = tql_c
    twisted_vector = tw_c


if __name__ == "__main__" and "--host" in sys.argv:
    host_check()
    worst = 0
    for N, mloc, ws in [(64, 200, 1), (64, 200, 10), (64, 30, 1), (64, 3, 1), (64, 1000, 3), (128, 1257, 1), (20, 5, 1), (16, 40, 1)]:
        for s in range(4 if N <= 64 else 2):
            G = make_G(N, mloc, s + 50, ws); G = 0.5 * (G + G.T)
            M, lam, info = transform_tridiag(G, passes=1)
            lr = np.linalg.eigvalsh(G)
            Mr = ref_M(G)
            err = np.abs(M - Mr).max() / np.abs(Mr).max()
            worst = max(worst, err)
            print(f"host N={N} mloc={mloc} ws={ws} err={err:.2e} lamerr={np.abs(lam-lr).max()/np.abs(lr).max():.1e} rot={info['steps']} res={info['maxres']:.1e}")
    print("host worst", worst)
