"""Prototype (numpy) of the eigen-free route used by the CUDA eig kernel:
   A = I + G = L L^T ; one-sided Jacobi on the columns of W=L (odd-even block ordering,
   blocks of 2 columns) ; Z = W V = U Sigma ; M = (I+G)^-1/2 = Z diag(sigma^-3) Z^T.
Used to choose sweep counts / stopping rule before writing the kernel."""
import numpy as np, sys

def make_G(N=64, mloc=200, seed=0, wscale=1.0):
    rng = np.random.default_rng(seed)
    A = rng.normal(size=(mloc, N)) * 0.5
    A -= A.mean(1, keepdims=True)
    A /= np.sqrt(N - 1.0)
    r = (0.05 * (1 + 0.5 * rng.uniform(size=mloc))) ** 2
    d = rng.uniform(0, 8e3, size=mloc) * np.sqrt(rng.uniform(size=mloc))
    w = np.exp(-(d / 4e3) ** 2) * wscale
    coef = w * w / r
    return A.T @ (coef[:, None] * A)

def jacobi_oe(W, tol_stop=1e-7, max_sweeps=12, float_params=True, verbose=False):
    """odd-even transposition ordering on blocks of 2 columns; returns Z, sweeps, history"""
    N = W.shape[1]
    W = W.copy()
    NB = N // 2
    pos = [[2 * b, 2 * b + 1] for b in range(NB)]  # block -> its two column indices
    hist = []
    def rot(p, q, st):
        x = W[:, p].copy(); y = W[:, q].copy()
        a = x @ x; b = y @ y; g = x @ y
        cosang = abs(g) / np.sqrt(a * b)
        st[0] = max(st[0], cosang)
        if cosang < 1e-15:
            return
        d = b - a; g2 = 2 * g
        if float_params:
            d = np.float32(d); g2 = np.float32(g2)
            h = np.sqrt(d * d + g2 * g2)
            t = float(g2 / (d + np.copysign(h, d))) if d != 0 else float(np.sign(g2))
        else:
            h = np.hypot(d, g2)
            t = g2 / (d + np.copysign(h, d)) if d != 0 else np.sign(g2)
        c = 1 / np.sqrt(1 + t * t); s = t * c
        W[:, p] = c * x - s * y
        W[:, q] = s * x + c * y
    for sweep in range(max_sweeps):
        st = [0.0]
        # intra-block pairs
        for b in range(NB):
            rot(pos[b][0], pos[b][1], st)
        for step in range(NB):
            start = step % 2
            for b in range(start, NB - 1, 2):
                A_, B_ = pos[b], pos[b + 1]
                rot(A_[0], B_[0], st); rot(A_[1], B_[1], st)
                rot(A_[0], B_[1], st); rot(A_[1], B_[0], st)
                pos[b], pos[b + 1] = pos[b + 1], pos[b]  # swap positions
        hist.append(st[0])
        if verbose: print("sweep", sweep, "max cos", st[0])
        if st[0] < tol_stop:
            break
    return W, sweep + 1, hist

def check(N=64, mloc=200, seed=0, wscale=1.0, **kw):
    G = make_G(N, mloc, seed, wscale)
    lam, U = np.linalg.eigh(G)
    lam = np.maximum(lam, 0)
    Mref = (U / np.sqrt(1 + lam)) @ U.T
    L = np.linalg.cholesky(np.eye(N) + G)
    Z, ns, hist = jacobi_oe(L, **kw)
    s2 = (Z * Z).sum(0)
    M = (Z / (s2 * np.sqrt(np.maximum(s2, 1)))) @ Z.T
    err = np.abs(M - Mref).max() / np.abs(Mref).max()
    return ns, err, hist, lam.max()

if __name__ == "__main__":
    for N, mloc, ws in [(64, 200, 1), (64, 200, 10), (64, 30, 1), (64, 1000, 3), (128, 1257, 1), (20, 5, 1), (12, 5, 1)]:
        N2 = N + (N % 2)
        for tol in (1e-6, 1e-7, 1e-8):
            res = [check(N2 if N2 == N else N2, mloc, s, ws, tol_stop=tol) for s in range(3 if N > 64 else 6)]
            print(f"N={N} mloc={mloc} ws={ws} tol={tol:g}: sweeps={[r[0] for r in res]} maxerr={max(r[1] for r in res):.2e} lammax={res[0][3]:.1f} hist0={['%.1e'%h for h in res[0][2]]}")

def pivoted_chol(A):
    A = A.copy(); n = A.shape[0]; perm = np.arange(n); L = np.zeros_like(A)
    for j in range(n):
        p = j + np.argmax(np.diag(A)[j:])
        if p != j:
            A[[j, p]] = A[[p, j]]; A[:, [j, p]] = A[:, [p, j]]; L[[j, p]] = L[[p, j]]; perm[[j, p]] = perm[[p, j]]
        L[j, j] = np.sqrt(A[j, j]); L[j+1:, j] = A[j+1:, j] / L[j, j]
        A[j+1:, j+1:] -= np.outer(L[j+1:, j], L[j+1:, j])
    return L, perm

def variants(N=64, mloc=200, seed=0, ws=1.0, tol=1e-7):
    G = make_G(N, mloc, seed, ws)
    A = np.eye(N) + G
    lam, U = np.linalg.eigh(G); lam = np.maximum(lam, 0)
    Mref = (U / np.sqrt(1 + lam)) @ U.T
    out = {}
    L = np.linalg.cholesky(A)
    Lp, perm = pivoted_chol(A)
    assert np.allclose(Lp @ Lp.T, A[np.ix_(perm, perm)])
    for name, W, pm in [("L", L, None), ("LT", L.T.copy(), None), ("A", A, None), ("Lpiv", Lp, perm), ("LpivT", Lp.T.copy(), perm)]:
        for fp in (True, False):
            Z, ns, hist = jacobi_oe(W, tol_stop=tol, float_params=fp)
            s2 = (Z * Z).sum(0)
            if name == "A":
                # A V = U Lambda: columns have norm lambda(A); u=z/|z|
                nz = np.sqrt(s2); M = (Z / (nz**2 * np.sqrt(nz))) @ Z.T
            elif name.endswith("T"):
                M = None  # would need V; only count sweeps
            else:
                M = (Z / (s2 * np.sqrt(np.maximum(s2, 1)))) @ Z.T
                if pm is not None:
                    Mi = np.empty_like(M); Mi[np.ix_(pm, pm)] = M; M = Mi
            err = np.nan if M is None else np.abs(M - Mref).max() / np.abs(Mref).max()
            out[(name, fp)] = (ns, err, ['%.0e' % h for h in hist])
    return out

if __name__ == "__main__" and len(sys.argv) > 1:
    for cfg in [(64, 200, 0, 1.0), (64, 200, 1, 10.0), (64, 30, 0, 1.0), (128, 1257, 0, 1.0)]:
        print(cfg)
        for k, v in variants(*cfg).items():
            print("  ", k, v)
