"""Closed-form inputs of the reference's own known-answer tests, regenerated (no files needed).

test/test_rrsqrt.F90:283-322   m=5, n=10, N=12, Ef=sin(3 i^2), H=reshape(1..50), y=1..5, R=2I
test/test_assim.F90:59-125     3x3 grid x 2 variables (n=18), N=10, Ef=sin(3 i), 1 obs at (2,2)=1, R=2
test/test_cellgrid.F90:6-8     integer lattices 4^2 / 20^2 / 1000^2, query (2,2), maxdist 2/3/5
test/test_covariance.F90:612-617 locfun known answers
"""
import numpy as np


def rrsqrt_case():
    m, n, N = 5, 10, 12
    i = np.arange(1, n * N + 1, dtype=np.float64)
    Ef = np.sin(3.0 * i * i).reshape((n, N), order="F")
    H = np.arange(1, m * n + 1, dtype=np.float64).reshape((m, n), order="F")
    y = np.arange(1, m + 1, dtype=np.float64)
    xf = Ef.sum(axis=1) / N
    Sf = (Ef - xf[:, None]) / np.sqrt(N - 1.0)
    var = np.full(m, 2.0)
    return dict(m=m, n=n, N=N, Ef=Ef, H=H, y=y, xf=xf, Sf=np.asfortranarray(Sf), var=var,
                HSf=np.asfortranarray(H @ Sf), Hxf=H @ xf,
                xmod=np.arange(n) / (n - 1.0), xobs=np.arange(m) / (m - 1.0), length=0.21)


def kalman_check(xf, Sf, H, y, R):
    Pf = Sf @ Sf.T
    K = Pf @ H.T @ np.linalg.inv(H @ Pf @ H.T + R)
    return xf + K @ (y - H @ xf), Pf - K @ H @ Pf


def assim_case():
    imax, jmax, nvar, N, m = 3, 3, 2, 10, 1
    n = imax * jmax * nvar
    i = np.arange(1, n * N + 1, dtype=np.float64)
    Ef = np.sin(3.0 * i).reshape((n, N), order="F")
    H = np.zeros((m, n))
    H[0, (2 - 1) + imax * (2 - 1)] = 1.0  # sub2ind([imax,jmax],[2,2])
    x = np.tile(np.arange(1, imax + 1, dtype=np.float64), jmax * nvar)
    yy = np.tile(np.repeat(np.arange(1, jmax + 1, dtype=np.float64), imax), nvar)
    return dict(n=n, N=N, m=m, Ef=np.asfortranarray(Ef), H=H, yo=np.array([1.0]), var=np.array([2.0]),
                x=x, y=yy, obsx=np.array([2.0]), obsy=np.array([2.0]))


LOCFUN_GOLDEN = [(0.0, 1.0), (0.4, 0.783573333333333), (1.5, 0.0164930555555556), (2.5, 0.0)]
CELLGRID_CASES = [(4, 2.0), (20, 3.0), (1000, 5.0)]
