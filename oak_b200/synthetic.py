"""Synthetic inputs of the benchmark configurations, reproducible anywhere.

Values come from a counter-based hash (splitmix64 on the global (row, member) index), so any rank can
generate any slab of the ensemble without communication, with numpy on the host or torch on the device.

Geometry ("C3-style"): nx x ny horizontal grid, nz levels, one variable, all sea, dx = dy = 1000 m,
x = i*dx, y = j*dy (i,j 1-based), zone = water column, zone label i + nx*(j-1), so the zone-permuted
state row of (i,j,k) is ((i-1) + nx*(j-1))*nz + (k-1) and the zone's first element is the surface level.
Observations: m points uniformly random in the horizontal domain at the surface, H = 4-point bilinear
rows, yo = H mu + 0.05 g, rmse = 0.05 (1 + 0.5 u).
"""
import math

import numpy as np

_M64 = (1 << 64) - 1


def _s64(v):
    v &= _M64
    return v - (1 << 64) if v >= (1 << 63) else v


_C1, _C2, _C3 = _s64(0x9E3779B97F4A7C15), _s64(0xBF58476D1CE4E5B9), _s64(0x94D049BB133111EB)


def _lsr(xp, z, s):
    return (z >> s) & ((1 << (64 - s)) - 1)


def _mix(xp, x):
    """splitmix64 finaliser on int64 arrays with wrap-around arithmetic (numpy or torch)."""
    z = x + _C1
    z = (z ^ _lsr(xp, z, 30)) * _C2
    z = (z ^ _lsr(xp, z, 27)) * _C3
    return z ^ _lsr(xp, z, 31)


def _uniform(xp, key):
    h = _mix(xp, key)
    u = _lsr(xp, h, 11)
    if xp is np:
        return u.astype(np.float64) * (1.0 / (1 << 53))
    return u.to(xp.float64) * (1.0 / (1 << 53))


def uniform(xp, idx, stream, seed=20261017):
    """u in [0,1) for int64 index array idx; `stream` separates arrays."""
    with np.errstate(over="ignore"):
        key = idx * 2 + _s64(seed * 0x1000003 + stream * 0x10001)
        return _uniform(xp, key * _s64(0xD1342543DE82EF95) + 1)


def normal(xp, idx, stream, seed=20261017):
    with np.errstate(over="ignore"):
        u1 = uniform(xp, idx * 2, stream, seed)
        u2 = uniform(xp, idx * 2 + 1, stream, seed)
    return xp.sqrt(-2.0 * xp.log(1.0 - u1)) * xp.cos((2.0 * math.pi) * u2)


class Grid:
    def __init__(self, nx, ny, nz, dx=1000.0):
        self.nx, self.ny, self.nz, self.dx = nx, ny, nz, dx
        self.nzones = nx * ny
        self.n = nx * ny * nz
        self.Lx, self.Ly = nx * dx, ny * dx

    def zone_xy(self, xp, zones):
        i = zones % self.nx
        j = zones // self.nx
        f = (lambda a: a.astype(np.float64)) if xp is np else (lambda a: a.to(xp.float64))
        return (f(i) + 1.0) * self.dx, (f(j) + 1.0) * self.dx

    def mu_rows(self, xp, rows):
        """background mean at zone-permuted state rows (int64 array)"""
        zones = rows // self.nz
        k = rows % self.nz
        x, y = self.zone_xy(xp, zones)
        kf = k.astype(np.float64) if xp is np else k.to(xp.float64)
        return xp.sin((2 * math.pi / self.Lx) * x) * xp.cos((2 * math.pi / self.Ly) * y) * xp.exp(-(kf + 1.0) / 10.0)


def ensemble_rows(xp, grid, rows, N, seed=20261017, members=None):
    """E[rows, members] = mu + 0.5 g, returned member-major: shape (len(members), len(rows))."""
    if members is not None:
        mem = members
    elif xp is np:
        mem = np.arange(N, dtype=np.int64)
    else:
        mem = xp.arange(N, dtype=rows.dtype, device=rows.device)
    mu = grid.mu_rows(xp, rows)
    idx = rows[None, :] * 4096 + (mem[:, None] if xp is np else mem[:, None].to(rows.dtype))
    return mu[None, :] + 0.5 * normal(xp, idx, 1, seed)


def observations(xp, grid, m, seed=20261017):
    """positions, bilinear operator (4 state rows + weights per obs), values, variances"""
    l = xp.arange(m, dtype=np.int64) if xp is np else xp.arange(m, dtype=xp.int64)
    ox = grid.dx * (1.0 + (grid.nx - 1) * uniform(xp, l, 2, seed))
    oy = grid.dx * (1.0 + (grid.ny - 1) * uniform(xp, l, 3, seed))
    fx = ox / grid.dx - 1.0
    fy = oy / grid.dx - 1.0
    i0 = xp.floor(fx)
    j0 = xp.floor(fy)
    if xp is np:
        i0 = np.minimum(i0, grid.nx - 2).astype(np.int64)
        j0 = np.minimum(j0, grid.ny - 2).astype(np.int64)
        ax = fx - i0
        ay = fy - j0
    else:
        i0 = xp.clamp(i0, max=grid.nx - 2).to(xp.int64)
        j0 = xp.clamp(j0, max=grid.ny - 2).to(xp.int64)
        ax = fx - i0.to(xp.float64)
        ay = fy - j0.to(xp.float64)
    rows = xp.stack([(i0 + grid.nx * j0) * grid.nz, (i0 + 1 + grid.nx * j0) * grid.nz,
                     (i0 + grid.nx * (j0 + 1)) * grid.nz, (i0 + 1 + grid.nx * (j0 + 1)) * grid.nz])
    wts = xp.stack([(1 - ax) * (1 - ay), ax * (1 - ay), (1 - ax) * ay, ax * ay])
    rmse = 0.05 * (1.0 + 0.5 * uniform(xp, l, 4, seed))
    noise = 0.05 * normal(xp, l, 5, seed)
    return dict(ox=ox, oy=oy, rows=rows, wts=wts, var=rmse * rmse, noise=noise)


def obs_space(xp, grid, obs, N, seed=20261017):
    """HE (member-major (N, m)), yo: applies the bilinear H to the synthetic ensemble and mean."""
    HE = None
    Hmu = None
    for c in range(4):
        r = obs["rows"][c]
        e = ensemble_rows(xp, grid, r, N, seed)
        HE = e * obs["wts"][c][None, :] if HE is None else HE + e * obs["wts"][c][None, :]
        mu = grid.mu_rows(xp, r) * obs["wts"][c]
        Hmu = mu if Hmu is None else Hmu + mu
    return HE, Hmu + obs["noise"]


def anomalies(xp, E):
    """mean and scaled anomalies of a member-major (N, rows) array (assimilation.F90:3127-3134)."""
    N = E.shape[0]
    mean = E.sum(0) / N if xp is np else E.sum(dim=0) / N
    return mean, (E - mean[None, :]) / math.sqrt(N - 1.0)


def coo_operator(grid, obs):
    """H as COO triplets (1-based) from the bilinear rows — for the assim_ensemble entry point."""
    m = obs["ox"].shape[0]
    Hi = np.tile(np.arange(1, m + 1, dtype=np.int32), 4)
    Hj = (np.asarray(obs["rows"]).reshape(-1) + 1).astype(np.int32)
    Hs = np.asarray(obs["wts"]).reshape(-1).astype(np.float64)
    return Hi, Hj, Hs


def small_case(nx=24, ny=20, nz=3, N=16, m=300, corr=3000.0, maxlen=6000.0, seed=7):
    """Everything on the host (numpy) for parity tests against the oracle."""
    g = Grid(nx, ny, nz)
    rows = np.arange(g.n, dtype=np.int64)
    E = ensemble_rows(np, g, rows, N, seed)          # (N, n)
    obs = observations(np, g, m, seed)
    HE, yo = obs_space(np, g, obs, N, seed)
    xf, Sf = anomalies(np, E)
    Hxf, HSf = anomalies(np, HE)
    zx, zy = g.zone_xy(np, np.arange(g.nzones, dtype=np.int64))
    return dict(grid=g, N=N, m=m, E=np.asfortranarray(E.T), xf=xf, Sf=np.asfortranarray(Sf.T), Hxf=Hxf,
                HSf=np.asfortranarray(HSf.T), yo=yo, var=obs["var"], obs=obs, zx=zx, zy=zy,
                zoneSize=np.full(g.nzones, nz, dtype=np.int32), corr=corr, maxlen=maxlen)
