"""CPU oracle for OAK's local ensemble analysis — TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this package.  The product (oak_b200/) never does; it fails loudly without its CUDA library.

`oracle.lib()` loads oracle/liboak_oracle.so (C restatement, see oak_oracle.c for the
reference file:line citations) and wires it to the scipy-bundled OpenBLAS for dsyev/dgemm.
"""
import ctypes as C
import glob
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

c_dp = C.POINTER(C.c_double)
c_ip = C.POINTER(C.c_int32)
c_u8p = C.POINTER(C.c_uint8)


class ObsT(C.Structure):
    _fields_ = [("m", C.c_int32), ("obsx", c_dp), ("obsy", c_dp), ("obsz", c_dp), ("obst", c_dp),
                ("loctype", C.c_int32), ("metrictype", C.c_int32), ("weightfun", C.c_int32),
                ("trig", C.c_int32)]


def build(force=False):
    so = os.path.join(_HERE, "liboak_oracle.so")
    src = os.path.join(_HERE, "oak_oracle.c")
    src2 = os.path.join(_HERE, "oak_ndgrid.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(src), os.path.getmtime(src2)):
        subprocess.check_call(["make", "-s", "-C", _HERE, "-B", "liboak_oracle.so"])
    return so


def _openblas_path():
    import scipy
    base = os.path.join(os.path.dirname(scipy.__file__), os.pardir, "scipy.libs")
    c = sorted(glob.glob(os.path.join(base, "libscipy_openblas*.so")))
    c = [p for p in c if "64_" not in os.path.basename(p)]
    if not c:
        raise RuntimeError("scipy-bundled OpenBLAS (LP64) not found")
    return os.path.abspath(c[0])


def lib():
    global _LIB
    if _LIB is not None:
        return _LIB
    so = build()
    L = C.CDLL(so)
    L.oracle_init_blas.argtypes = [C.c_char_p]
    rc = L.oracle_init_blas(_openblas_path().encode())
    if rc != 0:
        raise RuntimeError(f"oracle_init_blas failed: {rc}")
    L.oracle_locfun.restype = C.c_double
    L.oracle_locfun.argtypes = [C.c_double]
    L.oracle_distance.restype = C.c_double
    L.oracle_distance.argtypes = [C.c_int, C.c_int] + [C.c_double] * 4
    L.oracle_select_observations.restype = C.c_int
    L.oracle_select_observations.argtypes = [C.POINTER(ObsT)] + [C.c_double] * 6 + [c_dp, c_u8p]
    L.oracle_analysis.restype = C.c_int
    L.oracle_loc_analysis.restype = C.c_int
    L.oracle_assim_ensemble.restype = C.c_int
    L.oracle_analysis_increment.restype = C.c_int
    L.oracle_max_threads.restype = C.c_int
    L.oracle_set_threads.restype = C.c_int
    L.oracle_set_threads.argtypes = [C.c_int]
    if L.oracle_ndgrid_init(_openblas_path().encode()) != 0:
        raise RuntimeError("oracle_ndgrid_init failed (dgesvd not found)")
    L.oracle_ndgrid_create.restype = C.c_void_p
    L.oracle_ndgrid_create.argtypes = [C.c_int, c_ip, c_dp, c_u8p]
    L.oracle_ndgrid_destroy.argtypes = [C.c_void_p]
    L.oracle_cinterp.argtypes = [C.c_void_p, C.c_int, c_dp, c_ip, c_dp, c_ip]
    L.oracle_ndgrid_nsimplex.restype = C.c_int
    L.oracle_ndgrid_tetrahedra.argtypes = [C.c_int, c_dp]
    _LIB = L
    return L


def _dp(a):
    if a is None:
        return None
    assert a.dtype == np.float64 and (a.flags.f_contiguous or a.flags.c_contiguous)
    return a.ctypes.data_as(c_dp)


def _ip(a):
    if a is None:
        return None
    assert a.dtype == np.int32
    return a.ctypes.data_as(c_ip)


def _f(a):
    return None if a is None else np.asfortranarray(a, dtype=np.float64)


def make_obs(m, obsx=None, obsy=None, obsz=None, obst=None, loctype=1, metrictype=0, weightfun=0,
             trig=1):
    keep = [_f(v) for v in (obsx, obsy, obsz, obst)]
    o = ObsT(int(m), _dp(keep[0]), _dp(keep[1]), _dp(keep[2]), _dp(keep[3]), loctype, metrictype,
             weightfun, trig)
    o._keep = keep
    return o


def locfun(r):
    return lib().oracle_locfun(float(r))


def distance(metrictype, p0, p1, trig=0):
    return lib().oracle_distance(metrictype, trig, float(p0[0]), float(p0[1]), float(p1[0]),
                                 float(p1[1]))


def select_observations(obs, zone_pos, corrLen, maxLen):
    """one zone: returns (weight[m], relevant[m] bool) — assimilation.F90:3683-3771"""
    w = np.zeros(obs.m)
    r = np.zeros(obs.m, dtype=np.uint8)
    zp = list(zone_pos) + [0.0] * (4 - len(zone_pos))
    lib().oracle_select_observations(C.byref(obs), zp[0], zp[1], zp[2], zp[3], float(corrLen),
                                     float(maxLen), _dp(w), r.ctypes.data_as(c_u8p))
    return w, r.astype(bool)


def analysis(xf, Hxf, yo, Sf, HSf, var):
    """global analysis — rrsqrt.F90:196-208"""
    Sf = _f(Sf); HSf = _f(HSf)
    n, N = Sf.shape
    m = HSf.shape[0]
    xa = np.zeros(n); Sa = np.zeros((n, N), order="F"); ampl = np.zeros(N)
    rc = lib().oracle_analysis(m, n, N, _dp(_f(xf)), _dp(_f(Hxf)), _dp(_f(yo)), _dp(Sf), n, _dp(HSf),
                               max(m, 1), _dp(_f(var)), _dp(xa), _dp(Sa), n, _dp(ampl))
    if rc:
        raise RuntimeError(f"oracle_analysis status {rc}")
    return xa, Sa, ampl


def loc_analysis_cellgrid(zoneSize, zone_pos, corrLen, maxLen, obs, xf, Hxf, yo, Sf, HSf, var, e01=None):
    """locAnalysis with a CPU cell grid in front of the exact predicate (the "fair" CPU baseline;
    Cartesian metric, Gaussian weights with a finite cut-off only).  Same results as loc_analysis."""
    Sf = _f(Sf); HSf = _f(HSf)
    n, N = Sf.shape
    m = obs.m
    zs = np.ascontiguousarray(zoneSize, dtype=np.int32)
    nz = zs.size
    zx, zy = _f(zone_pos["x"]), _f(zone_pos["y"])
    cl = _f(np.broadcast_to(corrLen, (nz,)).copy())
    ml = _f(np.broadcast_to(maxLen, (nz,)).copy())
    xa = np.zeros(n); Sa = np.zeros((n, N), order="F")
    mloc = np.zeros(nz, dtype=np.int32)
    xf, Hxf, yo, var, e01 = [_f(a) for a in (xf, Hxf, yo, var, e01)]
    f = lib().oracle_loc_analysis_cellgrid
    f.restype = C.c_int
    rc = f(C.c_int(nz), _ip(zs), _dp(zx), _dp(zy), _dp(cl), _dp(ml), C.byref(obs), C.c_int(n), C.c_int(N), _dp(xf),
           _dp(Hxf), _dp(yo), _dp(Sf), C.c_int(n), _dp(HSf), C.c_int(max(m, 1)), _dp(var), _dp(e01), _dp(xa), _dp(Sa),
           C.c_int(n), _ip(mloc))
    if rc:
        raise RuntimeError(f"oracle_loc_analysis_cellgrid status {rc}")
    return xa, Sa, None, mloc


def loc_analysis(zoneSize, zone_pos, corrLen, maxLen, obs, xf, Hxf, yo, Sf, HSf, var, e01=None,
                 local_obs=True, zone_list=None, want_ampl=False):
    """locAnalysis — rrsqrt.F90:433-466.  zone_pos = dict(x=,y=,z=,t=) of per-zone arrays."""
    Sf = _f(Sf); HSf = _f(HSf)
    n, N = Sf.shape
    m = obs.m
    zs = np.ascontiguousarray(zoneSize, dtype=np.int32)
    nz = zs.size
    zp = [_f(zone_pos.get(k)) if zone_pos.get(k) is not None else None for k in "xyzt"]
    cl = _f(np.broadcast_to(corrLen, (nz,)).copy())
    ml = _f(np.broadcast_to(maxLen, (nz,)).copy())
    xa = np.zeros(n); Sa = np.zeros((n, N), order="F")
    ampl = np.zeros((N, nz), order="F") if want_ampl else None
    mloc = np.zeros(nz, dtype=np.int32)
    zl = None if zone_list is None else np.ascontiguousarray(zone_list, dtype=np.int32)
    args = (xf, Hxf, yo, var, e01)
    xf, Hxf, yo, var, e01 = [_f(a) for a in args]
    rc = lib().oracle_loc_analysis(
        C.c_int(nz), _ip(zs), _dp(zp[0]), _dp(zp[1]), _dp(zp[2]), _dp(zp[3]), _dp(cl), _dp(ml),
        C.byref(obs), C.c_int(1 if local_obs else 0), C.c_int(n), C.c_int(N), _dp(xf), _dp(Hxf),
        _dp(yo), _dp(Sf), C.c_int(n), _dp(HSf), C.c_int(max(m, 1)), _dp(var), _dp(e01), _dp(xa),
        _dp(Sa), C.c_int(n), _dp(ampl), _ip(zl), C.c_int(0 if zl is None else zl.size), _ip(mloc))
    if rc:
        raise RuntimeError(f"oracle_loc_analysis status {rc}")
    return xa, Sa, ampl, mloc


def interp1(x, y, xi):
    """anamorphosis.F90:304-339 — returns (yi, out)"""
    x = _f(x); y = _f(y)
    out = C.c_int(0)
    f = lib().oracle_interp1
    f.restype = C.c_double
    yi = f(C.c_int(x.size), _dp(x), _dp(y), C.c_double(xi), C.byref(out))
    return yi, bool(out.value)


def anamtransform(forward, anamtype, x, table=None):
    """assimilation.F90:4516-4576 on a vector (one variable): type 1 identity, 2 log/exp, 3 tabulated
    (table = K x 2: physical values, transformed values)"""
    f = lib().oracle_anam
    f.restype = C.c_double
    tab = _f(np.asfortranarray(table)) if table is not None else None
    K = 0 if tab is None else tab.shape[0]
    return np.array([f(C.c_int(anamtype), C.c_int(1 if forward else 0), C.c_int(K), _dp(tab), C.c_double(v))
                     for v in np.asarray(x, dtype=np.float64).ravel()])


def assim_ensemble(zoneSize, zone_pos, corrLen, maxLen, obs, E, Hi, Hj, Hs, Hshift, yo, var,
                   e01=None, anamtype=1, inflation=1.0, maxCorrection=None, anamtable=None, anamvars=None):
    """ensemble branch of Assim with the local scheme — assimilation.F90:3106-3134, :3235, :3301-3357"""
    tab = _f(np.asfortranarray(anamtable)) if anamtable is not None else None
    lib().oracle_set_anam_table(C.c_int(0 if tab is None else tab.shape[0]), _dp(tab))
    if anamvars is not None:   # per-variable transforms: (rowvar 0-based per row, [(type, table or None), ...]); anamtype 0
        rowvar, specs = anamvars
        rv = np.ascontiguousarray(rowvar, dtype=np.int32)
        vt = np.array([t for t, _ in specs], dtype=np.int32)
        tabs = [np.asfortranarray(tb, dtype=np.float64) if tb is not None else np.zeros((0, 2), order="F") for _, tb in specs]
        vK = np.array([tb.shape[0] for tb in tabs], dtype=np.int32)
        voff = np.concatenate([[0], np.cumsum(2 * vK)[:-1]]).astype(np.int32)
        flat = np.concatenate([tb.ravel(order="F") for tb in tabs] + [np.zeros(1)])
        _keep = (rv, vt, voff, vK, flat)
        lib().oracle_set_anam_vars(_ip(rv), _ip(vt), _ip(voff), _ip(vK), _dp(flat))
        anamtype = 0
    E = _f(E)
    n, N = E.shape
    zs = np.ascontiguousarray(zoneSize, dtype=np.int32)
    nz = zs.size
    zp = [_f(zone_pos.get(k)) if zone_pos.get(k) is not None else None for k in "xyzt"]
    cl = _f(np.broadcast_to(corrLen, (nz,)).copy())
    ml = _f(np.broadcast_to(maxLen, (nz,)).copy())
    Hi = np.ascontiguousarray(Hi, dtype=np.int32); Hj = np.ascontiguousarray(Hj, dtype=np.int32)
    Hs = _f(Hs); Hshift = _f(Hshift); yo = _f(yo); var = _f(var); e01 = _f(e01)
    maxCorrection = _f(maxCorrection)
    Ea = np.zeros((n, N), order="F"); xf = np.zeros(n); xa = np.zeros(n)
    rc = lib().oracle_assim_ensemble(
        C.c_int(nz), _ip(zs), _dp(zp[0]), _dp(zp[1]), _dp(zp[2]), _dp(zp[3]), _dp(cl), _dp(ml),
        C.byref(obs), C.c_int(n), C.c_int(N), _dp(E), C.c_int(n), C.c_int64(Hs.size), _ip(Hi),
        _ip(Hj), _dp(Hs), _dp(Hshift), _dp(yo), _dp(var), _dp(e01), C.c_int(anamtype),
        C.c_double(inflation), _dp(maxCorrection), _dp(Ea), C.c_int(n), _dp(xf), _dp(xa))
    if rc:
        raise RuntimeError(f"oracle_assim_ensemble status {rc}")
    return Ea, xf, xa


def rotate_vector(w, v):
    N = len(w)
    Om = np.zeros((N, N), order="F")
    lib().oracle_rotate_vector(C.c_int(N), _dp(_f(w)), _dp(_f(v)), _dp(Om))
    return Om


def init_partition(partition, nzones):
    p = np.ascontiguousarray(partition, dtype=np.int32)
    n = p.size
    zs = np.zeros(nzones, dtype=np.int32); zi = np.zeros(n, dtype=np.int32); izi = np.zeros(n, dtype=np.int32)
    lib().oracle_init_partition(C.c_int(n), _ip(p), C.c_int(nzones), _ip(zs), _ip(zi), _ip(izi))
    return zs, zi, izi


def max_threads():
    return lib().oracle_max_threads()


def set_threads(n=0):
    """OpenMP threads of the zone loop (n <= 0: all online processors, whatever OMP_NUM_THREADS says)."""
    return lib().oracle_set_threads(int(n))


# ---------------------------------------------------------------------------------------------------
# n-dimensional grid interpolation (ndgrid.F90: cinterp), the arithmetic of the observation operator
# ---------------------------------------------------------------------------------------------------
def ndgrid_full_coords(gshape, axes=None, coords=None):
    """Coordinates of every grid point, [n][prod(gshape)] with the first dimension fastest (Fortran order):
    from separable axes (one 1-D array per dimension) or from explicit arrays of the grid's shape."""
    gshape = tuple(int(g) for g in gshape)
    n = len(gshape)
    total = int(np.prod(gshape))
    out = np.empty((n, total))
    for d in range(n):
        if coords is not None:
            out[d] = np.asarray(coords[d], dtype=np.float64).reshape(gshape, order="F").ravel(order="F")
        else:
            shp = [1] * n
            shp[d] = gshape[d]
            out[d] = np.broadcast_to(np.asarray(axes[d], dtype=np.float64).reshape(shp), gshape).ravel(order="F")
    return out


def tetrahedra(n):
    """tetrahedron(2^n, n+1, nbth) of split (ndgrid.F90:357-435), as [nbth][n+1][2^n]"""
    L = lib()
    nb = L.oracle_ndgrid_nsimplex(n)
    t = np.zeros((nb, n + 1, 1 << n))
    L.oracle_ndgrid_tetrahedra(n, _dp(t))
    return t


def cinterp(gshape, coord_full, xi, masked=None):
    """cinterp (ndgrid.F90:1183-1257) for the points xi[m][n]: (indexes[m][2^n][n] 1-based, coeff[m][2^n], nbp[m])"""
    L = lib()
    gs = np.ascontiguousarray(gshape, dtype=np.int32)
    n = gs.size
    cf = np.ascontiguousarray(coord_full, dtype=np.float64)
    mk = None if masked is None else np.ascontiguousarray(masked, dtype=np.uint8)
    xi = np.ascontiguousarray(xi, dtype=np.float64).reshape(-1, n)
    m = xi.shape[0]
    idx = np.zeros((m, 1 << n, n), dtype=np.int32)
    co = np.zeros((m, 1 << n))
    nbp = np.zeros(m, dtype=np.int32)
    g = L.oracle_ndgrid_create(n, gs.ctypes.data_as(c_ip), _dp(cf), None if mk is None else mk.ctypes.data_as(c_u8p))
    try:
        L.oracle_cinterp(g, m, _dp(xi), idx.ctypes.data_as(c_ip), _dp(co), nbp.ctypes.data_as(c_ip))
    finally:
        L.oracle_ndgrid_destroy(g)
    return idx, co, nbp


class NdGrid:
    """Keeps the oracle's grid (and its lazily built databox tree, as the reference keeps it in type(grid)) alive
    across calls: the CPU baseline of bench.py --config hgen times cinterp with the tree already built."""

    def __init__(self, gshape, coord_full, masked=None):
        self.L = lib()
        self.gs = np.ascontiguousarray(gshape, dtype=np.int32)
        self.n = self.gs.size
        self.cf = np.ascontiguousarray(coord_full, dtype=np.float64)
        self.mk = None if masked is None else np.ascontiguousarray(masked, dtype=np.uint8)
        self.g = self.L.oracle_ndgrid_create(self.n, self.gs.ctypes.data_as(c_ip), _dp(self.cf),
                                             None if self.mk is None else self.mk.ctypes.data_as(c_u8p))

    def cinterp(self, xi):
        xi = np.ascontiguousarray(xi, dtype=np.float64).reshape(-1, self.n)
        m = xi.shape[0]
        idx = np.zeros((m, 1 << self.n, self.n), dtype=np.int32)
        co = np.zeros((m, 1 << self.n))
        nbp = np.zeros(m, dtype=np.int32)
        self.L.oracle_cinterp(self.g, m, _dp(xi), idx.ctypes.data_as(c_ip), _dp(co), nbp.ctypes.data_as(c_ip))
        return idx, co, nbp

    def close(self):
        if self.g:
            self.L.oracle_ndgrid_destroy(self.g)
            self.g = None
