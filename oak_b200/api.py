"""Host-side mirror of the reference's interface for the local-analysis path.

Reference signatures mirrored (all file:line under the OAK source tree):
    locAnalysis(zoneSize,selectObservations,xf,Hxf,yo,Sf,HSf,R,xa,Sa,amplitudes)    rrsqrt.F90:433-466
    selectObservations(ind,weight,relevantObs)                                      assimilation.F90:3683-3771
    DiagCovar / DCDCovar                                                            covariance.F90:70-79,:109-118
    ensemble branch of Assim                                                        assimilation.F90:3106-3134,:3235,:3301-3357
    parallPartion                                                                   parall.F90:166-186

Everything is executed by liboak_b200.so through the C ABI (include/oak_b200.h); errors raise
OakB200Error carrying the library's message (the reference prints and exits, ppdef.h:22).
"""
import ctypes as C

import numpy as np

from . import _lib

LOC_HORIZONTAL, LOC_DEPTH, LOC_TIME = 1, 2, 3
METRIC_CARTESIAN, METRIC_SPHERICAL, METRIC_SPHERICAL_APPROX = 0, 1, 2
WEIGHT_GAUSSIAN, WEIGHT_GASPARI_COHN, WEIGHT_UNIFORM = 0, 1, 2


class OakB200Error(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"oak_b200 error {code}: {msg}")
        self.code = code


def _check(rc):
    if rc != 0:
        raise OakB200Error(rc, _lib.lib().oakb200_last_error().decode(errors="replace"))


def _ptr(a):
    return None if a is None else C.c_void_p(a.ctypes.data)


class DiagCovar:
    """R = diag(D), D = error variances (covariance.F90:70-79; built as rmse**2 at assimilation.F90:2120-2127)."""

    def __init__(self, D):
        self.D = np.ascontiguousarray(D, dtype=np.float64)


class DCDCovar:
    """R^-1 x = D * (C^-1 (D * x)) (covariance.F90:109-118,:612-619); Assim wraps R with D in {0,1} for
    excluded observations (assimilation.F90:3086-3092)."""

    def __init__(self, D, Cov):
        if not isinstance(Cov, DiagCovar):
            raise OakB200Error(-6, "DCDCovar: only a DiagCovar inner covariance is supported (diagonal R)")
        self.D = np.ascontiguousarray(D, dtype=np.float64)
        self.C = Cov


def _flatten_R(R, m):
    """class(Covar) cannot cross the C ABI: extract (variance, d01) like the Fortran shim's `select type`."""
    if isinstance(R, DiagCovar):
        var, d01 = R.D, None
    elif isinstance(R, DCDCovar):
        var, d01 = R.C.D, R.D
    elif isinstance(R, np.ndarray) or np.isscalar(R):
        var, d01 = np.broadcast_to(np.asarray(R, dtype=np.float64), (m,)).copy(), None
    else:
        raise OakB200Error(-6, f"unsupported observation error covariance {type(R).__name__} (diagonal R only)")
    if var.shape != (m,):
        raise OakB200Error(-2, f"R has {var.shape} entries, expected ({m},)")
    return np.ascontiguousarray(var), None if d01 is None else np.ascontiguousarray(d01)


class Selector:
    """The state the reference's selectObservations callback reads from module globals
    (assimilation.F90:216-229,:3713-3767): per-zone position of the zone's first element,
    per-zone correlation / cut-off length, observation positions, loctype, metrictype."""

    def __init__(self, zone_x=None, zone_y=None, zone_z=None, zone_t=None, corrLen=1.0, maxLen=np.inf,
                 obs_x=None, obs_y=None, obs_z=None, obs_t=None, loctype=LOC_HORIZONTAL,
                 metrictype=METRIC_SPHERICAL, weightfun=WEIGHT_GAUSSIAN):
        self.zone = [None if v is None else np.ascontiguousarray(v, dtype=np.float64)
                     for v in (zone_x, zone_y, zone_z, zone_t)]
        self.obs = [None if v is None else np.ascontiguousarray(v, dtype=np.float64)
                    for v in (obs_x, obs_y, obs_z, obs_t)]
        self.corrLen, self.maxLen = corrLen, maxLen
        self.loctype, self.metrictype, self.weightfun = int(loctype), int(metrictype), int(weightfun)


class Handle:
    """One device context (one per GPU; replaces one MPI rank of parall.F90)."""

    def __init__(self, device=0, **options):
        self._L = _lib.lib()
        self._h = C.c_void_p()
        _check(self._L.oakb200_create(int(device), C.byref(self._h)))
        self.device = int(device)
        self.nzones = 0
        self.m = 0
        self._keep = []
        for k, v in options.items():
            self.set_option(k, v)

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self._L.oakb200_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def set_option(self, key, value):
        _check(self._L.oakb200_set_option(self._h, key.encode(), float(value)))

    # -- zones / observations ------------------------------------------------------------------
    def set_zones(self, zoneSize, zone_x=None, zone_y=None, zone_z=None, zone_t=None, corrLen=1.0,
                  maxLen=np.inf, loctype=LOC_HORIZONTAL, metrictype=METRIC_SPHERICAL,
                  weightfun=WEIGHT_GAUSSIAN):
        zs = np.ascontiguousarray(zoneSize, dtype=np.int32)
        nz = zs.size
        arrs = [None if v is None else np.ascontiguousarray(v, dtype=np.float64) for v in
                (zone_x, zone_y, zone_z, zone_t)]
        for a in arrs:
            if a is not None and a.shape != (nz,):
                raise OakB200Error(-2, f"zone coordinate array of shape {a.shape}, expected ({nz},)")
        cl = np.ascontiguousarray(np.broadcast_to(np.asarray(corrLen, dtype=np.float64), (nz,)))
        ml = np.ascontiguousarray(np.broadcast_to(np.asarray(maxLen, dtype=np.float64), (nz,)))
        _check(self._L.oakb200_set_zones(self._h, nz, _ptr(zs), _ptr(arrs[0]), _ptr(arrs[1]), _ptr(arrs[2]),
                                         _ptr(arrs[3]), _ptr(cl), _ptr(ml), int(loctype), int(metrictype),
                                         int(weightfun)))
        self.nzones = nz
        self.zoneSize = zs
        self.nrows = int(zs.astype(np.int64).sum())

    def set_observations(self, obs_x=None, obs_y=None, obs_z=None, obs_t=None):
        arrs = [None if v is None else np.ascontiguousarray(v, dtype=np.float64) for v in
                (obs_x, obs_y, obs_z, obs_t)]
        sizes = {a.size for a in arrs if a is not None}
        if len(sizes) > 1:
            raise OakB200Error(-2, "observation coordinate arrays differ in length")
        m = sizes.pop() if sizes else 0
        _check(self._L.oakb200_set_observations(self._h, m, _ptr(arrs[0]), _ptr(arrs[1]), _ptr(arrs[2]),
                                                _ptr(arrs[3])))
        self.m = m

    def configure(self, zoneSize, sel):
        """set_zones + set_observations from a Selector."""
        self.set_zones(zoneSize, *sel.zone, corrLen=sel.corrLen, maxLen=sel.maxLen, loctype=sel.loctype,
                       metrictype=sel.metrictype, weightfun=sel.weightfun)
        self.set_observations(*sel.obs)

    def select_observations(self, zone_first=0, zone_count=None):
        """selectObservations for a zone range: (offsets[zc+1], idx 1-based increasing, weight)."""
        zc = self.nzones - zone_first if zone_count is None else int(zone_count)
        offsets = np.zeros(zc + 1, dtype=np.int64)
        cap = max(1024, 64 * zc)
        while True:
            idx = np.zeros(cap, dtype=np.int32)
            w = np.zeros(cap, dtype=np.float64)
            rc = self._L.oakb200_select_observations(self._h, int(zone_first), zc, cap, _ptr(offsets), _ptr(idx),
                                                     _ptr(w))
            if rc == -5 and offsets[zc] > cap:
                cap = int(offsets[zc])
                continue
            _check(rc)
            tot = int(offsets[zc])
            return offsets, idx[:tot], w[:tot]

    # -- analysis --------------------------------------------------------------------------------
    def local_analysis(self, xf, Hxf, yo, Sf, HSf, R, out_Sa=None, want_amplitudes=False):
        """locAnalysis with host (numpy) arrays. Returns xa, Sa, amplitudes, stats."""
        Sf = np.asfortranarray(Sf, dtype=np.float64)
        HSf = np.asfortranarray(HSf, dtype=np.float64)
        n, N = Sf.shape
        m = HSf.shape[0]
        if HSf.shape[1] != N:
            raise OakB200Error(-2, f"HSf has {HSf.shape[1]} columns, Sf has {N}")
        xf = np.ascontiguousarray(xf, dtype=np.float64)
        Hxf = np.ascontiguousarray(Hxf, dtype=np.float64)
        yo = np.ascontiguousarray(yo, dtype=np.float64)
        var, d01 = _flatten_R(R, m)
        xa = np.empty(n)
        Sa = out_Sa if out_Sa is not None else np.empty((n, N), order="F")
        ampl = np.empty((N, self.nzones), order="F") if want_amplitudes else None
        st = _lib.Stats()
        _check(self._L.oakb200_local_analysis(self._h, n, N, m, _ptr(xf), _ptr(Hxf), _ptr(yo), _ptr(Sf), max(n, 1),
                                              _ptr(HSf), max(m, 1), _ptr(var), _ptr(d01), _ptr(xa), _ptr(Sa),
                                              max(n, 1), _ptr(ampl), C.byref(st)))
        return xa, Sa, ampl, st.asdict()

    def global_analysis(self, xf, Hxf, yo, Sf, HSf, R, out_Sa=None):
        """analysis (rrsqrt.F90:196-208): the global scheme with host (numpy) arrays; needs neither zones nor
        observation positions.  Returns xa, Sa, amplitudes (N), stats."""
        Sf = np.asfortranarray(Sf, dtype=np.float64)
        HSf = np.asfortranarray(HSf, dtype=np.float64)
        n, N = Sf.shape
        m = HSf.shape[0]
        if HSf.shape[1] != N:
            raise OakB200Error(-2, f"HSf has {HSf.shape[1]} columns, Sf has {N}")
        xf = np.ascontiguousarray(xf, dtype=np.float64)
        Hxf = np.ascontiguousarray(Hxf, dtype=np.float64)
        yo = np.ascontiguousarray(yo, dtype=np.float64)
        var, d01 = _flatten_R(R, m)
        xa = np.empty(n)
        Sa = out_Sa if out_Sa is not None else np.empty((n, N), order="F")
        ampl = np.empty(N)
        st = _lib.Stats()
        _check(self._L.oakb200_global_analysis(self._h, n, N, m, _ptr(xf), _ptr(Hxf), _ptr(yo), _ptr(Sf), max(n, 1),
                                               _ptr(HSf), max(m, 1), _ptr(var), _ptr(d01), _ptr(xa), _ptr(Sa),
                                               max(n, 1), _ptr(ampl), C.byref(st)))
        return xa, Sa, ampl, st.asdict()

    def global_analysis_dev(self, xf, Hxf, yo, Sf, HSf, Rdiag, xa, Sa, d01=None, ampl=None, stream=None):
        """The global scheme on CUDA tensors resident on this device (member-major like local_analysis_dev)."""
        import torch
        N, n = Sf.shape
        m = HSf.shape[1]
        for t in (xf, Hxf, yo, Sf, HSf, Rdiag, xa, Sa) + tuple(x for x in (d01, ampl) if x is not None):
            if not (t.is_cuda and t.dtype == torch.float64 and t.is_contiguous() and t.device.index == self.device):
                raise OakB200Error(-2, "global_analysis_dev needs contiguous fp64 CUDA tensors on the handle's device")
        if stream is None:
            stream = torch.cuda.current_stream(self.device).cuda_stream
        st = _lib.Stats()
        _check(self._L.oakb200_global_analysis_dev(
            self._h, n, N, m, C.c_void_p(xf.data_ptr()), C.c_void_p(Hxf.data_ptr()), C.c_void_p(yo.data_ptr()),
            C.c_void_p(Sf.data_ptr()), max(n, 1), C.c_void_p(HSf.data_ptr()), max(m, 1),
            C.c_void_p(Rdiag.data_ptr()), None if d01 is None else C.c_void_p(d01.data_ptr()),
            C.c_void_p(xa.data_ptr()), C.c_void_p(Sa.data_ptr()), max(n, 1),
            None if ampl is None else C.c_void_p(ampl.data_ptr()), C.c_void_p(stream), C.byref(st)))
        return st.asdict()

    def local_analysis_dev(self, xf, Hxf, yo, Sf, HSf, Rdiag, xa, Sa, d01=None, stream=None):
        """locAnalysis on CUDA tensors resident on this device (fp64).  Matrices are member-major:
        Sf/Sa of shape (N, n) and HSf of shape (N, m), contiguous — the column-major n x N / m x N
        arrays of the Fortran side.  Sa may be Sf.  Returns the stats dict."""
        import torch
        N, n = Sf.shape
        m = HSf.shape[1]
        for t in (xf, Hxf, yo, Sf, HSf, Rdiag, xa, Sa) + ((d01,) if d01 is not None else ()):
            if not (t.is_cuda and t.dtype == torch.float64 and t.is_contiguous() and t.device.index == self.device):
                raise OakB200Error(-2, "local_analysis_dev needs contiguous fp64 CUDA tensors on the handle's device")
        if stream is None:
            stream = torch.cuda.current_stream(self.device).cuda_stream
        st = _lib.Stats()
        _check(self._L.oakb200_local_analysis_dev(
            self._h, n, N, m, C.c_void_p(xf.data_ptr()), C.c_void_p(Hxf.data_ptr()), C.c_void_p(yo.data_ptr()),
            C.c_void_p(Sf.data_ptr()), max(n, 1), C.c_void_p(HSf.data_ptr()), max(m, 1),
            C.c_void_p(Rdiag.data_ptr()), None if d01 is None else C.c_void_p(d01.data_ptr()),
            C.c_void_p(xa.data_ptr()), C.c_void_p(Sa.data_ptr()), max(n, 1), None, C.c_void_p(stream),
            C.byref(st)))
        return st.asdict()

    def assim_ensemble_dev(self, E, Hi, Hj, Hs, Hshift, yo, Rdiag, Ea, anamtype=1, inflation=1.0, maxCorrection=None,
                           d01=None, xf_out=None, xa_out=None, stream=None):
        """Ensemble branch of Assim on CUDA tensors of this device: E, Ea member-major (N, n) (Ea may be E), Hi / Hj
        int32 1-based COO indices, Hs / Hshift / yo / Rdiag fp64.  Returns the stats dict."""
        import torch
        N, n = E.shape
        m = yo.numel()
        if stream is None:
            stream = torch.cuda.current_stream(self.device).cuda_stream
        p = lambda t: None if t is None else C.c_void_p(t.data_ptr())
        st = _lib.Stats()
        _check(self._L.oakb200_assim_ensemble_dev(
            self._h, n, N, m, p(E), max(n, 1), int(Hs.numel()), p(Hi), p(Hj), p(Hs), p(Hshift), p(yo), p(Rdiag), p(d01),
            int(anamtype), float(inflation), p(maxCorrection), p(Ea), max(n, 1), p(xf_out), p(xa_out),
            C.c_void_p(stream), C.byref(st)))
        return st.asdict()

    def set_peer_outputs(self, Sa_ptrs, xa_ptrs, ld, row0):
        """Fused all-gather (oakb200_set_peer_outputs): device pointers (ints) of the (N, n) member-major result
        array and of the mean vector of every destination; rows of this rank start at row0.  [] switches off."""
        k = len(Sa_ptrs)
        arr_S = (C.c_void_p * max(k, 1))(*[C.c_void_p(p) for p in Sa_ptrs])
        arr_x = (C.c_void_p * max(k, 1))(*[C.c_void_p(p) for p in xa_ptrs])
        _check(self._L.oakb200_set_peer_outputs(self._h, k, arr_S, arr_x, int(ld), int(row0)))

    def set_multicast_output(self, Sa_mc, xa_mc, ld, row0):
        """Fused gather through NVSwitch multicast (oakb200_set_multicast_output): multicast addresses (ints) of the
        (N, n) member-major result array and of the mean vector; rows of this rank start at row0.  None switches off."""
        _check(self._L.oakb200_set_multicast_output(self._h, C.c_void_p(Sa_mc) if Sa_mc else None,
                                                    C.c_void_p(xa_mc) if xa_mc else None, int(ld), int(row0)))

    def ipc_alloc(self, nbytes):
        """Device buffer other processes can map: returns (pointer, 64-byte handle)."""
        ptr = C.c_void_p()
        hd = C.create_string_buffer(64)
        _check(self._L.oakb200_ipc_alloc(self._h, int(nbytes), C.byref(ptr), hd))
        return ptr.value, hd.raw

    def ipc_open(self, handle):
        ptr = C.c_void_p()
        _check(self._L.oakb200_ipc_open(self._h, handle, C.byref(ptr)))
        return ptr.value

    def ipc_close(self, ptr):
        _check(self._L.oakb200_ipc_close(self._h, C.c_void_p(ptr)))

    def ipc_free(self, ptr):
        _check(self._L.oakb200_ipc_free(self._h, C.c_void_p(ptr)))

    def synchronize(self):
        """Completes an asynchronous local_analysis_dev (option async=1): status + stats."""
        st = _lib.Stats()
        _check(self._L.oakb200_synchronize(self._h, C.byref(st)))
        return st.asdict()

    def local_analysis_pinned(self, xf, Hxf, yo, Sf, HSf, Rdiag, xa, Sa, d01=None):
        """The host-buffer entry point on (pinned) CPU torch tensors, member-major like
        local_analysis_dev.  This is the end-to-end path: the library streams the state through the GPU."""
        N, n = Sf.shape
        m = HSf.shape[1]
        st = _lib.Stats()
        _check(self._L.oakb200_local_analysis(
            self._h, n, N, m, C.c_void_p(xf.data_ptr()), C.c_void_p(Hxf.data_ptr()), C.c_void_p(yo.data_ptr()),
            C.c_void_p(Sf.data_ptr()), max(n, 1), C.c_void_p(HSf.data_ptr()), max(m, 1),
            C.c_void_p(Rdiag.data_ptr()), None if d01 is None else C.c_void_p(d01.data_ptr()),
            C.c_void_p(xa.data_ptr()), C.c_void_p(Sa.data_ptr()), max(n, 1), None, C.byref(st)))
        return st.asdict()

    def set_anamorphosis_table(self, table):
        """AnamTrans%anam(v)%transform of the tabulated anamorphosis (type 3): K x 2, physical values and
        transformed values (assimilation.F90:4539-4567).  None clears it."""
        if table is None:
            _check(self._L.oakb200_set_anamorphosis_table(self._h, 0, None))
            return
        t = np.asfortranarray(table, dtype=np.float64)
        if t.ndim != 2 or t.shape[1] != 2:
            raise OakB200Error(-2, "anamorphosis table must be K x 2")
        _check(self._L.oakb200_set_anamorphosis_table(self._h, t.shape[0], _ptr(t)))

    def set_anamorphosis_vars(self, rowvar, specs):
        """Per-variable anamorphosis (assimilation.F90:4531-4567): rowvar = 1-based variable number of every row of the
        zone-permuted state, specs = [(type, table or None), ...] per variable.  Select with anamtype=0."""
        rv = np.ascontiguousarray(rowvar, dtype=np.int32)
        vt = np.array([t for t, _ in specs], dtype=np.int32)
        tabs = [np.asfortranarray(tb, dtype=np.float64) if tb is not None else np.zeros((0, 2), order="F") for _, tb in specs]
        vK = np.array([tb.shape[0] for tb in tabs], dtype=np.int32)
        flat = np.concatenate([tb.ravel(order="F") for tb in tabs] + [np.zeros(1)])
        _check(self._L.oakb200_set_anamorphosis_vars(self._h, len(specs), _ptr(vt), _ptr(vK), _ptr(flat), rv.size, _ptr(rv)))

    def assim_ensemble(self, E, Hi, Hj, Hs, Hshift, yo, R, anamtype=1, inflation=1.0, maxCorrection=None,
                       anamtable=None):
        """Ensemble branch of Assim (local scheme) with host arrays. Returns Ea, xf, xa, stats."""
        if anamtable is not None:
            self.set_anamorphosis_table(anamtable)
        E = np.asfortranarray(E, dtype=np.float64)
        n, N = E.shape
        yo = np.ascontiguousarray(yo, dtype=np.float64)
        m = yo.size
        Hi = np.ascontiguousarray(Hi, dtype=np.int32)
        Hj = np.ascontiguousarray(Hj, dtype=np.int32)
        Hs = np.ascontiguousarray(Hs, dtype=np.float64)
        Hshift = None if Hshift is None else np.ascontiguousarray(Hshift, dtype=np.float64)
        mc = None if maxCorrection is None else np.ascontiguousarray(maxCorrection, dtype=np.float64)
        var, d01 = _flatten_R(R, m)
        Ea = np.empty((n, N), order="F")
        xf = np.empty(n)
        xa = np.empty(n)
        st = _lib.Stats()
        _check(self._L.oakb200_assim_ensemble(self._h, n, N, m, _ptr(E), max(n, 1), Hs.size, _ptr(Hi), _ptr(Hj),
                                              _ptr(Hs), _ptr(Hshift), _ptr(yo), _ptr(var), _ptr(d01),
                                              int(anamtype), float(inflation), _ptr(mc), _ptr(Ea), max(n, 1),
                                              _ptr(xf), _ptr(xa), C.byref(st)))
        return Ea, xf, xa, st.asdict()

    def zone_counts(self):
        """Relevant observations per zone as the Gram kernel of the last analysis counted them."""
        out = np.zeros(self.nzones, dtype=np.int32)
        _check(self._L.oakb200_zone_counts(self._h, _ptr(out)))
        return out

    def cinterp(self, gshape, axes, xi, masked=None):
        """Batched cinterp (ndgrid.F90:1183-1257) on the device for one model grid with separable axes: returns
        (indexes[m][2^n][n] 1-based, coeff[m][2^n], nbp[m]).  Raises when observations fall into degenerate cells."""
        gs = np.ascontiguousarray(gshape, dtype=np.int32)
        n = gs.size
        ax = np.ascontiguousarray(np.concatenate([np.asarray(a, dtype=np.float64).ravel() for a in axes]))
        if ax.size != int(gs.sum()):
            raise OakB200Error(-2, f"axes hold {ax.size} values, gshape needs {int(gs.sum())}")
        mk = None if masked is None else np.ascontiguousarray(np.asarray(masked).ravel(order="F"), dtype=np.uint8)
        if mk is not None and mk.size != int(np.prod(gs)):
            raise OakB200Error(-2, "mask size does not match gshape")
        xi = np.ascontiguousarray(xi, dtype=np.float64).reshape(-1, n)
        m = xi.shape[0]
        idx = np.zeros((m, 1 << n, n), dtype=np.int32)
        co = np.zeros((m, 1 << n))
        nbp = np.zeros(m, dtype=np.int32)
        ndeg = C.c_int32(0)
        rc = self._L.oakb200_cinterp(self._h, n, _ptr(gs), _ptr(ax), _ptr(mk), m, _ptr(xi), _ptr(idx), _ptr(co),
                                     _ptr(nbp), C.byref(ndeg))
        self.last_cinterp = (idx, co, nbp, ndeg.value)
        _check(rc)
        return idx, co, nbp

    def cinterp_dev(self, gshape, axes, masked, xi, indexes, coeff, nbp):
        """Device-resident form (torch tensors): axes (float64, concatenated), masked (uint8 or None), xi (m x n),
        indexes (m x 2^n x n int32), coeff (m x 2^n), nbp (m int32).  Returns the number of degenerate observations."""
        gs = np.ascontiguousarray(gshape, dtype=np.int32)
        ndeg = C.c_int32(0)
        dp = lambda t: None if t is None else C.c_void_p(t.data_ptr())
        rc = self._L.oakb200_cinterp_dev(self._h, gs.size, _ptr(gs), dp(axes), dp(masked), int(xi.shape[0]), dp(xi),
                                         dp(indexes), dp(coeff), dp(nbp), C.byref(ndeg), None)
        _check(rc)
        return ndeg.value

    def fp64_peak(self, mode=0):
        v = C.c_double()
        _check(self._L.oakb200_fp64_peak(self._h, int(mode), C.byref(v)))
        return v.value


def partition_zones(nzones, nranks):
    """parallPartion with unit speeds (parall.F90:176-177): first[p]..first[p+1] (0-based) per rank."""
    first = np.zeros(nranks + 1, dtype=np.int32)
    _check(_lib.lib().oakb200_partition_zones(int(nzones), int(nranks), _ptr(first)))
    return first


def locanalysis(zoneSize, selectObservations, xf, Hxf, yo, Sf, HSf, R, handle=None, device=0,
                want_amplitudes=True, localise_obs=True):
    """locAnalysis (rrsqrt.F90:433-466): returns (xa, Sa, amplitudes).

    `selectObservations` is a Selector (the Fortran callback reads the same data from module
    globals); `R` a DiagCovar or DCDCovar.  amplitudes is zero, as in the reference's default
    local_obs branch (rrsqrt.F90:324); with localise_obs=False (rrsqrt.F90:374-385) every analysed zone uses all
    observations with their weights and amplitudes(:, zone) is filled."""
    own = handle is None
    h = Handle(device) if own else handle
    try:
        h.set_option("localise_obs", 1.0 if localise_obs else 0.0)
        h.configure(zoneSize, selectObservations)
        xa, Sa, ampl, _ = h.local_analysis(xf, Hxf, yo, Sf, HSf, R, want_amplitudes=want_amplitudes)
        return xa, Sa, ampl
    finally:
        if own:
            h.close()


def analysis(xf, Hxf, yo, Sf, HSf, R, handle=None, device=0):
    """analysis (rrsqrt.F90:196-208), the global scheme: returns (xa, Sa, amplitudes)."""
    own = handle is None
    h = Handle(device) if own else handle
    try:
        xa, Sa, ampl, _ = h.global_analysis(xf, Hxf, yo, Sf, HSf, R)
        return xa, Sa, ampl
    finally:
        if own:
            h.close()


def assim_ensemble(zoneSize, selectObservations, E, Hi, Hj, Hs, Hshift, yo, R, anamtype=1, inflation=1.0,
                   maxCorrection=None, handle=None, device=0, anamtable=None, anamvars=None):
    """Ensemble in, analysed ensemble out (assimilation.F90:3106-3134,:3235-3236,:3301-3357,:3558-3562).
    zoneSize = None selects the global scheme (schemetype = 0: `analysis` instead of `locanalysis`)."""
    own = handle is None
    h = Handle(device) if own else handle
    try:
        if zoneSize is None:          # global scheme (schemetype = 0): no zones, no observation positions
            h.set_option("scheme", 0)
        else:
            h.set_option("scheme", 1)
            h.configure(zoneSize, selectObservations)
        if anamvars is not None:      # per-variable transforms: (rowvar 1-based, [(type, table or None), ...])
            h.set_anamorphosis_vars(*anamvars)
            anamtype = 0
        Ea, xf, xa, _ = h.assim_ensemble(E, Hi, Hj, Hs, Hshift, yo, R, anamtype, inflation, maxCorrection,
                                         anamtable=anamtable)
        return Ea, xf, xa
    finally:
        if own:
            h.close()


def gen_observation_oper(handle, gshape, axes, xi, masked=None, seaindex=None):
    """The part of genObservationOper (assimilation.F90:2471-2656) that follows cinterp, for observations of ONE model
    variable: COO triplets (Hi, Hj, Hs), 1-based, 2^n entries per observation that lies in the grid and a single zero
    entry with model index -1 otherwise (:2597-2611).  Hj is the linear index of the corner in the variable's grid (first
    subscript fastest), mapped through `seaindex` (full -> packed, assimilation.F90:2379-2381) when given."""
    idx, co, nbp = handle.cinterp(gshape, axes, xi, masked)
    gs = np.asarray(gshape, dtype=np.int64)
    ioff = np.concatenate([[1], np.cumprod(gs[:-1])])
    m, twon, n = idx.shape
    lin = ((idx.astype(np.int64) - 1) * ioff).sum(axis=2) + 1          # [m][2^n], 1-based
    if seaindex is not None:
        lin = np.asarray(seaindex, dtype=np.int64)[lin - 1]
    inside = nbp > 0
    cnt = np.where(inside, twon, 1)
    start = np.concatenate([[0], np.cumsum(cnt)])
    nnz = int(start[-1])
    Hi = np.repeat(np.arange(1, m + 1, dtype=np.int32), cnt)
    Hj = np.full(nnz, -1, dtype=np.int32)
    Hs = np.zeros(nnz)
    pos = (start[:-1][inside][:, None] + np.arange(twon)[None, :]).ravel()
    Hj[pos] = lin[inside].ravel()
    Hs[pos] = co[inside].ravel()
    return Hi, Hj, Hs
