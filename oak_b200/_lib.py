"""ctypes binding of include/oak_b200.h.  Fails loudly when the library is missing."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# OAK_B200_LIB selects another build of the same library (A/B kernel experiments); default in-tree
LIB_PATH = os.environ.get("OAK_B200_LIB") or os.path.join(_HERE, "liboak_b200.so")
_LIB = None

c_dp = C.POINTER(C.c_double)
c_ip = C.POINTER(C.c_int32)
c_lp = C.POINTER(C.c_int64)


class Stats(C.Structure):
    _fields_ = [("zones_total", C.c_int64), ("zones_skipped", C.c_int64), ("obs_relevant_sum", C.c_int64),
                ("obs_candidate_sum", C.c_int64), ("jacobi_sweeps_sum", C.c_int64), ("h2d_bytes", C.c_int64),
                ("d2h_bytes", C.c_int64), ("ms_total", C.c_double), ("ms_pack", C.c_double),
                ("ms_gram", C.c_double), ("ms_eig", C.c_double), ("ms_apply", C.c_double),
                ("launches", C.c_int64), ("zones_fallback", C.c_int64),
                ("ms_tridiag", C.c_double), ("ms_tql", C.c_double), ("ms_tvec", C.c_double)]

    def asdict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


# name -> (restype, argtypes) ; must list every symbol include/oak_b200.h declares
SIGNATURES = {
    "oakb200_last_error": (C.c_char_p, []),
    "oakb200_version": (C.c_int, []),
    "oakb200_create": (C.c_int, [C.c_int, C.POINTER(C.c_void_p)]),
    "oakb200_destroy": (C.c_int, [C.c_void_p]),
    "oakb200_set_option": (C.c_int, [C.c_void_p, C.c_char_p, C.c_double]),
    "oakb200_set_zones": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                    C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32]),
    "oakb200_set_observations": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "oakb200_select_observations": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int64, C.c_void_p,
                                              C.c_void_p, C.c_void_p]),
    "oakb200_local_analysis": (C.c_int, [C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p,
                                         C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p,
                                         C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p,
                                         C.POINTER(Stats)]),
    "oakb200_local_analysis_dev": (C.c_int, [C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p,
                                             C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64,
                                             C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64,
                                             C.c_void_p, C.c_void_p, C.POINTER(Stats)]),
    "oakb200_global_analysis": (C.c_int, [C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p,
                                          C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p,
                                          C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p,
                                          C.POINTER(Stats)]),
    "oakb200_global_analysis_dev": (C.c_int, [C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p,
                                              C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64,
                                              C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64,
                                              C.c_void_p, C.c_void_p, C.POINTER(Stats)]),
    "oakb200_set_anamorphosis_table": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p]),
    "oakb200_host_alloc": (C.c_int, [C.c_int64, C.POINTER(C.c_void_p)]),
    "oakb200_host_free": (C.c_int, [C.c_void_p]),
    "oakb200_set_anamorphosis_vars": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64,
                                              C.c_void_p]),
    "oakb200_set_peer_outputs": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64]),
    "oakb200_set_multicast_output": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64]),
    "oakb200_ipc_alloc": (C.c_int, [C.c_void_p, C.c_int64, C.POINTER(C.c_void_p), C.c_char_p]),
    "oakb200_ipc_open": (C.c_int, [C.c_void_p, C.c_char_p, C.POINTER(C.c_void_p)]),
    "oakb200_ipc_close": (C.c_int, [C.c_void_p, C.c_void_p]),
    "oakb200_ipc_free": (C.c_int, [C.c_void_p, C.c_void_p]),
    "oakb200_assim_ensemble": (C.c_int, [C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_void_p, C.c_int64,
                                         C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                         C.c_void_p, C.c_void_p, C.c_int32, C.c_double, C.c_void_p, C.c_void_p,
                                         C.c_int64, C.c_void_p, C.c_void_p, C.POINTER(Stats)]),
    "oakb200_assim_ensemble_dev": (C.c_int, [C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_void_p, C.c_int64,
                                             C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                             C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_double,
                                             C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p,
                                             C.c_void_p, C.POINTER(Stats)]),
    "oakb200_synchronize": (C.c_int, [C.c_void_p, C.POINTER(Stats)]),
    "oakb200_partition_zones": (C.c_int, [C.c_int32, C.c_int32, C.c_void_p]),
    "oakb200_zone_counts": (C.c_int, [C.c_void_p, C.c_void_p]),
    "oakb200_cinterp": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p,
                                  C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "oakb200_cinterp_dev": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p,
                                      C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "oakb200_fp64_peak": (C.c_int, [C.c_void_p, C.c_int32, c_dp]),
}


def lib():
    """Loads liboak_b200.so; raises (no fallback) when it has not been built."""
    global _LIB
    if _LIB is not None:
        return _LIB
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python -m oak_b200.build` (needs nvcc). "
            "oak_b200 has no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    if hasattr(L, "oakb200_emulated") and os.environ.get("OAK_B200_TEST_EMU") != "1":
        raise RuntimeError(
            f"{LIB_PATH} is the CPU emulation build of the kernels (tools/cuemu, a test harness); it is never a "
            "product library: oak_b200 has no CPU fallback.")
    for name, (res, args) in SIGNATURES.items():
        f = getattr(L, name)  # AttributeError if the symbol is not exported
        f.restype = res
        f.argtypes = args
    _LIB = L
    return L
