"""oak_b200 — B200-native local ensemble analysis for OAK (hot path only).

Host-side mirror of the reference interface for this path:

    locanalysis(zoneSize, selectObservations, xf, Hxf, yo, Sf, HSf, R, ...)      rrsqrt.F90:433
    assim_ensemble(...)   the ensemble branch of Assim around it                  assimilation.F90:3106-3357
    analysis(xf, Hxf, yo, Sf, HSf, R)   the global scheme on the same primitives                rrsqrt.F90:196
    Handle.cinterp / gen_observation_oper   interpolation weights of the observation operator   ndgrid.F90:1183, assimilation.F90:2471

implemented by the CUDA library oak_b200/liboak_b200.so through its C ABI (include/oak_b200.h).
There is no CPU fallback: importing works anywhere, calling needs the built library and a B200.
"""
from .api import (DiagCovar, DCDCovar, Handle, OakB200Error, Selector, analysis, assim_ensemble, gen_observation_oper,
                  locanalysis, partition_zones)

__all__ = ["Handle", "Selector", "DiagCovar", "DCDCovar", "locanalysis", "analysis", "assim_ensemble",
           "partition_zones", "gen_observation_oper", "OakB200Error"]
