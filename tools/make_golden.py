"""Writes tests/golden/rrsqrt_known_answers.npz: the reference's own known answers for its test inputs
(test/test_rrsqrt.F90: global analysis vs the Kalman gain :57-74, local analysis with Gaspari-Cohn weights vs the
explicit Pa formula :193-233 — closed forms, independent of LAPACK and of this repository's oracle), plus the
oracle's analysed anomalies Sa for the local cases, which no shipped assertion of the reference pins (regression
values for the CUDA path).      python tools/make_golden.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle  # noqa: E402
from refcases import kalman_check, rrsqrt_case  # noqa: E402

c = rrsqrt_case()
n, m = c["n"], c["m"]
xa_global, Pa_global = kalman_check(c["xf"], c["Sf"], c["H"], c["y"], np.diag(c["var"]))
# local analysis, one zone per state element, Gaspari-Cohn weights (test/test_rrsqrt.F90:193-233)
Pf = c["Sf"] @ c["Sf"].T
R = np.diag(c["var"])
xa_gc = np.zeros(n)
for i in range(n):
    w = np.array([oracle.locfun(abs(c["xmod"][i] - xo) / c["length"]) for xo in c["xobs"]])
    iloc = np.where(w != 0)[0]
    if len(iloc) == 0:
        xa_gc[i] = c["xf"][i]
        continue
    invR = np.linalg.inv(R[np.ix_(iloc, iloc)]) * np.outer(w[iloc], w[iloc])
    Hl = c["H"][iloc]
    Pa = np.linalg.inv(np.linalg.inv(Pf) + Hl.T @ invR @ Hl)
    xa_gc[i] = c["xf"][i] + Pa[i] @ (Hl.T @ (invR @ (c["y"][iloc] - Hl @ c["xf"])))
obs = oracle.make_obs(m, obsx=c["xobs"], obsy=np.zeros(m), weightfun=1)
xo, So, _, mloc = oracle.loc_analysis([1] * n, dict(x=c["xmod"], y=np.zeros(n)), c["length"], 1e30, obs, c["xf"],
                                      c["Hxf"], c["y"], c["Sf"], c["HSf"], c["var"])
assert np.abs(xo - xa_gc).max() < 1e-8
xg, Sg, _ = oracle.analysis(c["xf"], c["Hxf"], c["y"], c["Sf"], c["HSf"], c["var"])
assert np.abs(xg - xa_global).max() < 1e-8 and np.abs(Sg @ Sg.T - Pa_global).max() < 1e-8
out = os.path.join(ROOT, "tests", "golden", "rrsqrt_known_answers.npz")
np.savez(out, xa_global=xa_global, Pa_global=Pa_global, xa_gc_local=xa_gc, mloc_gc_local=mloc,
         Sa_gc_local_oracle=So, xa_gc_local_oracle=xo, Sa_global_oracle=Sg)
print("wrote", out, os.path.getsize(out), "bytes")
