"""Builds the compile-time variants tools/r2_ab.sh times (no GPU needed: nvcc cross-compiles) into
oak_b200/variants/; the .so files are git-ignored but travel to the GPU box with the gpurun snapshot."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oak_b200.build import build_variant  # noqa: E402

out = os.path.join(ROOT, "oak_b200", "variants")
os.makedirs(out, exist_ok=True)
for name, defs in (("liboak_tw2.so", ["TVEC_TWISTED2=1"]), ("liboak_tqll.so", ["TQL_LOCAL=1"]),
                   ("liboak_tw2_tqll.so", ["TVEC_TWISTED2=1", "TQL_LOCAL=1"])):
    print(build_variant(os.path.join(out, name), defs))
