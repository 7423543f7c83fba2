"""Runs a selection of the GPU parity tests against the CPU emulation build of the kernel sources.

tools/cuemu compiles oak_b200/csrc/*.cu a second time with g++ against a fiber-based emulation of the CUDA
execution model (blocks, warps, shuffles, barriers, mma.m8n8k4.f64 fragments) into a library with the same C
ABI.  It is a functional checker for the kernels in a container without a GPU — block decomposition, shared
memory indexing, fragment layouts — and nothing else: the product never loads it (oak_b200/_lib.py refuses it
unless OAK_B200_TEST_EMU=1), and the authoritative parity run is `pytest -m gpu` on the B200.
"""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools", "cuemu"))


@pytest.fixture(scope="module")
def emu_lib():
    import build_emu
    return build_emu.build()


def _run(emu_lib, expr, extra_env=None):
    env = dict(os.environ, OAK_B200_LIB=emu_lib, OAK_B200_TEST_EMU="1")
    env.update(extra_env or {})
    p = subprocess.run([sys.executable, "-m", "pytest", os.path.join(ROOT, "tests", "test_gpu_parity.py"),
                        os.path.join(ROOT, "tests", "test_variants_gpu.py"), "-q", "-x",
                        "-p", "no:cacheprovider", "-k", expr], env=env, capture_output=True, text=True, timeout=900)
    assert p.returncode == 0, p.stdout[-3000:] + p.stderr[-2000:]
    assert " passed" in p.stdout and "failed" not in p.stdout, p.stdout[-2000:]
    return p.stdout


def test_product_loader_refuses_the_emulated_library(emu_lib):
    code = "import oak_b200._lib as L; L.lib()"
    env = dict(os.environ, OAK_B200_LIB=emu_lib)
    env.pop("OAK_B200_TEST_EMU", None)
    p = subprocess.run([sys.executable, "-c", code], env=env, cwd=ROOT, capture_output=True, text=True)
    assert p.returncode != 0 and "no CPU fallback" in p.stderr


def test_reference_known_answers_through_every_transform_kernel(emu_lib):
    # test/test_rrsqrt.F90 and test/test_assim.F90 known answers, all five eig kernels (tridiagonal route, three
    # register-resident Jacobi variants, the simple cross-check kernel)
    _run(emu_lib, "rrsqrt or assim_case")


def test_kernels_under_a_permuted_thread_schedule(emu_lib):
    # the fibers of a block run in a shuffled order: a missing barrier shows up as a wrong result
    _run(emu_lib, "rrsqrt_gaspari_cohn and (eig_tridiag or eig_fast)", {"CUEMU_SHUFFLE": "7"})


def test_tensor_core_gram_variants(emu_lib):
    _run(emu_lib, "gram_tensor_core and 64-60")


def test_fused_apply_in_the_transform_kernel(emu_lib):
    # option fuse_apply (k_tvec updates the zone rows from the factored transform on mma tiles), incl. the zones it
    # leaves to k_apply (no observation / Jacobi fallback)
    _run(emu_lib, "fused_apply and (N20 or degenerate)")


def test_pushes_to_the_peers_in_pieces(emu_lib):
    # host logic of the fused gather (copy-engine flavour) with option push_pieces, alone and with fuse_apply
    _run(emu_lib, "pushes_in_pieces")


def test_eigenvector_kernel_split(emu_lib):
    # option tvec_split: vectors of T from a low-register kernel of their own, incl. its hand-over of flagged zones
    _run(emu_lib, "split_in_two and 24")


def test_global_scheme(emu_lib):
    # analysis (rrsqrt.F90:196-208): tall-skinny Gram in partial matrices, one transform, apply over row blocks
    _run(emu_lib, "global_scheme")   # also selects test_assim_case_through_the_global_scheme (BASELINE config 1)


def test_apply_on_tensor_core_tiles(emu_lib):
    # option apply_kernel = 1 (k_apply_mma): local zones of 1..75 rows and the row blocks of the global scheme
    _run(emu_lib, "apply_on_tensor")


def test_ensemble_prologue_and_epilogue_inside_the_apply_kernel(emu_lib):
    # option ens_fuse: bit-identical to the three-pass form (k_mean_anom, analysis, k_epilogue) and within 1e-9 of the oracle
    _run(emu_lib, "fused_in_the_apply")


def test_observation_operator_weights(emu_lib):
    # hgen.cu: batched cinterp against the oracle's restatement of ndgrid.F90 (cells, weights, tie-breaking on faces,
    # degenerate cells reported)
    env = dict(os.environ, OAK_B200_LIB=emu_lib, OAK_B200_TEST_EMU="1")
    p = subprocess.run([sys.executable, "-m", "pytest", os.path.join(ROOT, "tests", "test_hgen_gpu.py"), "-q", "-x",
                        "-p", "no:cacheprovider"], env=env, capture_output=True, text=True, timeout=900)
    assert p.returncode == 0 and " passed" in p.stdout, p.stdout[-3000:] + p.stderr[-2000:]
