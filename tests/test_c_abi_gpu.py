"""The C ABI without ctypes: a plain C program (tests/c_abi/test_rrsqrt_abi.c) is compiled against
include/oak_b200.h, linked with -loak_b200 and run on the GPU box; it checks the reference's test_rrsqrt
known answers (test/test_rrsqrt.F90:57-74,:142-159) computed inside the C program itself."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "c_abi", "test_rrsqrt_abi.c")


def _build(tmp_path):
    exe = str(tmp_path / "test_rrsqrt_abi")
    libdir = os.path.join(ROOT, "oak_b200")
    subprocess.check_call(["gcc", "-O1", "-Wall", "-I", os.path.join(ROOT, "include"), SRC, "-L", libdir, "-loak_b200",
                           "-Wl,-rpath," + libdir, "-lm", "-o", exe])
    return exe


def test_c_program_compiles_and_links_against_the_header(tmp_path):
    """CPU: the header is valid C and every symbol the program uses resolves at link time."""
    from oak_b200 import build
    build.build()
    assert os.path.exists(_build(tmp_path))


@pytest.mark.gpu
def test_c_program_runs_the_rrsqrt_known_answers(tmp_path):
    p = subprocess.run([_build(tmp_path)], capture_output=True, text=True, timeout=300)
    assert p.returncode == 0 and "OK" in p.stdout, p.stdout + p.stderr
