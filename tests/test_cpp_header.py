"""include/superintervals.hpp: compiles and links against the C ABI library (CPU), and runs the
reference's hot-path unit tests restated in tests/cpp/hotpath.cpp (GPU)."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "tests", "cpp", "hotpath")


def _compile():
    cmd = ["g++", "-std=c++17", "-O1", "-Wall", os.path.join(ROOT, "tests", "cpp", "hotpath.cpp"),
           "-I" + os.path.join(ROOT, "include"), "-L" + os.path.join(ROOT, "superintervals_b200"),
           "-lsuperintervals_b200", "-pthread", "-Wl,-rpath," + os.path.join(ROOT, "superintervals_b200"), "-o", EXE]
    out = subprocess.run(cmd, capture_output=True, text=True)
    assert out.returncode == 0, out.stderr


def test_cpp_header_compiles_and_links():
    _compile()
    assert os.path.exists(EXE)


@pytest.mark.gpu
def test_cpp_hot_path_unit_tests_pass_on_gpu():
    _compile()
    out = subprocess.run([EXE], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "All query tests passed" in out.stdout
    assert "All set operation tests passed" in out.stdout


REF_TESTS = os.path.join(ROOT, "oracle", "_ref", "run-tests-b200")


@pytest.mark.gpu
@pytest.mark.skipif(not os.path.exists(REF_TESTS), reason="oracle/_ref/run-tests-b200 not built (needs /root/reference at build time)")
def test_the_references_own_test_program_passes_against_this_library():
    """reference test/tests.cpp, UNMODIFIED, compiled by oracle/Makefile against include/superintervals.hpp and
    libsuperintervals_b200.so: its asserts (query families, edge cases, set operations) run on the GPU."""
    out = subprocess.run([REF_TESTS], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "All query tests passed" in out.stdout and "All set operation tests passed" in out.stdout
