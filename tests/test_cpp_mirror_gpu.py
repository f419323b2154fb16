"""Builds tests/cpp/reference_suite.cpp (the reference's integration tests written against the C++
host mirror of the crate API) with g++, links it against the in-tree C-ABI library and runs it."""
import os
import subprocess

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_cpp_reference_suite(gpu_ctx, tmp_path):
    import spectrum_analyzer_b200 as sa
    lib_dir = os.path.dirname(sa.LIB_PATH)
    exe = str(tmp_path / "reference_suite")
    env = {k: v for k, v in os.environ.items() if k not in ("CC", "CXX")}
    subprocess.run(["g++", "-std=c++17", "-O1", "-o", exe, os.path.join(ROOT, "tests", "cpp", "reference_suite.cpp"),
                    "-L" + lib_dir, "-lspectrum_analyzer_cuda", "-Wl,-rpath," + lib_dir], check=True, env=env)
    out = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "reference_suite OK" in out.stdout
