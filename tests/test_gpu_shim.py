"""The C++ host-side mirror of the reference API (visfd_b200/csrc/visfd_cuda_shim.hpp),
driven like filter_mrc would drive it (tests/cpp/shim_check.cpp)."""
import os
import subprocess
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
def test_cpp_shim_against_oracle():
    exe = os.path.join(ROOT, "tests", "cpp", "shim_check")
    if not os.path.exists(exe):
        subprocess.check_call(["make", "-s", "-C", os.path.dirname(exe)])
    r = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    print(r.stdout, r.stderr)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "OK (0 failures)" in r.stdout


def test_cpp_shim_compiles():
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "tests", "cpp")])
