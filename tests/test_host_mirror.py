"""The C++ host-side mirror of PHDNavigator (monorfs_b200/host/phd_navigator.hpp): it must compile and
link against librbphd.so on CPU, fail loudly without a device, and pass the restated reference tests on GPU."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "cpp", "test_host_mirror.cpp")
LIBDIR = os.path.join(ROOT, "monorfs_b200", "_build")
EXE = os.path.join(ROOT, "tests", "cpp", "_build", "test_host_mirror")


def _compile():
    from monorfs_b200 import build
    build.build()
    os.makedirs(os.path.dirname(EXE), exist_ok=True)
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-o", EXE, SRC, "-L", LIBDIR, "-lrbphd",
                           "-Wl,-rpath," + LIBDIR])
    return EXE


def test_host_mirror_compiles_and_refuses_without_device():
    exe = _compile()
    import torch
    if torch.cuda.is_available():
        pytest.skip("a device is present")
    proc = subprocess.run([exe], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert proc.returncode != 0
    assert "no CPU fallback" in proc.stdout


@pytest.mark.gpu
def test_host_mirror_runs_reference_tests_on_gpu():
    exe = _compile()
    proc = subprocess.run([exe], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=300)
    assert proc.returncode == 0, proc.stdout
    assert "host mirror ok" in proc.stdout
