"""The C++ host mirror (hpp-fcl_b200/host/hppfcl_b200.hpp): compiles and links against the C-ABI
library everywhere; on the GPU box the reference's own box_box_distance / capsule / collide cases
run through it (tests/cpp/test_host_api.cpp)."""
import os
import subprocess

import pytest

from tests.common import ROOT

SRC = os.path.join(ROOT, "tests", "cpp", "test_host_api.cpp")
EXE = os.path.join(ROOT, "tests", "cpp", "test_host_api.bin")


def _build():
    lib = os.path.join(ROOT, "hpp-fcl_b200")
    subprocess.check_call(["/usr/bin/g++", "-std=c++17", "-O1", "-Wall", "-o", EXE, SRC, "-L", lib,
                           "-lhppfcl_b200", "-Wl,-rpath," + lib])


def test_cpp_host_api_compiles_and_refuses_cpu(built):
    import torch
    _build()
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    r = subprocess.run([EXE], capture_output=True, text=True)
    assert r.returncode != 0 and "no CPU fallback" in r.stderr


@pytest.mark.gpu
def test_cpp_host_api_on_gpu(built):
    _build()
    r = subprocess.run([EXE], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "HOST-API-OK" in r.stdout, r.stdout + r.stderr
