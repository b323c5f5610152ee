"""Committed golden vectors (tests/golden/*.npz; tests/golden/make_golden.py writes them with the oracle and
checks them there against the reference build, so they are the reference's own outputs):
the oracle and the CPU emulation of the device code reproduce them byte for byte on CPU, the CUDA path
through the C ABI reproduces them on the GPU.  Geometry is rebuilt from the fixture alone."""
import os

import numpy as np
import pytest

from tests.common import P, make_scenes
from tests.golden.make_golden import REQUESTS, rebuild as _rebuild, run

HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
NAMES = ["primitives", "convex", "mesh", "fcl_meshes"]


def rebuild(sc, name, z):
    _rebuild(sc, name, {k: z[k] for k in z.files})


def same_records(a, b):
    # NaN payloads aside, every field must match bit for bit
    for f in a.dtype.names:
        x, y = a[f], b[f]
        if x.dtype.kind == "f":
            ok = (x.view(np.uint64) == y.view(np.uint64)) | (np.isnan(x) & np.isnan(y)) | ((x == 0) & (y == 0))
        else:
            ok = x == y
        if f == "_pad":
            continue
        assert np.all(ok), "field %s differs at rows %s" % (f, np.unique(np.nonzero(~ok)[0])[:8])


def check(name, backend_key, gpu):
    z = np.load(os.path.join(HERE, name + ".npz"))
    sc = make_scenes(gpu=gpu, emu=not gpu)
    rebuild(sc, name, z)
    for rname, (kind, kw) in REQUESTS.items():
        got = run(sc.b[backend_key], kind, kw, z["h1"], z["tf1"], z["h2"], z["tf2"])
        same_records(z["res_" + rname], got)


@pytest.mark.parametrize("name", NAMES)
def test_fixture_files_exist_and_are_nontrivial(name):
    z = np.load(os.path.join(HERE, name + ".npz"))
    d = z["res_distance_default"]
    assert len(d) >= 240 and (d["min_distance"] > 0).sum() > 50
    assert (z["res_collide_default"]["num_contacts"] == 1).sum() > 20


@pytest.mark.parametrize("name", NAMES)
def test_oracle_reproduces_golden(name):
    check(name, "oracle", gpu=False)


@pytest.mark.parametrize("name", NAMES)
def test_emulated_device_code_reproduces_golden(name):
    check(name, "emu", gpu=False)


@pytest.mark.gpu
@pytest.mark.parametrize("name", NAMES)
def test_gpu_reproduces_golden(name):
    check(name, "gpu", gpu=True)
