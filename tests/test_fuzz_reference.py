"""Adversarial batches (tests/tools/fuzz_ref.py): axis-aligned and coincident poses, touching surfaces, extreme
scales, symmetric hulls, degenerate meshes and random request fields, compared record by record between the
reference build (oracle/_ref, when present), the oracle and the device code -- compiled for the host here, the
real kernels under `-m gpu`.  The rows on which the reference itself is undefined (Project::ProjectResult read
uninitialised for an exactly degenerate simplex, internal/intersect.h:58-70) are recognised and excluded.
"""
import os

import pytest

from oracle import oracle_lib
from tests.tools import fuzz_ref


def _ref():
    if os.path.isdir("/root/reference/src"):
        oracle_lib.build_ref()
    return oracle_lib.ref_available()


@pytest.mark.parametrize("first", [1, 41, 81])
def test_fuzz_oracle_reference_and_device_code(first):
    use_ref = _ref()
    for seed in range(first, first + 40):
        ok, tag = fuzz_ref.one_round(seed, 1500, use_ref, True)
        assert ok, tag


@pytest.mark.gpu
def test_fuzz_on_the_gpu():
    import hppfcl_b200 as hf
    use_ref = oracle_lib.ref_available()
    for seed in range(1, 25):
        ok, tag = fuzz_ref.one_round(seed, 4000, use_ref, hf.Engine(0))
        assert ok, tag
