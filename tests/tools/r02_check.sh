#!/bin/bash
# quick GPU check of the host mirrors and the host entry points after a change on the host side of the library
set -u
timeout 900 python -m pytest tests/test_bvh_parity.py tests/test_cpp_host_api.py tests/test_plane_halfspace.py tests/test_gpu_parity.py tests/test_broadphase.py -m gpu -x -q 2>&1 | tail -4
timeout 200 python tests/tools/bench_e2e.py 2>/dev/null | tail -1
