#!/bin/bash
set -u
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_cabi_and_host.py tests/test_broadphase.py -m gpu -x -q 2>&1 | tail -3
for p in 1 0 1; do HFB_HOST_PIPE=$p timeout 200 python tests/tools/bench_e2e.py 2>/dev/null | tail -1; done
