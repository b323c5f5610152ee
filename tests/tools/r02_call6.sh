#!/bin/bash
set -u
out=gpurun_out/r02f
mkdir -p "$out"
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k object_table 2>&1 | tail -40 > "$out/pytest_obj.txt"
grep -E "Error|error|assert" "$out/pytest_obj.txt" | head -10
