#!/bin/bash
set -u
mkdir -p gpurun_out/r02_fuzz
timeout 400 python tests/tools/fuzz_ref.py --gpu --minutes 4 --seed 20260924 > gpurun_out/r02_fuzz/fuzz_gpu.txt 2>&1; tail -5 gpurun_out/r02_fuzz/fuzz_gpu.txt
timeout 200 python tests/tools/fuzz_ref.py --gpu --minutes 2 --contacts --big-meshes --seed 77 > gpurun_out/r02_fuzz/fuzz_gpu_meshes.txt 2>&1; tail -3 gpurun_out/r02_fuzz/fuzz_gpu_meshes.txt
