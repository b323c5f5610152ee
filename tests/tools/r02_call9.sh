#!/bin/bash
set -u
out=gpurun_out/r02i
mkdir -p "$out"
nvidia-smi -L
timeout 600 python -m pytest tests/test_multi_gpu.py -m gpu -x -q 2>&1 | tail -15 > "$out/pytest_mgpu.txt"
tail -5 "$out/pytest_mgpu.txt"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > "$out/bench_n2.json" 2> "$out/bench_n2.err"
tail -3 "$out/bench_n2.err"; cut -c1-700 "$out/bench_n2.json"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --workload config5 --steps 10 --warmup 3 > "$out/bench5_n2.json" 2> "$out/bench5_n2.err"
tail -3 "$out/bench5_n2.err"; cut -c1-900 "$out/bench5_n2.json"
timeout 600 python bench.py --workload config5 --steps 10 --warmup 3 2>/dev/null | cut -c1-300
