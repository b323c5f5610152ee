#!/bin/bash
set -u
out=gpurun_out/r02_n8
mkdir -p "$out"
nvidia-smi -L | wc -l
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 8 --steps 10 --warmup 3 > "$out/bench_n8.json" 2> "$out/bench_n8.err"; tail -c 300 "$out/bench_n8.json"; echo; tail -3 "$out/bench_n8.err"
