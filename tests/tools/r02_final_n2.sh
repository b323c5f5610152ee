#!/bin/bash
# two GPUs of one box: the multi-GPU tests and the weak-scaling lines (C-ABI collectives)
set -u
out=gpurun_out/r02_final_n2
mkdir -p "$out"
timeout 900 python -m pytest tests/test_multi_gpu.py -m gpu -q 2>&1 | tail -3 > "$out/pytest.txt"; cat "$out/pytest.txt"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 > "$out/bench_n2.json" 2> "$out/bench_n2.err"; tail -c 500 "$out/bench_n2.json"; echo
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 2 --workload config5 > "$out/bench_config5_n2.json" 2> "$out/bench_config5_n2.err"; tail -c 500 "$out/bench_config5_n2.json"; echo
