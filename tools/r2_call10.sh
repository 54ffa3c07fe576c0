#!/bin/bash
# Round-2 GPU call 10 (8 GPUs): the driver's scaling launch at N = 8 — NCCL communicator inside the library with 8 ranks,
# sharded hash_tree_root, weak-scaling headline and the strong-scaling epoch batch.
set -u
O=gpurun_out/r2c10; mkdir -p $O
B200_SSZ_TRACE=1 timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --steps 3 --warmup 3 --skip-rlc --skip-single > $O/bench_n8.json 2> $O/bench_n8.err; echo "rc=$?" >> $O/bench_n8.err
ls -la $O
