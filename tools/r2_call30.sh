#!/bin/bash
# Round-2 GPU call 30 (4 GPUs): bench.py under torchrun at N = 4 exactly as the driver launches it (final build).
set -u
O=gpurun_out/r2c30; mkdir -p $O
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 4 --steps 5 --warmup 3 > $O/bench_n4.json 2> $O/bench_n4.err; echo "rc=$?" >> $O/bench_n4.err
ls -la $O; tail -2 $O/bench_n4.err
