#!/bin/bash
# Round-2 GPU call 19 (2 GPUs): final multi-rank validation — the two-process NCCL test through the C ABI (no torch) and bench.py
# under torchrun at N = 2 exactly as the driver launches it.
set -u
O=gpurun_out/r2c19; mkdir -p $O
timeout 900 python -m pytest tests/test_config_scale_gpu.py tests/test_rlc_gpu.py -m gpu -x -q --durations=5 > $O/pytest_2gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_2gpu.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 5 --warmup 3 > $O/bench_n2.json 2> $O/bench_n2.err; echo "rc=$?" >> $O/bench_n2.err
ls -la $O
