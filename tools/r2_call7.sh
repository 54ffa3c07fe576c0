#!/bin/bash
# Round-2 GPU call 7 (2 GPUs): RLC whole-batch check (1 GPU + sharded over 2), pipelined one-shot SSZ, new defaults
# (signature/message kernels under K1), full GPU suite, bench at N=1 and N=2 with SSZ trace.
set -u
O=gpurun_out/r2c7; mkdir -p $O
timeout 1800 python -m pytest tests -m gpu -x -q --durations=8 > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
B200_BLS_TRACE=1 timeout 900 python bench.py --steps 5 --warmup 3 > $O/bench.json 2> $O/bench.err; echo "bench rc=$?" >> $O/bench.err
B200_SSZ_TRACE=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 2 --steps 5 --warmup 3 --skip-rlc --skip-single > $O/bench_n2.json 2> $O/bench_n2.err; echo "rc=$?" >> $O/bench_n2.err
for c in 64 128; do echo "== CTA $c"; B200_SMALL_CTA=$c B200_BLS_TRACE=1 timeout 600 python bench.py --steps 3 --warmup 3 --skip-ssz --skip-strong --skip-single --skip-rlc 2>&1 >/dev/null | grep "b200 bls" | sed -n "5,5p"; done > $O/cta.txt 2>&1
ls -la $O
