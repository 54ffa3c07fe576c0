#!/bin/bash
# Round-2 GPU call 21 (1 GPU): shift-and-add field inverse (Kaliski) on the device — full GPU suite (b200_fp_selftest compares it with
# the exponentiation), latency of the VM kernels / registry step at three batch sizes, bench.
set -u
O=gpurun_out/r2c21; mkdir -p $O
timeout 1800 python -m pytest tests -m gpu -x -q --durations=5 > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
for t in 4096 1024 256; do timeout 600 python tools/probe_vm_blobs.py tools/vm_blobs_default $t; done > $O/vm_latency.txt 2>&1
B200_BLS_TRACE=1 timeout 900 python bench.py --steps 5 --warmup 3 > $O/bench.json 2> $O/bench.err; echo "bench rc=$?" >> $O/bench.err
ls -la $O
