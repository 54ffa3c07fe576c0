#!/bin/bash
# Round-2 GPU call 16 (1 GPU): hardware carry-chain add / subtract (fp.cuh) in every kernel; A/B of call-based products in the
# pairing VM (libb200_consensus_vmcall.so); full GPU suite; bench (block-signature-set leg, SSZ side-stream stage chains).
set -u
O=gpurun_out/r2c16; mkdir -p $O
timeout 1800 python -m pytest tests -m gpu -x -q --durations=5 > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
for lib in libb200_consensus.so libb200_consensus_vmcall.so; do
  for t in 4096 1024; do
    echo "== $lib T=$t"
    B200_LIB=$PWD/ethereum_consensus_b200/$lib B200_PROBE_QUICK=1 B200_BLS_TRACE=1 timeout 600 python tools/probe_chunks.py $t 512 2> $O/trace_${lib}_$t.err
    grep "b200 bls" $O/trace_${lib}_$t.err | sed -n "5p;\$p"
  done
done > $O/ab_vmcall.txt 2>&1
B200_LIB=$PWD/ethereum_consensus_b200/libb200_consensus_vmcall.so timeout 900 python -m pytest tests/test_bls_gpu.py tests/test_rlc_gpu.py -x -q > $O/pytest_vmcall.log 2>&1
for t in 4096 1024 256; do timeout 600 python tools/probe_vm_blobs.py tools/vm_blobs $t; done > $O/vm_schedules.txt 2>&1
for t in 4096 1024; do B200_LIB=$PWD/ethereum_consensus_b200/libb200_consensus_vmcall.so timeout 600 python tools/probe_vm_blobs.py tools/vm_blobs $t; done > $O/vm_schedules_vmcall.txt 2>&1
B200_BLS_TRACE=1 timeout 900 python bench.py --steps 5 --warmup 3 > $O/bench.json 2> $O/bench.err; echo "bench rc=$?" >> $O/bench.err
ls -la $O
