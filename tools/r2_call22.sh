#!/bin/bash
# Round-2 GPU call 22 (1 GPU): the latency-bound one-thread G2 kernels (signature decode, hash_to_G2) built three other ways —
# Fp2 products inlined into the curve routines, 255 instead of 128 registers, both — against the shipping build.
set -u
O=gpurun_out/r2c22; mkdir -p $O
for lib in libb200_consensus.so libb200_consensus_g2inl.so libb200_consensus_g2r255.so libb200_consensus_g2inl255.so; do
  for t in 4096 256; do
    echo "== $lib T=$t"
    B200_LIB=$PWD/ethereum_consensus_b200/$lib B200_BLS_TRACE=1 timeout 600 python tools/probe_vm_blobs.py tools/vm_blobs_default $t 2> $O/err_${lib}_$t.txt
  done
done > $O/g2_variants.txt 2>&1
for lib in libb200_consensus_g2inl.so libb200_consensus_g2inl255.so; do
  B200_LIB=$PWD/ethereum_consensus_b200/$lib timeout 900 python -m pytest tests/test_bls_gpu.py -x -q 2>&1 | tail -1
done > $O/pytest_variants.txt 2>&1
ls -la $O
