#!/bin/bash
# Round-2 GPU call 23 (1 GPU): G2 kernels with inlined Fp2 products + 255 registers (libb200_consensus_g2inl255.so) against the
# shipping build in STRICT mode (they run under the per-key kernel there) and in registry mode at T = 1024 / 2048.
set -u
O=gpurun_out/r2c23; mkdir -p $O
for lib in libb200_consensus.so libb200_consensus_g2inl255.so libb200_consensus.so libb200_consensus_g2inl255.so; do
  echo "== $lib"
  B200_LIB=$PWD/ethereum_consensus_b200/$lib B200_PROBE_QUICK=1 timeout 600 python tools/probe_chunks.py 4096 512 2>/dev/null
  B200_LIB=$PWD/ethereum_consensus_b200/$lib B200_PROBE_QUICK=1 timeout 600 python tools/probe_chunks.py 1024 512 2>/dev/null
done > $O/strict_ab.txt 2>&1
ls -la $O
