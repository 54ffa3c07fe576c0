#!/bin/bash
# Round-2 GPU call 24 (1 GPU): G2 kernels with the curve routines inlined as well (libb200_consensus_g2tow.so) vs the shipping build.
set -u
O=gpurun_out/r2c24; mkdir -p $O
for rep in 1 2; do
for lib in libb200_consensus.so libb200_consensus_g2tow.so; do
  for t in 4096 1024 256; do
    echo "== $lib T=$t"
    B200_LIB=$PWD/ethereum_consensus_b200/$lib B200_BLS_TRACE=1 timeout 600 python tools/probe_vm_blobs.py tools/vm_blobs_default $t 2>/dev/null
  done
done; done > $O/g2_tower.txt 2>&1
B200_LIB=$PWD/ethereum_consensus_b200/libb200_consensus_g2tow.so timeout 900 python -m pytest tests/test_bls_gpu.py tests/test_rlc_gpu.py -x -q 2>&1 | tail -1 > $O/pytest_variant.txt
B200_LIB=$PWD/ethereum_consensus_b200/libb200_consensus_g2tow.so B200_PROBE_QUICK=1 timeout 600 python tools/probe_chunks.py 4096 512 2>/dev/null >> $O/g2_tower.txt
ls -la $O
