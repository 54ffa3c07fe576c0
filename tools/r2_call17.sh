#!/bin/bash
# Round-2 GPU call 17 (1 GPU): pairing-VM translation unit at ptxas -O1 (shipping) / -O2 / -O3 after the carry-chain change.
set -u
O=gpurun_out/r2c17; mkdir -p $O
for lib in libb200_consensus.so libb200_consensus_vmo2.so libb200_consensus_vmo3.so; do
  for t in 4096 1024 256; do
    echo "== $lib T=$t"
    B200_LIB=$PWD/ethereum_consensus_b200/$lib timeout 600 python tools/probe_vm_blobs.py tools/vm_blobs_default $t
  done
done > $O/vm_ptxas_opt.txt 2>&1
timeout 600 python -m pytest tests/test_ssz_gpu.py -x -q > $O/pytest_ssz.log 2>&1
timeout 600 python bench.py --steps 3 --warmup 3 --skip-strong --skip-single --skip-rlc --skip-block > $O/bench.json 2> $O/bench.err
ls -la $O
