#!/bin/bash
# Round-2 GPU call 25 (1 GPU): pairing-VM register files with padded 24-word slots (100 B per slot instead of 112) so that 14 instead of
# 10-12 warps of Miller teams fit an SM — with a 40-slot schedule and 64-thread CTAs the 8 192 teams of T = 4096 are ONE wave.
set -u
O=gpurun_out/r2c25; mkdir -p $O
{
for cta in 32 64; do
  echo "== shipping layout, vm_cta=$cta"; B200_VM_CTA=$cta timeout 600 python tools/probe_vm_blobs.py tools/vm_blobs_pad 4096
  echo "== padded layout, vm_cta=$cta"; B200_LIB=$PWD/ethereum_consensus_b200/libb200_consensus_pad4.so B200_VM_CTA=$cta timeout 600 python tools/probe_vm_blobs.py tools/vm_blobs_pad 4096
done
echo "== padded layout, vm_cta=64, T=2048 / 8192"
for t in 2048 8192; do B200_LIB=$PWD/ethereum_consensus_b200/libb200_consensus_pad4.so B200_VM_CTA=64 B200_VM_TEAM16_MAX=0 timeout 600 python tools/probe_vm_blobs.py tools/vm_blobs_pad $t; done
echo "== shipping, T=2048 / 8192 (8-lane teams forced)"
for t in 2048 8192; do B200_VM_TEAM16_MAX=0 timeout 600 python tools/probe_vm_blobs.py tools/vm_blobs_pad $t; done
} > $O/pad4.txt 2>&1
B200_LIB=$PWD/ethereum_consensus_b200/libb200_consensus_pad4.so timeout 900 python -m pytest tests/test_bls_gpu.py tests/test_vm_blob.py -x -q 2>&1 | tail -1 > $O/pytest_pad4.txt
ls -la $O
