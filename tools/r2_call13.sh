#!/bin/bash
# Round-2 GPU call 13 (1 GPU): 16-lane VM teams + three 128-thread K1 CTAs per SM for small batches; full GPU suite; bench.
set -u
O=gpurun_out/r2c13; mkdir -p $O
{
for t in 256 1024 4096; do
  echo "== T=$t default"; B200_BLS_TRACE=1 timeout 600 python bench.py --steps 3 --warmup 3 --skip-ssz --skip-strong --skip-single --skip-rlc --tuples $t 2>&1 >/dev/null | grep "b200 bls" | sed -n "5,5p"
done
echo "== T=256 team 8 forced, 384-thread K1 CTAs forced"; B200_VM_TEAM16_MAX=0 B200_G1_SMALL_N=0 B200_BLS_TRACE=1 timeout 600 python bench.py --steps 3 --warmup 3 --skip-ssz --skip-strong --skip-single --skip-rlc --tuples 256 2>&1 >/dev/null | grep "b200 bls" | sed -n "5,5p"
echo "== T=1024 team 16 forced"; B200_VM_TEAM16_MAX=4096 B200_BLS_TRACE=1 timeout 600 python bench.py --steps 3 --warmup 3 --skip-ssz --skip-strong --skip-single --skip-rlc --tuples 1024 2>&1 >/dev/null | grep "b200 bls" | sed -n "5,5p"
} > $O/small_batches.txt 2>&1
timeout 1800 python -m pytest tests -m gpu -x -q --durations=6 > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
B200_BLS_TRACE=1 timeout 900 python bench.py --steps 5 --warmup 3 > $O/bench.json 2> $O/bench.err; echo "bench rc=$?" >> $O/bench.err
ls -la $O
