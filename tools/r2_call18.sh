#!/bin/bash
# Round-2 GPU call 18 (1 GPU): full GPU suite (mixed registry mode, split key copy on by default); split-copy and per-key-kernel
# variant A/B; bench; ncu of the per-key kernel after the carry-chain change.
set -u
O=gpurun_out/r2c18; mkdir -p $O
timeout 1800 python -m pytest tests -m gpu -x -q --durations=5 > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
{ timeout 600 python tools/probe_split.py 4096 512; B200_G1_VARIANT=0 timeout 600 python tools/probe_split.py 4096 512; } > $O/split_ab.txt 2>&1
B200_BLS_TRACE=1 timeout 900 python bench.py --steps 10 --warmup 5 > $O/bench.json 2> $O/bench.err; echo "bench rc=$?" >> $O/bench.err
B200_BLS_KEY_SPLIT=0 timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_g1_validate -s 2 -c 1 -o $O/k1_r2l \
   python bench.py --steps 1 --warmup 3 --skip-ssz --skip-strong --skip-single --skip-rlc --skip-block > /dev/null 2> $O/k1_ncu.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_bench_T4096.csv \
   python bench.py --steps 2 --warmup 3 --skip-strong --skip-single --skip-rlc --skip-block > $O/bench_under_ncu.json 2> $O/bench_under_ncu.err
ls -la $O
