#!/bin/bash
# Round-2 GPU call 5 (1 GPU): call-based K1 at 12 and 16 warps, the reworked incremental root, full GPU suite, bench + ncu.
set -u
O=gpurun_out/r2c5; mkdir -p $O
for v in 7 6 0; do B200_G1_VARIANT=$v timeout 300 python tools/tune_k1.py; done > $O/k1_variants.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q --durations=8 > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
timeout 600 python tools/probe_incremental.py > $O/incremental_trace.txt 2>&1
timeout 900 python bench.py --steps 5 --warmup 3 > $O/bench.json 2> $O/bench.err; echo "bench rc=$?" >> $O/bench.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_g1_validate -s 2 -c 1 -o $O/k1_call_r168 \
   python bench.py --steps 1 --warmup 3 --skip-ssz --skip-strong --skip-single > /dev/null 2> $O/k1_ncu.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_bench_T4096.csv \
   python bench.py --steps 2 --warmup 3 --skip-strong --skip-single > $O/bench_under_ncu.json 2> $O/bench_under_ncu.err
ls -la $O
