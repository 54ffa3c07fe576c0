#!/bin/bash
# Round-2 GPU call 11 (1 GPU): VM instruction prefetch at T = 4096 / 256; full GPU suite; final bench + launch list + ncu.
set -u
O=gpurun_out/r2c11; mkdir -p $O
for t in 4096 256; do echo "== T=$t"; B200_BLS_TRACE=1 timeout 600 python bench.py --steps 3 --warmup 3 --skip-ssz --skip-strong --skip-single --skip-rlc --tuples $t 2>&1 >/dev/null | grep "b200 bls" | sed -n "5,5p"; done > $O/vm_prefetch.txt 2>&1
timeout 1800 python -m pytest tests -m gpu -x -q --durations=6 > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
B200_BLS_TRACE=1 timeout 900 python bench.py --steps 10 --warmup 5 > $O/bench.json 2> $O/bench.err; echo "bench rc=$?" >> $O/bench.err
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_reference.json 2> $O/bench_reference.err; echo "rc=$?" >> $O/bench_reference.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_bench_T4096.csv \
   python bench.py --steps 2 --warmup 3 --skip-strong --skip-single --skip-rlc > $O/bench_under_ncu.json 2> $O/bench_under_ncu.err
ls -la $O
