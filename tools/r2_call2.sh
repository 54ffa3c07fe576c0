#!/bin/bash
# Round-2 GPU call 2 (1 GPU): the full GPU suite incl. the config-scale tests, the bench line of the new default build
# (variant 0 + PTX square, -O1 for g1/g2/vm), its launch list and one full ncu capture of the per-key kernel.
set -u
O=gpurun_out/r2c2; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q --durations=15 > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
timeout 900 python bench.py --steps 5 --warmup 3 > $O/bench.json 2> $O/bench.err; echo "bench rc=$?" >> $O/bench.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_bench_T4096.csv \
   python bench.py --steps 2 --warmup 3 --skip-strong --skip-single > $O/bench_under_ncu.json 2> $O/bench_under_ncu.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_g1_validate -s 2 -c 1 -o $O/k1_main \
   python bench.py --steps 1 --warmup 3 --skip-ssz --skip-strong --skip-single > /dev/null 2> $O/k1_ncu.err
python tools/int_peaks.py > $O/int_peaks.json 2> $O/int_peaks.err
ls -la $O
