#!/bin/bash
# Round-2 GPU call 26 (1 GPU): final state — full GPU suite, smoke(), the default bench line, launch list of the bench command, SSZ phase trace.
set -u
O=gpurun_out/r2c26; mkdir -p $O
timeout 1800 python -m pytest tests -m gpu -x -q --durations=5 > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/smoke.log 2>&1; echo "smoke rc=$?" >> $O/smoke.log
B200_BLS_TRACE=1 timeout 900 python bench.py > $O/bench_default.json 2> $O/bench_default.err; echo "bench rc=$?" >> $O/bench_default.err
B200_SSZ_TRACE=1 timeout 600 python bench.py --steps 2 --warmup 3 --tuples 256 --skip-strong --skip-single --skip-rlc --skip-block > /dev/null 2> $O/ssz_trace.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_bench_T4096.csv \
   python bench.py --steps 2 --warmup 3 --skip-strong --skip-single --skip-rlc --skip-block > $O/bench_under_ncu.json 2> $O/bench_under_ncu.err
ls -la $O
