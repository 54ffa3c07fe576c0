#!/bin/bash
# Round-2 GPU call 1: parity of the shipped build, A/B of the prepared variants, the fixed IMAD.WIDE microbenchmark,
# launch list + one full ncu capture of the per-key kernel of the build that ships.  Everything lands in gpurun_out/.
set -u
O=gpurun_out/r2c1; mkdir -p $O
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv -lms 500 > $O/clocks.csv &
SMI=$!
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
python tools/int_peaks.py > $O/int_peaks.json 2> $O/int_peaks.err
for lib in ethereum_consensus_b200/libb200_consensus*.so; do
  for var in 7 0; do
    echo "== $lib G1_VARIANT=$var"
    B200_G1_VARIANT=$var B200_LIB=$PWD/$lib B200_BLS_TRACE=1 timeout 600 python bench.py --steps 3 --warmup 3 --skip-ssz 2> /tmp/ab.err \
      | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('tuples/s', round(d['value']), 'ms/step', round(d['ms_per_step'],2), 'K1', round(d['roofline']['kernel_ms'],2), 'registry', round(d['registry_mode']['tuples_per_s']), 'reg ms', round(d['registry_mode']['ms_per_step'],2))"
    grep "b200 bls" /tmp/ab.err | sed -n "5,5p;12,12p"
  done
  B200_LIB=$PWD/$lib timeout 600 python -m pytest tests/test_bls_gpu.py -x -q 2>&1 | tail -1
done > $O/ab.txt 2>&1
# launch list of the default bench command (warm-up launches skipped by count is fragile: list everything of a short run)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_bench_T4096.csv \
   python bench.py --steps 2 --warmup 3 > $O/bench_under_ncu.json 2> $O/bench_under_ncu.err
# one full capture of the per-key kernel (3rd launch = past warm-up)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_g1_validate -s 2 -c 1 -o $O/k1_o1 \
   python bench.py --steps 1 --warmup 3 --skip-ssz > /dev/null 2> $O/k1_ncu.err
# the plain bench line (not under a profiler)
timeout 900 python bench.py --steps 5 --warmup 3 > $O/bench.json 2> $O/bench.err
kill $SMI
ls -la $O
