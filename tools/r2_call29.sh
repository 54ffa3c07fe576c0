#!/bin/bash
# Round-2 GPU call 29 (1 GPU): racecheck again on the dirty-path tests after the barrier fix in k_validator_roots_sparse; SSZ suite.
set -u
O=gpurun_out/r2c29; mkdir -p $O
timeout 600 python -m pytest tests/test_ssz_gpu.py -x -q > $O/pytest_ssz.log 2>&1; echo "pytest rc=$?" >> $O/pytest_ssz.log
timeout 1200 compute-sanitizer --tool racecheck --print-limit 8 --error-exitcode 9 python -m pytest -x -q tests/test_ssz_gpu.py -k "incremental or resident or update or dirty" > $O/racecheck.log 2>&1; echo "racecheck rc=$?" >> $O/racecheck.log
timeout 300 python bench.py --steps 2 --warmup 3 --tuples 256 --skip-strong --skip-single --skip-rlc --skip-block 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1]); s = d['ssz']
print('resident', round(s['value_ms_device_resident'], 4), 'incremental', s['incremental'])" > $O/ssz_bench.txt 2>&1
tail -3 $O/pytest_ssz.log; tail -4 $O/racecheck.log; cat $O/ssz_bench.txt
