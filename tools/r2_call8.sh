#!/bin/bash
# Round-2 GPU call 8 (1 GPU): reworked RLC flow (signature pair inside the T+1 Miller launch, split scaling roles).
set -u
O=gpurun_out/r2c8; mkdir -p $O
timeout 900 python -m pytest tests/test_rlc_gpu.py tests/test_bls_gpu.py -x -q --durations=5 > $O/pytest.log 2>&1; echo "rc=$?" >> $O/pytest.log
B200_BLS_TRACE=1 timeout 900 python bench.py --steps 5 --warmup 3 --skip-ssz --skip-strong > $O/bench.json 2> $O/bench.err; echo "bench rc=$?" >> $O/bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_rlc.csv \
   python bench.py --steps 1 --warmup 3 --skip-ssz --skip-strong --skip-single > /dev/null 2> $O/ncu.err
ls -la $O
