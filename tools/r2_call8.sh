#!/bin/bash
# Round-2 GPU call 8 (1 GPU): reworked RLC flow (signature pair inside the T+1 Miller launch, split scaling roles).
set -u
O=gpurun_out/r2c8; mkdir -p $O
timeout 900 python -m pytest tests/test_rlc_gpu.py tests/test_bls_gpu.py -x -q --durations=5 > $O/pytest.log 2>&1; echo "rc=$?" >> $O/pytest.log
B200_BLS_TRACE=1 timeout 900 python bench.py --steps 5 --warmup 3 --skip-ssz --skip-strong > $O/bench.json 2> $O/bench.err; echo "bench rc=$?" >> $O/bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_rlc.csv \
   python bench.py --steps 1 --warmup 3 --skip-ssz --skip-strong --skip-single > /dev/null 2> $O/ncu.err
ls -la $O
# VM register-file slot stride A/B (25 scalar vs 24 / 28 with 128-bit accesses)
for lib in ethereum_consensus_b200/libb200_consensus.so ethereum_consensus_b200/libb200_consensus_vm24.so ethereum_consensus_b200/libb200_consensus_vm28.so; do
  echo "== $lib"
  B200_LIB=$PWD/$lib B200_BLS_TRACE=1 timeout 600 python bench.py --steps 3 --warmup 3 --skip-ssz --skip-strong --skip-single --skip-rlc 2> /tmp/ab.err \
    | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('tuples/s', round(d['value']), 'ms/step', round(d['ms_per_step'],2), 'registry ms', round(d['registry_mode']['ms_per_step'],2))"
  grep "b200 bls" /tmp/ab.err | sed -n "5,5p"
  B200_LIB=$PWD/$lib timeout 600 python -m pytest tests/test_bls_gpu.py -x -q 2>&1 | tail -1
done > gpurun_out/r2c8/vm_slots.txt 2>&1
