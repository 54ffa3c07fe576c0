#!/bin/bash
# Round-2 GPU call 9 (1 GPU): VM register-file slot stride A/B on the shipping (unfused) programs; full GPU suite; smoke; bench.
set -u
O=gpurun_out/r2c9; mkdir -p $O
for lib in ethereum_consensus_b200/libb200_consensus.so ethereum_consensus_b200/libb200_consensus_vm24.so ethereum_consensus_b200/libb200_consensus_vm28.so; do
  echo "== $lib"
  B200_LIB=$PWD/$lib B200_BLS_TRACE=1 timeout 600 python bench.py --steps 3 --warmup 3 --skip-ssz --skip-strong --skip-single --skip-rlc 2> /tmp/ab.err \
    | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('tuples/s', round(d['value']), 'ms/step', round(d['ms_per_step'],2), 'registry ms', round(d['registry_mode']['ms_per_step'],2))"
  grep "b200 bls" /tmp/ab.err | sed -n "5,5p"
  B200_LIB=$PWD/$lib timeout 600 python -m pytest tests/test_bls_gpu.py -x -q 2>&1 | tail -1
done > $O/vm_slots.txt 2>&1
timeout 1800 python -m pytest tests -m gpu -x -q --durations=6 > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?" >> $O/smoke.log
B200_BLS_TRACE=1 timeout 900 python bench.py --steps 5 --warmup 3 > $O/bench.json 2> $O/bench.err; echo "bench rc=$?" >> $O/bench.err
ls -la $O
