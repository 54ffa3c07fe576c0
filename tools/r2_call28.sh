#!/bin/bash
# Round-2 GPU call 28 (1 GPU): compute-sanitizer initcheck (uninitialised global reads) and racecheck (shared-memory hazards: the
# pairing VM's register files, the aggregate kernel's trees, the cooperative Merkle levels) on the small parity tests.
set -u
O=gpurun_out/r2c28; mkdir -p $O
timeout 1200 compute-sanitizer --tool initcheck --print-limit 8 --error-exitcode 9 python -m pytest -x -q \
   "tests/test_bls_gpu.py::test_fast_aggregate_verify_batch_all_golden" "tests/test_bls_gpu.py::test_mixed_mode_extra_keys_validated_in_call" \
   "tests/test_rlc_gpu.py::test_rlc_on_golden_cases" tests/test_ssz_gpu.py > $O/initcheck.log 2>&1; echo "initcheck rc=$?" >> $O/initcheck.log
timeout 1500 compute-sanitizer --tool racecheck --print-limit 8 --error-exitcode 9 python -m pytest -x -q \
   "tests/test_bls_gpu.py::test_fast_aggregate_verify_batch_all_golden" "tests/test_rlc_gpu.py::test_rlc_on_golden_cases" \
   "tests/test_ssz_gpu.py" -k "not large" > $O/racecheck.log 2>&1; echo "racecheck rc=$?" >> $O/racecheck.log
ls -la $O; tail -5 $O/initcheck.log; tail -5 $O/racecheck.log
