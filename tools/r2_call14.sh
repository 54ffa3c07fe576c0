#!/bin/bash
# Round-2 GPU call 14 (1 GPU): compute-sanitizer memcheck over the round's new kernels (RLC, shuffling, fused sparse Merkle
# levels, two-pass finisher, 16-lane VM teams) on small inputs; final bench + ncu capture of the shipping per-key kernel.
set -u
O=gpurun_out/r2c14; mkdir -p $O
timeout 1500 compute-sanitizer --tool memcheck --print-limit 5 --error-exitcode 9 python -m pytest -x -q \
   "tests/test_rlc_gpu.py::test_rlc_on_golden_cases" "tests/test_rlc_gpu.py::test_rlc_swapped_signatures_do_not_cancel" \
   "tests/test_shuffle_gpu.py::test_shuffled_indices_small_sizes" "tests/test_shuffle_gpu.py::test_active_indices_and_state_resident_shuffle" \
   tests/test_ssz_gpu.py "tests/test_bls_gpu.py::test_fast_aggregate_verify_batch_all_golden" "tests/test_bls_gpu.py::test_registry_mode_matches_strict" \
   > $O/memcheck.log 2>&1; echo "memcheck rc=$?" >> $O/memcheck.log
B200_BLS_TRACE=1 timeout 900 python bench.py --steps 10 --warmup 5 > $O/bench.json 2> $O/bench.err; echo "bench rc=$?" >> $O/bench.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_g1_validate -s 2 -c 1 -o $O/k1_final \
   python bench.py --steps 1 --warmup 3 --skip-ssz --skip-strong --skip-single --skip-rlc > /dev/null 2> $O/k1_ncu.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_bench_T4096.csv \
   python bench.py --steps 2 --warmup 3 --skip-strong --skip-single --skip-rlc > $O/bench_under_ncu.json 2> $O/bench_under_ncu.err
ls -la $O
