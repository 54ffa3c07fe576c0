#!/bin/bash
# Round-2 GPU call 20 (1 GPU): folded upper Merkle levels (k_merkle_coop) — SSZ parity suite, memcheck, A/B against
# B200_SSZ_FOLD=0; memcheck over the round's late BLS additions (mixed registry mode, split key copy, run-time VM schedules).
set -u
O=gpurun_out/r2c20; mkdir -p $O
timeout 900 python -m pytest tests/test_ssz_gpu.py tests/test_shuffle_gpu.py "tests/test_config_scale_gpu.py::test_configs2_full_state_root" -x -q > $O/pytest_ssz.log 2>&1; echo "pytest rc=$?" >> $O/pytest_ssz.log
timeout 1500 compute-sanitizer --tool memcheck --print-limit 5 --error-exitcode 9 python -m pytest -x -q tests/test_ssz_gpu.py \
   "tests/test_bls_gpu.py::test_mixed_mode_extra_keys_validated_in_call" "tests/test_bls_gpu.py::test_chunked_pipeline_same_codes_on_ragged_golden_batch" \
   tests/test_vm_blob.py > $O/memcheck.log 2>&1; echo "memcheck rc=$?" >> $O/memcheck.log
for f in 1 0; do
  echo "== B200_SSZ_FOLD=$f"
  B200_SSZ_FOLD=$f timeout 600 python bench.py --steps 2 --warmup 3 --tuples 256 --skip-strong --skip-single --skip-rlc --skip-block 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1]); s = d['ssz']
print('resident', round(s['value_ms_device_resident'], 4), 'e2e', round(s['e2e_ms_from_pinned_host'], 4), 'incremental', s['incremental']['ms_device_root_only'], 'alu frac', round(s['roofline']['frac'], 3), 'launches', d['gpu_launches'])"
done > $O/fold_ab.txt 2>&1
ls -la $O
