#!/bin/bash
# Round-2 GPU call 15 (1 GPU): chunked strict pipeline (per-key kernel of key range c+1 over range c's pairing chain) —
# parity on the ragged golden batch + config scale, A/B of the knobs; SSZ resident re-hash with side-stream stage chains;
# bench with the configs[3] block-signature-set leg.
set -u
O=gpurun_out/r2c15; mkdir -p $O
timeout 1500 python -m pytest -x -q tests/test_bls_gpu.py tests/test_ssz_gpu.py tests/test_config_scale_gpu.py tests/test_rlc_gpu.py --durations=5 > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
timeout 900 python tools/probe_chunks.py 4096 512 > $O/chunks_T4096.txt 2>&1
timeout 900 python tools/probe_chunks.py 2048 512 > $O/chunks_T2048.txt 2>&1
timeout 900 python tools/probe_chunks.py 1024 512 > $O/chunks_T1024.txt 2>&1
B200_BLS_TRACE=1 timeout 900 python bench.py --steps 5 --warmup 3 > $O/bench.json 2> $O/bench.err; echo "bench rc=$?" >> $O/bench.err
ls -la $O
