#!/bin/bash
# Round-2 GPU call 31 (1 GPU): collector packs validator indices at collection time (block.py) — spec-shape block test + the block leg of the bench.
set -u
O=gpurun_out/r2c31; mkdir -p $O
timeout 600 python -m pytest tests/test_config_scale_gpu.py -k "configs3 or get_domain" -x -q > $O/pytest_block.log 2>&1; echo "rc=$?" >> $O/pytest_block.log
timeout 600 python bench.py --steps 2 --warmup 3 --tuples 256 --skip-ssz --skip-strong --skip-single --skip-rlc 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['block_signature_set'])" > $O/block.txt 2>&1
tail -2 $O/pytest_block.log; cat $O/block.txt
