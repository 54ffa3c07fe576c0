#!/bin/bash
# Round-2 GPU call 34 (1 GPU): the bench command once more after the last host-side edits (mulx/adx CPU arm), short form.
set -u
O=gpurun_out/r2c34; mkdir -p $O
timeout 400 python bench.py --steps 3 --warmup 3 --skip-strong --skip-rlc > $O/bench.json 2> $O/bench.err; echo "rc=$?" >> $O/bench.err
tail -1 $O/bench.err; python -c "
import json; d=json.loads(open('$O/bench.json').read().strip().splitlines()[-1]); print(d['value'], d['e2e']['value'], d['cpu_baseline'], d['block_signature_set']['ms_per_block_registry'], d['ssz']['value_ms_device_resident'])"
